#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c56; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bandlimited.py -x -q > $O/t.log 2>&1; tail -3 $O/t.log
for r in 20 25 30 32; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg1 --disk-r $r --no-cpu-baseline --no-e2e --no-parity > $O/r_$r.log 2> $O/r_$r.err
done
