#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c55; mkdir -p $O
for r in 9 12.5 16 20 25 30 32 35; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg1 --disk-r $r --no-cpu-baseline --no-e2e --no-parity > $O/r_$r.log 2> $O/r_$r.err
  MVTB_TC=0 timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg1 --disk-r $r --no-cpu-baseline --no-e2e --no-parity > $O/cc_$r.log 2> $O/cc_$r.err
done
