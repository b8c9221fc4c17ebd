#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c32; mkdir -p $O
timeout 900 python bench.py --steps 5 --warmup 3 --workload cfg3s --no-cpu-baseline > $O/bench_cfg3s.log 2> $O/bench_cfg3s.err
MVTB_TC_INV=1 timeout 900 python bench.py --steps 5 --warmup 3 --workload cfg3s --no-cpu-baseline --no-e2e > $O/bench_cfg3s_tci.log 2> $O/bench_cfg3s_tci.err
timeout 300 python tools/small_calls.py > $O/small.txt 2>&1; head -3 $O/small.txt
tail -2 $O/*.err
