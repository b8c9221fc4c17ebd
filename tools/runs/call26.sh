#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c26; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "select_warps" > $O/pytest_tc.log 2>&1; echo "rc=$?" >> $O/pytest_tc.log
tail -25 $O/pytest_tc.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
MVTB_TC_INV=2 $B > $O/b_fsel.log 2> $O/b_fsel.err; tail -2 $O/b_fsel.err
