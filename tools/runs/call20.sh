#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c20; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q > $O/pytest_tc.log 2>&1; echo "rc=$?" >> $O/pytest_tc.log
tail -30 $O/pytest_tc.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
MVTB_TC=1 $B > $O/b_tc.log 2> $O/b_tc.err
tail -3 $O/b_tc.err
