#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c4; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q > $O/pytest_tc.log 2>&1; echo "rc=$?" >> $O/pytest_tc.log
tail -30 $O/pytest_tc.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
MVTB_TWO_CALLS=1 $B > $O/b_tc_two.log 2>&1
$B > $O/b_tc_fused.log 2>&1
MVTB_NO_TC=1 MVTB_TWO_CALLS=1 $B > $O/b_notc_two.log 2>&1
tail -c 1500 $O/b_tc_two.log
