#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c18; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q > $O/pytest_tc.log 2>&1; echo "rc=$?" >> $O/pytest_tc.log
tail -5 $O/pytest_tc.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
MVTB_TC=1 $B > $O/b_tc.log 2>&1
MVTB_TC=1 MVTB_TC_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/b_prof.log 2> $O/b_prof.err
