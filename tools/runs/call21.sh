#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c21; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q > $O/pytest_tc.log 2>&1; echo "rc=$?" >> $O/pytest_tc.log
tail -5 $O/pytest_tc.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
for f in 0 1 2 4 6; do MVTB_TC=1 MVTB_TCI_FLAGS=$f $B > $O/b_f$f.log 2> $O/b_f$f.err; done
MVTB_TC=1 MVTB_TWO_CALLS=1 $B > $O/b_two.log 2> $O/b_two.err
