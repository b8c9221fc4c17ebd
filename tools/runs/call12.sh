#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c12; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_chain_sp.py tests/test_gpu_sparse_sp.py -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
MVTB_TWO_CALLS=1 $B > $O/b_two.log 2>&1
$B > $O/b_fused.log 2>&1
$B --workload cfg3 > $O/b_cfg3_fused.log 2>&1
MVTB_TWO_CALLS=1 $B --workload cfg3 > $O/b_cfg3_two.log 2>&1
