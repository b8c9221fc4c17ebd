#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c10; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_chain_sp.py tests/test_gpu_sparse_sp.py -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
MVTB_TWO_CALLS=1 $B > $O/b_two.log 2>&1
i=0
for v in "MVTB_IS_HS=1 MVTB_IS_STORE=2" "MVTB_IS_HS=2 MVTB_IS_STORE=2" "MVTB_IS_HS=1 MVTB_IS_STORE=2 MVTB_IS_LAG=74 MVTB_IS_SPREAD=50" "MVTB_IS_HS=2 MVTB_IS_STORE=2 MVTB_IS_LAG=74 MVTB_IS_SPREAD=50" "MVTB_IS_HS=1 MVTB_IS_STORE=2 MVTB_IS_LAG=30 MVTB_IS_SPREAD=30" "MVTB_IS_HS=2 MVTB_IS_STORE=2 MVTB_IS_LAG=0 MVTB_IS_SPREAD=30" "MVTB_IS_HS=1 MVTB_IS_STORE=1" "MVTB_IS_HS=1 MVTB_IS_STORE=2 MVTB_IS_LAG=600"; do
  i=$((i+1)); echo "$v" > $O/v$i.txt; env $v $B > $O/v$i.log 2>&1
done
