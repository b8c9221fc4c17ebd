#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c9; mkdir -p $O
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
i=0
for v in "MVTB_IS_HS=1 MVTB_IS_STORE=2" "MVTB_IS_HS=1 MVTB_IS_STORE=2 MVTB_IS_DEBUG=1" "MVTB_IS_HS=1 MVTB_IS_STORE=2 MVTB_IS_DEBUG=3" "MVTB_IS_HS=1 MVTB_IS_STORE=2 MVTB_IS_DEBUG=7" "MVTB_IS_HS=1 MVTB_IS_STORE=0 MVTB_IS_DEBUG=7" "MVTB_IS_HS=2 MVTB_IS_STORE=2 MVTB_IS_DEBUG=7" "MVTB_IS_HS=2 MVTB_IS_STORE=2 MVTB_IS_DEBUG=3" "MVTB_IS_HS=1 MVTB_IS_STORE=2 MVTB_IS_DEBUG=4"; do
  i=$((i+1)); echo "$v" > $O/v$i.txt; env $v $B > $O/v$i.log 2>&1
done
