#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c39; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
run() { name=$1; shift; env "$@" $B > $O/b_$name.log 2> $O/b_$name.err; }
run base X=1
run hs1 MVTB_IS_HS=1
run hs3 MVTB_IS_HS=3
run hs4 MVTB_IS_HS=4
run lag150 MVTB_IS_LAG=150
run lag600 MVTB_IS_LAG=600
run lag1200 MVTB_IS_LAG=1200
run store1 MVTB_IS_STORE=1
run store0 MVTB_IS_STORE=0
run spread50 MVTB_IS_SPREAD=50
run spread200 MVTB_IS_SPREAD=200
