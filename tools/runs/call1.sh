#!/bin/bash
# round 2, GPU call 1: fused inverse + select kernel — parity, first timings, queue / store-policy sweep, ncu
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 900 python -m pytest tests/test_gpu_chain_sp.py -x -q > $O/pytest_chain_sp.log 2>&1; echo "pytest rc=$?" >> $O/pytest_chain_sp.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
$B > $O/b_fused.log 2>$O/b_fused.err
MVTB_TWO_CALLS=1 $B > $O/b_two.log 2>&1
MVTB_IS_STORE=0 $B > $O/b_store0.log 2>&1
MVTB_IS_STORE=2 $B > $O/b_store2.log 2>&1
MVTB_IS_HS=2 $B > $O/b_hs2.log 2>&1
MVTB_IS_HS=1 $B > $O/b_hs1.log 2>&1
MVTB_IS_HS=6 $B > $O/b_hs6.log 2>&1
MVTB_IS_LAG=0 $B > $O/b_lag0.log 2>&1
MVTB_IS_LAG=150 $B > $O/b_lag150.log 2>&1
MVTB_IS_LAG=600 $B > $O/b_lag600.log 2>&1
MVTB_IS_SPREAD=30 $B > $O/b_spread30.log 2>&1
MVTB_IS_CHUNK=16 $B > $O/b_chunk16.log 2>&1
MVTB_IS_CHUNK=8 $B > $O/b_chunk8.log 2>&1
# ncu: DRAM traffic of the fused kernel with caches left alone between kernels (what the live pipeline sees)
S="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$S > $O/plain_ncu.log 2>&1 && \
ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_bl_inv_sp -s 3 -c 2 -o $O/is_full $S > $O/ncu_is.log 2>&1
ls -la $O
