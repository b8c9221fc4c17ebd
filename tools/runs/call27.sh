#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c27; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
MVTB_TC_INV=2 MVTB_TCI_DEBUG=1 $B > $O/b_nosel.log 2> $O/b_nosel.err
MVTB_TC_INV=2 MVTB_TCI_PVF=2 $B > $O/b_pv2.log 2> $O/b_pv2.err
MVTB_TC_INV=2 MVTB_TCI_PVF=2 MVTB_TCI_DEBUG=1 $B > $O/b_pv2_nosel.log 2> $O/b_pv2_nosel.err
MVTB_TC_INV=2 MVTB_TCI_PVF=4 $B > $O/b_pv4.log 2> $O/b_pv4.err
MVTB_TC_INV=2 MVTB_TC_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/b_prof.log 2> $O/b_prof.err
