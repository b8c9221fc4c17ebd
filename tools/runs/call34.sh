#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c34; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
MVTB_TC_INV=1 MVTB_BITS_OVERLAP=0 $B > $O/b_seq.log 2> $O/b_seq.err
MVTB_TC_INV=1 $B > $O/b_ovl.log 2> $O/b_ovl.err
MVTB_TC_INV=1 MVTB_BITS_OVERLAP=0 MVTB_TCI_PV=2 $B > $O/b_seq_pv2.log 2> $O/b_seq_pv2.err
MVTB_TC_INV=1 MVTB_BITS_OVERLAP=0 MVTB_TCI_PV=8 $B > $O/b_seq_pv8.log 2> $O/b_seq_pv8.err
