#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c28; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
MVTB_TC_INV=2 MVTB_TCI_DEBUG=2 $B > $O/b_nofence.log 2> $O/b_nofence.err
MVTB_TC_INV=2 MVTB_TCI_DEBUG=3 $B > $O/b_nofence_nosel.log 2> $O/b_nofence_nosel.err
