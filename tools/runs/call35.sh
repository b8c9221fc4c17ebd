#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c35; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
MVTB_TC_INV=1 MVTB_BITS_OVERLAP=0 MVTB_TCI_DEBUG=4 $B > $O/b_2terms.log 2> $O/b_2terms.err
MVTB_TC_INV=1 MVTB_BITS_OVERLAP=0 MVTB_TC_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/b_prof.log 2> $O/b_prof.err
