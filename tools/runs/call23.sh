#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c23; mkdir -p $O
B="timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
MVTB_TC=1 MVTB_TC_PROF=1 $B > $O/b_prof.log 2> $O/b_prof.err
