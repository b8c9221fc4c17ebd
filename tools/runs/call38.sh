#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c38; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
$B > $O/b_default.log 2> $O/b_default.err
MVTB_IS_DEBUG=8 $B > $O/b_staged.log 2> $O/b_staged.err
