#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c11; mkdir -p $O
timeout 120 tools/_build/tc_rate2 > $O/tc_rate2.txt 2>&1; echo "rc=$?" >> $O/tc_rate2.txt
cat $O/tc_rate2.txt
