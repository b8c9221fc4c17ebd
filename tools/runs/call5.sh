#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c5; mkdir -p $O
timeout 120 tools/_build/tc_rate > $O/tc_rate.txt 2>&1; echo "rc=$?" >> $O/tc_rate.txt
cat $O/tc_rate.txt
