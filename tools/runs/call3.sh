#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c3; mkdir -p $O
timeout 120 tools/_build/tc_probe > $O/tc_probe.txt 2>&1; echo "rc=$?" >> $O/tc_probe.txt
cat $O/tc_probe.txt
