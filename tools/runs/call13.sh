#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c13; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_prologue_dice.py -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -30 $O/pytest.log
