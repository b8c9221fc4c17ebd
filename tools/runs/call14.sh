#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c14; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -3 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_default.log 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
tail -c 3000 $O/bench_default.log
