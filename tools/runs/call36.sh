#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c36; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg5 > $O/bench_cfg5.log 2>&1
timeout 300 python tools/prof_cfg5.py > $O/prof_cfg5.txt 2>&1; cat $O/prof_cfg5.txt
