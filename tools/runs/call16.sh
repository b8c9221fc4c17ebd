#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c16; mkdir -p $O
S="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
$S > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:k_ -s 9 -c 12 --csv --log-file $O/launches_traffic.csv $S > $O/ncu1.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_default.log 2> $O/bench_default.err
tail -c 600 $O/bench_default.log
