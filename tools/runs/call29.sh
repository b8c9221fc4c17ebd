#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c29; mkdir -p $O
S="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
$S > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:k_ -s 9 -c 12 --csv --log-file $O/launches_traffic.csv $S > $O/ncu1.log 2>&1
$S > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_bl_fwd_tc -s 3 -c 1 -o $O/fwd_tc_full $S > $O/ncu2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_default.log 2> $O/bench_default.err
timeout 900 python bench.py --steps 5 --warmup 3 --workload cfg3s --no-cpu-baseline > $O/bench_cfg3s.log 2> $O/bench_cfg3s.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg4 > $O/bench_cfg4.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg5 > $O/bench_cfg5.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --workload cfg1 --no-cpu-baseline > $O/bench_cfg1.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1
ls -la $O
