#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c54; mkdir -p $O
S="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
$S > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:k_ -s 9 -c 12 --csv --log-file $O/launches_traffic.csv $S > $O/ncu1.log 2>&1
$S > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_bl_inv_sp -s 3 -c 1 -o $O/inv_sp_full $S > $O/ncu2.log 2>&1
ls -la $O
