#!/bin/bash
# round 2, GPU call 2: does freshly written data survive in L2 for a sparse read-modify-write? + DRAM bytes of fused variants
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c2; mkdir -p $O
timeout 300 tools/_build/l2probe > $O/l2probe.txt 2>&1
timeout 120 tools/_build/l2probe ncu > $O/l2probe_plain.txt 2>&1 && \
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --csv --log-file $O/l2probe_ncu.csv tools/_build/l2probe ncu > $O/l2probe_ncu.log 2>&1
S="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
M="--cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum -k regex:k_bl_inv_sp -s 3 -c 2 --csv"
for v in "MVTB_IS_HS=1" "MVTB_IS_HS=1 MVTB_IS_STORE=2" "MVTB_IS_HS=1 MVTB_IS_LAG=0 MVTB_IS_SPREAD=5" "MVTB_IS_HS=4 MVTB_IS_LAG=0 MVTB_IS_SPREAD=5" "MVTB_IS_HS=4 MVTB_IS_STORE=2 MVTB_IS_LAG=0 MVTB_IS_SPREAD=5"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v $S > $O/plain_$tag.log 2>&1 && env $v ncu $M --log-file $O/ncu_$tag.csv $S > $O/ncu_$tag.log 2>&1
done
ls -la $O
