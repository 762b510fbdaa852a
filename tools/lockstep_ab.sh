#!/usr/bin/env bash
# A/B of the TMA producers' lock-step window at the full C3 size: interleaved un-profiled kernel times first
# (tools/ab_knobs.py), then one ncu pass per variant for the scorer's DRAM bytes / L2 hit rate.
# VARIANTS: space-separated RDB_* assignments (comma-joined inside one variant); "" = defaults.
set -u
N=${RDB_BENCH_N:-10000000}
VARIANTS=${VARIANTS:-"tc_lockstep=0 tc_lockstep_spins=256 tc_lockstep_spins=4096 tc_lockstep_spins=65536"}
AB_ROUNDS=${AB_ROUNDS:-3} python tools/ab_knobs.py $N 768 65536 10 bf16 IP $VARIANTS
for v in $VARIANTS; do
  tag=$(echo "$v" | tr '=,' '__')
  env RDB_BENCH_N=$N $(echo "$v" | tr ',' ' ') timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:score_select_tc -s 3 -c 1 --csv --log-file gpurun_out/ls_${tag}_ncu.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  echo "ncu $v:"
  grep -E "dram__bytes|gpu__time|lts__t" gpurun_out/ls_${tag}_ncu.csv | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $NF}'
done
