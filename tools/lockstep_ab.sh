#!/usr/bin/env bash
# A/B of the TMA producers' lock-step window (RDB_TC_LOCKSTEP = window in groups of 8 tiles, 0 = off):
# un-profiled bench line first, then one ncu pass of the scorer for its DRAM bytes / L2 hit rate.
set -u
export RDB_BENCH_N=${RDB_BENCH_N:-4000000}
for w in ${WINDOWS:-0 8}; do
  export RDB_TC_LOCKSTEP=$w
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ls_w${w}.json 2>/dev/null
  python - <<PY
import json
j = json.load(open("gpurun_out/ls_w${w}.json"))
print("window $w:", round(j["roofline"]["achieved"], 1), "TFLOP/s", j["clocks"]["sm_mhz"], "MHz recall", j.get("recall_at_10_vs_fp32_bruteforce_256q"))
PY
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:score_select_tc -s 3 -c 1 --csv --log-file gpurun_out/ls_w${w}_ncu.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time|lts__t" gpurun_out/ls_w${w}_ncu.csv | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $NF}'
done
