#!/usr/bin/env python
"""Summarise one `ncu --set full` raw-page CSV of the headline kernel into profiles/r01_ncu_c3_traffic.json (the file
bench.py reads `roofline.traffic` from).  Usage: python tools/ncu_traffic_json.py RAW.csv "workload string" OUT.json"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "Ghz": 1.0, "Mhz": 1e-3, "%": 1.0}


def main():
    raw, workload, out = sys.argv[1:4]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    vals = rows[2 if len(sys.argv) < 5 else 2 + int(sys.argv[4])]
    col = {h: i for i, h in enumerate(hdr)}

    def get(name):
        i = col[name]
        return float(vals[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    j = {"kernel": vals[col["Kernel Name"]], "workload": workload, "dram_bytes_read": rd, "dram_bytes_write": wr,
         "traffic_bytes_per_launch": rd + wr,
         "tensor_pipe_active_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
         "l2_hit_rate_pct": get("lts__t_sector_hit_rate.pct"),
         "duration_ms_under_ncu": get("gpu__time_duration.sum"),
         "sm_clock_ghz_under_ncu": get("sm__cycles_elapsed.avg.per_second"),
         "source": "ncu --set full --clock-control none -k regex:score_select_tc -s 3 -c 1 python bench.py --steps 1 "
                   "--warmup 3 --no-cpu-baseline (gpurun, 1 B200); raw page: " + raw}
    json.dump(j, open(out, "w"), indent=1)
    print(json.dumps(j))


if __name__ == "__main__":
    main()
