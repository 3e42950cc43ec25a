#!/usr/bin/env python
"""profiles/r1_traffic.json from the two ncu csv captures of tools/refresh_profiles.sh (gpurun_out/traffic_resident{,3}.csv)."""
import csv, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALG_PER_PROBLEM = 135213 * (156.0 * 60800 + 132.0 * 152)   # C1: active px x algorithmic bytes per pixel per solve


def parse(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    vals = {r[hdr.index("Metric Name")]: float(r[hdr.index("Metric Value")].replace(",", "")) for r in rows[1:]}
    return vals, rows[1][hdr.index("Kernel Name")].split("::")[-1], rows[1][hdr.index("Grid Size")], rows[1][hdr.index("Block Size")]


out = {"what": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_resident "
               "--launch-skip 1 --launch-count 1 on: python bench.py --steps 1 --warmup 1 --batch {3,4} --no-cpu-baseline; one full "
               "19x8x400 launch of N co-resident C1 problems", "launches": {}}
for n, name in ((4, "traffic_resident.csv"), (3, "traffic_resident3.csv")):
    path = os.path.join(ROOT, "gpurun_out", name)
    if not os.path.exists(path):
        continue
    v, k, g, b = parse(path)
    out["launches"][str(n)] = {"kernel": f"{k}, grid {g}, block {b}", "problems_per_launch": n,
                               "dram_bytes_read": int(v["dram__bytes_read.sum"]), "dram_bytes_write": int(v["dram__bytes_write.sum"]),
                               "dram_bytes_per_launch": int(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]),
                               "duration_ms_under_ncu": v["gpu__time_duration.sum"] / 1e6,
                               "algorithmic_bytes_per_launch": ALG_PER_PROBLEM * n}
    shutil.copy(path, os.path.join(ROOT, "profiles", "r1_traffic_resident_ncu.csv" if n == 4 else "r1_traffic_resident3_ncu.csv"))
json.dump(out, open(os.path.join(ROOT, "profiles", "r1_traffic.json"), "w"), indent=1)
print(json.dumps(out["launches"], indent=1))
