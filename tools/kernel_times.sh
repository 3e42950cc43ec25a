#!/bin/bash
# per-kernel durations (warm caches, no replay) of a small ncu_target.py run: tools/kernel_times.sh stream C4 8 1
set -e
timeout 200 python tools/ncu_target.py "$@"
timeout 300 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/ktimes.csv python tools/ncu_target.py "$@" > gpurun_out/ktimes.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/ktimes.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]; hdr=rows[hi]
ik,iv=hdr.index("Kernel Name"),hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)!=len(hdr): continue
    k=r[ik].split("(")[0][-30:]; agg.setdefault(k,[]).append(float(r[iv].replace(",","")))
for k,v in agg.items(): print(f"{k:32s} n={len(v):3d} mean={sum(v)/len(v)/1000:8.2f} us  min={min(v)/1000:8.2f}  max={max(v)/1000:8.2f}")
PY
