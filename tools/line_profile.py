#!/usr/bin/env python
"""Per-source-line executed-instruction profile of one kernel: ncu SourceCounters (SASS page, csv) joined with
`nvdisasm -g` line info of the cubin extracted from libarapb200.so.
  cuobjdump -xelf solver_resident.sm_100a.cubin arap_flow_b200/libarapb200.so
  ncu -i rep.ncu-rep --page source --print-source sass --csv > sass.csv
  tools/line_profile.py solver_resident.sm_100a.cubin k_resident_tILi160ELi3ELb0 sass.csv solver_resident.cu <warp-iterations>"""
import collections, csv, re, subprocess, sys

cubin, kern, sass_csv, srcname, denom = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(out) if l.startswith(".text.") and kern in l][0]
end = next(i for i in range(start + 1, len(out)) if out[i].startswith("//---------------------"))
chain, ins = [], []
for l in out[start:end]:
    if "//## File" in l:
        chain = [(f, int(n)) for f, n in re.findall(r'File ".*?([^/"]+)", line (\d+)', l)]
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append(chain)
rows = list(csv.reader(open(sass_csv)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
hdr = rows[starts[0] + 1]
data = rows[starts[0] + 2:(starts[1] if len(starts) > 1 else len(rows))]
ie = hdr.index("Instructions Executed")
assert len(ins) == len(data), (len(ins), len(data))
agg = collections.Counter()
for ch, r in zip(ins, data):
    key = next((ln for f, ln in reversed(ch) if f == srcname), ch[-1] if ch else None)
    agg[key] += int(r[ie])
tot = sum(agg.values())
src = open(sys.argv[6] if len(sys.argv) > 6 else "arap_flow_b200/csrc/" + srcname).read().splitlines()
print(f"total {tot} warp-instructions = {tot / denom:.0f} per warp-iteration")
for key, v in agg.most_common(40):
    txt = src[key - 1].strip()[:100] if isinstance(key, int) else str(key)
    print(f"{v / tot:6.3f} {v / denom:7.1f}  L{key}: {txt}")
