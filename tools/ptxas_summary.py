#!/usr/bin/env python
"""Registers / spills / shared memory of every kernel, from the ptxas -v logs the build leaves in csrc/build/, plus a short
SASS excerpt that shows which instructions the cross-CTA protocol of the resident kernel compiles to.
  python tools/ptxas_summary.py > profiles/r2_ptxas_summary.txt"""
import glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
logs = sorted(glob.glob(os.path.join(ROOT, "arap_flow_b200", "csrc", "build", "*.ptxas.log")))
print("# ptxas -v summary (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false), one line per kernel")
for log in logs:
    txt = open(log).read().splitlines()
    name = None
    spill = ""
    for l in txt:
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", l)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"arapb200::\(anonymous namespace\)::", "", name)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", l)
        if m:
            spill = f"stack {m.group(1)} B, spill st {m.group(2)} B / ld {m.group(3)} B"
        m = re.search(r"Used (\d+) registers.*?(\d+) bytes smem", l)
        if m and name:
            print(f"{os.path.basename(log)[:-10]:18s} {m.group(1):>4s} regs  {spill:44s} static smem {m.group(2):>5s} B  {name[:110]}")
            name = None
so = os.path.join(ROOT, "arap_flow_b200", "libarapb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
print("\n# SASS mnemonics of the resident kernel's cross-CTA protocol and of the hot loops (count over libarapb200.so)")
for pat, what in [(r"STG\.E\.128\.STRONG\.GPU", "halo publication: st.relaxed.gpu.global.v2.b64 (two 64-bit (float, tag) elements)"),
                  (r"LDG\.E\.128\.STRONG\.GPU", "halo fetch: ld.relaxed.gpu.global.v2.b64"),
                  (r"LDG\.E\.64\.STRONG\.GPU", "barrier poll: ld.relaxed.gpu.global.u64"),
                  (r"REDG\.E\.ADD\.64\.STRONG\.GPU", "barrier arrival / wide accumulators: red.relaxed.gpu.global.add.u64"),
                  (r"REDUX\.SUM", "warp-level integer limb sums"),
                  (r"MEMBAR", "memory fences (none expected in the solver kernels)"),
                  (r"LDS\.128", "tile reads"), (r"UBLKCP|UTMALDG", "TMA bulk copies")]:
    n = len(re.findall(pat, sass))
    print(f"{n:6d}  {pat:32s} {what}")
