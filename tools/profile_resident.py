#!/usr/bin/env python
"""Print the resident kernel's own cycle accounting for one synthetic problem (GPU box only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arap_flow_b200 import lib, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
nPCG = int(sys.argv[2]) if len(sys.argv) > 2 else 400
copies = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # > 1: the first of that many co-resident identical problems
sp = synth.config(cfg)
for rep in range(2):
    prof, info, ms = lib.debug_resident_profile(sp.masks[0], sp.matches, 1, 2, nPCG, copies=copies)
it = 2 * nPCG
names = ["phase1(JTJ)", "phase2(update)", "phase3(p,halo)", "other", "bar:arrive", "bar:poll", "bar:fold", None,
         " arrive:limbs", " arrive:redux+stage", " arrive:cta-sync", " arrive:sum+red", " fold:decode"]
print(f"{cfg} x{copies}: {info}, launch {ms:.3f} ms, {ms * 1e3 / it:.2f} us per PCG iteration, barriers/CTA {int(prof[0, 7])}")
clk = 1.9e3  # cycles per us (approx.)
for i, n in enumerate(names):
    if n is None:
        continue
    v = prof[:, i].astype(np.float64) / it
    if i > 7 and not v.any():
        continue   # the finer split is only filled by builds that carry the extra time stamps
    print(f"  {n:16s} cycles/iter: mean {v.mean():9.0f}  min {v.min():9.0f}  max {v.max():9.0f}   (~{v.mean() / clk:.2f} us)")
