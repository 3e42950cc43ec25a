#!/usr/bin/env python
"""Small fixed workload for ncu captures: C1 (854x480) through the chosen back-end at a reduced schedule."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arap_flow_b200 import lib, synth

backend = {"resident": lib.BACKEND_RESIDENT, "stream": lib.BACKEND_STREAM}[sys.argv[1] if len(sys.argv) > 1 else "resident"]
cfg = sys.argv[2] if len(sys.argv) > 2 else "C1"
nPCG = int(sys.argv[3]) if len(sys.argv) > 3 else 50
sp = synth.config(cfg)
for _ in range(2):
    flow, rgb, m, costs = lib.deform(sp.rgb, sp.masks[0], sp.matches, nCont=1, nGN=1, nPCG=nPCG, backend=backend)
print("ok", float(costs[-1, -1]))
