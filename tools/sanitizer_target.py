#!/usr/bin/env python
"""Tiny workloads for compute-sanitizer: both back-ends + warp on a 96x80 problem with a ragged mask."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arap_flow_b200 import lib, synth

sp = synth.synth(96, 80, 2, 2, 5)
mask = sp.masks[0].copy()
mask[::7, ::5] = 255
for backend in (lib.BACKEND_RESIDENT, lib.BACKEND_STREAM):
    flow, rgb, m, costs = lib.deform(sp.rgb, mask, sp.matches, nCont=2, nGN=1, nPCG=6, backend=backend)
    print("backend", backend, "cost", float(costs[-1, -1]))
b = lib.Batch(96, 80, 3, 1, 1, 5)
outs = [b.submit(i, sp.rgb, sp.masks[i % 2], sp.matches) for i in range(3)]
b.run()
print("batch ok", [float(o["costs"][-1, -1]) for o in outs])
