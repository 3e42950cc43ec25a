#!/usr/bin/env python
"""Small fixed workloads for ncu captures (reduced schedules so that ~40 replays stay short).
  ncu_target.py resident C1 50 3   -> 3 co-resident 854x480 problems, 1 x 1 x 50 PCG iterations
  ncu_target.py stream  C4 6 1     -> 1920x1080 through the streaming back-end
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arap_flow_b200 import lib, synth

backend = {"resident": lib.BACKEND_RESIDENT, "stream": lib.BACKEND_STREAM}[sys.argv[1] if len(sys.argv) > 1 else "resident"]
cfg = sys.argv[2] if len(sys.argv) > 2 else "C1"
nPCG = int(sys.argv[3]) if len(sys.argv) > 3 else 50
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
pairs = [synth.config(cfg, i) for i in range(B)]
W, H = pairs[0].W, pairs[0].H
b = lib.Batch(W, H, B, 1, 1, nPCG, backend)
for _ in range(2):
    outs = [b.submit(i, p.rgb, p.masks[0], p.matches) for i, p in enumerate(pairs)]
    b.run()
print("ok", float(outs[0]["costs"][-1, -1]), b.timing_ms(), "launches", b.launches())
