#!/usr/bin/env python
"""Repeat batched solves many times and check every result against the first (protocol stress test)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arap_flow_b200 import lib, synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cases = [("C1", 4, dict(nCont=1, nGN=2, nPCG=200)), ("C3", 3, dict(nCont=1, nGN=2, nPCG=200)), ("C0", 8, dict(nCont=2, nGN=2, nPCG=100))]
for cfg, B, kw in cases:
    pairs = [synth.config(cfg, i) for i in range(B)]
    W, H = pairs[0].W, pairs[0].H
    b = lib.Batch(W, H, B, kw["nCont"], kw["nGN"], kw["nPCG"], lib.BACKEND_RESIDENT)
    ref = None
    t0 = time.time()
    for r in range(reps):
        outs = [b.submit(i, p.rgb, p.masks[0], p.matches) for i, p in enumerate(pairs)]
        b.run()
        cur = [(o["flow"].copy(), o["costs"].copy()) for o in outs]
        if ref is None:
            ref = cur
        else:
            for (f0, c0), (f1, c1) in zip(ref, cur):
                assert np.array_equal(f0, f1) and np.array_equal(c0, c1), f"{cfg}: run {r} differs from run 0"
    print(f"{cfg} x{B}: {reps} identical repetitions in {time.time() - t0:.1f} s")
    b.close()
print("stress ok")
