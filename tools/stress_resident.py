#!/usr/bin/env python
"""Repeat batched solves many times and check every result against the first (protocol stress test)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arap_flow_b200 import lib, synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cases = [("C1", 4, dict(nCont=1, nGN=2, nPCG=200)), ("C3", 3, dict(nCont=1, nGN=2, nPCG=200)), ("C0", 8, dict(nCont=2, nGN=2, nPCG=100))]
cases += [("C2", 2, dict(nCont=1, nGN=2, nPCG=200)),          # ragged group: compact 1-D cooperative grid
          ("C4", 1, dict(nCont=1, nGN=1, nPCG=60))]           # streaming back-end (fence-free wide accumulators)
for cfg, B, kw in cases:
    pairs = [synth.config(cfg, i) for i in range(B)]
    W, H = pairs[0].W, pairs[0].H
    problems = [(p, m) for p in pairs for m in p.masks]
    b = lib.Batch(W, H, len(problems), kw["nCont"], kw["nGN"], kw["nPCG"], lib.BACKEND_AUTO if cfg == "C4" else lib.BACKEND_RESIDENT)
    ref = None
    t0 = time.time()
    for r in range(reps):
        outs = [b.submit(i, p.rgb, m, p.matches) for i, (p, m) in enumerate(problems)]
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
