#!/usr/bin/env python
"""Throughput of the batched pipeline vs batch size (GPU box only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arap_flow_b200 import lib, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,3,4").split(",")]
nCont = int(sys.argv[3]) if len(sys.argv) > 3 else 2
base = {"C1": (854, 480, 1, 1, 1000), "C3": (1024, 436, 1, 5, 3000), "C0": (64, 64, 1, 1, 0)}[cfg]
pairs = [synth.synth(base[0], base[1], base[2], base[3], base[4] + i) for i in range(max(sizes))]
for B in sizes:
    b = lib.Batch(base[0], base[1], B, nCont, 8, 400)
    for rep in range(2):
        outs = [b.submit(i, pairs[i].rgb, pairs[i].masks[0], pairs[i].matches) for i in range(B)]
        t0 = time.perf_counter()
        b.run()
        dt = time.perf_counter() - t0
    tm = b.timing_ms()
    its = nCont * 8 * 400
    print(f"{cfg} B={B}: wall {dt*1e3:.1f} ms, solve {tm['solve']:.1f} ms -> {tm['solve']*1e3/its:.2f} us per PCG iteration (all {B}), "
          f"{tm['solve']*1e3/its/B:.2f} us per problem-iteration, full-schedule est {B/(tm['solve']*19/nCont/1e3 + tm['warp']/1e3):.2f} pairs/s, launches {b.launches()}")
    b.close()
