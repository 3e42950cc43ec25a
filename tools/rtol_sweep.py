#!/usr/bin/env python
"""Opt-in convergence-aware schedule: speed / accuracy of pcg_rtol on the bench workload, full
19x8x400 budget as the reference point.  tools/rtol_sweep.py [C1] [pairs]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arap_flow_b200 import lib, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if cfg == "cat512":   # the reference's README example: 9 hand-placed constraints, 101 k active px (chaotic regime)
    from types import SimpleNamespace
    from arap_flow_b200 import flowio
    g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    one = SimpleNamespace(W=512, H=512, rgb=flowio.read_png_rgb(os.path.join(g, "cat512_iRGB.png")),
                          masks=[flowio.read_png_mask_red(os.path.join(g, "cat512_iMsk.png"))],
                          matches=flowio.read_constraints(os.path.join(g, "cat512_iCstr.txt")))
    pairs = [one] * B
else:
    pairs = [synth.config(cfg, i) for i in range(B)]
W, H = pairs[0].W, pairs[0].H
b = lib.Batch(W, H, B, 19, 8, 400, lib.BACKEND_RESIDENT)
ref = None
rows = []
for rtol, gtol in ((0.0, 0.0), (1e-5, 0.0), (1e-4, 0.0), (1e-3, 0.0), (1e-2, 0.0), (1e-1, 0.0),
                   (1e-3, 1e-4), (1e-3, 1e-3), (1e-3, 1e-2), (1e-2, 1e-2)):
    b.set_option("pcg_rtol", rtol)
    b.set_option("gn_rtol", gtol)
    for _ in range(2):
        outs = [b.submit(i, p.rgb, p.masks[0], p.matches) for i, p in enumerate(pairs)]
        b.run()
    ms = b.timing_ms()["solve"]
    flows = [o["flow"].copy() for o in outs]
    costs = [float(o["costs"][-1, -1]) for o in outs]
    if ref is None:
        ref = flows
    epe = [float(np.linalg.norm(f - r, axis=-1)[p.masks[0] == 0].mean()) for f, r, p in zip(flows, ref, pairs)]
    mag = [float(np.linalg.norm(r, axis=-1)[p.masks[0] == 0].mean()) for r, p in zip(ref, pairs)]
    rows.append(dict(pcg_rtol=rtol, gn_rtol=gtol, solve_ms_per_pair=ms / B, pairs_per_s=1000.0 * B / ms, mean_epe_vs_full_px=float(np.mean(epe)),
                     max_pair_epe_px=float(np.max(epe)), mean_flow_px=float(np.mean(mag)), final_cost_mean=float(np.mean(costs))))
    print(json.dumps(rows[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/rtol_sweep_%s.json" % cfg, "w"), indent=1)
