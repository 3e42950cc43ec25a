#!/usr/bin/env python
"""The three schedules on one workload (default C1, seed 1000): the reference's fixed budget (gaussNewtonGPU, 19 x 8 x 400),
the opt-in early exits of the tuned kernels (pcg_rtol / gn_rtol) and the opt-in "LMGPU" solver kind (solver_lm.cu).
Prints one JSON line per schedule: wall ms of the batch call (one problem), final energy, mean EPE of the flow against
the fixed budget's.  GPU only."""
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from arap_flow_b200 import lib, synth  # noqa: E402


def run(sp, mask, **opts):
    b = lib.Batch(sp.W, sp.H, 1)
    for k, v in opts.items():
        b.set_option(k, v)
    o = b.submit(0, sp.rgb, mask, sp.matches)
    b.run()                                    # warm-up (module load, buffers)
    t0 = time.perf_counter()
    o = b.submit(0, sp.rgb, mask, sp.matches, out=o)
    b.run()
    dt = time.perf_counter() - t0
    n = b.launches()
    b.close()
    return o, dt, n


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C1"
    sp = synth.config(wl)
    mask = sp.masks[0]
    act = mask == 0
    ref, dt, n = run(sp, mask)
    rows = [("gaussNewtonGPU 19x8x400 (reference schedule)", ref, dt, n)]
    for name, opts in (("gaussNewtonGPU + pcg_rtol 1e-3", dict(pcg_rtol=1e-3)),
                       ("gaussNewtonGPU + pcg_rtol 1e-3 + gn_rtol 1e-4", dict(pcg_rtol=1e-3, gn_rtol=1e-4)),
                       ("LMGPU (default parameters)", dict(lm=1))):
        o, dt, n = run(sp, mask, **opts)
        rows.append((name, o, dt, n))
    for name, o, dt, n in rows:
        d = np.hypot(*(np.moveaxis(o["flow"] - ref["flow"], -1, 0)))[act]
        print(json.dumps({"workload": wl, "schedule": name, "ms_one_problem": round(1e3 * dt, 1),
                          "final_energy": float(o["costs"][-1, -1]), "mean_epe_vs_fixed_budget_px": float(d.mean()),
                          "max_epe_px": float(d.max())}), flush=True)


if __name__ == "__main__":
    main()
