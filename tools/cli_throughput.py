#!/usr/bin/env python
"""End-to-end throughput of the arap_deform binary on a list file of synthetic 854x480 pairs (files in, files out)."""
import os, sys, time, tempfile, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arap_flow_b200 import driver, flowio, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
d = tempfile.mkdtemp(prefix="arapcli_")
items = []
for i in range(n):
    sp = synth.config("C1", i)
    p = [os.path.join(d, f"{i}_{k}") for k in ("rgb.png", "msk.png", "cstr.txt", "out.flo", "wrgb.png", "wmsk.png")]
    flowio.write_png(p[0], sp.rgb)
    flowio.write_png(p[1], np.repeat(sp.masks[0][..., None], 3, axis=2))
    flowio.write_constraints(p[2], sp.matches)
    items.append(tuple(p))
lst = os.path.join(d, "list.txt")
driver.write_list_file(lst, items)
for rtol in (None, "1e-3"):
    env = dict(os.environ, ARAP_PLAN=driver.PLAN, ARAP_TIMING="1")
    if rtol:
        env["ARAP_PCG_RTOL"] = rtol
    t0 = time.time()
    subprocess.check_call([driver.ARAP_BIN, lst], env=env, stdout=subprocess.DEVNULL)
    dt = time.time() - t0
    print(f"arap_deform{' ARAP_PCG_RTOL=' + rtol if rtol else ''}: {n} pairs (PNG in, .flo + PNG out) in {dt:.2f} s = "
          f"{n / dt:.2f} pairs/s including process start and plan build")
