#!/usr/bin/env python
"""Files in -> files out throughput of the arap_deform binary, the way para_gen.py drives it.

  python tools/cli_throughput.py dispatch [--workload C1 --dispatches 8 --pairs 8]
      para_gen-sized dispatches on ONE GPU: a fresh solver process per dispatch (para_gen.py:178-200) against the
      resident worker (`arap_deform --serve`, clients with ARAP_SERVER).  Process start, PNG decode, H2D, solve,
      warp, D2H, PNG/.flo encode are all inside the measured wall time.
  python tools/cli_throughput.py shard [--workload C3 --pairs 64 --gpus 1 2 4 8]
      BASELINE config C3: 64 pairs sharded para_gen --gpu style over N GPUs, strong scaling; one solver
      process per GPU, and the same through one resident worker per GPU.
Prints one JSON line per measurement.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from arap_flow_b200 import driver, flowio, synth  # noqa: E402


def write_items(d, workload, n):
    items = []
    for i in range(n):
        sp = synth.config(workload, i)
        for s, mask in enumerate(sp.masks):
            p = [os.path.join(d, f"{i}_{s}_{k}") for k in ("rgb.png", "msk.png", "cstr.txt", "out.flo", "wrgb.png", "wmsk.png")]
            flowio.write_png(p[0], sp.rgb)
            flowio.write_png(p[1], np.repeat(mask[..., None], 3, axis=2))
            flowio.write_constraints(p[2], sp.matches)
            items.append(tuple(p))
    return items


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["dispatch", "shard"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--dispatches", type=int, default=8)
    ap.add_argument("--pairs", type=int, default=None)
    ap.add_argument("--gpus", type=int, nargs="+", default=[1])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--served-only", action="store_true", help="dispatch mode: skip the process-per-dispatch baseline")
    ap.add_argument("--no-warm", action="store_true", help="start the workers without --warm WxH")
    a = ap.parse_args()
    d = tempfile.mkdtemp(prefix="arapcli_")
    tmp = os.path.join(d, "tmp")
    try:
        if a.mode == "dispatch":
            wl = a.workload or "C1"
            per = a.pairs or 8
            items = write_items(d, wl, a.dispatches * per)
            dispatches = [items[k * per:(k + 1) * per] for k in range(a.dispatches)]
            batch = a.batch or (8 if per >= 8 else per)
            # (a) the reference's way: one process per dispatch
            if not a.served_only:
                t0 = time.time()
                for disp in dispatches:
                    driver.do_arap(disp, 0, tmp, batch=batch)
                dt = time.time() - t0
                print(json.dumps({"mode": "process per dispatch", "workload": wl, "dispatches": a.dispatches, "pairs_per_dispatch": per,
                                  "batch": batch, "seconds": dt, "pairs_per_s": len(items) / dt}), flush=True)
            # (b) resident worker
            spool = os.path.join(d, "spool")
            t0 = time.time()
            wh = None if a.no_warm else (synth.config(wl).W, synth.config(wl).H)
            with driver.Server(0, spool, batch=batch, warm=wh):
                t_up = time.time() - t0
                t1 = time.time()
                for disp in dispatches:
                    driver.do_arap(disp, 0, tmp, server=spool)
                dt = time.time() - t1
            print(json.dumps({"mode": "resident worker (arap_deform --serve)", "workload": wl, "dispatches": a.dispatches,
                              "pairs_per_dispatch": per, "batch": batch, "warm": wh, "seconds": dt, "pairs_per_s": len(items) / dt,
                              "worker_start_seconds_once": t_up, "pairs_per_s_including_worker_start": len(items) / (dt + t_up)}),
                  flush=True)
        else:
            wl = a.workload or "C3"
            n = a.pairs or 64
            items = write_items(d, wl, n)
            base = None
            for g in a.gpus:
                gpus = list(range(g))
                batch = a.batch or 8
                dt = driver.run_sharded(items, gpus, tmp, batch=batch)
                spools = [os.path.join(d, f"spool{r}") for r in gpus]
                wh = None if a.no_warm else (synth.config(wl).W, synth.config(wl).H)
                servers = [driver.Server(r, spools[r], batch=batch, warm=wh) for r in gpus]
                try:
                    dts = driver.run_sharded(items, gpus, tmp, servers=spools)
                finally:
                    for s in servers:
                        s.close()
                base = base or (dt, dts)
                print(json.dumps({"mode": "shard", "workload": wl, "pairs": n, "gpus": g, "scaling": "strong", "batch": batch, "warm": wh,
                                  "process_per_gpu": {"seconds": dt, "pairs_per_s": n / dt, "efficiency_vs_1gpu": base[0] / (g * dt)},
                                  "resident_worker_per_gpu": {"seconds": dts, "pairs_per_s": n / dts,
                                                              "efficiency_vs_1gpu": base[1] / (g * dts)}}), flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
