#!/usr/bin/env python
"""Bound the unpinnable solve: how far from the contract path (oracle == CUDA kernels, bit for bit) does an
implementation land that differs in everything the reference leaves unspecified?

Runs oracle/arap_literal.c (reference-like unfused 3-kernel schedule, residual-centric derivatives, libm sinf/cosf,
compiler-chosen FMA contraction, fp32 per-warp partial sums added in a seeded SHUFFLED order -- what the reference's
float atomics do, ARAP/API/src/util.t:528-531, 612-623) at the FULL 19 x 8 x 400 schedule for several shuffle seeds
and compares flow and final energy with oracle/arap_oracle.c on the same inputs.  North-star tolerances: mean EPE
< 1e-3 px, final energy within 1e-4 relative.  CPU only; minutes per case.

  python tools/literal_bound.py [--cases C1 C2 C3 cat512] [--seeds 1 2 3] [--out profiles/r2_literal_bound.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arap_flow_b200 import flowio, synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def problems(case):
    """yield (name, mask, matches)"""
    if case == "cat512":
        msk = flowio.read_png_mask_red(os.path.join(GOLD, "cat512_iMsk.png"))
        cstr = flowio.read_constraints(os.path.join(GOLD, "cat512_iCstr.txt"))
        yield "cat512", msk, np.asarray(cstr, np.int32)
        return
    sp = synth.config(case)
    for s, m in enumerate(sp.masks):
        yield f"{case}:{sp.seed}:{s}", m, sp.matches


def compare(mask, Xa, ca, Xb, cb):
    act = mask == 0
    d = np.hypot(Xa[..., 0] - Xb[..., 0], Xa[..., 1] - Xb[..., 1])
    return {"mean_epe_px": float(d[act].mean()), "max_epe_px": float(d[act].max()),
            "rel_energy_diff": float(abs(float(ca[-1, -1]) - float(cb[-1, -1])) / abs(float(cb[-1, -1]))),
            "final_energy": float(ca[-1, -1])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", nargs="+", default=["C1", "C2", "C3", "cat512"])
    ap.add_argument("--seeds", nargs="+", type=int, default=[1, 2, 3])
    ap.add_argument("--schedule", nargs=3, type=int, default=[19, 8, 400])
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_literal_bound.json"))
    args = ap.parse_args()
    nCont, nGN, nPCG = args.schedule
    O.build()
    db = json.load(open(args.out)) if os.path.exists(args.out) else {}
    db["_what"] = ("oracle/arap_literal.c (unfused reference-like schedule, residual-centric derivatives, libm sinf/cosf, "
                   "-ffp-contract=fast, fp32 per-warp sums added in a seeded shuffled order) vs oracle/arap_oracle.c "
                   "(the arithmetic contract the CUDA kernels match bit for bit); full schedule unless stated")
    db["_tolerances"] = {"mean_epe_px": 1e-3, "rel_energy_diff": 1e-4}
    for case in args.cases:
        for name, mask, matches in problems(case):
            if name in db and all(str(s) in db[name]["literal_vs_contract"] for s in args.seeds):
                continue
            t0 = time.time()
            Xo, Ao, co = O.solve(mask, matches, nCont=nCont, nGN=nGN, nPCG=nPCG)
            ent = db.setdefault(name, {"schedule": [nCont, nGN, nPCG], "active_px": int((mask == 0).sum()),
                                       "n_matches": int(len(matches)), "contract_final_energy": float(co[-1, -1]),
                                       "mean_flow_px": float(np.hypot(*np.moveaxis(O.flow(Xo), -1, 0))[mask == 0].mean()),
                                       "literal_vs_contract": {}, "literal_seed_vs_seed": {}})
            lits = {}
            for s in args.seeds:
                Xl, Al, cl = O.literal_solve(mask, matches, nCont=nCont, nGN=nGN, nPCG=nPCG, seed=s)
                lits[s] = (Xl, cl)
                ent["literal_vs_contract"][str(s)] = compare(mask, Xl, cl, Xo, co)
            ss = sorted(lits)
            for a, b in zip(ss, ss[1:]):
                ent["literal_seed_vs_seed"][f"{a}-{b}"] = compare(mask, lits[a][0], lits[a][1], lits[b][0], lits[b][1])
            ent["cpu_seconds"] = round(time.time() - t0, 1)
            tmp = args.out + ".tmp"
            with open(tmp, "w") as f:
                json.dump(db, f, indent=1, sort_keys=True)
            os.replace(tmp, args.out)
            print(name, json.dumps(ent["literal_vs_contract"]), flush=True)


if __name__ == "__main__":
    main()
