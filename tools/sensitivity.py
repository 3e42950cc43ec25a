#!/usr/bin/env python
"""How far can an implementation that differs from ours only in rounding order be?  The reference's own solver (fp32
atomics in arrival order, LLVM-chosen association) cannot be run here, so bit-level parity with it is unpinned
(DESIGN.md section 2).  This tool measures the CONDITIONING of the full 19x8x400 schedule instead: the same problem is
solved through Opt.h with the two weights moved by one ulp (four combinations), and the flows are compared with the
unperturbed one.  On DeepMatching-like inputs the spread is orders of magnitude below the 1e-3 px parity tolerance; on
the 9-constraint README example it is not (chaotic regime).

  tools/sensitivity.py [C1|C3|cat512] [nCont nGN nPCG]
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from arap_flow_b200 import flowio, lib, synth


def problem(cfg):
    if cfg == "cat512":
        g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
        mask = flowio.read_png_mask_red(os.path.join(g, "cat512_iMsk.png"))
        return mask, flowio.read_constraints(os.path.join(g, "cat512_iCstr.txt"))
    sp = synth.config(cfg)
    return sp.masks[0], sp.matches


def constraint_image(mask, matches, alpha):
    """CombinedSolver.h:223-242 + main.cpp:130-136 in numpy (later entries overwrite, border pins last)."""
    H, W = mask.shape
    Cn = np.full((H, W, 2), -1.0, np.float32)
    m = np.asarray(matches, np.int64).reshape(-1, 4)
    ys, xs = np.mgrid[0:H, 0:W]
    border = (ys == 0) | (xs == 0) | (ys == H - 1) | (xs == W - 1)
    pins = np.stack([xs[border], ys[border], xs[border], ys[border]], 1)
    a = np.float32(alpha)
    for x1, y1, x2, y2 in np.concatenate([m, pins]):
        if 0 <= x1 < W and 0 <= y1 < H and mask[y1, x1] == 0:
            Cn[y1, x1] = ((np.float32(1) - a) * np.float32(x1) + a * np.float32(x2),
                          (np.float32(1) - a) * np.float32(y1) + a * np.float32(y2))
    return Cn


def solve(mask, cimgs, wf, wr, nGN, nPCG):
    L = lib.load()
    H, W = mask.shape
    dev = torch.device("cuda:0")
    ys, xs = np.mgrid[0:H, 0:W]
    U = np.ascontiguousarray(np.stack([xs, ys], -1).astype(np.float32))
    tU = torch.from_numpy(U).to(dev)
    tX = tU.clone()
    tA = torch.zeros((H, W), dtype=torch.float32, device=dev)
    tM = torch.from_numpy(mask.astype(np.float32)).to(dev)
    st = L.Opt_NewState(lib.OptInitializationParameters(0, 0, 0, 0))
    prob = L.Opt_ProblemDefine(st, os.path.join(os.path.dirname(lib.LIB_PATH), "arap_plan.t").encode(), b"gaussNewtonGPU")
    plan = L.Opt_ProblemPlan(st, prob, (C.c_uint * 2)(W, H))
    cwf, cwr = C.c_float(float(wf)), C.c_float(float(wr))
    n1, n2 = C.c_uint(nGN), C.c_uint(nPCG)
    cost = 0.0
    for Cn in cimgs:
        tC = torch.from_numpy(Cn).to(dev)
        torch.cuda.synchronize()
        L.Opt_SetSolverParameter(st, plan, b"nIterations", C.byref(n1))
        L.Opt_SetSolverParameter(st, plan, b"lIterations", C.byref(n2))
        pp = (C.c_void_p * 7)(tX.data_ptr(), tA.data_ptr(), tU.data_ptr(), tC.data_ptr(), tM.data_ptr(),
                              C.cast(C.byref(cwf), C.c_void_p), C.cast(C.byref(cwr), C.c_void_p))
        L.Opt_ProblemSolve(st, plan, pp)
        cost = L.Opt_ProblemCurrentCost(st, plan)
    L.Opt_PlanFree(st, plan)
    L.Opt_ProblemDelete(st, prob)
    return (tX - tU).cpu().numpy(), float(cost)


def run(cfg, nCont=19, nGN=8, nPCG=400):
    mask, matches = problem(cfg)
    cimgs = [constraint_image(mask, matches, (t + 1) / np.float32(nCont)) for t in range(nCont)]
    wf0, wr0 = np.sqrt(np.float32(100.0)), np.sqrt(np.float32(0.01))
    base, c0 = solve(mask, cimgs, wf0, wr0, nGN, nPCG)
    act = mask == 0
    rows = []
    for df, dr in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        wf = np.nextafter(wf0, np.float32(np.inf * df)) if df else wf0
        wr = np.nextafter(wr0, np.float32(np.inf * dr)) if dr else wr0
        fl, c = solve(mask, cimgs, wf, wr, nGN, nPCG)
        epe = np.linalg.norm(fl - base, axis=-1)[act]
        rows.append(dict(ulp_wf=df, ulp_wr=dr, mean_epe_px=float(epe.mean()), max_epe_px=float(epe.max()),
                         rel_cost_diff=abs(c - c0) / abs(c0)))
    out = dict(workload=cfg, schedule=[nCont, nGN, nPCG], active_px=int(act.sum()),
               mean_flow_px=float(np.linalg.norm(base, axis=-1)[act].mean()), final_cost=c0, perturbations=rows,
               worst_mean_epe_px=max(r["mean_epe_px"] for r in rows), worst_rel_cost_diff=max(r["rel_cost_diff"] for r in rows))
    return out


if __name__ == "__main__":
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
    sched = [int(v) for v in sys.argv[2:5]] if len(sys.argv) >= 5 else [19, 8, 400]
    res = run(cfg, *sched)
    print(json.dumps(res))
