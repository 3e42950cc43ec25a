"""Bounding the unpinnable solve.  The reference solver cannot be built here, so bit-level parity of the
solve is unpinned; what is shown instead is that an implementation which differs from the contract path in everything the
reference leaves unspecified -- oracle/arap_literal.c: the reference's unfused 3-kernel schedule, residual-centric
derivatives, libm sinf/cosf, compiler-chosen FMA contraction, fp32 per-warp partial sums added in a seeded SHUFFLED order
like the reference's float atomics (ARAP/API/src/util.t:528-531, 612-623) -- lands inside the north-star tolerances
(flow 1e-3 px mean EPE, final energy 1e-4 relative) of the contract path (oracle/arap_oracle.c == the CUDA kernels, bit for
bit) on DeepMatching-like inputs.  CPU only."""
import glob
import json
import os

import numpy as np
import pytest

from arap_flow_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_EPE, TOL_E = 1e-3, 1e-4


def _cmp(mask, Xa, ca, Xb, cb):
    act = mask == 0
    d = np.hypot(Xa[..., 0] - Xb[..., 0], Xa[..., 1] - Xb[..., 1])[act]
    return d.mean(), abs(float(ca[-1, -1]) - float(cb[-1, -1])) / abs(float(cb[-1, -1]))


@pytest.mark.parametrize("W,H,nseg,fd,seed,kw", [(160, 128, 1, 2, 5, dict(nCont=4, nGN=3, nPCG=150)),
                                                  (192, 108, 4, 3, 9, dict(nCont=3, nGN=3, nPCG=120))])
def test_literal_implementation_lands_inside_the_tolerances(oracle, W, H, nseg, fd, seed, kw):
    sp = synth.synth(W, H, nseg, fd, seed)
    mask = sp.masks[-1]
    Xo, Ao, co = oracle.solve(mask, sp.matches, **kw)
    runs = []
    for s in (1, 2, 3):
        Xl, Al, cl = oracle.literal_solve(mask, sp.matches, seed=s, **kw)
        epe, de = _cmp(mask, Xl, cl, Xo, co)
        assert epe < TOL_EPE and de < TOL_E, (s, epe, de)
        assert (Xl[mask != 0] == oracle.grid(W, H)[mask != 0]).all()       # excluded pixels never move
        runs.append(Xl)
    # the shuffled arrival order really changes the arithmetic (otherwise the bound would be vacuous) ...
    assert any(not np.array_equal(runs[0], r) for r in runs[1:])
    # ... and a fixed order is reproducible
    a = oracle.literal_solve(mask, sp.matches, seed=7, flags=oracle.LIT_ORDERED_ATOMICS, **kw)[0]
    b = oracle.literal_solve(mask, sp.matches, seed=8, flags=oracle.LIT_ORDERED_ATOMICS, **kw)[0]
    assert np.array_equal(a, b)


def test_sincos_table_equals_per_iteration_evaluation(oracle):
    """The reference evaluates sin/cos of the stencil's angles in every PCGStep1; Angle is constant inside a Gauss-Newton
    step, so a per-step table of the same libm values is the same arithmetic: bit-identical here."""
    sp = synth.synth(96, 80, 1, 2, 3)
    kw = dict(nCont=2, nGN=2, nPCG=40, seed=4)
    Xa, Aa, ca = oracle.literal_solve(sp.masks[0], sp.matches, **kw)
    Xb, Ab, cb = oracle.literal_solve(sp.masks[0], sp.matches, flags=oracle.LIT_SINCOS_EVERY_ITER, **kw)
    assert np.array_equal(Xa, Xb) and np.array_equal(Aa, Ab) and np.array_equal(ca, cb)


def test_committed_full_schedule_bounds():
    """profiles/r2_literal_bound*.json (tools/literal_bound.py, minutes of CPU per case): the FULL 19 x 8 x 400 schedule on
    the BASELINE configurations, three shuffle seeds each."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_literal_bound*.json")))
    assert files
    seen = set()
    for f in files:
        db = json.load(open(f))
        for name, ent in db.items():
            if name.startswith("_"):
                continue
            assert ent["schedule"] == [19, 8, 400]
            vs = ent["literal_vs_contract"]
            assert len(vs) >= 3
            seen.add(name.split(":")[0])
            if name == "cat512":
                # 9 hand-placed constraints: the fixed-budget trajectory is chaotic -- ANY two rounding orders
                # differ by a few tenths of a pixel, the same distance as oracle <-> shipped golden (0.27 px).  Reported, not gated.
                assert all(0.02 < v["mean_epe_px"] < 1.0 for v in vs.values())
                assert all(0.02 < v["mean_epe_px"] < 1.0 for v in ent["literal_seed_vs_seed"].values())
                continue
            for seed, v in vs.items():
                assert v["mean_epe_px"] < TOL_EPE and v["rel_energy_diff"] < TOL_E, (name, seed, v)
    assert {"C1", "C2"} <= seen
