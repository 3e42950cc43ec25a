"""Protocol stress: the resident kernel's fence-free grid barrier and tagged halo exchange, and the streaming back-end's
wide accumulators, must give the SAME bits on every repetition.  compute-sanitizer's racecheck is closed on this pool, so
determinism under repetition -- together with the bit-exact oracle comparisons and the in-kernel watchdogs -- is the guard
against a torn or stale halo word.  (Was tools/stress_resident.py; VERDICT r1 "what's weak" 3.)"""
import numpy as np
import pytest

from arap_flow_b200 import lib, synth

pytestmark = pytest.mark.gpu

REPS = 30
CASES = [
    ("C1", 4, dict(nCont=1, nGN=2, nPCG=200), lib.BACKEND_RESIDENT, {}),     # (G, 4) grid, 128-register variant
    ("C1", 3, dict(nCont=1, nGN=2, nPCG=200), lib.BACKEND_RESIDENT, {}),     # (G, 3) grid, 168-register variant
    ("C3", 3, dict(nCont=1, nGN=2, nPCG=200), lib.BACKEND_RESIDENT, {}),
    ("C0", 8, dict(nCont=2, nGN=2, nPCG=100), lib.BACKEND_RESIDENT, {}),
    ("C2", 2, dict(nCont=1, nGN=2, nPCG=200), lib.BACKEND_RESIDENT, {}),     # ragged group: compact 1-D cooperative grid
    ("C4", 1, dict(nCont=1, nGN=1, nPCG=60), lib.BACKEND_AUTO, {}),          # streaming back-end
    # opt-in early exits: the path with the extra separating barrier (solver_resident.cu early exit)
    ("C1", 3, dict(nCont=2, nGN=3, nPCG=200), lib.BACKEND_RESIDENT, {"pcg_rtol": 1e-2}),
    ("C2", 1, dict(nCont=2, nGN=3, nPCG=200), lib.BACKEND_RESIDENT, {"pcg_rtol": 1e-1, "gn_rtol": 1e-2}),
    ("C0", 8, dict(nCont=2, nGN=2, nPCG=0), lib.BACKEND_RESIDENT, {}),       # no PCG iteration at all
]


@pytest.mark.parametrize("cfg,B,kw,backend,opts", CASES)
def test_repeated_batched_solves_are_bit_identical(cfg, B, kw, backend, opts):
    pairs = [synth.config(cfg, i) for i in range(B)]
    W, H = pairs[0].W, pairs[0].H
    problems = [(p, m) for p in pairs for m in p.masks]
    b = lib.Batch(W, H, len(problems), kw["nCont"], kw["nGN"], kw["nPCG"], backend)
    for k, v in opts.items():
        b.set_option(k, v)
    ref = None
    for r in range(REPS):
        outs = [b.submit(i, p.rgb, m, p.matches) for i, (p, m) in enumerate(problems)]
        b.run()     # raises on a watchdog abort (a stalled halo fetch / barrier would end here, not hang)
        cur = [(o["flow"].copy(), o["costs"].copy()) for o in outs]
        if ref is None:
            ref = cur
            assert all(np.isfinite(f).all() and np.isfinite(c).all() for f, c in ref)
        else:
            for (f0, c0), (f1, c1) in zip(ref, cur):
                assert np.array_equal(f0, f1) and np.array_equal(c0, c1), f"{cfg}: repetition {r} differs from repetition 0"
    b.close()
