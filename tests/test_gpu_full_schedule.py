"""The launch shapes bench.py times, at the FULL 19 x 8 x 400 schedule, against committed oracle fixtures.

tests/golden/full_schedule.json (tools/make_full_schedule_golden.py) holds, per problem, the oracle's 19 x 9 cost
table (bit patterns) and SHA-256 digests of flow / warped RGB / warped mask.  Every test here runs what a benchmark
line runs -- same Batch size, hence the same cooperative grid, kernel variant and co-residency -- and asserts
equality, so the timed shape is also the verified shape (VERDICT r1 "what's weak" 1).  A case whose fixture has not
been generated is reported as a skip, never silently passed.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from arap_flow_b200 import lib, synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full_schedule.json")
AXES = {"C4": (0.46, 0.46), "C1s": (0.15, 0.17)}
SHAPE = {"C0": (64, 64, 1, 1), "C1": (854, 480, 1, 1), "C2": (854, 480, 4, 3), "C3": (1024, 436, 1, 5),
         "C4": (1920, 1080, 1, 1), "C1s": (854, 480, 1, 1)}


def _db():
    with open(GOLD) as f:
        return json.load(f)


def _digest(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        a = a + np.float32(0.0)
    return hashlib.sha256(a.tobytes()).hexdigest()


def _pair(wl, seed):
    W, H, nseg, fd = SHAPE[wl]
    return synth.synth(W, H, nseg, fd, seed, axes=AXES.get(wl))


def _run_and_check(wl, seeds, expect_variant=None, expect_grid_y=None, backend=lib.BACKEND_AUTO, expect_streamed=False,
                   expect_compact=False):
    db = _db()
    pairs = [_pair(wl, s) for s in seeds]
    jobs = [(p, si, m) for p in pairs for si, m in enumerate(p.masks)]
    keys = [f"{wl}:{p.seed}:{si}" for p, si, m in jobs]
    missing = [k for k in keys if k not in db]
    if missing:
        pytest.skip("no full-schedule oracle fixture yet for " + ", ".join(missing))
    nCont, nGN, nPCG = db[keys[0]]["schedule"]
    b = lib.Batch(pairs[0].W, pairs[0].H, len(jobs), nCont, nGN, nPCG, backend)
    outs = [b.submit(i, p.rgb, m, p.matches) for i, (p, si, m) in enumerate(jobs)]
    b.run()
    info = b.launch_info()
    if expect_streamed:
        assert b.resident_count() == 0
    else:
        assert b.resident_count() == len(jobs)
        if expect_variant:
            assert info["variant"] == expect_variant, info
        if expect_grid_y:
            assert info["grid"][1] == expect_grid_y and info["problems_per_launch"] == expect_grid_y, info
        if expect_compact:
            assert info["grid"][1] == 1 and info["problems_per_launch"] > 1, info
    for k, o in zip(keys, outs):
        want = db[k]
        costs_bits = np.ascontiguousarray(o["costs"], np.float32).view(np.uint32)
        assert np.array_equal(costs_bits, np.asarray(want["costs_bits"], np.uint32)), (k, "cost table differs")
        assert _digest(o["flow"]) == want["flow_sha256"], (k, "flow differs")
        assert _digest(o["rgb"]) == want["rgb_sha256"] and _digest(o["mask"]) == want["mask_sha256"], (k, "warp differs")
    b.close()
    return info


# Kernel variants are named by their launch bounds (max threads, min CTAs per SM): with 128-thread CTAs the (384, 1)
# instantiation is the 168-register, spill-free one (three CTAs per SM) and (160, 3) the 128-register one (four).
def test_c1_three_coresident_problems_168_register_variant():
    """launches of three co-resident problems, (G, 3) grid, 168 registers (bench.py --batch 9)"""
    _run_and_check("C1", [1000, 1001, 1002], expect_variant=(384, 1), expect_grid_y=3)


def test_c1_four_coresident_problems_128_register_variant():
    _run_and_check("C1", [1000, 1001, 1002, 1003], expect_variant=(160, 3), expect_grid_y=4)


def test_c1_default_bench_batch_of_eight():
    """exactly bench.py --workload C1 (rank 0): Batch(854, 480, 8), seeds 1000..1007, two launches of four (128 registers)"""
    _run_and_check("C1", list(range(1000, 1008)), expect_variant=(160, 3), expect_grid_y=4)


def test_c1_batch_of_nine():
    """Batch(854, 480, 9), seeds 1000..1008: three launches of three (168 registers)"""
    _run_and_check("C1", list(range(1000, 1009)), expect_variant=(384, 1), expect_grid_y=3)


def test_c1_single_problem_launch():
    """what Opt_ProblemSolve / a one-pair dispatch runs: one problem alone in the grid"""
    _run_and_check("C1", [1000], expect_grid_y=1)


def test_c2_ragged_compact_grid_one_pair():
    """--multseg: the four segments of one pair share one compact 1-D cooperative grid"""
    _run_and_check("C2", [2000], expect_compact=True)


def test_c2_two_pairs_eight_problems():
    """2 pairs = 8 ragged problems in one compact grid (168-register variant)"""
    _run_and_check("C2", [2000, 2001], expect_compact=True)


def test_c2_default_bench_three_pairs_twelve_problems():
    """exactly bench.py --workload C2: 3 pairs = 12 ragged problems in one compact grid (128-register variant)"""
    _run_and_check("C2", [2000, 2001, 2002], expect_compact=True)


def test_c3_batch_three():
    _run_and_check("C3", [3000, 3001, 3002])


def test_c3_batch_of_eight():
    """3 + 3 + 2 (bench.py --workload C3 runs launches of three whatever the batch: test_c3_batch_three is its launch shape)"""
    _run_and_check("C3", list(range(3000, 3008)))


def test_c4_streaming_one_full_continuation_step():
    """1920x1080 through the streaming back-end: 8 Gauss-Newton steps x 400 PCG iterations"""
    _run_and_check("C4", [4000], expect_streamed=True)


def test_c1s_small_objects_share_a_launch():
    """DAVIS-typical small object (8 % coverage)"""
    _run_and_check("C1s", [1000, 1001])


def test_c1s_default_bench_batch_of_fourteen():
    """exactly bench.py --workload C1s: fourteen problems of 41 CTAs in one (41, 14) launch, 128-register variant"""
    _run_and_check("C1s", list(range(1000, 1014)), expect_variant=(160, 3), expect_grid_y=14)


def test_c0_batch_of_eight():
    """bench.py --workload C0: eight 64x64 problems of 3 CTAs each"""
    _run_and_check("C0", list(range(0, 8)))
