"""The "LMGPU" solver kind (csrc/solver_lm.cu) through the reference's own Opt.h call sequence -- Opt_ProblemDefine(...,
"LMGPU"), Opt_SetSolverParameter, Opt_ProblemInit / Opt_ProblemStep or Opt_ProblemSolve, Opt_ProblemCurrentCost -- against
the oracle's restatement (oracle/arap_oracle.c: arap_oracle_lm_solve), bit for bit: costs, unknowns, trust-region radius,
number of linear iterations, accept / revert / stop decisions, model cost and Q."""
import ctypes as C
import os

import numpy as np
import pytest

from arap_flow_b200 import lib
from tests.helpers import synth_gn_problem

pytestmark = pytest.mark.gpu

PLAN = os.path.join(os.path.dirname(lib.LIB_PATH), "arap_plan.t")
FLOAT_PARAMS = ("min_relative_decrease", "min_trust_region_radius", "max_trust_region_radius", "q_tolerance",
                "function_tolerance", "trust_region_radius", "radius_decrease_factor", "min_lm_diagonal", "max_lm_diagonal")


class LmPlan:
    def __init__(self, W, H, nGN, nPCG, **params):
        L = self.L = lib.load()
        L.arapb200_plan_lm_info.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.arapb200_plan_error.argtypes = [C.c_void_p]
        self.st = L.Opt_NewState(lib.OptInitializationParameters(0, 0, 0, 0))
        self.prob = L.Opt_ProblemDefine(self.st, PLAN.encode(), b"LMGPU")
        self.plan = L.Opt_ProblemPlan(self.st, self.prob, (C.c_uint * 2)(W, H))
        assert self.st and self.prob and self.plan
        self.keep = [C.c_uint(nGN), C.c_uint(nPCG)]
        L.Opt_SetSolverParameter(self.st, self.plan, b"nIterations", C.byref(self.keep[0]))
        L.Opt_SetSolverParameter(self.st, self.plan, b"lIterations", C.byref(self.keep[1]))
        for k, v in params.items():
            ref = C.c_float(v) if k in FLOAT_PARAMS else C.c_int(v)
            assert k in FLOAT_PARAMS or k == "residual_reset_period"
            L.Opt_SetSolverParameter(self.st, self.plan, k.encode(), C.byref(ref))

    def bind(self, pr, oracle):
        import torch
        dev = torch.device("cuda:0")
        self.d = {k: torch.from_numpy(np.ascontiguousarray(pr[k])).to(dev) for k in ("X", "A", "U", "C", "M")}
        self.wf, self.wr = C.c_float(float(oracle.WF)), C.c_float(float(oracle.WR))
        d = self.d
        self.pp = (C.c_void_p * 7)(d["X"].data_ptr(), d["A"].data_ptr(), d["U"].data_ptr(), d["C"].data_ptr(),
                                   d["M"].data_ptr(), C.cast(C.byref(self.wf), C.c_void_p),
                                   C.cast(C.byref(self.wr), C.c_void_p))

    def cost(self):
        return np.float32(self.L.Opt_ProblemCurrentCost(self.st, self.plan))

    def run_stepwise(self):
        """launchProfiledSolve's loop (ARAP/shared/OptUtils.h:47-64): init, then step until it returns 0."""
        L = self.L
        L.Opt_ProblemInit(self.st, self.plan, self.pp)
        costs, stats = [self.cost()], []
        while True:
            more = L.Opt_ProblemStep(self.st, self.plan, self.pp)
            info = (C.c_float * 6)()
            if not more and len(costs) - 1 >= self.keep[0].value:
                break                                    # the iteration budget ended the loop: no step was taken
            assert L.arapb200_plan_lm_info(self.plan, info) == 0
            stats.append(list(info))
            costs.append(self.cost())
            if not more:
                break
        assert L.arapb200_plan_error(self.plan) == 0
        return np.float32(costs), np.float32(stats).reshape(-1, 6)

    def unknowns(self):
        return self.d["X"].cpu().numpy(), self.d["A"].cpu().numpy()

    def close(self):
        self.L.Opt_PlanFree(self.st, self.plan)
        self.L.Opt_ProblemDelete(self.st, self.prob)


def _check(oracle, pr, nGN, nPCG, **params):
    Xo, Ao, co, so = oracle.lm_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], nGN, nPCG, **params)
    p = LmPlan(pr["W"], pr["H"], nGN, nPCG, **params)
    p.bind(pr, oracle)
    costs, stats = p.run_stepwise()
    X, A = p.unknowns()
    p.close()
    n = len(so)
    assert stats.shape == so.shape, (stats, so)
    assert np.array_equal(stats.view(np.uint32), so.view(np.uint32)), (stats, so)
    assert np.array_equal(costs.view(np.uint32), co[:n + 1].view(np.uint32)), (costs, co)
    assert np.array_equal(X, Xo) and np.array_equal(A, Ao)
    return co, so


@pytest.mark.parametrize("W,H,seed,fd,nGN,nPCG", [(96, 80, 42, 2, 6, 100), (160, 120, 7, 3, 5, 64), (203, 77, 11, 1, 20, 400)])
def test_lm_defaults_bit_exact(oracle, W, H, seed, fd, nGN, nPCG):
    pr = synth_gn_problem(oracle, W, H, seed=seed, fd=fd)
    co, so = _check(oracle, pr, nGN, nPCG)
    assert np.all(so[:, 1] < nPCG) or nPCG < 100            # the Q test ended the linear loops on the device
    assert co[-1] < co[0]


def test_lm_function_tolerance_stop(oracle):
    pr = synth_gn_problem(oracle, 64, 48, seed=3, fd=2)
    co, so = _check(oracle, pr, 20, 400)
    assert so[-1, 2] == 2.0 and len(so) < 20


def test_lm_reverts_and_minimum_radius_stop(oracle):
    pr = synth_gn_problem(oracle, 96, 80, seed=6, fd=2)
    co, so = _check(oracle, pr, 10, 30, min_relative_decrease=2.0, min_trust_region_radius=100.0)
    assert list(so[:, 2]) == [0.0, 0.0, 0.0, 3.0]


def test_lm_residual_refresh_and_full_linear_budget(oracle):
    pr = synth_gn_problem(oracle, 128, 96, seed=8, fd=2)
    co, so = _check(oracle, pr, 3, 50, residual_reset_period=3, q_tolerance=-1e30)
    assert list(so[:, 1]) == [50.0, 50.0, 50.0]
    _check(oracle, pr, 2, 33, residual_reset_period=1, q_tolerance=-1e30, trust_region_radius=10.0)
    # more linear iterations than the accumulator ring holds (64): the ring wraps twice
    co, so = _check(oracle, pr, 1, 150, q_tolerance=-1e30)
    assert list(so[:, 1]) == [150.0]


def test_lm_general_urshape_and_weights(oracle):
    """any UrShape image (d = u_i - u_j a runtime vector) and other weights"""
    pr = synth_gn_problem(oracle, 120, 90, seed=12, fd=2)
    rng = np.random.default_rng(5)
    U = (pr["U"] * np.float32(1.25) + rng.normal(0, 0.05, pr["U"].shape)).astype(np.float32)
    pr = dict(pr, U=U, X=U.copy())
    pr["C"] = np.where(pr["C"] >= 0, pr["C"] * np.float32(1.25), pr["C"]).astype(np.float32)
    _check(oracle, pr, 4, 80)


def test_lm_solve_entry_point_and_reuse_of_a_plan(oracle):
    """Opt_ProblemSolve = init + steps (o.t:2548-2551); a second solve on the same plan restarts the trust region from the
    solver parameters (:996-1001) and re-saves the Jacobi scaling."""
    pr = synth_gn_problem(oracle, 96, 80, seed=21, fd=2)
    p = LmPlan(96, 80, 5, 100)
    for rep in range(2):
        p.bind(pr, oracle)
        p.L.Opt_ProblemSolve(p.st, p.plan, p.pp)
        Xo, Ao, co, so = oracle.lm_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], 5, 100)
        X, A = p.unknowns()
        assert p.cost() == co[len(so)] and np.array_equal(X, Xo) and np.array_equal(A, Ao), rep
        pr = synth_gn_problem(oracle, 96, 80, seed=22, fd=3)
    p.close()


def test_lm_c1_sized_problem(oracle):
    """a benchmark-sized image (C1, 854x480, 135 k active pixels): 3 LM steps against the oracle"""
    from arap_flow_b200 import synth
    sp = synth.config("C1")
    m = oracle.with_border_pins(sp.matches, sp.W, sp.H)
    U = oracle.grid(sp.W, sp.H)
    pr = dict(W=sp.W, H=sp.H, M=sp.masks[0].astype(np.float32), U=U, X=U.copy(), A=np.zeros((sp.H, sp.W), np.float32),
              C=oracle.constraint_image(sp.masks[0], m, 1.0 / 19))
    _check(oracle, pr, 3, 400)


def test_lm_through_the_batch_pipeline_and_cli_option(oracle, tmp_path):
    """arapb200_batch_set_option("lm", 1) / ARAP_SOLVER=LMGPU: the whole arap_deform schedule with the LM solver kind ==
    the oracle's (constraint image per continuation step + arap_oracle_lm_solve), flow / warp / cost table bit for bit."""
    import subprocess
    from arap_flow_b200 import flowio, synth
    sp = synth.synth(160, 120, 2, 2, 31)
    nCont, nGN, nPCG = 5, 4, 100
    b = lib.Batch(160, 120, 2, nCont, nGN, nPCG)
    b.set_option("lm", 1)
    outs = [b.submit(i, sp.rgb, m, sp.matches) for i, m in enumerate(sp.masks)]
    b.run()
    its_total = 0
    for m, o in zip(sp.masks, outs):
        Xo, Ao, co, its = oracle.solve_lm(m, sp.matches, nCont, nGN, nPCG)
        its_total += sum(its)
        assert np.array_equal(o["costs"].view(np.uint32), co.view(np.uint32))
        assert np.array_equal(o["flow"], oracle.flow(Xo))
        rgb_o, m_o, _ = oracle.warp(Xo, sp.rgb, m)
        assert np.array_equal(o["rgb"], rgb_o) and np.array_equal(o["mask"], m_o)
    assert its_total < 0.5 * 2 * nCont * nGN * nPCG          # the schedule really is convergence-aware
    b.set_option("lm", 0)                                    # and back: the default solver kind on the same batch
    o = b.submit(0, sp.rgb, sp.masks[0], sp.matches)
    b.run()
    Xo, Ao, co = oracle.solve(sp.masks[0], sp.matches, nCont, nGN, nPCG)
    assert np.array_equal(o["costs"], co) and np.array_equal(o["flow"], oracle.flow(Xo))
    b.close()
    # the command-line tool: ARAP_SOLVER=LMGPU, full default schedule on a small image
    sp = synth.synth(96, 80, 1, 2, 32)
    p = {k: str(tmp_path / k) for k in ("rgb.png", "msk.png", "cstr.txt", "out.flo", "wrgb.png", "wmsk.png")}
    flowio.write_png(p["rgb.png"], sp.rgb)
    flowio.write_png(p["msk.png"], np.repeat(sp.masks[0][..., None], 3, axis=2))
    flowio.write_constraints(p["cstr.txt"], sp.matches)
    exe = os.path.join(os.path.dirname(lib.LIB_PATH), "bin", "arap_deform")
    env = dict(os.environ, ARAP_PLAN=PLAN, ARAP_SOLVER="LMGPU")
    r = subprocess.run([exe] + [p[k] for k in ("rgb.png", "msk.png", "cstr.txt", "out.flo", "wrgb.png", "wmsk.png")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    Xo, Ao, co, its = oracle.solve_lm(sp.masks[0], sp.matches)
    assert np.array_equal(flowio.read_flo(p["out.flo"]), oracle.flow(Xo))
    r = subprocess.run([exe, p["rgb.png"], p["msk.png"], p["cstr.txt"], p["out.flo"], p["wrgb.png"], p["wmsk.png"]],
                       env=dict(env, ARAP_SOLVER="nope"), capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "ARAP_SOLVER" in r.stderr
