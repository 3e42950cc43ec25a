"""Opt_InitializationParameters.collectPerKernelTimingInfo / verbosityLevel (ARAP/API/release/include/Opt.h:16-28): the
reference's per-kernel timer (ARAP/API/src/util.t:404-510) -- one event pair per launch under the reference's kernel
names, aggregated and printed at the end of a solve -- for both solver kinds; and the results do not change."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from arap_flow_b200 import lib
from tests.helpers import synth_gn_problem

pytestmark = pytest.mark.gpu
PLAN = os.path.join(os.path.dirname(lib.LIB_PATH), "arap_plan.t")


def _solve(oracle, pr, kind, nGN, nPCG, verbosity, collect):
    import torch
    L = lib.load()
    L.arapb200_plan_timing_report.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.arapb200_plan_timing_report.restype = C.c_size_t
    st = L.Opt_NewState(lib.OptInitializationParameters(0, verbosity, collect, 0))
    prob = L.Opt_ProblemDefine(st, PLAN.encode(), kind)
    plan = L.Opt_ProblemPlan(st, prob, (C.c_uint * 2)(pr["W"], pr["H"]))
    assert st and prob and plan
    a, b = C.c_uint(nGN), C.c_uint(nPCG)
    L.Opt_SetSolverParameter(st, plan, b"nIterations", C.byref(a))
    L.Opt_SetSolverParameter(st, plan, b"lIterations", C.byref(b))
    dev = torch.device("cuda:0")
    d = {k: torch.from_numpy(np.ascontiguousarray(pr[k])).to(dev) for k in ("X", "A", "U", "C", "M")}
    wf, wr = C.c_float(float(oracle.WF)), C.c_float(float(oracle.WR))
    pp = (C.c_void_p * 7)(d["X"].data_ptr(), d["A"].data_ptr(), d["U"].data_ptr(), d["C"].data_ptr(), d["M"].data_ptr(),
                          C.cast(C.byref(wf), C.c_void_p), C.cast(C.byref(wr), C.c_void_p))
    L.Opt_ProblemSolve(st, plan, pp)
    cost = np.float32(L.Opt_ProblemCurrentCost(st, plan))
    n = L.arapb200_plan_timing_report(plan, None, 0)
    buf = C.create_string_buffer(n + 1)
    L.arapb200_plan_timing_report(plan, buf, n + 1)
    X = d["X"].cpu().numpy()
    L.Opt_PlanFree(st, plan)
    L.Opt_ProblemDelete(st, prob)
    return cost, X, buf.value.decode()


def _rows(report):
    rows = {}
    for m in re.finditer(r"^ (\S+)\s+\|\s+(\d+)\s+\|\s+([0-9.]+)ms\|\s+([0-9.]+)ms$", report, re.M):
        rows[m.group(1)] = (int(m.group(2)), float(m.group(3)), float(m.group(4)))
    return rows


def test_gauss_newton_kernel_timing_table(oracle, capfd):
    pr = synth_gn_problem(oracle, 200, 160, seed=4, fd=2)
    nGN, nPCG = 3, 25
    c0, X0, rep0 = _solve(oracle, pr, b"gaussNewtonGPU", nGN, nPCG, 0, 0)
    assert rep0 == ""                                       # nothing is collected by default
    capfd.readouterr()
    c1, X1, rep = _solve(oracle, pr, b"gaussNewtonGPU", nGN, nPCG, 0, 1)
    assert capfd.readouterr().out == ""                     # verbosity 0: collected, not printed (util.t:452)
    assert c1 == c0 and np.array_equal(X1, X0)              # the timed (streaming, eager) path gives the same bits
    rows = _rows(rep)
    assert rows["overall"][0] == 1
    assert rows["PCGInit1"][0] == nGN and rows["PCGLinearUpdate"][0] == nGN
    assert rows["PCGStep1"][0] == nGN * nPCG and rows["PCGStep2"][0] == nGN * nPCG
    assert rows["computeCost"][0] == nGN + 1
    assert all(v[1] > 0 for v in rows.values())
    assert sum(v[1] for k, v in rows.items() if k != "overall") <= rows["overall"][1] * 1.001
    assert "TIMING " in rep and "Per-iter times ms (nonlinear,linear):" in rep
    # verbosity > 0: the reference prints the per-step costs, "final cost" and the table
    c2, X2, rep2 = _solve(oracle, pr, b"gaussNewtonGPU", nGN, nPCG, 1, 1)
    out = capfd.readouterr().out
    assert c2 == c0 and out.count("cost: ") == nGN and "final cost=" in out
    assert "        Kernel        |   Count  |   Total   | Average " in out and " PCGStep1 " in out
    # verbosity only: just the "overall" row
    c3, X3, rep3 = _solve(oracle, pr, b"gaussNewtonGPU", nGN, nPCG, 1, 0)
    capfd.readouterr()
    assert list(_rows(rep3)) == ["overall"] and c3 == c0


def test_lm_kernel_timing_table(oracle, capfd):
    pr = synth_gn_problem(oracle, 160, 120, seed=7, fd=3)
    Xo, Ao, co, so = oracle.lm_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], 4, 64, residual_reset_period=10)
    c, X, rep = _solve(oracle, pr, b"LMGPU", 4, 64, 0, 1)
    assert c == co[len(so)] and np.array_equal(X, Xo)
    rows = _rows(rep)
    its = int(so[:, 1].sum())
    assert rows["PCGInit1"][0] == len(so) and rows["computeModelCost"][0] == len(so)
    assert rows["PCGStep1"][0] >= its                        # launches after the Q test fired are no-ops, but are launched
    assert rows["PCGStep2"][0] + rows.get("PCGStep2_2ndHalf", (0,))[0] == rows["PCGStep1"][0]
    assert rows["PCGStep3"][0] == rows["PCGStep1"][0]
    assert rows.get("computeAdelta", (0,))[0] == rows.get("PCGStep2_1stHalf", (0,))[0]
