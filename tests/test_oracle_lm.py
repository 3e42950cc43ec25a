"""The oracle's restatement of the reference's dormant "LMGPU" solver kind (solverGPUGaussNewton.t with UsesLambda():
:616-680, :956-1007, :1016-1177; o.t:2174-2202, :2255-2288).  The tree holds no vectors for it (the app never requests
it), so these are property checks of the restatement; the CUDA path is compared with it bit for bit in test_gpu_lm.py."""
import numpy as np

from tests.helpers import synth_gn_problem


def test_lm_reaches_the_gauss_newton_energy_with_far_fewer_linear_iterations(oracle):
    pr = synth_gn_problem(oracle, 64, 48, seed=3, fd=2)
    Xg, Ag, cg, _ = oracle.gn_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], 8, 400)
    Xl, Al, cl, st = oracle.lm_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], 20, 400)
    assert cl[0] == cg[0]                                   # same initial cost (same cost function)
    assert np.all(np.diff(cl) <= 0)                         # prevCost only moves on accepted steps
    assert abs(float(cl[-1]) - float(cg[-1])) <= 1e-4 * float(cg[-1])
    assert st[:, 1].sum() < 0.25 * 8 * 400                  # the Q test ends the linear loops early (:1093-1101)
    assert st[-1, 2] == 2.0                                 # "Function tolerance reached" (:1129-1133)
    assert np.abs(Xl - Xg).max() < 5e-2


def test_lm_model_cost_predicts_the_cost_of_a_small_step(oracle):
    """model cost = 0.5 |F + J delta|^2: for an accepted step with a tiny trust region (small delta) the nonlinear cost at
    the trial point is close to the model's prediction."""
    pr = synth_gn_problem(oracle, 48, 40, seed=5, fd=1)
    _, _, c, st = oracle.lm_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], 1, 50, trust_region_radius=1e-3)
    model, new = float(st[0, 3]), float(st[0, 4])
    assert st[0, 2] == 1.0 and new < float(c[0])
    assert abs(new - model) < 1e-3 * abs(float(c[0]) - new)


def test_lm_reverted_steps_restore_the_unknowns_exactly(oracle):
    pr = synth_gn_problem(oracle, 48, 40, seed=6, fd=2)
    # no step can reach a relative decrease of 2 (it is at most about 1): every step is reverted, the radius shrinks by
    # 2, 4, 8, ... (:1143-1155) until it passes the minimum
    X, A, c, st = oracle.lm_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], 10, 30, min_relative_decrease=2.0,
                                  min_trust_region_radius=100.0)
    assert np.array_equal(X, pr["X"]) and np.array_equal(A, pr["A"])
    assert np.all(c == c[0])
    assert list(st[:, 2]) == [0.0, 0.0, 0.0, 3.0]
    assert list(st[:, 0]) == [5000.0, 1250.0, 156.25, 9.765625]


def test_lm_residual_refresh_changes_nothing_in_exact_arithmetic(oracle):
    """r = b - (J^T J + CtC) delta every residual_reset_period iterations (:1077-1086) is the same residual up to rounding."""
    pr = synth_gn_problem(oracle, 48, 40, seed=7, fd=2)
    a = oracle.lm_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], 2, 60, residual_reset_period=3, q_tolerance=-1e30)
    b = oracle.lm_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], 2, 60, residual_reset_period=1000, q_tolerance=-1e30)
    assert list(a[3][:, 1]) == [60.0, 60.0] and list(b[3][:, 1]) == [60.0, 60.0]
    assert abs(float(a[2][-1]) - float(b[2][-1])) <= 1e-4 * float(b[2][-1])
    assert not np.array_equal(a[0], b[0])                   # but it is a different rounding path
