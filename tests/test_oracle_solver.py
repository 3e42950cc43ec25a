"""Oracle self-checks: derivatives against finite differences of the residual definition, exact sums,
contract sincos, constraint image and whole-solve properties -- CPU only."""
import math

import numpy as np

from tests.helpers import random_problem
from arap_flow_b200 import synth


def _fd_jacobian(oracle, pr, eps=1e-6):
    H, W = pr["M"].shape
    Xd, Ad = pr["X"].astype(np.float64), pr["A"].astype(np.float64)

    def F(v):
        return oracle.residuals_f64(v[:2 * W * H].reshape(H, W, 2), v[2 * W * H:].reshape(H, W), pr["U"], pr["C"],
                                    pr["M"]).ravel()
    v0 = np.concatenate([Xd.ravel(), Ad.ravel()])
    f0 = F(v0)
    J = np.zeros((f0.size, v0.size))
    for k in range(v0.size):
        vp, vm = v0.copy(), v0.copy()
        vp[k] += eps
        vm[k] -= eps
        J[:, k] = (F(vp) - F(vm)) / (2 * eps)
    return J, f0


def test_derivatives_against_finite_differences(oracle):
    pr = random_problem(9, 7, seed=3)
    H, W = pr["M"].shape
    J, f0 = _fd_jacobian(oracle, pr)
    act = pr["M"] == 0
    un = lambda v: (v[:2 * W * H].reshape(H, W, 2), v[2 * W * H:].reshape(H, W))
    r, pre = oracle.eval_jtf(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"])
    gX, gA = un(J.T @ f0)
    scale = np.abs(J.T @ f0).max()
    assert np.abs(-r[..., :2][act] - gX[act]).max() < 2e-6 * scale
    assert np.abs(-r[..., 2][act] - gA[act]).max() < 2e-6 * scale
    DX, DA = un(np.einsum("ij,ij->j", J, J))
    assert np.abs(pre[..., 0][act] - 1 / (1 + np.sqrt(DX[..., 0][act])) ** 2).max() < 1e-6
    assert np.abs(pre[..., 2][act] - 1 / (1 + np.sqrt(DA[act])) ** 2).max() < 1e-6
    pv = np.concatenate([pr["p"][..., :2].ravel(), pr["p"][..., 2].ravel()]).astype(np.float64)
    qX, qA = un(J.T @ (J @ pv))
    q, d = oracle.apply_jtj(pr["A"], pr["U"], pr["C"], pr["M"], pr["p"])
    qs = np.abs(J.T @ (J @ pv)).max()
    assert np.abs(q[..., :2][act] - qX[act]).max() < 2e-6 * qs
    assert np.abs(q[..., 2][act] - qA[act]).max() < 2e-6 * qs
    assert abs(float(d) - float(pv @ (J.T @ (J @ pv)))) < 1e-5 * abs(float(d))
    c = oracle.cost(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"])
    assert abs(float(c) - 0.5 * float(f0 @ f0)) < 1e-5 * float(c)


def test_exact_sum_is_exact(oracle):
    rng = np.random.default_rng(0)
    for n in (1, 7, 1000, 100000):
        t = (rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 8, n)).astype(np.float32)
        want = np.float32(math.fsum(t.astype(np.float64)))
        assert oracle.exact_sum(t) == want
        assert oracle.exact_sum(t[::-1].copy()) == want          # order independent
    assert oracle.exact_sum(np.zeros(5, np.float32)) == 0.0


def test_contract_sincos(oracle):
    a = np.concatenate([np.linspace(-12, 12, 4001), [0.0, 1e-8, -1e-8, 100.5, -777.25]]).astype(np.float32)
    for x in a:
        s, c = oracle.sincos(float(x))
        assert abs(float(s) - math.sin(float(x))) <= 6.1e-8
        assert abs(float(c) - math.cos(float(x))) <= 6.1e-8
    assert oracle.sincos(0.0) == (np.float32(0.0), np.float32(1.0))


def test_constraint_image_rules(oracle):
    """CombinedSolver.h:223-242: only where mask(src)==0, later entries override, lerp by alpha; pins last."""
    W, H = 8, 6
    mask = np.full((H, W), 255, np.uint8)
    mask[1:5, 1:7] = 0
    mask[0, 3] = 0  # an object pixel on the border gets pinned
    m = np.array([[2, 2, 6, 2], [2, 2, 4, 4], [7, 5, 0, 0], [0, 3, 5, 5]], np.int32)
    allm = oracle.with_border_pins(m, W, H)
    assert len(allm) == 4 + oracle.lib().arap_oracle_border_pin_count(W, H)
    Cn = oracle.constraint_image(mask, allm, 0.5)
    assert tuple(Cn[2, 2]) == (3.0, 3.0)          # the later duplicate wins
    assert tuple(Cn[5, 7]) == (-1.0, -1.0)        # source off-object: dropped
    assert tuple(Cn[0, 3]) == (3.0, 0.0)          # border pin
    assert (Cn[mask != 0] == -1).all()


def test_zero_constraints_zero_flow_and_translation(oracle):
    sp = synth.synth(48, 40, 1, 1, 5)
    mask = sp.masks[0]
    X, A, costs = oracle.solve(mask, np.zeros((0, 4), np.int32), nCont=2, nGN=2, nPCG=20)
    assert np.array_equal(oracle.flow(X), np.zeros((40, 48, 2), np.float32)) and costs.max() == 0.0
    # a pure translation of every lattice point is matched rigidly with ~zero energy
    src = np.argwhere(mask == 0)[::7]
    m = np.array([[x, y, x + 3, y - 2] for y, x in src], np.int32)
    X, A, costs = oracle.solve(mask, m, nCont=4, nGN=3, nPCG=60)
    fl = oracle.flow(X)
    act = mask == 0
    assert np.abs(fl[act] - np.array([3.0, -2.0], np.float32)).max() < 2e-2
    assert np.abs(A[act]).max() < 1e-2 and costs[-1, -1] < 1e-3
    assert (fl[~act] == 0).all()


def test_cost_decreases_within_a_solve(oracle):
    from tests.helpers import synth_gn_problem
    pr = synth_gn_problem(oracle, 64, 48, seed=9, fd=2)
    X, A, costs, scal = oracle.gn_solve(pr["X"], pr["A"], pr["U"], pr["C"], pr["M"], 4, 50, trace=True)
    assert costs[-1] < costs[0] and scal.shape == (4, 50, 3) and (scal[..., 0] > 0).all()
