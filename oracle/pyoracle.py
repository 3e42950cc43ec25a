"""ctypes loader for the CPU oracle (oracle/arap_oracle.c).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from arap_flow_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
REF_WARP_BIN = os.path.join(_HERE, "_ref", "warp_image_ref")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    src = os.path.join(_HERE, "arap_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(src) > os.path.getmtime(_LIB):
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.arap_oracle_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.arap_oracle_exact_sum.argtypes = [_f32p, C.c_size_t]
        L.arap_oracle_exact_sum.restype = C.c_float
        L.arap_oracle_cost.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float]
        L.arap_oracle_cost.restype = C.c_float
        L.arap_oracle_eval_jtf.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float,
                                           C.c_float, _f32p, _f32p]
        L.arap_oracle_apply_jtj.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float,
                                            _f32p, _f32p]
        L.arap_oracle_apply_jtj.restype = C.c_float
        L.arap_oracle_residuals_f64.argtypes = [C.c_int, C.c_int, _f64p, _f64p, _f32p, _f32p, _f32p, C.c_double,
                                                C.c_double, _f64p]
        L.arap_oracle_gn_solve.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float,
                                           C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.arap_oracle_gn_solve.restype = C.c_int
        L.arap_oracle_lm_solve.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float,
                                           C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.arap_oracle_lm_solve.restype = C.c_int
        L.arap_oracle_constraint_image.argtypes = [C.c_int, C.c_int, _u8p, _i32p, C.c_int, C.c_float, _f32p]
        L.arap_oracle_border_pin_count.argtypes = [C.c_int, C.c_int]
        L.arap_oracle_border_pin_count.restype = C.c_int
        L.arap_oracle_solve.argtypes = [C.c_int, C.c_int, _u8p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        _f32p, _f32p, C.c_void_p]
        L.arap_oracle_solve.restype = C.c_int
        L.arap_oracle_flow.argtypes = [C.c_int, C.c_int, _f32p, _f32p]
        L.arap_oracle_warp.argtypes = [C.c_int, C.c_int, _f32p, _u8p, _u8p, _u8p, _u8p, _u32p]
        L.arap_oracle_flow_to_pos.argtypes = [C.c_int, C.c_int, _f32p, _f32p]
        L.arap_oracle_num_threads.restype = C.c_int
        L.arap_oracle_set_rtol.argtypes = [C.c_float, C.c_float]
        L.arap_oracle_set_rtol.restype = None
        _lib = L
    return _lib


WF = np.float32(np.sqrt(np.float32(100.0)))
WR = np.float32(np.sqrt(np.float32(0.01)))


def sincos(a: float):
    s, c = C.c_float(), C.c_float()
    lib().arap_oracle_sincos(C.c_float(a), C.byref(s), C.byref(c))
    return np.float32(s.value), np.float32(c.value)


def exact_sum(t: np.ndarray) -> np.float32:
    t = np.ascontiguousarray(t, dtype=np.float32).ravel()
    return np.float32(lib().arap_oracle_exact_sum(t, t.size))


def grid(W: int, H: int) -> np.ndarray:
    yy, xx = np.mgrid[0:H, 0:W]
    return np.ascontiguousarray(np.stack([xx, yy], axis=-1).astype(np.float32))


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


def cost(X, A, U, Cn, M, wf=WF, wr=WR) -> np.float32:
    H, W = M.shape
    return np.float32(lib().arap_oracle_cost(W, H, _c(X), _c(A), _c(U), _c(Cn), _c(M), wf, wr))


def eval_jtf(X, A, U, Cn, M, wf=WF, wr=WR):
    H, W = M.shape
    r = np.zeros((H, W, 3), np.float32)
    pre = np.zeros((H, W, 3), np.float32)
    lib().arap_oracle_eval_jtf(W, H, _c(X), _c(A), _c(U), _c(Cn), _c(M), wf, wr, r, pre)
    return r, pre


def apply_jtj(A, U, Cn, M, p, wf=WF, wr=WR):
    H, W = M.shape
    q = np.zeros((H, W, 3), np.float32)
    d = lib().arap_oracle_apply_jtj(W, H, _c(A), _c(U), _c(Cn), _c(M), wf, wr, _c(p), q)
    return q, np.float32(d)


def residuals_f64(X, A, U, Cn, M, wf=float(WF), wr=float(WR)):
    H, W = M.shape
    out = np.zeros((H, W, 10), np.float64)
    lib().arap_oracle_residuals_f64(W, H, _c(X, np.float64), _c(A, np.float64), _c(U), _c(Cn), _c(M), wf, wr, out)
    return out


def gn_solve(X, A, U, Cn, M, nGN, nPCG, wf=WF, wr=WR, trace=False):
    """One Opt_ProblemSolve.  Returns (X, A, costs[nGN+1], scal[nGN, nPCG, 3] or None); inputs untouched."""
    H, W = M.shape
    X = _c(X).copy()
    A = _c(A).copy()
    costs = np.zeros(nGN + 1, np.float32)
    scal = np.zeros((nGN, nPCG, 3), np.float32) if trace else None
    rc = lib().arap_oracle_gn_solve(W, H, X, A, _c(U), _c(Cn), _c(M), wf, wr, nGN, nPCG,
                                    costs.ctypes.data, scal.ctypes.data if trace else None)
    assert rc == 0
    return X, A, costs, scal


# solverGPUGaussNewton.t:26-39, in the order arap_oracle_lm_solve takes them
LM_DEFAULTS = {"min_relative_decrease": 1e-3, "min_trust_region_radius": 1e-32, "max_trust_region_radius": 1e16,
               "q_tolerance": 0.0001, "function_tolerance": 0.000001, "trust_region_radius": 1e4,
               "radius_decrease_factor": 2.0, "min_lm_diagonal": 1e-6, "max_lm_diagonal": 1e32,
               "residual_reset_period": 10}


def lm_solve(X, A, U, Cn, M, nGN, nPCG, wf=WF, wr=WR, **params):
    """One Opt_ProblemSolve on an "LMGPU" plan (the reference's dormant Levenberg-Marquardt kind).  Returns
    (X, A, costs[nGN+1], stats[steps, 6]); stats rows = (radius after, PCG iterations, verdict 1 accepted / 0 reverted /
    2 function tolerance / 3 minimum radius, model cost, new cost, last Q).  Inputs untouched."""
    H, W = M.shape
    X = _c(X).copy()
    A = _c(A).copy()
    pv = dict(LM_DEFAULTS)
    for k, v in params.items():
        assert k in pv, k
        pv[k] = v
    p10 = np.array([pv[k] for k in LM_DEFAULTS], np.float32)
    costs = np.zeros(nGN + 1, np.float32)
    stats = np.zeros((max(nGN, 1), 6), np.float32)
    n = lib().arap_oracle_lm_solve(W, H, X, A, _c(U), _c(Cn), _c(M), wf, wr, nGN, nPCG, p10.ctypes.data,
                                   costs.ctypes.data, stats.ctypes.data)
    assert n >= 0
    return X, A, costs, stats[:n]


def constraint_image(mask_red, matches, alpha):
    H, W = mask_red.shape
    Cn = np.zeros((H, W, 2), np.float32)
    m = _c(matches, np.int32).reshape(-1, 4)
    lib().arap_oracle_constraint_image(W, H, _c(mask_red, np.uint8), m, len(m), np.float32(alpha), Cn)
    return Cn


def with_border_pins(matches, W, H):
    """matches + the pins arap_deform appends (ARAP/deformation/src/main.cpp:130-136), same order."""
    pins = [(x, y, x, y) for y in range(H) for x in range(W) if y == 0 or x == 0 or y == H - 1 or x == W - 1]
    m = np.asarray(matches, np.int32).reshape(-1, 4)
    return np.concatenate([m, np.asarray(pins, np.int32).reshape(-1, 4)], axis=0)


def solve(mask_red, matches, nCont=19, nGN=8, nPCG=400):
    """Whole arap_deform solve for one image/segment.  Returns (X[H,W,2], A[H,W], costs[nCont, nGN+1])."""
    H, W = mask_red.shape
    X = np.zeros((H, W, 2), np.float32)
    A = np.zeros((H, W), np.float32)
    costs = np.zeros((nCont, nGN + 1), np.float32)
    m = _c(matches, np.int32).reshape(-1, 4)
    rc = lib().arap_oracle_solve(W, H, _c(mask_red, np.uint8), m, len(m), nCont, nGN, nPCG, X, A,
                                 costs.ctypes.data)
    assert rc == 0
    return X, A, costs


def solve_lm(mask_red, matches, nCont=19, nGN=8, nPCG=400, **params):
    """The arap_deform schedule with every Opt_ProblemSolve run by the "LMGPU" solver kind (default parameters unless
    given): resetGPU, then per continuation step the constraint image and one lm_solve.  Returns
    (X[H,W,2], A[H,W], costs[nCont, nGN+1], linear iterations per continuation step)."""
    H, W = mask_red.shape
    m = with_border_pins(matches, W, H)
    U = grid(W, H)
    M = _c(mask_red, np.uint8).astype(np.float32)
    X, A = U.copy(), np.zeros((H, W), np.float32)
    costs = np.zeros((nCont, nGN + 1), np.float32)
    its = []
    for t in range(nCont):
        alpha = np.float32(t + 1) / np.float32(nCont)
        X, A, costs[t], st = lm_solve(X, A, U, constraint_image(mask_red, m, alpha), M, nGN, nPCG, **params)
        its.append(int(st[:, 1].sum()))
    return X, A, costs, its


def set_rtol(pcg_rtol=0.0, gn_rtol=0.0):
    """Opt-in early exits (N4), mirrored from the resident kernel; (0, 0) restores the reference's fixed budget."""
    lib().arap_oracle_set_rtol(np.float32(pcg_rtol), np.float32(gn_rtol))


def flow(X):
    H, W, _ = X.shape
    out = np.zeros_like(X, dtype=np.float32)
    lib().arap_oracle_flow(W, H, _c(X), out)
    return out


def flow_to_pos(fl):
    H, W, _ = fl.shape
    out = np.zeros((H, W, 2), np.float32)
    lib().arap_oracle_flow_to_pos(W, H, _c(fl), out)
    return out


def warp(pos, rgb, mask_red):
    """Returns (warped_rgb[H,W,3] u8, warped_mask[H,W] u8 0/255, splat[H,W] u32)."""
    H, W = mask_red.shape
    o_rgb = np.zeros((H, W, 3), np.uint8)
    o_m = np.zeros((H, W), np.uint8)
    sp = np.zeros((H, W), np.uint32)
    lib().arap_oracle_warp(W, H, _c(pos), _c(rgb, np.uint8), _c(mask_red, np.uint8), o_rgb, o_m, sp)
    return o_rgb, o_m, sp


def num_threads() -> int:
    return int(lib().arap_oracle_num_threads())


# ---- the independently-ordered "literal" implementation (oracle/arap_literal.c) ---------------------------------
_LIT = os.path.join(_HERE, "libliteral.so")
LIT_SINCOS_EVERY_ITER, LIT_ORDERED_ATOMICS = 1, 2
_lit = None


def literal_lib():
    global _lit
    if _lit is None:
        src = os.path.join(_HERE, "arap_literal.c")
        if not os.path.exists(_LIT) or os.path.getmtime(src) > os.path.getmtime(_LIT):
            subprocess.check_call(["make", "-C", _HERE, "--no-print-directory", "libliteral.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(_LIT)
        L.arap_literal_solve.argtypes = [C.c_int, C.c_int, _u8p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_ulonglong, C.c_int, _f32p, _f32p, C.c_void_p]
        L.arap_literal_solve.restype = C.c_int
        _lit = L
    return _lit


def literal_solve(mask_red, matches, nCont=19, nGN=8, nPCG=400, seed=0, flags=0):
    """Same interface as solve(); reference-like unfused schedule, libm sin/cos, compiler-chosen contraction, fp32
    per-warp partial sums accumulated in a seeded shuffled order (arap_literal.c header)."""
    H, W = mask_red.shape
    X = np.zeros((H, W, 2), np.float32)
    A = np.zeros((H, W), np.float32)
    costs = np.zeros((nCont, nGN + 1), np.float32)
    m = _c(matches, np.int32).reshape(-1, 4)
    rc = literal_lib().arap_literal_solve(W, H, _c(mask_red, np.uint8), m, len(m), nCont, nGN, nPCG, seed, flags, X, A,
                                          costs.ctypes.data)
    assert rc == 0
    return X, A, costs
