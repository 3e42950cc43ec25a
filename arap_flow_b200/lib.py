"""ctypes binding of libarapb200.so (include/Opt.h + include/arapb200.h).

The shared library is the product; this module only marshals numpy buffers into its C ABI.  There is no
fallback of any kind: if the library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ARAPB200_LIB selects another build of the same library (A/B measurements, tools/build_variant.sh)
LIB_PATH = os.environ.get("ARAPB200_LIB") or os.path.join(_HERE, "libarapb200.so")

BACKEND_AUTO, BACKEND_STREAM, BACKEND_RESIDENT = 0, 1, 2

# every symbol include/Opt.h and include/arapb200.h declare
OPT_SYMBOLS = [
    "Opt_NewState", "Opt_ProblemDefine", "Opt_ProblemDelete", "Opt_ProblemPlan", "Opt_PlanFree",
    "Opt_SetSolverParameter", "Opt_ProblemSolve", "Opt_ProblemInit", "Opt_ProblemStep", "Opt_ProblemCurrentCost",
]
ARAP_SYMBOLS = [
    "arapb200_device_info", "arapb200_version", "arapb200_warp", "arapb200_warp_flow", "arapb200_deform",
    "arapb200_batch_create", "arapb200_batch_destroy", "arapb200_batch_submit", "arapb200_batch_run",
    "arapb200_batch_timing", "arapb200_batch_launches", "arapb200_debug_gn_solve", "arapb200_debug_eval_jtf",
    "arapb200_debug_apply_jtj", "arapb200_debug_cost", "arapb200_debug_sincos", "arapb200_debug_exact_sum",
    "arapb200_debug_resident_profile", "arapb200_debug_resident_profile_group", "arapb200_flatten", "arapb200_filter_matches", "arapb200_segment_mask",
    "arapb200_batch_set_option", "arapb200_batch_resident_count", "arapb200_debug_wide_sum", "arapb200_plan_error", "arapb200_plan_lm_info", "arapb200_plan_timing_report",
    "arapb200_batch_launch_info",
]


class OptInitializationParameters(C.Structure):
    _fields_ = [("doublePrecision", C.c_int), ("verbosityLevel", C.c_int),
                ("collectPerKernelTimingInfo", C.c_int), ("threadsPerBlock", C.c_int)]


_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

_lib = None


def load():
    """Load the library (raises if it has not been built: run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with arap_flow_b200/csrc/Makefile (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    L.arapb200_version.restype = C.c_char_p
    L.arapb200_device_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.arapb200_warp.argtypes = [C.c_int, C.c_int, _f32p, _u8p, _u8p, _u8p, _u8p, C.c_void_p]
    L.arapb200_warp_flow.argtypes = [C.c_int, C.c_int, _f32p, _u8p, _u8p, _u8p, _u8p, C.c_void_p]
    L.arapb200_deform.argtypes = [C.c_int, C.c_int, _u8p, _u8p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  _f32p, _u8p, _u8p, C.c_void_p]
    L.arapb200_batch_create.argtypes = [C.c_int] * 7
    L.arapb200_batch_create.restype = C.c_void_p
    L.arapb200_batch_destroy.argtypes = [C.c_void_p]
    L.arapb200_batch_destroy.restype = None
    L.arapb200_batch_submit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.arapb200_batch_run.argtypes = [C.c_void_p]
    L.arapb200_batch_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.arapb200_batch_launches.argtypes = [C.c_void_p]
    L.arapb200_batch_launches.restype = C.c_longlong
    L.arapb200_debug_gn_solve.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float,
                                          C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.arapb200_debug_eval_jtf.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float,
                                          _f32p, _f32p]
    L.arapb200_debug_apply_jtj.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float,
                                           _f32p, _f32p, C.POINTER(C.c_float)]
    L.arapb200_debug_cost.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float,
                                      C.POINTER(C.c_float)]
    L.arapb200_debug_sincos.argtypes = [C.c_int, _f32p, _f32p, _f32p]
    L.arapb200_debug_exact_sum.argtypes = [C.c_size_t, _f32p, C.POINTER(C.c_float)]
    # Opt.h
    L.Opt_NewState.argtypes = [OptInitializationParameters]
    L.Opt_NewState.restype = C.c_void_p
    L.Opt_ProblemDefine.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.Opt_ProblemDefine.restype = C.c_void_p
    L.Opt_ProblemDelete.argtypes = [C.c_void_p, C.c_void_p]
    L.Opt_ProblemDelete.restype = None
    L.Opt_ProblemPlan.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint)]
    L.Opt_ProblemPlan.restype = C.c_void_p
    L.Opt_PlanFree.argtypes = [C.c_void_p, C.c_void_p]
    L.Opt_PlanFree.restype = None
    L.Opt_SetSolverParameter.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p]
    L.Opt_SetSolverParameter.restype = None
    for n in ("Opt_ProblemSolve", "Opt_ProblemInit"):
        getattr(L, n).argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        getattr(L, n).restype = None
    L.Opt_ProblemStep.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.Opt_ProblemStep.restype = C.c_int
    L.Opt_ProblemCurrentCost.argtypes = [C.c_void_p, C.c_void_p]
    L.Opt_ProblemCurrentCost.restype = C.c_double
    _lib = L
    return L


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed with code {rc}")


def device_info():
    sm, l2, ma, mi = C.c_int(), C.c_size_t(), C.c_int(), C.c_int()
    _check(load().arapb200_device_info(C.byref(sm), C.byref(l2), C.byref(ma), C.byref(mi)), "arapb200_device_info")
    return dict(sm_count=sm.value, l2_bytes=l2.value, cc=(ma.value, mi.value))


def warp(pos, rgb, mask_red, want_splat=True):
    """Forward warp from absolute positions (== CombinedSolver::copyResultToCPU).  Returns (rgb, mask, splat)."""
    H, W = mask_red.shape
    o_rgb = np.zeros((H, W, 3), np.uint8)
    o_m = np.zeros((H, W), np.uint8)
    sp = np.zeros((H, W), np.uint32) if want_splat else None
    _check(load().arapb200_warp(W, H, _c(pos, np.float32), _c(rgb, np.uint8), _c(mask_red, np.uint8), o_rgb, o_m,
                                sp.ctypes.data if want_splat else None), "arapb200_warp")
    return o_rgb, o_m, sp


def warp_flow(flow, rgb, mask_red, want_splat=True):
    """warp_image: forward warp from a flow field (ARAP/warping/src/main.cpp:145-225)."""
    H, W = mask_red.shape
    o_rgb = np.zeros((H, W, 3), np.uint8)
    o_m = np.zeros((H, W), np.uint8)
    sp = np.zeros((H, W), np.uint32) if want_splat else None
    _check(load().arapb200_warp_flow(W, H, _c(flow, np.float32), _c(rgb, np.uint8), _c(mask_red, np.uint8), o_rgb, o_m,
                                     sp.ctypes.data if want_splat else None), "arapb200_warp_flow")
    return o_rgb, o_m, sp


def deform(rgb, mask_red, matches, nCont=19, nGN=8, nPCG=400, backend=BACKEND_AUTO):
    """arap_deform for one image/segment: returns (flow[H,W,2], warped_rgb, warped_mask, costs[nCont,nGN+1])."""
    H, W = mask_red.shape
    m = _c(matches, np.int32).reshape(-1, 4)
    flow = np.zeros((H, W, 2), np.float32)
    o_rgb = np.zeros((H, W, 3), np.uint8)
    o_m = np.zeros((H, W), np.uint8)
    costs = np.zeros((nCont, nGN + 1), np.float32)
    _check(load().arapb200_deform(W, H, _c(rgb, np.uint8), _c(mask_red, np.uint8), m, len(m), nCont, nGN, nPCG,
                                  backend, flow, o_rgb, o_m, costs.ctypes.data), "arapb200_deform")
    return flow, o_rgb, o_m, costs


def flatten(flows, rgbs, masks, background=None):
    """Layer the per-segment results of a --multseg pair and composite the background (para_gen.py:136-175, 50-61)."""
    n = len(flows)
    H, W = masks[0].shape
    fl = [_c(f, np.float32) for f in flows]
    rg = [_c(r, np.uint8) for r in rgbs]
    mk = [_c(m, np.uint8) for m in masks]
    P = C.c_void_p * n
    bg = _c(background, np.uint8) if background is not None else None
    o_f = np.zeros((H, W, 2), np.float32)
    o_r = np.zeros((H, W, 3), np.uint8)
    o_m = np.zeros((H, W), np.uint8)
    L = load()
    L.arapb200_flatten.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    _check(L.arapb200_flatten(W, H, n, P(*[a.ctypes.data for a in fl]), P(*[a.ctypes.data for a in rg]),
                              P(*[a.ctypes.data for a in mk]), bg.ctypes.data if bg is not None else None,
                              o_f.ctypes.data, o_r.ctypes.data, o_m.ctypes.data), "arapb200_flatten")
    return o_f, o_r, o_m


def filter_matches(matches, labels1, labels2):
    """N3: keep the matches para_gen.valid_cnstr keeps (para_gen.py:216-223, 468-482), in order.
    Returns (matches int32[k,4], labels uint8[k])."""
    m = _c(np.asarray(matches, np.int32).reshape(-1, 4), np.int32)
    l1, l2 = _c(labels1, np.uint8), _c(labels2, np.uint8)
    n = m.shape[0]
    out = np.zeros((max(n, 1), 4), np.int32)
    lab = np.zeros(max(n, 1), np.uint8)
    cnt = C.c_int(0)
    L = load()
    L.arapb200_filter_matches.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    _check(L.arapb200_filter_matches(l1.shape[1], l1.shape[0], l1.ctypes.data, l2.shape[1], l2.shape[0], l2.ctypes.data,
                                     m.ctypes.data, n, out.ctypes.data, lab.ctypes.data, C.byref(cnt)),
           "arapb200_filter_matches")
    return out[:cnt.value].copy(), lab[:cnt.value].copy()


def segment_mask(labels, segment=0):
    """N3: the mask image arap_deform expects (0 = solve, 255 = ARAP_BG): one label (--multseg) or all (segment=0)."""
    l = _c(labels, np.uint8)
    out = np.zeros(l.shape, np.uint8)
    L = load()
    L.arapb200_segment_mask.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    _check(L.arapb200_segment_mask(l.shape[1], l.shape[0], l.ctypes.data, int(segment), out.ctypes.data),
           "arapb200_segment_mask")
    return out


class Batch:
    """Many independent (image, segment) problems on the current device (arapb200_batch_*)."""

    def __init__(self, maxW, maxH, max_problems, nCont=19, nGN=8, nPCG=400, backend=BACKEND_AUTO):
        self.L = load()
        self.nCont, self.nGN = nCont, nGN
        self.h = self.L.arapb200_batch_create(maxW, maxH, max_problems, nCont, nGN, nPCG, backend)
        if not self.h:
            raise RuntimeError("arapb200_batch_create failed")
        self._keep = {}

    def set_option(self, name: str, value: float):
        """Opt-in behaviour beyond the reference (include/arapb200.h): e.g. set_option("pcg_rtol", 1e-3)."""
        self.L.arapb200_batch_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        _check(self.L.arapb200_batch_set_option(self.h, name.encode(), float(value)), "arapb200_batch_set_option")

    def submit(self, slot, rgb, mask_red, matches, out=None):
        """Queue one problem.  `out` may be the dict a previous submit() of the same image size returned: its arrays are
        then written in place instead of allocating fresh ones (what a caller looping over many pairs does)."""
        H, W = mask_red.shape
        rgb = _c(rgb, np.uint8)
        mask_red = _c(mask_red, np.uint8)
        m = _c(matches, np.int32).reshape(-1, 4)
        if out is None or out["flow"].shape != (H, W, 2):
            out = dict(flow=np.zeros((H, W, 2), np.float32), rgb=np.zeros((H, W, 3), np.uint8),
                       mask=np.zeros((H, W), np.uint8), costs=np.zeros((self.nCont, self.nGN + 1), np.float32))
        self._keep[slot] = (rgb, mask_red, m, out)
        _check(self.L.arapb200_batch_submit(self.h, slot, W, H, rgb.ctypes.data, mask_red.ctypes.data, m.ctypes.data,
                                            len(m), out["flow"].ctypes.data, out["rgb"].ctypes.data,
                                            out["mask"].ctypes.data, out["costs"].ctypes.data), "arapb200_batch_submit")
        return out

    def run(self):
        _check(self.L.arapb200_batch_run(self.h), "arapb200_batch_run")

    def timing_ms(self):
        ms = (C.c_float * 3)()
        self.L.arapb200_batch_timing(self.h, ms)
        return dict(total=ms[0], solve=ms[1], warp=ms[2])

    def launches(self):
        return int(self.L.arapb200_batch_launches(self.h))

    def resident_count(self):
        """Problems of the last run() that the resident back-end solved (the others streamed)."""
        self.L.arapb200_batch_resident_count.argtypes = [C.c_void_p]
        return int(self.L.arapb200_batch_resident_count(self.h))

    def launch_info(self):
        """Shape of the last cooperative launch of the last run() (zeros if every problem streamed)."""
        info = (C.c_int * 6)()
        self.L.arapb200_batch_launch_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        _check(self.L.arapb200_batch_launch_info(self.h, info), "arapb200_batch_launch_info")
        return dict(problems_per_launch=info[0], variant=(info[1], info[2]), grid=(info[3], info[4]), threads=info[5])

    def close(self):
        if self.h:
            self.L.arapb200_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- unit-level parity entry points -----------------------------------------------------------------
def debug_gn_solve(X, A, U, Cn, M, nGN, nPCG, wf, wr, backend=BACKEND_AUTO, trace=False):
    H, W = M.shape
    X = _c(X, np.float32).copy()
    A = _c(A, np.float32).copy()
    costs = np.zeros(nGN + 1, np.float32)
    scal = np.zeros((nGN, nPCG, 3), np.float32) if trace else None
    _check(load().arapb200_debug_gn_solve(W, H, X, A, _c(U, np.float32), _c(Cn, np.float32), _c(M, np.float32),
                                          wf, wr, nGN, nPCG, backend, costs.ctypes.data,
                                          scal.ctypes.data if trace else None), "arapb200_debug_gn_solve")
    return X, A, costs, scal


def debug_eval_jtf(X, A, U, Cn, M, wf, wr):
    H, W = M.shape
    r = np.zeros((H, W, 3), np.float32)
    pre = np.zeros((H, W, 3), np.float32)
    _check(load().arapb200_debug_eval_jtf(W, H, _c(X, np.float32), _c(A, np.float32), _c(U, np.float32),
                                          _c(Cn, np.float32), _c(M, np.float32), wf, wr, r, pre), "debug_eval_jtf")
    return r, pre


def debug_apply_jtj(A, U, Cn, M, p, wf, wr):
    H, W = M.shape
    q = np.zeros((H, W, 3), np.float32)
    d = C.c_float()
    _check(load().arapb200_debug_apply_jtj(W, H, _c(A, np.float32), _c(U, np.float32), _c(Cn, np.float32),
                                           _c(M, np.float32), wf, wr, _c(p, np.float32), q, C.byref(d)), "debug_apply_jtj")
    return q, np.float32(d.value)


def debug_cost(X, A, U, Cn, M, wf, wr):
    H, W = M.shape
    c = C.c_float()
    _check(load().arapb200_debug_cost(W, H, _c(X, np.float32), _c(A, np.float32), _c(U, np.float32),
                                      _c(Cn, np.float32), _c(M, np.float32), wf, wr, C.byref(c)), "debug_cost")
    return np.float32(c.value)


def debug_resident_profile(mask_red, matches, nCont, nGN, nPCG, copies=1):
    """Cycle accounting of the resident kernel (thread 0 of every CTA) for one problem, or for the first of `copies`
    identical problems sharing one cooperative launch.  Returns (prof[G, 16], info, ms)."""
    H, W = mask_red.shape
    m = _c(matches, np.int32).reshape(-1, 4)
    prof = np.zeros((160, 16), np.uint64)
    info = (C.c_int * 3)()
    ms = C.c_float()
    L = load()
    L.arapb200_debug_resident_profile.argtypes = [C.c_int, C.c_int, _u8p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    L.arapb200_debug_resident_profile_group.argtypes = [C.c_int, C.c_int, _u8p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                        C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    if copies > 1:
        _check(L.arapb200_debug_resident_profile_group(W, H, _c(mask_red, np.uint8), m, len(m), copies, nCont, nGN, nPCG,
                                                       prof.ctypes.data, info, C.byref(ms)), "debug_resident_profile_group")
    else:
        _check(L.arapb200_debug_resident_profile(W, H, _c(mask_red, np.uint8), m, len(m), nCont, nGN, nPCG,
                                                 prof.ctypes.data, info, C.byref(ms)), "debug_resident_profile")
    G = info[1]
    return prof[:G], dict(strips=info[0], ctas=info[1], warps=info[2]), ms.value


def debug_sincos(a):
    a = _c(a, np.float32).ravel()
    s = np.zeros_like(a)
    c = np.zeros_like(a)
    _check(load().arapb200_debug_sincos(a.size, a, s, c), "debug_sincos")
    return s, c


def debug_wide_sum(t):
    t = _c(t, np.float32).ravel()
    out = C.c_float()
    L = load()
    L.arapb200_debug_wide_sum.argtypes = [C.c_size_t, _f32p, C.POINTER(C.c_float)]
    _check(L.arapb200_debug_wide_sum(t.size, t, C.byref(out)), "debug_wide_sum")
    return np.float32(out.value)


def debug_exact_sum(t):
    t = _c(t, np.float32).ravel()
    out = C.c_float()
    _check(load().arapb200_debug_exact_sum(t.size, t, C.byref(out)), "debug_exact_sum")
    return np.float32(out.value)
