"""Host-side mirror of the reference's solve orchestration over the Opt.h C ABI.

`OptSolver` = ARAP/shared/OptSolver.h:43-91 (RAII over Opt_NewState / ProblemDefine / ProblemPlan, `solve` =
setAllSolverParameters + Opt_ProblemSolve + Opt_ProblemCurrentCost).  `CombinedSolver` = ARAP/deformation/src/
CombinedSolver.h:99-390 + ARAP/shared/CombinedSolverBase.h:99-120: it owns the five device images, uploads
mask / UrShape, rebuilds and uploads the constraint image before each of the `numIter` continuation steps
(`setConstraintImage`, :223-242), calls the solver once per step, downloads the warp field and rasterises it.

This is the path a caller of the reference's OWN library API takes: 19 synchronous `Opt_ProblemSolve` calls per image,
one problem at a time, with the reference's host-side constraint lerp and full-image upload per continuation step.
(The batched path, `arap_flow_b200.lib.Batch`, does all of that inside one kernel launch for several problems.)
The rasteriser is the library's GPU forward warp (`arapb200_warp`), not the reference's CPU loop.

torch provides device memory only; every computation is libarapb200's.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import lib

PLAN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "arap_plan.t")


def with_border_pins(matches, W, H):
    """loadData (ARAP/deformation/src/main.cpp:130-136): append a pin (x, y, x, y) for every border pixel, row-major."""
    m = np.asarray(matches, np.int32).reshape(-1, 4)
    ys, xs = np.mgrid[0:H, 0:W]
    b = (ys == 0) | (xs == 0) | (ys == H - 1) | (xs == W - 1)
    pins = np.stack([xs[b], ys[b], xs[b], ys[b]], axis=1).astype(np.int32)
    return np.concatenate([m, pins], axis=0)


class OptSolver:
    """ARAP/shared/OptSolver.h:43-91"""

    def __init__(self, W: int, H: int, plan_file: str = PLAN, solver_kind: str = "gaussNewtonGPU"):
        self.L = lib.load()
        self.state = self.L.Opt_NewState(lib.OptInitializationParameters(0, 0, 0, 0))
        self.problem = self.L.Opt_ProblemDefine(self.state, plan_file.encode(), solver_kind.encode())
        dims = (C.c_uint * 2)(W, H)
        self.plan = self.L.Opt_ProblemPlan(self.state, self.problem, dims) if self.problem else None
        assert self.state and self.problem and self.plan, "Opt_NewState / Opt_ProblemDefine / Opt_ProblemPlan failed"
        self.L.arapb200_plan_error.argtypes = [C.c_void_p]

    def solve(self, solver_params: dict, problem_params) -> float:
        for name, ref in solver_params.items():     # setAllSolverParameters (OptUtils.h:104-108)
            self.L.Opt_SetSolverParameter(self.state, self.plan, name.encode(), C.byref(ref))
        self.L.Opt_ProblemSolve(self.state, self.plan, problem_params)
        err = self.L.arapb200_plan_error(self.plan)
        if err:
            raise RuntimeError(f"Opt_ProblemSolve failed with code {err}")
        return float(self.L.Opt_ProblemCurrentCost(self.state, self.plan))

    def close(self):
        if self.plan:
            self.L.Opt_PlanFree(self.state, self.plan)
            self.plan = None
        if self.problem:
            self.L.Opt_ProblemDelete(self.state, self.problem)
            self.problem = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CombinedSolver:
    """ARAP/deformation/src/CombinedSolver.h:99-390 (numIter = 19, nonLinearIter = 8, linearIter = 400:
    ARAP/deformation/src/main.cpp:215-221)."""

    def __init__(self, W: int, H: int, plan_file: str = PLAN, numIter=19, nonLinearIter=8, linearIter=400):
        import torch
        self.torch = torch
        self.W, self.H = W, H
        self.numIter = numIter
        self.nIterations, self.lIterations = C.c_uint(nonLinearIter), C.c_uint(linearIter)
        self.solver = OptSolver(W, H, plan_file)
        dev = torch.device("cuda", torch.cuda.current_device())
        f32 = torch.float32
        # the five OptImages (CombinedSolver.h:161-166) + pinned staging for the uploads / the download
        self.d_offset = torch.empty((H, W, 2), dtype=f32, device=dev)
        self.d_angle = torch.empty((H, W), dtype=f32, device=dev)
        self.d_urshape = torch.empty((H, W, 2), dtype=f32, device=dev)
        self.d_constraints = torch.empty((H, W, 2), dtype=f32, device=dev)
        self.d_mask = torch.empty((H, W), dtype=f32, device=dev)
        self.h_constraints = torch.empty((H, W, 2), dtype=f32).pin_memory()
        self.h_field = torch.empty((H, W, 2), dtype=f32).pin_memory()
        self.h_mask = torch.empty((H, W), dtype=f32).pin_memory()
        yy, xx = np.mgrid[0:H, 0:W]
        self.h_urshape = torch.from_numpy(np.ascontiguousarray(np.stack([xx, yy], -1).astype(np.float32))).pin_memory()
        self.w_fit = C.c_float(float(np.sqrt(np.float32(100.0))))     # combinedSolveInit, :172-177
        self.w_reg = C.c_float(float(np.sqrt(np.float32(0.01))))
        self.costs = []
        self.h2d_bytes = self.d2h_bytes = 0
        self.seconds = {"reset": 0.0, "constraints": 0.0, "solve": 0.0, "download+warp": 0.0}   # wall time per stage, last image

    # ---- addImage (:139-170): keep the inputs, resolve which constraint wins at every source pixel once -------------
    def add_image(self, rgb, mask_red, constraints):
        H, W = self.H, self.W
        self.rgb = np.ascontiguousarray(rgb, np.uint8)
        self.mask_red = np.ascontiguousarray(mask_red, np.uint8)
        m = np.asarray(constraints, np.int32).reshape(-1, 4)
        ok = (m[:, 0] >= 0) & (m[:, 0] < W) & (m[:, 1] >= 0) & (m[:, 1] < H)
        m = m[ok]
        m = m[self.mask_red[m[:, 1], m[:, 0]] == 0]                   # m_orgMask(x, y).x == 0  (:234)
        idx = m[:, 1].astype(np.int64) * W + m[:, 0]
        # later entries overwrite earlier ones (:229-241): keep the LAST occurrence of every source pixel
        _, first_of_reversed = np.unique(idx[::-1], return_index=True)
        keep = np.sort(len(idx) - 1 - first_of_reversed)
        self._m = m[keep].astype(np.float32)
        self._idx = idx[keep]

    def _set_constraint_image(self, alpha):
        """setConstraintImage (:223-242): host lerp in binary32, full-image upload."""
        a = np.float32(alpha)
        om = np.float32(1.0) - a
        Cn = self.h_constraints.numpy().reshape(-1, 2)
        Cn[:] = -1.0
        Cn[self._idx, 0] = om * self._m[:, 0] + a * self._m[:, 2]
        Cn[self._idx, 1] = om * self._m[:, 1] + a * self._m[:, 3]
        self.d_constraints.copy_(self.h_constraints)
        self.h2d_bytes += Cn.nbytes

    def _reset_gpu(self):
        """resetGPU (:207-221)"""
        self.h_mask.numpy()[:] = self.mask_red
        self._set_constraint_image(1.0)
        self.d_urshape.copy_(self.h_urshape)
        self.d_offset.copy_(self.h_urshape)
        self.d_mask.copy_(self.h_mask)
        self.d_angle.zero_()
        self.h2d_bytes += 2 * self.h_urshape.numpy().nbytes + self.h_mask.numpy().nbytes

    def solve_all(self):
        """CombinedSolverBase::singleSolve (CombinedSolverBase.h:99-120)"""
        import time
        self.h2d_bytes = self.d2h_bytes = 0
        t0 = time.perf_counter()
        self._reset_gpu()                                             # preSingleSolve
        t1 = time.perf_counter()
        sec = {"reset": t1 - t0, "constraints": 0.0, "solve": 0.0, "download+warp": 0.0}
        pp = (C.c_void_p * 7)(self.d_offset.data_ptr(), self.d_angle.data_ptr(), self.d_urshape.data_ptr(),
                              self.d_constraints.data_ptr(), self.d_mask.data_ptr(),
                              C.cast(C.byref(self.w_fit), C.c_void_p), C.cast(C.byref(self.w_reg), C.c_void_p))
        sp = {"nIterations": self.nIterations, "lIterations": self.lIterations}
        self.costs = []
        for i in range(self.numIter):
            t0 = time.perf_counter()
            self._set_constraint_image(np.float32(i + 1) / np.float32(self.numIter))   # preNonlinearSolve (:199-201)
            t1 = time.perf_counter()
            self.costs.append(self.solver.solve(sp, pp))
            t2 = time.perf_counter()
            sec["constraints"] += t1 - t0
            sec["solve"] += t2 - t1
        t0 = time.perf_counter()
        # postSingleSolve -> copyResultToCPU (:280-342): download the warp field, rasterise
        self.h_field.copy_(self.d_offset)
        self.d2h_bytes += self.h_field.numpy().nbytes
        pos = self.h_field.numpy()
        self.warped_rgb, self.warped_mask, _ = lib.warp(pos, self.rgb, self.mask_red, want_splat=False)
        N = self.W * self.H
        self.h2d_bytes += 8 * N + 3 * N + N
        self.d2h_bytes += 3 * N + N
        sec["download+warp"] = time.perf_counter() - t0
        self.seconds = sec
        return self.costs[-1]

    def warp_field(self):
        """warpField (:352-366): flow = position - pixel"""
        return self.h_field.numpy() - self.h_urshape.numpy()

    def close(self):
        self.solver.close()
