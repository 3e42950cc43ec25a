-- arap_b200.t -- Terra binding of libarapb200 (the north star's "host code stays in Terra/Lua and reaches CUDA through a
-- thin C-ABI layer").  Terra consumes C headers directly, the way the reference's own wrapper pulls in its C pieces
-- (ARAP/API/src/createwrapper.t:68 terralib.includec), so the whole FFI is the two headers of this repository.
--
-- NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no terra / luajit (SURVEY.md 8c).  The C ABI underneath is
-- exercised by the reference's own C++ host program (oracle/_ref/arap_deform_refhost) and by ctypes (tests/).
--
--   terra arap_b200.t          -- from a directory where include/ and libarapb200.so are reachable, or set
--   ARAPB200_INCLUDE=/path/to/include ARAPB200_LIB=/path/to/libarapb200.so
local inc = os.getenv("ARAPB200_INCLUDE") or "include"
local lib = os.getenv("ARAPB200_LIB") or "arap_flow_b200/libarapb200.so"

local C = terralib.includecstring([[
  #include "Opt.h"        // the reference's own 10-function ABI (ARAP/API/release/include/Opt.h:34-70)
  #include "arapb200.h"   // whole-image calls
]], {"-I", inc})
terralib.linklibrary(lib)

local M = {}

-- whole-image call: replaces deformSingle (ARAP/deformation/src/main.cpp:140-160) with the reference's fixed budget
-- (main.cpp:215-221: 19 continuation steps x 8 Gauss-Newton steps x 400 PCG iterations)
terra M.deform(W : int, H : int, rgb : &uint8, mask_red : &uint8, matches : &int32, n : int,
               flow : &float, wrgb : &uint8, wmask : &uint8) : int
  return C.arapb200_deform(W, H, rgb, mask_red, matches, n, 19, 8, 400, C.ARAPB200_BACKEND_AUTO,
                           flow, wrgb, wmask, nil)
end

-- forward warp from a flow field: replaces ARAP/warping/src/main.cpp:145-225
terra M.warp_flow(W : int, H : int, flow : &float, rgb : &uint8, mask_red : &uint8, wrgb : &uint8, wmask : &uint8) : int
  return C.arapb200_warp_flow(W, H, flow, rgb, mask_red, wrgb, wmask, nil)
end

-- the Opt.h level, as the reference's wrapper exposes it (ARAP/shared/OptSolver.h:43-91): one Opt_ProblemSolve on the
-- caller's device images; params = {Offset, Angle, UrShape, Constraints, Mask (device), &w_fitSqrt, &w_regSqrt (host)}
-- solver_kind: "gaussNewtonGPU" (what the ARAP app asks for) or "LMGPU" (the reference's other kind, o.t:121-124)
terra M.solve_kind(plan_file : rawstring, solver_kind : rawstring, dims : &uint32, params : &&opaque) : double
  var init : C.Opt_InitializationParameters
  init.doublePrecision, init.verbosityLevel, init.collectPerKernelTimingInfo, init.threadsPerBlock = 0, 0, 0, 0
  var st = C.Opt_NewState(init)
  var pr = C.Opt_ProblemDefine(st, plan_file, solver_kind)
  if pr == nil then return -1.0 end
  var pl = C.Opt_ProblemPlan(st, pr, dims)
  if pl == nil then C.Opt_ProblemDelete(st, pr); return -1.0 end
  var nIt : uint32, lIt : uint32 = 8, 400
  C.Opt_SetSolverParameter(st, pl, "nIterations", &nIt)
  C.Opt_SetSolverParameter(st, pl, "lIterations", &lIt)
  C.Opt_ProblemSolve(st, pl, params)
  var c = C.Opt_ProblemCurrentCost(st, pl)
  if C.arapb200_plan_error(pl) ~= 0 then c = -1.0 end
  C.Opt_PlanFree(st, pl)
  C.Opt_ProblemDelete(st, pr)
  return c
end

terra M.solve_once(plan_file : rawstring, dims : &uint32, params : &&opaque) : double
  return M.solve_kind(plan_file, "gaussNewtonGPU", dims, params)
end

return M
