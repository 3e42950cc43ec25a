/*
 * Opt.h -- C ABI of the solver library, source-compatible with the reference's
 * ARAP/API/release/include/Opt.h:1-71 (10 entry points + Opt_InitializationParameters), so that
 * the reference's callers (ARAP/shared/OptSolver.h:43-91, OptUtils.h:47-64,104-108) compile and link
 * against libarapb200 unchanged.  In the reference these symbols are Terra thunks that forward into a
 * LuaJIT state (ARAP/API/src/createwrapper.t:124-220); here they are a plain extern "C" CUDA library.
 *
 * Scope: the ARAP problem of arap_plan.t solved by "gaussNewtonGPU" on a W x H image domain.
 * Anything else (other energies, "LMGPU", graphs, double precision) is refused loudly.
 */
#ifndef ARAPB200_OPT_H
#define ARAPB200_OPT_H

#ifndef ARAPB200_API
#define ARAPB200_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct Opt_State Opt_State;
typedef struct Opt_Plan Opt_Plan;
typedef struct Opt_Problem Opt_Problem;

/* replaces Opt.h:10-30.  doublePrecision must be 0 (the app never sets it:
 * ARAP/shared/CombinedSolverParameters.h:14); verbosityLevel > 0 prints one cost line per
 * Gauss-Newton step; collectPerKernelTimingInfo and threadsPerBlock are accepted and ignored. */
struct Opt_InitializationParameters {
    int doublePrecision;
    int verbosityLevel;
    int collectPerKernelTimingInfo;
    int threadsPerBlock;
};
typedef struct Opt_InitializationParameters Opt_InitializationParameters;

/* replaces Opt.h:34 (createwrapper.t:124-212).  Several states may be alive at once; there is no
 * matching free in the reference ABI (the state leaks by design, CombinedSolver.h:155-159). */
ARAPB200_API Opt_State* Opt_NewState(Opt_InitializationParameters params);

/* replaces Opt.h:39-40 (o.t:2521-2529).  `filename` must name a readable file holding the ARAP energy
 * (it is identified, not interpreted); `solverkind` must be "gaussNewtonGPU".  NULL on refusal. */
ARAPB200_API Opt_Problem* Opt_ProblemDefine(Opt_State* state, const char* filename, const char* solverkind);
ARAPB200_API void Opt_ProblemDelete(Opt_State* state, Opt_Problem* problem);

/* replaces Opt.h:45-46 (o.t:2530-2540, solverGPUGaussNewton.t:1254-1284): dimensions = {W, H}. */
ARAPB200_API Opt_Plan* Opt_ProblemPlan(Opt_State* state, Opt_Problem* problem, unsigned int* dimensions);
ARAPB200_API void Opt_PlanFree(Opt_State* state, Opt_Plan* plan);

/* replaces Opt.h:50 (solverGPUGaussNewton.t:1205-1221): "nIterations", "lIterations" read as int;
 * the LM knobs are accepted and ignored; unknown names warn (verbosity > 0) and are ignored. */
ARAPB200_API void Opt_SetSolverParameter(Opt_State* state, Opt_Plan* plan, const char* name, void* value);

/* replaces Opt.h:55, 62, 65 (o.t:2542-2551, solverGPUGaussNewton.t:956-1007, 1016-1177).
 * problemparams[0..6] = device float2* Offset (in/out), device float* Angle (in/out),
 * device float2* UrShape, device float2* Constraints, device float* Mask,
 * host float* w_fitSqrt, host float* w_regSqrt (arap_plan.t:2-8).  Every call returns with the
 * device work complete.  Opt_ProblemStep returns 0 once nIterations steps were taken. */
ARAPB200_API void Opt_ProblemSolve(Opt_State* state, Opt_Plan* plan, void** problemparams);
ARAPB200_API void Opt_ProblemInit(Opt_State* state, Opt_Plan* plan, void** problemparams);
ARAPB200_API int Opt_ProblemStep(Opt_State* state, Opt_Plan* plan, void** problemparams);

/* replaces Opt.h:70 (solverGPUGaussNewton.t:1179-1182): cost after the last completed init/step. */
ARAPB200_API double Opt_ProblemCurrentCost(Opt_State* state, Opt_Plan* plan);

#ifdef __cplusplus
}
#endif
#endif
