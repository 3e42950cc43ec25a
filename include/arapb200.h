/*
 * arapb200.h -- flat C ABI above the Opt.h level: what the reference's L3/L4 code does around the
 * solver (ARAP/deformation/src/CombinedSolver.h, ARAP/deformation/src/main.cpp,
 * ARAP/warping/src/main.cpp), as whole-image calls with plain pointers and sizes.  This is the
 * surface a Terra host reaches with terralib.includec and the one bench.py / the
 * Python mirror bind with ctypes.  All functions return 0 on success, non-zero on failure (message on
 * stderr); nothing in the library calls exit(): failures inside the Opt_* entry points are recorded on the
 * plan (arapb200_plan_error).
 */
#ifndef ARAPB200_H
#define ARAPB200_H

#include <stddef.h>
#include <stdint.h>

#ifndef ARAPB200_API
#define ARAPB200_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ARAPB200_BACKEND_AUTO 0     /* resident when the problem fits on chip, else streaming */
#define ARAPB200_BACKEND_STREAM 1   /* graph-captured 2-kernel PCG iteration, state in L2/HBM */
#define ARAPB200_BACKEND_RESIDENT 2 /* persistent cooperative kernel, state in registers/smem */

/* library / device info: fills sm_count, l2_bytes, cc_major, cc_minor (any may be NULL) */
ARAPB200_API int arapb200_device_info(int* sm_count, size_t* l2_bytes, int* cc_major, int* cc_minor);
ARAPB200_API const char* arapb200_version(void);

/* ---- forward warp (replaces ARAP/warping/src/main.cpp:145-225 == CombinedSolver.h:280-342) ----
 * Host buffers.  pos float2[W*H] absolute positions, rgb uint8x3[W*H], mask_red uint8[W*H]
 * (0 = object).  Outputs: out_rgb uint8x3[W*H], out_mask uint8[W*H] (255 where something landed),
 * out_splat uint32[W*H] = 1 + 2*(y*W+x) + t of the winning triangle (0 = empty); out_splat may be NULL. */
ARAPB200_API int arapb200_warp(int W, int H, const float* pos, const uint8_t* rgb, const uint8_t* mask_red,
                  uint8_t* out_rgb, uint8_t* out_mask, uint32_t* out_splat);
/* warp tool front half (warping/src/main.cpp:160-166): positions = grid + flow, then warp */
ARAPB200_API int arapb200_warp_flow(int W, int H, const float* flow, const uint8_t* rgb, const uint8_t* mask_red,
                       uint8_t* out_rgb, uint8_t* out_mask, uint32_t* out_splat);

/* ---- one image / segment, end to end (replaces deformSingle, deformation/src/main.cpp:140-160:
 * border pins :130-136, resetGPU + continuation CombinedSolver.h:172-242, solve, warp, flow :352-366).
 * Host buffers in, host buffers out.  matches int32[4*n] (x1 y1 x2 y2) WITHOUT the border pins.
 * out_flow float2[W*H]; out_rgb/out_mask as arapb200_warp; out_costs float[nCont*(nGN+1)] (may be
 * NULL): cost at init and after every Gauss-Newton step of every continuation step. */
ARAPB200_API int arapb200_deform(int W, int H, const uint8_t* rgb, const uint8_t* mask_red, const int32_t* matches,
                    int n_matches, int nCont, int nGN, int nPCG, int backend, float* out_flow,
                    uint8_t* out_rgb, uint8_t* out_mask, float* out_costs);

/* ---- "next" row N1: layer flatten + background composite (replaces para_gen.py:136-175 flatten and :50-61 add_bg).
 * n_layers per-segment results of one --multseg pair, in segment order: flows[s] float2[W*H], rgbs[s] uint8x3[W*H],
 * masks[s] uint8[W*H] (warped masks, non-zero = object).  Later segments overwrite earlier ones where their mask is
 * non-zero; where the final mask is 0 the colour is taken from `background` (uint8x3[W*H], may be NULL). */
ARAPB200_API int arapb200_flatten(int W, int H, int n_layers, const float* const* flows, const uint8_t* const* rgbs,
                                  const uint8_t* const* masks, const uint8_t* background, float* out_flow,
                                  uint8_t* out_rgb, uint8_t* out_mask);

/* ---- "next" row N3: match ingestion on the device (replaces the Python loops of para_gen.py:216-223 valid_cnstr,
 * :468-482 filter, :513-537 per-segment masks).
 * labels1 uint8[W1*H1], labels2 uint8[W2*H2]: the segment-label images of frame 1 and 2 (0 = background).
 * A raw match (x1 y1 x2 y2) is kept iff both points are inside their images, 0 < |p2 - p1| < 60 px, the label under
 * p1 is non-zero and equals the label under p2.  Order is preserved (later constraints overwrite earlier ones
 * downstream, CombinedSolver.h:223-242).  out_matches int32[4*n], out_labels uint8[n] (label under p1 of every kept
 * match; may be NULL), *n_out = number kept.  Negative coordinates are rejected (the reference would wrap them). */
ARAPB200_API int arapb200_filter_matches(int W1, int H1, const uint8_t* labels1, int W2, int H2, const uint8_t* labels2,
                                         const int32_t* matches, int n, int32_t* out_matches, uint8_t* out_labels,
                                         int* n_out);
/* The mask image arap_deform expects (red channel 0 = solve here, 255 = ARAP_BG, para_gen.py:30): segment > 0 selects
 * one label (--multseg, :526-527), segment == 0 selects every non-zero label (:515-516).  out_mask uint8[W*H]. */
ARAPB200_API int arapb200_segment_mask(int W, int H, const uint8_t* labels, int segment, uint8_t* out_mask);

/* ---- batched, pipelined variant: many independent (image, segment) problems per GPU --------- */
typedef struct arapb200_batch arapb200_batch;
/* max_problems problems of at most maxW x maxH in flight on the current device */
ARAPB200_API arapb200_batch* arapb200_batch_create(int maxW, int maxH, int max_problems, int nCont, int nGN, int nPCG,
                                      int backend);
ARAPB200_API void arapb200_batch_destroy(arapb200_batch* b);
/* enqueue problem `slot` (0 <= slot < max_problems); host pointers must stay valid until _wait */
ARAPB200_API int arapb200_batch_submit(arapb200_batch* b, int slot, int W, int H, const uint8_t* rgb,
                          const uint8_t* mask_red, const int32_t* matches, int n_matches, float* out_flow,
                          uint8_t* out_rgb, uint8_t* out_mask, float* out_costs);
/* run everything submitted since the last run; returns after all outputs are in host memory */
ARAPB200_API int arapb200_batch_run(arapb200_batch* b);
/* device milliseconds of the last run: [0] total, [1] solve kernels only, [2] warp kernels only */
ARAPB200_API int arapb200_batch_timing(arapb200_batch* b, float* ms3);
/* number of kernel launches issued by the last run */
ARAPB200_API long long arapb200_batch_launches(arapb200_batch* b);

/* how many problems of the last run were solved by the resident (on-chip) back-end; the rest took the streaming one */
ARAPB200_API int arapb200_batch_resident_count(arapb200_batch* b);

/* shape of the last cooperative launch of the last run: info6 = {problems sharing the launch, kernel variant's max
 * threads, variant's min CTAs per SM, grid.x, grid.y, threads per CTA}; zeros when every problem streamed.  Lets a test
 * assert that it exercised the launch shape a benchmark times. */
ARAPB200_API int arapb200_batch_launch_info(arapb200_batch* b, int* info6);

/* Options beyond the reference's behaviour; every one defaults to "off" and none is on the parity path.
 *   "pcg_rtol" (SURVEY.md 8f N4): 0 <= value < 1.  > 0: a PCG loop ends as soon as r.z <= value^2 * (r.z at its start)
 *   instead of always running lIterations iterations.  Changes results (by design).  Both back-ends honour it (the
 *   streaming one turns the remaining kernels of its captured graph into no-ops).
 *   "gn_rtol": 0 <= value < 1.  > 0: the Gauss-Newton steps of a continuation step end as soon as one of them lowers
 *   the cost by less than value (relative) instead of always running nIterations steps; the skipped entries of
 *   out_costs repeat the last cost.  Same scope and caveats as "pcg_rtol" (both back-ends).
 *   "cluster_barrier": 1 = problems that fit one thread-block cluster (<= 16 CTAs of <= 12 strips) run with a cluster-scope
 *   barrier (limb sums pushed through distributed shared memory, mbarrier completion); 0 (default) = every resident
 *   problem uses the L2 barrier.  Same results bit for bit; measured slower on B200, kept as an option.
 *   "lm": 1 = every Opt_ProblemSolve of the schedule is run by the reference's other solver kind, "LMGPU"
 *   (Levenberg-Marquardt trust region, Q-based exit of the linear loops, step acceptance / revert;
 *   ARAP/API/src/solverGPUGaussNewton.t with UsesLambda()) with its default solver parameters (:26-39), one problem at a
 *   time; nIterations / lIterations stay upper bounds.  Changes results (a different algorithm); out_costs holds the
 *   accepted cost after init and after every step, entries of steps not taken repeat the last one.  0 (default) =
 *   gaussNewtonGPU, what the ARAP app requests (CombinedSolverBase.h:76).
 * Returns non-zero for unknown names / bad values. */
ARAPB200_API int arapb200_batch_set_option(arapb200_batch* b, const char* name, double value);

/* Error state of a plan made by Opt_ProblemPlan (include/Opt.h).  Opt_ProblemInit / Step / Solve have no error return
 * in the reference's ABI (ARAP/API/release/include/Opt.h:60-68) and the reference exits the process on a CUDA failure
 * (ARAP/API/src/solverGPUGaussNewton.t:59-73); this library records the failure instead: 0 = fine, otherwise the code of
 * the first failure (Opt_ProblemCurrentCost then returns NaN and Opt_ProblemStep reports "finished"). */
struct Opt_Plan;
ARAPB200_API int arapb200_plan_error(struct Opt_Plan* plan);

/* Opt_InitializationParameters.collectPerKernelTimingInfo (ARAP/API/release/include/Opt.h:22-24, util.t:404-510): every
 * kernel launch of a solve is bracketed by a CUDA event pair under the reference's kernel name and aggregated at the end of
 * the solve (when Opt_ProblemStep returns 0 / Opt_ProblemSolve returns).  Such a plan runs on the streaming back-end with
 * eager launches (a persistent kernel has nothing to time per kernel; results are the same bits).  With verbosityLevel > 0
 * the table is printed like the reference's Timer:evaluate does; this call returns it in any case: copies at most
 * cap - 1 characters into buf (may be NULL), returns the full length (0 = no finished solve / no timing). */
ARAPB200_API size_t arapb200_plan_timing_report(struct Opt_Plan* plan, char* buf, size_t cap);

/* Opt_ProblemDefine(state, file, "LMGPU") selects the reference's other solver kind (ARAP/API/src/o.t:121-124,
 * ARAP/API/src/solverGPUGaussNewton.t with UsesLambda(): Levenberg-Marquardt trust region around the same PCG, Q-based
 * early exit of the linear loop, step acceptance / revert; never requested by the ARAP app).  Its solver parameters
 * (Opt_SetSolverParameter, :26-39, :140-157) are floats -- min_relative_decrease, min_trust_region_radius,
 * max_trust_region_radius, q_tolerance, function_tolerance, trust_region_radius, radius_decrease_factor,
 * min_lm_diagonal, max_lm_diagonal -- and the int residual_reset_period.  This call reports what the last
 * Opt_ProblemStep of such a plan did: info6 = trust-region radius after the step, linear iterations run, verdict
 * (1 accepted, 0 reverted, 2 function tolerance reached, 3 radius below the minimum), model cost, cost at the trial
 * point, last Q.  Returns 0, or 1 when the plan is not an LM plan. */
ARAPB200_API int arapb200_plan_lm_info(struct Opt_Plan* plan, float info6[6]);

/* ---- debug / parity entry points (unit-level comparison against the oracle) ----------------- */
/* one Opt_ProblemSolve on host buffers: X float2[N] and A float[N] in/out; U, C float2[N]; M float[N];
 * costs float[nGN+1] (may be NULL); scal float[nGN*nPCG*3] (den, num, bnum per PCG iteration; may be NULL) */
ARAPB200_API int arapb200_debug_gn_solve(int W, int H, float* X, float* A, const float* U, const float* C, const float* M,
                            float wf, float wr, int nGN, int nPCG, int backend, float* costs, float* scal);
/* r = -J^T F and pre (float3[N] each, zero on inactive pixels) */
ARAPB200_API int arapb200_debug_eval_jtf(int W, int H, const float* X, const float* A, const float* U, const float* C,
                            const float* M, float wf, float wr, float* r3, float* pre3);
/* q = J^T J p (float3[N]); *dot = p.q per the summation contract */
ARAPB200_API int arapb200_debug_apply_jtj(int W, int H, const float* A, const float* U, const float* C, const float* M,
                             float wf, float wr, const float* p3, float* q3, float* dot);
/* cost per the contract */
ARAPB200_API int arapb200_debug_cost(int W, int H, const float* X, const float* A, const float* U, const float* C,
                        const float* M, float wf, float wr, float* cost);
/* cycle accounting of the resident kernel: prof uint64[160*16] (per CTA: cycles in PCG phase 1/2/3, other,
 * barrier skew / poll / fold, barrier count), info int[3] = {strips, CTAs, warps per CTA}, *ms = launch time */
ARAPB200_API int arapb200_debug_resident_profile(int W, int H, const uint8_t* mask_red, const int32_t* matches,
                                                 int n_matches, int nCont, int nGN, int nPCG,
                                                 unsigned long long* prof, int* info, float* ms);
/* the same for the first of `copies` (1..16) identical problems that share ONE cooperative launch: what one problem's
 * phases and barriers cost next to co-resident neighbours; returns 4 when they do not fit one launch */
ARAPB200_API int arapb200_debug_resident_profile_group(int W, int H, const uint8_t* mask_red, const int32_t* matches,
                                                       int n_matches, int copies, int nCont, int nGN, int nPCG,
                                                       unsigned long long* prof, int* info, float* ms);
/* contract sincos and exact sum on the device */
ARAPB200_API int arapb200_debug_sincos(int n, const float* a, float* s, float* c);
ARAPB200_API int arapb200_debug_exact_sum(size_t n, const float* t, float* sum);
/* the same sum through the streaming back-end's fence-free wide fixed-point accumulators (one block per 256 terms) */
ARAPB200_API int arapb200_debug_wide_sum(size_t n, const float* t, float* sum);

#ifdef __cplusplus
}
#endif
#endif
