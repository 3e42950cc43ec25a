/*
 * arap_literal.c -- a SECOND, independently-ordered CPU implementation of the reference ARAP solve, written
 * to behave like the reference's generated CUDA rather than like this repo's arithmetic contract.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as arap_oracle.c): tests/, tools/literal_bound.py.  Never the product.
 *
 * Purpose.  The reference solver (Terra/Opt-generated CUDA) cannot be built here, so bit-level parity of the
 * solve is unpinned.  What CAN be established is a bound: how far can an implementation
 * land from the contract path (oracle == CUDA kernels, bit for bit) when it differs in everything the reference
 * leaves unspecified -- association order, FMA contraction, the sin/cos routine, and the order in which
 * per-warp partial sums reach the accumulator?  This file differs from arap_oracle.c in all of them at once:
 *
 *   schedule      the reference's kernel sequence, unfused: PCGInit1 | (PCGStep1, PCGStep2, PCGStep3) x lIterations |
 *                 PCGLinearUpdate | computeCost, AoS state vectors delta, r, z, p, Ap_X, preconditioner
 *                 (ARAP/API/src/solverGPUGaussNewton.t:361-397, 421-434, 446-489, 537-557, 580-592, 1016-1177)
 *   derivatives   residual-centric, the way createjtfcentered / createjtjcentered build them
 *                 (ARAP/API/src/o.t:2129-2172, 2029-2089): for every residual that contains the unknown,
 *                 dr/dx * r   and   dr/dx * sum_u dr/du p_u   -- NOT the collapsed "S-form" of the contract;
 *                 diag(J^T J) is the literal sum of squared derivatives (with the actual sin/cos), not |d|^2
 *   transcend.    libm sinf / cosf / sqrtf (the reference: CUDA 7.5 libdevice __nv_sinf/__nv_cosf,
 *                 ARAP/API/src/util.t:160-174).  The reference re-evaluates sin/cos of the 5 stencil angles in
 *                 every PCGStep1; Angle does not change inside a Gauss-Newton step and sinf/cosf are pure
 *                 functions, so a per-GN-step table of the SAME libm values is the same arithmetic.  Flag
 *                 LIT_SINCOS_EVERY_ITER evaluates them in every iteration anyway (slow; tests use it on a small
 *                 case to show the two are bit-identical).
 *   contraction   built with -O2 -ffp-contract=fast -mfma: the compiler fuses what it likes, as LLVM's NVPTX
 *                 back-end did for the reference
 *   reductions    util.t:612-623 + 528-531 + solverGPUGaussNewton.t:312-317: 16x16 thread blocks, a warp = two
 *                 16-pixel rows of a block, shfl.down tree inside the warp, then ONE fp32 atomic add per warp
 *                 into a single word.  The arrival order of those atomics is scheduling-dependent on the GPU;
 *                 here it is a seeded pseudo-random permutation of the warps, different for every reduction.
 *                 fp32 accumulation throughout, nothing exact.
 *
 * Continuation / constraint image / border pins / reset follow ARAP/deformation/src/CombinedSolver.h:172-242,
 * ARAP/shared/CombinedSolverBase.h:99-120 and ARAP/deformation/src/main.cpp:130-136 (those are host code in
 * the reference and have only one reasonable reading).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LIT_API __attribute__((visibility("default")))
#define LIT_SINCOS_EVERY_ITER 1
#define LIT_ORDERED_ATOMICS 2 /* debug: warps arrive in launch order instead of a shuffled one */

#define BLK 16
#define NPERM 61

typedef struct {
    int W, H, BX, BY, nwarps;
    const float *U, *C, *M; /* UrShape float2, Constraints float2, Mask float */
    float wf, wr;
    float *X, *A;                           /* unknowns: Offset float2[N], Angle float[N] */
    float *delta, *r, *z, *p, *Ap, *pre;    /* float3[N] each */
    float *cosA, *sinA;                     /* libm table of the current Gauss-Newton step */
    float *wsum;                            /* one partial per warp */
    int *perm;                              /* NPERM permutations of the warps */
    uint64_t rng;
    int flags;
} Lit;

static uint64_t rng_next(uint64_t *s)
{
    uint64_t x = *s;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    *s = x;
    return x * 2685821657736338717ull;
}

static inline int lit_active(const Lit *L, int x, int y) { return L->M[(size_t)y * L->W + x] == 0.0f; }

/* neighbour n of (x, y) in arap_plan.t:14 order; valid = InBounds * eq(Mask(x,y),0) * eq(Mask(0,0),0) (:17) */
static inline int lit_nb(const Lit *L, int x, int y, int n, int *xj, int *yj)
{
    static const int dx[4] = {1, -1, 0, 0}, dy[4] = {0, 0, 1, -1};
    *xj = x + dx[n];
    *yj = y + dy[n];
    if (*xj < 0 || *xj >= L->W || *yj < 0 || *yj >= L->H) return 0;
    return lit_active(L, *xj, *yj) && lit_active(L, x, y);
}

static inline void lit_cs(const Lit *L, size_t i, float *c, float *s)
{
    if (L->flags & LIT_SINCOS_EVERY_ITER) { *c = cosf(L->A[i]); *s = sinf(L->A[i]); }
    else { *c = L->cosA[i]; *s = L->sinA[i]; }
}

/* ---- the reference's grid-wide reduction: warp shuffle tree, then one fp32 atomic per warp ---------------- */
/* val[] holds one value per pixel (0 for excluded pixels); returns what the accumulator word would hold */
static float lit_reduce(Lit *L, const float *val)
{
    const int W = L->W, H = L->H;
#pragma omp parallel for schedule(static)
    for (int by = 0; by < L->BY; ++by)
        for (int bx = 0; bx < L->BX; ++bx)
            for (int w = 0; w < BLK * BLK / 32; ++w) {
                float v[32];
                for (int l = 0; l < 32; ++l) {
                    const int x = bx * BLK + (l & 15), y = by * BLK + 2 * w + (l >> 4);
                    v[l] = (x < W && y < H) ? val[(size_t)y * W + x] : 0.0f;
                }
                for (int off = 16; off > 0; off >>= 1)  /* util.t:612-623 */
                    for (int l = 0; l < off; ++l) v[l] = v[l] + v[l + off];
                L->wsum[((size_t)by * L->BX + bx) * (BLK * BLK / 32) + w] = v[0];
            }
    const int n = L->nwarps;
    float acc = 0.0f; /* cudaMemset before the kernel */
    if (L->flags & LIT_ORDERED_ATOMICS) {
        for (int k = 0; k < n; ++k) acc = acc + L->wsum[k];
        return acc;
    }
    const uint64_t r = rng_next(&L->rng);
    const int *perm = L->perm + (size_t)(r % NPERM) * n;
    const int off = (int)((r >> 20) % (uint64_t)n);
    for (int k = 0; k < n; ++k) {
        int j = k + off;
        if (j >= n) j -= n;
        acc = acc + L->wsum[perm[j]]; /* util.t:528-531 red.global.add.f32 */
    }
    return acc;
}

/* ---- derived functions, residual-centric ------------------------------------------------------------------ */
/* e = w_regSqrt * ((Offset(0,0) - Offset(n)) - Rotate2D(Angle(0,0), UrShape(0,0) - UrShape(n)))  (arap_plan.t:15-16) */
static inline void lit_ereg(const Lit *L, size_t i, size_t j, float c, float s, float e[2], float Rd[2], float dRd[2])
{
    const float dx = L->U[2 * i] - L->U[2 * j], dy = L->U[2 * i + 1] - L->U[2 * j + 1];
    Rd[0] = c * dx + (-s) * dy;   /* lib.t:92-96: matrix = (cos, -sin, sin, cos) */
    Rd[1] = s * dx + c * dy;
    dRd[0] = (-s) * dx + (-c) * dy; /* d/da */
    dRd[1] = c * dx + (-s) * dy;
    e[0] = L->wr * ((L->X[2 * i] - L->X[2 * j]) - Rd[0]);
    e[1] = L->wr * ((L->X[2 * i + 1] - L->X[2 * j + 1]) - Rd[1]);
}

static inline int lit_fit(const Lit *L, size_t i) { return L->C[2 * i] >= 0.0f && L->C[2 * i + 1] >= 0.0f; } /* arap_plan.t:22 */

/* evalJTF (o.t:2129-2172): F_hat = sum dr/dx00 * r, P_hat = sum (dr/dx00)^2 over every residual containing x00 */
static void lit_eval_jtf(const Lit *L, int x, int y, float g[3], float D[3])
{
    const size_t i = (size_t)y * L->W + x;
    float ci, si;
    lit_cs(L, i, &ci, &si);
    g[0] = g[1] = g[2] = 0.0f;
    D[0] = D[1] = D[2] = 0.0f;
    const float wr = L->wr, wf = L->wf;
    for (int n = 0; n < 4; ++n) {
        int xj, yj;
        if (!lit_nb(L, x, y, n, &xj, &yj)) continue;
        const size_t j = (size_t)yj * L->W + xj;
        float e[2], Rd[2], dRd[2];
        /* own residual r(i, n): d/dOffset_i = wr, d/dAngle_i = -wr * R'(a_i) d */
        lit_ereg(L, i, j, ci, si, e, Rd, dRd);
        g[0] += wr * e[0];
        g[1] += wr * e[1];
        D[0] += wr * wr;
        D[1] += wr * wr;
        const float da0 = -wr * dRd[0], da1 = -wr * dRd[1];
        g[2] += da0 * e[0];
        g[2] += da1 * e[1];
        D[2] += da0 * da0;
        D[2] += da1 * da1;
        /* the neighbour's residual r(j, -n) that contains Offset_i: d/dOffset_i = -wr */
        float cj, sj, ej[2];
        lit_cs(L, j, &cj, &sj);
        lit_ereg(L, j, i, cj, sj, ej, Rd, dRd);
        g[0] += (-wr) * ej[0];
        g[1] += (-wr) * ej[1];
        D[0] += wr * wr;
        D[1] += wr * wr;
    }
    if (lit_fit(L, i)) { /* w_fitSqrt * (Offset - Constraints) */
        g[0] += wf * (wf * (L->X[2 * i] - L->C[2 * i]));
        g[1] += wf * (wf * (L->X[2 * i + 1] - L->C[2 * i + 1]));
        D[0] += wf * wf;
        D[1] += wf * wf;
    }
}

/* applyJTJ (o.t:2029-2089): for every residual r containing x00: dr/dx00 * (sum_u dr/du * P_u) */
static void lit_apply_jtj(const Lit *L, int x, int y, float q[3])
{
    const size_t i = (size_t)y * L->W + x;
    float ci, si;
    lit_cs(L, i, &ci, &si);
    q[0] = q[1] = q[2] = 0.0f;
    const float wr = L->wr, wf = L->wf;
    const float *p = L->p;
    for (int n = 0; n < 4; ++n) {
        int xj, yj;
        if (!lit_nb(L, x, y, n, &xj, &yj)) continue;
        const size_t j = (size_t)yj * L->W + xj;
        const float dx = L->U[2 * i] - L->U[2 * j], dy = L->U[2 * i + 1] - L->U[2 * j + 1];
        /* own residual */
        const float dRi0 = (-si) * dx + (-ci) * dy, dRi1 = ci * dx + (-si) * dy;
        const float Jp0 = wr * p[3 * i] + (-wr) * p[3 * j] + (-wr * dRi0) * p[3 * i + 2];
        const float Jp1 = wr * p[3 * i + 1] + (-wr) * p[3 * j + 1] + (-wr * dRi1) * p[3 * i + 2];
        q[0] += wr * Jp0;
        q[1] += wr * Jp1;
        q[2] += (-wr * dRi0) * Jp0;
        q[2] += (-wr * dRi1) * Jp1;
        /* neighbour's residual r(j, -n): unknowns Offset_j (+wr), Offset_i (-wr), Angle_j (-wr R'(a_j) d_ji) */
        float cj, sj;
        lit_cs(L, j, &cj, &sj);
        const float ex = -dx, ey = -dy; /* d_ji */
        const float dRj0 = (-sj) * ex + (-cj) * ey, dRj1 = cj * ex + (-sj) * ey;
        const float Kp0 = wr * p[3 * j] + (-wr) * p[3 * i] + (-wr * dRj0) * p[3 * j + 2];
        const float Kp1 = wr * p[3 * j + 1] + (-wr) * p[3 * i + 1] + (-wr * dRj1) * p[3 * j + 2];
        q[0] += (-wr) * Kp0;
        q[1] += (-wr) * Kp1;
    }
    if (lit_fit(L, i)) {
        q[0] += wf * (wf * p[3 * i]);
        q[1] += wf * (wf * p[3 * i + 1]);
    }
}

/* cost (o.t:2375-2385): 0.5 * sum of squared residuals owned by the pixel */
static float lit_cost_px(const Lit *L, int x, int y)
{
    const size_t i = (size_t)y * L->W + x;
    float ci, si, sum = 0.0f;
    lit_cs(L, i, &ci, &si);
    for (int n = 0; n < 4; ++n) {
        int xj, yj;
        if (!lit_nb(L, x, y, n, &xj, &yj)) continue;
        float e[2], Rd[2], dRd[2];
        lit_ereg(L, i, (size_t)yj * L->W + xj, ci, si, e, Rd, dRd);
        sum += e[0] * e[0];
        sum += e[1] * e[1];
    }
    if (lit_fit(L, i)) {
        const float f0 = L->wf * (L->X[2 * i] - L->C[2 * i]), f1 = L->wf * (L->X[2 * i + 1] - L->C[2 * i + 1]);
        sum += f0 * f0;
        sum += f1 * f1;
    }
    return 0.5f * sum;
}

static void lit_fill_cs(Lit *L)
{
    const size_t N = (size_t)L->W * L->H;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; ++i) {
        L->cosA[i] = cosf(L->A[i]);
        L->sinA[i] = sinf(L->A[i]);
    }
}

static float lit_cost(Lit *L, float *scratch)
{
    lit_fill_cs(L);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < L->H; ++y)
        for (int x = 0; x < L->W; ++x)
            scratch[(size_t)y * L->W + x] = lit_active(L, x, y) ? lit_cost_px(L, x, y) : 0.0f;
    return lit_reduce(L, scratch);
}

static inline float guarded_invert(float v) /* solverGPUGaussNewton.t:323-332 */
{
    const float t = 1.0f + sqrtf(v);
    return 1.0f / (t * t);
}

/* one Gauss-Newton step = solverGPUGaussNewton.t:1016-1177 for gaussNewtonGPU */
static float lit_gn_step(Lit *L, int nPCG, float *scratch)
{
    const int W = L->W, H = L->H;
    lit_fill_cs(L);
    /* PCGInit1 */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            float d = 0.0f;
            if (lit_active(L, x, y)) {
                float g[3], D[3];
                lit_eval_jtf(L, x, y, g, D);
                for (int k = 0; k < 3; ++k) {
                    L->delta[3 * i + k] = 0.0f;
                    L->r[3 * i + k] = -g[k];
                    L->pre[3 * i + k] = guarded_invert(D[k]);
                    L->p[3 * i + k] = L->pre[3 * i + k] * L->r[3 * i + k];
                }
                d = L->r[3 * i] * L->p[3 * i] + L->r[3 * i + 1] * L->p[3 * i + 1] + L->r[3 * i + 2] * L->p[3 * i + 2];
            } else {
                L->pre[3 * i] = L->pre[3 * i + 1] = L->pre[3 * i + 2] = 0.0f; /* :392 */
            }
            scratch[i] = d;
        }
    float num = lit_reduce(L, scratch); /* scanAlphaNumerator */
    for (int it = 0; it < nPCG; ++it) {
        /* PCGStep1 */
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t i = (size_t)y * W + x;
                float d = 0.0f;
                if (lit_active(L, x, y)) {
                    float q[3];
                    lit_apply_jtj(L, x, y, q);
                    L->Ap[3 * i] = q[0]; L->Ap[3 * i + 1] = q[1]; L->Ap[3 * i + 2] = q[2];
                    d = L->p[3 * i] * q[0] + L->p[3 * i + 1] * q[1] + L->p[3 * i + 2] * q[2];
                }
                scratch[i] = d;
            }
        const float den = lit_reduce(L, scratch); /* scanAlphaDenominator */
        /* PCGStep2 */
        float alpha = 0.0f;
        if (den > 0.0f) alpha = num / den; /* :456-459 */
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t i = (size_t)y * W + x;
                float b = 0.0f;
                if (lit_active(L, x, y)) {
                    for (int k = 0; k < 3; ++k) {
                        L->delta[3 * i + k] = L->delta[3 * i + k] + alpha * L->p[3 * i + k];
                        L->r[3 * i + k] = L->r[3 * i + k] - alpha * L->Ap[3 * i + k];
                        L->z[3 * i + k] = L->pre[3 * i + k] * L->r[3 * i + k];
                    }
                    b = L->z[3 * i] * L->r[3 * i] + L->z[3 * i + 1] * L->r[3 * i + 1] + L->z[3 * i + 2] * L->r[3 * i + 2];
                }
                scratch[i] = b;
            }
        const float bnum = lit_reduce(L, scratch); /* scanBetaNumerator */
        /* PCGStep3 */
        float beta = 0.0f;
        if (num > 0.0f) beta = bnum / num; /* :544-547 */
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t i = (size_t)y * W + x;
                if (lit_active(L, x, y))
                    for (int k = 0; k < 3; ++k) L->p[3 * i + k] = L->z[3 * i + k] + beta * L->p[3 * i + k];
            }
        num = bnum; /* :1091 */
    }
    /* PCGLinearUpdate */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            if (lit_active(L, x, y)) {
                L->X[2 * i] = L->X[2 * i] + L->delta[3 * i];
                L->X[2 * i + 1] = L->X[2 * i + 1] + L->delta[3 * i + 1];
                L->A[i] = L->A[i] + L->delta[3 * i + 2];
            }
        }
    return lit_cost(L, scratch);
}

/* Whole arap_deform solve of one image / segment.  matches int32[4*n] WITHOUT border pins (added here, main.cpp:130-136).
 * X float2[N], A float[N] out; costs float[nCont*(nGN+1)] or NULL.  seed selects the atomic arrival orders. */
LIT_API int arap_literal_solve(int W, int H, const unsigned char *mask_red, const int *matches, int n_matches, int nCont,
                               int nGN, int nPCG, unsigned long long seed, int flags, float *X, float *A, float *costs)
{
    const size_t N = (size_t)W * H;
    Lit L;
    memset(&L, 0, sizeof(L));
    L.W = W; L.H = H; L.BX = (W - 1) / BLK + 1; L.BY = (H - 1) / BLK + 1; /* util.t:822-841 */
    L.nwarps = L.BX * L.BY * (BLK * BLK / 32);
    L.flags = flags;
    L.rng = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    if (L.rng == 0) L.rng = 1;
    float *U = (float *)malloc(2 * N * sizeof(float)), *C = (float *)malloc(2 * N * sizeof(float));
    float *M = (float *)malloc(N * sizeof(float)), *scratch = (float *)malloc(N * sizeof(float));
    L.delta = (float *)calloc(3 * N, sizeof(float)); L.r = (float *)calloc(3 * N, sizeof(float));
    L.z = (float *)calloc(3 * N, sizeof(float)); L.p = (float *)calloc(3 * N, sizeof(float));
    L.Ap = (float *)calloc(3 * N, sizeof(float)); L.pre = (float *)calloc(3 * N, sizeof(float));
    L.cosA = (float *)malloc(N * sizeof(float)); L.sinA = (float *)malloc(N * sizeof(float));
    L.wsum = (float *)malloc((size_t)L.nwarps * sizeof(float));
    L.perm = (int *)malloc((size_t)NPERM * L.nwarps * sizeof(int));
    if (!U || !C || !M || !scratch || !L.delta || !L.r || !L.z || !L.p || !L.Ap || !L.pre || !L.cosA || !L.sinA || !L.wsum || !L.perm)
        return 1;
    for (int k = 0; k < NPERM; ++k) { /* Fisher-Yates */
        int *pm = L.perm + (size_t)k * L.nwarps;
        for (int j = 0; j < L.nwarps; ++j) pm[j] = j;
        for (int j = L.nwarps - 1; j > 0; --j) {
            const int t = (int)(rng_next(&L.rng) % (uint64_t)(j + 1));
            const int tmp = pm[j]; pm[j] = pm[t]; pm[t] = tmp;
        }
    }
    /* resetGPU (CombinedSolver.h:207-221) */
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            U[2 * i] = (float)x; U[2 * i + 1] = (float)y;
            X[2 * i] = (float)x; X[2 * i + 1] = (float)y;
            A[i] = 0.0f;
            M[i] = (float)mask_red[i];
        }
    L.U = U; L.C = C; L.M = M; L.X = X; L.A = A;
    L.wf = sqrtf(100.0f); L.wr = sqrtf(0.01f); /* CombinedSolver.h:172-177 */
    for (int t = 0; t < nCont; ++t) {
        const float alpha = (float)(t + 1) / (float)nCont; /* :199-201 */
        /* setConstraintImage (:223-242), then the border pins appended by main.cpp:130-136 */
        for (size_t i = 0; i < N; ++i) { C[2 * i] = -1.0f; C[2 * i + 1] = -1.0f; }
        for (int k = 0; k < n_matches; ++k) {
            const int x = matches[4 * k], y = matches[4 * k + 1];
            if (x < 0 || x >= W || y < 0 || y >= H) continue; /* the reference would read out of bounds */
            if (mask_red[(size_t)y * W + x] == 0) {
                C[2 * ((size_t)y * W + x)] = (1.0f - alpha) * (float)x + alpha * (float)matches[4 * k + 2];
                C[2 * ((size_t)y * W + x) + 1] = (1.0f - alpha) * (float)y + alpha * (float)matches[4 * k + 3];
            }
        }
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                if ((y == 0 || x == 0 || y == H - 1 || x == W - 1) && mask_red[(size_t)y * W + x] == 0) {
                    C[2 * ((size_t)y * W + x)] = (1.0f - alpha) * (float)x + alpha * (float)x;
                    C[2 * ((size_t)y * W + x) + 1] = (1.0f - alpha) * (float)y + alpha * (float)y;
                }
        const float c0 = lit_cost(&L, scratch); /* init, :956-1007 */
        if (costs) costs[(size_t)t * (nGN + 1)] = c0;
        for (int g = 0; g < nGN; ++g) {
            const float c = lit_gn_step(&L, nPCG, scratch);
            if (costs) costs[(size_t)t * (nGN + 1) + g + 1] = c;
        }
    }
    free(U); free(C); free(M); free(scratch); free(L.delta); free(L.r); free(L.z); free(L.p); free(L.Ap); free(L.pre);
    free(L.cosA); free(L.sinA); free(L.wsum); free(L.perm);
    return 0;
}

LIT_API int arap_literal_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
