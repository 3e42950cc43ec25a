/*
 * arap_oracle.c -- CPU restatement of the reference ARAP solve + forward warp.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (arap_flow_b200/, include/)
 * may link, import or execute this file.  Callers allowed: tests/, __graft_entry__.smoke(),
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status
 *   - forward warp: PINNED.  Bit-exact against the reference's own warp tool
 *     (ARAP/warping/src/main.cpp, compiled by oracle/Makefile into oracle/_ref/) and against
 *     the shipped golden cat512_wMsk.png (tests/test_oracle_golden.py).
 *   - solve: the reference solver (Terra/Opt generated CUDA) cannot be built or run here
 *     (needs Terra release-2016-03-25 + CUDA 7.5 libdevice), and the tree holds no
 *     per-iteration vectors, so bit-level parity of the solve is UNPINNED.  It is anchored on the
 *     one end-to-end golden the tree ships (cat512_iCstr.txt -> cat512_iFlo.flo, weak because the
 *     fixed-budget GN/PCG trajectory is chaotic on 9 constraints; at energy level the shipped
 *     result and the oracle's agree to 1.5 %: 44.76 vs 45.41, test_solve_energy_pin_vs_shipped_golden)
 *     and on a finite-difference check of J^T F / J^T J p against the residual function
 *     (tests/test_oracle_solver.py).
 *
 * What is restated (all paths relative to /root/reference):
 *   energy                      arap_plan.t:1-23, ARAP/API/src/lib.t:92-96 (Rotate2D)
 *   bounds/validity wrapping    ARAP/API/src/o.t:1895-1936
 *   J^T F + diag(J^T J)         ARAP/API/src/o.t:2129-2172      (createjtfcentered)
 *   J^T J p                     ARAP/API/src/o.t:2029-2089      (createjtjcentered)
 *   cost                        ARAP/API/src/o.t:2375-2385
 *   exclude                     ARAP/API/src/o.t:2452-2465, arap_plan.t:11
 *   guarded invert (CERES)      ARAP/API/src/solverGPUGaussNewton.t:323-332
 *   PCGInit1/Step1/2/3/Update   ARAP/API/src/solverGPUGaussNewton.t:361-397, 421-434, 446-489, 537-557
 *   GN step / init control flow ARAP/API/src/solverGPUGaussNewton.t:956-1007, 1016-1177
 *   continuation, weights,
 *   constraint image, reset     ARAP/shared/CombinedSolverBase.h:99-120,
 *                               ARAP/deformation/src/CombinedSolver.h:172-242
 *   border pins                 ARAP/deformation/src/main.cpp:130-136
 *   flow extraction             ARAP/deformation/src/CombinedSolver.h:352-366
 *   forward warp                ARAP/warping/src/main.cpp:68-104, 110-142, 145-225
 *                               (in-app twin ARAP/deformation/src/CombinedSolver.h:61-97, 248-342)
 *
 * ARITHMETIC CONTRACT (shared with the CUDA kernels).
 * The reference's own fp32 association order (chosen by Opt's simplifier + LLVM) and its atomic
 * summation order are not recoverable, so this file fixes one:
 *   C1  all state and per-pixel arithmetic is IEEE binary32, round-to-nearest, no implicit
 *       contraction (build with -ffp-contract=off); fused multiply-adds appear only where
 *       fmaf() is written.
 *   C2  cos/sin of the angle are evaluated once per Gauss-Newton step by contract_sincos()
 *       (binary64 Cody-Waite + fdlibm kernel polynomials, fma/mul/add only, rounded to binary32),
 *       so CPU and GPU produce identical bits without depending on any libm.
 *   C3  every global dot product / cost sum is the EXACT sum of "group terms", rounded once to
 *       binary32.  A group is an aligned vertical quad (rows 4k..4k+3 of one column); its term is
 *       the binary32 sum, in increasing row order, of the per-pixel terms of its active pixels.
 *       "Exact" is realised by binned binary64 accumulation (exact hi part + tiny lo part); any
 *       implementation / summation order yields the same binary32 result (failure probability
 *       < 2^-40 per reduction).
 *   C4  ||R'(a) d||^2 is taken as ||d||^2 (they are equal for an exact rotation).
 *   C5  neighbour order (+x, -x, +y, -y) (arap_plan.t:14); accumulators start at +0 and skip
 *       invalid neighbours.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ C2: sincos ------------- */
/* fdlibm-style kernels on |r| <= pi/4 (coefficients are the published fdlibm k_sin.c / k_cos.c
 * minimax constants); reduction r = a - k*pi/2 with a two-term Cody-Waite split. */
static const double PIO2_HI = 1.57079632673412561417e+00; /* first 33 bits of pi/2 */
static const double PIO2_LO = 6.07710050650619224932e-11; /* pi/2 - PIO2_HI */
static const double TWO_OVER_PI = 6.36619772367581382433e-01;

ORACLE_API void arap_oracle_sincos(float a, float *s_out, float *c_out)
{
    double x = (double)a;
    double k = rint(x * TWO_OVER_PI);
    double r = fma(-k, PIO2_HI, x);
    r = fma(-k, PIO2_LO, r);
    double z = r * r;
    /* sin(r) = r + r*z*(S1 + z*(S2 + ... S6)) */
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    double sr = fma(r * z, ps, r);
    /* cos(r) = 1 - z/2 + z*z*(C1 + z*(C2 + ... C6)) */
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    long long q = (long long)k;
    double s, c;
    switch ((int)(q & 3)) {
    case 0: s = sr; c = cr; break;
    case 1: s = cr; c = -sr; break;
    case 2: s = -sr; c = -cr; break;
    default: s = -cr; c = sr; break;
    }
    *s_out = (float)s;
    *c_out = (float)c;
}

/* ------------------------------------------------------------------ C3: exact sums --------- */
static int ceil_log2_u64(uint64_t n)
{
    int b = 0;
    while (((uint64_t)1 << b) < n) ++b;
    return b;
}

/* Exact sum of n binary32 terms, rounded to binary32.  Two passes: max, then binning with
 * B = 1.5*2^k chosen so that every partial sum of the hi parts is exactly representable. */
ORACLE_API float arap_oracle_exact_sum(const float *t, size_t n)
{
    float m = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float v = fabsf(t[i]);
        if (v > m) m = v;
    }
    if (!(m > 0.0f)) return 0.0f; /* all zero (NaNs are not expected here) */
    int k = ilogbf(m) + ceil_log2_u64(n ? n : 1) + 2;
    volatile double B = ldexp(1.5, k);
    double H = 0.0, L = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double v = (double)t[i];
        volatile double tmp = B + v;
        double hi = tmp - B;
        double lo = v - hi;
        H += hi;
        L += lo;
    }
    return (float)(H + L);
}

/* ------------------------------------------------------------------ problem view ----------- */
typedef struct {
    int W, H;
    const float *U; /* float2[N]  UrShape */
    const float *C; /* float2[N]  Constraints, (-1,-1) = none */
    const float *M; /* float[N]   Mask, 0 = active */
    float wf, wr, wf2, wr2;
    /* active bounding box in quad rows */
    int x0, x1, y0, y1; /* pixel bbox, inclusive, y0 multiple of 4 */
    int gw, gh;         /* group grid inside the bbox */
} Prob;

static void prob_init(Prob *P, int W, int H, const float *U, const float *C, const float *M,
                      float wf, float wr)
{
    P->W = W; P->H = H; P->U = U; P->C = C; P->M = M;
    P->wf = wf; P->wr = wr; P->wf2 = wf * wf; P->wr2 = wr * wr;
    int x0 = W, x1 = -1, y0 = H, y1 = -1;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (M[(size_t)y * W + x] == 0.0f) {
                if (x < x0) x0 = x;
                if (x > x1) x1 = x;
                if (y < y0) y0 = y;
                if (y > y1) y1 = y;
            }
    if (x1 < 0) { x0 = 0; x1 = -1; y0 = 0; y1 = -1; }
    y0 &= ~3;
    P->x0 = x0; P->x1 = x1; P->y0 = y0; P->y1 = y1;
    P->gw = x1 - x0 + 1; if (P->gw < 0) P->gw = 0;
    P->gh = (y1 >= y0) ? (y1 - y0) / 4 + 1 : 0;
}

static inline int active(const Prob *P, int x, int y)
{
    return P->M[(size_t)y * P->W + x] == 0.0f;
}

/* neighbour offsets in the order of arap_plan.t:14 */
static const int NX[4] = {1, -1, 0, 0};
static const int NY[4] = {0, 0, 1, -1};

static inline int nb_valid(const Prob *P, int x, int y, int n, int *xj, int *yj)
{
    int a = x + NX[n], b = y + NY[n];
    *xj = a; *yj = b;
    if (a < 0 || a >= P->W || b < 0 || b >= P->H) return 0;
    return active(P, a, b);
}

static inline int has_fit(const Prob *P, size_t i)
{
    return P->C[2 * i] >= 0.0f && P->C[2 * i + 1] >= 0.0f; /* arap_plan.t:22 */
}

/* Sum the per-pixel terms T (full-image array, only active entries meaningful) per contract C3 */
static float reduce_terms(const Prob *P, const float *T, float *gbuf)
{
    const int W = P->W, H = P->H;
    if (P->gw == 0) return 0.0f;
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < P->gh; ++gy) {
        for (int gx = 0; gx < P->gw; ++gx) {
            int x = P->x0 + gx;
            float g = 0.0f;
            for (int r = 0; r < 4; ++r) {
                int y = P->y0 + 4 * gy + r;
                if (y < H && active(P, x, y)) g = g + T[(size_t)y * W + x];
            }
            gbuf[(size_t)gy * P->gw + gx] = g;
        }
    }
    return arap_oracle_exact_sum(gbuf, (size_t)P->gw * P->gh);
}

/* ------------------------------------------------------------------ derived functions ------ */
/* cost term of one active pixel (o.t:2375-2385; residuals arap_plan.t:15-23) */
static inline float cost_pixel(const Prob *P, const float *X, const float *cs, int x, int y)
{
    const int W = P->W;
    size_t i = (size_t)y * W + x;
    float ci = cs[2 * i], si = cs[2 * i + 1];
    float acc = 0.0f;
    for (int n = 0; n < 4; ++n) {
        int xj, yj;
        if (!nb_valid(P, x, y, n, &xj, &yj)) continue;
        size_t j = (size_t)yj * W + xj;
        float dx = P->U[2 * i] - P->U[2 * j], dy = P->U[2 * i + 1] - P->U[2 * j + 1];
        float dX0 = X[2 * i] - X[2 * j], dX1 = X[2 * i + 1] - X[2 * j + 1];
        float Ri0 = ci * dx - si * dy, Ri1 = si * dx + ci * dy;
        float e0 = dX0 - Ri0, e1 = dX1 - Ri1;
        float w0 = P->wr * e0, w1 = P->wr * e1;
        acc = fmaf(w0, w0, acc);
        acc = fmaf(w1, w1, acc);
    }
    if (has_fit(P, i)) {
        float f0 = P->wf * (X[2 * i] - P->C[2 * i]);
        float f1 = P->wf * (X[2 * i + 1] - P->C[2 * i + 1]);
        acc = fmaf(f0, f0, acc);
        acc = fmaf(f1, f1, acc);
    }
    return acc;
}

/* J^T F and diag(J^T J) of one active pixel (o.t:2129-2172), closed form SURVEY.md a-3 */
static inline void jtf_pixel(const Prob *P, const float *X, const float *cs, int x, int y,
                             float g[3], float D[2])
{
    const int W = P->W;
    size_t i = (size_t)y * W + x;
    float ci = cs[2 * i], si = cs[2 * i + 1];
    float gx0 = 0.0f, gx1 = 0.0f, ga = 0.0f, nd = 0.0f, nv = 0.0f;
    for (int n = 0; n < 4; ++n) {
        int xj, yj;
        if (!nb_valid(P, x, y, n, &xj, &yj)) continue;
        size_t j = (size_t)yj * W + xj;
        float cj = cs[2 * j], sj = cs[2 * j + 1];
        float dx = P->U[2 * i] - P->U[2 * j], dy = P->U[2 * i + 1] - P->U[2 * j + 1];
        float dX0 = X[2 * i] - X[2 * j], dX1 = X[2 * i + 1] - X[2 * j + 1];
        float Ri0 = ci * dx - si * dy, Ri1 = si * dx + ci * dy;
        float Rj0 = cj * dx - sj * dy, Rj1 = sj * dx + cj * dy;
        float e0 = dX0 - Ri0, e1 = dX1 - Ri1;
        float t0 = (dX0 + dX0) - Ri0; t0 = t0 - Rj0;
        float t1 = (dX1 + dX1) - Ri1; t1 = t1 - Rj1;
        gx0 = gx0 + t0;
        gx1 = gx1 + t1;
        float Q0 = (-(si * dx)) - ci * dy, Q1 = ci * dx - si * dy; /* R'(a_i) d */
        ga = ga + fmaf(Q1, e1, Q0 * e0);
        nd = nd + (dx * dx + dy * dy); /* C4 */
        nv = nv + 1.0f;
    }
    float g0 = P->wr2 * gx0, g1 = P->wr2 * gx1;
    float DX = (P->wr2 + P->wr2) * nv;
    if (has_fit(P, i)) {
        g0 = fmaf(P->wf2, X[2 * i] - P->C[2 * i], g0);
        g1 = fmaf(P->wf2, X[2 * i + 1] - P->C[2 * i + 1], g1);
        DX = DX + P->wf2;
    }
    g[0] = g0; g[1] = g1; g[2] = -(P->wr2 * ga);
    D[0] = DX; D[1] = P->wr2 * nd;
}

/* (J^T J p) of one active pixel (o.t:2029-2089), closed form SURVEY.md a-4, S-form */
static inline void jtj_pixel(const Prob *P, const float *cs, const float *p, int x, int y, float q[3])
{
    const int W = P->W;
    size_t i = (size_t)y * W + x;
    float ci = cs[2 * i], si = cs[2 * i + 1];
    float pi0 = p[3 * i], pi1 = p[3 * i + 1], pai = p[3 * i + 2];
    float sd0 = 0.0f, sd1 = 0.0f, nb0 = 0.0f, nb1 = 0.0f, dd = 0.0f, dc = 0.0f;
    float Sx = 0.0f, Sy = 0.0f, nd = 0.0f;
    for (int n = 0; n < 4; ++n) {
        int xj, yj;
        if (!nb_valid(P, x, y, n, &xj, &yj)) continue;
        size_t j = (size_t)yj * W + xj;
        float cj = cs[2 * j], sj = cs[2 * j + 1];
        float dx = P->U[2 * i] - P->U[2 * j], dy = P->U[2 * i + 1] - P->U[2 * j + 1];
        float dp0 = pi0 - p[3 * j], dp1 = pi1 - p[3 * j + 1];
        float paj = p[3 * j + 2];
        sd0 = sd0 + dp0;
        sd1 = sd1 + dp1;
        float Qj0 = (-(sj * dx)) - cj * dy, Qj1 = cj * dx - sj * dy; /* R'(a_j) d */
        nb0 = nb0 + Qj0 * paj;
        nb1 = nb1 + Qj1 * paj;
        dd = dd + (dx * dp0 + dy * dp1);
        dc = dc + (dx * dp1 - dy * dp0);
        Sx = Sx + dx;
        Sy = Sy + dy;
        nd = nd + (dx * dx + dy * dy);
    }
    float E0 = (-(si * Sx)) - ci * Sy, E1 = ci * Sx - si * Sy; /* R'(a_i) sum_j d_ij */
    float own0 = E0 * pai, own1 = E1 * pai;
    float t0 = (sd0 + sd0) - own0; t0 = t0 - nb0;
    float t1 = (sd1 + sd1) - own1; t1 = t1 - nb1;
    float q0 = P->wr2 * t0, q1 = P->wr2 * t1;
    float rdp = fmaf(ci, dc, -(si * dd)); /* sum_j (R'(a_i) d_ij) . dp_ij */
    float qa = P->wr2 * fmaf(nd, pai, -rdp);
    if (has_fit(P, i)) {
        q0 = fmaf(P->wf2, pi0, q0);
        q1 = fmaf(P->wf2, pi1, q1);
    }
    q[0] = q0; q[1] = q1; q[2] = qa;
}

static inline float guarded_invert(float d) /* solverGPUGaussNewton.t:323-332 (CERES flavour) */
{
    float t = 1.0f + sqrtf(d);
    return 1.0f / (t * t);
}

static inline float dot3(const float *a, const float *b)
{
    return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0]));
}

static void fill_cs(const Prob *P, const float *A, float *cs)
{
    const int W = P->W, H = P->H;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            if (active(P, x, y)) arap_oracle_sincos(A[i], &cs[2 * i + 1], &cs[2 * i]);
            else { cs[2 * i] = 1.0f; cs[2 * i + 1] = 0.0f; }
        }
}

/* ------------------------------------------------------------------ unit-level entry points */
ORACLE_API float arap_oracle_cost(int W, int H, const float *X, const float *A, const float *U,
                                  const float *C, const float *M, float wf, float wr)
{
    Prob P; prob_init(&P, W, H, U, C, M, wf, wr);
    size_t N = (size_t)W * H;
    float *cs = (float *)malloc(2 * N * sizeof(float));
    float *T = (float *)calloc(N, sizeof(float));
    float *gb = (float *)malloc(((size_t)P.gw * P.gh + 1) * sizeof(float));
    fill_cs(&P, A, cs);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (active(&P, x, y)) T[(size_t)y * W + x] = cost_pixel(&P, X, cs, x, y);
    float s = reduce_terms(&P, T, gb);
    free(cs); free(T); free(gb);
    return 0.5f * s;
}

/* r = -J^T F, pre = guardedInvert(diag), both float3[N] (X0,X1,A); inactive entries = 0 */
ORACLE_API void arap_oracle_eval_jtf(int W, int H, const float *X, const float *A, const float *U,
                                     const float *C, const float *M, float wf, float wr,
                                     float *r3, float *pre3)
{
    Prob P; prob_init(&P, W, H, U, C, M, wf, wr);
    size_t N = (size_t)W * H;
    float *cs = (float *)malloc(2 * N * sizeof(float));
    fill_cs(&P, A, cs);
    memset(r3, 0, 3 * N * sizeof(float));
    memset(pre3, 0, 3 * N * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (active(&P, x, y)) {
                size_t i = (size_t)y * W + x;
                float g[3], D[2];
                jtf_pixel(&P, X, cs, x, y, g, D);
                r3[3 * i] = -g[0]; r3[3 * i + 1] = -g[1]; r3[3 * i + 2] = -g[2];
                pre3[3 * i] = pre3[3 * i + 1] = guarded_invert(D[0]);
                pre3[3 * i + 2] = guarded_invert(D[1]);
            }
    free(cs);
}

/* q = J^T J p, float3[N]; returns dot(p,q) per contract C3 */
ORACLE_API float arap_oracle_apply_jtj(int W, int H, const float *A, const float *U, const float *C,
                                       const float *M, float wf, float wr, const float *p3, float *q3)
{
    Prob P; prob_init(&P, W, H, U, C, M, wf, wr);
    size_t N = (size_t)W * H;
    float *cs = (float *)malloc(2 * N * sizeof(float));
    float *T = (float *)calloc(N, sizeof(float));
    float *gb = (float *)malloc(((size_t)P.gw * P.gh + 1) * sizeof(float));
    fill_cs(&P, A, cs);
    memset(q3, 0, 3 * N * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (active(&P, x, y)) {
                size_t i = (size_t)y * W + x;
                jtj_pixel(&P, cs, p3, x, y, &q3[3 * i]);
                T[i] = dot3(&p3[3 * i], &q3[3 * i]);
            }
    float s = reduce_terms(&P, T, gb);
    free(cs); free(T); free(gb);
    return s;
}

/* Raw residual vector for the finite-difference test: 10 residuals per pixel
 * (4 neighbours x 2 comps, then fit x 2), zero where invalid.  Evaluated in double from the
 * definition in arap_plan.t:13-23 using libm sin/cos -- deliberately NOT the closed forms. */
ORACLE_API void arap_oracle_residuals_f64(int W, int H, const double *X, const double *A,
                                          const float *U, const float *C, const float *M,
                                          double wf, double wr, double *res10)
{
    Prob P; prob_init(&P, W, H, U, C, M, (float)wf, (float)wr);
    size_t N = (size_t)W * H;
    memset(res10, 0, 10 * N * sizeof(double));
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (!active(&P, x, y)) continue;
            size_t i = (size_t)y * W + x;
            double c = cos(A[i]), s = sin(A[i]);
            for (int n = 0; n < 4; ++n) {
                int xj, yj;
                if (!nb_valid(&P, x, y, n, &xj, &yj)) continue;
                size_t j = (size_t)yj * W + xj;
                double dx = (double)U[2 * i] - U[2 * j], dy = (double)U[2 * i + 1] - U[2 * j + 1];
                res10[10 * i + 2 * n] = wr * ((X[2 * i] - X[2 * j]) - (c * dx - s * dy));
                res10[10 * i + 2 * n + 1] = wr * ((X[2 * i + 1] - X[2 * j + 1]) - (s * dx + c * dy));
            }
            if (has_fit(&P, i)) {
                res10[10 * i + 8] = wf * (X[2 * i] - C[2 * i]);
                res10[10 * i + 9] = wf * (X[2 * i + 1] - C[2 * i + 1]);
            }
        }
}

/* ------------------------------------------------------------------ GN / PCG --------------- */
typedef struct {
    size_t N;
    float *cs, *r, *p, *q, *pre, *delta, *T, *gb;
} Work;

static int work_alloc(Work *w, const Prob *P)
{
    size_t N = (size_t)P->W * P->H;
    w->N = N;
    w->cs = (float *)malloc(2 * N * sizeof(float));
    w->r = (float *)calloc(3 * N, sizeof(float));
    w->p = (float *)calloc(3 * N, sizeof(float));
    w->q = (float *)calloc(3 * N, sizeof(float));
    w->pre = (float *)calloc(3 * N, sizeof(float));
    w->delta = (float *)calloc(3 * N, sizeof(float));
    w->T = (float *)calloc(N, sizeof(float));
    w->gb = (float *)malloc(((size_t)P->gw * P->gh + 1) * sizeof(float));
    return w->cs && w->r && w->p && w->q && w->pre && w->delta && w->T && w->gb;
}
static void work_free(Work *w)
{
    free(w->cs); free(w->r); free(w->p); free(w->q); free(w->pre); free(w->delta); free(w->T); free(w->gb);
}

static float cost_all(const Prob *P, Work *w, const float *X, const float *A)
{
    const int W = P->W;
    fill_cs(P, A, w->cs);
#pragma omp parallel for schedule(static)
    for (int y = P->y0; y <= P->y1; ++y)
        for (int x = P->x0; x <= P->x1; ++x)
            if (active(P, x, y)) w->T[(size_t)y * W + x] = cost_pixel(P, X, w->cs, x, y);
    return 0.5f * reduce_terms(P, w->T, w->gb);
}

/* Opt-in convergence-aware schedule (NOT reference behaviour, default off): the same two rules the
 * resident CUDA kernel implements, so that the early-exit path is checked against this oracle and not against itself.
 *   pcg_rtol > 0: a PCG loop ends after the iteration whose r.z <= pcg_rtol^2 * (r.z at PCGInit1)
 *   gn_rtol  > 0: the Gauss-Newton steps of one Opt_ProblemSolve end once a step lowers the cost by less than gn_rtol
 *                 (relative); the remaining cost entries repeat the last cost */
static float g_pcg_rtol = 0.0f, g_gn_rtol = 0.0f;
ORACLE_API void arap_oracle_set_rtol(float pcg_rtol, float gn_rtol)
{
    g_pcg_rtol = pcg_rtol > 0.0f ? pcg_rtol : 0.0f;
    g_gn_rtol = gn_rtol > 0.0f ? gn_rtol : 0.0f;
}

/* One Gauss-Newton step = solverGPUGaussNewton.t:1016-1177 with UsesLambda()==false.
 * Optional trace: scal[3*it+0..2] = (den, alpha-numerator used, beta numerator) per PCG iteration. */
static void gn_step(const Prob *P, Work *w, float *X, float *A, int nPCG, float *scal)
{
    const int W = P->W;
    float *r = w->r, *p = w->p, *q = w->q, *pre = w->pre, *delta = w->delta, *T = w->T;
    fill_cs(P, A, w->cs);
    /* PCGInit1 (:361-397) */
#pragma omp parallel for schedule(static)
    for (int y = P->y0; y <= P->y1; ++y)
        for (int x = P->x0; x <= P->x1; ++x)
            if (active(P, x, y)) {
                size_t i = (size_t)y * W + x;
                float g[3], D[2];
                jtf_pixel(P, X, w->cs, x, y, g, D);
                float px = guarded_invert(D[0]), pa = guarded_invert(D[1]);
                pre[3 * i] = px; pre[3 * i + 1] = px; pre[3 * i + 2] = pa;
                for (int k = 0; k < 3; ++k) {
                    r[3 * i + k] = -g[k];
                    p[3 * i + k] = pre[3 * i + k] * r[3 * i + k];
                    delta[3 * i + k] = 0.0f;
                }
                T[i] = dot3(&r[3 * i], &p[3 * i]);
            }
    float num = reduce_terms(P, T, w->gb);
    const float rtol2 = g_pcg_rtol * g_pcg_rtol;
    const float stop = (rtol2 > 0.0f) ? rtol2 * num : -1.0f;
    for (int it = 0; it < nPCG; ++it) {
        /* PCGStep1 (:421-434) */
#pragma omp parallel for schedule(static)
        for (int y = P->y0; y <= P->y1; ++y)
            for (int x = P->x0; x <= P->x1; ++x)
                if (active(P, x, y)) {
                    size_t i = (size_t)y * W + x;
                    jtj_pixel(P, w->cs, p, x, y, &q[3 * i]);
                    T[i] = dot3(&p[3 * i], &q[3 * i]);
                }
        float den = reduce_terms(P, T, w->gb);
        float alpha = (den > 0.0f) ? num / den : 0.0f; /* :456-459 */
        /* PCGStep2 (:446-489); z is not stored, it is recomputed as pre*r in step 3 */
#pragma omp parallel for schedule(static)
        for (int y = P->y0; y <= P->y1; ++y)
            for (int x = P->x0; x <= P->x1; ++x)
                if (active(P, x, y)) {
                    size_t i = (size_t)y * W + x;
                    float z[3];
                    for (int k = 0; k < 3; ++k) {
                        delta[3 * i + k] = fmaf(alpha, p[3 * i + k], delta[3 * i + k]);
                        r[3 * i + k] = fmaf(-alpha, q[3 * i + k], r[3 * i + k]);
                        z[k] = pre[3 * i + k] * r[3 * i + k];
                    }
                    T[i] = dot3(z, &r[3 * i]);
                }
        float bnum = reduce_terms(P, T, w->gb);
        float beta = (num > 0.0f) ? bnum / num : 0.0f; /* :537-550 */
        if (it + 1 < nPCG && bnum <= stop) { /* opt-in early exit: the direction is not needed any more */
            if (scal) { scal[3 * it] = den; scal[3 * it + 1] = num; scal[3 * it + 2] = bnum; }
            break;
        }
#pragma omp parallel for schedule(static)
        for (int y = P->y0; y <= P->y1; ++y)
            for (int x = P->x0; x <= P->x1; ++x)
                if (active(P, x, y)) {
                    size_t i = (size_t)y * W + x;
                    for (int k = 0; k < 3; ++k) {
                        float z = pre[3 * i + k] * r[3 * i + k];
                        p[3 * i + k] = fmaf(beta, p[3 * i + k], z);
                    }
                }
        if (scal) { scal[3 * it] = den; scal[3 * it + 1] = num; scal[3 * it + 2] = bnum; }
        num = bnum; /* :1091 */
    }
    /* PCGLinearUpdate (:552-557) */
#pragma omp parallel for schedule(static)
    for (int y = P->y0; y <= P->y1; ++y)
        for (int x = P->x0; x <= P->x1; ++x)
            if (active(P, x, y)) {
                size_t i = (size_t)y * W + x;
                X[2 * i] = X[2 * i] + delta[3 * i];
                X[2 * i + 1] = X[2 * i + 1] + delta[3 * i + 1];
                A[i] = A[i] + delta[3 * i + 2];
            }
}

/* == Opt_ProblemSolve (o.t:2548-2551): init (cost) then nGN steps, cost after each.
 * X float2[N] and A float[N] are updated in place.  costs[0..nGN] (may be NULL).
 * scal (may be NULL): 3*nPCG floats per GN step. */
ORACLE_API int arap_oracle_gn_solve(int W, int H, float *X, float *A, const float *U, const float *C,
                                    const float *M, float wf, float wr, int nGN, int nPCG,
                                    float *costs, float *scal)
{
    Prob P; prob_init(&P, W, H, U, C, M, wf, wr);
    Work w;
    if (!work_alloc(&w, &P)) return -1;
    float c0 = cost_all(&P, &w, X, A);
    if (costs) costs[0] = c0;
    float prev = c0;
    for (int g = 0; g < nGN; ++g) {
        gn_step(&P, &w, X, A, nPCG, scal ? scal + (size_t)3 * nPCG * g : NULL);
        float c = cost_all(&P, &w, X, A);
        if (costs) costs[g + 1] = c;
        if (g_gn_rtol > 0.0f && g + 1 < nGN && !((prev - c) > g_gn_rtol * prev)) { /* opt-in: this step gained too little */
            if (costs)
                for (int gg = g + 2; gg <= nGN; ++gg) costs[gg] = c;
            break;
        }
        prev = c;
    }
    work_free(&w);
    return 0;
}

/* ------------------------------------------------------------------ Levenberg-Marquardt ---- */
/* The reference's dormant "LMGPU" solver kind: solverGPUGaussNewton.t with UsesLambda() == true.
 * The ARAP app never requests it (CombinedSolverBase.h:76), the tree holds no vectors for it: PARITY UNPINNED, this
 * restatement and the CUDA path (csrc/solver_lm.cu) are checked against each other only.  Same arithmetic contract as
 * the Gauss-Newton path; the additional per-pixel expressions are fixed here:
 *   unclamped CtC_k = D_k * (1 / radius)                         (o.t:2255-2288; D = diag(J^T J) of jtf_pixel)
 *   (J^T J + CtC) v = fmaf(CtC_k, v_k, (J^T J v)_k)              (o.t:2076-2082)
 *   model cost term = sum over the pixel's residuals of (F + J delta)^2, accumulated like cost_pixel (o.t:2174-2202)
 *   Q term          = 0.5 * (delta . (r + b))                    (solverGPUGaussNewton.t:483, :527)
 * Host-side trust-region arithmetic is the reference's: binary32 fields, binary64 where its literals promote. */
typedef struct {
    float min_relative_decrease, min_trust_region_radius, max_trust_region_radius, q_tolerance, function_tolerance,
        trust_region_radius, radius_decrease_factor, min_lm_diagonal, max_lm_diagonal;
    int residual_reset_period;
} LmParams;

static void lm_defaults(LmParams *p) /* solverGPUGaussNewton.t:26-39 */
{
    p->min_relative_decrease = (float)1e-3;
    p->min_trust_region_radius = (float)1e-32;
    p->max_trust_region_radius = (float)1e16;
    p->q_tolerance = (float)0.0001;
    p->function_tolerance = (float)0.000001;
    p->trust_region_radius = (float)1e4;
    p->radius_decrease_factor = (float)2.0;
    p->min_lm_diagonal = (float)1e-6;
    p->max_lm_diagonal = (float)1e32;
    p->residual_reset_period = 10;
}

/* sum over the residuals owned by pixel (x, y) of (F + J delta)^2; delta = float3[N] (dX0, dX1, dA) */
static inline float modelcost_pixel(const Prob *P, const float *X, const float *cs, const float *delta, int x, int y)
{
    const int W = P->W;
    size_t i = (size_t)y * W + x;
    float ci = cs[2 * i], si = cs[2 * i + 1];
    float di0 = delta[3 * i], di1 = delta[3 * i + 1], dai = delta[3 * i + 2];
    float acc = 0.0f;
    for (int n = 0; n < 4; ++n) {
        int xj, yj;
        if (!nb_valid(P, x, y, n, &xj, &yj)) continue;
        size_t j = (size_t)yj * W + xj;
        float dx = P->U[2 * i] - P->U[2 * j], dy = P->U[2 * i + 1] - P->U[2 * j + 1];
        float dX0 = X[2 * i] - X[2 * j], dX1 = X[2 * i + 1] - X[2 * j + 1];
        float Ri0 = ci * dx - si * dy, Ri1 = si * dx + ci * dy;
        float e0 = dX0 - Ri0, e1 = dX1 - Ri1;
        float dd0 = di0 - delta[3 * j], dd1 = di1 - delta[3 * j + 1];
        float Q0 = (-(si * dx)) - ci * dy, Q1 = ci * dx - si * dy; /* R'(a_i) d */
        float m0 = (e0 + dd0) - Q0 * dai, m1 = (e1 + dd1) - Q1 * dai;
        float w0 = P->wr * m0, w1 = P->wr * m1;
        acc = fmaf(w0, w0, acc);
        acc = fmaf(w1, w1, acc);
    }
    if (has_fit(P, i)) {
        float f0 = P->wf * ((X[2 * i] - P->C[2 * i]) + di0);
        float f1 = P->wf * ((X[2 * i + 1] - P->C[2 * i + 1]) + di1);
        acc = fmaf(f0, f0, acc);
        acc = fmaf(f1, f1, acc);
    }
    return acc;
}

typedef struct {
    float *ctc, *b, *ssq, *adelta, *prevX, *prevA, *T2;
} LmWork;

/* out = (J^T J + CtC) v for every active pixel */
static void lm_apply(const Prob *P, Work *w, const LmWork *lw, const float *v, float *out)
{
    const int W = P->W;
#pragma omp parallel for schedule(static)
    for (int y = P->y0; y <= P->y1; ++y)
        for (int x = P->x0; x <= P->x1; ++x)
            if (active(P, x, y)) {
                size_t i = (size_t)y * W + x;
                float q[3];
                jtj_pixel(P, w->cs, v, x, y, q);
                for (int k = 0; k < 3; ++k) out[3 * i + k] = fmaf(lw->ctc[3 * i + k], v[3 * i + k], q[k]);
            }
}

/* One LM iteration = solverGPUGaussNewton.t:1016-1177 with UsesLambda() == true.  Returns 1 to continue, 0 when the
 * solver decided to stop (function tolerance / minimum radius).  stat[0..5] = radius after the step, PCG iterations
 * run, verdict (1 accepted, 0 reverted, 2 function tolerance reached, 3 radius below minimum), model cost, new cost,
 * last Q. */
static int lm_step(const Prob *P, Work *w, LmWork *lw, LmParams *lp, float *X, float *A, int nPCG, int first,
                   float *prevCost, float *stat)
{
    const int W = P->W;
    float *r = w->r, *p = w->p, *q = w->q, *pre = w->pre, *delta = w->delta, *T = w->T, *T2 = lw->T2;
    const float radius = lp->trust_region_radius;
    const float inv_radius = 1.0f / radius;
    fill_cs(P, A, w->cs);
    /* PCGInit1 (:361-397), PCGSaveSSq (:624-629), PCGComputeCtC (:616-622), PCGFinalizeDiagonal (:631-664) */
#pragma omp parallel for schedule(static)
    for (int y = P->y0; y <= P->y1; ++y)
        for (int x = P->x0; x <= P->x1; ++x)
            if (active(P, x, y)) {
                size_t i = (size_t)y * W + x;
                float g[3], D2[2];
                jtf_pixel(P, X, w->cs, x, y, g, D2);
                const float D[3] = {D2[0], D2[0], D2[1]};
                for (int k = 0; k < 3; ++k) {
                    if (first) lw->ssq[3 * i + k] = guarded_invert(D[k]); /* jacobiScaling ONCE_PER_SOLVE */
                    float u = D[k] * inv_radius;
                    float invS = 1.0f / lw->ssq[3 * i + k];
                    float mult = invS / radius;
                    float lo = lp->min_lm_diagonal * mult, hi = lp->max_lm_diagonal * mult;
                    float c = fminf(fmaxf(u, lo), hi);
                    lw->ctc[3 * i + k] = c;
                    pre[3 * i + k] = 1.0f / (c + radius * u);
                    r[3 * i + k] = -g[k];
                    lw->b[3 * i + k] = r[3 * i + k];
                    p[3 * i + k] = pre[3 * i + k] * r[3 * i + k];
                    delta[3 * i + k] = 0.0f;
                }
                T[i] = dot3(&r[3 * i], &p[3 * i]);
                float rr[3] = {r[3 * i] + r[3 * i], r[3 * i + 1] + r[3 * i + 1], r[3 * i + 2] + r[3 * i + 2]};
                T2[i] = 0.5f * dot3(&delta[3 * i], rr);
            }
    float num = reduce_terms(P, T, w->gb);
    float Q0 = reduce_terms(P, T2, w->gb);
    float Q1 = Q0;
    int it = 0;
    for (; it < nPCG;) {
        lm_apply(P, w, lw, p, q); /* PCGStep1 (:421-434) */
#pragma omp parallel for schedule(static)
        for (int y = P->y0; y <= P->y1; ++y)
            for (int x = P->x0; x <= P->x1; ++x)
                if (active(P, x, y)) {
                    size_t i = (size_t)y * W + x;
                    T[i] = dot3(&p[3 * i], &q[3 * i]);
                }
        float den = reduce_terms(P, T, w->gb);
        float alpha = (den > 0.0f) ? num / den : 0.0f;
        const int reset = ((it + 1) % lp->residual_reset_period) == 0; /* :1077-1086 */
        if (reset) {
#pragma omp parallel for schedule(static)
            for (int y = P->y0; y <= P->y1; ++y)
                for (int x = P->x0; x <= P->x1; ++x)
                    if (active(P, x, y)) {
                        size_t i = (size_t)y * W + x;
                        for (int k = 0; k < 3; ++k) delta[3 * i + k] = fmaf(alpha, p[3 * i + k], delta[3 * i + k]);
                    }
            lm_apply(P, w, lw, delta, lw->adelta); /* computeAdelta (:566-571) */
        }
#pragma omp parallel for schedule(static)
        for (int y = P->y0; y <= P->y1; ++y)
            for (int x = P->x0; x <= P->x1; ++x)
                if (active(P, x, y)) {
                    size_t i = (size_t)y * W + x;
                    float z[3], rb[3];
                    for (int k = 0; k < 3; ++k) {
                        if (reset) {
                            r[3 * i + k] = lw->b[3 * i + k] - lw->adelta[3 * i + k]; /* PCGStep2_2ndHalf (:505-534) */
                        } else {
                            delta[3 * i + k] = fmaf(alpha, p[3 * i + k], delta[3 * i + k]); /* PCGStep2 (:446-489) */
                            r[3 * i + k] = fmaf(-alpha, q[3 * i + k], r[3 * i + k]);
                        }
                        z[k] = pre[3 * i + k] * r[3 * i + k];
                        rb[k] = r[3 * i + k] + lw->b[3 * i + k];
                    }
                    T[i] = dot3(z, &r[3 * i]);
                    T2[i] = 0.5f * dot3(&delta[3 * i], rb);
                }
        float bnum = reduce_terms(P, T, w->gb);
        Q1 = reduce_terms(P, T2, w->gb);
        float beta = (num > 0.0f) ? bnum / num : 0.0f;
#pragma omp parallel for schedule(static)
        for (int y = P->y0; y <= P->y1; ++y)
            for (int x = P->x0; x <= P->x1; ++x)
                if (active(P, x, y)) {
                    size_t i = (size_t)y * W + x;
                    for (int k = 0; k < 3; ++k) {
                        float z = pre[3 * i + k] * r[3 * i + k];
                        p[3 * i + k] = fmaf(beta, p[3 * i + k], z);
                    }
                }
        num = bnum;
        ++it;
        float zeta = ((float)it * (Q1 - Q0)) / Q1; /* :1093-1101 */
        if (zeta < lp->q_tolerance) break;
        Q0 = Q1;
    }
    /* computeModelCostChange (:818-827), before the update */
#pragma omp parallel for schedule(static)
    for (int y = P->y0; y <= P->y1; ++y)
        for (int x = P->x0; x <= P->x1; ++x)
            if (active(P, x, y)) T[(size_t)y * W + x] = modelcost_pixel(P, X, w->cs, delta, x, y);
    float model_cost = 0.5f * reduce_terms(P, T, w->gb);
    float model_cost_change = *prevCost - model_cost;
    /* savePreviousUnknowns (:573-578), PCGLinearUpdate (:552-557) */
#pragma omp parallel for schedule(static)
    for (int y = P->y0; y <= P->y1; ++y)
        for (int x = P->x0; x <= P->x1; ++x)
            if (active(P, x, y)) {
                size_t i = (size_t)y * W + x;
                lw->prevX[2 * i] = X[2 * i]; lw->prevX[2 * i + 1] = X[2 * i + 1]; lw->prevA[i] = A[i];
                X[2 * i] = X[2 * i] + delta[3 * i];
                X[2 * i + 1] = X[2 * i + 1] + delta[3 * i + 1];
                A[i] = A[i] + delta[3 * i + 2];
            }
    float newCost = cost_all(P, w, X, A);
    stat[1] = (float)it; stat[3] = model_cost; stat[4] = newCost; stat[5] = Q1;
    float cost_change = *prevCost - newCost;
    float relative_decrease = cost_change / model_cost_change;
    if (cost_change >= 0.0f && relative_decrease > lp->min_relative_decrease) { /* :1127-1142 */
        float absolute_function_tolerance = *prevCost * lp->function_tolerance;
        if (cost_change <= absolute_function_tolerance) {
            stat[0] = lp->trust_region_radius; stat[2] = 2.0f;
            return 0; /* the update stays, prevCost does not move: the reference returns before both */
        }
        double step_quality = (double)relative_decrease;
        double min_factor = 1.0 / 3.0;
        double tmp_factor = 1.0 - pow(2.0 * step_quality - 1.0, 3.0);
        lp->trust_region_radius = (float)((double)lp->trust_region_radius / fmax(min_factor, tmp_factor));
        lp->trust_region_radius = (float)fmin((double)lp->trust_region_radius, (double)lp->max_trust_region_radius);
        lp->radius_decrease_factor = 2.0f;
        *prevCost = newCost;
        stat[2] = 1.0f;
    } else { /* :1143-1155 */
#pragma omp parallel for schedule(static)
        for (int y = P->y0; y <= P->y1; ++y)
            for (int x = P->x0; x <= P->x1; ++x)
                if (active(P, x, y)) {
                    size_t i = (size_t)y * W + x;
                    X[2 * i] = lw->prevX[2 * i]; X[2 * i + 1] = lw->prevX[2 * i + 1]; A[i] = lw->prevA[i];
                }
        lp->trust_region_radius = lp->trust_region_radius / lp->radius_decrease_factor;
        lp->radius_decrease_factor = (float)(2.0 * (double)lp->radius_decrease_factor);
        if (lp->trust_region_radius <= lp->min_trust_region_radius) {
            stat[0] = lp->trust_region_radius; stat[2] = 3.0f;
            return 0;
        }
        stat[2] = 0.0f;
    }
    stat[0] = lp->trust_region_radius;
    return 1;
}

/* == Opt_ProblemSolve on an "LMGPU" plan: init (:956-1007: parameters from the solver parameters, prevCost) then
 * steps until one returns 0 or nGN were taken.  params10 (may be NULL = defaults): the nine float parameters in the
 * order of LmParams, then residual_reset_period as a float.  costs[0..nGN] = prevCost after init and after every step
 * (entries of steps not taken repeat the last one); stats (may be NULL) = 6 floats per step taken.
 * Returns the number of steps taken, or -1. */
ORACLE_API int arap_oracle_lm_solve(int W, int H, float *X, float *A, const float *U, const float *C, const float *M,
                                    float wf, float wr, int nGN, int nPCG, const float *params10, float *costs,
                                    float *stats)
{
    Prob P; prob_init(&P, W, H, U, C, M, wf, wr);
    Work w;
    if (!work_alloc(&w, &P)) return -1;
    size_t N = (size_t)W * H;
    LmWork lw;
    lw.ctc = (float *)calloc(3 * N, sizeof(float));
    lw.b = (float *)calloc(3 * N, sizeof(float));
    lw.ssq = (float *)calloc(3 * N, sizeof(float));
    lw.adelta = (float *)calloc(3 * N, sizeof(float));
    lw.prevX = (float *)calloc(2 * N, sizeof(float));
    lw.prevA = (float *)calloc(N, sizeof(float));
    lw.T2 = (float *)calloc(N, sizeof(float));
    if (!lw.ctc || !lw.b || !lw.ssq || !lw.adelta || !lw.prevX || !lw.prevA || !lw.T2) return -1;
    LmParams lp;
    lm_defaults(&lp);
    if (params10) {
        lp.min_relative_decrease = params10[0]; lp.min_trust_region_radius = params10[1];
        lp.max_trust_region_radius = params10[2]; lp.q_tolerance = params10[3]; lp.function_tolerance = params10[4];
        lp.trust_region_radius = params10[5]; lp.radius_decrease_factor = params10[6];
        lp.min_lm_diagonal = params10[7]; lp.max_lm_diagonal = params10[8];
        lp.residual_reset_period = (int)params10[9];
        if (lp.residual_reset_period < 1) lp.residual_reset_period = 1;
    }
    float prevCost = cost_all(&P, &w, X, A);
    if (costs) costs[0] = prevCost;
    int taken = 0;
    for (int g = 0; g < nGN; ++g) {
        float st[6] = {0, 0, 0, 0, 0, 0};
        int more = lm_step(&P, &w, &lw, &lp, X, A, nPCG, g == 0, &prevCost, st);
        if (stats) memcpy(stats + 6 * (size_t)g, st, sizeof(st));
        if (costs) costs[g + 1] = prevCost;
        ++taken;
        if (!more) {
            if (costs) for (int gg = g + 2; gg <= nGN; ++gg) costs[gg] = prevCost;
            break;
        }
    }
    free(lw.ctc); free(lw.b); free(lw.ssq); free(lw.adelta); free(lw.prevX); free(lw.prevA); free(lw.T2);
    work_free(&w);
    return taken;
}

/* constraint image for continuation weight alpha (CombinedSolver.h:223-242) */
ORACLE_API void arap_oracle_constraint_image(int W, int H, const uint8_t *mask_red, const int *matches,
                                             int n_matches, float alpha, float *C)
{
    size_t N = (size_t)W * H;
    for (size_t i = 0; i < N; ++i) { C[2 * i] = -1.0f; C[2 * i + 1] = -1.0f; }
    for (int k = 0; k < n_matches; ++k) {
        int x = matches[4 * k], y = matches[4 * k + 1];
        if (x < 0 || x >= W || y < 0 || y >= H) continue; /* the reference would read out of bounds */
        if (mask_red[(size_t)y * W + x] == 0) {
            float nx = (1.0f - alpha) * (float)x + alpha * (float)matches[4 * k + 2];
            float ny = (1.0f - alpha) * (float)y + alpha * (float)matches[4 * k + 3];
            C[2 * ((size_t)y * W + x)] = nx;
            C[2 * ((size_t)y * W + x) + 1] = ny;
        }
    }
}

/* number of border pins main.cpp:130-136 appends */
ORACLE_API int arap_oracle_border_pin_count(int W, int H)
{
    int n = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (y == 0 || x == 0 || y == H - 1 || x == W - 1) ++n;
    return n;
}

/* The whole per-image path of arap_deform: border pins (main.cpp:130-136), resetGPU
 * (CombinedSolver.h:207-221), weights (:172-177), continuation (CombinedSolverBase.h:108-117).
 * X_out float2[N] (absolute positions), A_out float[N], costs[nCont*(nGN+1)] (may be NULL). */
ORACLE_API int arap_oracle_solve(int W, int H, const uint8_t *mask_red, const int *matches, int n_matches,
                                 int nCont, int nGN, int nPCG, float *X_out, float *A_out, float *costs)
{
    size_t N = (size_t)W * H;
    int nb = arap_oracle_border_pin_count(W, H);
    int *all = (int *)malloc((size_t)4 * (n_matches + nb) * sizeof(int));
    float *U = (float *)malloc(2 * N * sizeof(float));
    float *M = (float *)malloc(N * sizeof(float));
    float *C = (float *)malloc(2 * N * sizeof(float));
    if (!all || !U || !M || !C) return -1;
    memcpy(all, matches, (size_t)4 * n_matches * sizeof(int));
    int k = n_matches;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (y == 0 || x == 0 || y == H - 1 || x == W - 1) {
                all[4 * k] = x; all[4 * k + 1] = y; all[4 * k + 2] = x; all[4 * k + 3] = y; ++k;
            }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            U[2 * i] = (float)x; U[2 * i + 1] = (float)y;
            X_out[2 * i] = (float)x; X_out[2 * i + 1] = (float)y;
            A_out[i] = 0.0f;
            M[i] = (float)mask_red[i];
        }
    float wf = sqrtf(100.0f), wr = sqrtf(0.01f);
    int rc = 0;
    for (int t = 0; t < nCont && rc == 0; ++t) {
        float alpha = (float)(t + 1) / (float)nCont;
        arap_oracle_constraint_image(W, H, mask_red, all, k, alpha, C);
        rc = arap_oracle_gn_solve(W, H, X_out, A_out, U, C, M, wf, wr, nGN, nPCG,
                                  costs ? costs + (size_t)t * (nGN + 1) : NULL, NULL);
    }
    free(all); free(U); free(M); free(C);
    return rc;
}

/* flow = X - grid (CombinedSolver.h:352-366) */
ORACLE_API void arap_oracle_flow(int W, int H, const float *X, float *flow)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            flow[2 * i] = X[2 * i] - (float)x;
            flow[2 * i + 1] = X[2 * i + 1] - (float)y;
        }
}

/* ------------------------------------------------------------------ forward warp ----------- */
/* warping/src/main.cpp:68-104 with w0=w1=w2=1 */
static inline int point_in_triangle_lk(float x0, float y0, float x1, float y1, float x2, float y2,
                                       float sx, float sy, float *b0, float *b1, float *b2)
{
    float X0 = x0 - sx * 1.0f, X1 = x1 - sx * 1.0f, X2 = x2 - sx * 1.0f;
    float Y0 = y0 - sy * 1.0f, Y1 = y1 - sy * 1.0f, Y2 = y2 - sy * 1.0f;
    float d01 = X0 * Y1 - Y0 * X1;
    float d12 = X1 * Y2 - Y1 * X2;
    float d20 = X2 * Y0 - Y2 * X0;
    if ((d01 < 0) & (d12 < 0) & (d20 < 0)) return 0;
    float inv = 1.f / (d01 + d12 + d20);
    d01 *= inv; d12 *= inv; d20 *= inv;
    *b0 = d12; *b1 = d20; *b2 = d01;
    return (d01 >= 0 && d12 >= 0 && d20 >= 0);
}

static void raster_tri(int W, int H, const float *pa, const float *pb, const float *pc,
                       const uint8_t *ca, const uint8_t *cb, const uint8_t *cc, uint32_t id,
                       uint8_t *out_rgb, uint8_t *out_mask, uint32_t *splat)
{
    /* warping/src/main.cpp:110-142 */
    float minx = floorf(fminf(pa[0], fminf(pb[0], pc[0])));
    float miny = floorf(fminf(pa[1], fminf(pb[1], pc[1])));
    float maxx = ceilf(fmaxf(pa[0], fmaxf(pb[0], pc[0])));
    float maxy = ceilf(fmaxf(pa[1], fmaxf(pb[1], pc[1])));
    /* guard the int conversion (the reference has UB for |coords| > 2^31); clip is equivalent */
    if (!(minx > -1e9f)) minx = -1e9f;
    if (!(miny > -1e9f)) miny = -1e9f;
    if (!(maxx < 1e9f)) maxx = 1e9f;
    if (!(maxy < 1e9f)) maxy = 1e9f;
    int xs = (int)minx, ys = (int)miny;
    if (xs < 0) xs = 0;
    if (ys < 0) ys = 0;
    for (int x = xs; x <= maxx && x < W; ++x)
        for (int y = ys; y <= maxy && y < H; ++y) {
            float b0, b1, b2;
            if (point_in_triangle_lk(pa[0], pa[1], pb[0], pb[1], pc[0], pc[1], (float)x, (float)y,
                                     &b0, &b1, &b2)) {
                size_t o = (size_t)y * W + x;
                for (int k = 0; k < 3; ++k) {
                    float v = (float)ca[k] * b0 + (float)cb[k] * b1 + (float)cc[k] * b2;
                    out_rgb[3 * o + k] = (uint8_t)(int)v; /* vec3f -> vec3uc cast, vec3.h:32-37 */
                }
                out_mask[o] = 255;
                splat[o] = id;
            }
        }
}

/* Sequential last-writer-wins rasteriser (main.cpp:145-225).  pos float2[N], rgb uint8x3[N],
 * mask_red uint8[N]; outputs out_rgb uint8x3[N], out_mask uint8[N] (0/255),
 * splat uint32[N] = 1 + 2*(y*W+x) + t of the winning triangle, 0 = nothing landed. */
ORACLE_API void arap_oracle_warp(int W, int H, const float *pos, const uint8_t *rgb, const uint8_t *mask_red,
                                 uint8_t *out_rgb, uint8_t *out_mask, uint32_t *splat)
{
    size_t N = (size_t)W * H;
    memset(out_rgb, 0, 3 * N);
    memset(out_mask, 0, N);
    memset(splat, 0, N * sizeof(uint32_t));
    for (int y = 0; y + 1 < H; ++y)
        for (int x = 0; x + 1 < W; ++x) {
            size_t i00 = (size_t)y * W + x, i01 = i00 + 1, i10 = i00 + W, i11 = i10 + 1;
            if (mask_red[i00] || mask_red[i01] || mask_red[i10] || mask_red[i11]) continue;
            uint32_t id = (uint32_t)(2 * i00 + 1);
            raster_tri(W, H, &pos[2 * i00], &pos[2 * i01], &pos[2 * i10], &rgb[3 * i00], &rgb[3 * i01],
                       &rgb[3 * i10], id, out_rgb, out_mask, splat);
            raster_tri(W, H, &pos[2 * i10], &pos[2 * i01], &pos[2 * i11], &rgb[3 * i10], &rgb[3 * i01],
                       &rgb[3 * i11], id + 1, out_rgb, out_mask, splat);
        }
}

/* warp tool front half: positions from a flow field (main.cpp:160-166) */
ORACLE_API void arap_oracle_flow_to_pos(int W, int H, const float *flow, float *pos)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            pos[2 * i] = (float)x + flow[2 * i];
            pos[2 * i + 1] = (float)y + flow[2 * i + 1];
        }
}

ORACLE_API int arap_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
