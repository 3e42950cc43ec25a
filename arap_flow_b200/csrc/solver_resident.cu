// solver_resident.cu -- persistent, on-chip-resident Gauss-Newton / PCG solver (see solver_resident.cuh).
//
// Replaces the whole host loop of ARAP/API/src/solverGPUGaussNewton.t:1016-1177 (and, in lerp mode, the
// continuation loop of ARAP/shared/CombinedSolverBase.h:99-120 with CombinedSolver.h:223-242) by ONE
// kernel launch.  Arithmetic is the contract of DESIGN.md section 3, i.e. bit-identical to the streaming
// back-end and to oracle/arap_oracle.c.
//
// Decomposition: the active part of the image is cut into 32x8-pixel strips; a warp owns one strip, a lane
// one column of 8 pixels (two contract-C3 groups).  r, delta, p_angle, cos/sin and the neighbour flags live
// in registers for the whole PCG loop; (p_x, p_y, sin*p_a, cos*p_a) of every pixel sits in a shared-memory
// tile with a one-pixel ring.  Strips of the same CTA read each other's tiles directly; strips of other
// CTAs exchange their boundary through a 2.5 KB global "outbox" per strip, piggy-backed on the grid barrier.
//
// Per PCG iteration there are exactly two grid-wide barriers, each carrying the exact (h, l) partial sum of
// the dot product (sum p.Ap, then sum z.r).  The direction update p = z + beta p that follows the second
// barrier needs the neighbours' NEW p; instead of a third barrier, every strip publishes (z, p_old) of its
// boundary pixels before the second barrier and the receiver applies beta itself.
#include "solver_resident.cuh"
#include "grid_math.cuh"
#include "solver_stream.cuh" // FLAG_* bit layout

namespace arapb200 {
namespace {

constexpr int TW = RS_STRIP_W + 2, TH = RS_STRIP_H + 2;
constexpr int RS_THREADS_MAX = 384;
constexpr unsigned long long SENTINEL = 0xFFFFFFFFFFFFFFFFull;

struct __align__(16) StripSmem {
    float4 T[TH][TW];                  // tile + ring
    float2 rcs[RS_OUTBOX_ENTRIES];     // cos/sin of the ring pixels (remote sides only)
};

struct Ctl {
    double red[64];
    float bc;
    int abort;
    float preX[10];
    float preA[5];
};

__device__ __forceinline__ double2 ld_slot(const double2* p)
{
    double2 v;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_slot(double2* p, double a, double b)
{
    asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float4* p) { return __ldcg(p); }

// Everything a warp needs to know about its strip
struct StripCtx {
    bool has;                 // this warp owns a strip
    int slot;                 // global strip id
    int x, y0;                // lane's column, first row
    float4* own;              // &T[0][0] of the own tile
    const float4* up_row;     // row y = -1, indexed by lane
    const float4* down_row;   // row y = 8
    const float4* lptr;       // pixel to the left of (lane, row 0); row stride TW
    const float4* rptr;       // pixel to the right
    float2* rcs;
    int rem[4];               // remote neighbour strip id per side (0 up, 1 down, 2 left, 3 right) or -1
    float4* outbox;           // own outbox
};

struct Cta {
    const ResProb* P;
    Ctl* ctl;
    int cta, G, lane, wid, nw;
    unsigned epoch;
    // optional cycle accounting (thread 0 only): [0] arrive skew, [1] publish+poll, [2] fold+broadcast
    unsigned long long acc[3];
    bool prof;
};

// ---- grid barrier carrying an exact sum ----------------------------------------------------------
// g0/g1: this thread's two group terms.  Returns the exact sum over the whole problem, rounded to
// binary32; *ok = false when the watchdog fired (every thread of every CTA then leaves the kernel).
__device__ __noinline__ float grid_sum(Cta& c, float g0, float g1, bool& ok)
{
    Ctl* ctl = c.ctl;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (c.prof) t0 = clock64();
    // warp: exact sum of 64 terms
    float m = warp_max(fmaxf(fabsf(g0), fabsf(g1)));
    const double B = bin_base(ilogb_f32(m) + 6 + 2);
    double h0, l0, h1, l1;
    bin_split(B, (double)g0, h0, l0);
    bin_split(B, (double)g1, h1, l1);
    double hs = warp_sum(__dadd_rn(h0, h1));
    double ls = warp_sum(__dadd_rn(l0, l1));
    if (c.lane == 0) {
        ctl->red[c.wid] = hs;
        ctl->red[32 + c.wid] = ls;
    }
    __syncthreads();
    if (c.prof) t1 = clock64();
    if (c.wid == 0) {
        HL v;
        v.h = (c.lane < c.nw) ? ctl->red[c.lane] : 0.0;
        v.l = (c.lane < c.nw) ? ctl->red[32 + c.lane] : 0.0;
        HL cta_sum = warp_combine(v);
        const ResProb& P = *c.P;
        double2* buf = P.slots + (size_t)(c.epoch % 3u) * c.G;
        double2* nxt = P.slots + (size_t)((c.epoch + 1u) % 3u) * c.G;
        if (c.lane == 0) {
            // recycle next epoch's slot first: whoever sees this epoch's value also sees the reset
            st_slot(nxt + c.cta, __longlong_as_double((long long)SENTINEL), __longlong_as_double((long long)SENTINEL));
            __threadfence();
            st_slot(buf + c.cta, cta_sum.h, cta_sum.l);
        }
        // gather every CTA's partial (<= 5 per lane)
        double2 v5[5];
        unsigned spins = 0;
        bool done;
        do {
            done = true;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int i = c.lane + 32 * j;
                if (i < c.G) {
                    v5[j] = ld_slot(buf + i);
                    if ((unsigned long long)__double_as_longlong(v5[j].x) == SENTINEL ||
                        (unsigned long long)__double_as_longlong(v5[j].y) == SENTINEL)
                        done = false;
                } else {
                    v5[j] = make_double2(0.0, 0.0);
                }
            }
            done = __all_sync(0xffffffffu, done);
            if (!done && ((++spins & 0x3ffu) == 0)) {
                // watchdog: ~4M polls (seconds) or a peer's abort => bail out instead of hanging the GPU
                int ab = *(volatile int*)P.status;
                if (ab || spins > (1u << 22)) {
                    if (c.lane == 0) {
                        atomicExch(P.status, 1);
                        atomicCAS(P.status + 1, 0, 100 + (int)(c.epoch & 0xffff));
                    }
                    ctl->abort = 1;
                    done = true;
                }
            }
        } while (!done);
        if (c.prof) t2 = clock64();
        __threadfence();
        // exact fold of the G partials, identical in every CTA
        double mm = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j) mm = fmax(mm, fabs(v5[j].x));
        mm = warp_max(mm);
        const double B2 = bin_base(ilogb_f64(mm) + 8 + 2); // <= 160 partials
        double H = 0.0, L = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            double hi, lo;
            bin_split(B2, v5[j].x, hi, lo);
            H = __dadd_rn(H, hi);
            L = __dadd_rn(L, __dadd_rn(lo, v5[j].y));
        }
        H = warp_sum(H);
        L = warp_sum(L);
        if (c.lane == 0) ctl->bc = (float)__dadd_rn(H, L);
    }
    __syncthreads();
    if (c.prof && threadIdx.x == 0) {
        const long long t3 = clock64();
        c.acc[0] += (unsigned long long)(t1 - t0);
        c.acc[1] += (unsigned long long)(t2 - t1);
        c.acc[2] += (unsigned long long)(t3 - t2);
    }
    ++c.epoch;
    ok = (ctl->abort == 0);
    return ctl->bc;
}

// ---- halo publication / reception ----------------------------------------------------------------
// Entry layout in the outbox: two float4 per boundary pixel.
__device__ __forceinline__ void publish_rowcol(const StripCtx& s, int lane, int k, bool top, bool bottom, bool left,
                                               bool right, float4 v0, float4 v1, bool two)
{
    // called once per row k (fully unrolled); v0/v1 are the entry of pixel (lane, k)
    if (top && k == 0) {
        s.outbox[2 * lane] = v0;
        if (two) s.outbox[2 * lane + 1] = v1;
    }
    if (bottom && k == RS_STRIP_H - 1) {
        s.outbox[2 * (32 + lane)] = v0;
        if (two) s.outbox[2 * (32 + lane) + 1] = v1;
    }
    if (left && lane == 0) {
        s.outbox[2 * (64 + k)] = v0;
        if (two) s.outbox[2 * (64 + k) + 1] = v1;
    }
    if (right && lane == 31) {
        s.outbox[2 * (72 + k)] = v0;
        if (two) s.outbox[2 * (72 + k) + 1] = v1;
    }
}

__device__ __forceinline__ const float4* remote_entry(const ResProb& P, int nslot, int e)
{
    return P.outbox + ((size_t)nslot * RS_OUTBOX_ENTRIES + e) * 2;
}

// ring <- neighbours' (X_x, X_y, cos, sin)
__device__ __forceinline__ void recv_x(const ResProb& P, const StripCtx& s, int lane)
{
    if (s.rem[0] >= 0) { // up neighbour: its bottom row
        float4 v = ldcg4(remote_entry(P, s.rem[0], 32 + lane));
        s.own[0 * TW + lane + 1] = v;
        s.rcs[lane] = make_float2(v.z, v.w);
    }
    if (s.rem[1] >= 0) { // down neighbour: its top row
        float4 v = ldcg4(remote_entry(P, s.rem[1], lane));
        s.own[(TH - 1) * TW + lane + 1] = v;
        s.rcs[32 + lane] = make_float2(v.z, v.w);
    }
    if (s.rem[2] >= 0 && lane < 8) { // left neighbour: its right column
        float4 v = ldcg4(remote_entry(P, s.rem[2], 72 + lane));
        s.own[(lane + 1) * TW + 0] = v;
        s.rcs[64 + lane] = make_float2(v.z, v.w);
    }
    if (s.rem[3] >= 0 && lane >= 8 && lane < 16) { // right neighbour: its left column
        float4 v = ldcg4(remote_entry(P, s.rem[3], 64 + lane - 8));
        s.own[(lane - 8 + 1) * TW + TW - 1] = v;
        s.rcs[72 + lane - 8] = make_float2(v.z, v.w);
    }
}

__device__ __forceinline__ float4 p_entry_from(const float4 v0, const float4 v1, float beta, float2 cs)
{
    // v0 = (z0, z1, z2, p0_old), v1 = (p1_old, pa_old, -, -); cs = (cos, sin)
    const float p0 = fmaf(beta, v0.w, v0.x);
    const float p1 = fmaf(beta, v1.x, v0.y);
    const float pa = fmaf(beta, v1.y, v0.z);
    return make_float4(p0, p1, cs.y * pa, cs.x * pa);
}

// ring <- neighbours' new direction, computed from their published (z, p_old)
__device__ __forceinline__ void recv_p(const ResProb& P, const StripCtx& s, int lane, float beta)
{
    if (s.rem[0] >= 0) {
        const float4* e = remote_entry(P, s.rem[0], 32 + lane);
        s.own[0 * TW + lane + 1] = p_entry_from(ldcg4(e), ldcg4(e + 1), beta, s.rcs[lane]);
    }
    if (s.rem[1] >= 0) {
        const float4* e = remote_entry(P, s.rem[1], lane);
        s.own[(TH - 1) * TW + lane + 1] = p_entry_from(ldcg4(e), ldcg4(e + 1), beta, s.rcs[32 + lane]);
    }
    if (s.rem[2] >= 0 && lane < 8) {
        const float4* e = remote_entry(P, s.rem[2], 72 + lane);
        s.own[(lane + 1) * TW + 0] = p_entry_from(ldcg4(e), ldcg4(e + 1), beta, s.rcs[64 + lane]);
    }
    if (s.rem[3] >= 0 && lane >= 8 && lane < 16) {
        const float4* e = remote_entry(P, s.rem[3], 64 + lane - 8);
        s.own[(lane - 8 + 1) * TW + TW - 1] = p_entry_from(ldcg4(e), ldcg4(e + 1), beta, s.rcs[72 + lane - 8]);
    }
}

// constraint of a pixel for the current continuation weight (CombinedSolver.h:236-239)
__device__ __forceinline__ float2 constraint_of(const ResProb& P, size_t i, int x, int y, float alpha)
{
    float2 c = P.C[i];
    if (P.lerp_mode) {
        const float om = 1.0f - alpha;
        c.x = om * (float)x + alpha * c.x;
        c.y = om * (float)y + alpha * c.y;
    }
    return c;
}

// =====================================================================================================
__global__ void __launch_bounds__(RS_THREADS_MAX, 1) k_resident(const ResProb* __restrict__ probs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Ctl ctl;
    const ResProb& P = probs[blockIdx.y];
    if ((int)blockIdx.x >= P.G) return;
    StripSmem* S = reinterpret_cast<StripSmem*>(smem_raw);

    Cta c;
    c.P = &P; c.ctl = &ctl; c.cta = blockIdx.x; c.G = P.G;
    c.lane = threadIdx.x & 31; c.wid = threadIdx.x >> 5; c.nw = blockDim.x >> 5; c.epoch = 0;
    c.acc[0] = c.acc[1] = c.acc[2] = 0; c.prof = (P.prof != nullptr);
    unsigned long long ph[4] = {0, 0, 0, 0}; // cycles in PCG phase 1, 2, 3 and everything else (thread 0)
    long long tk = c.prof ? clock64() : 0;
#define RS_TICK(slot) do { if (c.prof && threadIdx.x == 0) { const long long tn = clock64(); ph[slot] += (unsigned long long)(tn - tk); tk = tn; } } while (0)
#define RS_TOCK() do { if (c.prof && threadIdx.x == 0) tk = clock64(); } while (0)
    const int lane = c.lane, wid = c.wid;
    const int W = P.W, H = P.H;
    const float wr = P.wr, wf = P.wf, wr2 = P.wr2, wf2 = P.wf2;

    if (threadIdx.x == 0) ctl.abort = 0;
    if (threadIdx.x < 10) { // preconditioner values by (number of valid neighbours, fit)
        const int nv = threadIdx.x % 5, fit = threadIdx.x / 5;
        float DX = (wr2 + wr2) * (float)nv;
        if (fit) DX = DX + wf2;
        ctl.preX[threadIdx.x] = guarded_invert(DX);
        if (!fit) ctl.preA[nv] = guarded_invert(wr2 * (float)nv);
    }

    // ---- which strip is mine, who are my neighbours ----
    const int n = P.n_strips;
    const int s_begin = (int)(((long long)c.cta * n) / c.G), s_end = (int)(((long long)(c.cta + 1) * n) / c.G);
    StripCtx s;
    s.slot = s_begin + wid;
    s.has = s.slot < s_end;
    s.own = &S[wid].T[0][0];
    s.rcs = S[wid].rcs;
    s.up_row = s.own + 0 * TW + 1;
    s.down_row = s.own + (TH - 1) * TW + 1;
    s.lptr = s.own + 1 * TW + lane;
    s.rptr = s.own + 1 * TW + lane + 2;
    s.rem[0] = s.rem[1] = s.rem[2] = s.rem[3] = -1;
    s.outbox = nullptr;
    int sx = 0, sy = 0;
    if (s.has) {
        const int2 xy = P.strip_xy[s.slot];
        sx = xy.x; sy = xy.y;
        s.outbox = P.outbox + (size_t)s.slot * RS_OUTBOX_ENTRIES * 2;
        const int nu = (sy > 0) ? P.slot_of_strip[(sy - 1) * P.SX + sx] : -1;
        const int nd = (sy + 1 < P.SY) ? P.slot_of_strip[(sy + 1) * P.SX + sx] : -1;
        const int nl = (sx > 0) ? P.slot_of_strip[sy * P.SX + sx - 1] : -1;
        const int nr = (sx + 1 < P.SX) ? P.slot_of_strip[sy * P.SX + sx + 1] : -1;
        if (nu >= 0) { if (nu >= s_begin && nu < s_end) s.up_row = &S[nu - s_begin].T[TH - 2][1]; else s.rem[0] = nu; }
        if (nd >= 0) { if (nd >= s_begin && nd < s_end) s.down_row = &S[nd - s_begin].T[1][1]; else s.rem[1] = nd; }
        if (nl >= 0) { if (nl >= s_begin && nl < s_end) { if (lane == 0) s.lptr = &S[nl - s_begin].T[1][TW - 2]; } else s.rem[2] = nl; }
        if (nr >= 0) { if (nr >= s_begin && nr < s_end) { if (lane == 31) s.rptr = &S[nr - s_begin].T[1][1]; } else s.rem[3] = nr; }
    }
    s.x = sx * RS_STRIP_W + lane;
    s.y0 = sy * RS_STRIP_H;
    const bool pub_top = s.rem[0] >= 0, pub_bot = s.rem[1] >= 0, pub_left = s.rem[2] >= 0, pub_right = s.rem[3] >= 0;

    // ---- flags from the mask (constant for the whole launch except the fit bit) ----
    unsigned fl[RS_STRIP_H];
#pragma unroll
    for (int k = 0; k < RS_STRIP_H; ++k) {
        unsigned f = 0;
        const int x = s.x, y = s.y0 + k;
        if (s.has && x < W && y < H) {
            const size_t i = (size_t)y * W + x;
            if (P.M[i] == 0.0f) {
                f = FLAG_ACTIVE;
                if (x + 1 < W && P.M[i + 1] == 0.0f) f |= 1u;
                if (x > 0 && P.M[i - 1] == 0.0f) f |= 2u;
                if (y + 1 < H && P.M[i + W] == 0.0f) f |= 4u;
                if (y > 0 && P.M[i - W] == 0.0f) f |= 8u;
            }
        }
        fl[k] = f;
    }

    // registers that live across the PCG loop
    float r0[RS_STRIP_H], r1[RS_STRIP_H], r2[RS_STRIP_H];
    float d0[RS_STRIP_H], d1[RS_STRIP_H], d2[RS_STRIP_H];
    float pa[RS_STRIP_H], cc[RS_STRIP_H], ss[RS_STRIP_H];
    float q0[RS_STRIP_H], q1[RS_STRIP_H], qa[RS_STRIP_H];
    bool ok = true;
    __syncthreads();

    for (int t = 0; t < P.nCont && ok; ++t) {
        const float alpha_c = (float)(t + 1) / (float)P.nCont; // CombinedSolver.h:199-201
        for (int g = 0; g <= P.nGN && ok; ++g) {
            // ======== prologue: tile <- (X, cos, sin), ring exchange, cost of the current state ========
#pragma unroll
            for (int k = 0; k < RS_STRIP_H; ++k) {
                float4 e = make_float4(0.f, 0.f, 1.f, 0.f);
                if (fl[k] & FLAG_ACTIVE) {
                    const size_t i = (size_t)(s.y0 + k) * W + s.x;
                    const float2 X = P.X[i];
                    float sn, cs;
                    contract_sincos(P.A[i], sn, cs);
                    cc[k] = cs; ss[k] = sn;
                    e = make_float4(X.x, X.y, cs, sn);
                }
                s.own[(k + 1) * TW + lane + 1] = e;
                if (s.has) publish_rowcol(s, lane, k, pub_top, pub_bot, pub_left, pub_right, e, e, false);
            }
            (void)grid_sum(c, 0.f, 0.f, ok);
            if (!ok) break;
            if (s.has) recv_x(P, s, lane);
            __syncthreads();
            {
                float gs0 = 0.f, gs1 = 0.f;
                float4 up = s.up_row[lane], cur = s.own[1 * TW + lane + 1];
#pragma unroll
                for (int k = 0; k < RS_STRIP_H; ++k) {
                    const float4 dn = (k < RS_STRIP_H - 1) ? s.own[(k + 2) * TW + lane + 1] : s.down_row[lane];
                    const unsigned f = fl[k];
                    if (f & FLAG_ACTIVE) {
                        const float4 lf = s.lptr[k * TW], rt = s.rptr[k * TW];
                        float acc = 0.f;
                        if (f & 1u) acc = cost_nb<0>(acc, cur.x, cur.y, cur.z, cur.w, rt, wr);
                        if (f & 2u) acc = cost_nb<1>(acc, cur.x, cur.y, cur.z, cur.w, lf, wr);
                        if (f & 4u) acc = cost_nb<2>(acc, cur.x, cur.y, cur.z, cur.w, dn, wr);
                        if (f & 8u) acc = cost_nb<3>(acc, cur.x, cur.y, cur.z, cur.w, up, wr);
                        const size_t i = (size_t)(s.y0 + k) * W + s.x;
                        const float2 ct = constraint_of(P, i, s.x, s.y0 + k, alpha_c);
                        if (ct.x >= 0.f && ct.y >= 0.f) acc = cost_fit(acc, cur.x, cur.y, ct.x, ct.y, wf);
                        if (k < 4) gs0 = gs0 + acc; else gs1 = gs1 + acc;
                    }
                    up = cur;
                    cur = dn;
                }
                const float tot = grid_sum(c, gs0, gs1, ok);
                if (!ok) break;
                if (c.cta == 0 && threadIdx.x == 0) P.costs[(size_t)t * (P.nGN + 1) + g] = 0.5f * tot;
            }
            if (g == P.nGN) break; // the trailing prologue only produced the final cost

            // ======== PCGInit1: r = -J^T F, z = pre * r ========
            float num;
            {
                float gs0 = 0.f, gs1 = 0.f;
                float4 up = s.up_row[lane], cur = s.own[1 * TW + lane + 1];
#pragma unroll
                for (int k = 0; k < RS_STRIP_H; ++k) {
                    const float4 dn = (k < RS_STRIP_H - 1) ? s.own[(k + 2) * TW + lane + 1] : s.down_row[lane];
                    unsigned f = fl[k] & ~FLAG_FIT;
                    r0[k] = r1[k] = r2[k] = 0.f;
                    d0[k] = d1[k] = d2[k] = 0.f;
                    pa[k] = 0.f;
                    if (f & FLAG_ACTIVE) {
                        const float4 lf = s.lptr[k * TW], rt = s.rptr[k * TW];
                        JtfAcc a;
                        jtf_zero(a);
                        if (f & 1u) jtf_nb<0>(a, cur.x, cur.y, cur.z, cur.w, rt);
                        if (f & 2u) jtf_nb<1>(a, cur.x, cur.y, cur.z, cur.w, lf);
                        if (f & 4u) jtf_nb<2>(a, cur.x, cur.y, cur.z, cur.w, dn);
                        if (f & 8u) jtf_nb<3>(a, cur.x, cur.y, cur.z, cur.w, up);
                        const size_t i = (size_t)(s.y0 + k) * W + s.x;
                        const float2 ct = constraint_of(P, i, s.x, s.y0 + k, alpha_c);
                        const bool fit = (ct.x >= 0.f && ct.y >= 0.f); // arap_plan.t:22
                        if (fit) f |= FLAG_FIT;
                        float g0, g1, ga, DX, DA;
                        jtf_finish(a, cur.x, cur.y, fit, ct.x, ct.y, wr2, wf2, g0, g1, ga, DX, DA);
                        const float pX = guarded_invert(DX), pA = guarded_invert(DA);
                        r0[k] = -g0; r1[k] = -g1; r2[k] = -ga;
                        const float z0 = pX * r0[k], z1 = pX * r1[k], z2 = pA * r2[k];
                        const float term = dot3(r0[k], r1[k], r2[k], z0, z1, z2);
                        if (k < 4) gs0 = gs0 + term; else gs1 = gs1 + term;
                    }
                    fl[k] = f;
                    up = cur;
                    cur = dn;
                }
                // publish (z = p_0, p_old = 0) of the boundary so that neighbours rebuild p_0 with beta = 0
#pragma unroll
                for (int k = 0; k < RS_STRIP_H; ++k) {
                    if (s.has) {
                        const unsigned f = fl[k];
                        const int ix = __popc(f & 15u) + ((f & FLAG_FIT) ? 5 : 0);
                        const float pX = ctl.preX[ix], pA = ctl.preA[__popc(f & 15u)];
                        const float4 v0 = (f & FLAG_ACTIVE) ? make_float4(pX * r0[k], pX * r1[k], pA * r2[k], 0.f)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                        publish_rowcol(s, lane, k, pub_top, pub_bot, pub_left, pub_right, v0,
                                       make_float4(0.f, 0.f, 0.f, 0.f), true);
                    }
                }
                num = grid_sum(c, gs0, gs1, ok); // solverGPUGaussNewton.t:395 scanAlphaNumerator
                if (!ok) break;
                // p_0 = z_0 into the tile (all warps are past their J^T F reads: grid_sum synchronised)
#pragma unroll
                for (int k = 0; k < RS_STRIP_H; ++k) {
                    const unsigned f = fl[k];
                    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (f & FLAG_ACTIVE) {
                        const int ix = __popc(f & 15u) + ((f & FLAG_FIT) ? 5 : 0);
                        const float pX = ctl.preX[ix], pA = ctl.preA[__popc(f & 15u)];
                        const float p0 = pX * r0[k], p1 = pX * r1[k];
                        pa[k] = pA * r2[k];
                        e = make_float4(p0, p1, ss[k] * pa[k], cc[k] * pa[k]);
                    }
                    s.own[(k + 1) * TW + lane + 1] = e;
                }
                if (s.has) recv_p(P, s, lane, 0.0f);
                __syncthreads();
            }

            // ======== PCG iterations ========
            RS_TICK(3);
            for (int it = 0; it < P.nPCG; ++it) {
                // ---- PCGStep1: q = J^T J p, den = sum p.q ----
                float gs0 = 0.f, gs1 = 0.f;
                {
                    float4 up = s.up_row[lane], cur = s.own[1 * TW + lane + 1];
#pragma unroll
                    for (int k = 0; k < RS_STRIP_H; ++k) {
                        const float4 dn = (k < RS_STRIP_H - 1) ? s.own[(k + 2) * TW + lane + 1] : s.down_row[lane];
                        const unsigned f = fl[k];
                        q0[k] = q1[k] = qa[k] = 0.f;
                        if (f & FLAG_ACTIVE) {
                            const float4 lf = s.lptr[k * TW], rt = s.rptr[k * TW];
                            JtjAcc a;
                            jtj_zero(a);
                            if (f & 1u) jtj_nb<0>(a, cur.x, cur.y, rt);
                            if (f & 2u) jtj_nb<1>(a, cur.x, cur.y, lf);
                            if (f & 4u) jtj_nb<2>(a, cur.x, cur.y, dn);
                            if (f & 8u) jtj_nb<3>(a, cur.x, cur.y, up);
                            jtj_finish(a, cc[k], ss[k], cur.x, cur.y, pa[k], (f & FLAG_FIT) != 0, wr2, wf2, q0[k], q1[k],
                                       qa[k]);
                            const float term = dot3(cur.x, cur.y, pa[k], q0[k], q1[k], qa[k]);
                            if (k < 4) gs0 = gs0 + term; else gs1 = gs1 + term;
                        }
                        up = cur;
                        cur = dn;
                    }
                }
                RS_TICK(0);
                const float den = grid_sum(c, gs0, gs1, ok);
                RS_TOCK();
                if (!ok) break;
                const float alpha = (den > 0.0f) ? num / den : 0.0f; // :456-459

                // ---- PCGStep2: delta += alpha p, r -= alpha q, z = pre r, bnum = sum z.r ----
                gs0 = 0.f; gs1 = 0.f;
#pragma unroll
                for (int k = 0; k < RS_STRIP_H; ++k) {
                    const unsigned f = fl[k];
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (f & FLAG_ACTIVE) {
                        const float4 e = s.own[(k + 1) * TW + lane + 1];
                        d0[k] = fmaf(alpha, e.x, d0[k]);
                        d1[k] = fmaf(alpha, e.y, d1[k]);
                        d2[k] = fmaf(alpha, pa[k], d2[k]);
                        r0[k] = fmaf(-alpha, q0[k], r0[k]);
                        r1[k] = fmaf(-alpha, q1[k], r1[k]);
                        r2[k] = fmaf(-alpha, qa[k], r2[k]);
                        const int ix = __popc(f & 15u) + ((f & FLAG_FIT) ? 5 : 0);
                        const float pX = ctl.preX[ix], pA = ctl.preA[__popc(f & 15u)];
                        const float z0 = pX * r0[k], z1 = pX * r1[k], z2 = pA * r2[k];
                        const float term = dot3(z0, z1, z2, r0[k], r1[k], r2[k]);
                        if (k < 4) gs0 = gs0 + term; else gs1 = gs1 + term;
                        v0 = make_float4(z0, z1, z2, e.x);
                        v1 = make_float4(e.y, pa[k], 0.f, 0.f);
                    }
                    if (s.has && it + 1 < P.nPCG)
                        publish_rowcol(s, lane, k, pub_top, pub_bot, pub_left, pub_right, v0, v1, true);
                }
                RS_TICK(1);
                const float bnum = grid_sum(c, gs0, gs1, ok);
                RS_TOCK();
                if (!ok) break;
                if (P.trace && c.cta == 0 && threadIdx.x == 0) {
                    float* tr = P.trace + ((size_t)(t * P.nGN + g) * P.nPCG + it) * 3;
                    tr[0] = den; tr[1] = num; tr[2] = bnum;
                }
                const float beta = (num > 0.0f) ? bnum / num : 0.0f; // :544-547
                num = bnum;                                          // :1091
                if (it + 1 == P.nPCG) break; // the direction is not needed after the last iteration

                // ---- PCGStep3: p = z + beta p (own pixels, then the remote ring) ----
#pragma unroll
                for (int k = 0; k < RS_STRIP_H; ++k) {
                    const unsigned f = fl[k];
                    if (f & FLAG_ACTIVE) {
                        const float4 e = s.own[(k + 1) * TW + lane + 1];
                        const int ix = __popc(f & 15u) + ((f & FLAG_FIT) ? 5 : 0);
                        const float pX = ctl.preX[ix], pA = ctl.preA[__popc(f & 15u)];
                        const float p0 = fmaf(beta, e.x, pX * r0[k]);
                        const float p1 = fmaf(beta, e.y, pX * r1[k]);
                        pa[k] = fmaf(beta, pa[k], pA * r2[k]);
                        s.own[(k + 1) * TW + lane + 1] = make_float4(p0, p1, ss[k] * pa[k], cc[k] * pa[k]);
                    }
                }
                if (s.has) recv_p(P, s, lane, beta);
                __syncthreads();
                RS_TICK(2);
            }
            if (!ok) break;

            // ======== PCGLinearUpdate ========
#pragma unroll
            for (int k = 0; k < RS_STRIP_H; ++k) {
                if (fl[k] & FLAG_ACTIVE) {
                    const size_t i = (size_t)(s.y0 + k) * W + s.x;
                    float2 X = P.X[i];
                    X.x = X.x + d0[k];
                    X.y = X.y + d1[k];
                    P.X[i] = X;
                    P.A[i] = P.A[i] + d2[k];
                }
            }
            __syncthreads(); // tile entries (p) are overwritten by the next prologue
        }
    }
    if (c.prof && threadIdx.x == 0) {
        unsigned long long* o = P.prof + (size_t)c.cta * 8;
        o[0] = ph[0]; o[1] = ph[1]; o[2] = ph[2]; o[3] = ph[3];
        o[4] = c.acc[0]; o[5] = c.acc[1]; o[6] = c.acc[2]; o[7] = c.epoch;
    }
#undef RS_TICK
#undef RS_TOCK
}

// ---- strip table construction -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_strip_active(int W, int H, int SX, int SY, const float* __restrict__ M,
                                                       unsigned char* __restrict__ active)
{
    const int strip = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (strip >= SX * SY) return;
    const int sx = strip % SX, sy = strip / SX; // row-major id here; ordering happens in k_strip_compact
    const int x = sx * RS_STRIP_W + lane;
    bool any = false;
    if (x < W)
        for (int k = 0; k < RS_STRIP_H; ++k) {
            const int y = sy * RS_STRIP_H + k;
            if (y < H && M[(size_t)y * W + x] == 0.0f) any = true;
        }
    const unsigned b = __ballot_sync(0xffffffffu, any);
    if (lane == 0) active[strip] = b ? 1 : 0;
}

// single block: column-major ordered compaction
__global__ void __launch_bounds__(1024) k_strip_compact(int SX, int SY, const unsigned char* __restrict__ active,
                                                         int2* __restrict__ strip_xy, int* __restrict__ slot_of_strip,
                                                         int* __restrict__ count)
{
    __shared__ int wsum[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int total = SX * SY, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int start = 0; start < total; start += 1024) {
        const int j = start + threadIdx.x; // column-major: j = sx * SY + sy
        int sx = 0, sy = 0, a = 0;
        if (j < total) {
            sx = j / SY; sy = j - sx * SY;
            a = active[sy * SX + sx];
        }
        const unsigned b = __ballot_sync(0xffffffffu, a != 0);
        const int prefix = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) wsum[wid] = __popc(b);
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < wid; ++w) woff += wsum[w];
        int tot = 0;
        for (int w = 0; w < 32; ++w) tot += wsum[w];
        const int b0 = base;
        if (j < total) {
            if (a) {
                const int slot = b0 + woff + prefix;
                strip_xy[slot] = make_int2(sx, sy);
                slot_of_strip[sy * SX + sx] = slot;
            } else {
                slot_of_strip[sy * SX + sx] = -1;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) base = b0 + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

} // namespace

// ------------------------------------------------------------------------------------------------ host
ResidentSolver::ResidentSolver(int maxW, int maxH) : maxW_(maxW), maxH_(maxH)
{
    int dev = 0;
    ARAP_CUDA_OR_EXIT(cudaGetDevice(&dev));
    ARAP_CUDA_OR_EXIT(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, dev));
    const int SX = (maxW + RS_STRIP_W - 1) / RS_STRIP_W, SY = (maxH + RS_STRIP_H - 1) / RS_STRIP_H;
    const size_t ns = (size_t)SX * SY;
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_strip_xy_, ns * sizeof(int2)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_slot_of_strip_, ns * sizeof(int) + ns)); // + the active bytes
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_count_, sizeof(int)));
    outbox_cap_ = ns;
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_outbox_, ns * RS_OUTBOX_ENTRIES * 2 * sizeof(float4)));
    ARAP_CUDA_OR_EXIT(cudaMemset(d_outbox_, 0, ns * RS_OUTBOX_ENTRIES * 2 * sizeof(float4)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_slots_, 3 * RS_MAX_CTAS * sizeof(double2)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_status_, 2 * sizeof(int)));
    ARAP_CUDA_OR_EXIT(cudaMemset(d_status_, 0, 2 * sizeof(int)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_prob_, sizeof(ResProb)));
    ARAP_CUDA_OR_EXIT(cudaFuncSetAttribute(k_resident, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(RS_MAX_WARPS * sizeof(StripSmem))));
}

ResidentSolver::~ResidentSolver()
{
    cudaFree(d_strip_xy_); cudaFree(d_slot_of_strip_); cudaFree(d_count_); cudaFree(d_outbox_);
    cudaFree(d_slots_); cudaFree(d_status_); cudaFree(d_prob_);
}

bool ResidentSolver::prepare(int W, int H, const float* d_M, cudaStream_t stream)
{
    if ((size_t)W * H > (size_t)maxW_ * maxH_) return false;
    W_ = W; H_ = H;
    SX_ = (W + RS_STRIP_W - 1) / RS_STRIP_W;
    SY_ = (H + RS_STRIP_H - 1) / RS_STRIP_H;
    if ((size_t)SX_ * SY_ > outbox_cap_) return false;
    d_M_ = d_M;
    unsigned char* d_active = reinterpret_cast<unsigned char*>(d_slot_of_strip_ + (size_t)SX_ * SY_);
    k_strip_active<<<(SX_ * SY_ + 7) / 8, 256, 0, stream>>>(W, H, SX_, SY_, d_M, d_active);
    k_strip_compact<<<1, 1024, 0, stream>>>(SX_, SY_, d_active, d_strip_xy_, d_slot_of_strip_, d_count_);
    launches_ += 2;
    ARAP_CUDA_OR_EXIT(cudaMemcpyAsync(&n_strips_, d_count_, sizeof(int), cudaMemcpyDeviceToHost, stream));
    ARAP_CUDA_OR_EXIT(cudaStreamSynchronize(stream));
    const int max_warps = RS_THREADS_MAX / 32;
    if (n_strips_ == 0) { G_ = 1; NW_ = 1; return true; }
    NW_ = (n_strips_ + sm_count_ - 1) / sm_count_;
    if (NW_ < 4) NW_ = 4;
    if (NW_ > max_warps) return false; // does not fit on chip
    G_ = (n_strips_ + NW_ - 1) / NW_;
    if (G_ > sm_count_) G_ = sm_count_;
    if (G_ > RS_MAX_CTAS) return false;
    // balanced split: ceil(n/G) strips at most per CTA
    if ((n_strips_ + G_ - 1) / G_ > NW_) return false;
    return true;
}

void ResidentSolver::enqueue(float2* X, float* A, const float2* C, int lerp_mode, float wf, float wr, int nCont, int nGN,
                             int nPCG, float* d_costs, float* d_trace, cudaStream_t stream)
{
    ResProb p{};
    p.W = W_; p.H = H_; p.SX = SX_; p.SY = SY_; p.n_strips = n_strips_; p.G = G_;
    p.X = X; p.A = A; p.C = C; p.M = d_M_; p.lerp_mode = lerp_mode;
    p.wf = wf; p.wr = wr; p.wf2 = wf * wf; p.wr2 = wr * wr;
    p.strip_xy = d_strip_xy_; p.slot_of_strip = d_slot_of_strip_;
    p.outbox = d_outbox_; p.slots = d_slots_; p.costs = d_costs; p.trace = d_trace; p.status = d_status_;
    p.nCont = nCont; p.nGN = nGN; p.nPCG = nPCG;
    p.prof = d_prof_;
    ARAP_CUDA_OR_EXIT(cudaMemcpyAsync(d_prob_, &p, sizeof(p), cudaMemcpyHostToDevice, stream));
    ARAP_CUDA_OR_EXIT(cudaMemsetAsync(d_slots_, 0xFF, 3 * RS_MAX_CTAS * sizeof(double2), stream));
    const int threads = NW_ * 32;
    const size_t smem = (size_t)NW_ * sizeof(StripSmem);
    int per_sm = 0;
    ARAP_CUDA_OR_EXIT(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_resident, threads, smem));
    if (per_sm * sm_count_ < G_) {
        fprintf(stderr, "arapb200: resident kernel cannot be co-resident (%d CTAs, %d per SM)\n", G_, per_sm);
        exit(1);
    }
    const ResProb* dp = d_prob_;
    void* args[] = {(void*)&dp};
    ARAP_CUDA_OR_EXIT(cudaLaunchCooperativeKernel((const void*)k_resident, dim3(G_, 1, 1), dim3(threads, 1, 1), args,
                                                  smem, stream));
    launches_ += 1;
}

int ResidentSolver::status(cudaStream_t stream)
{
    int st[2] = {0, 0};
    ARAP_CUDA_OR_EXIT(cudaMemcpyAsync(st, d_status_, sizeof(st), cudaMemcpyDeviceToHost, stream));
    ARAP_CUDA_OR_EXIT(cudaStreamSynchronize(stream));
    if (st[0]) {
        ARAP_CUDA_OR_EXIT(cudaMemsetAsync(d_status_, 0, sizeof(st), stream));
        return st[1] ? st[1] : 1;
    }
    return 0;
}

} // namespace arapb200
