// solver_stream.cu -- streaming Gauss-Newton / PCG back-end.
//
// Replaces the generated kernels PCGInit1 / PCGStep1 / PCGStep2 / PCGStep3 / PCGLinearUpdate /
// computeCost and the host loop of ARAP/API/src/solverGPUGaussNewton.t:361-397, 421-434, 446-489,
// 537-557, 580-592, 1016-1177 with:
//   k_prep    flags (validity of the 4 neighbours, fit, active), cos/sin table, UrShape check
//   k_init    r = -J^T F, pre, p = pre*r, delta = 0, sum r.p                    (PCGInit1)
//   k_step_a  p = pre*r + beta*p (recomputed on the halo), q = J^T J p, sum p.q  (PCGStep3 + PCGStep1)
//   k_step_b  alpha; delta += alpha p; r -= alpha q; sum (pre*r).r              (PCGStep2)
//   k_update  X += delta, refresh cos/sin                                        (PCGLinearUpdate)
//   k_cost    0.5 * sum residual^2                                               (computeCost)
// Scalars (alpha/beta numerators and denominators) never leave the device; every reduction is the
// deterministic exact sum of contract C3; a whole GN step (2*nPCG + 3 kernels) is one graph launch.
#include "solver_stream.cuh"
#include "grid_math.cuh"

namespace arapb200 {

namespace {

constexpr int TS = ST_TILE + 2; // staged tile pitch (1-pixel halo)


__device__ __forceinline__ unsigned long long* acc_set(const StreamPlanes& pl, int set)
{
    return pl.acc + set * WA_WORDS;
}
__device__ __forceinline__ int bn_set(int j) // accumulator of r.z of PCG iteration j (j >= -2)
{
    return (j + 3) % 3;
}
// Block result (valid in warp 0) -> the grid-wide accumulator `set`.  Fire and forget: no fence, no counter; the
// consumers run in a later kernel.
__device__ __forceinline__ void publish(const StreamPlanes& pl, int set, HL blockval)
{
    if (threadIdx.x == 0) wide_add(acc_set(pl, set), blockIdx.x, blockval.h);
    if (threadIdx.x == 1) wide_add(acc_set(pl, set), blockIdx.x, blockval.l);
}
// Warp 0 of a block reads two accumulators at once: lanes 0..15 `set_lo`, lanes 16..31 `set_hi`.
__device__ __forceinline__ long long fetch2(const StreamPlanes& pl, int set_lo, int set_hi)
{
    const int lane = threadIdx.x & 31;
    return wide_fetch(acc_set(pl, lane < 16 ? set_lo : set_hi), lane & 15);
}
__device__ __forceinline__ void zero_set(const StreamPlanes& pl, int set, int first_thread)
{
    const int t = (int)threadIdx.x - first_thread;
    if (t >= 0 && t < WA_WORDS) acc_set(pl, set)[t] = 0ULL;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS) k_prep(const StreamDev* __restrict__ dpp)
{
    const StreamDev& dp = *dpp;
    const int W = dp.W, H = dp.H;
    const int x = (blockIdx.x % dp.tx) * ST_TILE + (threadIdx.x & 31);
    const int yb = (blockIdx.x / dp.tx) * ST_TILE + (threadIdx.x >> 5) * 4;
    unsigned bad = 0;
    int any = 0;
    if (x < W) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int y = yb + r;
            if (y >= H) break;
            const size_t i = (size_t)y * W + x;
            unsigned f = 0;
            if (dp.M[i] == 0.0f) {
                f = FLAG_ACTIVE;
                any = 1;
                if (x + 1 < W && dp.M[i + 1] == 0.0f) f |= 1u;
                if (x > 0 && dp.M[i - 1] == 0.0f) f |= 2u;
                if (y + 1 < H && dp.M[i + W] == 0.0f) f |= 4u;
                if (y > 0 && dp.M[i - W] == 0.0f) f |= 8u;
                const float2 c = dp.C[i];
                if (c.x >= 0.0f && c.y >= 0.0f) f |= FLAG_FIT; // arap_plan.t:22
                float s, co;
                contract_sincos(dp.A[i], s, co);
                dp.cs[0][i] = co;
                dp.cs[1][i] = s;
                const float2 u = dp.U[i];
                if (u.x != (float)x || u.y != (float)y) bad = 1;
            }
            dp.flags[i] = (unsigned char)f;
        }
    }
    if (bad) atomicAdd(&dp.sc->bad_u, 1u);
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) dp.tile_active[blockIdx.x] = any ? 1 : 0;
}

// Stage (X_x, X_y, cos, sin) of a tile + halo.  Inactive / out-of-image entries are never used.
__device__ __forceinline__ void stage_x_tile(const StreamDev& dp, float4 (*T)[TS], int x0, int y0)
{
    for (int e = threadIdx.x; e < TS * TS; e += ST_THREADS) {
        const int ly = e / TS, lx = e - ly * TS;
        const int x = x0 + lx - 1, y = y0 + ly - 1;
        float4 v = make_float4(0.f, 0.f, 1.f, 0.f);
        if (x >= 0 && x < dp.W && y >= 0 && y < dp.H) {
            const size_t i = (size_t)y * dp.W + x;
            if (dp.flags[i] & FLAG_ACTIVE) {
                const float2 X = dp.X[i];
                v = make_float4(X.x, X.y, dp.cs[0][i], dp.cs[1][i]);
            }
        }
        T[ly][lx] = v;
    }
}

// PCGInit1 (solverGPUGaussNewton.t:361-397)
__global__ void __launch_bounds__(ST_THREADS) k_init(const StreamDev* __restrict__ dpp)
{
    __shared__ float4 T[TS][TS];
    __shared__ double red[64];
    const StreamDev& dp = *dpp;
    const int W = dp.W, H = dp.H;
    const int x0 = (blockIdx.x % dp.tx) * ST_TILE, y0 = (blockIdx.x / dp.tx) * ST_TILE;
    stage_x_tile(dp, T, x0, y0);
    __syncthreads();
    const int lx = threadIdx.x & 31, lyb = (threadIdx.x >> 5) * 4;
    const int x = x0 + lx;
    float g = 0.0f;
    if (x < W) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int ly = lyb + r, y = y0 + ly;
            if (y >= H) break;
            const size_t i = (size_t)y * W + x;
            const unsigned f = dp.flags[i];
            if (!(f & FLAG_ACTIVE)) {
                // the branch-free PCG kernels rely on zeros here; a previous problem may have left values behind
                dp.pre[0][i] = 0.f; dp.pre[1][i] = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) { dp.r[k][i] = 0.f; dp.p[0][k][i] = 0.f; dp.p[1][k][i] = 0.f; dp.q[k][i] = 0.f; dp.d[k][i] = 0.f; }
                dp.cs[0][i] = 0.f; dp.cs[1][i] = 0.f;
                continue;
            }
            const float4 Ei = T[ly + 1][lx + 1];
            JtfAcc a;
            jtf_zero(a);
            if (f & 1u) jtf_nb<0>(a, Ei.x, Ei.y, Ei.z, Ei.w, T[ly + 1][lx + 2]);
            if (f & 2u) jtf_nb<1>(a, Ei.x, Ei.y, Ei.z, Ei.w, T[ly + 1][lx]);
            if (f & 4u) jtf_nb<2>(a, Ei.x, Ei.y, Ei.z, Ei.w, T[ly + 2][lx + 1]);
            if (f & 8u) jtf_nb<3>(a, Ei.x, Ei.y, Ei.z, Ei.w, T[ly][lx + 1]);
            const bool fit = (f & FLAG_FIT) != 0;
            float2 c = make_float2(0.f, 0.f);
            if (fit) c = dp.C[i];
            float g0, g1, ga, DX, DA;
            jtf_finish(a, Ei.x, Ei.y, fit, c.x, c.y, dp.wr2, dp.wf2, g0, g1, ga, DX, DA);
            const float pX = guarded_invert(DX), pA = guarded_invert(DA);
            const float r0 = -g0, r1 = -g1, r2 = -ga;
            const float p0 = pX * r0, p1 = pX * r1, p2 = pA * r2;
            dp.pre[0][i] = pX;
            dp.pre[1][i] = pA;
            dp.r[0][i] = r0; dp.r[1][i] = r1; dp.r[2][i] = r2;
            dp.p[0][0][i] = p0; dp.p[0][1][i] = p1; dp.p[0][2][i] = p2;
            dp.d[0][i] = 0.f; dp.d[1][i] = 0.f; dp.d[2][i] = 0.f;
            g = g + dot3(r0, r1, r2, p0, p1, p2);
        }
    }
    publish(dp, bn_set(-1), block_exact_sum(g, red));
}

// PCGStep3 of the previous iteration fused with PCGStep1 (solverGPUGaussNewton.t:537-550, 421-434)
template <bool FIRST>
__global__ void __launch_bounds__(ST_THREADS) k_step_a(const __grid_constant__ StreamPlanes pl,
                                                       const StreamDev* __restrict__ dpp, int it)
{
    __shared__ float4 T[TS][TS];
    __shared__ double red[64];
    __shared__ float s_beta;
    const bool tile_on = pl.tile_active[blockIdx.x] != 0; // tiles without object pixels have nothing to add
    if (!tile_on && blockIdx.x != 0) return;
    const float wr2 = dpp->wr2, wf2 = dpp->wf2;
    const int W = pl.W, H = pl.H;
    const int x0 = (blockIdx.x % pl.tx) * ST_TILE, y0 = (blockIdx.x / pl.tx) * ST_TILE;
    float beta = 0.0f;
    if (!FIRST) {
        if (threadIdx.x < 32) {
            const float v = wide_round(fetch2(pl, bn_set(it - 1), bn_set(it - 2)));
            const float bnum = __shfl_sync(0xffffffffu, v, 0), num = __shfl_sync(0xffffffffu, v, 16);
            if (threadIdx.x == 0) {
                s_beta = (num > 0.0f) ? bnum / num : 0.0f; // :544-547
                if (blockIdx.x == 0 && dpp->trace) dpp->trace[3 * (it - 1) + 2] = bnum;
            }
        }
        __syncthreads();
        beta = s_beta;
    }
    // p is ping-ponged between two buffers: the halo of a tile needs the OLD direction of pixels that
    // the neighbouring tile is updating in this very kernel.
    float* const* __restrict__ psrc = pl.p[FIRST ? 0 : ((it - 1) & 1)];
    float* const* __restrict__ pdst = pl.p[it & 1];
    // stage (p_x, p_y, sin*p_a, cos*p_a) of tile + halo, p being the NEW direction
    for (int e = threadIdx.x; tile_on && e < TS * TS; e += ST_THREADS) {
        const int ly = e / TS, lx = e - ly * TS;
        const int x = x0 + lx - 1, y = y0 + ly - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x >= 0 && x < W && y >= 0 && y < H) {
            // no activity test: every plane of an inactive pixel holds zeros, which flow through as zeros
            const size_t i = (size_t)y * W + x;
            float p0 = psrc[0][i], p1 = psrc[1][i], p2 = psrc[2][i];
            const float c = pl.cs[0][i], sn = pl.cs[1][i];
            if (!FIRST) {
                const float pX = pl.pre[0][i], pA = pl.pre[1][i];
                const float r0 = pl.r[0][i], r1 = pl.r[1][i], r2 = pl.r[2][i];
                p0 = fmaf(beta, p0, pX * r0);
                p1 = fmaf(beta, p1, pX * r1);
                p2 = fmaf(beta, p2, pA * r2);
                const bool interior = (lx >= 1 && lx <= ST_TILE && ly >= 1 && ly <= ST_TILE);
                if (interior) { pdst[0][i] = p0; pdst[1][i] = p1; pdst[2][i] = p2; }
            }
            v = make_float4(p0, p1, sn * p2, c * p2);
        }
        T[ly][lx] = v;
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, lyb = (threadIdx.x >> 5) * 4;
    const int x = x0 + lx;
    float g = 0.0f;
    if (x < W && tile_on) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int ly = lyb + r, y = y0 + ly;
            if (y >= H) break;
            const size_t i = (size_t)y * W + x;
            const unsigned f = pl.flags[i];
            if (!(f & FLAG_ACTIVE)) continue;
            const float4 Pi = T[ly + 1][lx + 1];
            float pa = psrc[2][i];
            if (!FIRST) pa = fmaf(beta, pa, pl.pre[1][i] * pl.r[2][i]);
            JtjAcc a;
            jtj_zero(a);
            if (f & 1u) jtj_nb<0>(a, Pi.x, Pi.y, T[ly + 1][lx + 2]);
            if (f & 2u) jtj_nb<1>(a, Pi.x, Pi.y, T[ly + 1][lx]);
            if (f & 4u) jtj_nb<2>(a, Pi.x, Pi.y, T[ly + 2][lx + 1]);
            if (f & 8u) jtj_nb<3>(a, Pi.x, Pi.y, T[ly][lx + 1]);
            float q0, q1, qa;
            jtj_finish(a, pl.cs[0][i], pl.cs[1][i], Pi.x, Pi.y, pa, (f & FLAG_FIT) != 0, wr2, wf2, q0, q1, qa);
            pl.q[0][i] = q0; pl.q[1][i] = q1; pl.q[2][i] = qa;
            g = g + dot3(Pi.x, Pi.y, pa, q0, q1, qa);
        }
    }
    publish(pl, ST_ACC_D0 + (it & 1), block_exact_sum(g, red));
}

// PCGStep2 (solverGPUGaussNewton.t:446-489).  Branch-free: the planes of inactive pixels hold zeros (they are
// zero-initialised and never written), so they flow through as exact zeros and add +0 to the group term;
// every load of the four rows is issued before the first use.
__global__ void __launch_bounds__(ST_THREADS) k_step_b(const __grid_constant__ StreamPlanes pl,
                                                       const StreamDev* __restrict__ dpp, int it)
{
    __shared__ double red[64];
    __shared__ float s_alpha;
    const bool tile_on = pl.tile_active[blockIdx.x] != 0;
    if (!tile_on && blockIdx.x != 0) return;
    const int W = pl.W, H = pl.H;
    const int x = (blockIdx.x % pl.tx) * ST_TILE + (threadIdx.x & 31);
    const int yb = (blockIdx.x / pl.tx) * ST_TILE + (threadIdx.x >> 5) * 4;
    // warp 0: r.z of the previous iteration and this iteration's p.q; fetched before the planes, decoded after
    long long raw = 0;
    if (threadIdx.x < 32) raw = fetch2(pl, bn_set(it - 1), ST_ACC_D0 + (it & 1));
    const bool on = tile_on && x < W;
    const float* __restrict__ pk0 = pl.p[it & 1][0];
    const float* __restrict__ pk1 = pl.p[it & 1][1];
    const float* __restrict__ pk2 = pl.p[it & 1][2];
    float pv[4][3], qv[4][3], rv[4][3], dv[4][3], pre[4][2];
    if (on) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int y = min(yb + r, H - 1); // rows past the image re-read the last row; they are not stored
            const size_t i = (size_t)y * W + x;
            pv[r][0] = pk0[i]; pv[r][1] = pk1[i]; pv[r][2] = pk2[i];
            qv[r][0] = pl.q[0][i]; qv[r][1] = pl.q[1][i]; qv[r][2] = pl.q[2][i];
            rv[r][0] = pl.r[0][i]; rv[r][1] = pl.r[1][i]; rv[r][2] = pl.r[2][i];
            dv[r][0] = pl.d[0][i]; dv[r][1] = pl.d[1][i]; dv[r][2] = pl.d[2][i];
            pre[r][0] = pl.pre[0][i]; pre[r][1] = pl.pre[1][i];
        }
    }
    if (threadIdx.x < 32) {
        const float v = wide_round(raw);
        const float num = __shfl_sync(0xffffffffu, v, 0), den = __shfl_sync(0xffffffffu, v, 16);
        if (threadIdx.x == 0) {
            s_alpha = (den > 0.0f) ? num / den : 0.0f; // :456-459
            if (blockIdx.x == 0 && dpp->trace) {
                dpp->trace[3 * it] = den;
                dpp->trace[3 * it + 1] = num;
            }
        }
    }
    if (blockIdx.x == 0) { // recycle the accumulators nobody reads any more (their next writers are later kernels)
        zero_set(pl, bn_set(it - 2), 64);
        zero_set(pl, ST_ACC_D0 + ((it + 1) & 1), 160);
    }
    __syncthreads();
    const float alpha = s_alpha;
    float g = 0.0f;
    if (on) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int y = yb + r;
            if (y < H) {
                const size_t i = (size_t)y * W + x;
                float rr[3], zz[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    pl.d[k][i] = fmaf(alpha, pv[r][k], dv[r][k]);
                    rr[k] = fmaf(-alpha, qv[r][k], rv[r][k]);
                    pl.r[k][i] = rr[k];
                    zz[k] = ((k < 2) ? pre[r][0] : pre[r][1]) * rr[k];
                }
                g = g + dot3(zz[0], zz[1], zz[2], rr[0], rr[1], rr[2]);
            }
        }
    }
    publish(pl, bn_set(it), block_exact_sum(g, red));
}

// PCGLinearUpdate (solverGPUGaussNewton.t:552-557) + cos/sin refresh for the new angles
__global__ void __launch_bounds__(ST_THREADS) k_update(const StreamDev* __restrict__ dpp)
{
    const StreamDev& dp = *dpp;
    const int W = dp.W, H = dp.H;
    const int x = (blockIdx.x % dp.tx) * ST_TILE + (threadIdx.x & 31);
    const int yb = (blockIdx.x / dp.tx) * ST_TILE + (threadIdx.x >> 5) * 4;
    if (x >= W) return;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int y = yb + r;
        if (y >= H) break;
        const size_t i = (size_t)y * W + x;
        if (!(dp.flags[i] & FLAG_ACTIVE)) continue;
        float2 X = dp.X[i];
        X.x = X.x + dp.d[0][i];
        X.y = X.y + dp.d[1][i];
        dp.X[i] = X;
        const float a = dp.A[i] + dp.d[2][i];
        dp.A[i] = a;
        float s, c;
        contract_sincos(a, s, c);
        dp.cs[0][i] = c;
        dp.cs[1][i] = s;
    }
}

// computeCost (solverGPUGaussNewton.t:580-592, o.t:2375-2385)
__global__ void __launch_bounds__(ST_THREADS) k_cost(const StreamDev* __restrict__ dpp)
{
    __shared__ float4 T[TS][TS];
    __shared__ double red[64];
    const StreamDev& dp = *dpp;
    const int W = dp.W, H = dp.H;
    const int x0 = (blockIdx.x % dp.tx) * ST_TILE, y0 = (blockIdx.x / dp.tx) * ST_TILE;
    stage_x_tile(dp, T, x0, y0);
    __syncthreads();
    const int lx = threadIdx.x & 31, lyb = (threadIdx.x >> 5) * 4;
    const int x = x0 + lx;
    float g = 0.0f;
    if (x < W) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int ly = lyb + r, y = y0 + ly;
            if (y >= H) break;
            const size_t i = (size_t)y * W + x;
            const unsigned f = dp.flags[i];
            if (!(f & FLAG_ACTIVE)) continue;
            const float4 Ei = T[ly + 1][lx + 1];
            float acc = 0.0f;
            if (f & 1u) acc = cost_nb<0>(acc, Ei.x, Ei.y, Ei.z, Ei.w, T[ly + 1][lx + 2], dp.wr);
            if (f & 2u) acc = cost_nb<1>(acc, Ei.x, Ei.y, Ei.z, Ei.w, T[ly + 1][lx], dp.wr);
            if (f & 4u) acc = cost_nb<2>(acc, Ei.x, Ei.y, Ei.z, Ei.w, T[ly + 2][lx + 1], dp.wr);
            if (f & 8u) acc = cost_nb<3>(acc, Ei.x, Ei.y, Ei.z, Ei.w, T[ly][lx + 1], dp.wr);
            if (f & FLAG_FIT) {
                const float2 c = dp.C[i];
                acc = cost_fit(acc, Ei.x, Ei.y, c.x, c.y, dp.wf);
            }
            g = g + acc;
        }
    }
    publish(dp, ST_ACC_COST, block_exact_sum(g, red));
}

// Cost accumulator -> scalars; with tracing, also the last iteration's r.z
__global__ void __launch_bounds__(32) k_finish(const __grid_constant__ StreamPlanes pl, const StreamDev* __restrict__ dpp,
                                               int last_it)
{
    const float v = wide_round(fetch2(pl, ST_ACC_COST, bn_set(last_it < 0 ? 0 : last_it)));
    if (threadIdx.x == 0) pl.sc->cost = 0.5f * v;
    if (threadIdx.x == 16 && last_it >= 0 && dpp->trace) dpp->trace[3 * last_it + 2] = v;
}

// debug: one accumulator -> a float
__global__ void __launch_bounds__(32) k_decode(const __grid_constant__ StreamPlanes pl, int set, float* dst)
{
    const float v = wide_round(fetch2(pl, set, set));
    if (threadIdx.x == 0) *dst = v;
}

} // namespace

// ------------------------------------------------------------------------------------------ host
static constexpr size_t ACC_BYTES = (size_t)ST_ACC_SETS * WA_WORDS * sizeof(unsigned long long);

StreamSolver::StreamSolver(int W, int H)
{
    h_.W = W;
    h_.H = H;
    h_.tx = (W + ST_TILE - 1) / ST_TILE;
    h_.ty = (H + ST_TILE - 1) / ST_TILE;
    h_.ntiles = h_.tx * h_.ty;
    const size_t N = (size_t)W * H;
    const size_t Np = (N + 63) & ~(size_t)63; // keep every plane 256-byte aligned
    ARAP_CUDA_OR_EXIT(cudaMalloc(&planes_, 19 * Np * sizeof(float)));
    ARAP_CUDA_OR_EXIT(cudaMemset(planes_, 0, 19 * Np * sizeof(float)));
    float* b = planes_;
    for (int k = 0; k < 3; ++k) { h_.r[k] = b; b += Np; }
    for (int k = 0; k < 3; ++k) { h_.p[0][k] = b; b += Np; }
    for (int k = 0; k < 3; ++k) { h_.p[1][k] = b; b += Np; }
    for (int k = 0; k < 3; ++k) { h_.q[k] = b; b += Np; }
    for (int k = 0; k < 3; ++k) { h_.d[k] = b; b += Np; }
    for (int k = 0; k < 2; ++k) { h_.cs[k] = b; b += Np; }
    for (int k = 0; k < 2; ++k) { h_.pre[k] = b; b += Np; }
    ARAP_CUDA_OR_EXIT(cudaMalloc(&h_.flags, Np));
    ARAP_CUDA_OR_EXIT(cudaMemset(h_.flags, 0, Np));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&h_.tile_active, (size_t)h_.ntiles));
    ARAP_CUDA_OR_EXIT(cudaMemset(h_.tile_active, 1, (size_t)h_.ntiles));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&h_.acc, ACC_BYTES));
    ARAP_CUDA_OR_EXIT(cudaMemset(h_.acc, 0, ACC_BYTES));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&h_.sc, sizeof(StreamScalars)));
    ARAP_CUDA_OR_EXIT(cudaMemset(h_.sc, 0, sizeof(StreamScalars)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_, sizeof(StreamDev)));
    h_.trace = nullptr;
}

StreamSolver::~StreamSolver()
{
    if (graph_) cudaGraphExecDestroy(graph_);
    cudaFree(planes_);
    cudaFree(h_.flags);
    cudaFree(h_.tile_active);
    cudaFree(h_.acc);
    cudaFree(h_.sc);
    cudaFree(d_);
}

void StreamSolver::upload(cudaStream_t stream)
{
    ARAP_CUDA_OR_EXIT(cudaMemcpyAsync(d_, &h_, sizeof(StreamDev), cudaMemcpyHostToDevice, stream));
}

void StreamSolver::bind(float2* X, float* A, const float2* U, const float2* C, const float* M, float wf,
                        float wr, cudaStream_t stream)
{
    h_.X = X; h_.A = A; h_.U = U; h_.C = C; h_.M = M;
    h_.wf = wf; h_.wr = wr; h_.wf2 = wf * wf; h_.wr2 = wr * wr;
    h_.trace = nullptr;
    upload(stream);
}

void StreamSolver::enqueue_prep(cudaStream_t stream)
{
    ARAP_CUDA_OR_EXIT(cudaMemsetAsync(h_.acc, 0, ACC_BYTES, stream));
    k_prep<<<h_.ntiles, ST_THREADS, 0, stream>>>(d_);
    ++launches_;
}

void StreamSolver::enqueue_pcg_init(cudaStream_t stream)
{
    k_init<<<h_.ntiles, ST_THREADS, 0, stream>>>(d_);
    ++launches_;
}

void StreamSolver::enqueue_step_a(bool first, int it, cudaStream_t stream)
{
    const StreamPlanes& pl = h_;
    if (first) k_step_a<true><<<h_.ntiles, ST_THREADS, 0, stream>>>(pl, d_, it);
    else k_step_a<false><<<h_.ntiles, ST_THREADS, 0, stream>>>(pl, d_, it);
    k_decode<<<1, 32, 0, stream>>>(pl, ST_ACC_D0 + (it & 1), &h_.sc->den);
    launches_ += 2;
}

void StreamSolver::enqueue_init(cudaStream_t stream)
{
    const StreamPlanes& pl = h_;
    ARAP_CUDA_OR_EXIT(cudaMemsetAsync(&h_.sc->bad_u, 0, sizeof(unsigned), stream));
    enqueue_prep(stream);
    k_cost<<<h_.ntiles, ST_THREADS, 0, stream>>>(d_);
    k_finish<<<1, 32, 0, stream>>>(pl, d_, -1);
    launches_ += 2;
    ARAP_CUDA_OR_EXIT(cudaGetLastError());
}

void StreamSolver::launch_gn_body(int nPCG, cudaStream_t stream, bool tracing)
{
    const StreamPlanes& pl = h_;
    ARAP_CUDA_OR_EXIT(cudaMemsetAsync(h_.acc, 0, ACC_BYTES, stream));
    // the caller may have changed the constraint image / mask between steps (Opt.h:58-60): refresh flags
    k_prep<<<h_.ntiles, ST_THREADS, 0, stream>>>(d_);
    k_init<<<h_.ntiles, ST_THREADS, 0, stream>>>(d_);
    for (int it = 0; it < nPCG; ++it) {
        if (it == 0) k_step_a<true><<<h_.ntiles, ST_THREADS, 0, stream>>>(pl, d_, it);
        else k_step_a<false><<<h_.ntiles, ST_THREADS, 0, stream>>>(pl, d_, it);
        k_step_b<<<h_.ntiles, ST_THREADS, 0, stream>>>(pl, d_, it);
    }
    k_update<<<h_.ntiles, ST_THREADS, 0, stream>>>(d_);
    k_cost<<<h_.ntiles, ST_THREADS, 0, stream>>>(d_);
    k_finish<<<1, 32, 0, stream>>>(pl, d_, nPCG - 1);
    (void)tracing;
}

void StreamSolver::enqueue_gn_step(int nPCG, cudaStream_t stream, float* d_trace)
{
    const long long nodes = 2LL * nPCG + 5;
    if (d_trace) {
        h_.trace = d_trace;
        upload(stream);
        launch_gn_body(nPCG, stream, true);
        ARAP_CUDA_OR_EXIT(cudaGetLastError());
        h_.trace = nullptr;
        upload(stream);
        launches_ += nodes;
        return;
    }
    if (!graph_ || graph_npcg_ != nPCG) {
        if (graph_) { cudaGraphExecDestroy(graph_); graph_ = nullptr; }
        cudaStream_t cap;
        ARAP_CUDA_OR_EXIT(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        cudaGraph_t g;
        ARAP_CUDA_OR_EXIT(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        launch_gn_body(nPCG, cap, false);
        ARAP_CUDA_OR_EXIT(cudaStreamEndCapture(cap, &g));
        ARAP_CUDA_OR_EXIT(cudaGraphInstantiate(&graph_, g, 0));
        ARAP_CUDA_OR_EXIT(cudaGraphDestroy(g));
        ARAP_CUDA_OR_EXIT(cudaStreamDestroy(cap));
        graph_npcg_ = nPCG;
        graph_nodes_ = nodes;
    }
    ARAP_CUDA_OR_EXIT(cudaGraphLaunch(graph_, stream));
    launches_ += graph_nodes_;
}

void StreamSolver::read_back(cudaStream_t stream, float* cost, unsigned* bad_u)
{
    StreamScalars s;
    ARAP_CUDA_OR_EXIT(cudaMemcpyAsync(&s, h_.sc, sizeof(s), cudaMemcpyDeviceToHost, stream));
    ARAP_CUDA_OR_EXIT(cudaStreamSynchronize(stream));
    if (cost) *cost = s.cost;
    if (bad_u) *bad_u = s.bad_u;
}

} // namespace arapb200
