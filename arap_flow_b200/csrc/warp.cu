// warp.cu -- forward warp: rasterise the deformed pixel grid into warped RGB, warped mask and the
// per-pixel splat index.
//
// Replaces the single-threaded CPU rasteriser of ARAP/deformation/src/CombinedSolver.h:248-342
// (in-app) == ARAP/warping/src/main.cpp:110-225 (stand-alone tool); edge function PointInTriangleLK
// CombinedSolver.h:61-97 / main.cpp:68-104.  The reference walks the pixel quads in row-major order,
// two triangles per quad, and lets the LAST writer win.  Here the same result is produced without any
// ordering between threads: pass 1 does an integer atomicMax of (triangle sequence number + 1) into a
// uint32 z-buffer for every covered output pixel, pass 2 re-evaluates the winning triangle's
// barycentrics and resolves colour + mask.  The outcome is therefore deterministic and bit-identical
// to the sequential reference.  All float arithmetic uses explicit round-to-nearest intrinsics (no
// FMA contraction), because the reference's colour bytes change under contraction.
#include "warp.cuh"

namespace arapb200 {
namespace {

struct Bary {
    float b0, b1, b2;
    bool in;
};

// PointInTriangleLK with w0 = w1 = w2 = 1 (main.cpp:68-104), operation for operation.
__device__ __forceinline__ Bary lk(float x0, float y0, float x1, float y1, float x2, float y2, float sx, float sy)
{
    const float X0 = __fsub_rn(x0, sx), X1 = __fsub_rn(x1, sx), X2 = __fsub_rn(x2, sx);
    const float Y0 = __fsub_rn(y0, sy), Y1 = __fsub_rn(y1, sy), Y2 = __fsub_rn(y2, sy);
    float d01 = __fsub_rn(__fmul_rn(X0, Y1), __fmul_rn(Y0, X1));
    float d12 = __fsub_rn(__fmul_rn(X1, Y2), __fmul_rn(Y1, X2));
    float d20 = __fsub_rn(__fmul_rn(X2, Y0), __fmul_rn(Y2, X0));
    Bary r;
    r.in = false;
    r.b0 = r.b1 = r.b2 = 0.f;
    if ((d01 < 0.f) & (d12 < 0.f) & (d20 < 0.f)) return r; // backfacing
    const float inv = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(d01, d12), d20));
    d01 = __fmul_rn(d01, inv);
    d12 = __fmul_rn(d12, inv);
    d20 = __fmul_rn(d20, inv);
    r.b0 = d12;
    r.b1 = d20;
    r.b2 = d01;
    r.in = (d01 >= 0.f && d12 >= 0.f && d20 >= 0.f); // NaN => not covered
    return r;
}

__device__ __forceinline__ void splat_tri(int W, int H, float2 a, float2 b, float2 c, unsigned id, unsigned* z)
{
    // bbox floor(min) .. ceil(max) inclusive, clipped to the image (main.cpp:123-127)
    float minx = floorf(fminf(a.x, fminf(b.x, c.x))), miny = floorf(fminf(a.y, fminf(b.y, c.y)));
    float maxx = ceilf(fmaxf(a.x, fmaxf(b.x, c.x))), maxy = ceilf(fmaxf(a.y, fmaxf(b.y, c.y)));
    if (!(minx > -1e9f)) minx = -1e9f;
    if (!(miny > -1e9f)) miny = -1e9f;
    if (!(maxx < 1e9f)) maxx = 1e9f;
    if (!(maxy < 1e9f)) maxy = 1e9f;
    const int xs = max(0, (int)minx), ys = max(0, (int)miny);
    const int xe = min(W - 1, (int)maxx), ye = min(H - 1, (int)maxy);
    for (int y = ys; y <= ye; ++y)
        for (int x = xs; x <= xe; ++x)
            if (lk(a.x, a.y, b.x, b.y, c.x, c.y, (float)x, (float)y).in) atomicMax(&z[(size_t)y * W + x], id);
}

// pass 1: one thread per pixel quad (x, y), x + 1 < W, y + 1 < H, all four corners on the object
__global__ void __launch_bounds__(256) k_splat(int W, int H, const float2* __restrict__ pos,
                                                const unsigned char* __restrict__ mask_red,
                                                unsigned* __restrict__ z)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x + 1 >= W || y + 1 >= H) return;
    const size_t i00 = (size_t)y * W + x, i01 = i00 + 1, i10 = i00 + W, i11 = i10 + 1;
    if (mask_red[i00] | mask_red[i01] | mask_red[i10] | mask_red[i11]) return; // main.cpp:178, 191-196
    const float2 p00 = pos[i00], p01 = pos[i01], p10 = pos[i10], p11 = pos[i11];
    const unsigned id = (unsigned)(2 * i00 + 1);
    splat_tri(W, H, p00, p01, p10, id, z);     // (pos00, pos01, pos10)  main.cpp:197-199
    splat_tri(W, H, p10, p01, p11, id + 1, z); // (pos10, pos01, pos11)  main.cpp:200-202
}

// pass 2: one thread per output pixel
__global__ void __launch_bounds__(256) k_resolve(int W, int H, const float2* __restrict__ pos,
                                                  const unsigned char* __restrict__ rgb,
                                                  const unsigned* __restrict__ z, unsigned char* __restrict__ out_rgb,
                                                  unsigned char* __restrict__ out_mask)
{
    const size_t N = (size_t)W * H;
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const unsigned id = z[o];
    unsigned char c0 = 0, c1 = 0, c2 = 0, m = 0;
    if (id) {
        const unsigned t = (id - 1) & 1u;
        const size_t i00 = (id - 1) >> 1, i01 = i00 + 1, i10 = i00 + W, i11 = i10 + 1;
        const size_t ia = t ? i10 : i00, ib = i01, ic = t ? i11 : i10;
        const float2 a = pos[ia], b = pos[ib], c = pos[ic];
        const int x = (int)(o % W), y = (int)(o / W);
        const Bary w = lk(a.x, a.y, b.x, b.y, c.x, c.y, (float)x, (float)y);
        // vec3f val = c0*b0 + c1*b1 + c2*b2, then the truncating vec3f -> vec3uc cast (vec3.h:32-37)
        float v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float ca = (float)rgb[3 * ia + k], cb = (float)rgb[3 * ib + k], cc = (float)rgb[3 * ic + k];
            v[k] = __fadd_rn(__fadd_rn(__fmul_rn(ca, w.b0), __fmul_rn(cb, w.b1)), __fmul_rn(cc, w.b2));
        }
        c0 = (unsigned char)__float2int_rz(v[0]);
        c1 = (unsigned char)__float2int_rz(v[1]);
        c2 = (unsigned char)__float2int_rz(v[2]);
        m = 255;
    }
    out_rgb[3 * o] = c0;
    out_rgb[3 * o + 1] = c1;
    out_rgb[3 * o + 2] = c2;
    out_mask[o] = m;
}

// positions from a flow field (main.cpp:160-166); also used for flow extraction X - grid
// (CombinedSolver.h:352-366) with sign = -1
__global__ void __launch_bounds__(256) k_grid_add(int W, int H, const float2* __restrict__ in, float2* __restrict__ out,
                                                   float sign)
{
    const size_t N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float gx = (float)(int)(i % W), gy = (float)(int)(i / W);
    const float2 v = in[i];
    float2 r;
    if (sign > 0.f) {
        r.x = __fadd_rn(gx, v.x);
        r.y = __fadd_rn(gy, v.y);
    } else {
        r.x = __fsub_rn(v.x, gx);
        r.y = __fsub_rn(v.y, gy);
    }
    out[i] = r;
}

} // namespace

int warp_launches_per_call() { return 2; }

void enqueue_warp(int W, int H, const float2* d_pos, const unsigned char* d_rgb, const unsigned char* d_mask_red,
                  unsigned* d_z, unsigned char* d_out_rgb, unsigned char* d_out_mask, cudaStream_t stream)
{
    const size_t N = (size_t)W * H;
    ARAP_CUDA_CHECK(cudaMemsetAsync(d_z, 0, N * sizeof(unsigned), stream));
    dim3 g1((W + 31) / 32, (H + 7) / 8);
    k_splat<<<g1, 256, 0, stream>>>(W, H, d_pos, d_mask_red, d_z);
    k_resolve<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(W, H, d_pos, d_rgb, d_z, d_out_rgb, d_out_mask);
    ARAP_CUDA_CHECK(cudaGetLastError());
}

void enqueue_flow_to_pos(int W, int H, const float2* d_flow, float2* d_pos, cudaStream_t stream)
{
    const size_t N = (size_t)W * H;
    k_grid_add<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(W, H, d_flow, d_pos, 1.f);
    ARAP_CUDA_CHECK(cudaGetLastError());
}

void enqueue_pos_to_flow(int W, int H, const float2* d_pos, float2* d_flow, cudaStream_t stream)
{
    const size_t N = (size_t)W * H;
    k_grid_add<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(W, H, d_pos, d_flow, -1.f);
    ARAP_CUDA_CHECK(cudaGetLastError());
}

} // namespace arapb200
