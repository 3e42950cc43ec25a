// stream_probe.cu -- what limits the streaming PCG kernels at 1080p?  Same traffic as k_step_b (14 planes read, 6 written,
// 80 B/px) under different thread mappings and reduction tails.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// -o gpurun_out/stream_probe tools/probes/stream_probe.cu ; run on the GPU box, prints GB/s per variant.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct P { float* pl[20]; int W, H, tx, ty; double2* partials; unsigned* counter; unsigned long long* limbs; float* out; };

__device__ __forceinline__ float body(const P& p, size_t i0, size_t stride, int rows, float alpha)
{
    float v[4][14];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const size_t i = i0 + (size_t)min(r, rows - 1) * stride;
#pragma unroll
        for (int k = 0; k < 14; ++k) v[r][k] = p.pl[k][i];
    }
    float g = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (r < rows) {
            const size_t i = i0 + (size_t)r * stride;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                p.pl[9 + k][i] = fmaf(alpha, v[r][k], v[r][9 + k]);
                const float rr = fmaf(-alpha, v[r][3 + k], v[r][6 + k]);
                p.pl[6 + k][i] = rr;
                g = fmaf(v[r][12 + (k >> 1)] * rr, rr, g);
            }
        }
    }
    return g;
}

__device__ __forceinline__ float block_sum(float g, float* sm)
{
    for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(~0u, g, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = g;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(~0u, t, o);
    }
    return t;
}

__device__ void tail_fence(const P& p, float t, float* sm)
{
    __shared__ bool last;
    if (threadIdx.x == 0) {
        p.partials[blockIdx.x] = make_double2((double)t, 0.0);
        __threadfence();
        last = atomicInc(p.counter, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double s = 0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += __ldcg(&p.partials[i].x);
    float f = block_sum((float)s, sm);
    if (threadIdx.x == 0) *p.out = f;
}

__device__ void tail_red(const P& p, float t)
{
    if (threadIdx.x == 0) {
        // 4 limbs, spread over 32 copies to keep same-address atomics short
        const long long q = (long long)((double)t * 1048576.0);
        unsigned long long* L = p.limbs + (size_t)(blockIdx.x & 31) * 16;
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(L), "l"((unsigned long long)(q & 0xffffff)) : "memory");
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(L + 1), "l"((unsigned long long)((q >> 24) & 0xffffff)) : "memory");
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(L + 2), "l"((unsigned long long)((q >> 48) & 0xffff)) : "memory");
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(L + 3), "l"(1ull) : "memory");
    }
}

// mode: 0 none, 1 block sum only, 2 fence tail, 3 red tail
template <int MODE>
__global__ void __launch_bounds__(256) k_tile32(P p, float alpha)
{
    __shared__ float sm[32];
    const int x = (blockIdx.x % p.tx) * 32 + (threadIdx.x & 31);
    const int yb = (blockIdx.x / p.tx) * 32 + (threadIdx.x >> 5) * 4;
    float g = 0.f;
    if (x < p.W && yb < p.H) g = body(p, (size_t)yb * p.W + x, p.W, min(4, p.H - yb), alpha);
    if (MODE == 0) { if (g == 123.456f) *p.out = g; return; }
    float t = block_sum(g, sm);
    if (MODE == 1) { if (threadIdx.x == 0 && t == 123.456f) *p.out = t; return; }
    if (MODE == 2) tail_fence(p, t, sm);
    if (MODE == 3) tail_red(p, t);
}

// 256 x 4 strips: block = 256 consecutive columns, 4 rows
template <int MODE>
__global__ void __launch_bounds__(256) k_strip(P p, float alpha)
{
    __shared__ float sm[32];
    const int sx = (p.W + 255) / 256;
    const int x = (blockIdx.x % sx) * 256 + threadIdx.x;
    const int yb = (blockIdx.x / sx) * 4;
    float g = 0.f;
    if (x < p.W && yb < p.H) g = body(p, (size_t)yb * p.W + x, p.W, min(4, p.H - yb), alpha);
    if (MODE == 0) { if (g == 123.456f) *p.out = g; return; }
    float t = block_sum(g, sm);
    if (MODE == 2) tail_fence(p, t, sm);
    if (MODE == 3) tail_red(p, t);
}

// linear: every thread one float4 of each plane (no quad structure) -- the copy-like ideal
__global__ void __launch_bounds__(256) k_linear(P p, float alpha, size_t n4)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    float4 v[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) v[k] = reinterpret_cast<const float4*>(p.pl[k])[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float4 d = v[9 + k], r = v[6 + k];
        d.x = fmaf(alpha, v[k].x, d.x); d.y = fmaf(alpha, v[k].y, d.y); d.z = fmaf(alpha, v[k].z, d.z); d.w = fmaf(alpha, v[k].w, d.w);
        r.x = fmaf(-alpha, v[3 + k].x, r.x); r.y = fmaf(-alpha, v[3 + k].y, r.y); r.z = fmaf(-alpha, v[3 + k].z, r.z); r.w = fmaf(-alpha, v[3 + k].w, r.w);
        r.x += v[12 + (k >> 1)].x * 1e-30f;
        reinterpret_cast<float4*>(p.pl[9 + k])[i] = d;
        reinterpret_cast<float4*>(p.pl[6 + k])[i] = r;
    }
}

// persistent: grid = SMs * k CTAs, each loops over 32x32 tiles with stride gridDim
template <int MODE>
__global__ void __launch_bounds__(256) k_persist(P p, float alpha)
{
    __shared__ float sm[32];
    float g = 0.f;
    for (int t = blockIdx.x; t < p.tx * p.ty; t += gridDim.x) {
        const int x = (t % p.tx) * 32 + (threadIdx.x & 31);
        const int yb = (t / p.tx) * 32 + (threadIdx.x >> 5) * 4;
        if (x < p.W && yb < p.H) g += body(p, (size_t)yb * p.W + x, p.W, min(4, p.H - yb), alpha);
    }
    float t = block_sum(g, sm);
    if (MODE == 2) tail_fence(p, t, sm);
    if (MODE == 3) tail_red(p, t);
}

int main(int argc, char** argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 1920, H = argc > 2 ? atoi(argv[2]) : 1080;
    const size_t N = (size_t)W * H;
    P p{};
    p.W = W; p.H = H; p.tx = (W + 31) / 32; p.ty = (H + 31) / 32;
    for (int k = 0; k < 20; ++k) { CK(cudaMalloc(&p.pl[k], N * 4 + 1024)); CK(cudaMemset(p.pl[k], 0, N * 4 + 1024)); }
    CK(cudaMalloc(&p.partials, 65536 * sizeof(double2)));
    CK(cudaMalloc(&p.counter, 4)); CK(cudaMemset(p.counter, 0, 4));
    CK(cudaMalloc(&p.limbs, 32 * 16 * 8)); CK(cudaMemset(p.limbs, 0, 32 * 16 * 8));
    CK(cudaMalloc(&p.out, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 200;
    const double bytes = (double)N * 80.0;
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 10; ++i) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-28s %8.2f us/launch  %8.1f GB/s\n", name, 1000.0 * ms / reps, bytes * reps / (ms * 1e-3) / 1e9);
    };
    const int nt = p.tx * p.ty, sx = (W + 255) / 256, ns = sx * ((H + 3) / 4);
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("%dx%d: %d tiles, %d strips, %d SMs, %.1f MB per pass\n", W, H, nt, ns, sms, bytes / 1e6);
    run("linear float4", [&] { k_linear<<<(unsigned)((N / 4 + 255) / 256), 256>>>(p, 0.5f, N / 4); });
    run("tile32 no-reduce", [&] { k_tile32<0><<<nt, 256>>>(p, 0.5f); });
    run("tile32 block-sum", [&] { k_tile32<1><<<nt, 256>>>(p, 0.5f); });
    run("tile32 fence tail", [&] { k_tile32<2><<<nt, 256>>>(p, 0.5f); });
    run("tile32 red tail", [&] { k_tile32<3><<<nt, 256>>>(p, 0.5f); });
    run("strip256x4 no-reduce", [&] { k_strip<0><<<ns, 256>>>(p, 0.5f); });
    run("strip256x4 fence tail", [&] { k_strip<2><<<ns, 256>>>(p, 0.5f); });
    run("strip256x4 red tail", [&] { k_strip<3><<<ns, 256>>>(p, 0.5f); });
    for (int k = 1; k <= 4; ++k) {
        char nm[64];
        snprintf(nm, 64, "persist x%d fence tail", k);
        run(nm, [&] { k_persist<2><<<sms * k, 256>>>(p, 0.5f); });
        snprintf(nm, 64, "persist x%d red tail", k);
        run(nm, [&] { k_persist<3><<<sms * k, 256>>>(p, 0.5f); });
    }
    return 0;
}
