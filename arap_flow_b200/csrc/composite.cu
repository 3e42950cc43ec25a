// composite.cu -- "next" row N1: the step right after the warp in the reference's driver.
//   flatten (para_gen.py:136-175): the per-segment outputs of a --multseg pair are layered in segment order; where a
//     later segment's warped mask is non-zero its flow / colour / mask replace what is below (segment 0 is the base).
//   add_bg  (para_gen.py:50-61, 206-212): where the final mask is 0 the colour comes from the background image.
//   match filter + segment masks (para_gen.py:216-223, 468-482, 513-537): "next" row N3, the step right before the solve.
// One thread per pixel; the winning layer is the LAST segment whose mask is non-zero (else segment 0), so the result
// does not depend on any ordering between threads.  Pure select arithmetic: bit-identical to the numpy original.
#include "../../include/arapb200.h"
#include "common.cuh"

#include <mutex>
#include <vector>

namespace arapb200 {
namespace {

constexpr int MAX_LAYERS = 32;
struct Layers {
    const float2* flow[MAX_LAYERS];
    const unsigned char* rgb[MAX_LAYERS];
    const unsigned char* mask[MAX_LAYERS];
    int n;
};

__global__ void __launch_bounds__(256) k_flatten(size_t N, Layers L, const unsigned char* __restrict__ bg,
                                                  float2* __restrict__ out_flow, unsigned char* __restrict__ out_rgb,
                                                  unsigned char* __restrict__ out_mask)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int win = 0;
    for (int s = L.n - 1; s >= 1; --s)
        if (L.mask[s][i] != 0) { win = s; break; }
    const unsigned char m = L.mask[win][i];
    out_flow[i] = L.flow[win][i];
    out_mask[i] = m;
    const unsigned char* src = (bg != nullptr && m == 0) ? bg : L.rgb[win];
    out_rgb[3 * i] = src[3 * i];
    out_rgb[3 * i + 1] = src[3 * i + 1];
    out_rgb[3 * i + 2] = src[3 * i + 2];
}

// ---- N3: match filter (para_gen.py:216-223) and per-segment masks (:513-537) ----
__global__ void __launch_bounds__(256) k_match_valid(int n, const int4* __restrict__ m, int W1, int H1,
                                                     const unsigned char* __restrict__ l1, int W2, int H2,
                                                     const unsigned char* __restrict__ l2, unsigned char* __restrict__ label)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 v = m[i];
    unsigned char keep = 0;
    if (v.x >= 0 && v.y >= 0 && v.z >= 0 && v.w >= 0 && v.x < W1 && v.z < W2 && v.y < H1 && v.w < H2) {
        const long long dx = (long long)v.z - v.x, dy = (long long)v.w - v.y;
        const long long d2 = dx * dx + dy * dy; // sqrt(d2) < 60 and > 0  <=>  0 < d2 < 3600 for integers
        const unsigned char a = l1[(size_t)v.y * W1 + v.x];
        if (d2 > 0 && d2 < 3600 && a > 0 && a == l2[(size_t)v.w * W2 + v.z]) keep = a;
    }
    label[i] = keep; // 0 = dropped (a kept match always has a non-zero label)
}

// order-preserving compaction by one block (n is a few thousand)
__global__ void __launch_bounds__(1024) k_match_compact(int n, const int4* __restrict__ m,
                                                        const unsigned char* __restrict__ label, int4* __restrict__ out_m,
                                                        unsigned char* __restrict__ out_l, int* __restrict__ count)
{
    __shared__ int wsum[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int start = 0; start < n; start += 1024) {
        const int i = start + threadIdx.x;
        const unsigned char l = (i < n) ? label[i] : 0;
        const unsigned b = __ballot_sync(0xffffffffu, l != 0);
        if (lane == 0) wsum[wid] = __popc(b);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            if (w < wid) woff += wsum[w];
            tot += wsum[w];
        }
        const int b0 = base;
        if (l != 0) {
            const int o = b0 + woff + __popc(b & ((1u << lane) - 1u));
            out_m[o] = m[i];
            out_l[o] = l;
        }
        __syncthreads();
        if (threadIdx.x == 0) base = b0 + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

__global__ void __launch_bounds__(256) k_segment_mask(size_t N, const unsigned char* __restrict__ labels, int segment,
                                                      unsigned char* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned char l = labels[i];
    const bool object = segment > 0 ? (l == (unsigned char)segment) : (l != 0);
    out[i] = object ? 0 : 255; // ARAP_BG = 255 (para_gen.py:30)
}

} // namespace
} // namespace arapb200

using namespace arapb200;

namespace {
struct Scratch { // RAII device allocations of one call
    std::vector<void*> p;
    ~Scratch() { for (void* d : p) cudaFree(d); }
    template <class T> T* get(size_t n)
    {
        void* d = nullptr;
        if (cudaMalloc(&d, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
        p.push_back(d);
        return (T*)d;
    }
};
} // namespace

extern "C" int arapb200_filter_matches(int W1, int H1, const uint8_t* labels1, int W2, int H2, const uint8_t* labels2,
                                       const int32_t* matches, int n, int32_t* out_matches, uint8_t* out_labels,
                                       int* n_out)
{
    if (W1 <= 0 || H1 <= 0 || W2 <= 0 || H2 <= 0 || !labels1 || !labels2 || n < 0 || (n > 0 && (!matches || !out_matches)) || !n_out)
        return 1;
    *n_out = 0;
    if (n == 0) return 0;
    Scratch s;
    const size_t N1 = (size_t)W1 * H1, N2 = (size_t)W2 * H2;
    unsigned char* d_l1 = s.get<unsigned char>(N1);
    unsigned char* d_l2 = s.get<unsigned char>(N2);
    int4* d_m = s.get<int4>(n);
    int4* d_o = s.get<int4>(n);
    unsigned char* d_lab = s.get<unsigned char>(n);
    unsigned char* d_ol = s.get<unsigned char>(n);
    int* d_cnt = s.get<int>(1);
    if (!d_l1 || !d_l2 || !d_m || !d_o || !d_lab || !d_ol || !d_cnt) return 2;
    ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(d_l1, labels1, N1, cudaMemcpyHostToDevice, nullptr));
    ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(d_l2, labels2, N2, cudaMemcpyHostToDevice, nullptr));
    ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(d_m, matches, (size_t)n * sizeof(int4), cudaMemcpyHostToDevice, nullptr));
    k_match_valid<<<(n + 255) / 256, 256>>>(n, d_m, W1, H1, d_l1, W2, H2, d_l2, d_lab);
    k_match_compact<<<1, 1024>>>(n, d_m, d_lab, d_o, d_ol, d_cnt);
    int cnt = 0;
    ARAP_CUDA_OR_RETURN(cudaMemcpy(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost));
    if (cnt > 0) {
        ARAP_CUDA_OR_RETURN(cudaMemcpy(out_matches, d_o, (size_t)cnt * sizeof(int4), cudaMemcpyDeviceToHost));
        if (out_labels) ARAP_CUDA_OR_RETURN(cudaMemcpy(out_labels, d_ol, (size_t)cnt, cudaMemcpyDeviceToHost));
    }
    *n_out = cnt;
    return 0;
}

extern "C" int arapb200_segment_mask(int W, int H, const uint8_t* labels, int segment, uint8_t* out_mask)
{
    if (W <= 0 || H <= 0 || !labels || !out_mask || segment < 0 || segment > 255) return 1;
    Scratch s;
    const size_t N = (size_t)W * H;
    unsigned char* d_l = s.get<unsigned char>(N);
    unsigned char* d_o = s.get<unsigned char>(N);
    if (!d_l || !d_o) return 2;
    ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(d_l, labels, N, cudaMemcpyHostToDevice, nullptr));
    k_segment_mask<<<(unsigned)((N + 255) / 256), 256>>>(N, d_l, segment, d_o);
    ARAP_CUDA_OR_RETURN(cudaMemcpy(out_mask, d_o, N, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int arapb200_flatten(int W, int H, int n_layers, const float* const* flows, const uint8_t* const* rgbs,
                                const uint8_t* const* masks, const uint8_t* background, float* out_flow,
                                uint8_t* out_rgb, uint8_t* out_mask)
{
    if (W <= 0 || H <= 0 || n_layers < 1 || n_layers > MAX_LAYERS || !flows || !rgbs || !masks) return 1;
    const size_t N = (size_t)W * H;
    Layers L{};
    L.n = n_layers;
    // One grow-only device arena per process (cudaMalloc / cudaFree of 15 multi-megabyte blocks per call cost far more
    // than the kernel): [layers: flow, rgb, mask][background][outputs], every block 256-byte aligned.
    static std::mutex mu;
    static unsigned char* arena = nullptr;
    static size_t arena_bytes = 0;
    static int arena_dev = -1;
    std::lock_guard<std::mutex> lock(mu);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t per_layer = al(N * sizeof(float2)) + al(3 * N) + al(N);
    const size_t need = (size_t)(n_layers + 1) * per_layer + al(3 * N);
    int dev = 0;
    int rc = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) rc = 2;
    if (!rc && (need > arena_bytes || dev != arena_dev)) {
        if (arena) { cudaFree(arena); arena = nullptr; arena_bytes = 0; }
        if (cudaMalloc(&arena, need) != cudaSuccess) rc = 2;
        else { arena_bytes = need; arena_dev = dev; }
    }
    unsigned char* cur = arena;
    auto take = [&](size_t bytes) { unsigned char* p = cur; cur += al(bytes); return p; };
    auto up = [&](const void* h, size_t bytes) -> void* {
        void* d = take(bytes);
        if (cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, nullptr) != cudaSuccess) rc = 2;
        return d;
    };
    for (int s = 0; s < n_layers && !rc; ++s) {
        L.flow[s] = (const float2*)up(flows[s], N * sizeof(float2));
        L.rgb[s] = (const unsigned char*)up(rgbs[s], 3 * N);
        L.mask[s] = (const unsigned char*)up(masks[s], N);
    }
    const unsigned char* d_bg = nullptr;
    if (!rc && background) d_bg = (const unsigned char*)up(background, 3 * N);
    if (!rc) {
        float2* d_of = (float2*)take(N * sizeof(float2));
        unsigned char* d_or = take(3 * N);
        unsigned char* d_om = take(N);
        k_flatten<<<(unsigned)((N + 255) / 256), 256>>>(N, L, d_bg, d_of, d_or, d_om);
        if (cudaMemcpy(out_flow, d_of, N * sizeof(float2), cudaMemcpyDeviceToHost) ||
            cudaMemcpy(out_rgb, d_or, 3 * N, cudaMemcpyDeviceToHost) || cudaMemcpy(out_mask, d_om, N, cudaMemcpyDeviceToHost))
            rc = 3;
    }
    if (rc) fprintf(stderr, "arapb200_flatten: CUDA failure (%s)\n", cudaGetErrorString(cudaGetLastError()));
    return rc;
}
