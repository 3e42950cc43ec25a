// composite.cu -- "next" row N1: the step right after the warp in the reference's driver.
//   flatten (para_gen.py:136-175): the per-segment outputs of a --multseg pair are layered in segment order; where a
//     later segment's warped mask is non-zero its flow / colour / mask replace what is below (segment 0 is the base).
//   add_bg  (para_gen.py:50-61, 206-212): where the final mask is 0 the colour comes from the background image.
// One thread per pixel; the winning layer is the LAST segment whose mask is non-zero (else segment 0), so the result
// does not depend on any ordering between threads.  Pure select arithmetic: bit-identical to the numpy original.
#include "../../include/arapb200.h"
#include "common.cuh"

#include <mutex>

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

} // namespace
} // namespace arapb200

using namespace arapb200;

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
