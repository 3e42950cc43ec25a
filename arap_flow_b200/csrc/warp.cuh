// warp.cuh -- forward-warp kernels (see warp.cu).  Device pointers, enqueue-only.
#pragma once
#include "common.cuh"

namespace arapb200 {

// z: uint32[W*H] scratch that ends up holding the splat index (1 + 2*(y*W+x) + t, 0 = empty)
void enqueue_warp(int W, int H, const float2* d_pos, const unsigned char* d_rgb, const unsigned char* d_mask_red,
                  unsigned* d_z, unsigned char* d_out_rgb, unsigned char* d_out_mask, cudaStream_t stream);
void enqueue_flow_to_pos(int W, int H, const float2* d_flow, float2* d_pos, cudaStream_t stream);
void enqueue_pos_to_flow(int W, int H, const float2* d_pos, float2* d_flow, cudaStream_t stream);
int warp_launches_per_call();

} // namespace arapb200
