// pipeline.cu -- see pipeline.cuh
#include "pipeline.cuh"
#include <cmath>
#include <cstring>
#include <unordered_map>

namespace arapb200 {

void build_match_records(int W, int H, const unsigned char* mask_red, const int* matches, int n_matches,
                         std::vector<MatchRec>& out)
{
    out.clear();
    std::unordered_map<int, int> where; // pixel -> position in out (later entries override earlier)
    auto add = [&](int x1, int y1, int x2, int y2) {
        if (x1 < 0 || x1 >= W || y1 < 0 || y1 >= H) return; // the reference would index out of bounds
        const int idx = y1 * W + x1;
        if (mask_red[idx] != 0) return; // CombinedSolver.h:234
        MatchRec r{idx, (float)x1, (float)y1, (float)x2, (float)y2};
        auto it = where.find(idx);
        if (it == where.end()) {
            where.emplace(idx, (int)out.size());
            out.push_back(r);
        } else {
            out[it->second] = r;
        }
    };
    for (int k = 0; k < n_matches; ++k) add(matches[4 * k], matches[4 * k + 1], matches[4 * k + 2], matches[4 * k + 3]);
    // border pins in row-major order (main.cpp:130-136); they come last, so they override
    for (int y = 0; y < H; ++y) {
        if (y == 0 || y == H - 1) {
            for (int x = 0; x < W; ++x) add(x, y, x, y);
        } else {
            add(0, y, 0, y);
            if (W > 1) add(W - 1, y, W - 1, y);
        }
    }
}

namespace {

// resetGPU (CombinedSolver.h:207-221): UrShape = X = pixel grid, Angle = 0, Mask = float(red)
__global__ void __launch_bounds__(256) k_reset(int W, int H, const unsigned char* __restrict__ mask_red,
                                                float2* __restrict__ X, float2* __restrict__ U,
                                                float* __restrict__ A, float* __restrict__ M)
{
    const size_t N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float2 g = make_float2((float)(int)(i % W), (float)(int)(i / W));
    X[i] = g;
    U[i] = g;
    A[i] = 0.f;
    M[i] = (float)mask_red[i];
}

__global__ void __launch_bounds__(256) k_fill_c(size_t N, float2* __restrict__ C)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) C[i] = make_float2(-1.f, -1.f);
}

// setConstraintImage (CombinedSolver.h:223-242) on the compact, already filtered match list
__global__ void __launch_bounds__(256) k_scatter_c(const MatchRec* __restrict__ m, int n, float alpha,
                                                    float2* __restrict__ C)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const MatchRec r = m[k];
    const float om = 1.0f - alpha;
    C[r.idx] = make_float2(om * r.x1 + alpha * r.x2, om * r.y1 + alpha * r.y2);
}

// resident back-end: per-pixel match TARGET image; the kernel applies the continuation lerp itself
__global__ void __launch_bounds__(256) k_fill_t(size_t N, float2* __restrict__ C)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) C[i] = make_float2(-1e30f, -1e30f);
}
__global__ void __launch_bounds__(256) k_scatter_t(const MatchRec* __restrict__ m, int n, float2* __restrict__ C)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const MatchRec r = m[k];
    C[r.idx] = make_float2(r.x2, r.y2);
}

__global__ void k_copy_cost(const StreamScalars* __restrict__ sc, float* __restrict__ dst) { *dst = sc->cost; }

} // namespace

void enqueue_reset_state(int W, int H, const unsigned char* d_mask_red, float2* d_X, float2* d_U, float* d_A,
                         float* d_M, cudaStream_t stream)
{
    const size_t N = (size_t)W * H;
    k_reset<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(W, H, d_mask_red, d_X, d_U, d_A, d_M);
}

void enqueue_constraint_image(int W, int H, const MatchRec* d_matches, int n, float alpha, float2* d_C,
                              cudaStream_t stream)
{
    const size_t N = (size_t)W * H;
    k_fill_c<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(N, d_C);
    if (n > 0) k_scatter_c<<<(n + 255) / 256, 256, 0, stream>>>(d_matches, n, alpha, d_C);
}

void enqueue_target_image(int W, int H, const MatchRec* d_matches, int n, float2* d_C, cudaStream_t stream)
{
    const size_t N = (size_t)W * H;
    k_fill_t<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(N, d_C);
    if (n > 0) k_scatter_t<<<(n + 255) / 256, 256, 0, stream>>>(d_matches, n, d_C);
}

DeformPipeline::DeformPipeline(int maxW, int maxH, int nCont, int nGN, int nPCG, int backend)
    : maxW_(maxW), maxH_(maxH), nCont_(nCont), nGN_(nGN), nPCG_(nPCG), backend_(backend)
{
    const size_t N = (size_t)maxW * maxH;
    ARAP_CUDA_OR_EXIT(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    for (auto& e : ev_) ARAP_CUDA_OR_EXIT(cudaEventCreate(&e));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_X_, N * sizeof(float2)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_U_, N * sizeof(float2)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_C_, N * sizeof(float2)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_flow_, N * sizeof(float2)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_A_, N * sizeof(float)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_M_, N * sizeof(float)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_costs_, (size_t)nCont * (nGN + 1) * sizeof(float)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_rgb_, 3 * N));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_mask_, N));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_orgb_, 3 * N));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_omask_, N));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_z_, N * sizeof(unsigned)));
    h_in_bytes_ = 4 * N;
    h_out_bytes_ = N * sizeof(float2) + 4 * N + (size_t)nCont * (nGN + 1) * sizeof(float);
    ARAP_CUDA_OR_EXIT(cudaMallocHost(&h_in_, h_in_bytes_));
    ARAP_CUDA_OR_EXIT(cudaMallocHost(&h_out_, h_out_bytes_));
}

DeformPipeline::~DeformPipeline()
{
    cudaStreamSynchronize(stream_);
    delete solver_;
    delete resident_;
    cudaFree(d_X_); cudaFree(d_U_); cudaFree(d_C_); cudaFree(d_flow_); cudaFree(d_A_); cudaFree(d_M_);
    cudaFree(d_costs_); cudaFree(d_rgb_); cudaFree(d_mask_); cudaFree(d_orgb_); cudaFree(d_omask_);
    cudaFree(d_z_); cudaFree(d_matches_);
    cudaFreeHost(h_in_); cudaFreeHost(h_out_);
    for (auto& e : ev_) cudaEventDestroy(e);
    cudaStreamDestroy(stream_);
}

int DeformPipeline::run(const HostProblem& hp)
{
    const int W = hp.W, H = hp.H;
    if (W <= 0 || H <= 0 || (size_t)W * H > (size_t)maxW_ * maxH_) {
        fprintf(stderr, "arapb200: problem %dx%d does not fit the pipeline (%dx%d)\n", W, H, maxW_, maxH_);
        return 1;
    }
    const size_t N = (size_t)W * H;
    if (curW_ != W || curH_ != H) { // re-plan on a size change (CombinedSolver.h:149-160)
        delete solver_;
        solver_ = nullptr;
        curW_ = W;
        curH_ = H;
    }
    if (backend_ != ARAPB200_BACKEND_STREAM && !resident_) resident_ = new ResidentSolver(maxW_, maxH_);
    const long long l0 = (solver_ ? solver_->launches() : 0) + (resident_ ? resident_->launches() : 0);
    std::vector<MatchRec> recs;
    build_match_records(W, H, hp.mask_red, hp.matches, hp.n_matches, recs);
    if (recs.size() > matches_cap_) {
        cudaFree(d_matches_);
        matches_cap_ = recs.size() * 2 + 1024;
        ARAP_CUDA_OR_RETURN(cudaMalloc(&d_matches_, matches_cap_ * sizeof(MatchRec)));
    }
    // stage inputs through pinned memory
    memcpy(h_in_, hp.rgb, 3 * N);
    memcpy(h_in_ + 3 * N, hp.mask_red, N);
    ARAP_CUDA_OR_RETURN(cudaEventRecord(ev_[0], stream_));
    ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(d_rgb_, h_in_, 3 * N, cudaMemcpyHostToDevice, stream_));
    ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(d_mask_, h_in_ + 3 * N, N, cudaMemcpyHostToDevice, stream_));
    if (!recs.empty())
        ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(d_matches_, recs.data(), recs.size() * sizeof(MatchRec),
                                            cudaMemcpyHostToDevice, stream_));
    enqueue_reset_state(W, H, d_mask_, d_X_, d_U_, d_A_, d_M_, stream_);
    launches_ += 1;
    // weights: CombinedSolver.h:172-177
    const float wf = sqrtf(100.0f), wr = sqrtf(0.01f);
    bool use_res = false;
    if (backend_ != ARAPB200_BACKEND_STREAM) {
        use_res = resident_->prepare(W, H, d_M_, stream_);
        if (!use_res && backend_ == ARAPB200_BACKEND_RESIDENT) {
            fprintf(stderr, "arapb200: problem %dx%d (%d strips) does not fit the resident back-end\n", W, H,
                    resident_->n_strips());
            return 3;
        }
    }
    last_resident_ = use_res;
    if (!use_res && !solver_) solver_ = new StreamSolver(W, H);
    if (!use_res) solver_->bind(d_X_, d_A_, d_U_, d_C_, d_M_, wf, wr, stream_);
    ARAP_CUDA_OR_RETURN(cudaEventRecord(ev_[1], stream_));
    if (use_res) {
        // the whole continuation schedule (CombinedSolverBase.h:99-120) is ONE kernel launch
        enqueue_target_image(W, H, d_matches_, (int)recs.size(), d_C_, stream_);
        launches_ += recs.empty() ? 1 : 2;
        resident_->enqueue(d_X_, d_A_, d_C_, 1, wf, wr, nCont_, nGN_, nPCG_, d_costs_, nullptr, stream_);
    }
    for (int t = 0; t < nCont_ && !use_res; ++t) {
        const float alpha = (float)(t + 1) / (float)nCont_; // CombinedSolver.h:199-201
        enqueue_constraint_image(W, H, d_matches_, (int)recs.size(), alpha, d_C_, stream_);
        launches_ += recs.empty() ? 1 : 2;
        solver_->enqueue_init(stream_);
        k_copy_cost<<<1, 1, 0, stream_>>>(solver_->d_scalars(), d_costs_ + (size_t)t * (nGN_ + 1));
        for (int g = 0; g < nGN_; ++g) {
            solver_->enqueue_gn_step(nPCG_, stream_);
            k_copy_cost<<<1, 1, 0, stream_>>>(solver_->d_scalars(), d_costs_ + (size_t)t * (nGN_ + 1) + g + 1);
        }
        launches_ += 1 + nGN_;
    }
    ARAP_CUDA_OR_RETURN(cudaEventRecord(ev_[2], stream_));
    enqueue_pos_to_flow(W, H, d_X_, d_flow_, stream_);
    enqueue_warp(W, H, d_X_, d_rgb_, d_mask_, d_z_, d_orgb_, d_omask_, stream_);
    launches_ += 1 + warp_launches_per_call();
    ARAP_CUDA_OR_RETURN(cudaEventRecord(ev_[3], stream_));
    unsigned char* o = h_out_;
    ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(o, d_flow_, N * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
    ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(o + 8 * N, d_orgb_, 3 * N, cudaMemcpyDeviceToHost, stream_));
    ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(o + 11 * N, d_omask_, N, cudaMemcpyDeviceToHost, stream_));
    const size_t cbytes = (size_t)nCont_ * (nGN_ + 1) * sizeof(float);
    ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(o + 12 * N, d_costs_, cbytes, cudaMemcpyDeviceToHost, stream_));
    ARAP_CUDA_OR_RETURN(cudaStreamSynchronize(stream_));
    if (use_res) {
        if (int st = resident_->status(stream_)) {
            fprintf(stderr, "arapb200: resident solver aborted (watchdog, code %d)\n", st);
            return 4;
        }
    } else {
        unsigned bad = 0;
        solver_->read_back(stream_, nullptr, &bad);
        if (bad) {
            fprintf(stderr, "arapb200: internal error: UrShape check failed\n");
            return 2;
        }
    }
    if (hp.out_flow) memcpy(hp.out_flow, o, N * sizeof(float2));
    if (hp.out_rgb) memcpy(hp.out_rgb, o + 8 * N, 3 * N);
    if (hp.out_mask) memcpy(hp.out_mask, o + 11 * N, N);
    if (hp.out_costs) memcpy(hp.out_costs, o + 12 * N, cbytes);
    launches_ += (solver_ ? solver_->launches() : 0) + (resident_ ? resident_->launches() : 0) - l0;
    cudaEventElapsedTime(&ms_total_, ev_[0], ev_[3]);
    cudaEventElapsedTime(&ms_solve_, ev_[1], ev_[2]);
    cudaEventElapsedTime(&ms_warp_, ev_[2], ev_[3]);
    return 0;
}

} // namespace arapb200
