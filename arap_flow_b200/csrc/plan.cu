// plan.cu -- see plan.cuh
#include "plan.cuh"
#include <cstring>

namespace arapb200 {

namespace {
constexpr int kMaxCostLog = 4096;
// the resident and streaming kernels are specialised for UrShape == pixel grid: verify it
__global__ void __launch_bounds__(256) k_check_grid(int W, int H, const float2* __restrict__ U,
                                                     const float* __restrict__ M, unsigned* __restrict__ bad)
{
    const size_t N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || M[i] != 0.0f) return;
    const float2 u = U[i];
    if (u.x != (float)(int)(i % W) || u.y != (float)(int)(i / W)) atomicAdd(bad, 1u);
}
} // namespace

GnPlan::GnPlan(int W, int H, int verbosity, int backend)
    : W_(W), H_(H), verbosity_(verbosity), backend_(backend)
{
    ARAP_CUDA_OR_EXIT(cudaStreamCreateWithFlags(&stream_h_, cudaStreamNonBlocking));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_costs_, kMaxCostLog * sizeof(float)));
    ARAP_CUDA_OR_EXIT(cudaMalloc(&d_bad_u_, sizeof(unsigned)));
}

long long GnPlan::launches() const
{
    return (stream_ ? stream_->launches() : 0) + (resident_ ? resident_->launches() : 0);
}

GnPlan::~GnPlan()
{
    if (stream_h_) {
        cudaStreamSynchronize(stream_h_);
        cudaStreamDestroy(stream_h_);
    }
    cudaFree(d_costs_);
    cudaFree(d_bad_u_);
}

bool GnPlan::set_parameter(const char* name, const void* value)
{
    if (strcmp(name, "nIterations") == 0) { n_iterations_ = *(const int*)value; return true; }
    if (strcmp(name, "lIterations") == 0) { l_iterations_ = *(const int*)value; return true; }
    // extension: float, relative PCG tolerance; 0 (default) = the reference's fixed budget
    if (strcmp(name, "pcg_rtol") == 0) { pcg_rtol_ = *(const float*)value; return true; }
    // Levenberg-Marquardt knobs of SolverParameters (:26-39): valid names, unused by gaussNewtonGPU
    static const char* lm[] = {"residual_reset_period", "min_relative_decrease", "min_trust_region_radius",
                               "max_trust_region_radius", "q_tolerance", "function_tolerance",
                               "trust_region_radius", "radius_decrease_factor", "min_lm_diagonal",
                               "max_lm_diagonal"};
    for (const char* k : lm)
        if (strcmp(name, k) == 0) return true;
    return false;
}

// AUTO: resident when the active part of the image fits on chip, else streaming.  An UrShape image other than the pixel
// grid (never bound by the ARAP app, CombinedSolver.h:207-221, but legal for an Opt.h caller) takes the streaming
// back-end's general-d kernels.
void GnPlan::choose_backend(void** pp)
{
    const size_t N = (size_t)W_ * H_;
    ARAP_CUDA_OR_EXIT(cudaMemsetAsync(d_bad_u_, 0, sizeof(unsigned), stream_h_));
    k_check_grid<<<(unsigned)((N + 255) / 256), 256, 0, stream_h_>>>(W_, H_, (const float2*)pp[2], (const float*)pp[4],
                                                                     d_bad_u_);
    unsigned bad = 0;
    ARAP_CUDA_OR_EXIT(cudaMemcpyAsync(&bad, d_bad_u_, sizeof(unsigned), cudaMemcpyDeviceToHost, stream_h_));
    ARAP_CUDA_OR_EXIT(cudaStreamSynchronize(stream_h_));
    general_ = bad != 0;
    use_resident_ = false;
    if (backend_ != ARAPB200_BACKEND_STREAM && !general_) {
        if (!resident_) resident_.reset(new ResidentSolver(W_, H_));
        use_resident_ = resident_->prepare(W_, H_, (const float*)pp[4], stream_h_);
        if (!use_resident_ && backend_ == ARAPB200_BACKEND_RESIDENT) {
            fprintf(stderr, "arapb200: problem %dx%d (%d strips) does not fit the resident back-end\n", W_, H_,
                    resident_->n_strips());
            exit(1);
        }
    }
    if (!use_resident_) {
        if (!stream_) stream_.reset(new StreamSolver(W_, H_));
        stream_->set_general(general_);
    }
}

// resident launch: cost before and after each of nGN Gauss-Newton steps; returns the last cost
float GnPlan::run_resident(void** pp, int nGN, float* trace)
{
    if (nGN + 1 > kMaxCostLog) { fprintf(stderr, "arapb200: nIterations too large\n"); exit(1); }
    resident_->set_pcg_rtol(pcg_rtol_);
    resident_->enqueue((float2*)pp[0], (float*)pp[1], (const float2*)pp[3], 0, *(const float*)pp[5],
                       *(const float*)pp[6], 1, nGN, l_iterations_, d_costs_, trace, stream_h_);
    float c = 0.f;
    ARAP_CUDA_OR_EXIT(cudaMemcpyAsync(&c, d_costs_ + nGN, sizeof(float), cudaMemcpyDeviceToHost, stream_h_));
    ARAP_CUDA_OR_EXIT(cudaStreamSynchronize(stream_h_));
    if (int st = resident_->status(stream_h_)) {
        fprintf(stderr, "arapb200: resident solver aborted (watchdog, code %d)\n", st);
        exit(3);
    }
    return c;
}

void GnPlan::bind(void** pp)
{
    // arap_plan.t:2-8: Offset, Angle, UrShape, Constraints, Mask, w_fitSqrt, w_regSqrt
    // (scalars are dereferenced on the host at init AND at every step: util.t:664-692)
    stream_->bind((float2*)pp[0], (float*)pp[1], (const float2*)pp[2], (const float2*)pp[3], (const float*)pp[4],
                  *(const float*)pp[5], *(const float*)pp[6], stream_h_);
}

void GnPlan::init(void** pp)
{
    choose_backend(pp);
    n_iter_ = 0;
    if (use_resident_) {
        prev_cost_ = run_resident(pp, 0, nullptr);
        return;
    }
    bind(pp);
    stream_->enqueue_init(stream_h_);
    stream_->read_back(stream_h_, &prev_cost_, nullptr);
}

int GnPlan::step(void** pp)
{
    if (n_iter_ >= n_iterations_) return 0;
    float* tr = d_trace_ ? d_trace_ + (size_t)3 * l_iterations_ * n_iter_ : nullptr;
    float c;
    if (use_resident_) {
        c = run_resident(pp, 1, tr);
    } else {
        bind(pp);
        stream_->enqueue_gn_step(l_iterations_, stream_h_, tr);
        stream_->read_back(stream_h_, &c, nullptr);
    }
    if (verbosity_ > 0) printf("cost: %f -> %f\n", prev_cost_, c); // :1158-1163
    prev_cost_ = c;
    ++n_iter_;
    return 1;
}

void GnPlan::solve(void** pp)
{
    if (verbosity_ > 0) { // keep the per-step prints: go through the stepwise path
        init(pp);
        while (step(pp)) {}
        return;
    }
    choose_backend(pp);
    n_iter_ = 0;
    if (use_resident_) {
        prev_cost_ = run_resident(pp, n_iterations_, d_trace_); // ONE launch for the whole Opt_ProblemSolve
        n_iter_ = n_iterations_;
        return;
    }
    bind(pp);
    stream_->enqueue_init(stream_h_);
    for (; n_iter_ < n_iterations_; ++n_iter_) {
        float* tr = d_trace_ ? d_trace_ + (size_t)3 * l_iterations_ * n_iter_ : nullptr;
        stream_->enqueue_gn_step(l_iterations_, stream_h_, tr);
    }
    stream_->read_back(stream_h_, &prev_cost_, nullptr);
}

} // namespace arapb200
