// plan.cu -- see plan.cuh
#include "plan.cuh"
#include <cstring>

namespace arapb200 {

GnPlan::GnPlan(int W, int H, int verbosity, int backend)
    : W_(W), H_(H), verbosity_(verbosity), backend_(backend), stream_(W, H)
{
    ARAP_CUDA_OR_EXIT(cudaStreamCreateWithFlags(&stream_h_, cudaStreamNonBlocking));
}

GnPlan::~GnPlan()
{
    if (stream_h_) {
        cudaStreamSynchronize(stream_h_);
        cudaStreamDestroy(stream_h_);
    }
}

bool GnPlan::set_parameter(const char* name, const void* value)
{
    if (strcmp(name, "nIterations") == 0) { n_iterations_ = *(const int*)value; return true; }
    if (strcmp(name, "lIterations") == 0) { l_iterations_ = *(const int*)value; return true; }
    // Levenberg-Marquardt knobs of SolverParameters (:26-39): valid names, unused by gaussNewtonGPU
    static const char* lm[] = {"residual_reset_period", "min_relative_decrease", "min_trust_region_radius",
                               "max_trust_region_radius", "q_tolerance", "function_tolerance",
                               "trust_region_radius", "radius_decrease_factor", "min_lm_diagonal",
                               "max_lm_diagonal"};
    for (const char* k : lm)
        if (strcmp(name, k) == 0) return true;
    return false;
}

void GnPlan::bind(void** pp)
{
    // arap_plan.t:2-8: Offset, Angle, UrShape, Constraints, Mask, w_fitSqrt, w_regSqrt
    // (scalars are dereferenced on the host at init AND at every step: util.t:664-692)
    stream_.bind((float2*)pp[0], (float*)pp[1], (const float2*)pp[2], (const float2*)pp[3], (const float*)pp[4],
                 *(const float*)pp[5], *(const float*)pp[6], stream_h_);
}

void GnPlan::check_grid(unsigned bad_u) const
{
    if (bad_u) {
        fprintf(stderr,
                "arapb200: UrShape differs from the pixel grid on %u active pixels; this build only supports "
                "the grid the ARAP app uploads (CombinedSolver.h:207-221)\n", bad_u);
        exit(1);
    }
}

void GnPlan::init(void** pp)
{
    bind(pp);
    n_iter_ = 0;
    stream_.enqueue_init(stream_h_);
    unsigned bad = 0;
    stream_.read_back(stream_h_, &prev_cost_, &bad);
    check_grid(bad);
}

int GnPlan::step(void** pp)
{
    if (n_iter_ >= n_iterations_) return 0;
    bind(pp);
    float* tr = d_trace_ ? d_trace_ + (size_t)3 * l_iterations_ * n_iter_ : nullptr;
    stream_.enqueue_gn_step(l_iterations_, stream_h_, tr);
    float c;
    stream_.read_back(stream_h_, &c, nullptr);
    if (verbosity_ > 0) printf("cost: %f -> %f\n", prev_cost_, c); // :1158-1163
    prev_cost_ = c;
    ++n_iter_;
    return 1;
}

void GnPlan::solve(void** pp)
{
    bind(pp);
    n_iter_ = 0;
    stream_.enqueue_init(stream_h_);
    if (verbosity_ > 0) { // keep the per-step prints: go through the stepwise path
        unsigned bad = 0;
        stream_.read_back(stream_h_, &prev_cost_, &bad);
        check_grid(bad);
        while (step(pp)) {}
        return;
    }
    for (; n_iter_ < n_iterations_; ++n_iter_) {
        float* tr = d_trace_ ? d_trace_ + (size_t)3 * l_iterations_ * n_iter_ : nullptr;
        stream_.enqueue_gn_step(l_iterations_, stream_h_, tr);
    }
    unsigned bad = 0;
    stream_.read_back(stream_h_, &prev_cost_, &bad);
    check_grid(bad);
}

} // namespace arapb200
