// pipeline.cuh -- per-image path of arap_deform on the device, batched: what CombinedSolver (ARAP/
// deformation/src/CombinedSolver.h:139-242, 280-366) and CombinedSolverBase::singleSolve (ARAP/shared/
// CombinedSolverBase.h:99-120) do around the solver, for many independent (image, segment) problems at once,
// with every host<->device hop removed except the upload of the inputs and the download of the outputs.
#pragma once
#include "plan.cuh"
#include "warp.cuh"
#include <vector>

namespace arapb200 {

struct HostProblem {
    int W = 0, H = 0;
    const unsigned char* rgb = nullptr;      // uint8x3[N]
    const unsigned char* mask_red = nullptr; // uint8[N]
    const int* matches = nullptr;            // int32[4*n], without border pins
    int n_matches = 0;
    float* out_flow = nullptr;               // float2[N]
    unsigned char* out_rgb = nullptr;        // uint8x3[N]
    unsigned char* out_mask = nullptr;       // uint8[N]
    float* out_costs = nullptr;              // float[nCont*(nGN+1)] or null
};

// compact constraint record: source pixel index and the match, after the reference's filtering
// (mask(src) == 0, later entries override earlier ones, border pins appended: main.cpp:130-136,
// CombinedSolver.h:223-242)
struct MatchRec {
    int idx;
    float x1, y1, x2, y2;
};

// host-side: filter + dedupe the match list exactly like setConstraintImage would resolve it
void build_match_records(int W, int H, const unsigned char* mask_red, const int* matches, int n_matches,
                         std::vector<MatchRec>& out);

class BatchPipeline {
public:
    // up to max_problems problems of at most maxW x maxH in flight
    BatchPipeline(int maxW, int maxH, int max_problems, int nCont, int nGN, int nPCG, int backend);
    ~BatchPipeline();
    BatchPipeline(const BatchPipeline&) = delete;
    BatchPipeline& operator=(const BatchPipeline&) = delete;
    // upload, solve (all continuation steps), flow, warp, download -- for every problem given.  Blocking.
    int run(const HostProblem* problems, int count);
    long long launches() const { return launches_; }
    float last_ms_total() const { return ms_total_; }
    float last_ms_solve() const { return ms_solve_; }
    float last_ms_warp() const { return ms_warp_; }
    int last_resident_count() const { return n_resident_; }
    int last_group_size() const { return group_size_; }
    // {problems in the largest cooperative launch, variant max threads, variant min CTAs/SM, grid.x, grid.y, threads} of
    // the last resident launch of the last run (zeros if everything streamed)
    void last_launch_info(int out[6]) const
    {
        out[0] = group_size_;
        int sh[5] = {0, 0, 0, 0, 0};
        if (resident_ && n_resident_ > 0) resident_->last_launch_shape(sh);
        for (int i = 0; i < 5; ++i) out[1 + i] = sh[i];
    }
    int max_problems() const { return (int)dev_.size(); }
    // opt-in early exits of the PCG loops / of the Gauss-Newton steps (both back-ends)
    void set_pcg_rtol(float rtol);
    void set_gn_rtol(float rtol);
    void set_cluster_barrier(bool on) { if (resident_) resident_->set_cluster_barrier(on); }
    // opt-in: every Opt_ProblemSolve of the schedule runs the reference's other solver kind, "LMGPU" (solver_lm.cuh:
    // trust region + Q-based exit of the linear loops) with its default parameters, one problem at a time
    void set_lm(bool on) { lm_on_ = on; }

private:
    struct Dev { // device + pinned staging of one problem slot
        float2 *X = nullptr, *U = nullptr, *C = nullptr, *flow = nullptr;
        float *A = nullptr, *M = nullptr, *costs = nullptr;
        unsigned char *rgb = nullptr, *mask = nullptr, *orgb = nullptr, *omask = nullptr;
        unsigned* z = nullptr;
        MatchRec* matches = nullptr;
        size_t matches_cap = 0;
        unsigned char *h_in = nullptr, *h_out = nullptr;
        std::vector<MatchRec> recs;
        bool resident = false;
    };
    void solve_streaming(const HostProblem& hp, Dev& d);
    void solve_lm(const HostProblem& hp, Dev& d);
    LmSolver* lm_ = nullptr;
    int lmW_ = 0, lmH_ = 0;
    bool lm_on_ = false;
    std::vector<float> lm_costs_;
    int maxW_, maxH_, nCont_, nGN_, nPCG_, backend_;
    std::vector<Dev> dev_;
    StreamSolver* solver_ = nullptr;
    int solverW_ = 0, solverH_ = 0;
    ResidentSolver* resident_ = nullptr;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
    long long launches_ = 0;
    float ms_total_ = 0, ms_solve_ = 0, ms_warp_ = 0;
    int n_resident_ = 0, group_size_ = 0;
    float pcg_rtol_ = 0.0f, gn_rtol_ = 0.0f;
};

// shared small kernels
void enqueue_reset_state(int W, int H, const unsigned char* d_mask_red, float2* d_X, float2* d_U, float* d_A,
                         float* d_M, cudaStream_t stream);
// resident back-end: per-pixel match target image ((-1e30, -1e30) = none); the kernel does the lerp
void enqueue_target_image(int W, int H, const MatchRec* d_matches, int n, float2* d_C, cudaStream_t stream);
void enqueue_constraint_image(int W, int H, const MatchRec* d_matches, int n, float alpha, float2* d_C,
                              cudaStream_t stream);

} // namespace arapb200
