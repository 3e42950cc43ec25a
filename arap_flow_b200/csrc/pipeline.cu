// pipeline.cu -- see pipeline.cuh
#include "pipeline.cuh"
#include <algorithm>
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

BatchPipeline::BatchPipeline(int maxW, int maxH, int max_problems, int nCont, int nGN, int nPCG, int backend)
    : maxW_(maxW), maxH_(maxH), nCont_(nCont), nGN_(nGN), nPCG_(nPCG), backend_(backend)
{
    const size_t N = (size_t)maxW * maxH;
    ARAP_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    for (auto& e : ev_) ARAP_CUDA_CHECK(cudaEventCreate(&e));
    dev_.resize(max_problems > 0 ? max_problems : 1);
    const size_t cbytes = (size_t)nCont * (nGN + 1) * sizeof(float);
    for (Dev& d : dev_) {
        ARAP_CUDA_CHECK(cudaMalloc(&d.X, N * sizeof(float2)));
        ARAP_CUDA_CHECK(cudaMalloc(&d.U, N * sizeof(float2)));
        ARAP_CUDA_CHECK(cudaMalloc(&d.C, N * sizeof(float2)));
        ARAP_CUDA_CHECK(cudaMalloc(&d.flow, N * sizeof(float2)));
        ARAP_CUDA_CHECK(cudaMalloc(&d.A, N * sizeof(float)));
        ARAP_CUDA_CHECK(cudaMalloc(&d.M, N * sizeof(float)));
        ARAP_CUDA_CHECK(cudaMalloc(&d.costs, cbytes));
        ARAP_CUDA_CHECK(cudaMalloc(&d.rgb, 3 * N));
        ARAP_CUDA_CHECK(cudaMalloc(&d.mask, N));
        ARAP_CUDA_CHECK(cudaMalloc(&d.orgb, 3 * N));
        ARAP_CUDA_CHECK(cudaMalloc(&d.omask, N));
        ARAP_CUDA_CHECK(cudaMalloc(&d.z, N * sizeof(unsigned)));
        ARAP_CUDA_CHECK(cudaMallocHost(&d.h_in, 4 * N));
        ARAP_CUDA_CHECK(cudaMallocHost(&d.h_out, 12 * N + cbytes));
    }
    if (backend_ != ARAPB200_BACKEND_STREAM) resident_ = new ResidentSolver(maxW, maxH, (int)dev_.size());
}

BatchPipeline::~BatchPipeline()
{
    cudaStreamSynchronize(stream_);
    delete solver_;
    delete resident_;
    delete lm_;
    for (Dev& d : dev_) {
        cudaFree(d.X); cudaFree(d.U); cudaFree(d.C); cudaFree(d.flow); cudaFree(d.A); cudaFree(d.M);
        cudaFree(d.costs); cudaFree(d.rgb); cudaFree(d.mask); cudaFree(d.orgb); cudaFree(d.omask);
        cudaFree(d.z); cudaFree(d.matches);
        cudaFreeHost(d.h_in); cudaFreeHost(d.h_out);
    }
    for (auto& e : ev_) cudaEventDestroy(e);
    cudaStreamDestroy(stream_);
}

// the streaming back-end: one graph launch per Gauss-Newton step (problems too large for the chip)
void BatchPipeline::solve_streaming(const HostProblem& hp, Dev& d)
{
    const int W = hp.W, H = hp.H;
    if (!solver_ || solverW_ != W || solverH_ != H) { // re-plan on a size change (CombinedSolver.h:149-160)
        if (solver_) launches_ += solver_->launches();
        delete solver_;
        solver_ = new StreamSolver(W, H);
        solverW_ = W;
        solverH_ = H;
    }
    const long long l0 = solver_->launches();
    const float wf = sqrtf(100.0f), wr = sqrtf(0.01f); // CombinedSolver.h:172-177
    solver_->set_pcg_rtol(pcg_rtol_);
    solver_->set_gn_rtol(gn_rtol_);
    solver_->bind(d.X, d.A, d.U, d.C, d.M, wf, wr, stream_);
    for (int t = 0; t < nCont_; ++t) {
        const float alpha = (float)(t + 1) / (float)nCont_; // CombinedSolver.h:199-201
        enqueue_constraint_image(W, H, d.matches, (int)d.recs.size(), alpha, d.C, stream_);
        launches_ += d.recs.empty() ? 1 : 2;
        solver_->enqueue_init(stream_);
        k_copy_cost<<<1, 1, 0, stream_>>>(solver_->d_scalars(), d.costs + (size_t)t * (nGN_ + 1));
        for (int g = 0; g < nGN_; ++g) {
            solver_->enqueue_gn_step(nPCG_, stream_);
            k_copy_cost<<<1, 1, 0, stream_>>>(solver_->d_scalars(), d.costs + (size_t)t * (nGN_ + 1) + g + 1);
        }
        launches_ += 1 + nGN_;
    }
    launches_ += solver_->launches() - l0;
}

// the "LMGPU" solver kind: per continuation step one Opt_ProblemSolve = init + up to nGN steps (o.t:2548-2551), the host
// deciding acceptance after each (solver_lm.cu).  Cost log: prevCost after init and after every step; entries of steps the
// solver did not take repeat the last one.
void BatchPipeline::solve_lm(const HostProblem& hp, Dev& d)
{
    const int W = hp.W, H = hp.H;
    if (!lm_ || lmW_ != W || lmH_ != H) {
        if (lm_) launches_ += lm_->launches();
        delete lm_;
        lm_ = nullptr;
        lm_ = new LmSolver(W, H);
        lmW_ = W;
        lmH_ = H;
    }
    const long long l0 = lm_->launches();
    const float wf = sqrtf(100.0f), wr = sqrtf(0.01f); // CombinedSolver.h:172-177
    lm_->bind(d.X, d.A, d.U, d.C, d.M, wf, wr);
    lm_costs_.assign((size_t)nCont_ * (nGN_ + 1), 0.0f);
    for (int t = 0; t < nCont_; ++t) {
        const float alpha = (float)(t + 1) / (float)nCont_; // CombinedSolver.h:199-201
        enqueue_constraint_image(W, H, d.matches, (int)d.recs.size(), alpha, d.C, stream_);
        launches_ += d.recs.empty() ? 1 : 2;
        float* c = lm_costs_.data() + (size_t)t * (nGN_ + 1);
        float prev = lm_->init(stream_);
        c[0] = prev;
        int g = 0;
        while (g < nGN_) {
            const int more = lm_->step(nPCG_, stream_, &prev);
            c[++g] = prev;
            if (!more) break;
        }
        for (int k = g + 1; k <= nGN_; ++k) c[k] = prev;
    }
    ARAP_CUDA_CHECK(cudaMemcpyAsync(d.costs, lm_costs_.data(), lm_costs_.size() * sizeof(float), cudaMemcpyHostToDevice, stream_));
    ARAP_CUDA_CHECK(cudaStreamSynchronize(stream_)); // lm_costs_ is reused by the next problem
    launches_ += lm_->launches() - l0;
}

void BatchPipeline::set_pcg_rtol(float rtol)
{
    pcg_rtol_ = rtol > 0.0f ? rtol : 0.0f;
    if (resident_) resident_->set_pcg_rtol(rtol);
}

void BatchPipeline::set_gn_rtol(float rtol)
{
    gn_rtol_ = rtol > 0.0f ? rtol : 0.0f;
    if (resident_) resident_->set_gn_rtol(rtol);
}

int BatchPipeline::run(const HostProblem* problems, int count)
{
    if (count <= 0) return 0;
    if (count > (int)dev_.size()) {
        fprintf(stderr, "arapb200: %d problems submitted, pipeline holds %d\n", count, (int)dev_.size());
        return 1;
    }
    const size_t cbytes = (size_t)nCont_ * (nGN_ + 1) * sizeof(float);
    const float wf = sqrtf(100.0f), wr = sqrtf(0.01f); // CombinedSolver.h:172-177
    const long long r0 = resident_ ? resident_->launches() : 0;
    // ---- stage + upload every problem, reset its state (resetGPU), build its strip tables ----
    for (int i = 0; i < count; ++i) {
        const HostProblem& hp = problems[i];
        Dev& d = dev_[i];
        if (hp.W <= 0 || hp.H <= 0 || (size_t)hp.W * hp.H > (size_t)maxW_ * maxH_) {
            fprintf(stderr, "arapb200: problem %dx%d does not fit the pipeline (%dx%d)\n", hp.W, hp.H, maxW_, maxH_);
            return 1;
        }
        const size_t N = (size_t)hp.W * hp.H;
        build_match_records(hp.W, hp.H, hp.mask_red, hp.matches, hp.n_matches, d.recs);
        if (d.recs.size() > d.matches_cap) {
            cudaFree(d.matches);
            d.matches_cap = d.recs.size() * 2 + 1024;
            ARAP_CUDA_OR_RETURN(cudaMalloc(&d.matches, d.matches_cap * sizeof(MatchRec)));
        }
        memcpy(d.h_in, hp.rgb, 3 * N);
        memcpy(d.h_in + 3 * N, hp.mask_red, N);
    }
    ARAP_CUDA_OR_RETURN(cudaEventRecord(ev_[0], stream_));
    for (int i = 0; i < count; ++i) {
        const HostProblem& hp = problems[i];
        Dev& d = dev_[i];
        const size_t N = (size_t)hp.W * hp.H;
        ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(d.rgb, d.h_in, 3 * N, cudaMemcpyHostToDevice, stream_));
        ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(d.mask, d.h_in + 3 * N, N, cudaMemcpyHostToDevice, stream_));
        if (!d.recs.empty())
            ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(d.matches, d.recs.data(), d.recs.size() * sizeof(MatchRec),
                                                cudaMemcpyHostToDevice, stream_));
        enqueue_reset_state(hp.W, hp.H, d.mask, d.X, d.U, d.A, d.M, stream_);
        launches_ += 1;
        d.resident = false;
        if (resident_ && !lm_on_) resident_->prepare_enqueue(i, hp.W, hp.H, d.M, stream_);
    }
    n_resident_ = 0;
    if (resident_ && !lm_on_) {
        ARAP_CUDA_OR_RETURN(cudaStreamSynchronize(stream_)); // strip counts are needed to size the launches
        for (int i = 0; i < count; ++i) {
            dev_[i].resident = resident_->prepare_finish(i);
            if (!dev_[i].resident && backend_ == ARAPB200_BACKEND_RESIDENT) {
                fprintf(stderr, "arapb200: problem %dx%d (%d strips) does not fit the resident back-end\n", problems[i].W,
                        problems[i].H, resident_->n_strips(i));
                return 3;
            }
            n_resident_ += dev_[i].resident ? 1 : 0;
        }
    }
    ARAP_CUDA_OR_RETURN(cudaEventRecord(ev_[1], stream_));
    // ---- solve.  Resident problems: whole continuation schedules, several problems per cooperative launch ----
    group_size_ = 0;
    for (int i = 0; i < count;) {
        Dev& d = dev_[i];
        if (lm_on_) {
            solve_lm(problems[i], d);
            ++i;
            continue;
        }
        if (!d.resident) {
            solve_streaming(problems[i], d);
            ++i;
            continue;
        }
        int run_len = 0; // consecutive resident problems
        while (i + run_len < count && dev_[i + run_len].resident) ++run_len;
        // small problems (each fits one thread-block cluster) share ONE launch, however many they are: their barriers run
        // through distributed shared memory and the problems need not be co-resident
        if (const int ncl = resident_->cluster_run(i, run_len)) {
            for (int j = 0; j < ncl; ++j) {
                Dev& dj = dev_[i + j];
                const HostProblem& hp = problems[i + j];
                enqueue_target_image(hp.W, hp.H, dj.matches, (int)dj.recs.size(), dj.C, stream_);
                launches_ += dj.recs.empty() ? 1 : 2;
                resident_->set_problem(i + j, dj.X, dj.A, dj.C, 1, wf, wr, dj.costs, nullptr);
            }
            resident_->enqueue_cluster(i, ncl, nCont_, nGN_, nPCG_, stream_);
            if (ncl > group_size_) group_size_ = ncl;
            i += ncl;
            continue;
        }
        // the cluster-eligible problems further down this run get their own launch: stop the co-resident group before them
        for (int j = 1; j < run_len; ++j)
            if (resident_->cluster_run(i + j, 1)) { run_len = j; break; }
        int gsz = resident_->group_size(i, run_len);
        if (gsz < run_len) { // several launches: split the run evenly instead of one full launch and a small remainder
            const int ngroups = (run_len + gsz - 1) / gsz;
            gsz = std::min(gsz, (run_len + ngroups - 1) / ngroups);
        }
        for (int j = 0; j < gsz; ++j) {
            Dev& dj = dev_[i + j];
            const HostProblem& hp = problems[i + j];
            enqueue_target_image(hp.W, hp.H, dj.matches, (int)dj.recs.size(), dj.C, stream_);
            launches_ += dj.recs.empty() ? 1 : 2;
            resident_->set_problem(i + j, dj.X, dj.A, dj.C, 1, wf, wr, dj.costs, nullptr);
        }
        resident_->enqueue_group(i, gsz, nCont_, nGN_, nPCG_, stream_);
        if (gsz > group_size_) group_size_ = gsz;
        i += gsz;
    }
    ARAP_CUDA_OR_RETURN(cudaEventRecord(ev_[2], stream_));
    // ---- flow extraction + forward warp + download ----
    for (int i = 0; i < count; ++i) {
        const HostProblem& hp = problems[i];
        Dev& d = dev_[i];
        enqueue_pos_to_flow(hp.W, hp.H, d.X, d.flow, stream_);
        enqueue_warp(hp.W, hp.H, d.X, d.rgb, d.mask, d.z, d.orgb, d.omask, stream_);
        launches_ += 1 + warp_launches_per_call();
    }
    ARAP_CUDA_OR_RETURN(cudaEventRecord(ev_[3], stream_));
    for (int i = 0; i < count; ++i) {
        const HostProblem& hp = problems[i];
        Dev& d = dev_[i];
        const size_t N = (size_t)hp.W * hp.H;
        unsigned char* o = d.h_out;
        ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(o, d.flow, N * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
        ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(o + 8 * N, d.orgb, 3 * N, cudaMemcpyDeviceToHost, stream_));
        ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(o + 11 * N, d.omask, N, cudaMemcpyDeviceToHost, stream_));
        ARAP_CUDA_OR_RETURN(cudaMemcpyAsync(o + 12 * N, d.costs, cbytes, cudaMemcpyDeviceToHost, stream_));
    }
    ARAP_CUDA_OR_RETURN(cudaStreamSynchronize(stream_));
    if (resident_ && n_resident_ > 0) {
        if (int st = resident_->status(stream_)) {
            fprintf(stderr, "arapb200: resident solver aborted (watchdog, code %d)\n", st);
            return 4;
        }
    }
    if (solver_ && n_resident_ < count) {
        unsigned bad = 0;
        solver_->read_back(stream_, nullptr, &bad);
        if (bad) {
            fprintf(stderr, "arapb200: internal error: UrShape check failed\n");
            return 2;
        }
    }
    for (int i = 0; i < count; ++i) {
        const HostProblem& hp = problems[i];
        Dev& d = dev_[i];
        const size_t N = (size_t)hp.W * hp.H;
        const unsigned char* o = d.h_out;
        if (hp.out_flow) memcpy(hp.out_flow, o, N * sizeof(float2));
        if (hp.out_rgb) memcpy(hp.out_rgb, o + 8 * N, 3 * N);
        if (hp.out_mask) memcpy(hp.out_mask, o + 11 * N, N);
        if (hp.out_costs) memcpy(hp.out_costs, o + 12 * N, cbytes);
    }
    if (resident_) launches_ += resident_->launches() - r0;
    cudaEventElapsedTime(&ms_total_, ev_[0], ev_[3]);
    cudaEventElapsedTime(&ms_solve_, ev_[1], ev_[2]);
    cudaEventElapsedTime(&ms_warp_, ev_[2], ev_[3]);
    return 0;
}

} // namespace arapb200
