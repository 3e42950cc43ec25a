// solver_resident.cuh -- the resident Gauss-Newton/PCG back-end: ONE persistent cooperative kernel runs
// whole Gauss-Newton solves (optionally whole 19-step continuation schedules) with the PCG state
// (r, delta, p, cos/sin, flags) held in registers and the search direction shared through shared-memory
// tiles.  Nothing but tile-boundary values and two 16-byte partial sums per CTA cross L2 per PCG iteration.
// See DESIGN.md section 4 for the decomposition, the barrier and the halo protocol.
#pragma once
#include "common.cuh"

namespace arapb200 {

constexpr int RS_STRIP_W = 32;   // a strip is 32 x 8 pixels: one warp, lane = column, 8 rows per lane
constexpr int RS_STRIP_H = 8;
constexpr int RS_MAX_WARPS = 16; // strips per CTA
constexpr int RS_MAX_CTAS = 160; // CTAs per problem (8-bit arrival count per barrier word)
constexpr int RS_OUTBOX_ENTRIES = 80; // 32 top + 32 bottom + 8 left + 8 right, 48 bytes each

// One problem as the kernel sees it (device memory, one per blockIdx.y)
struct ResProb {
    int W, H, SX, SY;
    int n_strips;          // active strips
    int G;                 // CTAs working on this problem (blockIdx.x >= G exits)
    float2* X;             // Offset (in/out)
    float* A;              // Angle (in/out)
    const float2* C;       // Constraints image, or (lerp_mode) per-pixel match target, (<0) = none
    const float* M;        // Mask (0 = active)
    int lerp_mode;         // 1: C holds targets; constraint = (1-a)*pixel + a*target, a = (t+1)/nCont
    float wf, wr, wf2, wr2;
    const int2* strip_xy;      // [n_strips] strip coordinates (sx, sy), column-major order
    const int* slot_of_strip;  // [SY*SX] -> slot or -1
    uint4* outbox;             // [n_strips][RS_OUTBOX_ENTRIES][3]: six (float, tag) words per boundary pixel
    unsigned long long* bar;   // [2][4] barrier words: 16-bit arrival count | 48-bit fixed-point limb sum
    float* costs;              // [nCont][nGN+1]
    float* trace;              // optional [nCont*nGN][nPCG][3]
    int* status;               // [0] abort flag, [1] error code
    int nCont, nGN, nPCG;
    unsigned long long* prof;  // optional [G][8] cycle counters (debug)
};

class ResidentSolver {
public:
    // capacity: largest image handled
    ResidentSolver(int maxW, int maxH);
    ~ResidentSolver();
    ResidentSolver(const ResidentSolver&) = delete;
    ResidentSolver& operator=(const ResidentSolver&) = delete;

    // Build the strip tables for mask M (device, float[N], 0 = active).  Blocking (4-byte read-back).
    // Returns false when the problem does not fit on chip (caller falls back to the streaming back-end).
    bool prepare(int W, int H, const float* d_M, cudaStream_t stream);
    // Enqueue a launch: nCont continuation steps x nGN Gauss-Newton steps x nPCG iterations; costs
    // (device, nCont*(nGN+1) floats) receives the cost before the first and after every GN step.
    void enqueue(float2* X, float* A, const float2* C, int lerp_mode, float wf, float wr, int nCont, int nGN,
                 int nPCG, float* d_costs, float* d_trace, cudaStream_t stream);
    // after the stream has been synchronised: non-zero = the kernel bailed out (watchdog)
    int status(cudaStream_t stream);
    long long launches() const { return launches_; }
    int n_strips() const { return n_strips_; }
    int ctas() const { return G_; }
    int warps() const { return NW_; }
    // debug: per-CTA cycle accounting of the next launches into d_prof ([ctas()][8] u64), or null to disable
    void set_profile(unsigned long long* d_prof) { d_prof_ = d_prof; }

private:
    int maxW_, maxH_;
    int W_ = 0, H_ = 0, SX_ = 0, SY_ = 0, n_strips_ = 0, G_ = 0, NW_ = 0;
    const float* d_M_ = nullptr;
    int2* d_strip_xy_ = nullptr;
    int* d_slot_of_strip_ = nullptr;
    int* d_count_ = nullptr;
    uint4* d_outbox_ = nullptr;
    size_t outbox_cap_ = 0;
    unsigned long long* d_bar_ = nullptr;
    int* d_status_ = nullptr;
    ResProb* d_prob_ = nullptr;
    int sm_count_ = 0;
    long long launches_ = 0;
    unsigned long long* d_prof_ = nullptr;
};

} // namespace arapb200
