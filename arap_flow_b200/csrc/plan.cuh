// plan.cuh -- the object behind Opt_Plan: solver parameters, iteration state and the back-end.
// Mirrors the plan vtable {init, step, cost, setsolverparameter, free} of ARAP/API/src/o.t:126-133 as
// filled by ARAP/API/src/solverGPUGaussNewton.t:1254-1284.
#pragma once
#include "../../include/arapb200.h"
#include "solver_stream.cuh"

namespace arapb200 {

class GnPlan {
public:
    GnPlan(int W, int H, int verbosity, int backend);
    ~GnPlan();
    // solverGPUGaussNewton.t:1205-1221.  false = unknown name.
    bool set_parameter(const char* name, const void* value);
    void init(void** problemparams);  // :956-1007
    int step(void** problemparams);   // :1016-1177
    void solve(void** problemparams); // o.t:2548-2551, with a single host sync at the end
    double current_cost() const { return (double)prev_cost_; }
    long long launches() const { return stream_.launches(); }
    // parity/debug: device buffer of 3*lIterations floats per GN step, or null
    void set_trace(float* d_trace) { d_trace_ = d_trace; }

private:
    void bind(void** problemparams);
    void check_grid(unsigned bad_u) const;
    int W_, H_, verbosity_, backend_;
    int n_iterations_ = 10, l_iterations_ = 10; // solver_parameter_defaults, :26-39
    int n_iter_ = 0;
    float prev_cost_ = 0.f;
    cudaStream_t stream_h_ = nullptr;
    StreamSolver stream_;
    float* d_trace_ = nullptr;
};

} // namespace arapb200
