// common.cuh -- shared device/host helpers: error handling, the arithmetic contract's scalar
// routines (sincos, guarded invert) and the exact-sum primitives.
//
// The whole library is compiled with -fmad=false -prec-div=true -prec-sqrt=true: fused
// multiply-adds exist only where fmaf()/fma() is written, divisions and square roots are IEEE.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

namespace arapb200 {

// Failures inside the library surface as a C++ exception that every extern "C" entry point catches and turns into
// its error convention (a non-zero return code for arapb200_*, NULL from Opt_ProblemPlan, "finished" from Opt_ProblemStep
// plus a sticky error on the plan).  The reference prints and exits the process (ARAP/API/src/solverGPUGaussNewton.t:
// 59-73, ARAP/shared/cudaUtil.h:26-31); a library embedded in somebody else's process must not.
struct ArapError {
    int code;
    ArapError(int c) : code(c ? c : 1) {}
};
[[noreturn]] inline void arap_fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline void arap_fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "arapb200: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    throw ArapError(code);
}
#define ARAP_CUDA_CHECK(call)                                                                          \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            ::arapb200::arap_fail((int)e__, "CUDA error %d (%s) at %s:%d: %s", (int)e__, cudaGetErrorString(e__), \
                                  __FILE__, __LINE__, #call);                                          \
    } while (0)

// Same, but returns the error code (flat arapb200_* API).
#define ARAP_CUDA_OR_RETURN(call)                                                                      \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            fprintf(stderr, "arapb200: CUDA error %d (%s) at %s:%d: %s\n", (int)e__,                   \
                    cudaGetErrorString(e__), __FILE__, __LINE__, #call);                               \
            return (int)e__;                                                                           \
        }                                                                                              \
    } while (0)

// ----------------------------------------------------------------------------------------------
// Contract C2: sincos in binary64 (Cody-Waite reduction + fdlibm kernel polynomials), fma/mul/add
// only, rounded to binary32.  Mirrors oracle/arap_oracle.c:arap_oracle_sincos bit for bit.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void contract_sincos(float a, float& s_out, float& c_out)
{
    const double PIO2_HI = 1.57079632673412561417e+00;
    const double PIO2_LO = 6.07710050650619224932e-11;
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    double x = (double)a;
    double k = rint(x * TWO_OVER_PI);
    double r = fma(-k, PIO2_HI, x);
    r = fma(-k, PIO2_LO, r);
    double z = r * r;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    double sr = fma(r * z, ps, r);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    long long q = (long long)k;
    double s, c;
    switch ((int)(q & 3)) {
    case 0: s = sr; c = cr; break;
    case 1: s = cr; c = -sr; break;
    case 2: s = -sr; c = -cr; break;
    default: s = -cr; c = sr; break;
    }
    s_out = (float)s;
    c_out = (float)c;
}

// solverGPUGaussNewton.t:323-332 (CERES flavour): 1 / (1 + sqrt(d))^2
__device__ __forceinline__ float guarded_invert(float d)
{
    float t = 1.0f + sqrtf(d);
    return 1.0f / (t * t);
}

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2)
{
    return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}

// ----------------------------------------------------------------------------------------------
// Contract C3: exact sums.  A value is carried as (h, l): h = exact sum of binned hi parts,
// l = plain binary64 sum of the (tiny) remainders.  Any combination order gives the same
// float(h + l) (see DESIGN.md for the probability bound).
// ----------------------------------------------------------------------------------------------
struct HL {
    double h, l;
};

__device__ __forceinline__ int ilogb_f32(float m) // m >= 0; subnormals/zero report -127
{
    return ((__float_as_int(m) >> 23) & 0xff) - 127;
}
__device__ __forceinline__ int ilogb_f64(double m) // m >= 0
{
    return ((__double2hiint(m) >> 20) & 0x7ff) - 1023;
}
// B = 1.5 * 2^k, k clamped to the normal range
__device__ __forceinline__ double bin_base(int k)
{
    k = max(-1000, min(1000, k));
    return __hiloint2double(((k + 1023) << 20) | 0x80000, 0);
}
// split v against base B: hi is a multiple of ulp(B), lo = v - hi exactly
__device__ __forceinline__ void bin_split(double B, double v, double& hi, double& lo)
{
    double t = __dadd_rn(B, v);
    hi = __dadd_rn(t, -B);
    lo = __dadd_rn(v, -hi);
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Exact sum over a warp of one binary32 term per lane (all 32 lanes must call).
// n_scale = number of terms the bin must be able to hold (>= 32); result valid in every lane.
__device__ __forceinline__ HL warp_exact_sum(float g)
{
    float m = warp_max(fabsf(g));
    HL out;
    int k = ilogb_f32(m) + 5 + 2; // 32 terms
    double B = bin_base(k);
    double hi, lo;
    bin_split(B, (double)g, hi, lo);
    out.h = warp_sum(hi);
    out.l = warp_sum(lo);
    return out;
}

// Combine n (h, l) pairs held one per lane (lanes >= n pass zeros) exactly; result in every lane.
__device__ __forceinline__ HL warp_combine(HL v)
{
    double m = warp_max(fabs(v.h));
    int k = ilogb_f64(m) + 5 + 2;
    double B = bin_base(k);
    double hi, lo;
    bin_split(B, v.h, hi, lo);
    HL out;
    out.h = warp_sum(hi);
    out.l = warp_sum(__dadd_rn(lo, v.l));
    return out;
}

// Block-level exact sum of one binary32 term per thread.  smem must hold 2*32 doubles.
// Requires blockDim.x to be a multiple of 32 and <= 1024.  Result valid in warp 0 (all lanes).
__device__ __forceinline__ HL block_exact_sum(float g, double* smem)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    HL w = warp_exact_sum(g);
    if (lane == 0) {
        smem[wid] = w.h;
        smem[32 + wid] = w.l;
    }
    __syncthreads();
    HL out = {0.0, 0.0};
    if (wid == 0) {
        HL v;
        v.h = (lane < nw) ? smem[lane] : 0.0;
        v.l = (lane < nw) ? smem[32 + lane] : 0.0;
        out = warp_combine(v);
    }
    return out;
}

// ----------------------------------------------------------------------------------------------
// Wide fixed-point accumulators (streaming back-end): a grid-wide exact sum WITHOUT a fence, a counter or a
// "last block".  Every value this library sums is a multiple of 2^-149 of magnitude < 2^150, so a 320-bit
// fixed-point number with LSB 2^-160 holds any such sum exactly.  It is kept as WA_LIMBS 32-bit limbs, each in its
// own signed 64-bit word (2^31 contributions cannot overflow a word), WA_COPIES copies to spread same-address
// atomics.  A producer adds the (at most three) non-zero limbs of a binary64 value with fire-and-forget
// red.global.add.u64; integer addition commutes, so the total is independent of arrival order.  A consumer in a
// LATER kernel sums the copies, propagates carries and rounds ONCE to binary32 (round-to-nearest-even).
// ----------------------------------------------------------------------------------------------
constexpr int WA_LIMBS = 10;
constexpr int WA_COPIES = 8;
constexpr int WA_WORDS = WA_LIMBS * WA_COPIES; // one accumulator, [copy][limb]
constexpr int WA_LSB = -160;

__device__ __forceinline__ void wide_add(unsigned long long* __restrict__ acc, unsigned copy, double v)
{
    if (v == 0.0) return;
    const long long bits = __double_as_longlong(v);
    const bool neg = bits < 0;
    const int ex = (int)((bits >> 52) & 0x7ff);
    unsigned long long mant = ((unsigned long long)bits & 0xfffffffffffffULL) | (ex ? (1ULL << 52) : 0ULL);
    int pos = (ex ? ex : 1) - 1075 - WA_LSB; // bit position of the mantissa's LSB
    if (pos < 0) {                           // multiples of 2^-149 have the trailing zeros to give
        if (pos < -52) return;
        mant >>= -pos;
        pos = 0;
    }
    const int idx = pos >> 5, off = pos & 31;
    const unsigned __int128 big = (unsigned __int128)mant << off;
    unsigned long long* a = acc + (copy % WA_COPIES) * WA_LIMBS;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned long long w = (unsigned)(big >> (32 * k));
        if (w != 0 && idx + k < WA_LIMBS) {
            const unsigned long long c = neg ? (0ULL - w) : w;
            asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(a + idx + k), "l"(c) : "memory");
        }
    }
}

// Read one accumulator.  Called by the 16 lanes of a half-warp (seg_lane = lane & 15): lanes 0..9 fetch one limb
// each; the value is returned in every lane of the half-warp.  Both halves of a warp may decode different
// accumulators in the same call (all 32 lanes must call).
__device__ __forceinline__ long long wide_fetch(const unsigned long long* __restrict__ acc, int seg_lane)
{
    long long s = 0;
    if (seg_lane < WA_LIMBS) {
        unsigned long long v[WA_COPIES];
#pragma unroll
        for (int c = 0; c < WA_COPIES; ++c) v[c] = __ldcg(acc + c * WA_LIMBS + seg_lane);
#pragma unroll
        for (int c = 0; c < WA_COPIES; ++c) s += (long long)v[c];
    }
    return s;
}
__device__ __forceinline__ float wide_round(long long s_mine)
{
    unsigned limb[WA_LIMBS];
    long long carry = 0;
#pragma unroll
    for (int k = 0; k < WA_LIMBS; ++k) {
        const long long t = __shfl_sync(0xffffffffu, s_mine, k, 16) + carry;
        limb[k] = (unsigned)t;
        carry = t >> 32;
    }
    const bool neg = carry < 0;
    if (neg) { // two's complement negate over the 320 bits
        unsigned c = 1;
#pragma unroll
        for (int k = 0; k < WA_LIMBS; ++k) {
            const unsigned v = ~limb[k];
            limb[k] = v + c;
            c = (c && v == 0xffffffffu) ? 1u : 0u;
        }
    }
    int j = -1;
    unsigned hi = 0, lo = 0;
    bool sticky = false;
#pragma unroll
    for (int k = WA_LIMBS - 1; k >= 0; --k) {
        if (j < 0) {
            if (limb[k] != 0) {
                j = k;
                hi = limb[k];
                lo = (k > 0) ? limb[k > 0 ? k - 1 : 0] : 0u;
            }
        } else if (k < j - 1) {
            sticky = sticky || (limb[k] != 0);
        }
    }
    if (j < 0) return 0.0f;
    // >= 33 significant bits; the sticky bit is jammed far below binary32's rounding position (round to odd)
    const unsigned long long top = ((unsigned long long)hi << 32) | lo | (sticky ? 1ULL : 0ULL);
    const float f = __ull2float_rn(top);
    const int e = 32 * (j - 1) + WA_LSB; // weight of top's LSB, in [-192, 96]
    const double d = (double)f * __hiloint2double((e + 1023) << 20, 0);
    const float r = (float)d; // exact for normal results; subnormal totals are exact multiples of 2^-149
    return neg ? -r : r;
}

__device__ __forceinline__ float hl_to_float(HL v)
{
    return (float)__dadd_rn(v.h, v.l);
}

} // namespace arapb200
