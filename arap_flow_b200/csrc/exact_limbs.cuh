// exact_limbs.cuh -- integer helpers of the resident kernel's exact grid-wide sums (contract C3):
// a binary32 term <-> four signed 24-bit limbs of a 96-bit fixed-point number, and the 4-limb total -> binary32 with one
// rounding.  Plain integer code, host + device, so that tests/test_exact_limbs.py can check it against exact rational
// arithmetic without a GPU (tests/exact_limbs_host.cpp).
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define ARAP_HD __host__ __device__ __forceinline__
#else
#define ARAP_HD inline
#endif

// On the B200 the integer limb conversion (to_limbs) and the 128-bit decode (limbs_to_float_int) are SLOWER than the
// float / double chains they were meant to replace (64- and 128-bit shifts by variable amounts are long 32-bit
// instruction sequences): C1 7.0 vs 7.8 pairs/s, profiles/r2_barrier_ab.txt.  They stay here, with their host-side
// exactness tests, as measured alternatives; the kernel can select a decode with ARAP_RS_INT_FOLD (0 = binary64 chain,
// 1 = 128-bit, 2 = the 64-bit fast path below).
#ifndef ARAP_RS_INT_FOLD
#define ARAP_RS_INT_FOLD 0
#endif

namespace arapb200 {

ARAP_HD int clz64(unsigned long long v) // v != 0
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

// One binary32 term -> four signed 24-bit limbs (units 2^(S-18), 2^(S-42), 2^(S-66), 2^(S-90)), in integer arithmetic
// (the float formulation is a chain of 4 x {FMUL, F2I, I2F, FSUB} on the slow conversion pipe, in front of every
// barrier arrival).  Exact for |g| < 2^(S+6) and ulp(g) >= 2^(S-90); bits below the LSB are rounded to nearest;
// |g| >= 2^(S+6), Inf and NaN report ovf and contribute nothing.
ARAP_HD void to_limbs(float g, int S, int& l0, int& l1, int& l2, int& l3, bool& ovf)
{
    unsigned b;
    memcpy(&b, &g, 4);
    const int e = (int)((b >> 23) & 0xffu);
    unsigned m = (b & 0x7fffffu) | (e ? 0x800000u : 0u);
    int p = (e > 1 ? e : 1) - 60 - S; // bit position of m's LSB: value = m * 2^(max(e,1) - 150), unit 2^(S - 90)
    ovf = (e == 255) || (m != 0u && p > 72);
    if (p < 0) { // below the LSB: round to nearest
        const int r = -p;
        m = (r <= 25) ? ((m + (1u << (r - 1))) >> r) : 0u;
        p = 0;
    }
    if (ovf) m = 0u;
    if (p > 72) p = 72;
    const int j = (p * 43) >> 10; // p / 24 for 0 <= p < 96
    const int o = p - 24 * j;
    const unsigned long long w = (unsigned long long)m << o; // < 2^48
    int lo = (int)(w & 0xffffffull), hi = (int)(w >> 24);
    if ((int)b < 0) { lo = -lo; hi = -hi; }
    l3 = (j == 0) ? lo : 0;
    l2 = (j == 0) ? hi : ((j == 1) ? lo : 0);
    l1 = (j == 1) ? hi : ((j == 2) ? lo : 0);
    l0 = (j == 2) ? hi : ((j == 3) ? lo : 0);
}

// T = L0*2^72 + L1*2^48 + L2*2^24 + L3 (|Lj| < 2^45 => |T| < 2^118, units 2^e_unit) -> binary32, rounded ONCE to
// nearest-even.  Returns false (nothing computed) when the result would leave binary32's normal range by a wide margin:
// the caller then takes the binary64 route.
ARAP_HD bool limbs_to_float_int(const long long L[4], int e_unit, bool& is_zero, float& out)
{
    const __int128 T = ((__int128)L[0] << 72) + ((__int128)L[1] << 48) + ((__int128)L[2] << 24) + (__int128)L[3];
    is_zero = (T == 0);
    out = 0.0f;
    if (is_zero) return true;
    const bool neg = T < 0;
    const unsigned __int128 a = neg ? (unsigned __int128)(-T) : (unsigned __int128)T;
    const unsigned long long hi = (unsigned long long)(a >> 64), lo = (unsigned long long)a;
    const int nb = hi ? 128 - clz64(hi) : 64 - clz64(lo); // significant bits of |T|
    if (nb + e_unit < -120 || nb + e_unit > 120) return false;
    unsigned q;
    int sh = nb - 24;
    if (sh <= 0) {
        q = (unsigned)lo;
        sh = 0;
    } else {
        const unsigned t25 = (unsigned)(a >> (sh - 1)); // q and the round bit
        const bool sticky = (a & ((((unsigned __int128)1) << (sh - 1)) - 1)) != 0;
        q = t25 >> 1;
        if ((t25 & 1u) && (sticky || (q & 1u))) ++q; // nearest, ties to even (q may become 2^24: still exact below)
    }
    // q * 2^(sh + e_unit): q <= 2^24 and the exponent keeps the product a normal binary32 => exact
    const int ex = sh + e_unit; // in [-144, 120]
    unsigned long long dbits = (unsigned long long)(ex + 1023) << 52;
    double scale;
    memcpy(&scale, &dbits, 8);
    const float r = (float)((double)q * scale);
    out = neg ? -r : r;
    return true;
}

// A lighter integer decode (ARAP_RS_INT_FOLD == 2).  The scale prediction keeps the total's leading bit inside the top two
// limbs, so: fold the carries of the low limbs upward with constant shifts, H = L0*2^24 + L1 + carries (64-bit), low part
// 0 <= Lo < 2^48 as a sticky bit only, then ONE variable 64-bit shift.  Everything 64-bit (the 128-bit version above needs
// 128-bit variable shifts: measured slower than the binary64 chain).  Returns false -- caller falls back to the binary64
// route -- when the fast path's preconditions do not hold (|L0| >= 2^37, fewer than 25 significant bits in H, or a result
// outside the comfortable binary32 range).
ARAP_HD bool limbs_to_float_i64(const long long L[4], int e_unit, bool& is_zero, float& out)
{
    out = 0.0f;
    is_zero = false;
    if (L[0] >= (1ll << 37) || L[0] <= -(1ll << 37)) return false;
    // low part: Lo = l2 * 2^24 + L3 with l2 = L2 mod 2^24 (floor), carries go up
    const long long c2 = L[2] >> 24;                 // floor
    const long long l2 = L[2] - (c2 << 24);          // [0, 2^24)
    long long Lo = (l2 << 24) + L[3];                // |L3| < 2^45  =>  -2^45 < Lo < 2^48 + 2^45
    const long long c = Lo >> 48;                    // floor: -1, 0 or 1
    Lo -= c << 48;                                   // [0, 2^48)
    const long long H = (L[0] << 24) + L[1] + c2 + c; // |.| < 2^61 + 2^45 + 2^21 + 1
    // total T = H * 2^48 + Lo, 0 <= Lo < 2^48
    const bool neg = H < 0;
    const bool low = Lo != 0;
    // |T| = a * 2^48 + (low ? something in (0, 2^48) : 0)
    const unsigned long long a = neg ? (unsigned long long)(-H) - (low ? 1ull : 0ull) : (unsigned long long)H;
    if (a == 0) {
        if (!low) { is_zero = true; return true; }
        return false; // the whole total sits in the low limbs: leave it to the general route
    }
    const int nb = 64 - clz64(a);
    if (nb < 25) return false;
    if (nb + 48 + e_unit < -120 || nb + 48 + e_unit > 120) return false;
    const int sh = nb - 24; // >= 1
    unsigned q = (unsigned)(a >> sh);
    const bool rbit = ((a >> (sh - 1)) & 1ull) != 0;
    const bool sticky = low || (a & ((1ull << (sh - 1)) - 1ull)) != 0;
    if (rbit && (sticky || (q & 1u))) ++q;
    const int ex = sh + 48 + e_unit;
    unsigned long long dbits = (unsigned long long)(ex + 1023) << 52;
    double scale;
    memcpy(&scale, &dbits, 8);
    const float r = (float)((double)q * scale);
    out = neg ? -r : r;
    return true;
}

} // namespace arapb200
