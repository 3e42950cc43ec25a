// Host build of arap_flow_b200/csrc/exact_limbs.cuh for tests/test_exact_limbs.py (g++, no CUDA).
#include "../arap_flow_b200/csrc/exact_limbs.cuh"

extern "C" {
int el_to_limbs(float g, int S, int* out4)
{
    bool ovf;
    arapb200::to_limbs(g, S, out4[0], out4[1], out4[2], out4[3], ovf);
    return ovf ? 1 : 0;
}
// returns 0 = computed, 1 = out of the integer path's range (caller falls back), 2 = computed and zero
int el_fold(const long long* L4, int e_unit, float* out)
{
    bool z;
    if (!arapb200::limbs_to_float_int(L4, e_unit, z, *out)) return 1;
    return z ? 2 : 0;
}
// the lighter 64-bit decode: 0 = computed, 1 = preconditions not met (fallback), 2 = computed and zero
int el_fold64(const long long* L4, int e_unit, float* out)
{
    bool z;
    if (!arapb200::limbs_to_float_i64(L4, e_unit, z, *out)) return 1;
    return z ? 2 : 0;
}
}
