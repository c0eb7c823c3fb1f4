// TEST INFRASTRUCTURE ONLY -- scalar/SIMD-friendly helpers shared by the stand-in inference headers.
//
// Eigen is a third-party dependency of la3dm that is NOT under /root/reference and not installed here (version
// unpinned -> "parity unpinned").  The stand-ins restate the Eigen EXPRESSIONS written in the
// reference's inference headers as plain fp32 loops.  Two flavours, chosen at compile time:
//   default                : libm cosf/sinf/expf per element, sequential fp32 sums.  Deterministic; this flavour
//                            generates the golden vectors (built with -ffp-contract=off, like a stock x86-64 build
//                            of the reference, which has no FMA contraction).
//   -DLA3DM_STANDIN_FAST   : branch-free polynomial sincos / exp in the style of the Cephes-derived packet math Eigen
//                            3.3 uses (psin/pcos/pexp), written so that gcc -O3 -march=native auto-vectorises the
//                            dense M x N kernel-matrix loops.  This flavour is the CPU *timing* baseline so that the
//                            reference is not sandbagged by scalar libm calls.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace la3dm_standin {

#ifdef LA3DM_STANDIN_FAST
// Cephes-style single-precision sin/cos, valid for 0 <= x < ~8192 (we only ever pass 2*pi*d, d in [0, ~50)).
static inline void sincos_f(float x, float &s, float &c) {
    const float FOPI = 1.27323954473516f;  // 4/pi
    const float DP1 = 0.78515625f, DP2 = 2.4187564849853515625e-4f, DP3 = 3.77489497744594108e-8f;
    int j = (int) (x * FOPI);
    j = (j + 1) & ~1;                        // j += j & 1
    const float y = (float) j;
    const float z0 = ((x - y * DP1) - y * DP2) - y * DP3;
    const float z = z0 * z0;
    const float ps = ((-1.9515295891E-4f * z + 8.3321608736E-3f) * z - 1.6666654611E-1f) * z * z0 + z0;
    const float pc = ((2.443315711809948E-005f * z - 1.388731625493765E-003f) * z + 4.166664568298827E-002f) * z * z
                     - 0.5f * z + 1.0f;
    const int q = j & 7;                     // octant pair
    const bool swap = (q == 2) || (q == 6);
    float ss = swap ? pc : ps;
    float cc = swap ? ps : pc;
    ss = (q & 4) ? -ss : ss;                 // sin sign: octants 4..7
    cc = (q == 2 || q == 4) ? -cc : cc;      // cos sign: octants 2..5 -> j in {2,4}
    s = ss; c = cc;
}
static inline float exp_f(float x) {
    // Cephes expf: x = n*ln2 + r, polynomial in r, scale by 2^n
    x = x < -87.0f ? -87.0f : (x > 88.0f ? 88.0f : x);
    const float LOG2EF = 1.44269504088896341f, C1 = 0.693359375f, C2 = -2.12194440e-4f;
    float fx = std::floor(x * LOG2EF + 0.5f);
    const float r = (x - fx * C1) - fx * C2;
    const float r2 = r * r;
    float p = 1.9875691500E-4f;
    p = p * r + 1.3981999507E-3f; p = p * r + 8.3334519073E-3f; p = p * r + 4.1665795894E-2f;
    p = p * r + 1.6666665459E-1f; p = p * r + 5.0000001201E-1f;
    p = p * r2 + r + 1.0f;
    std::int32_t e = ((std::int32_t) fx + 127) << 23;
    float sc; std::memcpy(&sc, &e, 4);
    return p * sc;
}
#else
static inline void sincos_f(float x, float &s, float &c) { s = sinf(x); c = cosf(x); }
static inline float exp_f(float x) { return expf(x); }
#endif

// covSparse element (bgkinference.h:115-116 and the -L/-LV copies): d is the (already ell-scaled) distance.
//   (((2 + cos(d*2*pi~)) * (1 - d) / 3) + sin(d*2*pi~) / (2*pi~)) * sf2,   pi~ = 3.1415926f
static inline float sparse_kernel_unclamped(float d, float sf2) {
    const float t = d * 2.0f * 3.1415926f;
    float s, c;
    sincos_f(t, s, c);
    return (((2.0f + c) * (1.0f - d) / 3.0f) + s / (2.0f * 3.1415926f)) * sf2;
}

// Eigen's rowwise().norm() on a 3-vector: sqrt(a0 + (a1 + a2)) (redux_novec_unroller halves [0] | [1,2]).
static inline float norm3(float dx, float dy, float dz) {
    return std::sqrt(dx * dx + (dy * dy + dz * dz));
}

}  // namespace la3dm_standin
