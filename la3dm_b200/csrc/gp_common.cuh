// la3dm_b200 -- GPOctoMap arithmetic shared by the SIMT kernels (predict_gp.cu) and the tensor-core kernel
// (predict_gp_tc.cu): libm-exact expf and the Matern-3/2 kernel element (include/gpoctomap/gpregressor.h:114-117).
#pragma once
#include "engine.cuh"

namespace la3dm_b200 {
namespace {

// expf as the reference's host libm computes it: glibc >= 2.27 (the algorithm of ARM's optimized-routines expf,
// sysdeps/ieee754/flt-32/e_expf.c): x N / ln2 = k + r, exp(x) = 2^(k/N) 2^(r/N) ~ T[k % N] 2^(k/N int part)
// (C0 r^3 + C1 r^2 + C2 r + 1), all in double, rounded once to float.  Restated here because the GP path amplifies a
// 1-ulp difference in K by cond(K) ~ 1e4; checked bit for bit against libm on 2e7 arguments.
__device__ const unsigned long long kExp2fTab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

__device__ __forceinline__ float expf_libm(float x) {
    if (!(x > -87.0f && x < 88.0f)) return (float) exp((double) x);     // outside the fast path of the algorithm
    const double InvLn2N = 0x1.71547652b82fep+0 * 32, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32,
                 C2 = 0x1.62e42ff0c52d6p-1 / 32;
    double z = InvLn2N * (double) x;
    double kd = z + SHIFT;
    const unsigned long long ki = (unsigned long long) __double_as_longlong(kd);
    kd -= SHIFT;
    const double r = z - kd;
    const unsigned long long t = kExp2fTab[ki % 32] + (ki << (52 - 5));
    const double s = __longlong_as_double((long long) t);
    z = C0 * r + C1;
    const double r2 = r * r;
    double y = C2 * r + 1;
    y = z * r2 + y;
    y = y * s;
    return (float) y;
}

// covMaterniso3 element (gpregressor.h:114-117) on pre-scaled coordinates
__device__ __forceinline__ float matern3(float ax, float ay, float az, float bx, float by, float bz, float sf2) {
    const float dx = bx - ax, dy = by - ay, dz = bz - az;
    const float r = sqrtf(dx * dx + (dy * dy + dz * dz));      // Eigen rowwise().norm() of a 3-vector
    const float e = expf_libm(-r);
    return ((1 + r) * e) * sf2;
}


}  // namespace
}  // namespace la3dm_b200
