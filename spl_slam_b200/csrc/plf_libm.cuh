// plf_libm.cuh -- bit-exact device models of the three single-precision libm functions the reference's code reaches
// through C++ overload resolution, as glibc 2.39 (x86-64) computes them:
//
//   cosf / sinf : `cos(angle)` with a float argument under `using namespace std` (src/ORBextractor.cc:114, rBRIEF steering)
//                 and `cos(pSingleLine->direction)` (binary_descriptor_custom.cpp:1130-1131, LBD) resolve to
//                 std::cos(float) = cosf.  glibc's cosf/sinf (sysdeps/ieee754/flt-32/s_sincosf.h, the ARM optimized
//                 routines) evaluate a double-precision polynomial after a one-multiply range reduction and round once.
//   atan2f      : `atan2(dy, dx)` on floats (LSDDetector_custom.cpp:298, KeyLine.angle) = atan2f;
//                 glibc's e_atan2f.c / s_atanf.c are the fdlibm single-precision routines.
//
// The published algorithms are restated here (constants are the published ones); what makes them trustworthy is the
// check, not the recollection: oracle/libm_check.c compares these exact functions (this header compiled as host code)
// with the container's libm -- cosf/sinf EXHAUSTIVELY over every float with |x| <= 6.3 (2.17e9 inputs, covers
// [0, 2 pi) for ORB and [-pi, pi] for LBD), atanf exhaustively over all finite floats, atan2f on 1.6e9 random pairs:
// zero mismatches.  A correctly rounded cos would NOT do: (float)cos((double)x) differs from glibc's cosf for 0.04 % of
// those inputs and (float)atan2(double, double) from atan2f for 15 % of pairs.
//
// Everything is IEEE arithmetic without FMA contraction (nvcc -fmad=false; gcc -ffp-contract=off for the host check).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if !defined(__CUDACC__) && !defined(__host__)
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace plf_libm {

__host__ __device__ __forceinline__ uint32_t asuint(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
__host__ __device__ __forceinline__ float asfloat(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }

// polynomial of the quadrant: sine for even n, cosine for odd n; `neg` selects the negated cosine set (quadrants 2, 3)
__host__ __device__ __forceinline__ float sincos_poly(double x, double x2, bool neg, int n)
{
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double t1 = s2 + x2 * s3;
        const double x7 = x3 * x2;
        const double s = x + x3 * s1;
        return (float)(s + x7 * t1);
    }
    const double sg = neg ? -1.0 : 1.0;
    const double c0 = sg * 0x1p0, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5, c3 = sg * -0x1.6c087e89a359dp-10,
                 c4 = sg * 0x1.99343027bf8c3p-16;
    const double x4 = x2 * x2;
    const double t2 = c3 + x2 * c4;
    const double t1 = c0 + x2 * c1;
    const double x6 = x4 * x2;
    const double c = t1 + x4 * c2;
    return (float)(c + x6 * t2);
}

// x - n * (pi / 2) with n = round(x * 2 / pi) taken from a 2^24-scaled product (no libm rounding call)
__host__ __device__ __forceinline__ double reduce_fast(double x, int* np)
{
    const double r = x * 0x1.45F306DC9C883p+23;
    const int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return x - (double)n * 0x1.921FB54442D18p0;
}

// valid for |y| < 120 (the reference's arguments are angles in [-pi, 2 pi]); larger arguments return NaN loudly
__host__ __device__ __forceinline__ float cosf_glibc(float y)
{
    double x = (double)y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return sincos_poly(x, x * x, false, 1);
    }
    if (abstop12(y) < abstop12(120.0f)) {
        int n;
        x = reduce_fast(x, &n);
        const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        return sincos_poly(x * s, x * x, (n & 2) != 0, n ^ 1);
    }
    return asfloat(0x7fc00000u);
}

__host__ __device__ __forceinline__ float sinf_glibc(float y)
{
    double x = (double)y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return sincos_poly(x, x * x, false, 0);
    }
    if (abstop12(y) < abstop12(120.0f)) {
        int n;
        x = reduce_fast(x, &n);
        const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        return sincos_poly(x * s, x * x, (n & 2) != 0, n);
    }
    return asfloat(0x7fc00000u);
}

// fdlibm s_atanf.c
__host__ __device__ __forceinline__ float atanf_glibc(float x)
{
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f,
                          -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f,
                          1.6285819933e-02f};
    const int32_t hx = (int32_t)asuint(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) { // |x| >= 2^25
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) { // |x| < 0.4375
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return hx < 0 ? -r : r;
}

// fdlibm e_atan2f.c (finite arguments; infinities do not occur: the arguments are pixel coordinate differences)
__host__ __device__ __forceinline__ float atan2f_glibc(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = (int32_t)asuint(x), ix = hx & 0x7fffffff;
    const int32_t hy = (int32_t)asuint(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanf_glibc(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        if (m < 2) return y;
        return m == 2 ? pi + tiny : -pi - tiny;
    }
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000 || iy == 0x7f800000) return asfloat(0x7fc00000u);
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = atanf_glibc(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

} // namespace plf_libm
