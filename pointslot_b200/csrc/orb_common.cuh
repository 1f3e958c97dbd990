// orb_common.cuh - device helpers shared by the two descriptor kernels (describe.cu: the ORB-SLAM extractor of
// ORBextractor.cc; cvorb.cu: OpenCV's own cv::ORB used for the dense object features, Frame.cc:2623-2627): the rBRIEF
// pattern table, the radius-15 disc, BORDER_REFLECT_101, cv::fastAtan2 and glibc's cosf / sinf.  Everything sits in an
// anonymous namespace: each translation unit gets its own copy of the tables.
#pragma once
#include <stdint.h>

namespace {

// pattern, re-laid out at compile time as [t][lane] words (test 8*lane + t; bytes x0,y0,x1,y1) so that a
// warp reads one coalesced 128-byte line per test index
struct PatTable { uint32_t w[256]; };
constexpr PatTable make_pat()
{
    constexpr int8_t raw[1024] = {
#include "../../include/pslot_orb_pattern.inc"
    };
    PatTable t{};
    for (int i = 0; i < 256; ++i) {
        const int tt = i >> 5, ln = i & 31;
        const int8_t *p = raw + 4 * (8 * ln + tt);
        t.w[i] = (uint32_t)(uint8_t)p[0] | ((uint32_t)(uint8_t)p[1] << 8) | ((uint32_t)(uint8_t)p[2] << 16) |
                 ((uint32_t)(uint8_t)p[3] << 24);
    }
    return t;
}
__device__ const PatTable g_pat = make_pat();
// the same table as floats, [t][lane] = (x0, y0, x1, y1) of test 8*lane + t: one coalesced 16-byte load per lane and
// test instead of four byte extractions and four int -> float conversions (the conversion unit is quarter rate)
struct PatTableF { float4 v[256]; };
constexpr PatTableF make_patf()
{
    constexpr int8_t raw[1024] = {
#include "../../include/pslot_orb_pattern.inc"
    };
    PatTableF t{};
    for (int i = 0; i < 256; ++i) {
        const int tt = i >> 5, ln = i & 31;
        const int8_t *p = raw + 4 * (8 * ln + tt);
        t.v[i].x = (float)p[0]; t.v[i].y = (float)p[1]; t.v[i].z = (float)p[2]; t.v[i].w = (float)p[3];
    }
    return t;
}
__device__ const PatTableF g_patf = make_patf();
// umax[v]: half-width of the radius-15 disc at row v (ORBextractor.cc:454-469)
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// cv::fastAtan2 (OpenCV 4.13 atanImpl<float>), degrees; float32, one rounding per operation.
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// a = cosf(angle*factorPI), b = sinf(...) exactly as the reference binary gets them (parity rule 3): `(float)cos(float)`
// at ORBextractor.cc:112-113 is glibc's cosf, whose algorithm (glibc >= 2.28, sysdeps/ieee754/flt-32/s_sincosf.h: reduction
// by pi/2 with a 2^24-prescaled multiply, even / odd polynomial in double, one rounding to float) is restated here with
// explicit IEEE double operations.  The CPU checker carries the same restatement (orc_glibc_cosf), verified
// against libm on every float in [0, 2 pi] (profiles/r02_sincosf_sweep.txt); tools/steer_sweep.py checks this device
// version against it on every float angle in [0, 360].
__device__ __forceinline__ float glibc_sincos_poly(double x, double x2, bool negcos, int n)
{
    if ((n & 1) == 0) {
        const double x3 = __dmul_rn(x, x2), s1 = __fma_rn(x2, -0x1.994eb3774cf24p-13, 0x1.1107605230bc4p-7);
        const double x7 = __dmul_rn(x3, x2), s = __fma_rn(x3, -0x1.555545995a603p-3, x);
        return __double2float_rn(__fma_rn(x7, s1, s));
    }
    const double g = negcos ? -1.0 : 1.0;
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __fma_rn(x2, g * 0x1.99343027bf8c3p-16, g * -0x1.6c087e89a359dp-10);
    const double c1 = __fma_rn(x2, g * -0x1.ffffffd0c621cp-2, g);
    const double x6 = __dmul_rn(x4, x2), c = __fma_rn(x4, g * 0x1.55553e1068f19p-5, c1);
    return __double2float_rn(__fma_rn(x6, c2, c));
}
__device__ __forceinline__ float glibc_sincosf(float y, bool cosine)
{
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ffu;
    double x = (double)y;
    if (top < 0x3f4u) {                      // |y| < pi/4: abstop12(0x1.921FB6p-1f)
        if (top < 0x398u) return cosine ? 1.0f : y; // |y| < 2^-12
        return glibc_sincos_poly(x, __dmul_rn(x, x), false, cosine ? 1 : 0);
    }
    const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __fma_rn(-(double)n, 0x1.921FB54442D18p0, x);
    const double sgn = ((n + 1) & 2) ? -1.0 : 1.0; // sign[n & 3] = {1, -1, -1, 1}
    return glibc_sincos_poly(__dmul_rn(x, sgn), __dmul_rn(x, x), (n & 2) != 0, cosine ? n ^ 1 : n);
}
__device__ __forceinline__ void steer_coeffs(float angle_deg, float &a, float &b)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
    const float ang = __fmul_rn(angle_deg, factorPI);
    a = glibc_sincosf(ang, true);
    b = glibc_sincosf(ang, false);
}

} // namespace
