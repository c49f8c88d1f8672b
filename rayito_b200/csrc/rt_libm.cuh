// sinf / cosf / powf with the reference's libm bits.
//
// The reference calls the C library's float cos/sin/pow for every sample direction
// (RSampling.h:449-518, RMaterial.h:246,283-287); they decide where rays go, so a
// last-bit difference sends a path elsewhere.  The reference's libm is a THIRD-PARTY
// dependency that is not in /root/reference: GNU libc 2.39 (Ubuntu 24.04,
// libm.so.6), whose sinf/cosf/powf are the published ARM Optimized Routines
// algorithms (glibc sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, e_powf.c, sincosf.h,
// sincosf_data.c, e_powf_log2_data.c, e_exp2f_data.c): double-precision polynomial
// kernels, one final rounding to float.  This file restates those algorithms with the
// library's constants and the operation fusing of its x86-64 FMA build (the variant
// glibc's ifunc selects on every AVX2+FMA host, i.e. the build box and the B200 box):
// every a*b+c of the source is ONE fused multiply-add there, so it is fma() here.
//
// Pinned by tests/test_libm_cpu.py: the host build of this very file is compared
// bit-for-bit with the C library over millions of arguments in the ranges the path
// uses.  Ranges outside the fast paths (|x| >= 120 for sin/cos, x <= 0 / subnormal /
// overflowing powf) fall back to double-precision CUDA math, rounded once.
#ifndef RAYITO_B200_RT_LIBM_CUH
#define RAYITO_B200_RT_LIBM_CUH

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define RT_HD __host__ __device__
#else
#define RT_HD
#endif

RT_HD static inline uint32_t rtm_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
RT_HD static inline float rtm_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
RT_HD static inline uint64_t rtm_d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
RT_HD static inline double rtm_u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

// __sincosf_table[2] of glibc 2.39: sign = {1,-1,-1,1}, hpi_inv = 2/pi * 2^24, hpi = pi/2,
// cosine coefficients c0..c4 (negated in the second table), sine coefficients s1..s3
#define RTM_HPI_INV 0x1.45F306DC9C883p+23
#define RTM_HPI 0x1.921FB54442D18p+0
#define RTM_C0 0x1p0
#define RTM_C1 -0x1.ffffffd0c621cp-2
#define RTM_C2 0x1.55553e1068f19p-5
#define RTM_C3 -0x1.6c087e89a359dp-10
#define RTM_C4 0x1.99343027bf8c3p-16
#define RTM_S1 -0x1.555545995a603p-3
#define RTM_S2 0x1.1107605230bc4p-7
#define RTM_S3 -0x1.994eb3774cf24p-13

// sinf_poly (sincosf.h): sine polynomial for even n, cosine polynomial for odd n;
// `flip` selects the second table (all cosine coefficients negated: exact)
RT_HD static inline float rtm_sinf_poly(double x, double x2, bool flip, int n)
{
    if ((n & 1) == 0)
    {
        double x3 = x * x2;
        double s1 = fma(x2, RTM_S3, RTM_S2);
        double x7 = x3 * x2;
        double s = fma(x3, RTM_S1, x);
        return (float)fma(s1, x7, s);
    }
    const double f = flip ? -1.0 : 1.0;
    double x4 = x2 * x2;
    double c2 = fma(x2, f * RTM_C4, f * RTM_C3);
    double c1 = fma(x2, f * RTM_C1, f * RTM_C0);
    double x6 = x4 * x2;
    double c = fma(x4, f * RTM_C2, c1);
    return (float)fma(c2, x6, c);
}

RT_HD static inline uint32_t rtm_abstop12(float x) { return (rtm_f2u(x) >> 20) & 0x7ff; }

// reduce_fast (sincosf.h): n = round(x * 2/pi), x - n*pi/2 with one fused operation
RT_HD static inline double rtm_reduce_fast(double x, int& n)
{
    double r = x * RTM_HPI_INV;
    n = ((int32_t)r + 0x800000) >> 24;
    return fma(-(double)n, RTM_HPI, x);
}

RT_HD static inline double rtm_quadrant_sign(int n) { return ((n + 1) & 2) ? -1.0 : 1.0; }   // {1,-1,-1,1}[n & 3]

// s_sinf.c, fast paths
RT_HD static inline float rtm_sinf(float y)
{
    double x = (double)y;
    uint32_t top = rtm_abstop12(y);
    if (top < 0x3f4)                         // |y| < pi/4
    {
        if (top < 0x398)                     // |y| < 2^-12
            return y;
        return rtm_sinf_poly(x, x * x, false, 0);
    }
    if (top < 0x42f)                         // |y| < 120
    {
        int n;
        x = rtm_reduce_fast(x, n);
        return rtm_sinf_poly(x * rtm_quadrant_sign(n), x * x, (n & 2) != 0, n);
    }
    return (float)sin(x);
}

// s_cosf.c, fast paths
RT_HD static inline float rtm_cosf(float y)
{
    double x = (double)y;
    uint32_t top = rtm_abstop12(y);
    if (top < 0x3f4)
    {
        if (top < 0x398)
            return 1.0f;
        return rtm_sinf_poly(x, x * x, false, 1);
    }
    if (top < 0x42f)
    {
        int n;
        x = rtm_reduce_fast(x, n);
        return rtm_sinf_poly(x * rtm_quadrant_sign(n), x * x, (n & 2) != 0, n ^ 1);
    }
    return (float)cos(x);
}

// sinf and cosf of the same argument share the range reduction (same bits as the two
// separate calls: the reduction is a pure function of the argument)
RT_HD static inline void rtm_sincosf(float y, float& sn, float& cs)
{
    double x = (double)y;
    uint32_t top = rtm_abstop12(y);
    if (top >= 0x3f4 && top < 0x42f)
    {
        int n;
        x = rtm_reduce_fast(x, n);
        double xs = x * rtm_quadrant_sign(n), x2 = x * x;
        bool flip = (n & 2) != 0;
        sn = rtm_sinf_poly(xs, x2, flip, n);
        cs = rtm_sinf_poly(xs, x2, flip, n ^ 1);
        return;
    }
    sn = rtm_sinf(y);
    cs = rtm_cosf(y);
}

// Tables: one copy in constant memory for the device, one for the host build
#define RTM_LOG2_TAB_INIT { \
        { 0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2 }, { 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2 }, \
        { 0x1.49539f0f010bp+0, -0x1.7418b0a1fb77bp-2 },  { 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2 }, \
        { 0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2 }, { 0x1.25e227b0b8eap+0, -0x1.97c1d1b3b7afp-3 }, \
        { 0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3 }, { 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4 }, \
        { 0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5 }, { 0x1p+0, 0x0p+0 }, \
        { 0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4 },  { 0x1.ca4b31f026aap-1, 0x1.476a9543891bap-3 }, \
        { 0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2 }, \
        { 0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2 },  { 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2 } }
#define RTM_EXP2_TAB_INIT { \
        0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
        0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
        0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
        0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
        0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
        0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
        0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
        0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull }
static const double rtm_log2_tab_host[16][2] = RTM_LOG2_TAB_INIT;
static const uint64_t rtm_exp2_tab_host[32] = RTM_EXP2_TAB_INIT;
#ifdef __CUDACC__
__constant__ double rtm_log2_tab_dev[16][2] = RTM_LOG2_TAB_INIT;
__constant__ uint64_t rtm_exp2_tab_dev[32] = RTM_EXP2_TAB_INIT;
#endif
#ifdef __CUDA_ARCH__
#define RTM_TAB(name) name##_dev
#else
#define RTM_TAB(name) name##_host
#endif

// __powf_log2_data.tab (invc, logc) and __exp2f_data.tab of glibc 2.39
RT_HD static inline void rtm_log2_entry(uint32_t i, double& invc, double& logc)
{
    invc = RTM_TAB(rtm_log2_tab)[i][0];
    logc = RTM_TAB(rtm_log2_tab)[i][1];
}

RT_HD static inline uint64_t rtm_exp2_entry(uint32_t i)
{
    return RTM_TAB(rtm_exp2_tab)[i];
}

// e_powf.c, main path: x positive and normal, no overflow / underflow of the result.
// Anything else goes through double-precision pow (never the case on the render path:
// bases are |cos| in (0, 1] or pixel values, exponents positive and modest).
RT_HD static inline float rtm_powf(float x, float y)
{
    uint32_t ix = rtm_f2u(x), iy = rtm_f2u(y);
    bool x_ok = ix - 0x00800000u < 0x7f800000u - 0x00800000u;                 // positive normal
    bool y_ok = (2u * iy - 1u) < (2u * 0x7f800000u - 1u);                     // not 0, inf, nan
    if (x_ok && y_ok)
    {
        // log2_inline
        uint32_t tmp = ix - 0x3f330000u;
        uint32_t i = (tmp >> 19) & 15u;
        uint32_t top = tmp & 0xff800000u;
        uint32_t iz = ix - top;
        int32_t k = (int32_t)top >> 23;
        double invc, logc;
        rtm_log2_entry(i, invc, logc);
        double z = (double)rtm_u2f(iz);
        double r = fma(z, invc, -1.0);
        double y0 = logc + (double)k;
        const double A0 = 0x1.27616c9496e0bp-2, A1 = -0x1.71969a075c67ap-2, A2 = 0x1.ec70a6ca7baddp-2,
                     A3 = -0x1.7154748bef6c8p-1, A4 = 0x1.71547652ab82bp+0;
        double r2 = r * r;
        double yy = fma(A0, r, A1);
        double p = fma(A2, r, A3);
        double r4 = r2 * r2;
        double q = fma(A4, r, y0);
        q = fma(p, r2, q);
        double logx = fma(yy, r4, q);
        double ylogx = (double)y * logx;
        if (((rtm_d2u(ylogx) >> 47) & 0xffff) >= 0x80bf)
        {
            // |y*log2(x)| >= 126 (e_powf.c): overflow, underflow, or still exp2_inline
            if (ylogx > 0x1.fffffffd1d571p+6)
                return rtm_u2f(0x7f800000u);                    // __math_oflowf
            if (ylogx <= -150.0)
                return 0.0f;                                    // __math_uflowf
            if (ylogx < -149.0)
                return 0x1.4p-75f * 0x1.4p-75f;                 // __math_may_uflowf
        }
        {
            // exp2_inline, sign_bias = 0
            const double SHIFT = 0x1.8p+47;                 // 0x1.8p+52 / 32
            const double C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
            double kd = ylogx + SHIFT;
            uint64_t ki = rtm_d2u(kd);
            kd -= SHIFT;
            double rr = ylogx - kd;
            uint64_t t = rtm_exp2_entry((uint32_t)(ki & 31u));
            t += ki << 47;
            double s = rtm_u2d(t);
            double zz = fma(C0, rr, C1);
            double rr2 = rr * rr;
            double yv = fma(C2, rr, 1.0);
            yv = fma(zz, rr2, yv);
            yv = yv * s;
            return (float)yv;
        }
    }
    return (float)pow((double)x, (double)y);
}

#endif // RAYITO_B200_RT_LIBM_CUH
