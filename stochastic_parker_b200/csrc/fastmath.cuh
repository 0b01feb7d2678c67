// fastmath.cuh -- branch-free FP64 elementary functions for the production push kernel.
//
// The push is one long dependent FP64 chain per particle (profiles/r01b_push_coop.md): the
// libdevice log/exp/division/sqrt sequences cost 50-90 SASS instructions each and every one
// ends in a slow-path branch that splits the basic block and stops the scheduler from
// interleaving independent work.  The versions here are straight-line code for the domain the
// kernel guarantees (positive, normal arguments), each within ~2 ulp:
//   rcp    MUFU.RCP64H seed (2^-23) + one cubic Newton step               4 instr
//   rsqrt  MUFU.RSQ64H seed (2^-22) + one cubic Newton step               6 instr
//   log    fdlibm e_log.c reduction and Lg1..Lg7 polynomial               ~36 instr
//   exp    round-to-nearest k via the 1.5*2^52 trick, degree-13 Taylor    ~24 instr
// __host__ versions (seeded from single-precision division) exist only so that
// tests/fastmath_host_check.cpp can measure the polynomial error on the CPU.
#pragma once
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define FM_HD __host__ __device__ __forceinline__
#else
#define FM_HD inline
#include <cmath>
#endif

namespace fm {

// Polynomial coefficients and magic numbers of log_pos / exp_mid.  As literals every FP64 constant costs two UMOV
// (uniform-register immediates) right before its use -- 50 of the 1070 warp-instructions of a C1 step
// (profiles/r02w_c1_by_line.txt); from a __constant__ table (FM_CONST_TABLE=1) one LDCU.128 brings two of them.
// MEASURED: no gain in the kernels of the named configs (profiles/README.md, call X: C1 +0.2 %, C3 +1.3 %, C4 -2.3 %,
// C5 +0.4 %), +1-1.5 % in the general-pusher kernels, whose steps take ten powers (call K2).  The default keeps the
// literals, the Makefile turns the table on for the two general-pusher translation units; the host build
// (tests/fastmath_host_check.cpp) always uses the literals.
#ifndef FM_CONST_TABLE
#define FM_CONST_TABLE 0
#endif
#define FM_TABLE_VALUES                                                                                         \
    1.531383769920937332e-01, 2.222219843214978396e-01, 3.999999999940941908e-01, 1.479819860511658591e-01,     \
    1.818357216161805012e-01, 2.857142874366239149e-01, 6.666666666666735130e-01, 4503601774854144.0,           \
    6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.4426950408889634074, 6755399441055744.0,          \
    1.0 / 479001600.0, 1.0 / 3628800.0, 1.0 / 6227020800.0, 1.0 / 39916800.0, 1.0 / 40320.0, 1.0 / 362880.0,    \
    1.0 / 720.0, 1.0 / 5040.0, 1.0 / 24.0, 1.0 / 120.0, 0.5, 1.0 / 6.0
enum : int { K_LG6 = 0, K_LG4, K_LG2, K_LG7, K_LG5, K_LG3, K_LG1, K_TWO52B, K_LN2HI, K_LN2LO, K_LOG2E, K_MAGIC,
             K_E12, K_E10, K_E13, K_E11, K_E8, K_E9, K_E6, K_E7, K_E4, K_E5, K_E2, K_E3, K_COUNT };
FM_HD constexpr double kval(int i)
{
    constexpr double t[K_COUNT] = {FM_TABLE_VALUES};
    return t[i];
}
#if defined(__CUDACC__) && FM_CONST_TABLE
static __constant__ double ktab_dev[K_COUNT] = {FM_TABLE_VALUES};
#endif
template <int I> FM_HD double K()
{
#if defined(__CUDA_ARCH__) && FM_CONST_TABLE
    return ktab_dev[I];
#else
    constexpr double v = kval(I);
    return v;
#endif
}

FM_HD int hi_word(double x)
{
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    uint64_t b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
FM_HD int lo_word(double x)
{
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    uint64_t b; memcpy(&b, &x, 8); return (int)(uint32_t)b;
#endif
}
FM_HD double make_double(int hi, int lo)
{
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &b, 8); return d;
#endif
}
FM_HD double fma_(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return fma(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}

FM_HD double rcp_seed(double x)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#else
    return (double)(1.0f / (float)x);
#endif
}
FM_HD double rsqrt_seed(double x)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#else
    return (double)(1.0f / sqrtf((float)x));
#endif
}

// 1/x for normal x (|x| in ~[1e-150, 1e150]); e = 1 - x r0 ~ 2^-23, r = r0 (1 + e + e^2)
FM_HD double rcp(double x)
{
    const double r0 = rcp_seed(x);
    const double e = fma_(-x, r0, 1.0);
    return fma_(r0, fma_(e, e, e), r0);
}

// x^-1/2 for normal x > 0; e = 1 - x y0^2, y = y0 (1 + e/2 + 3 e^2/8)
FM_HD double rsqrt(double x)
{
    const double y0 = rsqrt_seed(x);
    const double t = x * y0;
    const double e = fma_(-t, y0, 1.0);
    return fma_(y0 * e, fma_(0.375, e, 0.5), y0);
}

// sqrt(x), x >= 0 (0 -> 0); one Heron correction after x * rsqrt(x)
FM_HD double sqrt_pos(double x)
{
    const double y = rsqrt(x > 0.0 ? x : 1.0);
    const double s = x * y;
    return fma_(fma_(-s, s, x), 0.5 * y, s);
}

// natural log of a positive normal double (fdlibm e_log.c: f = m - 1 with m in
// [sqrt(1/2), sqrt(2)), s = f/(2+f), log(1+f) = f - hfsq + s (hfsq + R(s^2)))
FM_HD double log_pos(double x)
{
    int hi = hi_word(x);
    const int lo = lo_word(x);
    int k = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;  // m in [1, 2)
    const bool big = hi >= 0x3ff6a09f;     // m >= sqrt(2) (to 20 bits): use m/2
    hi = big ? hi - 0x00100000 : hi;
    k = big ? k + 1 : k;
    const double f = make_double(hi, lo) - 1.0;
    const double s = f * rcp(2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * fma_(w, fma_(w, K<K_LG6>(), K<K_LG4>()), K<K_LG2>());
    const double t2 = z * fma_(w, fma_(w, fma_(w, K<K_LG7>(), K<K_LG5>()), K<K_LG3>()), K<K_LG1>());
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    // k as a double without an integer->float conversion: 2^52 + 2^31 + k, minus the bias
    const double dk = make_double(0x43300000, k ^ (int)0x80000000) - K<K_TWO52B>();
    const double l1p = f - (hfsq - s * (hfsq + R));
    return fma_(dk, K<K_LN2HI>(), fma_(dk, K<K_LN2LO>(), l1p));
}

// exp(x) for |x| < 700 (result normal)
FM_HD double exp_mid(double x)
{
    const double kMagic = K<K_MAGIC>();  // 1.5 * 2^52: t's low word is round(x log2 e)
    const double t = fma_(x, K<K_LOG2E>(), kMagic);
    const int k = lo_word(t);
    const double kd = t - kMagic;
    double r = fma_(-kd, K<K_LN2HI>(), x);
    r = fma_(-kd, K<K_LN2LO>(), r);  // |r| <= 0.3466
    const double r2 = r * r;
    // Taylor to r^13, even and odd parts as two short Horner chains
    double pe = fma_(r2, K<K_E12>(), K<K_E10>());
    double po = fma_(r2, K<K_E13>(), K<K_E11>());
    pe = fma_(r2, pe, K<K_E8>());
    po = fma_(r2, po, K<K_E9>());
    pe = fma_(r2, pe, K<K_E6>());
    po = fma_(r2, po, K<K_E7>());
    pe = fma_(r2, pe, K<K_E4>());
    po = fma_(r2, po, K<K_E5>());
    pe = fma_(r2, pe, K<K_E2>());
    po = fma_(r2, po, K<K_E3>());
    // exp(r) = 1 + r + r^2 (pe + r po)
    const double p = fma_(r2, fma_(r, po, pe), r) + 1.0;
    return make_double(hi_word(p) + (k << 20), lo_word(p));
}

}  // namespace fm
