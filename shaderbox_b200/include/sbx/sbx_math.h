// sbx_math.h -- fp32 transcendentals of the device operator library, bit-compatible with the
// libm a C++ build of the reference links (glibc 2.39, x86-64 FMA variants).
//
// Why this exists.  The reference's value noise hashes with fract(sin(n)*753.5453123)
// (src/noise_iq.h:5-9): a 1-ulp difference in sin(n) moves the hash by ~1e-4..1e-3, and the march
// loops (src/app_clouds.h:182-198, src/app_planet.h:328-342) amplify that into visible pixel
// differences.  CUDA's own sinf/expf/powf are not bit-identical to glibc's, so the device library
// carries its own: the same published algorithms glibc 2.39 uses -- Arm Optimized Routines'
// sinf / cosf / expf / powf (double-precision polynomial kernels, Szabolcs Nagy / Wilco Dijkstra,
// sysdeps/ieee754/flt-32/{s_sinf,s_cosf,e_expf,e_powf}.c) -- evaluated with the same operation
// order and the same fused multiply-adds as the x86-64 `_fma` ifunc variants.  glibc is a
// third-party dependency of the reference's C++ build, not part of /root/reference; constants
// are the published table values (tools/dump_libm_consts.py shows where they sit in libm.so.6).
//
// The header is host+device: tests/ compile it with g++ (-mfma -ffp-contract=off) and compare it
// against libm over the whole float range; the kernels compile it with nvcc / NVRTC
// (--fmad=false).  Every fma() below is an explicit, required fusion.
#ifndef SBX_MATH_H_
#define SBX_MATH_H_

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define SBX_MATH_FN __device__ __forceinline__
#define SBX_MATH_COLD static __device__ __noinline__   /* rare paths: one copy per kernel, not one per call site */
#define SBX_DEVICE_CODE 1
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#define SBX_MATH_FN static inline
#define SBX_MATH_COLD static inline
#endif

#if defined(SBX_DEVICE_CODE)
typedef unsigned int sbx_u32;
typedef unsigned long long sbx_u64;
typedef long long sbx_i64;
SBX_MATH_FN sbx_u32 sbx_f2u(float f) { return __float_as_uint(f); }
SBX_MATH_FN float sbx_u2f(sbx_u32 u) { return __uint_as_float(u); }
SBX_MATH_FN sbx_u64 sbx_d2u(double d) { return (sbx_u64)__double_as_longlong(d); }
SBX_MATH_FN double sbx_u2d(sbx_u64 u) { return __longlong_as_double((long long)u); }
SBX_MATH_FN double sbx_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
SBX_MATH_FN float sbx_d2f(double d) { return __double2float_rn(d); }
SBX_MATH_FN float sbx_sqrtf(float a) { return __fsqrt_rn(a); }
SBX_MATH_FN float sbx_divf(float a, float b) { return __fdiv_rn(a, b); }
#else
typedef uint32_t sbx_u32;
typedef uint64_t sbx_u64;
typedef int64_t sbx_i64;
SBX_MATH_FN sbx_u32 sbx_f2u(float f) { sbx_u32 u; memcpy(&u, &f, 4); return u; }
SBX_MATH_FN float sbx_u2f(sbx_u32 u) { float f; memcpy(&f, &u, 4); return f; }
SBX_MATH_FN sbx_u64 sbx_d2u(double d) { sbx_u64 u; memcpy(&u, &d, 8); return u; }
SBX_MATH_FN double sbx_u2d(sbx_u64 u) { double d; memcpy(&d, &u, 8); return d; }
SBX_MATH_FN double sbx_fma(double a, double b, double c) { return __builtin_fma(a, b, c); }
SBX_MATH_FN float sbx_d2f(double d) { return (float)d; }
SBX_MATH_FN float sbx_sqrtf(float a) { return sqrtf(a); }
SBX_MATH_FN float sbx_divf(float a, float b) { return a / b; }
#endif

// ---------------------------------------------------------------------------------------------
// Tables.  On the device they live in shared memory: the kernel prologue stages one "LUT block"
// from HBM with a single TMA bulk copy (cp.async.bulk, see sbx_kernel.cuh) because per-lane
// table indices diverge and __constant__ memory would serialise them.  Layout of the block
// (all 8-byte words):   [0,32) exp2 table   [32,64) log2 table (invc,logc pairs)
// ---------------------------------------------------------------------------------------------
#define SBX_LUT_EXP2_WORDS 32
#define SBX_LUT_LOG2_WORDS 32
#define SBX_LUT_MATH_BYTES ((SBX_LUT_EXP2_WORDS + SBX_LUT_LOG2_WORDS) * 8)

// 2^(i/32) with the exponent contribution (i << 47) pre-subtracted (e_exp2f_data.c layout)
#define SBX_EXP2_TABLE_INIT                                                                      \
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,  \
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,  \
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,  \
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,  \
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,  \
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,  \
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,  \
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull

// log2 table: {1/c, log2(c)} for 16 sub-intervals of [0x1.66p-1, 0x1.66p0) (e_powf_log2_data.c)
#define SBX_LOG2_TABLE_INIT                                                                      \
    0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2,      \
    0x1.49539f0f010bp+0, -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2,       \
    0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8eap+0, -0x1.97c1d1b3b7afp-3,        \
    0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4,      \
    0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1p+0, 0x0p+0,                                   \
    0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4, 0x1.ca4b31f026aap-1, 0x1.476a9543891bap-3,         \
    0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3, 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2,        \
    0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2, 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2

#if defined(SBX_DEVICE_CODE)
// the kernel's dynamic shared memory starts with the LUT block
extern __shared__ __align__(16) unsigned char sbx_smem[];
SBX_MATH_FN sbx_u64 sbx_exp2_tab(sbx_u32 i) { return ((const sbx_u64*)sbx_smem)[i]; }
SBX_MATH_FN double sbx_log2_invc(sbx_u32 i) { return ((const double*)sbx_smem)[SBX_LUT_EXP2_WORDS + 2 * i]; }
SBX_MATH_FN double sbx_log2_logc(sbx_u32 i) { return ((const double*)sbx_smem)[SBX_LUT_EXP2_WORDS + 2 * i + 1]; }
#else
static const sbx_u64 sbx_exp2_table_[32] = {SBX_EXP2_TABLE_INIT};
static const double sbx_log2_table_[32] = {SBX_LOG2_TABLE_INIT};
SBX_MATH_FN sbx_u64 sbx_exp2_tab(sbx_u32 i) { return sbx_exp2_table_[i]; }
SBX_MATH_FN double sbx_log2_invc(sbx_u32 i) { return sbx_log2_table_[2 * i]; }
SBX_MATH_FN double sbx_log2_logc(sbx_u32 i) { return sbx_log2_table_[2 * i + 1]; }
#endif

// ---------------------------------------------------------------------------------------------
// sinf / cosf   (s_sinf.c, s_cosf.c, s_sincosf.h)
// ---------------------------------------------------------------------------------------------
// Polynomial on [-pi/4, pi/4] in double.  sgn = +1 for table 0, -1 for table 1 of s_sincosf_data.c
// (the second table is the first with c0,c1,c2.. / s.. negated where the quadrant requires it).
SBX_MATH_FN float sbx_sincos_poly(double x, double x2, int n, int neg) {
    // coefficients of table[0]; table[1] negates c0..c4 only
    const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5,
                 c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double t1 = sbx_fma(x2, s3, s2);
        double x7 = x3 * x2;
        double s = sbx_fma(x3, s1, x);
        return sbx_d2f(sbx_fma(x7, t1, s));
    } else {
        double k = neg ? -1.0 : 1.0;
        double x4 = x2 * x2;
        double t2 = sbx_fma(x2, k * c4, k * c3);
        double t1 = sbx_fma(x2, k * c1, k * c0);
        double x6 = x4 * x2;
        double c = sbx_fma(x4, k * c2, t1);
        return sbx_d2f(sbx_fma(x6, t2, c));
    }
}

// 4/pi as overlapping 32-bit words (__inv_pio4)
SBX_MATH_FN sbx_u32 sbx_inv_pio4(int i) {
#if defined(SBX_DEVICE_CODE)
    // 24 words, selected by a 4-bit index and +0/+4/+8: keep them in immediate-friendly form
    switch (i) {
#define SBX_IPW(k, v) case k: return v;
#else
    static const sbx_u32 w[24] = {
#define SBX_IPW(k, v) v,
#endif
        SBX_IPW(0, 0xa2u) SBX_IPW(1, 0xa2f9u) SBX_IPW(2, 0xa2f983u) SBX_IPW(3, 0xa2f9836eu)
        SBX_IPW(4, 0xf9836e4eu) SBX_IPW(5, 0x836e4e44u) SBX_IPW(6, 0x6e4e4415u) SBX_IPW(7, 0x4e441529u)
        SBX_IPW(8, 0x441529fcu) SBX_IPW(9, 0x1529fc27u) SBX_IPW(10, 0x29fc2757u) SBX_IPW(11, 0xfc2757d1u)
        SBX_IPW(12, 0x2757d1f5u) SBX_IPW(13, 0x57d1f534u) SBX_IPW(14, 0xd1f534ddu) SBX_IPW(15, 0xf534ddc0u)
        SBX_IPW(16, 0x34ddc0dbu) SBX_IPW(17, 0xddc0db62u) SBX_IPW(18, 0xc0db6295u) SBX_IPW(19, 0xdb629599u)
        SBX_IPW(20, 0x6295993cu) SBX_IPW(21, 0x95993c43u) SBX_IPW(22, 0x993c4390u) SBX_IPW(23, 0x3c439041u)
#undef SBX_IPW
#if defined(SBX_DEVICE_CODE)
    }
    return 0u;
#else
    };
    return w[i];
#endif
}

// |x| < 120: n = round(x * 2/pi) via a 2^24-scaled multiply, x - n*pi/2 in one fma
SBX_MATH_FN double sbx_reduce_fast(double x, int* np) {
    const double hpi_inv = 0x1.45f306dc9c883p+23, hpi = 0x1.921fb54442d18p+0;
    double r = x * hpi_inv;
    int n = ((int)r + 0x800000) >> 24;
    *np = n;
    return sbx_fma(-(double)n, hpi, x);
}

// |x| >= 120: 96 bits of 4/pi times the 24-bit mantissa, keep the top 2 bits as the quadrant
typedef struct sbx_reduced { double x; int n; } sbx_reduced;   // returned by value: the out-of-line copy then needs no stack slot in its callers
SBX_MATH_COLD sbx_reduced sbx_reduce_large_cold(sbx_u32 xi) {
    const int base = (int)((xi >> 26) & 15u);
    const int shift = (int)((xi >> 23) & 7u);
    sbx_u64 n, res0, res1, res2;
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    res0 = (sbx_u64)(sbx_u32)(xi * sbx_inv_pio4(base));
    res1 = (sbx_u64)xi * sbx_inv_pio4(base + 4);
    res2 = (sbx_u64)xi * sbx_inv_pio4(base + 8);
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ull << 61)) >> 62;
    res0 -= n << 62;
    sbx_reduced r;
    r.n = (int)n;
    r.x = (double)(sbx_i64)res0 * 0x1.921fb54442d18p-62;
    return r;
}
SBX_MATH_FN double sbx_reduce_large(sbx_u32 xi, int* np) {
    const sbx_reduced r = sbx_reduce_large_cold(xi);
    *np = r.n;
    return r.x;
}

SBX_MATH_FN float sbx_sinf(float y) {
    double x = (double)y;
    const sbx_u32 top = (sbx_f2u(y) >> 20) & 0x7ffu;
    int n;
    if (top < 0x3f4u) {                      // |y| < pi/4
        if (top < 0x398u) return y;          // |y| < 2^-12: sin y == y in fp32
        return sbx_sincos_poly(x, x * x, 0, 0);
    } else if (top < 0x42fu) {               // |y| < 120
        x = sbx_reduce_fast(x, &n);
        const double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   // sign[] = {1,-1,-1,1}
        return sbx_sincos_poly(x * sg, x * x, n, n & 2);
    } else if (top < 0x7f8u) {
        const sbx_u32 xi = sbx_f2u(y);
        const int sign = (int)(xi >> 31);
        x = sbx_reduce_large(xi, &n);
        const int q = (n + sign) & 3;
        const double sg = (q == 1 || q == 2) ? -1.0 : 1.0;
        return sbx_sincos_poly(x * sg, x * x, n, (n + sign) & 2);
    }
    return y - y;                            // inf / nan -> nan
}

SBX_MATH_FN float sbx_cosf(float y) {
    double x = (double)y;
    const sbx_u32 top = (sbx_f2u(y) >> 20) & 0x7ffu;
    int n;
    if (top < 0x3f4u) {
        if (top < 0x398u) return 1.0f;
        return sbx_sincos_poly(x, x * x, 1, 0);
    } else if (top < 0x42fu) {
        x = sbx_reduce_fast(x, &n);
        const double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        return sbx_sincos_poly(x * sg, x * x, n ^ 1, n & 2);
    } else if (top < 0x7f8u) {
        const sbx_u32 xi = sbx_f2u(y);
        x = sbx_reduce_large(xi, &n);
        const double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        return sbx_sincos_poly(x * sg, x * x, n ^ 1, n & 2);
    }
    return y - y;
}

// sin and cos of one argument from ONE argument reduction (what glibc's sincosf does with the same
// kernels).  Each output is bit-identical to sbx_sinf(y) / sbx_cosf(y): same n, same reduced x, same
// polynomial branch.  The rotation builders of util.h need both values of every angle.
SBX_MATH_FN void sbx_sincosf(float y, float* sin_out, float* cos_out) {
    const sbx_u32 bits = sbx_f2u(y);
    const sbx_u32 top = (bits >> 20) & 0x7ffu;
    double x = (double)y;
    if (top < 0x3f4u) {                              // |y| < pi/4: no reduction
        if (top < 0x398u) { *sin_out = y; *cos_out = 1.0f; return; }
        const double x2 = x * x;
        *sin_out = sbx_sincos_poly(x, x2, 0, 0);
        *cos_out = sbx_sincos_poly(x, x2, 1, 0);
        return;
    }
    if (top >= 0x7f8u) { *sin_out = *cos_out = y - y; return; }   // inf / nan
    int n, n_sin;
    if (top < 0x42fu) {                              // |y| < 120
        x = sbx_reduce_fast(x, &n);
        n_sin = n;
    } else {                                         // the sine folds the argument's sign into the quadrant
        x = sbx_reduce_large(bits, &n);
        n_sin = n + (int)(bits >> 31);
    }
    const double x2 = x * x;
    const int qs = n_sin & 3, qc = n & 3;
    const double xs = x * ((qs == 1 || qs == 2) ? -1.0 : 1.0);
    const double xc = x * ((qc == 1 || qc == 2) ? -1.0 : 1.0);
    *sin_out = sbx_sincos_poly(xs, x2, n, n_sin & 2);
    *cos_out = sbx_sincos_poly(xc, x2, n ^ 1, n & 2);
}

// ---------------------------------------------------------------------------------------------
// expf   (e_expf.c, N = 32, degree-3 polynomial)
// ---------------------------------------------------------------------------------------------
// the main path of expf: valid for |x| < 88 (biased exponent field <= 0x42a), where none of the special cases
// below can apply.  Callers that can bound their argument (a march loop whose optical depth per step is bounded by
// its uniforms) call this directly and save the range test per call.
#if defined(SBX_DEVICE_CODE) && defined(SBX_EXPF_CONSTANT_BANK)
// -DSBX_EXPF_CONSTANT_BANK (kernels with expf in their hot loop: native APP_CLOUDS): the four double constants live
// in constant memory and are fetched with one LDC.64 each, whereas a 64-bit immediate with a non-zero low word
// costs two moves per use (8 of the ~40 instructions of an inlined expf, seven times per in-cloud march step).
// It costs registers, so kernels that call expf once per pixel keep the immediates (PLANET, RAYTRACER: -3 %).
__constant__ double sbx_expf_k[4] = {0x1.71547652b82fep+5, 0x1.c6af84b912394p-20, 0x1.ebfce50fac4f3p-13, 0x1.62e42ff0c52d6p-6};
#define SBX_EXPF_INVLN2N sbx_expf_k[0]
#define SBX_EXPF_C0 sbx_expf_k[1]
#define SBX_EXPF_C1 sbx_expf_k[2]
#define SBX_EXPF_C2 sbx_expf_k[3]
#else
#define SBX_EXPF_INVLN2N 0x1.71547652b82fep+5
#define SBX_EXPF_C0 0x1.c6af84b912394p-20
#define SBX_EXPF_C1 0x1.ebfce50fac4f3p-13
#define SBX_EXPF_C2 0x1.62e42ff0c52d6p-6
#endif
SBX_MATH_FN float sbx_expf_core(float x) {
    const double InvLn2N = SBX_EXPF_INVLN2N, Shift = 0x1.8p52;
    const double C0 = SBX_EXPF_C0, C1 = SBX_EXPF_C1, C2 = SBX_EXPF_C2;
    const double xd = (double)x;
    double kd = sbx_fma(InvLn2N, xd, Shift);
    const sbx_u64 ki = sbx_d2u(kd);
    kd -= Shift;
    const double r = sbx_fma(InvLn2N, xd, -kd);
    sbx_u64 t = sbx_exp2_tab((sbx_u32)(ki & 31u));
    t += ki << 47;
    const double s = sbx_u2d(t);
    const double z = sbx_fma(r, C0, C1);
    const double r2 = r * r;
    double yv = sbx_fma(r, C2, 1.0);
    yv = sbx_fma(z, r2, yv);
    yv = yv * s;
    return sbx_d2f(yv);
}

SBX_MATH_FN float sbx_expf(float x) {
    const sbx_u32 top = (sbx_f2u(x) >> 20) & 0x7ffu;
    if (top > 0x42au) {                                      // |x| >= 88 or nan
        if (sbx_f2u(x) == 0xff800000u) return 0.0f;           // exp(-inf)
        if (top > 0x7f7u) return x + x;                       // +inf, nan
        if (x > 0x1.62e42ep6f) return sbx_u2f(0x7f800000u);   // overflow
        if (x < -0x1.9fe368p6f) return 0.0f;                  // underflow to zero
        if (x < -0x1.9d1d9ep6f) return 0x1p-149f;             // __math_may_uflowf: 0x1.4p-75f squared
        // -0x1.9d1d9ep6 <= x <= -88 falls through: subnormal result, rounded once from double
    }
    return sbx_expf_core(x);
}

// ---------------------------------------------------------------------------------------------
// powf   (e_powf.c: log2 via 16-entry table + degree-5 poly, exp2 via the table above)
// ---------------------------------------------------------------------------------------------
SBX_MATH_FN double sbx_log2_inline(sbx_u32 ix) {
    const double A0 = 0x1.27616c9496e0bp-2, A1 = -0x1.71969a075c67ap-2, A2 = 0x1.ec70a6ca7baddp-2,
                 A3 = -0x1.7154748bef6c8p-1, A4 = 0x1.71547652ab82bp+0;
    const sbx_u32 tmp = ix - 0x3f330000u;
    const sbx_u32 i = (tmp >> 19) & 15u;
    const sbx_u32 topb = tmp & 0xff800000u;
    const sbx_u32 iz = ix - topb;
    const int k = (int)topb >> 23;
    const double invc = sbx_log2_invc(i), logc = sbx_log2_logc(i);
    const double z = (double)sbx_u2f(iz);
    const double r = sbx_fma(z, invc, -1.0);
    const double y0 = logc + (double)k;
    const double r2 = r * r;
    double y = sbx_fma(r, A0, A1);
    const double p = sbx_fma(r, A2, A3);
    const double r4 = r2 * r2;
    double q = sbx_fma(r, A4, y0);
    q = sbx_fma(r2, p, q);
    y = sbx_fma(y, r4, q);
    return y;
}

SBX_MATH_FN float sbx_exp2_inline(double xd, sbx_u32 sign_bias) {
    const double Shift = 0x1.8p+47;
    const double C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
    double kd = xd + Shift;
    const sbx_u64 ki = sbx_d2u(kd);
    kd -= Shift;
    const double r = xd - kd;
    sbx_u64 t = sbx_exp2_tab((sbx_u32)(ki & 31u));
    const sbx_u64 ski = ki + (sbx_u64)sign_bias;
    t += ski << 47;
    const double s = sbx_u2d(t);
    const double z = sbx_fma(r, C0, C1);
    const double r2 = r * r;
    double y = sbx_fma(r, C2, 1.0);
    y = sbx_fma(z, r2, y);
    y = y * s;
    return sbx_d2f(y);
}

// 0: not an integer, 1: odd integer, 2: even integer
SBX_MATH_FN int sbx_checkint(sbx_u32 iy) {
    const int e = (int)((iy >> 23) & 0xffu);
    if (e < 0x7f) return 0;
    if (e > 0x7f + 23) return 2;
    if (iy & ((1u << (0x7f + 23 - e)) - 1u)) return 0;
    if (iy & (1u << (0x7f + 23 - e))) return 1;
    return 2;
}
SBX_MATH_FN int sbx_zeroinfnan(sbx_u32 ix) { return 2u * ix - 1u >= 2u * 0x7f800000u - 1u; }

SBX_MATH_FN float sbx_powf(float x, float y) {
    sbx_u32 sign_bias = 0;
    sbx_u32 ix = sbx_f2u(x);
    const sbx_u32 iy = sbx_f2u(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || sbx_zeroinfnan(iy)) {
        // x is subnormal, zero, inf, nan or negative; or y is zero, inf, nan
        if (sbx_zeroinfnan(iy)) {
            if (2u * iy == 0u) return ((ix ^ 0x00400000u) & 0x7fffffffu) > 0x7fc00000u ? x + y : 1.0f;
            if (ix == 0x3f800000u) return ((iy ^ 0x00400000u) & 0x7fffffffu) > 0x7fc00000u ? x + y : 1.0f;
            if (2u * ix > 2u * 0x7f800000u || 2u * iy > 2u * 0x7f800000u) return x + y;
            if (2u * ix == 2u * 0x3f800000u) return 1.0f;
            if ((2u * ix < 2u * 0x3f800000u) == !(iy >> 31)) return 0.0f;   // |x|<1 && y==inf or |x|>1 && y==-inf
            return y * y;
        }
        if (sbx_zeroinfnan(ix)) {
            float x2 = x * x;
            if ((ix >> 31) && sbx_checkint(iy) == 1) x2 = -x2;
            return (iy >> 31) ? sbx_divf(1.0f, x2) : x2;
        }
        if (ix >> 31) {                                   // x < 0
            const int yint = sbx_checkint(iy);
            if (yint == 0) return sbx_divf(x - x, x - x);  // nan
            if (yint == 1) sign_bias = 1u << 16;          // SIGN_BIAS = 1 << (EXP2F_TABLE_BITS + 11)
            ix &= 0x7fffffffu;
        }
        if (ix < 0x00800000u) {                            // subnormal x: normalise
            ix = sbx_f2u(x * 0x1p23f);
            ix &= 0x7fffffffu;
            ix -= 23u << 23;
        }
    }
    const double logx = sbx_log2_inline(ix);
    const double ylogx = (double)y * logx;
    if (((sbx_d2u(ylogx) >> 47) & 0xffffu) >= (sbx_d2u(126.0) >> 47)) {
        // |y*log2(x)| >= 126
        if (ylogx > 0x1.fffffffd1d571p+6) return sign_bias ? sbx_u2f(0xff800000u) : sbx_u2f(0x7f800000u);
        // (0x1.fffffffa3aae2p+6, 0x1.fffffffd1d571p+6]: overflow only in directed rounding modes
        if (ylogx <= -150.0) return sign_bias ? -0.0f : 0.0f;
        if (ylogx < -149.0) return sign_bias ? -0x1p-149f : 0x1p-149f;   // __math_may_uflowf
    }
    return sbx_exp2_inline(ylogx, sign_bias);
}

// ---------------------------------------------------------------------------------------------
// acosf / atanf / atan2f / tanf: glibc 2.39 still ships the fdlibm single-precision versions
// (Sun Microsystems 1993, float conversion by Ian Lance Taylor; sysdeps/ieee754/flt-32/
// {e_acosf,s_atanf,e_atan2f,s_tanf,k_tanf}.c).  Plain fp32 arithmetic in source order, no
// contraction: the same operations round the same way on SSE and on the SM.
// ---------------------------------------------------------------------------------------------
SBX_MATH_FN float sbx_fabsf(float x) { return sbx_u2f(sbx_f2u(x) & 0x7fffffffu); }

SBX_MATH_FN float sbx_acos_rational(float z) {
    // p(z)/q(z) ~ (asin(sqrt z)/sqrt z - 1)/z on [0, 0.25]
    const float pS0 = 0x1.555556p-3f, pS1 = -0x1.4d612p-2f, pS2 = 0x1.9c155p-3f, pS3 = -0x1.48228cp-5f,
                pS4 = 0x1.9efe08p-11f, pS5 = 0x1.23de1p-15f;
    const float qS1 = -0x1.33a272p+1f, qS2 = 0x1.02ae5ap+1f, qS3 = -0x1.6066c2p-1f, qS4 = 0x1.3b8c5cp-4f;
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    return sbx_divf(p, q);
}

SBX_MATH_FN float sbx_acosf(float x) {
    const float pi = 0x1.921fb4p+1f, pio2_hi = 0x1.921fb4p+0f, pio2_lo = 0x1.4442dp-24f;
    const sbx_u32 hx = sbx_f2u(x);
    const sbx_u32 ix = hx & 0x7fffffffu;
    if (ix == 0x3f800000u) {                         // |x| == 1
        if ((int)hx > 0) return 0.0f;
        return pi + 2.0f * pio2_lo;
    } else if (ix > 0x3f800000u) {
        return sbx_divf(x - x, x - x);               // |x| > 1 or nan -> nan
    }
    if (ix < 0x3f000000u) {                          // |x| < 0.5
        if (ix <= 0x32800000u) return pio2_hi + pio2_lo;
        const float z = x * x;
        const float r = sbx_acos_rational(z);
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if ((int)hx < 0) {                        // x < -0.5
        const float z = (1.0f + x) * 0.5f;
        const float r = sbx_acos_rational(z);
        const float s = sbx_sqrtf(z);
        const float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    } else {                                         // x > 0.5
        const float z = (1.0f - x) * 0.5f;
        const float s = sbx_sqrtf(z);
        const float df = sbx_u2f(sbx_f2u(s) & 0xfffff000u);
        const float c = sbx_divf(z - df * df, s + df);
        const float r = sbx_acos_rational(z);
        const float w = r * s + c;
        return 2.0f * (df + w);
    }
}

SBX_MATH_FN float sbx_atanf(float x) {
    const float aT0 = 0x1.555556p-2f, aT1 = -0x1.99999ap-3f, aT2 = 0x1.24924ap-3f, aT3 = -0x1.c71c7p-4f,
                aT4 = 0x1.745cdcp-4f, aT5 = -0x1.3b0f2ap-4f, aT6 = 0x1.10d66ap-4f, aT7 = -0x1.dde2d6p-5f,
                aT8 = 0x1.97b4b2p-5f, aT9 = -0x1.2b4442p-5f, aT10 = 0x1.0ad3aep-6f;
    const sbx_u32 hx = sbx_f2u(x);
    const sbx_u32 ix = hx & 0x7fffffffu;
    float hi = 0.0f, lo = 0.0f;
    int id;
    if (ix >= 0x4c000000u) {                         // |x| >= 2^25
        if (ix > 0x7f800000u) return x + x;          // nan
        if ((int)hx > 0) return 0x1.921fb4p+0f + 0x1.4442dp-24f;
        return -0x1.921fb4p+0f - 0x1.4442dp-24f;
    }
    if (ix < 0x3ee00000u) {                          // |x| < 0.4375
        if (ix < 0x31000000u) return x;              // |x| < 2^-29
        id = -1;
    } else {
        x = sbx_fabsf(x);
        if (ix < 0x3f980000u) {                      // |x| < 1.1875
            if (ix < 0x3f300000u) {                  // 7/16 <= |x| < 11/16
                id = 0; hi = 0x1.dac67p-2f; lo = 0x1.586ed2p-28f;
                x = sbx_divf(2.0f * x - 1.0f, 2.0f + x);
            } else {                                 // 11/16 <= |x| < 19/16
                id = 1; hi = 0x1.921fb4p-1f; lo = 0x1.4442dp-25f;
                x = sbx_divf(x - 1.0f, x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000u) {                  // |x| < 2.4375
                id = 2; hi = 0x1.f730bcp-1f; lo = 0x1.281f68p-25f;
                x = sbx_divf(x - 1.5f, 1.0f + 1.5f * x);
            } else {                                 // 2.4375 <= |x| < 2^25
                id = 3; hi = 0x1.921fb4p+0f; lo = 0x1.4442dp-24f;
                x = sbx_divf(-1.0f, x);
            }
        }
    }
    float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    z = hi - ((x * (s1 + s2) - lo) - x);
    return ((int)hx < 0) ? -z : z;
}

SBX_MATH_FN float sbx_atan2f(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = 0x1.921fb6p-1f, pi_o_2 = 0x1.921fb6p+0f, pi = 0x1.921fb6p+1f,
                pi_lo = -0x1.777a5cp-24f;
    const sbx_u32 hx = sbx_f2u(x), hy = sbx_f2u(y);
    const sbx_u32 ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;     // nan
    if (hx == 0x3f800000u) return sbx_atanf(y);                 // x == 1
    const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);   // 2*sign(x) + sign(y)
    if (iy == 0u) {
        switch (m) {
            case 0: case 1: return y;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (ix == 0u) return ((int)hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
                case 0: return pi_o_4 + tiny;
                case 1: return -pi_o_4 - tiny;
                case 2: return 3.0f * pi_o_4 + tiny;
                default: return -3.0f * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return pi + tiny;
                default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000u) return ((int)hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = ((int)iy - (int)ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;                      // |y/x| > 2^60
    else if ((int)hx < 0 && k < -60) z = 0.0f;                  // |y|/x < -2^60
    else z = sbx_atanf(sbx_fabsf(sbx_divf(y, x)));
    switch (m) {
        case 0: return z;
        case 1: return sbx_u2f(sbx_f2u(z) ^ 0x80000000u);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

// tan on [-pi/4, pi/4] with a tail y; iy = 1 -> tan, -1 -> -1/tan  (k_tanf.c)
SBX_MATH_FN float sbx_kernel_tanf(float x, float y, int iy) {
    const float pio4 = 0x1.921fb4p-1f, pio4lo = 0x1.4442dp-25f;
    const float T0 = 0x1.555556p-2f, T1 = 0x1.111112p-3f, T2 = 0x1.ba1ba2p-5f, T3 = 0x1.664f48p-6f,
                T4 = 0x1.226e3ep-7f, T5 = 0x1.d6d22cp-9f, T6 = 0x1.7dbc9p-10f, T7 = 0x1.344d9p-11f,
                T8 = 0x1.026f72p-12f, T9 = 0x1.47e88ap-14f, T10 = 0x1.2b80f4p-14f, T11 = -0x1.375cbep-16f,
                T12 = 0x1.b2a708p-16f;
    const sbx_u32 hx = sbx_f2u(x);
    const sbx_u32 ix = hx & 0x7fffffffu;
    float z, r, v, w, s;
    if (ix < 0x39000000u) {                          // |x| < 2^-13
        if ((int)x == 0) {
            if ((ix | (sbx_u32)(iy + 1)) == 0u) return sbx_divf(1.0f, sbx_fabsf(x));
            else if (iy == 1) return x;
            else return sbx_divf(-1.0f, x);
        }
    }
    if (ix >= 0x3f2ca140u) {                         // |x| >= 0.6744
        if ((int)hx < 0) { x = -x; y = -y; }
        z = pio4 - x;
        w = pio4lo - y;
        x = z + w;
        y = 0.0f;
        if (sbx_fabsf(x) < 0x1p-13f)
            return (float)((1 - (int)((hx >> 30) & 2u)) * iy) * (1.0f - (float)(2 * iy) * x);
    }
    z = x * x;
    w = z * z;
    r = T1 + w * (T3 + w * (T5 + w * (T7 + w * (T9 + w * T11))));
    v = z * (T2 + w * (T4 + w * (T6 + w * (T8 + w * (T10 + w * T12)))));
    s = z * x;
    r = y + z * (s * (r + v) + y);
    r += T0 * s;
    w = x + r;
    if (ix >= 0x3f2ca140u) {
        v = (float)iy;
        return (float)(1 - (int)((hx >> 30) & 2u)) * (v - 2.0f * (x - (sbx_divf(w * w, w + v) - r)));
    }
    if (iy == 1) return w;
    {   // -1/(x+r) with the error of the quotient corrected
        float a, t;
        z = sbx_u2f(sbx_f2u(w) & 0xfffff000u);
        v = r - (z - x);
        t = a = sbx_divf(-1.0f, w);
        t = sbx_u2f(sbx_f2u(t) & 0xfffff000u);
        s = 1.0f + t * z;
        return t + a * (s + t * v);
    }
}

SBX_MATH_FN float sbx_tanf(float x) {
    const sbx_u32 ix = sbx_f2u(x) & 0x7fffffffu;
    if (ix <= 0x3f490fdau) return sbx_kernel_tanf(x, 0.0f, 1);   // |x| <= pi/4
    if (ix >= 0x7f800000u) return x - x;                         // inf / nan
    // reduce in double like sinf (but with separate multiply and subtract: this entry point has
    // no FMA variant in glibc), then split the remainder into a float head and tail
    double xd = (double)x;
    int n;
    if (((sbx_f2u(x) >> 20) & 0x7ffu) < 0x42fu) {
        const double hpi_inv = 0x1.45f306dc9c883p+23, hpi = 0x1.921fb54442d18p+0;
        const double r = xd * hpi_inv;
        n = ((int)r + 0x800000) >> 24;
        xd = xd - (double)n * hpi;
    } else {
        xd = sbx_reduce_large(sbx_f2u(x), &n);
        if ((int)sbx_f2u(x) < 0) xd = -xd;
    }
    const float y0 = sbx_d2f(xd);
    const float y1 = sbx_d2f(xd - (double)y0);
    return sbx_kernel_tanf(y0, y1, 1 - ((n & 1) << 1));
}

// ---------------------------------------------------------------------------------------------
// IEEE helpers shared by the vector layer
// ---------------------------------------------------------------------------------------------
#if defined(SBX_DEVICE_CODE)
SBX_MATH_FN float sbx_floorf(float a) { return floorf(a); }
SBX_MATH_FN float sbx_fminf(float a, float b) { return fminf(a, b); }
SBX_MATH_FN float sbx_fmaxf(float a, float b) { return fmaxf(a, b); }
#else
SBX_MATH_FN float sbx_floorf(float a) { return floorf(a); }
SBX_MATH_FN float sbx_fminf(float a, float b) { return fminf(a, b); }
SBX_MATH_FN float sbx_fmaxf(float a, float b) { return fmaxf(a, b); }
#endif

#endif  // SBX_MATH_H_
