// gmath.cuh — bit-exact device restatements of the libm functions the
// reference's arithmetic goes through on linux-gnu (rustc lowers f32::powf /
// sin / cos / exp to glibc powf / sinf / cosf / expf).
//
// Why this exists (DESIGN.md §"Numerics"): the reference forms oscillator
// arguments as f32(idx*freq*2pi/sr) with freq derived from powf() envelopes
// (src/envelope.rs:21-26, src/gen/oscillator.rs:42-46).  A 1-ulp difference in
// powf is multiplied by the phase argument (hundreds to 1e5 rad), so CUDA's
// own powf (<=2 ulp, different rounding) breaks the 1e-5 parity bar.  These
// routines follow glibc 2.39's published double-precision algorithms
// (sysdeps/ieee754/flt-32/{e_powf,e_expf,s_sinf,s_cosf}.c, ARM optimized
// routines lineage) with the x86-64 FMA ifunc variant's contraction pattern,
// so results are bit-identical to the host libm the reference links.
// Table constants are those of the platform libm (verified in tests/).
//
// Provenance / licence: the algorithms and table constants restate third-party code, not the reference — glibc 2.39
// (GNU LGPL-2.1-or-later) `sysdeps/ieee754/flt-32/`: e_powf.c, e_expf.c, s_sinf.c, s_cosf.c, sincosf.h and their data tables
// (derived from ARM Optimized Routines, MIT / Apache-2.0-with-LLVM-exception, (c) Arm Ltd.), and the fdlibm single-precision
// routines glibc still carries for tanf / tanhf / expm1f ((c) 1993 Sun Microsystems, "freely granted" permission notice).
// Nothing here is copied from gooey-audio/libgooey.
//
// The same header compiles for the host (g++, tests only) with GM_HD empty.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define GM_HD __host__ __device__ __forceinline__
#define GM_CONST __constant__
#else
#define GM_HD static inline
#define GM_CONST static const
#endif

namespace gm {

#ifdef __CUDA_ARCH__
GM_HD uint32_t asuint(float f) { return __float_as_uint(f); }
GM_HD float asfloat(uint32_t u) { return __uint_as_float(u); }
GM_HD uint64_t asuint64(double d) { return (uint64_t)__double_as_longlong(d); }
GM_HD double asdouble(uint64_t u) { return __longlong_as_double((long long)u); }
GM_HD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
GM_HD uint32_t asuint(float f) { uint32_t u; __builtin_memcpy(&u, &f, 4); return u; }
GM_HD float asfloat(uint32_t u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
GM_HD uint64_t asuint64(double d) { uint64_t u; __builtin_memcpy(&u, &d, 8); return u; }
GM_HD double asdouble(uint64_t u) { double d; __builtin_memcpy(&d, &u, 8); return d; }
GM_HD double dfma(double a, double b, double c) { return __builtin_fma(a, b, c); }
#endif

// 2^(i/32) with the exponent contribution of i/32 removed (exp2f_data.tab).
#define GM_INIT_K_EXP2_TAB { \
  0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
  0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
  0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
  0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
  0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
  0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
  0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
  0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull, \
}
#ifdef __CUDACC__
__constant__ uint64_t d_k_exp2_tab[32] = GM_INIT_K_EXP2_TAB;
#endif
static const uint64_t h_k_exp2_tab[32] = GM_INIT_K_EXP2_TAB;


// powf_log2_data: {1/c, log2(c)} for the 16 sub-intervals of [0x1.66p-1, 0x1.66p0).
#define GM_INIT_K_PLOG_TAB { \
  {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2}, \
  {0x1.49539f0f010bp+0, -0x1.7418b0a1fb77bp-2},  {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2}, \
  {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8eap+0, -0x1.97c1d1b3b7afp-3}, \
  {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4}, \
  {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1p+0, 0x0p+0}, \
  {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},  {0x1.ca4b31f026aap-1, 0x1.476a9543891bap-3}, \
  {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2}, \
  {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},  {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}, \
}
#ifdef __CUDACC__
__constant__ double d_k_plog_tab[16][2] = GM_INIT_K_PLOG_TAB;
#endif
static const double h_k_plog_tab[16][2] = GM_INIT_K_PLOG_TAB;


#define GM_INIT_K_INV_PIO4 { \
  0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, \
  0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, \
  0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041, \
}
#ifdef __CUDACC__
__constant__ uint32_t d_k_inv_pio4[24] = GM_INIT_K_INV_PIO4;
#endif
static const uint32_t h_k_inv_pio4[24] = GM_INIT_K_INV_PIO4;



#ifdef __CUDA_ARCH__
#define GM_TAB(name) d_##name
#else
#define GM_TAB(name) h_##name
#endif
// ---------------------------------------------------------------- expf ----
// glibc e_expf.c: exp(x) = 2^(k/32) * 2^(r/32), double arithmetic, one rounding.
GM_HD float expf_core(double xd) {
  const double InvLn2N = 0x1.71547652b82fep+5, SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-20, C1 = 0x1.ebfce50fac4f3p-13, C2 = 0x1.62e42ff0c52d6p-6;
  double z = InvLn2N * xd;
  double kd = z + SHIFT;
  uint64_t ki = asuint64(kd);
  kd -= SHIFT;
  double r = z - kd;
  uint64_t t = GM_TAB(k_exp2_tab)[ki & 31];
  t += ki << (52 - 5);
  double s = asdouble(t);
  z = dfma(C0, r, C1);
  double r2 = r * r;
  double y = dfma(C2, r, 1.0);
  y = dfma(z, r2, y);
  y = y * s;
  return (float)y;
}

GM_HD float g_expf(float x) {
  uint32_t abstop = (asuint(x) >> 20) & 0x7ff;
  if (abstop >= 0x42b) {  // |x| >= 88 or NaN/Inf
    if (asuint(x) == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8) return x + x;
    if (x > 0x1.62e42ep6f) return INFINITY;
    if (x < -0x1.9fe368p6f) return 0.0f;
    // -103.97 <= x <= -88 (subnormal results) and 88 <= x <= 88.72: same core.
  }
  return expf_core((double)x);
}

// ---------------------------------------------------------------- powf ----
GM_HD double plog2_inline(uint32_t ix) {
  const double A0 = 0x1.27616c9496e0bp-2, A1 = -0x1.71969a075c67ap-2, A2 = 0x1.ec70a6ca7baddp-2,
               A3 = -0x1.7154748bef6c8p-1, A4 = 0x1.71547652ab82bp+0;
  uint32_t tmp = ix - 0x3f330000u;
  int i = (tmp >> (23 - 4)) & 15;
  uint32_t top = tmp & 0xff800000u;
  uint32_t iz = ix - top;
  int k = (int32_t)top >> 23;
  double invc = GM_TAB(k_plog_tab)[i][0], logc = GM_TAB(k_plog_tab)[i][1];
  double z = (double)asfloat(iz);
  double r = dfma(z, invc, -1.0);
  double y0 = logc + (double)k;
  double r2 = r * r;
  double y = dfma(A0, r, A1);
  double p = dfma(A2, r, A3);
  double r4 = r2 * r2;
  double q = dfma(A4, r, y0);
  q = dfma(p, r2, q);
  y = dfma(y, r4, q);
  return y;
}

GM_HD float pexp2_inline(double xd, uint32_t sign_bias) {
  const double SHIFT = 0x1.8p+47;  // 0x1.8p52 / 32
  const double C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
  double kd = xd + SHIFT;
  uint64_t ki = asuint64(kd);
  kd -= SHIFT;
  double r = xd - kd;
  uint64_t t = GM_TAB(k_exp2_tab)[ki & 31];
  uint64_t ski = ki + sign_bias;
  t += ski << (52 - 5);
  double s = asdouble(t);
  double z = dfma(C0, r, C1);
  double r2 = r * r;
  double y = dfma(C2, r, 1.0);
  y = dfma(z, r2, y);
  y = y * s;
  return (float)y;
}

GM_HD int pow_checkint(uint32_t iy) {  // 0: not int, 1: odd, 2: even
  int e = iy >> 23 & 0xff;
  if (e < 0x7f) return 0;
  if (e > 0x7f + 23) return 2;
  if (iy & ((1u << (0x7f + 23 - e)) - 1)) return 0;
  if (iy & (1u << (0x7f + 23 - e))) return 1;
  return 2;
}
GM_HD bool pow_zeroinfnan(uint32_t ix) { return 2 * ix - 1 >= 2u * 0x7f800000u - 1; }

GM_HD float g_powf(float x, float y) {
  uint32_t sign_bias = 0;
  uint32_t ix = asuint(x), iy = asuint(y);
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || pow_zeroinfnan(iy)) {
    if (pow_zeroinfnan(iy)) {
      if (2 * iy == 0) return 1.0f;
      if (ix == 0x3f800000u) return 1.0f;
      if (2 * ix > 2u * 0x7f800000u || 2 * iy > 2u * 0x7f800000u) return x + y;
      if (2 * ix == 2 * 0x3f800000u) return 1.0f;
      if ((2 * ix < 2 * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;
      return y * y;
    }
    if (pow_zeroinfnan(ix)) {
      float x2 = x * x;
      if ((ix & 0x80000000u) && pow_checkint(iy) == 1) x2 = -x2;
      return (iy & 0x80000000u) ? 1.0f / x2 : x2;
    }
    if (ix & 0x80000000u) {
      int yint = pow_checkint(iy);
      if (yint == 0) return NAN;
      if (yint == 1) sign_bias = 1u << (5 + 11);
      ix &= 0x7fffffffu;
    }
    if (ix < 0x00800000u) {
      ix = asuint(x * 0x1p23f);
      ix &= 0x7fffffffu;
      ix -= 23u << 23;
    }
  }
  double logx = plog2_inline(ix);
  double ylogx = (double)y * logx;
  if ((asuint64(ylogx) >> 47 & 0xffff) >= (asuint64(126.0) >> 47)) {
    if (ylogx > 0x1.fffffffd1d571p+6) return sign_bias ? -INFINITY : INFINITY;
    if (ylogx <= -150.0) return sign_bias ? -0.0f : 0.0f;
  }
  return pexp2_inline(ylogx, sign_bias);
}

// ------------------------------------------------------------ sinf/cosf ----
struct SinCosTab { double c0, c1, c2, c3, c4, s1, s2, s3; };

GM_HD double sc_reduce_fast(double x, int* np) {
  const double hpi_inv = 0x1.45f306dc9c883p+23, hpi = 0x1.921fb54442d18p+0;
  double r = x * hpi_inv;
  int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return dfma(-(double)n, hpi, x);
}

GM_HD double sc_reduce_large(uint32_t xi, int* np) {
  const uint32_t* arr = &GM_TAB(k_inv_pio4)[(xi >> 26) & 15];
  int shift = (xi >> 23) & 7;
  uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;
  res0 = (uint32_t)(xi * arr[0]);
  res1 = (uint64_t)xi * arr[4];
  res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ull << 61)) >> 62;
  res0 -= n << 62;
  double x = (double)(int64_t)res0;
  *np = (int)n;
  return x * 0x1.921fb54442d18p-62;
}

// neg selects the negated-cosine table (quadrants 2,3); n&1 selects cos vs sin poly.
GM_HD float sc_poly(double x, double x2, bool neg, int n) {
  if ((n & 1) == 0) {
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    double x3 = x * x2;
    double t1 = dfma(x2, s3, s2);
    double x7 = x3 * x2;
    double s = dfma(x3, s1, x);
    return (float)dfma(x7, t1, s);
  } else {
    double c0 = 1.0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5,
           c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
    if (neg) { c0 = -c0; c1 = -c1; c2 = -c2; c3 = -c3; c4 = -c4; }
    double x4 = x2 * x2;
    double t2 = dfma(x2, c4, c3);
    double t1 = dfma(x2, c1, c0);
    double x6 = x4 * x2;
    double c = dfma(x4, c2, t1);
    return (float)dfma(x6, t2, c);
  }
}

GM_HD uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }

GM_HD float g_sinf(float y) {
  double x = (double)y;
  int n;
  uint32_t at = abstop12(y);
  if (at < 0x3f4) {           // |y| < pi/4  (abstop12(pio4) = 0x3f4)
    double s = x * x;
    if (at < 0x398) return y;  // |y| < 2^-12
    return sc_poly(x, s, false, 0);
  } else if (at < 0x42f) {    // |y| < 120
    x = sc_reduce_fast(x, &n);
    double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return sc_poly(x * s, x * x, (n & 2) != 0, n);
  } else if (at < 0x7f8) {
    uint32_t xi = asuint(y);
    int sign = xi >> 31;
    x = sc_reduce_large(xi, &n);
    int m = n + sign;
    double s = ((m & 3) == 1 || (m & 3) == 2) ? -1.0 : 1.0;
    return sc_poly(x * s, x * x, (m & 2) != 0, n);
  }
  return NAN;
}

GM_HD float g_cosf(float y) {
  double x = (double)y;
  int n;
  uint32_t at = abstop12(y);
  if (at < 0x3f4) {
    double x2 = x * x;
    if (at < 0x398) return 1.0f;
    return sc_poly(x, x2, false, 1);
  } else if (at < 0x42f) {
    x = sc_reduce_fast(x, &n);
    double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return sc_poly(x * s, x * x, (n & 2) != 0, n ^ 1);
  } else if (at < 0x7f8) {
    uint32_t xi = asuint(y);
    int sign = xi >> 31;
    x = sc_reduce_large(xi, &n);
    int m = n + sign;
    double s = ((m & 3) == 1 || (m & 3) == 2) ? -1.0 : 1.0;
    return sc_poly(x * s, x * x, (m & 2) != 0, n ^ 1);
  }
  return NAN;
}


// ------------------------------------------------- expm1f / tanhf / tanf ----
// glibc 2.39 still uses the fdlibm single-precision routines for these
// (sysdeps/ieee754/flt-32/{s_expm1f,s_tanhf,s_tanf,k_tanf,e_rem_pio2f}.c): plain f32
// arithmetic, no FMA variant.  Restated operation-for-operation; every product/sum
// below must stay unfused (-fmad=false on the device, -ffp-contract=off on the host).
GM_HD float g_expm1f(float x) {
  const float one = 1.0f, huge = 1.0e+30f, tiny = 1.0e-30f;
  const float o_threshold = 8.8721679688e+01f, ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f, invln2 = 1.4426950216e+00f;
  const float Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f, Q3 = -7.9365076090e-05f, Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
  float y, hi, lo, c = 0.0f, t, e, hxs, hfx, r1;
  int32_t k;
  uint32_t hx = asuint(x);
  uint32_t xsb = hx & 0x80000000u;
  hx &= 0x7fffffffu;
  if (hx >= 0x4195b844u) {            // |x| >= 27 ln2
    if (hx >= 0x42b17218u) {          // |x| >= 88.72
      if (hx > 0x7f800000u) return x + x;
      if (hx == 0x7f800000u) return xsb == 0 ? x : -1.0f;
      if (x > o_threshold) return huge * huge;
    }
    if (xsb != 0) return tiny - one;
  }
  if (hx > 0x3eb17218u) {             // |x| > 0.5 ln2
    if (hx < 0x3F851592u) {           // |x| < 1.5 ln2
      if (xsb == 0) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
      else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
    } else {
      k = (int32_t)(invln2 * x + (xsb == 0 ? 0.5f : -0.5f));
      t = (float)k;
      hi = x - t * ln2_hi;
      lo = t * ln2_lo;
    }
    x = hi - lo;
    c = (hi - x) - lo;
  } else if (hx < 0x33000000u) {      // |x| < 2^-25
    t = huge + x;
    return x - (t - (huge + x));
  } else k = 0;
  hfx = 0.5f * x;
  hxs = x * hfx;
  r1 = one + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
  t = 3.0f - r1 * hfx;
  e = hxs * ((r1 - t) / (6.0f - x * t));
  if (k == 0) return x - (x * e - hxs);
  e = (x * (e - c) - c);
  e -= hxs;
  if (k == -1) return 0.5f * (x - e) - 0.5f;
  if (k == 1) {
    if (x < -0.25f) return -2.0f * (e - (x + 0.5f));
    return one + 2.0f * (x - e);
  }
  if (k <= -2 || k > 56) {
    y = one - (e - x);
    y = asfloat(asuint(y) + ((uint32_t)k << 23));
    return y - one;
  }
  if (k < 23) {
    t = asfloat(0x3f800000u - (0x1000000u >> k));
    y = t - (e - x);
    y = asfloat(asuint(y) + ((uint32_t)k << 23));
  } else {
    t = asfloat((uint32_t)(0x7f - k) << 23);
    y = x - (e + t);
    y += one;
    y = asfloat(asuint(y) + ((uint32_t)k << 23));
  }
  return y;
}

GM_HD float g_tanhf(float x) {
  const float one = 1.0f, two = 2.0f, tiny = 1.0e-30f;
  float t, z;
  int32_t jx = (int32_t)asuint(x);
  int32_t ix = jx & 0x7fffffff;
  if (ix >= 0x7f800000) return jx >= 0 ? one / x + one : one / x - one;
  if (ix < 0x41b00000) {              // |x| < 22
    if (ix == 0) return x;
    if (ix < 0x24000000) return x * (one + x);
    if (ix >= 0x3f800000) { t = g_expm1f(two * fabsf(x)); z = one - two / (t + two); }
    else { t = g_expm1f(-two * fabsf(x)); z = -t / (t + two); }
  } else z = one - tiny;
  return jx >= 0 ? z : -z;
}

GM_HD float g_kernel_tanf(float x, float y, int iy) {
  const float one = 1.0f, pio4 = 7.8539812565e-01f, pio4lo = 3.7748947079e-08f;
  const float T0 = 3.3333334327e-01f, T1 = 1.3333334029e-01f, T2 = 5.3968254477e-02f, T3 = 2.1869488060e-02f,
              T4 = 8.8632395491e-03f, T5 = 3.5920790397e-03f, T6 = 1.4562094584e-03f, T7 = 5.8804126456e-04f,
              T8 = 2.4646313977e-04f, T9 = 7.8179444245e-05f, T10 = 7.1407252108e-05f, T11 = -1.8558637748e-05f,
              T12 = 2.5907305826e-05f;
  float z, r, v, w, s;
  int32_t hx = (int32_t)asuint(x);
  int32_t ix = hx & 0x7fffffff;
  if (ix < 0x39000000) {              // |x| < 2^-13
    if ((int)x == 0) {
      if ((ix | (iy + 1)) == 0) return one / fabsf(x);
      else if (iy == 1) return x;
      else return -one / x;
    }
  }
  if (ix >= 0x3f2ca140) {             // |x| >= 0.6744
    if (hx < 0) { x = -x; y = -y; }
    z = pio4 - x;
    w = pio4lo - y;
    x = z + w; y = 0.0f;
    if (fabsf(x) < 0x1p-13f) return (float)((1 - ((hx >> 30) & 2)) * iy) * (1.0f - (float)(2 * iy) * x);
  }
  z = x * x;
  w = z * z;
  r = T1 + w * (T3 + w * (T5 + w * (T7 + w * (T9 + w * T11))));
  v = z * (T2 + w * (T4 + w * (T6 + w * (T8 + w * (T10 + w * T12)))));
  s = z * x;
  r = y + z * (s * (r + v) + y);
  r += T0 * s;
  w = x + r;
  if (ix >= 0x3f2ca140) {
    v = (float)iy;
    return (float)(1 - ((hx >> 30) & 2)) * (v - 2.0f * (x - (w * w / (w + v) - r)));
  }
  if (iy == 1) return w;
  float a, t;
  z = asfloat(asuint(w) & 0xfffff000u);
  v = r - (z - x);
  t = a = -1.0f / w;
  t = asfloat(asuint(t) & 0xfffff000u);
  s = 1.0f + t * z;
  return t + a * (s + t * v);
}

// tanf for |x| <= 3pi/4 through the fdlibm small-argument reduction (e_rem_pio2f.c first branch); larger
// arguments never occur on this path (filter prewarp arguments are <= 0.45*pi) and fall back to NaN-safe tan.
GM_HD float g_tanf(float x) {
  const float pio2_1 = 1.5707855225e+00f, pio2_1t = 1.0804334124e-05f, pio2_2 = 1.0804273188e-05f, pio2_2t = 6.0770999344e-11f;
  int32_t hx = (int32_t)asuint(x);
  int32_t ix = hx & 0x7fffffff;
  if (ix <= 0x3f490fda) return g_kernel_tanf(x, 0.0f, 1);
  if (ix >= 0x7f800000) return x - x;
  if (ix <= 0x4016cbe3) {             // |x| ~< 3pi/4
    float z, y0, y1;
    if (hx > 0) {
      z = x - pio2_1;
      if ((ix & 0xfffffff0) != 0x3fc90fd0) { y0 = z - pio2_1t; y1 = (z - y0) - pio2_1t; }
      else { z -= pio2_2; y0 = z - pio2_2t; y1 = (z - y0) - pio2_2t; }
      return g_kernel_tanf(y0, y1, -1);
    } else {
      z = x + pio2_1;
      if ((ix & 0xfffffff0) != 0x3fc90fd0) { y0 = z + pio2_1t; y1 = (z - y0) + pio2_1t; }
      else { z += pio2_2; y0 = z + pio2_2t; y1 = (z - y0) + pio2_2t; }
      return g_kernel_tanf(y0, y1, -1);
    }
  }
  return tanf(x);
}


// sin(y) to ~1 ulp (1.2e-7 absolute, measured against f64 over [-pi/2, pi/2]) for |y| < 2^50, ~17 issue slots instead of the
// ~100 of the glibc port above: half-period reduction in f64 (k = rint(y / pi) by the 1.5 * 2^52 shift, whose low word is k;
// r = y - k pi, exact to 1e-10 for |y| < 3e7) and ONE odd minimax polynomial of degree 11 on [-pi/2, pi/2], sign from the
// parity of k.  NOT bit-identical to glibc (the port is): used only where DESIGN.md "Where approximation is allowed" applies
// and the parity tests keep their margin.
GM_HD float g_sinf_fast(float y) {
  const double yd = (double)y;
  const double shift = 6755399441055744.0;               // 1.5 * 2^52
  const double t = fma(yd, 0.31830988618379067154, shift);
  const double kd = t - shift;
  const double rd = fma(-kd, 3.14159265358979323846, yd);
#ifdef __CUDA_ARCH__
  const unsigned k = (unsigned)__double2loint(t);
#else
  const unsigned k = (unsigned)(long long)kd;
#endif
  const float r = (float)rd, z = r * r;
  const float p = fmaf(fmaf(fmaf(fmaf(-2.3791757897129173e-08f, z, 2.751906322373543e-06f), z, -0.00019840722961816937f), z, 0.008333330042660236f), z, -0.1666666716337204f);
  const float v = fmaf(r * z, p, r);
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(v) ^ (k << 31));
#else
  return (k & 1u) ? -v : v;
#endif
}
// a / b, correctly rounded, for a fixed divisor whose reciprocal y = RN(1 / b) the caller hoists (Markstein: q = RN(a y),
// r = a - b q exactly, q' = RN(q + r y)); bit-identical to the IEEE quotient away from overflow / underflow of the quotient
// (checked on 7.2e8 random numerators x 12 sample rates; tests/test_emu_cpu.py and tests/test_gmath_gpu.py hold smaller runs on the host and on the device).  3 instructions
// instead of the ~9 + slow-path branch of div.rn.f32.
GM_HD float g_div_by(float a, float b, float y) {
  const float q = a * y;
  const float r = fmaf(-q, b, a);
  return fmaf(r, y, q);
}
}  // namespace gm
