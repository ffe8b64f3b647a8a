/* crt_powf5.h -- x^5 exactly as glibc 2.39's powf(x, 5.0f) returns it, for the Fresnel term of
 * RayTracer::calculateRefraction (`std::powf(1.0f - cosineAlpha, 5)`, RayTracer.cpp:407).
 *
 * glibc's powf (sysdeps/ieee754/flt-32/e_powf.c; the ARM optimized-routines algorithm) is not correctly rounded:
 * it evaluates exp2(y * log2(x)) in binary64 with a 16-entry log2 table + degree-5 polynomial and a 32-entry exp2
 * table + cubic, and rounds once to binary32.  A float-only or plain `x*x*x*x*x` device version differs from it in the
 * last bit for ~0.07 % / 41 % of inputs, which breaks bit-identical float RGB on refractive
 * pixels.  This header restates that published algorithm for the fixed exponent 5; the table constants are the
 * mathematical constants of the algorithm (invc ~ 1/c, logc = log2(c); 2^(i/32)), as published with it.
 * tests/test_powf5.py checks bit-equality against the host libm over 2e7 inputs (incl. subnormals, negatives, 0).
 *
 * The x86-64 glibc dispatches powf to an FMA build on FMA-capable CPUs; CRT_POWF5_FMA(a,b,c) is therefore a fused
 * multiply-add where the C source has `a * b + c`.  (FMA vs non-FMA changes the binary64 value by <= 1 ulp(double), i.e.
 * flips the final float rounding with probability ~2^-28 per call.)
 */
#ifndef CRT_POWF5_H
#define CRT_POWF5_H
#include <stdint.h>

#ifdef __CUDA_ARCH__
#define CRT_P5_HD __device__ __forceinline__
#define CRT_POWF5_FMA(a, b, c) __fma_rn((a), (b), (c))
#define CRT_P5_MUL(a, b) __dmul_rn((a), (b))
#define CRT_P5_ADD(a, b) __dadd_rn((a), (b))
#define CRT_P5_SUB(a, b) __dsub_rn((a), (b))
#define CRT_P5_U2D(u) __longlong_as_double((long long)(u))
#define CRT_P5_D2U(d) ((uint64_t)__double_as_longlong(d))
#define CRT_P5_F2U(f) __float_as_uint(f)
#define CRT_P5_U2F(u) __uint_as_float(u)
#define CRT_P5_D2F(d) __double2float_rn(d)
#define CRT_P5_CONST __device__ static const
#else
#include <math.h>
#include <string.h>
#define CRT_P5_HD static inline
#ifndef CRT_POWF5_FMA
#define CRT_POWF5_FMA(a, b, c) fma((a), (b), (c))
#endif
#define CRT_P5_MUL(a, b) ((a) * (b))
#define CRT_P5_ADD(a, b) ((a) + (b))
#define CRT_P5_SUB(a, b) ((a) - (b))
static inline double crt_p5_u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t crt_p5_d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline uint32_t crt_p5_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float crt_p5_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define CRT_P5_U2D(u) crt_p5_u2d(u)
#define CRT_P5_D2U(d) crt_p5_d2u(d)
#define CRT_P5_F2U(f) crt_p5_f2u(f)
#define CRT_P5_U2F(u) crt_p5_u2f(u)
#define CRT_P5_D2F(d) ((float)(d))
#define CRT_P5_CONST static const
#endif

/* log2 table: {invc, logc}, c near the centre of [2^-1/2 ... 2^1/2) split in 16 */
CRT_P5_CONST double CRT_P5_LOG2_TAB[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2}, {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},  {0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},  {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2},
};
CRT_P5_CONST double CRT_P5_LOG2_POLY[5] = {0x1.27616c9496e0bp-2, -0x1.71969a075c67ap-2, 0x1.ec70a6ca7baddp-2,
                                          -0x1.7154748bef6c8p-1, 0x1.71547652ab82bp+0};
/* exp2 table: bits(2^(i/32)) - (i << 47) */
CRT_P5_CONST uint64_t CRT_P5_EXP2_TAB[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull,
    0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull,
    0x3feedea64c123422ull, 0x3feece086061892dull, 0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull,
    0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, 0x3feee89f995ad3adull,
    0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};
CRT_P5_CONST double CRT_P5_EXP2_POLY[3] = {0x1.c6af84b912394p-5, 0x1.ebfce50fac4f3p-3, 0x1.62e42ff0c52d6p-1};

CRT_P5_HD float crt_powf5(float x) {
  uint32_t ix = CRT_P5_F2U(x);
  uint64_t sign_bias = 0;
  if (x != x) return x + x;                      /* NaN */
  if ((ix << 1) == 0) return x;                  /* (+-0)^5 = +-0 (odd exponent keeps the sign) */
  if ((ix & 0x7fffffffu) == 0x7f800000u) return x; /* (+-inf)^5 */
  if (ix & 0x80000000u) {                        /* negative base, odd integer exponent */
    sign_bias = 1ull << (5 + 11);
    ix &= 0x7fffffffu;
  }
  if (ix < 0x00800000u) {                        /* subnormal: normalise */
    ix = CRT_P5_F2U(CRT_P5_U2F(ix) * 0x1p23f);
    ix &= 0x7fffffffu;
    ix -= 23u << 23;
  }
  /* log2_inline */
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> (23 - 4)) % 16u);
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ix - top;
  const int k = (int32_t)top >> 23;
  const double invc = CRT_P5_LOG2_TAB[i][0], logc = CRT_P5_LOG2_TAB[i][1];
  const double z = (double)CRT_P5_U2F(iz);
  const double r = CRT_POWF5_FMA(z, invc, -1.0);
  const double y0 = CRT_P5_ADD(logc, (double)k);
  const double r2 = CRT_P5_MUL(r, r);
  double y = CRT_POWF5_FMA(CRT_P5_LOG2_POLY[0], r, CRT_P5_LOG2_POLY[1]);
  const double p = CRT_POWF5_FMA(CRT_P5_LOG2_POLY[2], r, CRT_P5_LOG2_POLY[3]);
  const double r4 = CRT_P5_MUL(r2, r2);
  double q = CRT_POWF5_FMA(CRT_P5_LOG2_POLY[4], r, y0);
  q = CRT_POWF5_FMA(p, r2, q);
  y = CRT_POWF5_FMA(y, r4, q);
  const double ylogx = CRT_P5_MUL(5.0, y);
  /* range: |ylogx| >= 126 */
  if (((CRT_P5_D2U(ylogx) >> 47) & 0xffff) >= (CRT_P5_D2U(126.0) >> 47)) {
    if (ylogx > 0x1.fffffffd1d571p+6) return sign_bias ? -__builtin_inff() : __builtin_inff();
    if (ylogx <= -150.0) return sign_bias ? -0.0f : 0.0f;
  }
  /* exp2_inline */
  const double shift = 0x1.8p+52 / 32.0;
  double kd = CRT_P5_ADD(ylogx, shift);
  const uint64_t ki = CRT_P5_D2U(kd);
  kd = CRT_P5_SUB(kd, shift);
  const double rr = CRT_P5_SUB(ylogx, kd);
  uint64_t t = CRT_P5_EXP2_TAB[ki % 32];
  const uint64_t ski = ki + sign_bias;
  t += ski << (52 - 5);
  const double s = CRT_P5_U2D(t);
  const double zz = CRT_POWF5_FMA(CRT_P5_EXP2_POLY[0], rr, CRT_P5_EXP2_POLY[1]);
  const double rr2 = CRT_P5_MUL(rr, rr);
  double yy = CRT_POWF5_FMA(CRT_P5_EXP2_POLY[2], rr, 1.0);
  yy = CRT_POWF5_FMA(zz, rr2, yy);
  yy = CRT_P5_MUL(yy, s);
  return CRT_P5_D2F(yy);
}
#endif /* CRT_POWF5_H */
