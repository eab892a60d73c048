// vimg_libm.cuh -- glibc's float sinf / cosf / logf / log2f / powf / acosf / atanf / atan2f with the
// library's exact results (sine / cosine first; the others follow further down).
//
// The reference draws directions with std::sin / std::cos on floats
// (include/rng/sampling.h:15-22,69-79, background.h:120-127); g++ lowers those to
// glibc's sinf / cosf / sincosf.  A last-bit difference there changes the sampled
// direction, after which the GPU path and the CPU path stop sharing an RNG trajectory.
// CUDA's sinf/cosf are 1-2 ulp, glibc's (2.28+, the ARM optimized-routines algorithm)
// evaluate a degree-7/8 polynomial in DOUBLE after a one-step range reduction and round
// once -- which a GPU can reproduce operation by operation, because IEEE double
// add/mul/fma are exact on both sides.
//
// This restates the x86-64 `*_fma` ifunc variants of glibc 2.39 (the ones every
// FMA-capable host selects): same coefficients (sincosf table), same reduction
// n = round(x * 2/pi), xr = fma(-n, pi/2, x), same polynomial grouping and the same
// fused operations the compiled library uses.  oracle/tools/libm_check.c verifies it
// bit-for-bit against the host's sinf/cosf/sincosf over every float in (-120, 120).
// |x| >= 120 (never produced by the renderer: angles are in [0, 2 pi]) falls back to the
// correctly-rounded-in-practice double evaluation.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VLM_FN __host__ __device__ __forceinline__
#else
#define VLM_FN static inline
#endif

namespace vimg_libm {

// |x| >= 120 fallback; out of line on the device so the (never taken) slow path costs no I-cache
#if defined(__CUDA_ARCH__)
__device__ __noinline__ static double slow_sin(double x) { return sin(x); }
__device__ __noinline__ static double slow_cos(double x) { return cos(x); }
#else
static inline double slow_sin(double x) { return sin(x); }
static inline double slow_cos(double x) { return cos(x); }
#endif

// __sincosf_table[0] of glibc (sysdeps/ieee754/flt-32/s_sincosf_data.c).  Table [1], used
// for quadrants 2 and 3, is the same with c0..c4 negated; round-to-nearest is symmetric, so
// the negated polynomial is exactly the negated result and one coefficient set suffices.
#define VLM_HPI_INV 0x1.45F306DC9C883p+23 /* 2/pi * 2^24 */
#define VLM_HPI 0x1.921FB54442D18p0       /* pi/2 */
#define VLM_C0 0x1p0
#define VLM_C1 -0x1.ffffffd0c621cp-2
#define VLM_C2 0x1.55553e1068f19p-5
#define VLM_C3 -0x1.6c087e89a359dp-10
#define VLM_C4 0x1.99343027bf8c3p-16
#define VLM_S1 -0x1.555545995a603p-3
#define VLM_S2 0x1.1107605230bc4p-7
#define VLM_S3 -0x1.994eb3774cf24p-13

VLM_FN uint32_t abstop12(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  return (u >> 20) & 0x7ff;
}

// sinf_poly, even branch: s + x7 * s1 with s1 = S2 + x2*S3, s = x + x3*S1
VLM_FN double poly_sin(double x, double x2) {
  double s1 = fma(x2, VLM_S3, VLM_S2);
  double x3 = x2 * x;
  double x7 = x2 * x3;
  double s = fma(x3, VLM_S1, x);
  return fma(s1, x7, s);
}
// sinf_poly, odd branch: c + x6 * c2 with c1 = C0 + x2*C1, c2 = C3 + x2*C4, c = c1 + x4*C2
VLM_FN double poly_cos(double x2) {
  double x4 = x2 * x2;
  double c1 = fma(x2, VLM_C1, VLM_C0);
  double c2 = fma(x2, VLM_C4, VLM_C3);
  double x6 = x2 * x4;
  double c = fma(x4, VLM_C2, c1);
  return fma(c2, x6, c);
}

// reduce_fast: n = round(x * 2/pi) via the 2^24-scaled constant, xr = x - n*pi/2 (fused)
VLM_FN double reduce_fast(double x, int& n) {
  double r = x * VLM_HPI_INV;
  n = ((int32_t)r + 0x800000) >> 24;
  return fma(-(double)n, VLM_HPI, x);
}

// s_sinf.c / s_cosf.c after reduce_fast: s = sign[n & 3] = {1,-1,-1,1}, p = table[(n & 2) != 0],
// then sinf_poly(xr * s, xr * xr, p, n) for sinf and sinf_poly(..., n ^ 1) for cosf; the last
// argument's low bit picks the cosine polynomial.
VLM_FN float sincos_quadrant(double xr, int n, int use_cos_poly) {
  double x2 = xr * xr;
  if (!use_cos_poly) {
    const int q = n & 3;
    double sgn = (q == 1 || q == 2) ? -1.0 : 1.0;
    return (float)poly_sin(xr * sgn, x2);
  }
  double c = poly_cos(x2);
  return (float)((n & 2) ? -c : c);
}

VLM_FN float glibc_sinf(float y) {
  double x = y;
  const uint32_t top = abstop12(y);
  if (top <= 0x3f3) {            // |y| < pi/4
    if (top <= 0x397) return y;  // |y| < 2^-12
    return (float)poly_sin(x, x * x);
  }
  if (top <= 0x42e) {  // |y| < 120
    int n;
    double xr = reduce_fast(x, n);
    return sincos_quadrant(xr, n, n & 1);
  }
  return (float)slow_sin(x);
}

VLM_FN float glibc_cosf(float y) {
  double x = y;
  const uint32_t top = abstop12(y);
  if (top <= 0x3f3) {
    if (top <= 0x397) return 1.0f;
    return (float)poly_cos(x * x);
  }
  if (top <= 0x42e) {
    int n;
    double xr = reduce_fast(x, n);
    return sincos_quadrant(xr, n, (n ^ 1) & 1);  // cosf: sinf_poly(x * s, x2, p, n ^ 1), s and p from n
  }
  return (float)slow_cos(x);
}


// ------------------------------------------------------------------ powf / log2f ----
// glibc 2.39's powf and log2f (sysdeps/ieee754/flt-32/e_powf.c, e_log2f.c: the ARM
// optimized-routines algorithm), x86-64 `*_fma` ifunc variants.  Both evaluate in DOUBLE from a
// 16-entry {1/c, log2(c)} table and round once; powf continues with a 32-entry 2^(i/32) table.
// The compiled library contracts EVERY `a * b + c` of the source into one fused operation
// (checked in the disassembly of libm.so.6: vfmadd for r, y, p, q and the exp2 polynomial), which
// is what the explicit fma() calls below restate.  The reference calls them through std::pow /
// std::log2 on floats: sRGB OETF and agx (include/color_utils.h:50-68, src/tonemap/agx.cpp),
// sRGB -> linear (color_utils.h:28-47), texture LOD (ray cones).
// oracle/tools/libm_check.cpp compares them bit for bit with the host libm: log2f over every
// positive float, powf over every float of (0, 2] for the exponents the reference uses.
#define VLM_LOG2_TABLE(X)                                                                          \
  X(0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2) X(0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2)     \
  X(0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2) X(0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2)     \
  X(0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2) X(0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3)     \
  X(0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3) X(0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4)     \
  X(0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5) X(0x1p+0, 0x0p+0)                                  \
  X(0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4) X(0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3)       \
  X(0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3) X(0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2)       \
  X(0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2) X(0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2)
#define VLM_EXP2_TABLE                                                                             \
  0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,      \
  0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,      \
  0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,      \
  0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,      \
  0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,      \
  0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,      \
  0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,      \
  0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull

// __logf_data.tab[i].logc = ln(c); its invc column is the one above
#define VLM_LN_TABLE                                                                               \
  -0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3,      \
  -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c8100p-3, -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4,      \
  -0x1.252f438e10c1ep-5, 0x0p+0, 0x1.aa5aa5df25984p-5, 0x1.c5e53aa362eb4p-4,                       \
  0x1.526e57720db08p-3, 0x1.bc2860d224770p-3, 0x1.1058bc8a07ee1p-2, 0x1.4043057b6ee09p-2
#define VLM_X_PAIR(a, b) a, b,
static const double vlm_host_log2_tab[32] = {VLM_LOG2_TABLE(VLM_X_PAIR)};  // __log2f_data.tab == __powf_log2_data.tab
static const uint64_t vlm_host_exp2_tab[32] = {VLM_EXP2_TABLE};           // __exp2f_data.tab
static const double vlm_host_ln_tab[16] = {VLM_LN_TABLE};
#if defined(__CUDACC__)
__device__ static const double vlm_dev_log2_tab[32] = {VLM_LOG2_TABLE(VLM_X_PAIR)};
__device__ static const uint64_t vlm_dev_exp2_tab[32] = {VLM_EXP2_TABLE};
__device__ static const double vlm_dev_ln_tab[16] = {VLM_LN_TABLE};
#endif
#if defined(__CUDA_ARCH__)
#define VLM_LOG2_TAB vlm_dev_log2_tab
#define VLM_EXP2_TAB vlm_dev_exp2_tab
#define VLM_LN_TAB vlm_dev_ln_tab
__device__ __noinline__ static double slow_pow(double x, double y) { return pow(x, y); }
#else
#define VLM_LOG2_TAB vlm_host_log2_tab
#define VLM_EXP2_TAB vlm_host_exp2_tab
#define VLM_LN_TAB vlm_host_ln_tab
static inline double slow_pow(double x, double y) { return pow(x, y); }
#endif

VLM_FN uint32_t f2u(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  return u;
}
VLM_FN float u2f(uint32_t u) {
  float x;
  memcpy(&x, &u, 4);
  return x;
}
VLM_FN uint64_t d2u(double x) {
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
}
VLM_FN double u2d(uint64_t u) {
  double x;
  memcpy(&x, &u, 8);
  return x;
}

// log2_inline's table step, shared by log2f and powf: x = 2^k z with z in [OFF, 2 OFF),
// c = the centre of z's sub-interval; returns r = z/c - 1 (fused) and y0 = log2(c) + k
VLM_FN void log2_reduce(uint32_t ix, double& r, double& y0) {
  const uint32_t tmp = ix - 0x3f330000u;  // OFF
  const uint32_t i = (tmp >> 19) & 15u;
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ix - top;
  const int k = (int32_t)tmp >> 23;  // arithmetic shift
  const double invc = VLM_LOG2_TAB[2 * i], logc = VLM_LOG2_TAB[2 * i + 1];
  r = fma((double)u2f(iz), invc, -1.0);
  y0 = logc + (double)k;
}

// e_log2f.c: log2(x) = log1p(z/c - 1)/ln2 + log2(c) + k, degree-4 polynomial for the first term
VLM_FN float glibc_log2f(float x) {
  uint32_t ix = f2u(x);
  if (ix == 0x3f800000u) return 0.f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {  // x < 0x1p-126, inf or nan
    if (ix * 2u == 0) return u2f(0xff800000u);           // log2(+-0) = -inf
    if (ix == 0x7f800000u) return x;                     // log2(inf) = inf
    if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return u2f(0x7fc00000u);  // x < 0 or nan
    ix = f2u(x * 0x1p23f);  // subnormal: normalise
    ix -= 23u << 23;
  }
  double r, y0;
  log2_reduce(ix, r, y0);
  const double r2 = r * r;
  double y = fma(0x1.ecabf496832e0p-2, r, -0x1.715479ffae3dep-1);  // A[1] * r + A[2]
  y = fma(-0x1.712b6f70a7e4dp-2, r2, y);                           // A[0] * r2 + y
  const double p = fma(0x1.715475f35c8b8p+0, r, y0);               // A[3] * r + y0
  return (float)fma(y, r2, p);
}

// e_logf.c: ln(x) = log1p(z/c - 1) + ln(c) + k ln2, degree-3 polynomial for the first term
VLM_FN float glibc_logf(float x) {
  uint32_t ix = f2u(x);
  if (ix == 0x3f800000u) return 0.f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2u == 0) return u2f(0xff800000u);
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return u2f(0x7fc00000u);
    ix = f2u(x * 0x1p23f);
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const uint32_t i = (tmp >> 19) & 15u;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const int k = (int32_t)tmp >> 23;
  const double r = fma((double)u2f(iz), VLM_LOG2_TAB[2 * i], -1.0);
  const double y0 = fma((double)k, 0x1.62e42fefa39efp-1, VLM_LN_TAB[i]);  // logc + k * Ln2
  const double r2 = r * r;
  double y = fma(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);  // A[1] * r + A[2]
  y = fma(-0x1.00ea348b88334p-2, r2, y);                           // A[0] * r2 + y
  return (float)fma(y, r2, y0 + r);
}

// e_powf.c: 2^(y log2 x); log2 with a degree-5 polynomial (relative error 2^-58 or so), exp2f's
// table step with k = round(32 ylogx) taken from the low bits of ylogx + 0x1.8p+47.
// Restated: x a positive finite float (subnormals are normalised like the library does), y finite
// and not 0, overflow / underflow / subnormal results included.  Everything else (x <= 0, inf,
// nan, y = 0; the callers above have x in [0, 1], and x = 0 gives 0) is the double pow.
VLM_FN float glibc_powf(float x, float y) {
  uint32_t ix = f2u(x);
  const uint32_t iy = f2u(y);
  if (2u * iy - 1u >= 2u * 0x7f800000u - 1u) return (float)slow_pow((double)x, (double)y);  // y = 0, inf, nan
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix == 0 || ix >= 0x00800000u) return (float)slow_pow((double)x, (double)y);  // 0, negative, inf, nan
    ix = f2u(x * 0x1p23f) & 0x7fffffffu;  // subnormal: normalise
    ix -= 23u << 23;
  }
  double r, y0;
  log2_reduce(ix, r, y0);
  const double r2 = r * r;
  double yy = fma(0x1.27616c9496e0bp-2, r, -0x1.71969a075c67ap-2);  // A[0] * r + A[1]
  const double p = fma(0x1.ec70a6ca7baddp-2, r, -0x1.7154748bef6c8p-1);  // A[2] * r + A[3]
  const double r4 = r2 * r2;
  double q = fma(0x1.71547652ab82bp+0, r, y0);  // A[4] * r + y0
  q = fma(p, r2, q);
  const double logx = fma(yy, r4, q);
  const double ylogx = (double)y * logx;
  if (((d2u(ylogx) >> 47) & 0xffffu) >= 0x80bfu) {               // |ylogx| >= 126 (sign_bias = 0: x > 0)
    if (ylogx > 0x1.fffffffd1d571p+6) return u2f(0x7f800000u);  // __math_oflowf
    if (ylogx <= -150.0) return 0.f;                             // __math_uflowf
    if (ylogx < -149.0) return 0x1p-149f;                        // __math_may_uflowf: 0x1.4p-75f squared
    // else: a subnormal (or huge) result, rounded ONCE from the double product below
  }
  // exp2_inline: ylogx = k/32 + rr, |rr| <= 1/64
  double kd = ylogx + 0x1.8p+47;  // SHIFT / N
  const uint64_t ki = d2u(kd);
  kd -= 0x1.8p+47;
  const double rr = ylogx - kd;
  const uint64_t t = VLM_EXP2_TAB[ki & 31u] + (ki << 47);
  const double s = u2d(t);
  const double z = fma(0x1.c6af84b912394p-5, rr, 0x1.ebfce50fac4f3p-3);  // C[0] * r + C[1]
  const double rr2 = rr * rr;
  double e = fma(0x1.62e42ff0c52d6p-1, rr, 1.0);  // C[2] * r + 1
  e = fma(z, rr2, e);
  return (float)(e * s);
}

// ------------------------------------------------------------- acosf / atanf / atan2f ----
// glibc 2.39's acosf, atanf and atan2f are still the fdlibm float routines
// (sysdeps/ieee754/flt-32/e_acosf.c, s_atanf.c, e_atan2f.c); they have no `*_fma` variant, the
// library code is plain single-precision mulss / addss / divss / sqrtss (disassembly of libm.so.6),
// i.e. the source's expressions evaluated left to right with no contraction -- restated below with
// the same constants and the same association.  The reference calls them for the env-map
// direction <-> uv maps (background.h) and the primary ray cone's spread angle (tl_camera.cpp).
// This header is compiled with -fmad=false (device) / -ffp-contract=off (host check).
VLM_FN float acos_poly_ratio(float z) {  // r = p / q of e_acosf.c
  const float p = z * (u2f(0x3e2aaaabu) + z * (u2f(0xbea6b090u) + z * (u2f(0x3e4e0aa8u) + z * (u2f(0xbd241146u)
                  + z * (u2f(0x3a4f7f04u) + z * u2f(0x3811ef08u))))));
  const float q = 1.0f + z * (u2f(0xc019d139u) + z * (u2f(0x4001572du) + z * (u2f(0xbf303361u) + z * u2f(0x3d9dc62eu))));
  return p / q;
}
VLM_FN float glibc_acosf(float x) {
  const float pi = u2f(0x40490fdau), pio2_hi = u2f(0x3fc90fdau), pio2_lo = u2f(0x33a22168u);
  const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
  if (ix == 0x3f800000u) return (int32_t)hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;  // |x| == 1
  if (ix > 0x3f800000u) return u2f(0x7fc00000u);                                // |x| > 1 or nan
  if (ix < 0x3f000000u) {                                                       // |x| < 0.5
    if (ix <= 0x32800000u) return pio2_hi + pio2_lo;
    const float r = acos_poly_ratio(x * x);
    return pio2_hi - (x - (pio2_lo - r * x));
  }
  if ((int32_t)hx < 0) {  // x < -0.5
    const float z = (1.0f + x) * 0.5f;
    const float r = acos_poly_ratio(z), s = sqrtf(z);
    const float w = r * s - pio2_lo;
    return pi - 2.0f * (s + w);
  }
  const float z = (1.0f - x) * 0.5f;  // x > 0.5
  const float s = sqrtf(z);
  const float df = u2f(f2u(s) & 0xfffff000u);
  const float c = (z - df * df) / (s + df);
  const float r = acos_poly_ratio(z);
  const float w = r * s + c;
  return 2.0f * (df + w);
}

VLM_FN float glibc_atanf(float x) {
  const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
  float hi, lo;
  int id;
  if (ix >= 0x4c000000u) {  // |x| >= 2^25
    if (ix > 0x7f800000u) return x + x;
    const float hi3 = u2f(0x3fc90fdau), lo3 = u2f(0x33a22168u);
    return (int32_t)hx > 0 ? hi3 + lo3 : -hi3 - lo3;
  }
  if (ix < 0x3ee00000u) {  // |x| < 0.4375
    if (ix < 0x31000000u) return x;
    id = -1;
    hi = lo = 0.f;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000u) {    // |x| < 1.1875
      if (ix < 0x3f300000u) {  // 7/16 <= |x| < 11/16
        id = 0, hi = u2f(0x3eed6338u), lo = u2f(0x31ac3769u);
        x = (2.0f * x - 1.0f) / (2.0f + x);
      } else {
        id = 1, hi = u2f(0x3f490fdau), lo = u2f(0x33222168u);
        x = (x - 1.0f) / (x + 1.0f);
      }
    } else if (ix < 0x401c0000u) {  // |x| < 2.4375
      id = 2, hi = u2f(0x3f7b985eu), lo = u2f(0x33140fb4u);
      x = (x - 1.5f) / (1.0f + 1.5f * x);
    } else {
      id = 3, hi = u2f(0x3fc90fdau), lo = u2f(0x33a22168u);
      x = -1.0f / x;
    }
  }
  const float z = x * x, w = z * z;
  const float s1 = z * (u2f(0x3eaaaaabu) + w * (u2f(0x3e124925u) + w * (u2f(0x3dba2e6eu) + w * (u2f(0x3d886b35u)
                   + w * (u2f(0x3d4bda59u) + w * u2f(0x3c8569d7u))))));
  const float s2 = w * (u2f(0xbe4ccccdu) + w * (u2f(0xbde38e38u) + w * (u2f(0xbd9d8795u) + w * (u2f(0xbd6ef16bu)
                   + w * u2f(0xbd15a221u)))));
  if (id < 0) return x - x * (s1 + s2);
  const float r = hi - ((x * (s1 + s2) - lo) - x);
  return (int32_t)hx < 0 ? -r : r;
}

VLM_FN float glibc_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = u2f(0x3f490fdbu), pi_o_2 = u2f(0x3fc90fdbu), pi = u2f(0x40490fdbu),
              pi_lo = u2f(0xb3bbbd2eu);
  const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu, hy = f2u(y), iy = hy & 0x7fffffffu;
  if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;
  if (hx == 0x3f800000u) return glibc_atanf(y);
  const uint32_t m = ((hy >> 31) & 1u) | ((hx >> 30) & 2u);  // 2 * sign(x) + sign(y)
  if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
  if (ix == 0) return (int32_t)hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000u) {
    if (iy == 0x7f800000u) {
      switch (m) {
        case 0: return pi_o_4 + tiny;
        case 1: return -pi_o_4 - tiny;
        case 2: return 3.0f * pi_o_4 + tiny;
        default: return -3.0f * pi_o_4 - tiny;
      }
    }
    switch (m) {
      case 0: return 0.0f;
      case 1: return -0.0f;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (iy == 0x7f800000u) return (int32_t)hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int32_t k = ((int32_t)iy - (int32_t)ix) >> 23;
  float z;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if ((int32_t)hx < 0 && k < -60) z = 0.0f;
  else z = glibc_atanf(fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

// ------------------------------------------------------------------ log2 (double) ----
// glibc 2.39's double log2 (sysdeps/ieee754/dbl-64/e_log2.c, ARM optimized routines), `__log2_fma`
// variant: 64-entry {1/c, log2(c)} table, r = fma(z, 1/c, -1), r/ln2 split into hi + lo with an fma,
// degree-6 polynomial; a separate degree-10 path near 1.  Fusions as in the library's disassembly.
// The reference reaches it through unqualified log2() / std::log2 on doubles: the env map's mip
// level (background.h:100) and compute_texture_LOD's log2(width * height) (texture_RGB.h:141).
// oracle/tools/libm_check.cpp compares it with the host libm on 2^30 pseudo-random doubles (all
// binades, the neighbourhood of 1, subnormals) and on every float in [2^-20, 2^20] widened.
#define VLM_LOG2D_TABLE(X) \
  X(0x1.724286bb1acf8p+0, -0x1.1095feecdb000p-1) X(0x1.6e1f766d2cca1p+0, -0x1.08494bd76d000p-1) \
  X(0x1.6a13d0e30d48ap+0, -0x1.00143aee8f800p-1) X(0x1.661ec32d06c85p+0, -0x1.efec5360b4000p-2) \
  X(0x1.623fa951198f8p+0, -0x1.dfdd91ab7e000p-2) X(0x1.5e75ba4cf026cp+0, -0x1.cffae0cc79000p-2) \
  X(0x1.5ac055a214fb8p+0, -0x1.c043811fda000p-2) X(0x1.571ed0f166e1ep+0, -0x1.b0b67323ae000p-2) \
  X(0x1.53909590bf835p+0, -0x1.a152f5a2db000p-2) X(0x1.5014fed61adddp+0, -0x1.9217f5af86000p-2) \
  X(0x1.4cab88e487bd0p+0, -0x1.8304db0719000p-2) X(0x1.49539b4334feep+0, -0x1.74189f9a9e000p-2) \
  X(0x1.460cbdfafd569p+0, -0x1.6552bb5199000p-2) X(0x1.42d664ee4b953p+0, -0x1.56b23a29b1000p-2) \
  X(0x1.3fb01111dd8a6p+0, -0x1.483650f5fa000p-2) X(0x1.3c995b70c5836p+0, -0x1.39de937f6a000p-2) \
  X(0x1.3991c4ab6fd4ap+0, -0x1.2baa1538d6000p-2) X(0x1.3698e0ce099b5p+0, -0x1.1d98340ca4000p-2) \
  X(0x1.33ae48213e7b2p+0, -0x1.0fa853a40e000p-2) X(0x1.30d191985bdb1p+0, -0x1.01d9c32e73000p-2) \
  X(0x1.2e025cab271d7p+0, -0x1.e857da2fa6000p-3) X(0x1.2b404cf13cd82p+0, -0x1.cd3c8633d8000p-3) \
  X(0x1.288b02c7ccb50p+0, -0x1.b26034c14a000p-3) X(0x1.25e2263944de5p+0, -0x1.97c1c2f4fe000p-3) \
  X(0x1.234563d8615b1p+0, -0x1.7d6023f800000p-3) X(0x1.20b46e33eaf38p+0, -0x1.633a71a05e000p-3) \
  X(0x1.1e2eefdcda3ddp+0, -0x1.494f5e9570000p-3) X(0x1.1bb4a580b3930p+0, -0x1.2f9e424e0a000p-3) \
  X(0x1.19453847f2200p+0, -0x1.162595afdc000p-3) X(0x1.16e06c0d5d73cp+0, -0x1.f9c9a75bd8000p-4) \
  X(0x1.1485f47b7e4c2p+0, -0x1.c7b575bf9c000p-4) X(0x1.12358ad0085d1p+0, -0x1.960c60ff48000p-4) \
  X(0x1.0fef00f532227p+0, -0x1.64ce247b60000p-4) X(0x1.0db2077d03a8fp+0, -0x1.33f78b2014000p-4) \
  X(0x1.0b7e6d65980d9p+0, -0x1.0387d1a42c000p-4) X(0x1.0953efe7b408dp+0, -0x1.a6f9208b50000p-5) \
  X(0x1.07325cac53b83p+0, -0x1.47a954f770000p-5) X(0x1.05197e40d1b5cp+0, -0x1.d23a8c50c0000p-6) \
  X(0x1.03091c1208ea2p+0, -0x1.16a2629780000p-6) X(0x1.0101025b37e21p+0, -0x1.720f8d8e80000p-8) \
  X(0x1.fc07ef9caa76bp-1, 0x1.6fe53b1500000p-7) X(0x1.f4465d3f6f184p-1, 0x1.11ccce10f8000p-5) \
  X(0x1.ecc079f84107fp-1, 0x1.c4dfc8c8b8000p-5) X(0x1.e573a99975ae8p-1, 0x1.3aa321e574000p-4) \
  X(0x1.de5d6f0bd3de6p-1, 0x1.918a0d08b8000p-4) X(0x1.d77b681ff38b3p-1, 0x1.e72e9da044000p-4) \
  X(0x1.d0cb5724de943p-1, 0x1.1dcd2507f6000p-3) X(0x1.ca4b2dc0e7563p-1, 0x1.476ab03dea000p-3) \
  X(0x1.c3f8ee8d6cb51p-1, 0x1.7074377e22000p-3) X(0x1.bdd2b4f020c4cp-1, 0x1.98ede8ba94000p-3) \
  X(0x1.b7d6c006015cap-1, 0x1.c0db86ad2e000p-3) X(0x1.b20366e2e338fp-1, 0x1.e840aafcee000p-3) \
  X(0x1.ac57026295039p-1, 0x1.0790ab4678000p-2) X(0x1.a6d01bc2731ddp-1, 0x1.1ac056801c000p-2) \
  X(0x1.a16d3bc3ff18bp-1, 0x1.2db11d4fee000p-2) X(0x1.9c2d14967feadp-1, 0x1.406464ec58000p-2) \
  X(0x1.970e4f47c9902p-1, 0x1.52dbe093af000p-2) X(0x1.920fb3982bcf2p-1, 0x1.651902050d000p-2) \
  X(0x1.8d30187f759f1p-1, 0x1.771d2cdeaf000p-2) X(0x1.886e5ebb9f66dp-1, 0x1.88e9c857d9000p-2) \
  X(0x1.83c97b658b994p-1, 0x1.9a80155e16000p-2) X(0x1.7f405ffc61022p-1, 0x1.abe186ed3d000p-2) \
  X(0x1.7ad22181415cap-1, 0x1.bd0f2aea0e000p-2) X(0x1.767dcf99eff8cp-1, 0x1.ce0a43dbf4000p-2)
static const double vlm_host_log2d_tab[128] = {VLM_LOG2D_TABLE(VLM_X_PAIR)};
#if defined(__CUDACC__)
__device__ static const double vlm_dev_log2d_tab[128] = {VLM_LOG2D_TABLE(VLM_X_PAIR)};
#endif
#if defined(__CUDA_ARCH__)
#define VLM_LOG2D_TAB vlm_dev_log2d_tab
#else
#define VLM_LOG2D_TAB vlm_host_log2d_tab
#endif

VLM_FN double glibc_log2(double x) {
  const double InvLn2hi = 0x1.7154765200000p+0, InvLn2lo = 0x1.705fc2eefa200p-33;
  uint64_t ix = d2u(x);
  const uint64_t LO = 0x3feea4af00000000ull;  // asuint64(1.0 - 0x1.5b51p-5)
  if (ix - LO <= 0x210a9ffffffffull) {        // [LO, asuint64(1.0 + 0x1.6ab2p-5)): close to 1
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = x - 1.0;
    const double hi = r * InvLn2hi;
    double lo = fma(r, InvLn2lo, fma(r, InvLn2hi, -hi));
    const double r2 = r * r, r4 = r2 * r2;
    const double q = fma(r, 0x1.ec709dc3a03f7p-2, -0x1.71547652b82fep-1);  // B[0] + r * B[1]
    const double y = fma(q, r2, hi);                                         // hi + p, p = r2 * q
    lo = fma(q, r2, hi - y) + lo;                                            // lo += hi - y + p
    const double b23 = fma(r, 0x1.2776c50f05be4p-2, -0x1.71547652b7c3fp-2);
    const double b45 = fma(r, 0x1.a61761ec4e736p-3, -0x1.ec709dd768fe5p-3);
    const double b67 = fma(r, 0x1.484d154f01b4ap-3, -0x1.7153fbc64a79bp-3);
    const double b89 = fma(r, 0x1.0b32f285aee66p-3, -0x1.289e4a72c383cp-3);
    const double lowp = fma(b45, r2, b23);
    const double highp = fma(b89, r2, b67);
    lo = fma(fma(highp, r4, lowp), r4, lo);
    return y + lo;
  }
  const uint32_t top = (uint32_t)(ix >> 48);
  if (top - 0x0010u >= 0x7ff0u - 0x0010u) {  // x < 0x1p-1022, inf or nan
    if (ix * 2 == 0) return u2d(0xfff0000000000000ull);
    if (ix == 0x7ff0000000000000ull) return x;
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return u2d(0x7ff8000000000000ull);
    ix = d2u(x * 0x1p52);  // subnormal: normalise
    ix -= 52ull << 52;
  }
  const uint64_t tmp = ix - 0x3fe6000000000000ull;  // OFF
  const uint32_t i = (uint32_t)(tmp >> 46) & 63u;
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
  const double invc = VLM_LOG2D_TAB[2 * i], logc = VLM_LOG2D_TAB[2 * i + 1];
  const double z = u2d(iz), kd = (double)k;
  const double r = fma(z, invc, -1.0);
  const double t1 = r * InvLn2hi;
  const double t2 = fma(r, InvLn2lo, fma(r, InvLn2hi, -t1));
  const double t3 = kd + logc;
  const double hi = t3 + t1;
  const double lo = ((t3 - hi) + t1) + t2;
  const double r2 = r * r, r4 = r2 * r2;
  const double a01 = fma(r, 0x1.ec709dc3a04bep-2, -0x1.71547652b8339p-1);
  const double a23 = fma(r, 0x1.2776c50034c48p-2, -0x1.7154764702ffbp-2);
  const double a45 = fma(r, 0x1.a6225e117f92ep-3, -0x1.ec7b328ea92bcp-3);
  const double p = fma(a45, r4, fma(a23, r2, a01));
  return fma(r2, p, lo) + hi;
}

}  // namespace vimg_libm
