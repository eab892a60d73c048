// vimg_libm.cuh -- float sine/cosine with glibc's exact results.
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

}  // namespace vimg_libm
