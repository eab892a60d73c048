// vimg_math.cuh -- arithmetic with the reference's exact semantics.
//
// Bit-exact parity with the reference's scalar path
// fixes the ORDER of every float operation.  This header restates, for the
// device, the glm 1.0.1 scalar expressions the reference's hot path uses and
// the libstdc++ definitions of std::min/max/clamp/lerp.  The translation unit
// must be compiled with -fmad=false: the only fused multiply-adds are the
// explicit fmaf() calls where the reference calls std::fma
// (include/geometry/triangle.h:11-15, disney_common.h:40).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "vimg_libm.cuh"

#define VD __device__ __forceinline__
// VIMG_LEAN 1 (vimg_lean.cu): a build of the shade / END kernels for scenes WITHOUT image textures, normal maps,
// metallic-roughness textures, an env map and meshes whose uv sets differ -- the reference's own scenes.  The code of
// those features is what sets the register need of the kernels (Lambertian shade 167 -> 126, END 117 -> 91 without
// it), whether a scene uses them or not; vimg_api.cu looks at the scene (vimg_scene::lean) and launches this build
// when none is there.  Every test below is `VIMG_LEAN ||` / `!VIMG_LEAN &&` in front of the run-time test it replaces.
#ifndef VIMG_LEAN
#define VIMG_LEAN 0
#endif

namespace vimg {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

VD V2 mk2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
VD V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
VD V3 mk3(float s) { return mk3(s, s, s); }
VD V4 mk4(float x, float y, float z, float w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// component-wise, like glm's vec operators
VD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
VD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
VD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
VD V3 operator/(V3 a, V3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
VD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
VD V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
VD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
VD V3 operator+(V3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
VD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
VD V2 operator+(V2 a, V2 b) { return mk2(a.x + b.x, a.y + b.y); }
VD V2 operator-(V2 a, V2 b) { return mk2(a.x - b.x, a.y - b.y); }
VD V2 operator*(V2 a, V2 b) { return mk2(a.x * b.x, a.y * b.y); }
VD V2 operator*(V2 a, float s) { return mk2(a.x * s, a.y * s); }
VD V2 operator*(float s, V2 a) { return mk2(s * a.x, s * a.y); }
VD V2 operator-(V2 a) { return mk2(-a.x, -a.y); }
VD V4 operator+(V4 a, V4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
VD V4 operator*(V4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }

// glm::dot -- compute_dot<vec3>: tmp = a*b; (tmp.x + tmp.y) + tmp.z
VD float dot(V3 a, V3 b) {
  float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
  return tx + ty + tz;
}
VD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
VD float length2(V3 v) { return dot(v, v); }
VD float length(V3 v) { return sqrtf(dot(v, v)); }
// glm::normalize = v * inversesqrt(dot(v,v)), inversesqrt = 1 / sqrt
VD V3 normalize(V3 v) { return v * (1.0f / sqrtf(dot(v, v))); }
VD V2 normalize(V2 v) { return v * (1.0f / sqrtf(dot(v, v))); }
VD V3 cross(V3 a, V3 b) {
  return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
// glm::min(a,b) = (b < a) ? b : a ; glm::max(a,b) = (a < b) ? b : a
VD float glm_min(float a, float b) { return (b < a) ? b : a; }
VD float glm_max(float a, float b) { return (a < b) ? b : a; }
// std::min / std::max / std::clamp (libstdc++)
VD float std_min(float a, float b) { return (b < a) ? b : a; }
VD float std_max(float a, float b) { return (a < b) ? b : a; }
VD float std_clamp(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }
VD int std_clampi(int v, int lo, int hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }
// glm::mix(x, y, a) = x * (1 - a) + y * a
VD V3 mix(V3 x, V3 y, float a) { return x * (1.0f - a) + y * a; }
VD V2 mix(V2 x, V2 y, float a) { return x * (1.0f - a) + y * a; }
// std::lerp(float) as libstdc++ implements it (<cmath> __lerp)
VD float std_lerp(float a, float b, float t) {
  if ((a <= 0 && b >= 0) || (a >= 0 && b <= 0)) return t * b + (1 - t) * a;
  if (t == 1) return b;
  const float x = a + t * (b - a);
  return (t > 1) == (b > a) ? (b < x ? x : b) : (b > x ? x : b);
}

// glm mat4 (column-major, m[col*4+row]) times vec4:
//   (m0*v0 + m1*v1) + (m2*v2 + m3*v3)       type_mat4x4.inl operator*(mat, vec)
VD V4 mat4_mul(const float* m, V4 v) {
  V4 m0 = mk4(m[0], m[1], m[2], m[3]), m1 = mk4(m[4], m[5], m[6], m[7]);
  V4 m2 = mk4(m[8], m[9], m[10], m[11]), m3 = mk4(m[12], m[13], m[14], m[15]);
  V4 add0 = m0 * v.x + m1 * v.y;
  V4 add1 = m2 * v.z + m3 * v.w;
  return add0 + add1;
}

// ---- libm ------------------------------------------------------------------
// The reference calls glibc.  Every FLOAT libm function on the path -- sinf / cosf (sampled
// directions, thin-lens primary rays), logf (clearcoat), log2f (texture LOD), acosf / atan2f
// (env-map uv, ray-cone spread), powf (post-process, sRGB) -- is restated operation by operation
// in vimg_libm.cuh and checked against the host libm over all float inputs.  Double expressions
// of the reference call the double functions directly: log2 is glibc's, restated
// too; sin / cos / pow / acos / atan2 are CUDA's (<= 2 ulp in double, rounded to float afterwards).
// CUDA's double-precision library functions expand to hundreds of instructions each when
// inlined; the render kernels call them from many places, so they live in ONE out-of-line
// copy each (instruction-cache footprint, see DESIGN.md "Kernel structure").
#define VNI __device__ __noinline__
#ifndef VIMG_FAST
// `vdbl` marks the sub-expressions the reference evaluates in double
typedef double vdbl;
VNI double d_sin(double x) { return sin(x); }
VNI double d_cos(double x) { return cos(x); }
VNI double d_log(double x) { return log(x); }
#ifndef VIMG_FAST_LIBM
VNI double d_log2(double x) { return vimg_libm::glibc_log2(x); }  // glibc's double log2, restated exactly
#else
VNI double d_log2(double x) { return log2(x); }
#endif
VNI double d_pow(double x, double y) { return pow(x, y); }
VNI double d_acos(double x) { return acos(x); }
VNI double d_atan2(double y, double x) { return atan2(y, x); }
#else
// Tolerance build (vimg_fast.cu, -use_fast_math): the same expressions in float with the hardware's
// approximate division / sqrt / transcendental units.  Images agree with the reference within the
// stated relMSE tolerance (tests/test_gpu_fast.py), not bit for bit.
// (VIMG_FAST_DBL / VIMG_FAST_NO_INTRIN: bisection knobs of scripts/fast_bias.sh)
#ifdef VIMG_FAST_DBL
typedef double vdbl;
#else
typedef float vdbl;
#endif
#ifdef VIMG_FAST_NO_INTRIN
VD vdbl d_sin(vdbl x) { return sin(x); }
VD vdbl d_cos(vdbl x) { return cos(x); }
VD vdbl d_log(vdbl x) { return log(x); }
VD vdbl d_log2(vdbl x) { return log2(x); }
VD vdbl d_pow(vdbl x, vdbl y) { return pow(x, y); }
#else
VD vdbl d_sin(vdbl x) { return (vdbl)__sinf((float)x); }
VD vdbl d_cos(vdbl x) { return (vdbl)__cosf((float)x); }
VD vdbl d_log(vdbl x) { return (vdbl)__logf((float)x); }
VD vdbl d_log2(vdbl x) { return (vdbl)__log2f((float)x); }
VD vdbl d_pow(vdbl x, vdbl y) { return (vdbl)__powf((float)x, (float)y); }
#endif
VD vdbl d_acos(vdbl x) { return acos(x); }
VD vdbl d_atan2(vdbl y, vdbl x) { return atan2(y, x); }
#endif
// a literal of the reference's double arithmetic
#define VR(x) ((vdbl)(x))

#ifndef VIMG_FAST_LIBM
// glibc's float functions restated operation by operation (vimg_libm.cuh; every one checked against
// the host libm over all float inputs by oracle/tools/libm_check.cpp), one out-of-line copy each
VNI float g_logf(float x) { return vimg_libm::glibc_logf(x); }
VNI float g_log2f(float x) { return vimg_libm::glibc_log2f(x); }
VNI float g_acosf(float x) { return vimg_libm::glibc_acosf(x); }
VNI float g_atan2f(float y, float x) { return vimg_libm::glibc_atan2f(y, x); }
VNI float g_powf(float x, float y) { return vimg_libm::glibc_powf(x, y); }
VD float m_sinf(float x) { return vimg_libm::glibc_sinf(x); }  // bit-exact glibc sinf, |x| < 120
VD float m_cosf(float x) { return vimg_libm::glibc_cosf(x); }
VD float m_logf(float x) { return g_logf(x); }
VD float m_log2f(float x) { return g_log2f(x); }
VD float m_acosf(float x) { return g_acosf(x); }
VD float m_atan2f(float y, float x) { return g_atan2f(y, x); }
VD float m_powf(float x, float y) { return g_powf(x, y); }
#else
VD float m_sinf(float x) { return sinf(x); }
VD float m_cosf(float x) { return cosf(x); }
VD float m_logf(float x) { return logf(x); }
VD float m_log2f(float x) { return log2f(x); }
VD float m_acosf(float x) { return acosf(x); }
VD float m_atan2f(float y, float x) { return atan2f(y, x); }
VD float m_powf(float x, float y) { return powf(x, y); }
#endif

#define VIMG_PI_D VR(3.141592653589793238462643383279502884)  /* std::numbers::pi   */
#define VIMG_INV_PI_D VR(0.318309886183790671537767526745028724) /* std::numbers::inv_pi */
#define VIMG_INV_PI_F 0.318309886183790671537767526745028724f /* inv_pi_v<float>  */

VD bool is_inf(float x) { return fabsf(x) == CUDART_INF_F; }
VD bool is_nan(float x) { return x != x; }

}  // namespace vimg
