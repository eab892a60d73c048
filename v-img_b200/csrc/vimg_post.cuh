// vimg_post.cuh -- the post-process chain right after the render path:
// tonemapper, sRGB OETF and 8-bit quantisation of src/main.cpp:304-356 as ONE pass over the
// frame in HBM, so that what crosses PCIe is W*H*3 bytes instead of W*H*12.
//
//   simple_clamp              include/color_utils.h:21-25
//   aces                      src/tonemap/aces.cpp:5-31
//   agx (+ look, + eotf)      src/tonemap/agx.cpp:6-99
//   reinhard_lum              src/tonemap/reinhard.cpp:3-37 (needs the frame's largest luminance)
//   sRGB_gamma_correction     include/color_utils.h:50-68
//   NaN -> magenta, int(255.999 * v)   src/main.cpp:339-356
//
// Arithmetic: every glm expression written out per component in glm's order, no FMA contraction
// (the library is built with -fmad=false), IEEE division: the +, -, *, / part is bit-identical
// to the CPU.  std::pow / std::log2 on floats are glibc's powf / log2f there and the exact
// restatements of vimg_libm.cuh here, so the float image after gamma is bit-identical too.
// tests/test_gpu_post.py checks EVERY float input of the OETF's pow branch.
// HBM bound: 12 bytes read + 3 bytes written per pixel (+ 12 when the float image is kept).
#pragma once
#include "vimg_math.cuh"

namespace vimg {

VD float glm_clamp(float x, float lo, float hi) { return glm_min(glm_max(x, lo), hi); }

// luminance (color_utils.h:9-11)
VD float post_luminance(V3 c) { return dot(c, mk3(0.212671f, 0.715160f, 0.072169f)); }

// glm mat3 * vec3 (type_mat3x3.inl), m column-major
VD V3 post_mat3_mul(const float (&m)[9], V3 v) {
  return mk3(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
             m[2] * v.x + m[5] * v.y + m[8] * v.z);
}

VD float post_aces_fit(float v) {  // rtt_and_odt_fit, aces.cpp:5-9
  float a = v * (v + 0.0245786f) - 0.000090537f;
  float b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
  return a / b;
}
VD V3 post_aces(V3 v) {  // aces_fitted, aces.cpp:11-25
  const float in_m[9] = {0.59719f, 0.07600f, 0.02840f, 0.35458f, 0.90834f, 0.13383f, 0.04823f, 0.01566f, 0.83777f};
  const float out_m[9] = {1.60475f, -0.10208f, -0.00327f, -0.53108f, 1.10813f, -0.07276f,
                          -0.07367f, -0.00605f, 1.07602f};
  v = post_mat3_mul(in_m, v);
  v = mk3(post_aces_fit(v.x), post_aces_fit(v.y), post_aces_fit(v.z));
  return post_mat3_mul(out_m, v);
}

VD float post_agx_contrast(float x) {  // agxDefaultContrastApprox, agx.cpp:6-14
  float x2 = x * x, x4 = x2 * x2;
  return 15.5f * x4 * x2 - 40.14f * x4 * x + 31.96f * x4 - 6.868f * x2 * x + 0.4298f * x2 + 0.1191f * x
         - 0.00232f;
}
VD V3 post_agx(V3 val) {  // agx + agxLook + agxEotf, agx.cpp:16-99
  const float agx_mat[9] = {(float)0.842479062253094, (float)0.0423282422610123, (float)0.0423756549057051,
                            (float)0.0784335999999992, (float)0.878468636469772, (float)0.0784336,
                            (float)0.0792237451477643, (float)0.0791661274605434, (float)0.879142973793104};
  const float agx_mat_inv[9] = {(float)1.19687900512017, (float)-0.0528968517574562, (float)-0.0529716355144438,
                                (float)-0.0980208811401368, (float)1.15190312990417, (float)-0.0980434501171241,
                                (float)-0.0990297440797205, (float)-0.0989611768448433, (float)1.15107367264116};
  const float min_ev = -12.47393f, max_ev = 4.026069f;
  val = post_mat3_mul(agx_mat, val);
  val = mk3(glm_clamp(m_log2f(val.x), min_ev, max_ev), glm_clamp(m_log2f(val.y), min_ev, max_ev),
            glm_clamp(m_log2f(val.z), min_ev, max_ev));
  val = mk3((val.x - min_ev) / (max_ev - min_ev), (val.y - min_ev) / (max_ev - min_ev),
            (val.z - min_ev) / (max_ev - min_ev));
  val = mk3(post_agx_contrast(val.x), post_agx_contrast(val.y), post_agx_contrast(val.z));
  // agxLook: slope 1, offset 0, power 1, sat 1: pow(x * 1 + 0, 1) = x * 1 + 0
  val = mk3(val.x * 1.0f + 0.0f, val.y * 1.0f + 0.0f, val.z * 1.0f + 0.0f);
  const float luma = post_luminance(val);
  val = mk3(luma + 1.0f * (val.x - luma), luma + 1.0f * (val.y - luma), luma + 1.0f * (val.z - luma));
  val = post_mat3_mul(agx_mat_inv, val);
  if (val.x < 0.f) val.x = 0.f;
  if (val.y < 0.f) val.y = 0.f;
  if (val.z < 0.f) val.z = 0.f;
  return mk3(m_powf(val.x, 2.2f), m_powf(val.y, 2.2f), m_powf(val.z, 2.2f));
}

VD float post_srgb_oetf(float p) {  // color_utils.h:50-68
  p = glm_clamp(p, 0.0f, 1.0f);
  if (p < 0.0031308f) return p * 12.92f;
  return 1.055f * m_powf(p, 1.0f / 2.4f) - 0.055f;
}

VD uint32_t post_quantise(float v) {  // main.cpp:348-351
  // static_cast<int>(double): x86 cvttsd2si gives INT_MIN out of range, then std::clamp -> 0;
  // after the OETF v is in [0, 1] or NaN (NaN pixels are handled by the caller)
  const double d = 255.999 * (double)v;
  int i = (d >= 2147483648.0 || d < -2147483648.0 || d != d) ? (int)0x80000000 : (int)d;
  return (uint32_t)(i < 0 ? 0 : (i > 255 ? 255 : i));
}

// Byte path without pow: byte(p) = quantise(oetf(p)) is a non-decreasing step function of the
// clamped input p in [0, 1], so it is fully described by 256 thresholds T[b] = the smallest float
// whose byte is >= b.  They are found once per device by bisection over the float bit patterns
// WITH the functions above, so the byte of every float is by construction the one the pow path
// gives (which tests/test_gpu_post.py compares with the CPU's for every float of the pow branch);
// a pixel then costs three table look-ups in shared memory (an approximate estimate of the byte,
// settled by the exact thresholds) instead of three exact pow evaluations.
__global__ void post_build_thresholds_kernel(uint32_t* __restrict__ thr) {
  const uint32_t b = threadIdx.x;  // 256 threads
  uint32_t lo = 0u, hi = 0x3f800000u;  // bit patterns of 0.0f .. 1.0f order like the values
  if (b == 0) {
    thr[0] = 0u;
    return;
  }
  if (post_quantise(post_srgb_oetf(__uint_as_float(hi))) < b) {  // never reached (byte(1.0) = 255)
    thr[b] = 0xffffffffu;
    return;
  }
  while (lo < hi) {  // invariant: byte(hi) >= b
    const uint32_t mid = lo + (hi - lo) / 2u;
    if (post_quantise(post_srgb_oetf(__uint_as_float(mid))) >= b) hi = mid;
    else lo = mid + 1u;
  }
  thr[b] = hi;
}
VD uint32_t post_byte_from_thresholds(const uint32_t* s_thr, float c) {
  // glm::clamp(c, 0, 1) lets -0.0 through; + 0.0f turns it into +0.0 so that the bits order
  const float p = glm_clamp(c, 0.0f, 1.0f) + 0.0f;
  const uint32_t bits = __float_as_uint(p);
  // The byte is the largest b with T[b] <= bits.  Start from an ESTIMATE -- the same chain with the
  // hardware's approximate exp2 / log2 (two MUFU operations, ~1e-6 relative) -- and let the exact
  // table settle it: nearly always two shared-memory reads instead of an 8-step search.  Whatever
  // the estimate is, the loops end at the same b the search finds (T is non-decreasing, T[0] = 0,
  // T[256] = 0xffffffff).
  const float est = p < 0.0031308f ? p * 12.92f : 1.055f * __powf(p, 1.0f / 2.4f) - 0.055f;
  int b = (int)(255.999f * est);
  b = b < 0 ? 0 : (b > 255 ? 255 : b);
  while (s_thr[b] > bits) --b;
  while (s_thr[b + 1] <= bits) ++b;
  return (uint32_t)b;
}

// largest luminance of the frame (reinhard.cpp:3-13): `if (curr > largest)` from 0, so NaN and
// negative luminances never win and the result is a non-negative float, whose bit pattern
// orders like the value: atomicMax on the bits
__global__ void post_max_luminance_kernel(const float* __restrict__ rgb, uint64_t n_pixels,
                                          uint32_t* __restrict__ max_bits) {
  float m = 0.0f;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_pixels; i += (uint64_t)gridDim.x * blockDim.x) {
    const float l = post_luminance(mk3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
    if (l > m) m = l;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_down_sync(0xffffffffu, m, o);
    if (other > m) m = other;
  }
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(max_bits, __float_as_uint(m));
}

template <int TM>
VD V3 post_tonemap(V3 c, float largest_L) {
  if (TM == VIMG_TONEMAP_CLAMP) {
    c = mk3(glm_clamp(c.x, 0.f, 1.f), glm_clamp(c.y, 0.f, 1.f), glm_clamp(c.z, 0.f, 1.f));
  } else if (TM == VIMG_TONEMAP_AGX) {
    c = post_agx(c);
  } else if (TM == VIMG_TONEMAP_REINHARD) {  // reinhard.cpp:16-37
    const float in_L = post_luminance(c);
    const float numerator = in_L * (1.0f + (in_L / (largest_L * largest_L)));
    const float new_L = numerator / (1.0f + in_L);
    c = in_L > 0.f ? c * (new_L / in_L) : mk3(0.f);
  } else if (TM == VIMG_TONEMAP_ACES) {
    c = post_aces(c);
  }
  return c;
}

// bytes only: the OETF and the quantisation through the threshold table (NaN passes through
// glm's clamp and the OETF, so a NaN channel after the tonemapper is a magenta pixel)
template <int TM>
VD void post_pixel_bytes(V3 c, float largest_L, const uint32_t* s_thr, uint32_t& r8, uint32_t& g8, uint32_t& b8) {
  c = post_tonemap<TM>(c, largest_L);
  if (is_nan(c.x) || is_nan(c.y) || is_nan(c.z)) {
    r8 = 255;
    g8 = 0;
    b8 = 255;
  } else {
    r8 = post_byte_from_thresholds(s_thr, c.x);
    g8 = post_byte_from_thresholds(s_thr, c.y);
    b8 = post_byte_from_thresholds(s_thr, c.z);
  }
}

template <int TM>
VD void post_pixel(V3 c, float largest_L, V3& srgb, uint32_t& r8, uint32_t& g8, uint32_t& b8) {
  c = post_tonemap<TM>(c, largest_L);
  srgb = mk3(post_srgb_oetf(c.x), post_srgb_oetf(c.y), post_srgb_oetf(c.z));
  if (is_nan(srgb.x) || is_nan(srgb.y) || is_nan(srgb.z)) {  // main.cpp:342-346
    r8 = 255;
    g8 = 0;
    b8 = 255;
  } else {
    r8 = post_quantise(srgb.x);
    g8 = post_quantise(srgb.y);
    b8 = post_quantise(srgb.z);
  }
}

// thread <-> 4 consecutive pixels: 3 x 16-byte loads, 3 x 4-byte stores of the packed bytes
// (+ 3 x 16-byte stores when the float image is wanted)
template <int TM>
__global__ void __launch_bounds__(256) post_kernel(const float* __restrict__ rgb, uint64_t n_pixels,
                                                   const uint32_t* __restrict__ max_bits,
                                                   const uint32_t* __restrict__ thresholds,
                                                   float* __restrict__ out_srgb, uint8_t* __restrict__ out_rgb8) {
  __shared__ uint32_t s_thr[257];
  const bool bytes_only = out_srgb == nullptr && thresholds != nullptr;
  if (bytes_only) {
    s_thr[threadIdx.x] = thresholds[threadIdx.x];  // blockDim.x == 256
    if (threadIdx.x == 0) s_thr[256] = 0xffffffffu;
    __syncthreads();
  }
  const float largest_L = TM == VIMG_TONEMAP_REINHARD ? __uint_as_float(*max_bits) : 0.f;
  const uint64_t n_quads = n_pixels / 4;
  for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < n_quads; q += (uint64_t)gridDim.x * blockDim.x) {
    const float4* src = reinterpret_cast<const float4*>(rgb) + 3 * q;
    const float4 a = __ldcs(src), b = __ldcs(src + 1), c = __ldcs(src + 2);
    const float in[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    float of[12];
    uint32_t by[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (bytes_only) {
        post_pixel_bytes<TM>(mk3(in[3 * k], in[3 * k + 1], in[3 * k + 2]), largest_L, s_thr, by[3 * k], by[3 * k + 1],
                             by[3 * k + 2]);
        continue;
      }
      V3 s;
      post_pixel<TM>(mk3(in[3 * k], in[3 * k + 1], in[3 * k + 2]), largest_L, s, by[3 * k], by[3 * k + 1], by[3 * k + 2]);
      of[3 * k] = s.x;
      of[3 * k + 1] = s.y;
      of[3 * k + 2] = s.z;
    }
    if (out_srgb) {
      float4* dst = reinterpret_cast<float4*>(out_srgb) + 3 * q;
      __stcs(dst, make_float4(of[0], of[1], of[2], of[3]));
      __stcs(dst + 1, make_float4(of[4], of[5], of[6], of[7]));
      __stcs(dst + 2, make_float4(of[8], of[9], of[10], of[11]));
    }
    if (out_rgb8) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(out_rgb8) + 3 * q;
#pragma unroll
      for (int w = 0; w < 3; ++w)
        __stcs(dst + w, by[4 * w] | (by[4 * w + 1] << 8) | (by[4 * w + 2] << 16) | (by[4 * w + 3] << 24));
    }
  }
  // the last n_pixels % 4 pixels
  const uint64_t tail = n_quads * 4 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (tail < n_pixels) {
    V3 s = mk3(0.f);
    uint32_t r8, g8, b8;
    if (bytes_only)
      post_pixel_bytes<TM>(mk3(rgb[3 * tail], rgb[3 * tail + 1], rgb[3 * tail + 2]), largest_L, s_thr, r8, g8, b8);
    else
      post_pixel<TM>(mk3(rgb[3 * tail], rgb[3 * tail + 1], rgb[3 * tail + 2]), largest_L, s, r8, g8, b8);
    if (out_srgb) {
      out_srgb[3 * tail] = s.x;
      out_srgb[3 * tail + 1] = s.y;
      out_srgb[3 * tail + 2] = s.z;
    }
    if (out_rgb8) {
      out_rgb8[3 * tail] = (uint8_t)r8;
      out_rgb8[3 * tail + 1] = (uint8_t)g8;
      out_rgb8[3 * tail + 2] = (uint8_t)b8;
    }
  }
}

}  // namespace vimg
