// vimg_prep.cuh -- scene-preparation kernels, the step BEFORE the render path (SURVEY.md 8(f)
// rank 3): what the reference's constructors compute on the host in seconds per 2K texture.
//
//   mip chain        ImageTexture::ImageTexture (src/image_texture.cpp:59-130): every level is
//                    the previous one through Wronski's 8-tap downsampling filter, each tap a
//                    bilinear fetch with col_at_uv_mipmap (:132-161), negative results set to 0
//   env-map CDFs     ArraySampling2D / ArraySampling1D (include/rng/sampling.h:114-135,167-197):
//                    luminance * sin(elevation) per pixel, one running float sum per row, the
//                    row integrals summed the same way
//   byte conversions ImageTexture::convert_sRGB_to_linear / convert_RGB_to_normal
//                    (src/image_texture.cpp:257-275) on the decoded 8-bit image
//
// Both must be bit-compatible with the host versions or the parity of everything that reads a
// texture moves.  The filter is float +, *, and the bilinear fetch the render kernels already
// use.  The CDFs are SEQUENTIAL float sums (cdf[x] = cdf[x-1] + |f|): float addition is not
// associative, so the additions of a row are made by one lane in the reference's order (a warp per
// row loads and evaluates |f| in parallel, H rows in parallel), and the sin(elevation) factors --
// H values of std::sin(double) -- are computed on the host with the host libm, which is what the
// reference calls.
#pragma once
#include "vimg_hitinfo.cuh"

namespace vimg {

// ImageTexture::col_at_uv_mipmap on a bare level (same expressions as col_at_uv_mipmap above)
VD V3 bilinear_level(const float* __restrict__ lvl, uint32_t mip_w, uint32_t mip_h, uint32_t wrap_u,
                     uint32_t wrap_v, V2 uv) {
  float pixel_u = handle_wrapping(uv.x, wrap_u) * mip_w;
  float pixel_v = handle_wrapping(uv.y, wrap_v) * mip_h;
  int curr_x = std_clampi((int)pixel_u, 0, (int)mip_w - 1);
  int curr_y = std_clampi((int)pixel_v, 0, (int)mip_h - 1);
  int next_x = std_clampi(curr_x + 1, 0, (int)mip_w - 1);
  int next_y = std_clampi(curr_y + 1, 0, (int)mip_h - 1);
  float x_fraction = pixel_u - curr_x;
  float y_fraction = pixel_v - curr_y;
  V3 x0 = load3(lvl, curr_x + curr_y * mip_w), x1 = load3(lvl, next_x + curr_y * mip_w);
  V3 a = mix(x0, x1, x_fraction);
  V3 y0 = load3(lvl, curr_x + next_y * mip_w), y1 = load3(lvl, next_x + next_y * mip_w);
  V3 b = mix(y0, y1, x_fraction);
  return mix(a, b, y_fraction);
}

// one mip level from the previous one (image_texture.cpp:84-122); thread <-> output pixel
__global__ void __launch_bounds__(256)
prep_mip_level_kernel(const float* __restrict__ prev, uint32_t prev_w, uint32_t prev_h, float* __restrict__ next,
                      uint32_t next_w, uint32_t next_h, uint32_t wrap_u, uint32_t wrap_v) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= next_w || y >= next_h) return;
  const V2 inv_res = mk2(1.f / prev_w, 1.f / prev_h);
  const V2 curr_uv = mk2((float)(2 * x), (float)(2 * y)) * inv_res;
  // glm::vec2(-0.75777, ...) / (2.907, 0.0): double literals narrowed to float
  const float k = (float)0.75777, f = (float)2.907;
  const float tap[8][2] = {{-k, -k}, {k, -k}, {k, k}, {-k, k}, {-f, 0.f}, {f, 0.f}, {0.f, -f}, {0.f, f}};
  V3 sum = mk3(0.f);
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const float wgt = t < 4 ? 0.37487566f : -0.12487566f;
    const V2 uv = curr_uv + mk2(tap[t][0], tap[t][1]) * inv_res;
    sum = sum + wgt * bilinear_level(prev, prev_w, prev_h, wrap_u, wrap_v, uv);
  }
  if (sum.x < 0) sum.x = 0.f;
  if (sum.y < 0) sum.y = 0.f;
  if (sum.z < 0) sum.z = 0.f;
  float* o = next + 3 * ((size_t)x + (size_t)y * next_w);
  o[0] = sum.x;
  o[1] = sum.y;
  o[2] = sum.z;
}

// ArraySampling2D rows (sampling.h:176-193) + ArraySampling1D (:118-135): WARP <-> image row.
// The running sum cdf[x + 1] = cdf[x] + |f(x)| is sequential by contract (float addition is not
// associative), but only the additions are: the 32 lanes load a chunk of the row (coalesced) and
// evaluate |luminance * sin(elevation)| into shared memory, lane 0 turns the chunk into running sums
// in the reference's order (one dependent FADD per element, operands already on chip), and all lanes
// store the chunk and later normalise the row.  One thread per row did all of it with 24 KB-strided
// loads: 1.6 ms for a 2048x1024 map; this is bound by the 2048-add chain of one row instead.
#define PREP_CDF_CHUNK 2048
__global__ void __launch_bounds__(32)
prep_row_cdf_kernel(const float* __restrict__ image, uint32_t width, uint32_t height,
                    const float* __restrict__ sin_elevation, float* __restrict__ col_cdfs,
                    float* __restrict__ row_integrals) {
  __shared__ float s_v[PREP_CDF_CHUNK];
  const uint32_t y = blockIdx.x, lane = threadIdx.x;
  if (y >= height) return;
  const float se = sin_elevation[y];
  float* cdf = col_cdfs + (size_t)y * (width + 1);
  const float* row = image + 3 * (size_t)y * width;
  float run = 0.f;  // meaningful on lane 0
  if (lane == 0) cdf[0] = 0.f;
  for (uint32_t base = 0; base < width; base += PREP_CDF_CHUNK) {
    const uint32_t n = width - base < PREP_CDF_CHUNK ? width - base : PREP_CDF_CHUNK;
    for (uint32_t i = lane; i < n; i += 32) {
      const float* px = row + 3 * (size_t)(base + i);
      const float lum = dot(mk3(px[0], px[1], px[2]), mk3(0.212671f, 0.715160f, 0.072169f)) * se;
      s_v[i] = fabsf(lum);
    }
    __syncwarp();
    if (lane == 0) {
      for (uint32_t i = 0; i < n; ++i) {
        run = run + s_v[i];
        s_v[i] = run;
      }
    }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) cdf[base + i + 1] = s_v[i];
    __syncwarp();
  }
  const float func_int = __shfl_sync(0xffffffffu, run, 0);
  if (lane == 0) row_integrals[y] = func_int;
  __syncwarp();  // this warp's own stores to cdf[] are visible to it below
  if (func_int == 0.f) {
    for (uint32_t i = lane; i <= width; i += 32) cdf[i] = (float)i / (float)width;
  } else {
    for (uint32_t i = lane; i <= width; i += 32) cdf[i] = cdf[i] / func_int;
  }
}

// row_probabilities = ArraySampling1D(row integrals): H values, ONE warp, the reference's order
// (same chunk scheme: coalesced loads into shared memory, lane 0 adds)
__global__ void __launch_bounds__(32)
prep_row_prob_kernel(const float* __restrict__ row_integrals, uint32_t height, float* __restrict__ row_cdf) {
  __shared__ float s_v[PREP_CDF_CHUNK];
  if (blockIdx.x != 0) return;
  const uint32_t lane = threadIdx.x;
  float run = 0.f;
  if (lane == 0) row_cdf[0] = 0.f;
  for (uint32_t base = 0; base < height; base += PREP_CDF_CHUNK) {
    const uint32_t n = height - base < PREP_CDF_CHUNK ? height - base : PREP_CDF_CHUNK;
    for (uint32_t i = lane; i < n; i += 32) s_v[i] = fabsf(row_integrals[base + i]);
    __syncwarp();
    if (lane == 0) {
      for (uint32_t i = 0; i < n; ++i) {
        run = run + s_v[i];
        s_v[i] = run;
      }
    }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) row_cdf[base + i + 1] = s_v[i];
    __syncwarp();
  }
  run = __shfl_sync(0xffffffffu, run, 0);
  if (run == 0.f) {
    for (uint32_t i = lane; i <= height; i += 32) row_cdf[i] = (float)i / (float)height;
  } else {
    for (uint32_t i = lane; i <= height; i += 32) row_cdf[i] = row_cdf[i] / run;
  }
}

// ---- glTF texture conversions (src/scene_loading/gltf_loading.cpp:133-166) ----------------------
// make_texture turns the decoded 8-bit image into floats (:139-143) and calls
//   ImageTexture::convert_sRGB_to_linear (image_texture.cpp:257-263): p / 255.f, then
//       pix_sRGB_to_linear (color_utils.h:28-47) = p / 12.92f or powf((p + 0.055f) / 1.055f, 2.4f)
//   ImageTexture::convert_RGB_to_normal  (image_texture.cpp:265-275): p / 127.5f - 1, x and y
//       times scale, glm::normalize
// Here the BYTES go to the device (3 B per pixel over PCIe instead of 12) and the floats are
// produced there.  HBM bound: 3 bytes read + 12 written per pixel.

// the value of one colour byte; powf is glibc's, restated exactly (vimg_libm.cuh)
VD float prep_srgb_byte_to_linear(uint32_t b) {
  float p = (float)b;
  p = p / 255.f;
  if (p <= 0.04045f) return p / 12.92f;
  return vimg_libm::glibc_powf((p + 0.055f) / 1.055f, 2.4f);
}

// A channel has 256 possible inputs: every block evaluates the function once per byte value into
// shared memory (one powf per thread), then streams: thread <-> 16 consecutive channel bytes
// (one 16-byte load, four 16-byte stores).  n = number of channel bytes (3 * pixels).
__global__ void __launch_bounds__(256)
prep_srgb8_to_linear_kernel(const uint8_t* __restrict__ in, uint64_t n, float* __restrict__ out) {
  __shared__ float lut[256];
  lut[threadIdx.x] = prep_srgb_byte_to_linear(threadIdx.x);
  __syncthreads();
  const uint64_t n16 = n / 16, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float4* o = reinterpret_cast<float4*>(out) + 4 * i;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = make_float4(lut[w[k] & 255u], lut[(w[k] >> 8) & 255u], lut[(w[k] >> 16) & 255u], lut[w[k] >> 24]);
  }
  // tail: fewer than 16 bytes
  const uint64_t t = n16 * 16 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = lut[in[t]];
}

// thread <-> 4 pixels = 12 bytes in (three 4-byte loads), 48 bytes out (three 16-byte stores)
VD V3 prep_normal_of_bytes(uint32_t r, uint32_t g, uint32_t b, float scale) {
  V3 n = mk3((float)r / 127.5f - 1.f, (float)g / 127.5f - 1.f, (float)b / 127.5f - 1.f);
  n.x *= scale;
  n.y *= scale;
  return normalize(n);
}
__global__ void __launch_bounds__(256)
prep_rgb8_to_normal_kernel(const uint8_t* __restrict__ in, uint64_t n_pixels, float scale, float* __restrict__ out) {
  const uint64_t n4 = n_pixels / 4, stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t i = tid; i < n4; i += stride) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(in) + 3 * i;
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    const V3 a = prep_normal_of_bytes(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, scale);
    const V3 b = prep_normal_of_bytes(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u, scale);
    const V3 c = prep_normal_of_bytes((w1 >> 16) & 255u, w1 >> 24, w2 & 255u, scale);
    const V3 d = prep_normal_of_bytes((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24, scale);
    float4* o = reinterpret_cast<float4*>(out) + 3 * i;
    o[0] = make_float4(a.x, a.y, a.z, b.x);
    o[1] = make_float4(b.y, b.z, c.x, c.y);
    o[2] = make_float4(c.z, d.x, d.y, d.z);
  }
  const uint64_t t = n4 * 4 + tid;  // tail: fewer than 4 pixels
  if (t < n_pixels) {
    const V3 a = prep_normal_of_bytes(in[3 * t], in[3 * t + 1], in[3 * t + 2], scale);
    out[3 * t] = a.x;
    out[3 * t + 1] = a.y;
    out[3 * t + 2] = a.z;
  }
}

// parity hook (vimg_libm_eval): the DEVICE copies of glibc's float functions on caller-given inputs
__global__ void __launch_bounds__(256)
libm_eval_kernel(uint32_t fn, const float* __restrict__ x, const float* __restrict__ y, uint64_t n, float* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = x[i], b = y ? y[i] : 0.f;
    float r;
    switch (fn) {
      case 0: r = m_sinf(a); break;
      case 1: r = m_cosf(a); break;
      case 2: r = m_logf(a); break;
      case 3: r = m_log2f(a); break;
      case 4: r = m_powf(a, b); break;
      case 5: r = m_acosf(a); break;
      case 6: r = m_atan2f(a, b); break;
      case 7: r = vimg_libm::glibc_atanf(a); break;
      // the DOUBLE log2 of the widened input: low / high word of the result's bit pattern
      case 8: r = __uint_as_float((uint32_t)vimg_libm::d2u(d_log2((double)a))); break;
      default: r = __uint_as_float((uint32_t)(vimg_libm::d2u(d_log2((double)a)) >> 32)); break;
    }
    out[i] = r;
  }
}

}  // namespace vimg
