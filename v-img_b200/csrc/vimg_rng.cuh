// vimg_rng.cuh -- PCG32, dense random floats, R2 sequence and the shape samplers.
#pragma once
#include "vimg_types.cuh"

namespace vimg {

// pcg_state_setseq_64 (include/rng/pcg_rand.h:5-11)
struct Pcg32 {
  uint64_t state, inc;
};

// pcg32_random_r (include/rng/pcg_rand.h:15-21)
VD uint32_t pcg32_random(Pcg32& rng) {
  uint64_t oldstate = rng.state;
  rng.state = oldstate * 6364136223846793005ULL + rng.inc;
  uint32_t xorshifted = (uint32_t)(((oldstate >> 18u) ^ oldstate) >> 27u);
  uint32_t rot = (uint32_t)(oldstate >> 59u);
  return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31));
}

// pcg32_srandom_r (include/rng/pcg_rand.h:27-33)
VD void pcg32_seed(Pcg32& rng, uint64_t initstate, uint64_t initseq) {
  rng.state = 0U;
  rng.inc = (initseq << 1u) | 1u;
  pcg32_random(rng);
  rng.state += initstate;
  pcg32_random(rng);
}

// rand_float (include/rng/sampling.h:85-105): dense float in [0,1) from two draws
VD float rand_float(Pcg32& rng) {
  uint64_t r1 = pcg32_random(rng);
  uint64_t r2 = pcg32_random(rng);
  uint64_t u = (r1 << 32) | r2;
  uint32_t z = (uint32_t)__clzll((long long)u);
  if (z <= 40) {
    uint32_t e = 126 - z;
    uint32_t m = ((uint32_t)u) & 0x7fffff;
    return __uint_as_float(e << 23 | m);
  }
  return 0x1.0p-64f * (float)((uint32_t)u);
}

// random_x_y_r2 (include/rng/sampling.h:228-239).  `g` is a float constant, a1/a2
// are folded from double arithmetic on it and narrowed to float.
VD V2 random_x_y_r2(uint32_t n) {
  constexpr float g = 1.32471795724474602596;
  constexpr float a1 = 1.0 - (1.0 / g);
  constexpr float a2 = 1.0 - (1.0 / (g * g));
  float x = a1 * n;
  float y = a2 * n;
  return mk2(x - floorf(x), y - floorf(y));
}

// sample_disk (include/rng/sampling.h:15-22): phi computed in double, narrowed
VD V2 sample_disk(float rand1, float rand2) {
  float r = sqrtf(rand1);
  float phi = (float)(2.f * VIMG_PI_D * rand2);
  float x = r * m_cosf(phi);
  float y = r * m_sinf(phi);
  return mk2(x, y);
}

// sample_sphere (sampling.h:26-36): unqualified sqrt/cos/sin on float => double
VD V3 sample_sphere(float rand1, float rand2) {
  float phi = (float)(2 * VIMG_PI_D * rand1);
  float cos_theta = 2 * rand2 - 1;
  float sin_theta = (float)sqrt((vdbl)(1 - cos_theta * cos_theta));
  float x = (float)(d_cos((vdbl)phi) * sin_theta);
  float y = (float)(d_sin((vdbl)phi) * sin_theta);
  return mk3(x, y, cos_theta);
}

// sample_sphere_cap (sampling.h:40-51): sqrtf stays float, cos/sin are double
VD V3 sample_sphere_cap(float rand1, float rand2, float cos_theta_max) {
  float phi = (float)(2 * VIMG_PI_D * rand1);
  float cos_theta = std_lerp(cos_theta_max, 1.0f, rand2);
  float sin_theta = sqrtf(1 - cos_theta * cos_theta);
  float x = (float)(d_cos((vdbl)phi) * sin_theta);
  float y = (float)(d_sin((vdbl)phi) * sin_theta);
  return mk3(x, y, cos_theta);
}

// sample_hemisphere_cosine (sampling.h:69-79): std:: float functions
VD V3 sample_hemisphere_cosine(float rand1, float rand2) {
  float phi = (float)(2 * VIMG_PI_D * rand1);
  float cos_theta = sqrtf(rand2);
  float sin_theta = sqrtf(1 - cos_theta * cos_theta);
  float x = m_cosf(phi) * sin_theta;
  float y = m_sinf(phi) * sin_theta;
  return mk3(x, y, cos_theta);
}

// ArraySampling1D::sample (sampling.h:144-155): std::upper_bound over n+1 cdf entries
VD void sample_cdf_1d(const float* cdf, uint32_t n_plus_1, float u, uint32_t& index, float& du) {
  // first element > u
  uint32_t lo = 0, len = n_plus_1;
  while (len > 0) {
    uint32_t half = len >> 1;
    uint32_t mid = lo + half;
    if (u < cdf[mid]) {
      len = half;
    } else {
      lo = mid + 1;
      len = len - half - 1;
    }
  }
  index = lo - 1;
  du = u - cdf[index];
  float w = cdf[index + 1] - cdf[index];
  if (w > 0) du /= w;
}

}  // namespace vimg
