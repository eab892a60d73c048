// vimg_kernels.cuh -- the render kernels (replacement of scene_integrator<F>).
//
// include/integrators.h:101-137: for every pixel seed ONE PCG stream with the image
// index, then run the samples of that pixel in order (the stream carries over from
// sample to sample, so a pixel's samples are inherently sequential), sum in float and
// divide by spp.  One GPU thread owns one pixel; a warp owns an 8x4 pixel patch, two
// warps one of the reference's 8x8 tiles.  Tiles are numbered as the reference's
// work_list (x-major) and dealt round-robin to ranks: tile t belongs to rank t % n_ranks.
#pragma once
#include "vimg_integrators.cuh"

namespace vimg {

struct RenderArgs {
  uint32_t samples, depth;
  uint32_t rank, n_ranks;
  uint32_t tiles_x, tiles_y;
  uint32_t n_local_tiles;  // tiles owned by this rank
  int32_t single_x, single_y;  // >= 0: trace_pixel mode, one thread
};

VD void warp_add_counter(unsigned long long* dst, uint32_t v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, (unsigned long long)v);
}

}  // namespace vimg
#include "vimg_wavefront.cuh"
#include "vimg_queues.cuh"
namespace vimg {

// maps the calling thread to its pixel; returns false if it has none
VD bool thread_pixel(const RenderArgs& a, const DCamera& cam, uint32_t& x, uint32_t& y) {
  if (a.single_x >= 0) {
    x = (uint32_t)a.single_x;
    y = (uint32_t)a.single_y;
    return blockIdx.x == 0 && threadIdx.x == 0;
  }
  const uint32_t local_tile = blockIdx.x * (blockDim.x >> 6) + (threadIdx.x >> 6);
  if (local_tile >= a.n_local_tiles) return false;
  const uint32_t tile = a.rank + local_tile * a.n_ranks;
  const uint32_t tx = tile / a.tiles_y, ty = tile % a.tiles_y;
  const uint32_t in_tile = threadIdx.x & 63;
  x = tx * 8 + (in_tile & 7);
  y = ty * 8 + (in_tile >> 3);
  return x < (uint32_t)cam.res[0] && y < (uint32_t)cam.res[1];
}

template <int INTEGRATOR, bool COUNT>
__global__ void __launch_bounds__(128)
render_kernel(const __grid_constant__ DScene sc, const __grid_constant__ RenderArgs args,
              float* __restrict__ out, DCounters* __restrict__ counters) {
  uint32_t x, y;
  const bool active = thread_pixel(args, sc.cam, x, y);
  PathCounters pc = {};
  uint32_t nan_samples = 0, n_samples = 0;
  const uint32_t W = sc.cam.res[0], H = sc.cam.res[1];
  const size_t image_index = active ? x + ((size_t)(H - 1 - y) * W) : 0;
  const uint32_t image_seq_start = x + y;
  Pcg32 rng;
  pcg32_seed(rng, image_index, 0);
  V3 acc = mk3(0.0f);

  // camera ray of sample s (include/integrators.h:119-123); g++ evaluates the two rand_float
  // arguments right to left, so rand2 gets the first draw
  auto camera_ray = [&](uint32_t s) {
    V2 off = random_x_y_r2(image_seq_start + s);
    float rand2 = rand_float(rng);
    float rand1 = rand_float(rng);
    return generate_ray(sc.cam, (float)x + off.x, (float)y + off.y, rand1, rand2);
  };

  // the normal integrators trace one camera ray per sample; material / mis run in
  // wavefront_kernel (vimg_wavefront.cuh)
  if (active) {
    for (uint32_t s = 0; s < args.samples; s++) {
      Ray cam_ray = camera_ray(s);
      V3 col = normal_integrator<INTEGRATOR == VIMG_INTEGRATOR_G_NORMAL, COUNT>(sc, cam_ray, pc);
      if (is_nan(col.x) || is_nan(col.y) || is_nan(col.z)) nan_samples++;
      acc = acc + col;
      n_samples++;
    }
  }
  if (active) {
    acc = acc / (float)args.samples;
    float* o = out + (args.single_x >= 0 ? 0 : 3 * image_index);
    o[0] = acc.x;
    o[1] = acc.y;
    o[2] = acc.z;
  }
  warp_add_counter(&counters->rays_closest, pc.rays_closest);
  warp_add_counter(&counters->rays_shadow, pc.rays_shadow);
  warp_add_counter(&counters->nan_samples, nan_samples);
  warp_add_counter(&counters->samples, n_samples);
  if (COUNT) {
    warp_add_counter(&counters->inner_nodes, pc.trav.inner);
    warp_add_counter(&counters->leaf_nodes, pc.trav.leaf);
    warp_add_counter(&counters->tri_tests, pc.trav.tri);
    warp_add_counter(&counters->sphere_tests, pc.trav.sph);
  }
}

// turbo_colormap (src/integrators/heatmap.cpp:20-37): double literals narrowed to float by the
// glm constructors; glm::dot(vec4) = (x*x' + y*y') + (z*z' + w*w'), dot(vec2) = x*x' + y*y'
VD V3 turbo_colormap(float x) {
  const float r4[4] = {(float)0.13572138, (float)4.61539260, (float)-42.66032258, (float)132.13108234};
  const float g4[4] = {(float)0.09140261, (float)2.19418839, (float)4.84296658, (float)-14.18503333};
  const float b4[4] = {(float)0.10667330, (float)12.64194608, (float)-60.58204836, (float)110.36276771};
  const float r2[2] = {(float)-152.94239396, (float)59.28637943};
  const float g2[2] = {(float)4.27729857, (float)2.82956604};
  const float b2[2] = {(float)-89.90310912, (float)27.34824973};
  x = std_clamp(x, 0.f, 1.f);
  const float v4[4] = {1.0f, x, x * x, x * x * x};
  const float v2[2] = {v4[2] * v4[2], v4[3] * v4[2]};
  auto dot4 = [&](const float (&c)[4]) {
    float tx = v4[0] * c[0], ty = v4[1] * c[1], tz = v4[2] * c[2], tw = v4[3] * c[3];
    return (tx + ty) + (tz + tw);
  };
  auto dot2 = [&](const float (&c)[2]) { return v2[0] * c[0] + v2[1] * c[1]; };
  return mk3(dot4(r4) + dot2(r2), dot4(g4) + dot2(g2), dot4(b4) + dot2(b2));
}

// heatmap_img (src/integrators/heatmap.cpp:39-147): the integrators' camera rays, the cost
// BVH::hit<float> accumulates along the closest-hit traversal (bvh.h:128-131,159-163,189-192:
// 0.5 for the root box, 2 * 0.5 per inner node popped on the scalar path, 1 per primitive test --
// all dyadic, so the float sums are exact in any order), truncated mean over the samples,
// turbo colour map of mean / factor.  out_cost (nullable): the truncated mean, W*H floats.
__global__ void __launch_bounds__(128)
heatmap_kernel(const __grid_constant__ DScene sc, const __grid_constant__ RenderArgs args, const float factor,
               float* __restrict__ out, float* __restrict__ out_cost) {
  uint32_t x, y;
  if (!thread_pixel(args, sc.cam, x, y)) return;
  const uint32_t W = sc.cam.res[0], H = sc.cam.res[1];
  const size_t image_index = x + ((size_t)(H - 1 - y) * W);
  const uint32_t image_seq_start = x + y;
  Pcg32 rng;
  pcg32_seed(rng, image_index, 0);
  float acc = 0.f;
  for (uint32_t s = 0; s < args.samples; s++) {
    V2 off = random_x_y_r2(image_seq_start + s);
    float rand2 = rand_float(rng);
    float rand1 = rand_float(rng);
    Ray ray = generate_ray(sc.cam, (float)x + off.x, (float)y + off.y, rand1, rand2);
    PreHit ph;
    TravCounters tc = {};
    bvh_traverse<false, true>(sc, ray, ph, &tc);
    // an empty BVH returns 0 before the root test (bvh.h:100-102)
    const float cost = sc.n_nodes == 0 ? 0.f : 0.5f + (float)tc.inner + (float)(tc.tri + tc.sph);
    acc += cost;
  }
  const float hits = (float)(uint32_t)(acc / (float)args.samples);
  if (out_cost) out_cost[image_index] = hits;
  const V3 c = turbo_colormap(hits / factor);
  out[3 * image_index] = c.x;
  out[3 * image_index + 1] = c.y;
  out[3 * image_index + 2] = c.z;
}

// primary-hit primitive ids (parity hook): camera + closest-hit traversal only
__global__ void __launch_bounds__(128)
trace_ids_kernel(const __grid_constant__ DScene sc, const __grid_constant__ RenderArgs args,
                 int32_t* __restrict__ out_ids, float* __restrict__ out_t) {
  uint32_t x, y;
  if (!thread_pixel(args, sc.cam, x, y)) return;
  const uint32_t W = sc.cam.res[0], H = sc.cam.res[1];
  const size_t image_index = x + ((size_t)(H - 1 - y) * W);
  const uint32_t image_seq_start = x + y;
  Pcg32 rng;
  pcg32_seed(rng, image_index, 0);
  for (uint32_t s = 0; s < args.samples; s++) {
    V2 off = random_x_y_r2(image_seq_start + s);
    float rand2 = rand_float(rng);
    float rand1 = rand_float(rng);
    Ray ray = generate_ray(sc.cam, (float)x + off.x, (float)y + off.y, rand1, rand2);
    PreHit ph;
    bool hit = bvh_traverse<false, false>(sc, ray, ph, nullptr);
    out_ids[image_index * args.samples + s] = hit ? (int32_t)sc.leaf_prims[ph.leaf_slot].prim_id : -1;
    if (out_t) out_t[image_index * args.samples + s] = ray.maxT;
  }
}

}  // namespace vimg
