// vimg_wavefront.cuh -- persistent, block-local wavefront path tracer (material / mis).
//
// Measured (profiles/r1_*): a one-thread-per-pixel megakernel runs the Disney scene with
// ~6 of 32 lanes active -- a few lanes in the (long) Principled code serialise against many
// lanes in the (short) Lambertian code -- and thrashes the instruction cache.  This kernel
// keeps the reference's per-pixel sequential RNG stream (one path in flight per pixel,
// include/integrators.h:116-131) but regroups paths between stages:
//
//   * a block owns WF_SLOTS path slots (more than it has threads) whose whole state lives in
//     SHARED MEMORY (SoA);
//   * slot stage (thread i works on slots i, i + WF_THREADS, ...): pending shadow ray -> finish sample / pull the
//     next pixel / camera ray -> closest-hit traversal -> resolve (miss, emitter, Russian
//     roulette, depth) -> classify the hit material;
//   * block barrier; the live slots were bucketed by material class with warp ballots and
//     shared-memory counters, so that each warp of the next stage sees ONE material;
//   * shade stage (thread k works on the k-th slot of the bucketed list): light sampling,
//     BSDF sampling, both BSDF evaluations, next ray, shadow-ray request;
//   * block barrier, repeat.  Slots whose pixel is complete pull the next pixel of the frame
//     from a global counter (tile-major order), so there is no per-block tail.
//
// Every float operation and every RNG draw happens in the reference's order for its path
// (src/integrators/mis_integrator.cpp:18-189, mat_integrator.cpp:4-85), so the images are
// unchanged bit for bit (tests/test_gpu_parity.py).
#pragma once
#include "vimg_integrators.cuh"

namespace vimg {

#ifndef WF_THREADS
#define WF_THREADS 512
#endif
#ifndef WF_SLOTS
#define WF_SLOTS 1024  // path slots per block (a multiple of WF_THREADS): more slots than threads
#endif                // so that the shade stage has enough items of every class to fill all warps

enum : uint32_t {
  WF_ALIVE = 1u,       // slot owns a pixel with samples left
  WF_IN_PATH = 2u,     // a path is in flight (else: needs a camera ray)
  WF_FIRST = 4u,       // the ray in flight is the camera ray
  WF_NSB = 8u,         // non_specular_bounce
  WF_NEE = 16u,        // shadow ray pending (mis)
  WF_NEE_VALID = 32u,  // ... and its contribution passed the pdf checks
  WF_END = 64u,        // path ends before the next trace (result: L for mis, nee_c for material)
};

enum : uint32_t { WF_CLASS_LAMBERT = 0, WF_CLASS_PRINCIPLED = 1, WF_CLASS_DIELECTRIC = 2, WF_NUM_CLASSES = 3, WF_CLASS_NONE = 3 };

struct WfShared {
  float ray_o[3][WF_SLOTS], ray_d[3][WF_SLOTS], cone_w[WF_SLOTS], cone_s[WF_SLOTS];
  float thr[3][WF_SLOTS];
  float eta_scale[WF_SLOTS], last_pdf[WF_SLOTS];
  unsigned long long rng[WF_SLOTS];
  uint32_t flags[WF_SLOTS], d[WF_SLOTS];
  int32_t hit_slot[WF_SLOTS];
  float hit_t[WF_SLOTS], e0[WF_SLOTS], e1[WF_SLOTS], e2[WF_SLOTS], inv_det[WF_SLOTS];
  float nee_wi[3][WF_SLOTS], nee_dist[WF_SLOTS], nee_c[3][WF_SLOTS];
  float acc[3][WF_SLOTS], L[3][WF_SLOTS];  // pixel accumulator, bounce_result of the path in flight
  uint32_t s[WF_SLOTS], pixel[WF_SLOTS];   // sample counter, local pixel number
  uint16_t list[WF_NUM_CLASSES][WF_SLOTS];
  uint32_t class_count[2][WF_NUM_CLASSES];
};

// n / d for a divisor that is fixed for the launch, exact for every 32-bit n (Granlund-Montgomery: multiply-high
// with a 33-bit reciprocal split into a 32-bit multiplier and two shifts): 4 instructions where the hardware-less
// integer division is ~25.  The END kernels divide three times per item (pixel -> tile -> x, y; unit -> pixel).
struct FastDiv {
  uint32_t mul, sh1, sh2, d;
};
#ifdef __CUDACC__
VD uint32_t fast_div(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(n, f.mul);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}
#endif
inline FastDiv make_fast_div(uint32_t d) {  // d >= 1
  FastDiv f;
  uint32_t l = 0;
  while (l < 32 && (1ull << l) < d) ++l;  // ceil(log2 d)
  f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.sh1 = l < 1 ? l : 1;
  f.sh2 = l > 1 ? l - 1 : 0;
  f.d = d;
  return f;
}

struct WfArgs {
  uint32_t samples, depth;
  uint32_t rank, n_ranks;
  uint32_t tiles_x, tiles_y;
  uint32_t n_local_pixels;     // 64 * tiles owned by this rank
  int32_t single_x, single_y;  // >= 0: trace_pixel mode
  // queue wavefront only: sample ranges per pixel (vimg_render_params.spp_streams, >= 1), the
  // work units of this rank (n_local_pixels * streams; unit u = range u % streams of local pixel
  // u / streams) and, for streams > 1, the partial sums [image_index][range] x rgb
  uint32_t streams, n_local_units;
  float* partial;
  // sample-range partition over ranks (VIMG_PARTITION_SPP): this rank renders the ranges s_rank, s_rank +
  // s_n_ranks, ... of EVERY pixel -- my_streams of them; unit u = range s_rank + (u % my_streams) * s_n_ranks
  // of local pixel u / my_streams.  Tile partition: s_rank 0, s_n_ranks 1, my_streams = streams.
  uint32_t s_rank, s_n_ranks, my_streams;
  // 1: every traversal starts from an empty stack (VIMG_FLAG_FRESH_STACK, or streams > 1); 0: a path's
  // closest-hit traversal inherits what its shadow ray's traversal left (the reference's thread_stack)
  uint32_t fresh_stack;
  // the launch's fixed divisors (tiles_y, my_streams, streams) and, when a range boundary fits 32 bits
  // (samples < 65536: range <= 65535), the switch to divide it that way
  FastDiv div_tiles_y, div_my_streams, div_streams;
  uint32_t samples_small;
};
// first sample of sample range r: floor(r * samples / streams) (r = streams: one past the last)
#ifdef __CUDACC__
VD uint32_t wf_range_begin(const WfArgs& a, uint32_t r) {
  if (a.samples_small) return fast_div(r * a.samples, a.div_streams);
  return (uint32_t)(((uint64_t)r * a.samples) / a.streams);
}
#endif
// a slot's sample range number lives in the upper half of its flag word
#define WF_RANGE_SHIFT 16
#define WF_FLAG_MASK 0xffffu

VD V3 ld3(const float (*a)[WF_SLOTS], int i) { return mk3(a[0][i], a[1][i], a[2][i]); }
VD void st3(float (*a)[WF_SLOTS], int i, V3 v) {
  a[0][i] = v.x;
  a[1][i] = v.y;
  a[2][i] = v.z;
}

// local pixel number (tile-major within this rank's tile list) -> image coordinates
VD bool wf_pixel_xy(const WfArgs& a, const DCamera& cam, uint32_t p, uint32_t& x, uint32_t& y) {
  if (a.single_x >= 0) {
    x = (uint32_t)a.single_x;
    y = (uint32_t)a.single_y;
    return true;
  }
  const uint32_t tile = a.rank + (p >> 6) * a.n_ranks;
  const uint32_t tx = fast_div(tile, a.div_tiles_y), ty = tile - tx * a.tiles_y;
  x = tx * 8 + (p & 7);
  y = ty * 8 + ((p >> 3) & 7);
  return x < (uint32_t)cam.res[0] && y < (uint32_t)cam.res[1];
}

VD uint32_t wf_leaf_material(const DScene& sc, int32_t leaf_slot) { return sc.shade_recs[leaf_slot].material; }

template <bool NEED_FRAME = true>
VD HitInfo wf_hit_info(const DScene& sc, const DLeafPrim& lp, const DShadeRec& sr, const PreHit& ph, const Ray& ray) {
  if (lp.flags & VIMG_LEAF_SPHERE) return sphere_hit_info(lp, sr, ray);
  return triangle_hit_info<NEED_FRAME>(sc, lp, sr, ph);
}
VD HitInfo wf_hit_info(const DScene& sc, int32_t leaf_slot, const PreHit& ph, const Ray& ray) {
  const DLeafPrim lp = sc.leaf_prims[leaf_slot];
  const DShadeRec sr = load_shade_rec(sc, leaf_slot);
  return wf_hit_info<true>(sc, lp, sr, ph, ray);
}

VD uint32_t wf_material_class(uint32_t mat_type) {
  switch (mat_type) {
    case VIMG_MAT_LAMBERTIAN: return WF_CLASS_LAMBERT;
    case VIMG_MAT_PRINCIPLED: return WF_CLASS_PRINCIPLED;
    case VIMG_MAT_DIELECTRIC: return WF_CLASS_DIELECTRIC;
    default: return WF_CLASS_NONE;
  }
}

#ifdef WF_STAGE_CLOCKS
#define WF_CLK(k)                                  \
  do {                                             \
    long long now__ = clock64();                   \
    clk[k] += (unsigned long long)(now__ - t_prev); \
    t_prev = now__;                                \
  } while (0)
#else
#define WF_CLK(k) do {} while (0)
#endif

template <int INTEGRATOR, bool COUNT>
__global__ void __launch_bounds__(WF_THREADS, (WF_THREADS <= 256 ? 2 : 1))
wavefront_kernel(const __grid_constant__ DScene sc, const __grid_constant__ WfArgs args,
                 float* __restrict__ out, DCounters* __restrict__ counters,
                 unsigned int* __restrict__ next_pixel) {
  constexpr bool MIS = INTEGRATOR == VIMG_INTEGRATOR_MIS;
  extern __shared__ __align__(16) unsigned char wf_smem[];
  WfShared& sh = *reinterpret_cast<WfShared*>(wf_smem);
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31;
  const uint32_t lane_lt = (1u << lane) - 1u;
  const uint32_t W = sc.cam.res[0], H = sc.cam.res[1];

#ifdef WF_STAGE_CLOCKS
  unsigned long long clk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long t_prev = clock64();
#endif
  bool exhausted = false;  // the frame has no more pixels to hand out
  PathCounters pc = {};
  uint32_t nan_samples = 0, n_samples = 0;
  for (int i = tid; i < WF_SLOTS; i += WF_THREADS) sh.flags[i] = 0;
  if (tid < (int)WF_NUM_CLASSES) sh.class_count[0][tid] = sh.class_count[1][tid] = 0;
  __syncthreads();

  for (uint32_t iter = 0;; ++iter) {
    uint32_t* counts = sh.class_count[iter & 1];
    bool any_alive = false;

    // ================= slot stage (thread tid <-> slots tid, tid + WF_THREADS, ...) =================
#pragma unroll 1
    for (int i = tid; i < WF_SLOTS; i += WF_THREADS) {
      uint32_t f = sh.flags[i];
      uint32_t cls = WF_CLASS_NONE;

      // finish the current sample of this slot's pixel (include/integrators.h:128-137)
      auto finish = [&](V3 result) {
        if (is_nan(result.x) || is_nan(result.y) || is_nan(result.z)) nan_samples++;
        V3 acc = ld3(sh.acc, i) + result;
        n_samples++;
        f &= ~(WF_IN_PATH | WF_END | WF_NEE | WF_NEE_VALID);
        const uint32_t s = ++sh.s[i];
        if (s == args.samples) {
          acc = acc / (float)args.samples;
          uint32_t px, py;
          wf_pixel_xy(args, sc.cam, sh.pixel[i], px, py);
          const size_t image_index = px + ((size_t)(H - 1 - py) * W);
          float* o = out + (args.single_x >= 0 ? 0 : 3 * image_index);
          o[0] = acc.x;
          o[1] = acc.y;
          o[2] = acc.z;
          f &= ~WF_ALIVE;
        } else {
          st3(sh.acc, i, acc);
        }
      };

      // nodes an occluded shadow ray left unvisited: inherited by the closest-hit traversal of the
      // same path's next segment (the reference's shared thread_stack, see bvh_traverse)
      uint2 stk[VIMG_STACK_MAX];
      int stale = 0;
      // A. shadow ray requested by the last shade stage (mis_integrator.cpp:59-78)
      if (MIS && (f & WF_NEE)) {
        Ray sr;
        sr.o = ld3(sh.ray_o, i);  // the shade stage left hit_p here
        sr.dir = ld3(sh.nee_wi, i);
        sr.minT = 0.0001f;
        sr.maxT = sh.nee_dist[i] - 0.0001f;
        PreHit ph;
        pc.rays_shadow++;
        const bool occluded = bvh_traverse<true, COUNT>(sc, sr, ph, &pc.trav, stk, stale);
        if (!occluded && (f & WF_NEE_VALID)) st3(sh.L, i, ld3(sh.L, i) + ld3(sh.nee_c, i));
        f &= ~(WF_NEE | WF_NEE_VALID);
      }
      if (f & WF_END) {
        finish(MIS ? ld3(sh.L, i) : ld3(sh.nee_c, i));
        stale = 0;  // only within a path (the queue pipeline cannot carry a stack across samples)
      }
      if (args.fresh_stack) stale = 0;
      WF_CLK(0);

      // B. pull the next pixel of the frame, start the next sample
      {
        const bool want = !(f & WF_ALIVE) && !exhausted;
        const uint32_t m = __ballot_sync(0xffffffffu, want);
        if (m) {
          uint32_t base = 0;
          if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(next_pixel, (unsigned int)__popc(m));
          base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
          if (want) {
            const uint32_t p = base + __popc(m & lane_lt);
            uint32_t px, py;
            if (p >= args.n_local_pixels) {
              exhausted = true;
            } else if (wf_pixel_xy(args, sc.cam, p, px, py) && args.samples > 0) {
              Pcg32 r;
              pcg32_seed(r, px + ((size_t)(H - 1 - py) * W), 0);
              sh.rng[i] = r.state;
              st3(sh.acc, i, mk3(0.f));
              sh.s[i] = 0;
              sh.pixel[i] = p;
              f = WF_ALIVE;
            }
          }
        }
      }
      if ((f & WF_ALIVE) && !(f & WF_IN_PATH)) {
        // camera ray (include/integrators.h:119-123); rand2 gets the first draw
        uint32_t px, py;
        wf_pixel_xy(args, sc.cam, sh.pixel[i], px, py);
        Pcg32 r;
        r.state = sh.rng[i];
        r.inc = 1;
        V2 off = random_x_y_r2(px + py + sh.s[i]);
        float rand2 = rand_float(r);
        float rand1 = rand_float(r);
        Ray cam = generate_ray(sc.cam, (float)px + off.x, (float)py + off.y, rand1, rand2);
        sh.rng[i] = r.state;
        st3(sh.ray_o, i, cam.o);
        st3(sh.ray_d, i, cam.dir);
        sh.cone_w[i] = cam.cone.cone_width;
        sh.cone_s[i] = cam.cone.spread_angle;
        st3(sh.thr, i, mk3(1.0f));
        sh.eta_scale[i] = 1;
        sh.last_pdf[i] = 0.f;
        sh.d[i] = 0;
        st3(sh.L, i, mk3(0.0f));
        f = WF_ALIVE | WF_IN_PATH | WF_FIRST;
      }

      WF_CLK(1);
      // C. closest-hit traversal + D. resolve
      if (f & WF_IN_PATH) {
        const uint32_t d = sh.d[i];
        if (!MIS && d >= args.depth) {
          finish(mk3(0.0f));  // material: for (d = 0; d < depth; ...) ran out -> vec3(0)
        } else {
          Ray ray;
          ray.o = ld3(sh.ray_o, i);
          ray.dir = ld3(sh.ray_d, i);
          ray.minT = 0.0001f;
          ray.maxT = CUDART_INF_F;
          ray.cone.cone_width = sh.cone_w[i];
          ray.cone.spread_angle = sh.cone_s[i];
          PreHit ph;
          pc.rays_closest++;
          const bool found = bvh_traverse<false, COUNT>(sc, ray, ph, &pc.trav, stk, stale);
          WF_CLK(2);
          const V3 thr = ld3(sh.thr, i);
          if (!found) {
            V3 bg = background_emit(sc, ray.dir, ray.cone);
            if (!MIS) {
              finish(thr * bg);
            } else if (f & WF_FIRST) {
              finish(bg);
            } else {
              V3 L = ld3(sh.L, i);
              const float last_pdf = sh.last_pdf[i];
              if (last_pdf != 0 && sc.bg.emissive) {
                float light_pdf = background_pdf(sc, ray.dir) / sc.n_lights;
                float mis_weight = balance_heuristic(last_pdf, light_pdf);
                L = L + thr * mis_weight * bg;
              }
              finish(L);
            }
          } else {
            const vimg_material& nmat = sc.materials[wf_leaf_material(sc, ph.leaf_slot)];
            cls = wf_material_class(nmat.type);
            if (MIS) {
              if (mat_is_emissive(nmat)) {
                cls = WF_CLASS_NONE;
                const HitInfo next = wf_hit_info(sc, ph.leaf_slot, ph, ray);
                V3 emit = mat_emitted(nmat, ray.dir, next.hit_n_s);
                if (f & WF_FIRST) {
                  finish(emit);
                } else {
                  V3 L = ld3(sh.L, i);
                  const float last_pdf = sh.last_pdf[i];
                  if (last_pdf != 0) {
                    // ray.o is the previous hit point, ray.dir the sampled direction
                    float light_pdf = emitter_surf_pdf(sc, next.prim, ray.o, next.hit_p, ray.dir) / sc.n_lights;
                    float G = geometric_term(ray.o, next.hit_p, next.hit_n_g);
                    float mis_weight = balance_heuristic(last_pdf * G, light_pdf);
                    L = L + thr * mis_weight * emit;
                  } else {
                    L = L + thr * emit;
                  }
                  finish(L);
                }
              } else {
                bool go_on = true;
                uint32_t dn = d;
                if (!(f & WF_FIRST)) {
                  if (d > 5) {
                    Pcg32 r;
                    r.state = sh.rng[i];
                    r.inc = 1;
                    V3 t = thr;
                    go_on = russian_roulette(r, sh.eta_scale[i], t);
                    sh.rng[i] = r.state;
                    if (go_on) st3(sh.thr, i, t);
                  }
                  dn = d + 1;
                }
                if (!go_on || dn >= args.depth || cls == WF_CLASS_NONE) {
                  // RR break / loop condition d < depth / base Material (delta, sample_mat -> nullopt)
                  cls = WF_CLASS_NONE;
                  finish(ld3(sh.L, i));
                } else {
                  sh.d[i] = dn;
                }
              }
            } else if (cls == WF_CLASS_NONE) {
              // material integrator on DiffuseLight / base Material: sample_mat -> nullopt,
              // return throughput * emitted (mat_integrator.cpp:21-81)
              V3 emitted_col = mk3(0.f);
              if (mat_is_emissive(nmat)) {
                const HitInfo next = wf_hit_info(sc, ph.leaf_slot, ph, ray);
                emitted_col = mat_emitted(nmat, ray.dir, next.hit_n_s);
              }
              finish(thr * emitted_col);
            }
            if (cls != WF_CLASS_NONE) {
              sh.hit_slot[i] = ph.leaf_slot;
              sh.hit_t[i] = ray.maxT;
              sh.e0[i] = ph.e0;
              sh.e1[i] = ph.e1;
              sh.e2[i] = ph.e2;
              sh.inv_det[i] = ph.inv_det;
              f &= ~WF_FIRST;
            }
          }
        }
      }
      sh.flags[i] = f;
      any_alive |= (f & WF_ALIVE) != 0;

      // bucket the slots that need shading by material class (one list per class)
#pragma unroll
      for (uint32_t c = 0; c < WF_NUM_CLASSES; ++c) {
        const uint32_t m = __ballot_sync(0xffffffffu, cls == c);
        if (m) {
          uint32_t base = 0;
          if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(&counts[c], (uint32_t)__popc(m));
          base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
          if (cls == c) sh.list[c][base + __popc(m & lane_lt)] = (uint16_t)i;
        }
      }
    }
    WF_CLK(3);
    // barrier 1: lists complete; also the block-wide "anything left to do" vote
    const int live = __syncthreads_or(any_alive || !exhausted);
    WF_CLK(4);
    if (!live) break;
    if (tid < (int)WF_NUM_CLASSES) sh.class_count[(iter + 1) & 1][tid] = 0;  // next iteration's counters

    // ================= shade stage (thread k <-> k-th bucketed slot) =================
    {
      const uint32_t n0 = counts[0], n1 = counts[1], n2 = counts[2];
      // Principled first (longest), then Lambertian, then Dielectric; consecutive items share a
      // class, so whole warps do
#pragma unroll 1
      for (uint32_t item = (uint32_t)tid; item < n0 + n1 + n2; item += WF_THREADS) {
      int j;
      uint32_t k = item;
      if (k < n1) {
        j = sh.list[WF_CLASS_PRINCIPLED][k];
      } else if ((k -= n1) < n0) {
        j = sh.list[WF_CLASS_LAMBERT][k];
      } else {
        j = sh.list[WF_CLASS_DIELECTRIC][k - n0];
      }
      {
        uint32_t fj = sh.flags[j];
        Ray ray;
        ray.o = ld3(sh.ray_o, j);
        ray.dir = ld3(sh.ray_d, j);
        ray.minT = 0.0001f;
        ray.maxT = sh.hit_t[j];
        ray.cone.cone_width = sh.cone_w[j];
        ray.cone.spread_angle = sh.cone_s[j];
        PreHit ph;
        ph.leaf_slot = sh.hit_slot[j];
        ph.e0 = sh.e0[j];
        ph.e1 = sh.e1[j];
        ph.e2 = sh.e2[j];
        ph.inv_det = sh.inv_det[j];
        const HitInfo hit = wf_hit_info(sc, ph.leaf_slot, ph, ray);
        const vimg_material& mat = sc.materials[hit.mat];
        Pcg32 rng;
        rng.state = sh.rng[j];
        rng.inc = 1;
        V3 thr = ld3(sh.thr, j);
        bool nsb = (fj & WF_NSB) != 0;
        float hit_dist = length(ray.o - hit.hit_p);
        float surface_spread_angle
            = spread_angle_from_curvature(hit.mean_curvature, ray.cone.cone_width, ray.dir, hit.hit_n_s);

        if (MIS) {
          // ---- mis_integrator.cpp:45-117 at `hit` ----
          bool nee = false, nee_valid = false;
          EmitterInfo li;
          V3 light_col = mk3(0.f);
          if (!mat_is_delta(mat)) {
            light_col = lights_sample(sc, hit.hit_p, rng, li);
            nee = li.pdf != 0.f;
          }
          const RayCone nee_cone = propagate_reflect_cone(ray.cone, surface_spread_angle * 2.f, hit_dist);
          const bool nee_regularize = nsb;
          ScatterInfo sm;
          const bool scattered = mat_sample(sc, mat, ray.dir, hit, rng, nsb, sm);
          RayCone cone = ray.cone;
          float eta_scale = sh.eta_scale[j];
          if (scattered) {
            if (!sm.is_specular) nsb = true;
            if (sm.eta != 0.f) {
              eta_scale /= (sm.eta * sm.eta);
              cone = propagate_refract_cone(ray.cone, ray.dir, surface_spread_angle, sm.eta, sm.wo);
            } else {
              cone = nee_cone;
            }
          }
          // ONE inlined eval_pdf_pair site serves both calls of the loop body:
          //   q = 0 light direction (mis_integrator.cpp:68-71), q = 1 sampled direction (:104-106)
          V3 contrib = mk3(0.f), mat_sample_eval = mk3(0.f);
          float mat_sample_pdf = 0.f;
#pragma unroll 1
          for (int q = 0; q < 2; ++q) {
            if (!(q == 0 ? nee : scattered)) continue;
            V3 e;
            float p;
            mat_eval_pdf_pair(sc, mat, ray.dir, q == 0 ? li.wi : sm.wo, hit, q == 0 ? nee_cone : cone,
                              q == 0 ? nee_regularize : nsb, e, p);
            if (q == 0) {
              if (p != 0 && !is_nan(p)) {
                float G = li.G;
                float mis_weight = balance_heuristic(li.pdf, p * G);
                contrib = thr * e * mis_weight * G * light_col / li.pdf;
                nee_valid = true;
              }
            } else {
              mat_sample_eval = e;
              mat_sample_pdf = p;
            }
          }
          fj &= ~(WF_NSB | WF_NEE | WF_NEE_VALID | WF_END);
          if (nsb) fj |= WF_NSB;
          if (nee) {
            fj |= WF_NEE | (nee_valid ? WF_NEE_VALID : 0u);
            st3(sh.nee_wi, j, li.wi);
            sh.nee_dist[j] = li.dist;
            st3(sh.nee_c, j, contrib);
          }
          st3(sh.ray_o, j, hit.hit_p);  // origin of the shadow ray and of the next ray
          if (!scattered || is_nan(mat_sample_pdf)) {
            fj |= WF_END;  // return bounce_result (after the pending shadow ray)
          } else {
            thr = thr * (mat_sample_eval / mat_sample_pdf);
            st3(sh.thr, j, thr);
            sh.last_pdf[j] = mat_sample_pdf;
            sh.eta_scale[j] = eta_scale;
            st3(sh.ray_d, j, sm.wo);
            sh.cone_w[j] = cone.cone_width;
            sh.cone_s[j] = cone.spread_angle;
          }
        } else {
          // ---- material_integrator, mat_integrator.cpp:21-83 at `hit` ----
          V3 emitted_col = mat_emitted(mat, ray.dir, hit.hit_n_s);
          ScatterInfo sr;
          fj &= ~(WF_NSB | WF_END);
          if (!mat_sample(sc, mat, ray.dir, hit, rng, nsb, sr)) {
            st3(sh.nee_c, j, thr * emitted_col);
            fj |= WF_END;
          } else {
            if (!sr.is_specular) nsb = true;
            RayCone cone;
            float eta_scale = sh.eta_scale[j];
            if (sr.eta != 0.f) {
              eta_scale /= (sr.eta * sr.eta);
              cone = propagate_refract_cone(ray.cone, ray.dir, surface_spread_angle, sr.eta, sr.wo);
            } else {
              cone = propagate_reflect_cone(ray.cone, surface_spread_angle * 2.f, hit_dist);
            }
            thr = thr * (emitted_col + mat_eval_div_pdf(sc, mat, ray.dir, sr.wo, hit, cone, nsb));
            const uint32_t d = sh.d[j];
            bool go_on = true;
            if (d > 5) go_on = russian_roulette(rng, eta_scale, thr);
            if (!go_on) {
              st3(sh.nee_c, j, mk3(0.0f));  // `break`: return vec3(0)
              fj |= WF_END;
            } else {
              st3(sh.thr, j, thr);
              sh.eta_scale[j] = eta_scale;
              st3(sh.ray_o, j, hit.hit_p);
              st3(sh.ray_d, j, sr.wo);
              sh.cone_w[j] = cone.cone_width;
              sh.cone_s[j] = cone.spread_angle;
              sh.d[j] = d + 1;
            }
          }
          if (nsb) fj |= WF_NSB;
        }
        sh.rng[j] = rng.state;
        sh.flags[j] = fj;
      }
      }
    }
    WF_CLK(5);
    __syncthreads();  // barrier 2: shade results visible to the slots' home threads
    WF_CLK(6);
  }

#ifdef WF_STAGE_CLOCKS
  if (lane == 0)
    for (int k = 0; k < 8; ++k) atomicAdd(&counters->stage_clk[k], clk[k]);
#endif
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

}  // namespace vimg
