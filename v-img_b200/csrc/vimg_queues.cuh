// vimg_queues.cuh -- multi-kernel "queue wavefront" path tracer (material / mis).
//
// The block-local wavefront (vimg_wavefront.cuh) keeps all path state in shared memory, which
// pins it to ONE 512-thread block per SM at 128 registers: 4 warps per scheduler, issue slots
// 37 % busy, ~11 of 32 lanes active inside the traversal (profiles/r1f).  This pipeline trades
// shared-memory state for occupancy and coherence:
//
//   * path state of N slots lives in HBM/L2 as SoA (16-byte records per field group);
//   * TRACE kernel: small (<= 64 registers, 32+ warps per SM), persistent, pulls closest-hit and
//     shadow rays from two queues; a lane whose ray ends is refilled from the warp's chunk, so
//     the traversal loop runs with most lanes busy;
//   * CLASSIFY kernel (thread <-> traced closest ray): routes the slot by what the ray hit into
//     one SHADE queue per material class or into the END queue (miss, emitter, base Material);
//   * SHADE kernels, one per material class (Lambertian, Principled, Dielectric), each compiled
//     with only its own BSDF: pending shadow-ray result, Russian roulette and depth test, light
//     sampling, BSDF sampling, both eval_pdf_pair calls, next ray and shadow-ray request;
//   * END kernel (thread <-> ended path): miss / emitter contribution, sample accumulation,
//     next pixel, next camera ray.
//   Every kernel is driven by a queue of slot numbers, none walks the whole slot pool, so the
//   long tail of a frame (a few pixels with very long paths) costs launch latency only.
//
// Per path the sequence of float operations and RNG draws is the reference's
// (src/integrators/mis_integrator.cpp:18-189, mat_integrator.cpp:4-85) and identical to the
// block-local kernel's, so the images agree bit for bit (tests/test_gpu_parity.py runs both).
// The reference seeds one PCG stream per pixel and carries it over the pixel's samples
// (include/integrators.h:116-131), so one path per pixel is in flight; a slot whose pixel is
// complete pulls the next pixel of this rank's tile list.
#pragma once
#include "vimg_wavefront.cuh"

namespace vimg {

#ifndef QW_TRACE_THREADS
#define QW_TRACE_THREADS 256
#endif
#ifndef QW_TRACE_BLOCKS_PER_SM
#define QW_TRACE_BLOCKS_PER_SM 4  // 64 registers per thread, 32 warps per SM
#endif
#ifndef QW_REFILL_MIN
#define QW_REFILL_MIN 12  // refill a warp's idle lanes once at least this many are idle
#endif
#ifndef QW_INNER_MIN_LANES
#define QW_INNER_MIN_LANES 8  // leave the inner-node loop once fewer lanes than this are still in it
#endif
// a closest-ray queue entry with this bit carries the slot's shadow ray too (traced first, same lane)
#define QW_PAIR_BIT 0x80000000u
#ifndef QW_CHUNK
#define QW_CHUNK 64  // rays a warp takes from the global cursor with one atomic
#endif
#ifndef QW_THREADS
#define QW_THREADS 256  // classify / shade / end block size
#endif
#ifndef QW_END_MIN_BLOCKS
#define QW_END_MIN_BLOCKS 3
#endif

struct QwCounters {
  uint32_t n_closest[2], n_shadow[2];  // queue lengths, double buffered by iteration parity
  uint32_t n_end[2];                   // (qw_shade_kernel addresses these three pairs as &n_closest[0] + 2 k + parity)
  uint32_t n_shade[2][WF_NUM_CLASSES];  // by iteration parity
  uint32_t cursor;      // trace kernel's work cursor
  uint32_t pad_[3];
};

// Path state of one slot: five 32-byte records (one DRAM sector each), grouped by who reads and
// writes them together, so that every kernel moves whole sectors with 16-byte loads and stores.
struct __align__(32) QwRay {   // END / SHADE write; TRACE, SHADE, END read
  float4 o_w;  // origin xyz (hit point after shading), cone_width
  float4 d_s;  // direction xyz, cone spread_angle
};
struct __align__(32) QwHit {   // TRACE writes (closest ray: h0, h1; shadow ray: occl)
  float4 h0;   // leaf slot (int bits), t, e0, e1
  float2 h1;   // e2, inv_det
  uint32_t occl, pad_;
};
struct __align__(32) QwPath {  // SHADE and END read and write
  float4 thr_eta;  // throughput rgb, eta_scale
  unsigned long long rng;
  uint32_t flags, d;
};
struct __align__(32) QwNee {   // SHADE writes; TRACE reads wi_dist; SHADE / END read c_pdf
  float4 wi_dist;  // shadow ray direction xyz, light distance
  float4 c_pdf;    // pending NEE contribution rgb (material: result of an ended path), last_pdf
};
struct __align__(32) QwPix {   // END reads and writes; SHADE adds the NEE contribution to L
  float4 L_s;        // bounce_result of the path in flight rgb, sample counter (uint bits)
  float4 acc_pixel;  // pixel accumulator rgb, local pixel number (uint bits)
};

// The five records of a slot lie side by side (160 bytes): a kernel that gathers a slot's state
// touches one DRAM page and 2.5-3 64-byte bursts instead of five pages (measured: DRAM read
// traffic of the END / SHADE kernels was 2.5x their sector bytes with five separate arrays).
struct __align__(32) QwSlot {
  QwRay ray;
  QwHit hit;
  QwPath path;
  QwNee nee;
  QwPix pix;
};
static_assert(sizeof(QwSlot) == 160, "slot state must be 160 bytes");

struct QwState {
  uint32_t n_slots;
  QwSlot* slot;
  uint32_t* q_closest[2];
  int32_t* q_leaf;  // leaf record hit by closest ray i of the queue TRACE just consumed (-1: miss):
                    // CLASSIFY reads it next to q_closest[i], coalesced, instead of gathering slot state
  uint32_t* q_shadow[2];
  uint32_t* q_end[2];
  uint2* q_shade[WF_NUM_CLASSES];  // {slot, leaf record of the hit}: the shade launch fetches both at once
  QwCounters* cnt;
  uint32_t* next_pixel;  // next local pixel number to hand out (shared by all pools of a frame)
};

VD V3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }
VD float4 f4(V3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
// slot state is touched once per kernel: keep it out of L1 (which holds the scene tables)
VD float4 ldg4(const float4* p) { return __ldcg(p); }
VD void stg4(float4* p, float4 v) { __stcg(p, v); }
VD uint4 ldg4u(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

// Block-wide queue append.  Every thread of the block calls this (want may be false);
// one global atomic per queue per block.  `s_cnt`/`s_base` are NQ shared words each.
template <int NQ>
struct QwAppend {
  uint32_t off[NQ];
  VD void count(uint32_t* s_cnt, const bool (&want)[NQ], uint32_t lane, uint32_t lane_lt) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const uint32_t m = __ballot_sync(0xffffffffu, want[q]);
      uint32_t base = 0;
      if (m) {
        if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(&s_cnt[q], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
      }
      off[q] = base + __popc(m & lane_lt);
    }
  }
};

// ================================================================== TRACE ====
// BVH::hit<closest> and BVH::occlude (include/bvh.h:83-225) for every queued ray; same
// visiting order, box arithmetic and tie rules as bvh_traverse (vimg_bvh.cuh), restructured
// as a persistent loop whose idle lanes are refilled.
template <bool COUNT>
__global__ void __launch_bounds__(QW_TRACE_THREADS, QW_TRACE_BLOCKS_PER_SM)
qw_trace_kernel(const __grid_constant__ DScene sc, const __grid_constant__ QwState st,
                const uint32_t parity, DCounters* __restrict__ counters) {
  QwCounters* c = st.cnt;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // queues of the NEXT iteration (their last consumer, the previous trace launch, is done)
    c->n_closest[parity ^ 1u] = 0;
    c->n_shadow[parity ^ 1u] = 0;
    c->n_end[parity ^ 1u] = 0;
#pragma unroll
    for (uint32_t k = 0; k < WF_NUM_CLASSES; ++k) c->n_shade[parity ^ 1u][k] = 0;
  }
  const uint32_t n_c = c->n_closest[parity], total = n_c + c->n_shadow[parity];
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counters->stage_items[VIMG_STAGE_TRACE], (unsigned long long)total);
  const uint32_t* __restrict__ qc = st.q_closest[parity];
  const uint32_t* __restrict__ qs = st.q_shadow[parity];
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lane_lt = (1u << lane) - 1u;

  // rays a warp takes from the global cursor with one atomic: large when there is plenty of
  // work (few same-address atomics), one ray per lane when the queues are short (parallelism)
  const uint32_t n_warps = gridDim.x * (QW_TRACE_THREADS / 32);
  uint32_t chunk = (total / n_warps) & ~31u;
  chunk = chunk < 32u ? 32u : (chunk > (uint32_t)QW_CHUNK ? (uint32_t)QW_CHUNK : chunk);
  PathCounters pc = {};
  uint32_t chunk_next = 0, chunk_end = 0;  // this warp's private range of work items
  bool more = total > 0;

  // per-lane ray: `have` = in traversal, `done` = finished, result not stored yet
  // `pair`: this lane traces a shadow ray whose path goes on -- the closest ray of the same slot follows
  // on the same lane and inherits the stack the occlusion traversal leaves (see bvh_traverse)
  bool have = false, done = false, any_hit = false, occluded = false, pair = false;
  uint32_t slot = 0, qpos = 0;
  Ray ray;
  RayPre pre;
  PreHit ph;
  uint2 cur = make_uint2(0, 0);
  int sp = 0;
  uint2 stack[VIMG_STACK_MAX];
  ray.minT = 0.0001f;
  ray.maxT = 0.f;
  ray.o = ray.dir = mk3(0.f);
  pre.inv_dir = mk3(0.f);
  pre.has_inf = false;
  pre.kx = pre.ky = pre.kz = 0;
  pre.Sx = pre.Sy = pre.Sz = 0.f;
  ph.e0 = ph.e1 = ph.e2 = ph.inv_det = 0.f;
  ph.leaf_slot = -1;

  for (;;) {
    if (done) {  // the one place results are stored
      QwHit* h = &st.slot[slot].hit;
      bool store_closest = !any_hit;
      if (any_hit) {
        h->occl = occluded ? 1u : 0u;
        if (pair) {
          // mis_integrator.cpp:64 -> :119 on one thread_stack: BVH::hit pushes the root on top of what
          // BVH::occlude left when it returned early (sp entries; 0 after a completed traversal)
          pair = false;
          any_hit = false;
          occluded = false;
          const float4 d4 = ldg4(&st.slot[slot].ray.d_s);
          ray.dir = xyz(d4);
          ray.maxT = CUDART_INF_F;
          ph.e0 = ph.e1 = ph.e2 = ph.inv_det = 0.f;  // the occluder's record is not a hit of this ray
          ph.leaf_slot = -1;
          pc.rays_closest++;
          store_closest = true;
          if (sc.n_nodes != 0) {
            pre = ray_precompute(ray);
            const float root_hit = slab(sc.root_min, sc.root_max, ray, pre);
            if (!is_inf(root_hit)) {
              if (sp + (int)sc.stack_cap > VIMG_STACK_MAX) sp = 0;
              cur = make_uint2(sc.root_first, sc.root_count);
              have = true;
              store_closest = false;
            }
          }
        }
      }
      if (store_closest) {
        stg4(&h->h0, make_float4(__int_as_float(ph.leaf_slot), ray.maxT, ph.e0, ph.e1));
        h->h1 = make_float2(ph.e2, ph.inv_det);
        st.q_leaf[qpos] = ph.leaf_slot;
      }
      done = false;
    }
    const uint32_t idle = __ballot_sync(0xffffffffu, !have);
    if (idle == 0xffffffffu || (more && __popc(idle) >= QW_REFILL_MIN)) {
      if (more) {
        if (chunk_next == chunk_end) {
          uint32_t b = 0;
          if (lane == 0) b = atomicAdd(&c->cursor, chunk);
          b = __shfl_sync(0xffffffffu, b, 0);
          chunk_next = b < total ? b : total;
          chunk_end = b + chunk < total ? b + chunk : total;
          if (chunk_next == chunk_end) more = false;
        }
        if (!have) {
          const uint32_t idx = chunk_next + __popc(idle & lane_lt);
          if (idx < chunk_end) {
            ray.minT = 0.0001f;
            ph.e0 = ph.e1 = ph.e2 = ph.inv_det = 0.f;
            ph.leaf_slot = -1;
            occluded = false;
            sp = 0;  // an earlier shadow ray of this lane may have ended with entries left
            float4 o4, d4;
            if (idx < n_c) {
              const uint32_t entry = qc[idx];
              slot = entry & ~QW_PAIR_BIT;
              qpos = idx;
              o4 = ldg4(&st.slot[slot].ray.o_w);
              pair = (entry & QW_PAIR_BIT) != 0;
              any_hit = pair;
              if (pair) {  // the shadow ray first, then the closest ray from the same origin
                d4 = ldg4(&st.slot[slot].nee.wi_dist);
                ray.maxT = d4.w - 0.0001f;
                pc.rays_shadow++;
              } else {
                d4 = ldg4(&st.slot[slot].ray.d_s);
                ray.maxT = CUDART_INF_F;
                pc.rays_closest++;
              }
            } else {
              // shadow ray (mis_integrator.cpp:59-64): origin = hit point, maxT = dist - 1e-4
              any_hit = true;
              pair = false;
              slot = qs[idx - n_c];
              o4 = ldg4(&st.slot[slot].ray.o_w);
              d4 = ldg4(&st.slot[slot].nee.wi_dist);
              ray.maxT = d4.w - 0.0001f;
              pc.rays_shadow++;
            }
            ray.o = xyz(o4);
            ray.dir = xyz(d4);
            done = true;
            if (sc.n_nodes != 0) {
              pre = ray_precompute(ray);
              const float root_hit = slab(sc.root_min, sc.root_max, ray, pre);
              if (!is_inf(root_hit)) {
                cur = make_uint2(sc.root_first, sc.root_count);
                have = true;
                done = false;
              }
            }
          }
        }
        const uint32_t taken = chunk_next + __popc(idle);
        chunk_next = taken < chunk_end ? taken : chunk_end;
      }
      if (!__any_sync(0xffffffffu, have || done)) {
        if (!more) break;
        continue;
      }
    }

    // inner nodes: descend until this lane holds a leaf or its ray is done
    while (have && cur.y == 0) {
      if (COUNT) pc.trav.inner++;
      const float4* rec = reinterpret_cast<const float4*>(sc.pairs + ((cur.x - 1u) >> 1));
      const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2);
      const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(rec + 3));
      const float lmin[3] = {q0.x, q0.y, q0.z}, rmin[3] = {q0.w, q1.x, q1.y};
      const float lmax[3] = {q1.z, q1.w, q2.x}, rmax[3] = {q2.y, q2.z, q2.w};
      const float bb_hit1 = slab(lmin, lmax, ray, pre);
      const float bb_hit2 = slab(rmin, rmax, ray, pre);
      uint2 first_child = make_uint2(q3.x, q3.y), sec_child = make_uint2(q3.z, q3.w);
      const bool h1 = !is_inf(bb_hit1), h2 = !is_inf(bb_hit2);
      if (h2) {
        if (h1) {
          // closest hit: the more distant box is pushed first (tie: first child pushed first);
          // any hit: unordered, left pushed then right (bvh.h:195-212)
          if (!any_hit && bb_hit2 > bb_hit1) {
            const uint2 t = first_child;
            first_child = sec_child;
            sec_child = t;
          }
          stack[sp++] = first_child;
#ifdef QW_PREFETCH_PUSH
          // the pushed node is visited later: start moving its record towards the SM now
          if (first_child.y == 0)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(sc.pairs + ((first_child.x - 1u) >> 1)));
          else
            asm volatile("prefetch.global.L2 [%0];" ::"l"(sc.leaf_prims + first_child.x));
#endif
        }
        cur = sec_child;
      } else if (h1) {
        cur = first_child;
      } else if (sp > 0) {
        cur = stack[--sp];
      } else {
        have = false;
        done = true;
      }
      // most lanes of the warp wait with a leaf (or for a new ray): stop descending, serve them
      if (QW_INNER_MIN_LANES > 0 && __popc(__activemask()) < QW_INNER_MIN_LANES) break;
    }
    if (have && cur.y != 0) {  // leaf: primitives in obj_indices order, the last success wins
      if (COUNT) pc.trav.leaf++;
      for (uint32_t i = 0; i < cur.y; ++i) {
        const uint32_t ls = cur.x + i;
        const float4* rec = reinterpret_cast<const float4*>(sc.leaf_prims + ls);
        const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
        const float a[9] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
        const uint32_t flags = __float_as_uint(r2.z);
        bool hit;
        if (flags & VIMG_LEAF_SPHERE) {
          if (COUNT) pc.trav.sph++;
          hit = sphere_hit(a, ray);
          if (hit) ph.e0 = ph.e1 = ph.e2 = ph.inv_det = 0.f;
        } else {
          if (COUNT) pc.trav.tri++;
          hit = tri_hit(a, flags, ray, pre, ph);
        }
        if (hit) {
          ph.leaf_slot = (int32_t)ls;
          if (any_hit) {
            occluded = true;
            break;
          }
        }
      }
      if (!occluded && sp > 0) {
        cur = stack[--sp];
      } else {
        have = false;
        done = true;
      }
    }
  }

  warp_add_counter(&counters->rays_closest, pc.rays_closest);
  warp_add_counter(&counters->rays_shadow, pc.rays_shadow);
  if (COUNT) {
    warp_add_counter(&counters->inner_nodes, pc.trav.inner);
    warp_add_counter(&counters->leaf_nodes, pc.trav.leaf);
    warp_add_counter(&counters->tri_tests, pc.trav.tri);
    warp_add_counter(&counters->sphere_tests, pc.trav.sph);
  }
}

// =============================================================== CLASSIFY ====
// thread <-> closest ray of this iteration: route the slot by what the ray hit.  Shadeable
// materials go to their class's shade queue; misses, emitter hits and base-Material hits end
// the path and go to the END queue of this iteration.
__global__ void __launch_bounds__(QW_THREADS)
qw_classify_kernel(const __grid_constant__ DScene sc, const __grid_constant__ QwState st, const uint32_t parity) {
  __shared__ uint32_t s_cnt[WF_NUM_CLASSES + 1], s_base[WF_NUM_CLASSES + 1];
  QwCounters* c = st.cnt;
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  if (blockIdx.x == 0 && tid == 0) c->cursor = 0;  // the next trace launch starts at item 0
  const uint32_t n = c->n_closest[parity];
  const uint32_t* __restrict__ q = st.q_closest[parity];
  for (uint32_t base = blockIdx.x * QW_THREADS; base < n; base += gridDim.x * QW_THREADS) {
    if (tid < (int)WF_NUM_CLASSES + 1) s_cnt[tid] = 0;
    __syncthreads();
    const uint32_t item = base + tid;
    uint32_t cls = WF_CLASS_NONE + 1, j = 0;  // NONE + 1: no item
    int32_t leaf_slot = -1;
    if (item < n) {
      j = q[item] & ~QW_PAIR_BIT;
      leaf_slot = __ldcg(&st.q_leaf[item]);
      cls = WF_CLASS_NONE;
      if (leaf_slot >= 0) {
        const uint32_t lc = sc.leaf_class[leaf_slot];
        cls = (lc & VIMG_LEAF_CLASS_EMISSIVE) ? WF_CLASS_NONE : (lc & 3u);
      }
    }
    bool want[WF_NUM_CLASSES + 1];
#pragma unroll
    for (uint32_t k = 0; k <= WF_NUM_CLASSES; ++k) want[k] = cls == k;
    QwAppend<WF_NUM_CLASSES + 1> ap;
    ap.count(s_cnt, want, lane, lane_lt);
    __syncthreads();
    if (tid < (int)WF_NUM_CLASSES) {
      if (s_cnt[tid]) s_base[tid] = atomicAdd(&c->n_shade[parity][tid], s_cnt[tid]);
    } else if (tid == (int)WF_NUM_CLASSES) {
      if (s_cnt[tid]) s_base[tid] = atomicAdd(&c->n_end[parity], s_cnt[tid]);
    }
    __syncthreads();
    if (cls < WF_NUM_CLASSES) st.q_shade[cls][s_base[cls] + ap.off[cls]] = make_uint2(j, (uint32_t)leaf_slot);
    else if (cls == WF_CLASS_NONE) st.q_end[parity][s_base[WF_NUM_CLASSES] + ap.off[WF_NUM_CLASSES]] = j;
    __syncthreads();
  }
}

// ==================================================================== END ====
// thread <-> slot whose path ended (miss, emitter hit, Russian roulette, depth, no scattering)
// or that has no pixel yet.  Finishes the sample (include/integrators.h:128-137), hands out
// the next pixel of this rank's tile list and generates the next camera ray (:116-123).
// What END reads for one slot and what it leaves behind (layout: the kernel's business, see ShadeIO)
struct EndIO {
  // in
  float4 thr4, o4, d4, h0, cp4, L4, ap4;  // throughput | ray origin | ray direction | leaf, t, e0, e1 | pending contribution,
                                          // last_pdf | L, sample counter | pixel accumulator, local pixel number
  uint4 pa1, hb;                          // rng lo, rng hi, flags, depth counter | e2, inv_det, shadow ray occluded, -
  // out: rng / flags of the slot, its sample counter and accumulator, and the camera ray of its next sample
  uint4 rng_fd;
  float4 L_s, acc_pixel, cam_o, cam_d;
  bool new_ray;
};

// Finishes the sample of a path that ended (miss, emitter hit, Russian roulette, depth, no scattering;
// include/integrators.h:128-137), hands out the next pixel of this rank's tile list and generates the next camera
// ray (:116-123).  Called by EVERY thread of the warp (`valid`: this thread has a slot): the pixel hand-out votes.
template <int INTEGRATOR>
VD void end_item(const DScene& sc, const WfArgs& args, uint32_t* next_pixel, float* __restrict__ out, const bool valid,
                 bool& exhausted, const uint32_t lane, const uint32_t lane_lt, uint32_t& nan_samples, uint32_t& n_samples,
                 EndIO& io) {
  constexpr bool MIS = INTEGRATOR == VIMG_INTEGRATOR_MIS;
  const uint32_t W = sc.cam.res[0], H = sc.cam.res[1];
  const float4 thr4 = io.thr4, o4 = io.o4, d4 = io.d4, h0 = io.h0, cp4 = io.cp4, L4 = io.L4, ap4 = io.ap4;
  const uint4 pa1 = io.pa1, hb = io.hb;
  uint32_t f = pa1.z;
  const uint32_t d = pa1.w;
  Pcg32 rng;
  rng.state = ((unsigned long long)pa1.y << 32) | pa1.x;
  uint32_t range = f >> WF_RANGE_SHIFT;  // sample range of the pixel this slot works on
  rng.inc = ((uint64_t)range << 1) | 1u;
  V3 acc = xyz(ap4);
  uint32_t s = __float_as_uint(L4.w), pixel = __float_as_uint(ap4.w);
  const bool occl = hb.z != 0;
  bool new_ray = false;

  // finish the current sample of this slot's pixel (include/integrators.h:128-137)
  auto finish = [&](V3 result) {
    if (is_nan(result.x) || is_nan(result.y) || is_nan(result.z)) {
      nan_samples++;
#ifdef VIMG_FAST
      result = mk3(0.f);  // tolerance build: a NaN sample is counted and dropped (the reference adds it: the pixel turns NaN)
#endif
    }
    acc = acc + result;
    n_samples++;
    f &= ~(WF_IN_PATH | WF_END | WF_NEE | WF_NEE_VALID);
    s = s + 1;
    if (args.streams == 1) {
      if (s == args.samples) {
        const V3 px_col = acc / (float)args.samples;
        uint32_t px, py;
        wf_pixel_xy(args, sc.cam, pixel, px, py);
        const size_t image_index = px + ((size_t)(H - 1 - py) * W);
        float* o = out + (args.single_x >= 0 ? 0 : 3 * image_index);
        o[0] = px_col.x;
        o[1] = px_col.y;
        o[2] = px_col.z;
        f &= ~WF_ALIVE;
      }
    } else if (s == wf_range_begin(args, range + 1)) {
      // end of this sample range: its partial sum; qw_combine_kernel adds the ranges in order
      uint32_t px, py;
      wf_pixel_xy(args, sc.cam, pixel, px, py);
      const size_t image_index = px + ((size_t)(H - 1 - py) * W);
      float* o = args.partial + 3 * (image_index * args.streams + range);
      o[0] = acc.x;
      o[1] = acc.y;
      o[2] = acc.z;
      f &= ~WF_ALIVE;
    }
  };

  if (valid && (f & WF_IN_PATH)) {
    // pending shadow ray of the last bounce (mis_integrator.cpp:64-78)
    V3 L = xyz(L4);
#ifdef VIMG_DEBUG_NEE
    if (MIS) printf("D end: s=%u nee=%d valid=%d occl=%d c=%.9g %.9g %.9g L=%.9g %.9g %.9g END=%d leaf=%d lastpdf=%.9g\n", s, (f & WF_NEE) != 0, (f & WF_NEE_VALID) != 0, (int)occl,
                    cp4.x, cp4.y, cp4.z, L4.x, L4.y, L4.z, (f & WF_END) != 0, __float_as_int(h0.x), cp4.w);
#endif
    if (MIS && (f & WF_NEE) && (f & WF_NEE_VALID) && !occl) L = L + xyz(cp4);
    if (f & WF_END) {
      finish(MIS ? L : xyz(cp4));
    } else {
      Ray ray;
      ray.o = xyz(o4);
      ray.dir = xyz(d4);
      ray.minT = 0.0001f;
      ray.maxT = h0.y;
      ray.cone.cone_width = o4.w;
      ray.cone.spread_angle = d4.w;
      PreHit ph;
      ph.leaf_slot = __float_as_int(h0.x);
      const V3 thr = xyz(thr4);
      const float last_pdf = cp4.w;
      if (ph.leaf_slot < 0) {
        // miss (mis_integrator.cpp:36-38,173-185; mat_integrator.cpp:24-27)
        V3 bg = background_emit(sc, ray.dir, ray.cone);
        if (!MIS) {
          finish(thr * bg);
        } else if (f & WF_FIRST) {
          finish(bg);
        } else {
          if (last_pdf != 0 && sc.bg.emissive) {
            float light_pdf = background_pdf(sc, ray.dir) / sc.n_lights;
            float mis_weight = balance_heuristic(last_pdf, light_pdf);
#ifdef VIMG_DEBUG_NEE
            printf("D miss: thr=%.9g %.9g %.9g bg=%.9g %.9g %.9g light_pdf=%.9g w=%.9g dir=%.9g %.9g %.9g cone=%.9g %.9g\n", thr.x, thr.y, thr.z, bg.x, bg.y, bg.z,
                   light_pdf, mis_weight, ray.dir.x, ray.dir.y, ray.dir.z, ray.cone.cone_width, ray.cone.spread_angle);
#endif
            L = L + thr * mis_weight * bg;
          }
          finish(L);
        }
      } else {
        const vimg_material& nmat = sc.materials[wf_leaf_material(sc, ph.leaf_slot)];
        if (mat_is_emissive(nmat)) {
          ph.e0 = h0.z;
          ph.e1 = h0.w;
          ph.e2 = __uint_as_float(hb.x);
          ph.inv_det = __uint_as_float(hb.y);
          const DLeafPrim nlp = sc.leaf_prims[ph.leaf_slot];
          // an emitter's hit: position and normals only
          const HitInfo next = wf_hit_info<false>(sc, nlp, load_shade_rec(sc, ph.leaf_slot), ph, ray);
          V3 emit = mat_emitted(nmat, ray.dir, next.hit_n_s);
          if (!MIS) {
            finish(thr * emit);  // sample_mat -> nullopt: throughput * emitted
          } else if (f & WF_FIRST) {
            finish(emit);
          } else {
            if (last_pdf != 0) {
              // ray.o is the previous hit point, ray.dir the sampled direction
              float light_pdf = emitter_surf_pdf_leaf(nlp, ray.o, next.hit_p, ray.dir) / sc.n_lights;
              float G = geometric_term(ray.o, next.hit_p, next.hit_n_g);
              float mis_weight = balance_heuristic(last_pdf * G, light_pdf);
              L = L + thr * mis_weight * emit;
            } else {
              L = L + thr * emit;
            }
            finish(L);
          }
        } else if (!MIS) {
          finish(mk3(0.f));  // base Material: sample_mat -> nullopt, emitted = 0
        } else {
          // base Material under mis: the Russian-roulette draw still happens before
          // sample_mat -> nullopt ends the path (mis_integrator.cpp:154-167, 83-88)
          if (!(f & WF_FIRST) && d > 5) {
            V3 t = thr;
            russian_roulette(rng, thr4.w, t);
          }
          finish(L);
        }
      }
    }
  }

  // the next pixel of the frame for a slot without one
  {
    const bool want = valid && !(f & WF_ALIVE);
    const uint32_t m = __ballot_sync(0xffffffffu, want);
    if (m) {
      uint32_t pbase = 0;
      if (lane == (uint32_t)(__ffs(m) - 1)) {
        // no atomic once the frame is handed out (the counter was looked at when the launch
        // started; a warp that sees it run out remembers)
        pbase = exhausted ? args.n_local_units : atomicAdd(next_pixel, (uint32_t)__popc(m));
      }
      pbase = __shfl_sync(0xffffffffu, pbase, __ffs(m) - 1);
      exhausted = exhausted || pbase >= args.n_local_units;
      if (want) {
        // tiles overhanging the frame contain pixel numbers without a pixel: take the next
        uint32_t u = pbase + __popc(m & lane_lt), p = 0;
        uint32_t px = 0, py = 0;
        bool got = false;
        while (u < args.n_local_units && !(got = wf_pixel_xy(args, sc.cam, p = fast_div(u, args.div_my_streams), px, py)))
          u = atomicAdd(next_pixel, 1u);
        if (got) {
          range = args.s_rank + (u - p * args.my_streams) * args.s_n_ranks;
          if (!MIS && args.depth == 0) {
            // material_integrator with depth 0 never traces: every sample is vec3(0)
            const size_t image_index = px + ((size_t)(H - 1 - py) * W);
            float* o = out + (args.single_x >= 0 ? 0 : 3 * image_index);
            o[0] = o[1] = o[2] = 0.f;
            n_samples += args.samples;
          } else {
            pcg32_seed(rng, px + ((size_t)(H - 1 - py) * W), range);
            acc = mk3(0.f);
            s = wf_range_begin(args, range);
            pixel = p;
            f = WF_ALIVE | (range << WF_RANGE_SHIFT);
          }
        }
      }
    }
  }
  // camera ray of the next sample (include/integrators.h:119-123); rand2 gets the first draw
  if (f & WF_ALIVE) {
    uint32_t px, py;
    wf_pixel_xy(args, sc.cam, pixel, px, py);
    V2 off = random_x_y_r2(px + py + s);
    float rand2 = rand_float(rng);
    float rand1 = rand_float(rng);
    Ray cam = generate_ray(sc.cam, (float)px + off.x, (float)py + off.y, rand1, rand2);
    io.cam_o = f4(cam.o, cam.cone.cone_width);
    io.cam_d = f4(cam.dir, cam.cone.spread_angle);
    f = WF_ALIVE | WF_IN_PATH | WF_FIRST | (range << WF_RANGE_SHIFT);
    new_ray = true;
  }
  io.rng_fd = make_uint4((uint32_t)rng.state, (uint32_t)(rng.state >> 32), f, 0u);
  io.L_s = make_float4(0.f, 0.f, 0.f, __uint_as_float(s));
  io.acc_pixel = f4(acc, __uint_as_float(pixel));
  io.new_ray = new_ray;
}

template <int INTEGRATOR>
__global__ void __launch_bounds__(QW_THREADS, QW_END_MIN_BLOCKS)
qw_end_kernel(const __grid_constant__ DScene sc, const __grid_constant__ QwState st,
              const __grid_constant__ WfArgs args, const uint32_t parity,
              float* __restrict__ out, DCounters* __restrict__ counters) {
  constexpr bool MIS = INTEGRATOR == VIMG_INTEGRATOR_MIS;
  __shared__ uint32_t s_cnt[1], s_base[1];
  QwCounters* c = st.cnt;
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  const uint32_t W = sc.cam.res[0], H = sc.cam.res[1];
  const uint32_t n = c->n_end[parity];
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&counters->stage_items[VIMG_STAGE_END], (unsigned long long)n);
  const uint32_t* __restrict__ q = st.q_end[parity];
  uint32_t nan_samples = 0, n_samples = 0;
  // one look at the unit counter per warp and launch, in flight together with the first gathers
  bool exhausted = *reinterpret_cast<volatile uint32_t*>(st.next_pixel) >= args.n_local_units;

  for (uint32_t base = blockIdx.x * QW_THREADS; base < n; base += gridDim.x * QW_THREADS) {
    if (tid == 0) s_cnt[0] = 0;
    __syncthreads();
    const uint32_t item = base + tid;
    const bool valid = item < n;
    const uint32_t i = valid ? q[item] : 0u;
    // the whole slot state in one round trip
    uint4 pa1 = make_uint4(0, 0, 0, 0);
    float4 thr4 = make_float4(0.f, 0.f, 0.f, 0.f), o4 = thr4, d4 = thr4, h0 = thr4, cp4 = thr4, L4 = thr4, ap4 = thr4;
    uint4 hb = pa1;
    if (valid) {
      thr4 = ldg4(&st.slot[i].path.thr_eta);
      pa1 = ldg4u(&st.slot[i].path.rng);
      o4 = ldg4(&st.slot[i].ray.o_w);
      d4 = ldg4(&st.slot[i].ray.d_s);
      h0 = ldg4(&st.slot[i].hit.h0);
      hb = ldg4u(&st.slot[i].hit.h1);
      cp4 = ldg4(&st.slot[i].nee.c_pdf);
      L4 = ldg4(&st.slot[i].pix.L_s);
      ap4 = ldg4(&st.slot[i].pix.acc_pixel);
    }
    EndIO io;
    io.thr4 = thr4;
    io.o4 = o4;
    io.d4 = d4;
    io.h0 = h0;
    io.cp4 = cp4;
    io.L4 = L4;
    io.ap4 = ap4;
    io.pa1 = pa1;
    io.hb = hb;
    end_item<INTEGRATOR>(sc, args, st.next_pixel, out, valid, exhausted, lane, lane_lt, nan_samples, n_samples, io);
    const bool new_ray = io.new_ray;
    if (new_ray) {
      stg4(&st.slot[i].ray.o_w, io.cam_o);
      stg4(&st.slot[i].ray.d_s, io.cam_d);
    }
    if (valid) {
      // path: throughput 1, eta_scale 1, d 0; nee: last_pdf 0; pix: L 0, s, acc, pixel
      stg4(&st.slot[i].path.thr_eta, make_float4(1.f, 1.f, 1.f, 1.f));
      __stcg(reinterpret_cast<uint4*>(&st.slot[i].path.rng), io.rng_fd);
      stg4(&st.slot[i].nee.c_pdf, make_float4(0.f, 0.f, 0.f, 0.f));
      stg4(&st.slot[i].pix.L_s, io.L_s);
      stg4(&st.slot[i].pix.acc_pixel, io.acc_pixel);
    }

    const bool want[1] = {new_ray};
    QwAppend<1> ap;
    ap.count(s_cnt, want, lane, lane_lt);
    __syncthreads();
    if (tid == 0 && s_cnt[0]) s_base[0] = atomicAdd(&c->n_closest[parity ^ 1u], s_cnt[0]);
    __syncthreads();
    if (new_ray) st.q_closest[parity ^ 1u][s_base[0] + ap.off[0]] = i;
  }
  warp_add_counter(&counters->nan_samples, nan_samples);
  warp_add_counter(&counters->samples, n_samples);
}

// ================================================================== SHADE ====
template <uint32_t CLASS>
VD bool cls_sample(const DScene& sc, const vimg_material& m, V3 wi, const HitInfo& hit, Pcg32& rng,
                   bool regularize, ScatterInfo& out) {
  if constexpr (CLASS == WF_CLASS_LAMBERT) return lambertian_sample(wi, hit, rng, out);
  else if constexpr (CLASS == WF_CLASS_DIELECTRIC) return dielectric_sample(m, wi, hit, rng, out);
  else return principled_sample(sc, m, wi, hit, rng, regularize, out);
}
template <uint32_t CLASS>
VD void cls_eval_pdf_pair(const DScene& sc, const vimg_material& m, V3 wi, V3 wo, const HitInfo& hit,
                          RayCone cone, bool regularize, V3& eval, float& pdf) {
  if constexpr (CLASS == WF_CLASS_LAMBERT) {
    lambertian_eval_pdf(sc, m, wi, wo, hit, cone, eval, pdf);
  } else if constexpr (CLASS == WF_CLASS_PRINCIPLED) {
    principled_eval_pdf(sc, m, wi, wo, hit, cone, regularize, eval, pdf);
  } else {  // Dielectric inherits the base (0, 1) (material.h:56-60)
    eval = mk3(0.f);
    pdf = 1.0f;
  }
}
template <uint32_t CLASS>
VD V3 cls_eval_div_pdf(const DScene& sc, const vimg_material& m, V3 wi, V3 wo, const HitInfo& hit,
                       RayCone cone, bool regularize) {
  if constexpr (CLASS == WF_CLASS_LAMBERT) {
    return col_at_ray_hit(sc, m.texture, wi, cone, hit);
  } else if constexpr (CLASS == WF_CLASS_DIELECTRIC) {
    return mk3(1.f);
  } else {
    V3 e;
    float p;
    principled_eval_pdf(sc, m, wi, wo, hit, cone, regularize, e, p);
    return e / p;
  }
}

// vimg_stage of a material class's shade launches
VD constexpr uint32_t wf_class_stage(uint32_t cls) {
  return cls == WF_CLASS_PRINCIPLED ? VIMG_STAGE_SHADE_PRINCIPLED
                                    : (cls == WF_CLASS_LAMBERT ? VIMG_STAGE_SHADE_LAMBERTIAN : VIMG_STAGE_SHADE_DIELECTRIC);
}

#ifndef QW_WARP_APPEND_SHADE
#define QW_WARP_APPEND_SHADE 1  // 0: block-aggregated appends behind two barriers
#endif
#ifndef QW_SHADE_MIN_BLOCKS_P
#define QW_SHADE_MIN_BLOCKS_P 2  // Principled
#endif
#ifndef QW_SHADE_MIN_BLOCKS_L
#define QW_SHADE_MIN_BLOCKS_L 2  // Lambertian, Dielectric
#endif

// What the shade kernels read for one item and what they leave behind; the memory layout of both is the
// kernel's business (slot records here, queue-ordered records in the streamed pipeline of vimg_stream.cuh).
struct ShadeIO {
  // in: path state, ray, hit, pending NEE contribution, bounce_result so far (the records of the hit primitive are
  // separate arguments of shade_bounce: they hold arrays and live in local memory)
  float4 t4, o4, d4, h0, cp4, L4;  // throughput + eta_scale | origin + cone width | direction + spread | leaf, t, e0, e1 |
                                   // pending contribution + last_pdf | L + sample counter
  uint4 pa1, hb;                   // rng lo, rng hi, flags, depth counter | e2, inv_det, shadow ray occluded, -
  // out
  bool L_changed;    // L4 was updated (pending shadow ray of the last bounce was unoccluded)
  bool ended_early;  // Russian roulette / depth ended the path before shading: only rng_fd changed
  bool shaded;       // thr_eta, rng_fd, new_o, new_d, nee_wi, nee_c are the path's new state
  uint4 rng_fd;
  float4 thr_eta, new_o, new_d, nee_wi, nee_c;
  bool want_closest, want_shadow, want_end;
};

// the body of the bounce loop at one hit (mis_integrator.cpp:45-122, mat_integrator.cpp:21-83)
template <int INTEGRATOR, uint32_t CLASS>
VD void shade_bounce(const DScene& sc, const WfArgs& args, const DLeafPrim& lp, const DShadeRec& sr, ShadeIO& io) {
  constexpr bool MIS = INTEGRATOR == VIMG_INTEGRATOR_MIS;
  float4 t4 = io.t4, L4 = io.L4;
  const float4 o4 = io.o4, d4 = io.d4, h0 = io.h0, cp4 = io.cp4;
  const uint4 pa1 = io.pa1, hb = io.hb;
  bool want_closest = false, want_shadow = false, want_end = false;
  bool go = true;
  io.L_changed = io.ended_early = io.shaded = false;
  uint32_t fj = pa1.z, d = pa1.w;
  Pcg32 rng;
  rng.state = ((unsigned long long)pa1.y << 32) | pa1.x;
  rng.inc = ((uint64_t)(fj >> WF_RANGE_SHIFT) << 1) | 1u;
  if (go) {
    if (MIS) {
      // what mis_integrator does between the BVH::hit that found this surface and the next
      // loop body: the shadow ray of the previous bounce (mis_integrator.cpp:64-78), Russian
      // roulette (:154-167) and the loop condition d < depth
#ifdef VIMG_DEBUG_NEE
      if (fj & WF_NEE) printf("D shade: pending nee valid=%d occl=%u c=%.9g %.9g %.9g\n", (fj & WF_NEE_VALID) != 0, hb.z, cp4.x, cp4.y, cp4.z);
#endif
      if ((fj & WF_NEE) && (fj & WF_NEE_VALID) && hb.z == 0) {
        L4.x += cp4.x;
        L4.y += cp4.y;
        L4.z += cp4.z;
        io.L_changed = true;
      }
      fj &= ~(WF_NEE | WF_NEE_VALID);
      bool go_on = true;
      if (!(fj & WF_FIRST)) {
        if (d > 5) {
          V3 t = xyz(t4);
          go_on = russian_roulette(rng, t4.w, t);
          if (go_on) t4 = f4(t, t4.w);
        }
        d = d + 1;
      }
      fj &= ~WF_FIRST;
      if (!go_on || d >= args.depth) {
        // the path returns bounce_result: the next END launch finishes the sample
        fj |= WF_END;
        io.rng_fd = make_uint4((uint32_t)rng.state, (uint32_t)(rng.state >> 32), fj, d);
        io.ended_early = true;
        want_end = true;
        go = false;
      }
    } else {
      fj &= ~WF_FIRST;
    }
  }
  if (go) {
    Ray ray;
    ray.o = xyz(o4);
    ray.dir = xyz(d4);
    ray.minT = 0.0001f;
    ray.maxT = h0.y;
    ray.cone.cone_width = o4.w;
    ray.cone.spread_angle = d4.w;
    PreHit ph;
    ph.leaf_slot = __float_as_int(h0.x);
    ph.e0 = h0.z;
    ph.e1 = h0.w;
    ph.e2 = __uint_as_float(hb.x);
    ph.inv_det = __uint_as_float(hb.y);
    // the tangent frame is read by the Principled BSDF only
    const HitInfo hit = wf_hit_info<CLASS == WF_CLASS_PRINCIPLED>(sc, lp, sr, ph, ray);
    const vimg_material& mat = sc.materials[hit.mat];
    V3 thr = xyz(t4);
    float eta_scale = t4.w;
    bool nsb = (fj & WF_NSB) != 0;
    // ray cones: skipped when nothing in the scene reads them (DScene::uses_cones)
    const bool cones = !VIMG_LEAN && sc.uses_cones != 0;
    float hit_dist = cones ? length(ray.o - hit.hit_p) : 0.f;
    float surface_spread_angle
        = cones ? spread_angle_from_curvature(hit.mean_curvature, ray.cone.cone_width, ray.dir, hit.hit_n_s) : 0.f;
    // records written back at the end (whole sectors)
    float4 nee_wi = make_float4(0.f, 0.f, 0.f, 0.f), nee_c = nee_wi;
    float4 new_o = f4(hit.hit_p, o4.w), new_d = d4;  // origin of the shadow ray and of the next ray

    if (MIS) {
      // ---- mis_integrator.cpp:45-117 at `hit` ----
      bool nee = false, nee_valid = false;
      EmitterInfo li;
      V3 light_col = mk3(0.f);
      if (CLASS != WF_CLASS_DIELECTRIC) {  // !mat->is_delta()
        light_col = lights_sample(sc, hit.hit_p, rng, li);
        nee = li.pdf != 0.f;
      }
      const RayCone nee_cone = cones ? propagate_reflect_cone(ray.cone, surface_spread_angle * 2.f, hit_dist) : ray.cone;
      const bool nee_regularize = nsb;
      ScatterInfo sm;
      const bool scattered = cls_sample<CLASS>(sc, mat, ray.dir, hit, rng, nsb, sm);
      RayCone cone = ray.cone;
      if (scattered) {
        if (!sm.is_specular) nsb = true;
        if (sm.eta != 0.f) {
          eta_scale /= (sm.eta * sm.eta);
          if (cones) cone = propagate_refract_cone(ray.cone, ray.dir, surface_spread_angle, sm.eta, sm.wo);
        } else {
          cone = nee_cone;
        }
      }
      // ONE inlined eval_pdf_pair site serves both calls of the loop body:
      //   q = 0 light direction (mis_integrator.cpp:68-71), q = 1 sampled direction (:104-106)
      V3 contrib = mk3(0.f), mat_sample_eval = mk3(0.f);
      float mat_sample_pdf = 0.f;
#pragma unroll 1
      for (int qq = 0; qq < 2; ++qq) {
        if (!(qq == 0 ? nee : scattered)) continue;
        V3 e;
        float p;
        cls_eval_pdf_pair<CLASS>(sc, mat, ray.dir, qq == 0 ? li.wi : sm.wo, hit, qq == 0 ? nee_cone : cone,
                                 qq == 0 ? nee_regularize : nsb, e, p);
        if (qq == 0) {
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
#ifdef VIMG_DEBUG_NEE
      printf("D bounce d=%u class=%u hp=%.9g %.9g %.9g n=%.9g %.9g %.9g nee=%d wi=%.9g %.9g %.9g lpdf=%.9g dist=%g valid=%d c=%.9g %.9g %.9g scat=%d wo=%.9g %.9g %.9g mpdf=%.9g thr=%.9g %.9g %.9g\n",
             d, CLASS, hit.hit_p.x, hit.hit_p.y, hit.hit_p.z, hit.hit_n_s.x, hit.hit_n_s.y, hit.hit_n_s.z, (int)nee, li.wi.x, li.wi.y, li.wi.z, li.pdf, li.dist, (int)nee_valid,
             contrib.x, contrib.y, contrib.z, (int)scattered, sm.wo.x, sm.wo.y, sm.wo.z, mat_sample_pdf, thr.x, thr.y, thr.z);
#endif
      fj &= ~(WF_NSB | WF_NEE | WF_NEE_VALID | WF_END);
      if (nsb) fj |= WF_NSB;
      if (nee) {
        fj |= WF_NEE | (nee_valid ? WF_NEE_VALID : 0u);
        nee_wi = f4(li.wi, li.dist);
        nee_c = f4(contrib, 0.f);
        want_shadow = true;
      }
      if (!scattered || is_nan(mat_sample_pdf)) {
        fj |= WF_END;  // return bounce_result (after the pending shadow ray)
        want_end = true;
      } else {
        thr = thr * (mat_sample_eval / mat_sample_pdf);
        nee_c.w = mat_sample_pdf;  // last_pdf
        new_o.w = cone.cone_width;
        new_d = f4(sm.wo, cone.spread_angle);
        want_closest = true;
      }
    } else {
      // ---- material_integrator, mat_integrator.cpp:21-83 at `hit` ----
      V3 emitted_col = mat_emitted(mat, ray.dir, hit.hit_n_s);
      ScatterInfo sr;
      fj &= ~(WF_NSB | WF_END);
      if (!cls_sample<CLASS>(sc, mat, ray.dir, hit, rng, nsb, sr)) {
        nee_c = f4(thr * emitted_col, 0.f);
        fj |= WF_END;
        want_end = true;
      } else {
        if (!sr.is_specular) nsb = true;
        RayCone cone;
        if (sr.eta != 0.f) {
          eta_scale /= (sr.eta * sr.eta);
          cone = cones ? propagate_refract_cone(ray.cone, ray.dir, surface_spread_angle, sr.eta, sr.wo) : ray.cone;
        } else {
          cone = cones ? propagate_reflect_cone(ray.cone, surface_spread_angle * 2.f, hit_dist) : ray.cone;
        }
        thr = thr * (emitted_col + cls_eval_div_pdf<CLASS>(sc, mat, ray.dir, sr.wo, hit, cone, nsb));
        bool go_on = true;
        if (d > 5) go_on = russian_roulette(rng, eta_scale, thr);
        if (!go_on || d + 1 >= args.depth) {
          // `break`, or the for loop runs out before the next BVH::hit: return vec3(0)
          nee_c = make_float4(0.f, 0.f, 0.f, 0.f);
          fj |= WF_END;
          want_end = true;
        } else {
          new_o.w = cone.cone_width;
          new_d = f4(sr.wo, cone.spread_angle);
          d = d + 1;
          want_closest = true;
        }
      }
      if (nsb) fj |= WF_NSB;
    }
    io.thr_eta = f4(thr, eta_scale);
    io.rng_fd = make_uint4((uint32_t)rng.state, (uint32_t)(rng.state >> 32), fj, d);
    io.new_o = new_o;
    io.new_d = new_d;
    io.nee_wi = nee_wi;
    io.nee_c = nee_c;
    io.shaded = true;
  }

  io.L4 = L4;
  io.want_closest = want_closest;
  io.want_shadow = want_shadow;
  io.want_end = want_end;
}

// thread <-> item of the class's shade queue: the body of the bounce loop at one hit
// (mis_integrator.cpp:45-122, mat_integrator.cpp:21-83)
template <int INTEGRATOR, uint32_t CLASS>
__global__ void __launch_bounds__(QW_THREADS, CLASS == WF_CLASS_PRINCIPLED ? QW_SHADE_MIN_BLOCKS_P : QW_SHADE_MIN_BLOCKS_L)
qw_shade_kernel(const __grid_constant__ DScene sc, const __grid_constant__ QwState st,
                const __grid_constant__ WfArgs args, const uint32_t parity, DCounters* __restrict__ counters) {
  constexpr bool MIS = INTEGRATOR == VIMG_INTEGRATOR_MIS;
  __shared__ uint32_t s_cnt[3], s_base[3];
  QwCounters* c = st.cnt;
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  const uint32_t n = c->n_shade[parity][CLASS];
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&counters->stage_items[wf_class_stage(CLASS)], (unsigned long long)n);
  const uint2* __restrict__ q = st.q_shade[CLASS];

  for (uint32_t base = blockIdx.x * QW_THREADS; base < n; base += gridDim.x * QW_THREADS) {
    if (!QW_WARP_APPEND_SHADE) {
      if (tid < 3) s_cnt[tid] = 0;
      __syncthreads();
    }
    const uint32_t item = base + tid;
    bool want_closest = false, want_shadow = false, want_end = false;
    bool go = item < n;
    const uint2 qi = go ? q[item] : make_uint2(0u, 0u);
    const uint32_t j = qi.x;
    // the whole slot state and the leaf record of the hit in one round trip
    float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f), o4 = t4, d4 = t4, h0 = t4, cp4 = t4, L4 = t4;
    uint4 pa1 = make_uint4(0, 0, 0, 0), hb = pa1;
    DLeafPrim lp;
    DShadeRec sr;
    if (go) {
      {
        const float4* rec = reinterpret_cast<const float4*>(sc.leaf_prims + qi.y);
        float4* dst = reinterpret_cast<float4*>(&lp);
        dst[0] = __ldg(rec);
        dst[1] = __ldg(rec + 1);
        dst[2] = __ldg(rec + 2);
      }
      sr = load_shade_rec(sc, (int32_t)qi.y);
      t4 = ldg4(&st.slot[j].path.thr_eta);
      pa1 = ldg4u(&st.slot[j].path.rng);
      o4 = ldg4(&st.slot[j].ray.o_w);
      d4 = ldg4(&st.slot[j].ray.d_s);
      h0 = ldg4(&st.slot[j].hit.h0);
      hb = ldg4u(&st.slot[j].hit.h1);
      if (MIS) {
        cp4 = ldg4(&st.slot[j].nee.c_pdf);
        L4 = ldg4(&st.slot[j].pix.L_s);
      }
    }
    ShadeIO io;
    io.t4 = t4;
    io.o4 = o4;
    io.d4 = d4;
    io.h0 = h0;
    io.cp4 = cp4;
    io.L4 = L4;
    io.pa1 = pa1;
    io.hb = hb;
    io.want_closest = io.want_shadow = io.want_end = false;
    if (go) {
      shade_bounce<INTEGRATOR, CLASS>(sc, args, lp, sr, io);
      if (io.L_changed) stg4(&st.slot[j].pix.L_s, io.L4);
      if (io.ended_early) __stcg(reinterpret_cast<uint4*>(&st.slot[j].path.rng), io.rng_fd);
      if (io.shaded) {
        stg4(&st.slot[j].path.thr_eta, io.thr_eta);
        __stcg(reinterpret_cast<uint4*>(&st.slot[j].path.rng), io.rng_fd);
        stg4(&st.slot[j].ray.o_w, io.new_o);
        stg4(&st.slot[j].ray.d_s, io.new_d);
        stg4(&st.slot[j].nee.wi_dist, io.nee_wi);
        stg4(&st.slot[j].nee.c_pdf, io.nee_c);
      }
    }
    want_closest = io.want_closest;
    want_shadow = io.want_shadow;
    want_end = io.want_end;

    // every queue written here belongs to the NEXT iteration, so this launch is independent of
    // the END launch of the same iteration and the two run concurrently
    // a path that goes on after requesting a shadow ray: ONE trace item, the shadow ray first and the
    // closest ray after it on the same lane (QW_PAIR_BIT), because the reference's closest-hit
    // traversal inherits the stack its occlusion traversal left behind (see bvh_traverse)
    const bool pair = want_closest && want_shadow && !args.fresh_stack;
    if (pair) want_shadow = false;
    uint32_t pos_c, pos_s, pos_e;
    if (QW_WARP_APPEND_SHADE) {
      // one atomic instruction per warp: lane k reserves the warp's entries of queue k (closest, shadow, END; their
      // counters lie 8 bytes apart) -- no block barrier, one round trip (the order of a queue is not part of any result)
      const uint32_t m0 = __ballot_sync(0xffffffffu, want_closest), m1 = __ballot_sync(0xffffffffu, want_shadow),
                     m2 = __ballot_sync(0xffffffffu, want_end);
      const uint32_t mine = lane == 0 ? m0 : (lane == 1 ? m1 : (lane == 2 ? m2 : 0u));
      uint32_t b = 0;
      if (mine) b = atomicAdd(&c->n_closest[0] + 2u * lane + (parity ^ 1u), (uint32_t)__popc(mine));
      pos_c = __shfl_sync(0xffffffffu, b, 0) + __popc(m0 & lane_lt);
      pos_s = __shfl_sync(0xffffffffu, b, 1) + __popc(m1 & lane_lt);
      pos_e = __shfl_sync(0xffffffffu, b, 2) + __popc(m2 & lane_lt);
    } else {
      const bool want[3] = {want_closest, want_shadow, want_end};
      QwAppend<3> ap;
      ap.count(s_cnt, want, lane, lane_lt);
      __syncthreads();
      if (tid == 0 && s_cnt[0]) s_base[0] = atomicAdd(&c->n_closest[parity ^ 1u], s_cnt[0]);
      if (tid == 1 && s_cnt[1]) s_base[1] = atomicAdd(&c->n_shadow[parity ^ 1u], s_cnt[1]);
      if (tid == 2 && s_cnt[2]) s_base[2] = atomicAdd(&c->n_end[parity ^ 1u], s_cnt[2]);
      __syncthreads();
      pos_c = s_base[0] + ap.off[0];
      pos_s = s_base[1] + ap.off[1];
      pos_e = s_base[2] + ap.off[2];
    }
    if (want_closest) st.q_closest[parity ^ 1u][pos_c] = pair ? (j | QW_PAIR_BIT) : j;
    if (want_shadow) st.q_shadow[parity ^ 1u][pos_s] = j;
    if (want_end) st.q_end[parity ^ 1u][pos_e] = j;
  }
}


// The kernels of one build of this header.  The library holds two: the bit-exact build (vimg_api.cu:
// -fmad=false, IEEE division / sqrt, glibc's libm restated) and the tolerance build (vimg_fast.cu: the
// same sources with -use_fast_math, float where the reference promotes to double, plus the kernels
// above that the bit-exact contract rules out).  Launches go through cudaLaunchKernel on these entries.
struct QwKernelTable {
  const void* trace[2];       // [COUNT]
  const void* classify;
  const void* shade[2][WF_NUM_CLASSES];  // [material = 0, mis = 1][class]
  const void* end[2];                    // [material = 0, mis = 1]
  // tolerance build only: flat re-layout of a small scene (vimg_flat.cuh) and the streamed pipeline (vimg_stream.cuh)
  const void* build_flat;
  const void* sv_shade[2][WF_NUM_CLASSES];
  const void* sv_end[2];
  const void* sv_init;
  // 4-wide BVH for large scenes (vimg_wide.cuh)
  const void* wide_build;
  const void* trace_wide[2];  // [COUNT]
};
inline QwKernelTable qw_kernel_table() {
  QwKernelTable t{};
  t.trace[0] = (const void*)&qw_trace_kernel<false>;
  t.trace[1] = (const void*)&qw_trace_kernel<true>;
  t.classify = (const void*)&qw_classify_kernel;
  t.shade[0][WF_CLASS_LAMBERT] = (const void*)&qw_shade_kernel<VIMG_INTEGRATOR_MATERIAL, WF_CLASS_LAMBERT>;
  t.shade[0][WF_CLASS_PRINCIPLED] = (const void*)&qw_shade_kernel<VIMG_INTEGRATOR_MATERIAL, WF_CLASS_PRINCIPLED>;
  t.shade[0][WF_CLASS_DIELECTRIC] = (const void*)&qw_shade_kernel<VIMG_INTEGRATOR_MATERIAL, WF_CLASS_DIELECTRIC>;
  t.shade[1][WF_CLASS_LAMBERT] = (const void*)&qw_shade_kernel<VIMG_INTEGRATOR_MIS, WF_CLASS_LAMBERT>;
  t.shade[1][WF_CLASS_PRINCIPLED] = (const void*)&qw_shade_kernel<VIMG_INTEGRATOR_MIS, WF_CLASS_PRINCIPLED>;
  t.shade[1][WF_CLASS_DIELECTRIC] = (const void*)&qw_shade_kernel<VIMG_INTEGRATOR_MIS, WF_CLASS_DIELECTRIC>;
  t.end[0] = (const void*)&qw_end_kernel<VIMG_INTEGRATOR_MATERIAL>;
  t.end[1] = (const void*)&qw_end_kernel<VIMG_INTEGRATOR_MIS>;
  return t;
}

}  // namespace vimg
