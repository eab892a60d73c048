// vimg_stream.cuh -- the STREAMED queue wavefront (tolerance build, flat scenes): SoA ray / hit queues.
//
// What ncu showed about the slot-record pipeline of vimg_queues.cuh (profiles/r5_*): its kernels gather path
// state by slot number -- 32-byte sectors at random places of a 335 MB array -- and the memory system fetches
// them at 128-byte granularity.  A flat-list trace kernel fed from slots needed 32 - 48 bytes per ray and moved 218
// (dram__bytes_read 524 MB for 2.4 M rays); the shade / END kernels moved 2.2 - 2.6x their sector bytes and every
// item costs two dependent DRAM round trips (queue entry -> slot).  With the traversal itself reduced to a
// divergence-free loop, that traffic is the frame time.
//
// Here everything that flows from one kernel to the next flows through QUEUE-ORDERED records, written with
// warp- or block-aggregated appends and read back fully coalesced, 16 bytes per lane and array:
//   hit queues  (-> shade per material class) and the END queue, double buffered by iteration parity:
//                                        ha = slot, leaf record, t, occluded | hb = e0, e1, e2, inv_det   (32 B)
// and what stays with the path lives in ONE 128-byte, 128-byte-aligned line per slot (QvLine): throughput, RNG,
// flags, bounce_result, accumulator, the ray (hit-info needs it), pending NEE contribution.  One line fetch per
// shade / END item instead of two; full-sector writes only.
//
// There is no ray queue and no trace kernel: the kernel that makes a path segment's rays -- shade: closest ray and / or
// shadow ray, END: the camera ray of the slot's next sample -- traces them itself over the flat lists in shared memory
// (qv_trace_route) and appends the hit record to the queue of the hit material's class, or to the END queue.  The
// launches of an iteration read the queues of its parity and fill those of the other one, so they run side by side.
//
// Per path, the estimator, the RNG stream and its draw order are those of vimg_queues.cuh (the shade and END
// bodies are the same functions: shade_bounce, end_item).
#pragma once
#include "vimg_queues.cuh"
#ifdef VIMG_FAST
#include "vimg_flat.cuh"
#endif

namespace vimg {

#ifndef QV_LINE_BYTES
#define QV_LINE_BYTES 128  // 256 / 512: the address-range experiment of DESIGN.md 5.1c (same slots, 2x / 4x the range)
#endif
struct __align__(QV_LINE_BYTES) QvLine {
  float4 thr_eta;    // sector 0: throughput rgb, eta_scale
  uint4 rng_fd;      //           rng lo, rng hi, flags | sample range << 16, depth counter
  float4 L_s;        // sector 1: bounce_result rgb, sample counter
  float4 acc_pixel;  //           pixel accumulator rgb, local pixel number
  float4 o_w;        // sector 2: ray origin, cone width
  float4 d_s;        //           ray direction, cone spread
  float4 c_pdf;      // sector 3: pending NEE contribution rgb, last_pdf
  float4 spare;
};
static_assert(sizeof(QvLine) == QV_LINE_BYTES, "one line per slot");

struct QvQueueCounters {  // the queues one iteration fills (16 bytes: the host clears them with one memset)
  uint32_t n_end;
  uint32_t n_shade[WF_NUM_CLASSES];
};
struct QvCounters {
  QvQueueCounters q[2];  // by iteration parity
  uint32_t pad_[8];
};

// Streamed shade / END launches: 128-thread blocks pack the register file in finer steps than 256-thread ones
// (measured, disney_spheres 256 spp: 231.9 -> 226.4 ms at the same registers per thread); the Principled kernel
// gets 168 registers (3 blocks per SM: no spills, 224.3 ms).  More warps at fewer registers lose: Lambertian at 96
// registers (5 blocks, 178 B of spills) 235.8 ms, at 80 registers 242.4 ms.
#ifndef QV_THREADS
#define QV_THREADS 128
#endif
#ifndef QV_SHADE_MIN_BLOCKS_L
#define QV_SHADE_MIN_BLOCKS_L 4  // Lambertian, Dielectric: 128 registers
#endif
#ifndef QV_SHADE_MIN_BLOCKS_P
#define QV_SHADE_MIN_BLOCKS_P 3  // Principled: 168 registers
#endif
#ifndef QV_END_MIN_BLOCKS
#define QV_END_MIN_BLOCKS 6      // 80 registers
#endif
#define QV_HAS_CLOSEST 0x40000000u
#define QV_HAS_SHADOW 0x80000000u
#define QV_SLOT_MASK 0x3fffffffu

struct QvState {
  uint32_t n_slots;
  QvLine* line;
  float4 *ha[2][WF_NUM_CLASSES], *hb[2][WF_NUM_CLASSES];  // hit queues of the material classes, by iteration parity
  float4 *ea[2], *eb[2];                                  // END queue
  QvCounters* cnt;
  uint32_t* next_pixel;
};

#ifdef VIMG_FAST
#ifndef QV_PREFETCH
#define QV_PREFETCH 1  // 1: shade / END start the fetch of their NEXT item's line into L2 while they work on this one
#endif
// Queue appends: block-aggregated (one global atomic per block and queue, two __syncthreads per batch of items) or, 1,
// per warp (one atomic per warp and queue, no barrier: a block no longer waits for the slowest of its warps and for the
// round trip of one thread's atomic).  Measured (disney_spheres, 256 spp): the shade kernels gain with warp appends
// (Lambertian 118.6 -> 114.0 us per launch), END -- which also takes pixels from a global counter -- does not
// (86.6 -> 90.0 us; with the four-queue routing of qv_trace_route: frame 188.4 -> 189.9 ms).
#ifndef QV_WARP_APPEND
#define QV_WARP_APPEND 1        // shade
#endif
#ifndef QV_WARP_APPEND_END
#define QV_WARP_APPEND_END 0
#endif
VD void qv_prefetch_line(const QvLine* ln) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ln)); }
VD float4 ldq(const float4* p) { return __ldcs(p); }  // queue records are read once: evict first
VD void stq(float4* p, float4 v) { __stcs(p, v); }

// ============================================== the traces of a path segment, at the tail of its producer ====
// The kernel that made a path segment's rays -- a shade kernel (closest ray and / or shadow ray) or END (camera ray)
// -- traces them itself over the flat lists in shared memory and routes the result: the hit record goes straight
// into the hit queue of the hit material's class, or into the END queue (miss, emitter, base material, a path that
// only had its last shadow ray left, a path that ended with nothing pending).  There is no ray queue and no trace
// launch: the divergence-free test loop (issue bound as a kernel of its own, 65 % of the issue slots) fills the slots
// the shade / END bodies leave idle while they wait for their gathers (23 - 35 %), and a path segment costs one
// kernel instead of two.
//   valid: this lane has an item to route;  has_c / has_s: it has a closest / shadow ray to trace first
//   o: origin of both rays, d: closest-ray direction, w: shadow-ray direction + distance to the light sample
template <bool WARP_APPEND>
VD void qv_trace_route(const QfScene& fs, const QvState& st, QvCounters* c, const uint32_t q, uint32_t* s_cnt,
                       uint32_t* s_base, const bool valid, const bool has_c, const bool has_s, const V3 o, const V3 d,
                       const float4 w, const uint32_t slot, const int tid, const uint32_t lane, const uint32_t lane_lt,
                       uint32_t& rays_closest, uint32_t& rays_shadow) {
  // shadow ray (mis_integrator.cpp:59-64): origin = hit point, maxT = dist - 1e-4
  QfBest sh, best;
  sh.t = has_s ? w.w - 0.0001f : 0.f;
  sh.k = -1;
  sh.u = sh.v = 0.f;
  best.t = has_c ? CUDART_INF_F : 0.f;
  best.k = -1;
  best.u = best.v = 0.f;
  const bool any_s = __any_sync(0xffffffffu, has_s), any_c = __any_sync(0xffffffffu, has_c);
  if (any_s && any_c)
    qf_trace2(fs, o, d, xyz(w), 0.0001f, best, sh);  // both rays in one pass over the lists
  else if (any_c)
    qf_trace(fs, o, d, 0.0001f, best);
  else if (any_s)
    qf_trace(fs, o, xyz(w), 0.0001f, sh);
  const bool occl = sh.k >= 0;
  rays_shadow += has_s;
  rays_closest += has_c;
  uint32_t cls = WF_CLASS_NONE + 1;  // NONE + 1: no item
  int32_t leaf = -1;
  float4 hb4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
    cls = WF_CLASS_NONE;  // END
    if (best.k >= 0) {
      uint32_t lc;
      qf_decode(fs, best, leaf, lc, hb4);
      cls = (lc & VIMG_LEAF_CLASS_EMISSIVE) ? WF_CLASS_NONE : (lc & 3u);
    }
  }
  uint32_t pos = 0;
  if (WARP_APPEND) {
    // one atomic instruction for the warp: lane k reserves the warp's items of queue k (0 = END, 1 + class), so the
    // four reservations are one round trip instead of four dependent ones
    static_assert(WF_NUM_CLASSES == 3 && WF_CLASS_NONE == 3, "queue index = (class + 1) & 3");
    const uint32_t qi = (cls + 1u) & 3u;  // END -> 0, class k -> k + 1 (no item: cls = 4 -> 1, masked out by the ballots)
    uint32_t mine = 0, m_own = 0;
#pragma unroll
    for (uint32_t k = 0; k <= WF_NUM_CLASSES; ++k) {
      const uint32_t m = __ballot_sync(0xffffffffu, valid && qi == k);
      if (lane == k) mine = m;
      if (qi == k) m_own = m;
    }
    uint32_t base = 0;
    if (lane <= WF_NUM_CLASSES && mine) base = atomicAdd(&c->q[q].n_end + lane, (uint32_t)__popc(mine));
    base = __shfl_sync(0xffffffffu, base, qi);
    pos = base + __popc(m_own & lane_lt);
  } else {
    bool want[WF_NUM_CLASSES + 1];
#pragma unroll
    for (uint32_t k = 0; k <= WF_NUM_CLASSES; ++k) want[k] = cls == k;
    QwAppend<WF_NUM_CLASSES + 1> ap;
    ap.count(s_cnt, want, lane, lane_lt);
    __syncthreads();
    if (tid < (int)WF_NUM_CLASSES) {
      if (s_cnt[tid]) s_base[tid] = atomicAdd(&c->q[q].n_shade[tid], s_cnt[tid]);
    } else if (tid == (int)WF_NUM_CLASSES) {
      if (s_cnt[tid]) s_base[tid] = atomicAdd(&c->q[q].n_end, s_cnt[tid]);
    }
    __syncthreads();
    if (valid) pos = s_base[cls] + ap.off[cls];
  }
  if (valid) {
    const float4 ha4 = make_float4(__uint_as_float(slot), __int_as_float(leaf), best.k >= 0 ? best.t : CUDART_INF_F,
                                   __uint_as_float(occl ? 1u : 0u));
    float4* qa = cls < WF_NUM_CLASSES ? st.ha[q][cls] : st.ea[q];
    float4* qb = cls < WF_NUM_CLASSES ? st.hb[q][cls] : st.eb[q];
    stq(qa + pos, ha4);
    stq(qb + pos, hb4);
  }
}
VD void qv_stage_flat(QfScene& fs, const QfScene* __restrict__ flat, int tid) {
  const uint4* src = reinterpret_cast<const uint4*>(flat);
  uint4* dst = reinterpret_cast<uint4*>(&fs);
  for (uint32_t i = tid; i < sizeof(QfScene) / 16; i += QV_THREADS) dst[i] = __ldg(src + i);
  __syncthreads();  // the staged lists are visible
}

// ====================================================================== SHADE, streamed ====
// thread <-> item of the class's hit queue; reads the queues of `parity`, fills those of parity ^ 1
template <int INTEGRATOR, uint32_t CLASS>
__global__ void __launch_bounds__(QV_THREADS, CLASS == WF_CLASS_PRINCIPLED ? QV_SHADE_MIN_BLOCKS_P : QV_SHADE_MIN_BLOCKS_L)
qv_shade_kernel(const __grid_constant__ DScene sc, const __grid_constant__ QvState st,
                const __grid_constant__ WfArgs args, const uint32_t parity, const QfScene* __restrict__ flat,
                const uint32_t count, DCounters* __restrict__ counters) {
  __shared__ QfScene fs;
  __shared__ uint32_t s_cnt[WF_NUM_CLASSES + 1], s_base[WF_NUM_CLASSES + 1];
  QvCounters* c = st.cnt;
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  qv_stage_flat(fs, flat, tid);
  const uint32_t n = c->q[parity].n_shade[CLASS];
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&counters->stage_items[wf_class_stage(CLASS)], (unsigned long long)n);
  const float4* __restrict__ qa = st.ha[parity][CLASS];
  const float4* __restrict__ qb = st.hb[parity][CLASS];
  uint32_t rays_closest = 0, rays_shadow = 0;

  for (uint32_t base = blockIdx.x * QV_THREADS; base < n; base += gridDim.x * QV_THREADS) {
    if (!QV_WARP_APPEND) {
      if (tid < (int)WF_NUM_CLASSES + 1) s_cnt[tid] = 0;
      __syncthreads();
    }
    const uint32_t item = base + tid;
    const bool go = item < n;
    float4 ha4 = make_float4(0.f, 0.f, 0.f, 0.f), hb4 = ha4;
    if (go) {
      ha4 = ldq(qa + item);
      hb4 = ldq(qb + item);
    }
    const uint32_t j = __float_as_uint(ha4.x);
    const int32_t leaf = __float_as_int(ha4.y);
    QvLine* ln = st.line + j;
    if (QV_PREFETCH) {
      // (holding the next slot number in a register and prefetching after the item is shaded measured slower: spills)
      const uint32_t next = item + gridDim.x * QV_THREADS;
      if (next < n) qv_prefetch_line(st.line + __float_as_uint(__ldg(reinterpret_cast<const float*>(qa + next))));
    }
    ShadeIO io;
    io.t4 = io.o4 = io.d4 = io.cp4 = io.L4 = make_float4(0.f, 0.f, 0.f, 0.f);
    io.pa1 = make_uint4(0, 0, 0, 0);
    io.new_o = io.new_d = io.nee_wi = make_float4(0.f, 0.f, 1.f, 0.f);
    float4 ap4 = io.t4;
    DLeafPrim lp;
    DShadeRec sr;
    if (go) {
      {
        const float4* rec = reinterpret_cast<const float4*>(sc.leaf_prims + leaf);
        float4* dst = reinterpret_cast<float4*>(&lp);
        dst[0] = __ldg(rec);
        dst[1] = __ldg(rec + 1);
        dst[2] = __ldg(rec + 2);
      }
      sr = load_shade_rec(sc, leaf);
      io.t4 = ldg4(&ln->thr_eta);
      io.pa1 = ldg4u(&ln->rng_fd);
      io.L4 = ldg4(&ln->L_s);
      ap4 = ldg4(&ln->acc_pixel);
      io.o4 = ldg4(&ln->o_w);
      io.d4 = ldg4(&ln->d_s);
      io.cp4 = ldg4(&ln->c_pdf);
    }
    io.h0 = make_float4(ha4.y, ha4.z, hb4.x, hb4.y);  // leaf, t, e0, e1
    io.hb = make_uint4(__float_as_uint(hb4.z), __float_as_uint(hb4.w), __float_as_uint(ha4.w), 0u);  // e2, inv_det, occluded
    io.want_closest = io.want_shadow = io.want_end = false;
    io.shaded = io.ended_early = io.L_changed = false;
    if (go) {
      shade_bounce<INTEGRATOR, CLASS>(sc, args, lp, sr, io);
      if (io.L_changed) {  // the whole sector: no partial writes
        stg4(&ln->L_s, io.L4);
        stg4(&ln->acc_pixel, ap4);
      }
      if (io.ended_early) {
        stg4(&ln->thr_eta, io.t4);
        __stcg(&ln->rng_fd, io.rng_fd);
      }
      if (io.shaded) {
        stg4(&ln->thr_eta, io.thr_eta);
        __stcg(&ln->rng_fd, io.rng_fd);
        stg4(&ln->o_w, io.new_o);
        stg4(&ln->d_s, io.new_d);
        stg4(&ln->c_pdf, io.nee_c);
        stg4(&ln->spare, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    // the path segment's rays (closest ray and / or shadow ray), or straight to END (ended, nothing pending)
    const bool to_trace = io.want_closest || io.want_shadow;
    const bool to_end = io.want_end && !io.want_shadow;
    qv_trace_route<QV_WARP_APPEND != 0>(fs, st, c, parity ^ 1u, s_cnt, s_base, to_trace || to_end, io.want_closest, io.want_shadow,
                                        xyz(io.new_o), xyz(io.new_d), io.nee_wi, j, tid, lane, lane_lt, rays_closest, rays_shadow);
  }
  warp_add_counter(&counters->rays_closest, rays_closest);
  warp_add_counter(&counters->rays_shadow, rays_shadow);
  if (count) {
    warp_add_counter(&counters->tri_tests, (rays_closest + rays_shadow) * fs.n_tri);
    warp_add_counter(&counters->sphere_tests, (rays_closest + rays_shadow) * fs.n_sph);
  }
}

// ======================================================================== END, streamed ====
// thread <-> item of the END queue: a path that ended (or a slot that has no pixel yet)
template <int INTEGRATOR>
__global__ void __launch_bounds__(QV_THREADS, QV_END_MIN_BLOCKS)
qv_end_kernel(const __grid_constant__ DScene sc, const __grid_constant__ QvState st,
              const __grid_constant__ WfArgs args, const uint32_t parity, float* __restrict__ out,
              const QfScene* __restrict__ flat, const uint32_t count, DCounters* __restrict__ counters) {
  __shared__ QfScene fs;
  __shared__ uint32_t s_cnt[WF_NUM_CLASSES + 1], s_base[WF_NUM_CLASSES + 1];
  QvCounters* c = st.cnt;
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  qv_stage_flat(fs, flat, tid);
  const uint32_t n = c->q[parity].n_end;
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&counters->stage_items[VIMG_STAGE_END], (unsigned long long)n);
  const float4* __restrict__ qa = st.ea[parity];
  const float4* __restrict__ qb = st.eb[parity];
  uint32_t nan_samples = 0, n_samples = 0, rays_closest = 0, rays_shadow = 0;
  bool exhausted = *reinterpret_cast<volatile uint32_t*>(st.next_pixel) >= args.n_local_units;

  for (uint32_t base = blockIdx.x * QV_THREADS; base < n; base += gridDim.x * QV_THREADS) {
    if (!QV_WARP_APPEND_END) {
      if (tid < (int)WF_NUM_CLASSES + 1) s_cnt[tid] = 0;
      __syncthreads();
    }
    const uint32_t item = base + tid;
    const bool valid = item < n;
    float4 ha4 = make_float4(0.f, 0.f, 0.f, 0.f), hb4 = ha4;
    if (valid) {
      ha4 = ldq(qa + item);
      hb4 = ldq(qb + item);
    }
    const uint32_t i = __float_as_uint(ha4.x);
    QvLine* ln = st.line + i;
    if (QV_PREFETCH) {
      const uint32_t next = item + gridDim.x * QV_THREADS;
      if (next < n) qv_prefetch_line(st.line + __float_as_uint(__ldg(reinterpret_cast<const float*>(qa + next))));
    }
    EndIO io;
    io.thr4 = io.o4 = io.d4 = io.cp4 = io.L4 = io.ap4 = make_float4(0.f, 0.f, 0.f, 0.f);
    io.pa1 = make_uint4(0, 0, 0, 0);
    io.cam_o = io.cam_d = make_float4(0.f, 0.f, 1.f, 0.f);
    if (valid) {
      io.thr4 = ldg4(&ln->thr_eta);
      io.pa1 = ldg4u(&ln->rng_fd);
      io.L4 = ldg4(&ln->L_s);
      io.ap4 = ldg4(&ln->acc_pixel);
      io.o4 = ldg4(&ln->o_w);
      io.d4 = ldg4(&ln->d_s);
      io.cp4 = ldg4(&ln->c_pdf);
    }
    io.h0 = make_float4(ha4.y, ha4.z, hb4.x, hb4.y);
    io.hb = make_uint4(__float_as_uint(hb4.z), __float_as_uint(hb4.w), __float_as_uint(ha4.w), 0u);
    end_item<INTEGRATOR>(sc, args, st.next_pixel, out, valid, exhausted, lane, lane_lt, nan_samples, n_samples, io);
    if (valid) {
      // path: throughput 1, eta_scale 1, d 0; nee: last_pdf 0; pix: L 0, s, acc, pixel
      stg4(&ln->thr_eta, make_float4(1.f, 1.f, 1.f, 1.f));
      __stcg(&ln->rng_fd, io.rng_fd);
      stg4(&ln->L_s, io.L_s);
      stg4(&ln->acc_pixel, io.acc_pixel);
      stg4(&ln->c_pdf, make_float4(0.f, 0.f, 0.f, 0.f));
      stg4(&ln->spare, make_float4(0.f, 0.f, 0.f, 0.f));
      if (io.new_ray) {
        stg4(&ln->o_w, io.cam_o);
        stg4(&ln->d_s, io.cam_d);
      }
    }
    // the camera ray of the slot's next sample: traced here, its hit record routed to the queues of parity ^ 1
    qv_trace_route<QV_WARP_APPEND_END != 0>(fs, st, c, parity ^ 1u, s_cnt, s_base, io.new_ray, io.new_ray, false, xyz(io.cam_o),
                                            xyz(io.cam_d), make_float4(0.f, 0.f, 1.f, 0.f), i, tid, lane, lane_lt, rays_closest,
                                            rays_shadow);
  }
  warp_add_counter(&counters->nan_samples, nan_samples);
  warp_add_counter(&counters->samples, n_samples);
  warp_add_counter(&counters->rays_closest, rays_closest);
  if (count) {
    warp_add_counter(&counters->tri_tests, rays_closest * fs.n_tri);
    warp_add_counter(&counters->sphere_tests, rays_closest * fs.n_sph);
  }
}

// the first END queue: every slot, none with a pixel yet
__global__ void qv_init_kernel(float4* __restrict__ ea, float4* __restrict__ eb, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ea[i] = make_float4(__uint_as_float(i), __int_as_float(-1), CUDART_INF_F, __uint_as_float(0u));
  eb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

inline void qv_fill_kernel_table(QwKernelTable& t) {
  t.sv_shade[0][WF_CLASS_LAMBERT] = (const void*)&qv_shade_kernel<VIMG_INTEGRATOR_MATERIAL, WF_CLASS_LAMBERT>;
  t.sv_shade[0][WF_CLASS_PRINCIPLED] = (const void*)&qv_shade_kernel<VIMG_INTEGRATOR_MATERIAL, WF_CLASS_PRINCIPLED>;
  t.sv_shade[0][WF_CLASS_DIELECTRIC] = (const void*)&qv_shade_kernel<VIMG_INTEGRATOR_MATERIAL, WF_CLASS_DIELECTRIC>;
  t.sv_shade[1][WF_CLASS_LAMBERT] = (const void*)&qv_shade_kernel<VIMG_INTEGRATOR_MIS, WF_CLASS_LAMBERT>;
  t.sv_shade[1][WF_CLASS_PRINCIPLED] = (const void*)&qv_shade_kernel<VIMG_INTEGRATOR_MIS, WF_CLASS_PRINCIPLED>;
  t.sv_shade[1][WF_CLASS_DIELECTRIC] = (const void*)&qv_shade_kernel<VIMG_INTEGRATOR_MIS, WF_CLASS_DIELECTRIC>;
  t.sv_end[0] = (const void*)&qv_end_kernel<VIMG_INTEGRATOR_MATERIAL>;
  t.sv_end[1] = (const void*)&qv_end_kernel<VIMG_INTEGRATOR_MIS>;
  t.sv_init = (const void*)&qv_init_kernel;
}
#endif  // VIMG_FAST

}  // namespace vimg
