// vimg_flat.cuh -- tolerance build only (VIMG_FLAG_FAST): the pipeline for scenes of a few dozen
// primitives, i.e. every scene the reference ships (13 - 18 primitives).
//
// Device-side re-layout of the BVH.  For such a scene the useful re-layout
// of the reference's binary BVH is NO tree: qf_build_kernel turns the leaf array into two flat
// lists -- triangles as three plane equations (the supporting plane and the two barycentric
// planes: t, u, v are three dot products and one division), spheres as centre + radius -- that a
// block keeps in shared memory.  A ray tests every entry.  All lanes of a warp test the SAME entry
// at the same time, the record is one broadcast read, there is no stack, no box test and no
// divergence: the binary traversal of the same scene runs with 16 of 32 lanes per instruction
// (profiles/r4_trace_kernel_fresh_stack.txt), this loop with all of them.
//
// Because a ray's trace no longer diverges, it does not need a kernel of its own.  qf_path_kernel
// (thread <-> path slot) does everything between two BSDF evaluations of a path:
//   pending shadow ray of the last bounce (mis_integrator.cpp:59-78) -> closest-hit ray ->
//   miss / emitter contribution with its MIS weight (:36-38, :173-185, :123-152), sample
//   accumulation, pixel write, next pixel, next camera ray (include/integrators.h:116-137) and that
//   ray's trace -> route the path by the material class it hit into the shade queues.
// It replaces qw_trace_kernel + qw_classify_kernel + qw_end_kernel of the general pipeline: a path
// segment costs two kernels (path, shade) and two round trips of its state through HBM instead
// of four.  Per path the estimator, the RNG stream and the order of its draws are unchanged.
//
// What differs from the reference's traversal, and why this is the tolerance mode only: primitives
// the box tests would have culled are tested too, ties between primitives at the same distance
// resolve in list order, the triangle test is the plane form instead of the watertight one.  Each
// changes a hit only at float-rounding distance from an edge or a tie.
#pragma once
#include "vimg_queues.cuh"

namespace vimg {

#ifndef QF_FLAT_MAX
#define QF_FLAT_MAX 64  // leaf records a flat scene may have
#endif

struct QfScene {
  uint32_t n_tri, n_sph, pad_[2];
  float4 tri[QF_FLAT_MAX * 3];  // {N0, d0} {N1, d1} {N2, d2}: t = (d0 - N0.o) / (N0.dir); u = N1.P + d1; v = N2.P + d2
  float4 sph[QF_FLAT_MAX];      // centre, radius
  uint8_t tri_leaf[QF_FLAT_MAX], sph_leaf[QF_FLAT_MAX];    // leaf record (index into leaf_prims) of each entry
  uint8_t tri_class[QF_FLAT_MAX], sph_class[QF_FLAT_MAX];  // its leaf_class byte
};

// one block; thread <-> leaf record.  Degenerate triangles (triangle.h:89-92) are dropped.
__global__ void qf_build_kernel(const __grid_constant__ DScene sc, QfScene* __restrict__ out) {
  __shared__ uint32_t s_n[2];
  const uint32_t i = threadIdx.x;
  if (i < 2) s_n[i] = 0;
  __syncthreads();
  if (i < sc.n_leaf_recs) {
    const DLeafPrim lp = sc.leaf_prims[i];
    const uint8_t cls = sc.leaf_class[i];
    if (lp.flags & VIMG_LEAF_SPHERE) {
      const uint32_t k = atomicAdd(&s_n[1], 1u);
      out->sph[k] = make_float4(lp.a[0], lp.a[1], lp.a[2], lp.a[3]);
      out->sph_leaf[k] = (uint8_t)i;
      out->sph_class[k] = cls;
    } else if (!(lp.flags & VIMG_LEAF_DEGENERATE)) {
      const V3 p0 = mk3(lp.a[0], lp.a[1], lp.a[2]), p1 = mk3(lp.a[3], lp.a[4], lp.a[5]), p2 = mk3(lp.a[6], lp.a[7], lp.a[8]);
      const V3 e1 = p1 - p0, e2 = p2 - p0;
      const V3 n = cross(e1, e2);
      const float inv_n2 = 1.0f / dot(n, n);
      const V3 n1 = cross(e2, n) * inv_n2, n2 = cross(n, e1) * inv_n2;  // n1.e1 = 1, n1.e2 = 0; n2.e1 = 0, n2.e2 = 1
      const uint32_t k = atomicAdd(&s_n[0], 1u);
      out->tri[3 * k] = make_float4(n.x, n.y, n.z, dot(n, p0));
      out->tri[3 * k + 1] = make_float4(n1.x, n1.y, n1.z, -dot(n1, p0));
      out->tri[3 * k + 2] = make_float4(n2.x, n2.y, n2.z, -dot(n2, p0));
      out->tri_leaf[k] = (uint8_t)i;
      out->tri_class[k] = cls;
    }
  }
  __syncthreads();
  if (i == 0) {
    out->n_tri = s_n[0];
    out->n_sph = s_n[1];
  }
}

// what a flat trace found: entry k of the triangle list (k < n_tri) or of the sphere list (k - n_tri)
struct QfHit {
  int k;       // -1: nothing
  float u, v;  // barycentric weights of p1 and p2 (triangles)
};

// closest hit over the whole list.  Every lane of the warp runs every entry; `on` = this lane has a ray.
VD void qf_closest(const QfScene& fs, bool on, V3 o, V3 d, float tmin, float& tmax, QfHit& h) {
  h.k = -1;
  h.u = h.v = 0.f;
  const uint32_t nt = fs.n_tri, ns = fs.n_sph;
  for (uint32_t k = 0; k < nt; ++k) {
    const float4 a = fs.tri[3 * k], b = fs.tri[3 * k + 1], c = fs.tri[3 * k + 2];
    const float det = a.x * d.x + a.y * d.y + a.z * d.z;
    const float t = (a.w - (a.x * o.x + a.y * o.y + a.z * o.z)) / det;  // det = 0: inf or NaN, rejected below
    if (on && t > tmin && t < tmax) {
      const float px = o.x + t * d.x, py = o.y + t * d.y, pz = o.z + t * d.z;
      const float u = b.x * px + b.y * py + b.z * pz + b.w;
      const float v = c.x * px + c.y * py + c.z * pz + c.w;
      if (u >= 0.f && v >= 0.f && u + v <= 1.f) {
        tmax = t;
        h.k = (int)k;
        h.u = u;
        h.v = v;
      }
    }
  }
  for (uint32_t k = 0; k < ns; ++k) {
    const float4 s = fs.sph[k];
    const float a4[4] = {s.x, s.y, s.z, s.w};
    if (on) {
      Ray r;
      r.o = o;
      r.dir = d;
      r.minT = tmin;
      r.maxT = tmax;
      if (sphere_hit(a4, r)) {
        tmax = r.maxT;
        h.k = (int)(nt + k);
      }
    }
  }
}

// any hit in (tmin, tmax): BVH::occlude
VD bool qf_occluded(const QfScene& fs, bool on, V3 o, V3 d, float tmin, float tmax) {
  bool occl = false;
  const uint32_t nt = fs.n_tri, ns = fs.n_sph;
  for (uint32_t k = 0; k < nt; ++k) {
    const float4 a = fs.tri[3 * k], b = fs.tri[3 * k + 1], c = fs.tri[3 * k + 2];
    const float det = a.x * d.x + a.y * d.y + a.z * d.z;
    const float t = (a.w - (a.x * o.x + a.y * o.y + a.z * o.z)) / det;
    if (on && !occl && t > tmin && t < tmax) {
      const float px = o.x + t * d.x, py = o.y + t * d.y, pz = o.z + t * d.z;
      const float u = b.x * px + b.y * py + b.z * pz + b.w;
      const float v = c.x * px + c.y * py + c.z * pz + c.w;
      occl = u >= 0.f && v >= 0.f && u + v <= 1.f;
    }
  }
  for (uint32_t k = 0; k < ns; ++k) {
    const float4 s = fs.sph[k];
    const float a4[4] = {s.x, s.y, s.z, s.w};
    if (on && !occl) {
      Ray r;
      r.o = o;
      r.dir = d;
      r.minT = tmin;
      r.maxT = tmax;
      occl = sphere_hit(a4, r);
    }
  }
  return occl;
}

// ---- branch-free tests over the lists: every lane runs every entry, no lane waits for another ----
struct QfBest {
  float t;     // closest accepted distance so far (starts at the ray's maxT)
  int k;       // entry: triangle k, or n_tri + sphere k; -1 none
  float u, v;  // barycentric weights of p1 and p2 (triangles)
};

VD void qf_test_tri(const float4 a, const float4 b, const float4 c, V3 o, V3 d, float tmin, int k, QfBest& best) {
  const float det = a.x * d.x + a.y * d.y + a.z * d.z;
  const float t = __fdividef(a.w - (a.x * o.x + a.y * o.y + a.z * o.z), det);  // det = 0: +-inf or NaN, no hit
  const float px = o.x + t * d.x, py = o.y + t * d.y, pz = o.z + t * d.z;
  const float u = b.x * px + b.y * py + b.z * pz + b.w;
  const float v = c.x * px + c.y * py + c.z * pz + c.w;
  const bool h = (t > tmin) & (t < best.t) & (u >= 0.f) & (v >= 0.f) & (u + v <= 1.f);
  best.t = h ? t : best.t;
  best.k = h ? k : best.k;
  best.u = h ? u : best.u;
  best.v = h ? v : best.v;
}

// Sphere::sphere_hit_template + solveQuadratic (sphere.h:13-100) in float, without the early returns
VD void qf_test_sphere(const float4 s, V3 o, V3 d, float tmin, int k, QfBest& best) {
  const float r2 = s.w * s.w;
  const V3 f = mk3(o.x - s.x, o.y - s.y, o.z - s.z);
  const float aa = dot(d, d);
  const float bp = -dot(f, d);
  const float cc = dot(f, f) - r2;
  const V3 tmp = f + __fdividef(bp, aa) * d;
  const float disc = r2 - dot(tmp, tmp);
  const float sq = sqrtf(aa * fmaxf(disc, 0.f));
  const float q = bp + (bp > 0.f ? sq : -sq);
  const float ta = __fdividef(cc, q), tb = __fdividef(q, aa);
  const float t0 = fminf(ta, tb), t1 = fmaxf(ta, tb);
  const float t = (t0 >= tmin && t0 <= best.t) ? t0 : t1;
  const bool h = (disc >= 0.f) & (t >= tmin) & (t <= best.t);
  best.t = h ? t : best.t;
  best.k = h ? k : best.k;
}

VD void qf_trace(const QfScene& fs, V3 o, V3 d, float tmin, QfBest& best) {
  const int nt = (int)fs.n_tri, ns = (int)fs.n_sph;
#pragma unroll 4
  for (int k = 0; k < nt; ++k) qf_test_tri(fs.tri[3 * k], fs.tri[3 * k + 1], fs.tri[3 * k + 2], o, d, tmin, k, best);
#pragma unroll 2
  for (int k = 0; k < ns; ++k) qf_test_sphere(fs.sph[k], o, d, tmin, nt + k, best);
}

// ====================================================== TRACE + CLASSIFY over the flat lists ====
// thread <-> queued ray (closest rays, then shadow rays: a shadow ray is occluded when anything lies in
// (minT, dist - 1e-4), so both kinds run the same loop).  The material class of what a closest ray hit
// routes its path straight into the shade / END queues: this launch also does qw_classify_kernel's work.
template <bool COUNT>
__global__ void __launch_bounds__(QW_THREADS, 4)
qf_trace_flat_kernel(const __grid_constant__ DScene sc, const __grid_constant__ QwState st, const uint32_t parity,
                     const QfScene* __restrict__ flat, DCounters* __restrict__ counters) {
  __shared__ QfScene fs;
  __shared__ uint32_t s_cnt[WF_NUM_CLASSES + 1], s_base[WF_NUM_CLASSES + 1];
  QwCounters* c = st.cnt;
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  if (blockIdx.x == 0 && tid == 0) {
    // queues of the NEXT iteration (their last consumers ran before this launch)
    c->n_closest[parity ^ 1u] = 0;
    c->n_shadow[parity ^ 1u] = 0;
    c->n_end[parity ^ 1u] = 0;
#pragma unroll
    for (uint32_t k = 0; k < WF_NUM_CLASSES; ++k) c->n_shade[parity ^ 1u][k] = 0;
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(flat);
    uint4* dst = reinterpret_cast<uint4*>(&fs);
    for (uint32_t i = tid; i < sizeof(QfScene) / 16; i += QW_THREADS) dst[i] = __ldg(src + i);
  }
  const uint32_t n_c = c->n_closest[parity], total = n_c + c->n_shadow[parity];
  const uint32_t* __restrict__ qc = st.q_closest[parity];
  const uint32_t* __restrict__ qs = st.q_shadow[parity];
  uint32_t rays_closest = 0, rays_shadow = 0;

  for (uint32_t base = blockIdx.x * QW_THREADS; base < total; base += gridDim.x * QW_THREADS) {
    if (tid < (int)WF_NUM_CLASSES + 1) s_cnt[tid] = 0;
    __syncthreads();  // also: the staged lists are visible
    const uint32_t item = base + tid;
    const bool valid = item < total;
    const bool shadow = item >= n_c;
    uint32_t slot = 0;
    V3 o = mk3(0.f), d = mk3(0.f, 0.f, 1.f);
    QfBest best;
    best.t = 0.f;  // no item: nothing is accepted
    best.k = -1;
    best.u = best.v = 0.f;
    if (valid) {
      slot = shadow ? qs[item - n_c] : qc[item];
      const float4 o4 = ldg4(&st.slot[slot].ray.o_w);
      // shadow ray (mis_integrator.cpp:59-64): origin = hit point, maxT = dist - 1e-4
      const float4 d4 = ldg4(shadow ? &st.slot[slot].nee.wi_dist : &st.slot[slot].ray.d_s);
      o = xyz(o4);
      d = xyz(d4);
      best.t = shadow ? d4.w - 0.0001f : CUDART_INF_F;
      rays_shadow += shadow;
      rays_closest += !shadow;
    }
    qf_trace(fs, o, d, 0.0001f, best);
    uint32_t cls = WF_CLASS_NONE + 1;  // NONE + 1: nothing to route (shadow ray, or no item)
    int32_t leaf = -1;
    if (valid) {
      QwHit* h = &st.slot[slot].hit;
      if (shadow) {
        h->occl = best.k >= 0 ? 1u : 0u;
      } else {
        cls = WF_CLASS_NONE;
        float e0 = 0.f, e1 = 0.f, e2 = 0.f, inv_det = 0.f;
        if (best.k >= 0) {
          const bool is_tri = (uint32_t)best.k < fs.n_tri;
          const uint32_t e = is_tri ? (uint32_t)best.k : (uint32_t)best.k - fs.n_tri;
          leaf = is_tri ? fs.tri_leaf[e] : fs.sph_leaf[e];
          const uint32_t lc = is_tri ? fs.tri_class[e] : fs.sph_class[e];
          cls = (lc & VIMG_LEAF_CLASS_EMISSIVE) ? WF_CLASS_NONE : (lc & 3u);
          if (is_tri) {  // weights of p0, p1, p2 as Triangle::hit_info reads them (e_i * inv_det)
            e0 = 1.f - best.u - best.v;
            e1 = best.u;
            e2 = best.v;
            inv_det = 1.f;
          }
        }
        stg4(&h->h0, make_float4(__int_as_float(leaf), best.k >= 0 ? best.t : CUDART_INF_F, e0, e1));
        h->h1 = make_float2(e2, inv_det);
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
    if (cls < WF_NUM_CLASSES) st.q_shade[cls][s_base[cls] + ap.off[cls]] = make_uint2(slot, (uint32_t)leaf);
    else if (cls == WF_CLASS_NONE) st.q_end[parity][s_base[WF_NUM_CLASSES] + ap.off[WF_NUM_CLASSES]] = slot;
  }
  warp_add_counter(&counters->rays_closest, rays_closest);
  warp_add_counter(&counters->rays_shadow, rays_shadow);
  if (COUNT) {
    warp_add_counter(&counters->tri_tests, (rays_closest + rays_shadow) * fs.n_tri);
    warp_add_counter(&counters->sphere_tests, (rays_closest + rays_shadow) * fs.n_sph);
  }
}

#ifndef QF_PATH_MIN_BLOCKS
#define QF_PATH_MIN_BLOCKS 3
#endif
// A lane whose path ended has a fresh camera ray.  It is traced in the same launch when at least this
// many lanes of the warp have one (the list loop costs the same for 1 lane as for 32), at most
// QF_MAX_PASSES times; otherwise the slot waits for the next iteration's launch.
#ifndef QF_REPASS_MIN
#define QF_REPASS_MIN 12
#endif
#ifndef QF_MAX_PASSES
#define QF_MAX_PASSES 4
#endif

// thread <-> slot of q_closest[parity]: everything between two BSDF evaluations of the slot's path
// (see the header comment).  The shade launches of this iteration consume what it queues; they
// queue the slots whose paths go on, or end after their last shadow ray, for the next iteration.
template <int INTEGRATOR>
__global__ void __launch_bounds__(QW_THREADS, QF_PATH_MIN_BLOCKS)
qf_path_kernel(const __grid_constant__ DScene sc, const __grid_constant__ QwState st,
               const __grid_constant__ WfArgs args, const uint32_t parity, const QfScene* __restrict__ flat,
               float* __restrict__ out, DCounters* __restrict__ counters) {
  constexpr bool MIS = INTEGRATOR == VIMG_INTEGRATOR_MIS;
  __shared__ QfScene fs;
  __shared__ uint32_t s_cnt[WF_NUM_CLASSES + 1], s_base[WF_NUM_CLASSES + 1];
  __shared__ bool s_last;
  QwCounters* c = st.cnt;
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  const uint32_t W = sc.cam.res[0], H = sc.cam.res[1];
  if (blockIdx.x == 0 && tid == 0) {
    // shade queues of the NEXT iteration (their consumers, the last iteration's shade launches, are done).
    // n_closest[parity ^ 1], which this launch and then this iteration's shade launches append to, was
    // zeroed by the last block of the previous path launch (below).
#pragma unroll
    for (uint32_t k = 0; k < WF_NUM_CLASSES; ++k) c->n_shade[parity ^ 1u][k] = 0;
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(flat);
    uint4* dst = reinterpret_cast<uint4*>(&fs);
    for (uint32_t i = tid; i < sizeof(QfScene) / 16; i += QW_THREADS) dst[i] = __ldg(src + i);
  }
  const uint32_t n = c->n_closest[parity];
  const uint32_t* __restrict__ q = st.q_closest[parity];
  uint32_t nan_samples = 0, n_samples = 0, rays_closest = 0, rays_shadow = 0;
  bool exhausted = *reinterpret_cast<volatile uint32_t*>(st.next_pixel) >= args.n_local_units;

  for (uint32_t base = blockIdx.x * QW_THREADS; base < n; base += gridDim.x * QW_THREADS) {
    if (tid < (int)WF_NUM_CLASSES + 1) s_cnt[tid] = 0;
    __syncthreads();  // also: the staged lists are visible
    const uint32_t item = base + tid;
    const bool valid = item < n;
    const uint32_t i = valid ? q[item] : 0u;
    uint4 pa1 = make_uint4(0, 0, 0, 0);
    float4 thr4 = make_float4(0.f, 0.f, 0.f, 0.f), o4 = thr4, d4 = thr4, wi4 = thr4, cp4 = thr4, L4 = thr4, ap4 = thr4;
    if (valid) {
      thr4 = ldg4(&st.slot[i].path.thr_eta);
      pa1 = ldg4u(&st.slot[i].path.rng);
      o4 = ldg4(&st.slot[i].ray.o_w);
      d4 = ldg4(&st.slot[i].ray.d_s);
      wi4 = ldg4(&st.slot[i].nee.wi_dist);
      cp4 = ldg4(&st.slot[i].nee.c_pdf);
      L4 = ldg4(&st.slot[i].pix.L_s);
      ap4 = ldg4(&st.slot[i].pix.acc_pixel);
    }
    uint32_t f = pa1.z;
    uint32_t d = pa1.w;
    Pcg32 rng;
    rng.state = ((unsigned long long)pa1.y << 32) | pa1.x;
    uint32_t range = f >> WF_RANGE_SHIFT;
    rng.inc = ((uint64_t)range << 1) | 1u;
    V3 acc = xyz(ap4), L = xyz(L4), thr = xyz(thr4);
    float eta_scale = thr4.w, last_pdf = cp4.w;
    uint32_t s = __float_as_uint(L4.w), pixel = __float_as_uint(ap4.w);
    V3 ro = xyz(o4), rd = xyz(d4);
    RayCone cone;
    cone.cone_width = o4.w;
    cone.spread_angle = d4.w;
    bool new_ray = false;        // the slot's ray was replaced by a camera ray
    uint32_t cls = WF_CLASS_NONE;  // class queue the slot goes to (NONE: none)
    float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 h1 = make_float2(0.f, 0.f);

    // finish the current sample of this slot's pixel (include/integrators.h:128-137)
    auto finish = [&](V3 result) {
      if (is_nan(result.x) || is_nan(result.y) || is_nan(result.z)) {
        nan_samples++;
        result = mk3(0.f);  // counted and dropped (the reference adds it: the pixel turns NaN)
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
      } else if (s == (uint32_t)(((uint64_t)(range + 1) * args.samples) / args.streams)) {
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

    bool work = valid;  // this lane's slot still has something to do in this launch
    for (int pass = 1;; ++pass) {
      // ---- A. shadow ray of the last bounce (mis_integrator.cpp:59-78): origin = hit point, maxT = dist - 1e-4
      if (MIS) {
        const bool sh = work && (f & WF_NEE);
        if (__any_sync(0xffffffffu, sh)) {
          const bool occl = qf_occluded(fs, sh, ro, xyz(wi4), 0.0001f, wi4.w - 0.0001f);
          if (sh) {
            rays_shadow++;
            if ((f & WF_NEE_VALID) && !occl) L = L + xyz(cp4);
            f &= ~(WF_NEE | WF_NEE_VALID);
          }
        }
      }
      // ---- B. a path that ended at its last bounce (Russian roulette, depth, no scattering)
      if (work && (f & WF_IN_PATH) && (f & WF_END)) finish(MIS ? L : xyz(cp4));
      // ---- C. closest hit of the ray in flight
      const bool tr = work && (f & WF_IN_PATH);
      float tmax = CUDART_INF_F;
      QfHit hit;
      qf_closest(fs, tr, ro, rd, 0.0001f, tmax, hit);
      if (tr) {
        rays_closest++;
        if (hit.k < 0) {
          // miss (mis_integrator.cpp:36-38,173-185; mat_integrator.cpp:24-27)
          const V3 bg = background_emit(sc, rd, cone);
          if (!MIS) {
            finish(thr * bg);
          } else if (f & WF_FIRST) {
            finish(bg);
          } else {
            if (last_pdf != 0 && sc.bg.emissive) {
              const float light_pdf = background_pdf(sc, rd) / sc.n_lights;
              L = L + thr * balance_heuristic(last_pdf, light_pdf) * bg;
            }
            finish(L);
          }
        } else {
          const bool is_tri = (uint32_t)hit.k < fs.n_tri;
          const uint32_t e = is_tri ? (uint32_t)hit.k : (uint32_t)hit.k - fs.n_tri;
          const int32_t leaf = is_tri ? fs.tri_leaf[e] : fs.sph_leaf[e];
          const uint32_t lc = is_tri ? fs.tri_class[e] : fs.sph_class[e];
          PreHit ph;
          ph.leaf_slot = leaf;
          ph.e0 = is_tri ? 1.f - hit.u - hit.v : 0.f;  // weights of p0, p1, p2 (Triangle::hit_info: e_i * inv_det)
          ph.e1 = hit.u;
          ph.e2 = hit.v;
          ph.inv_det = is_tri ? 1.f : 0.f;
          if (lc & VIMG_LEAF_CLASS_EMISSIVE) {
            Ray ray;
            ray.o = ro;
            ray.dir = rd;
            ray.minT = 0.0001f;
            ray.maxT = tmax;
            ray.cone = cone;
            const DLeafPrim nlp = sc.leaf_prims[leaf];
            const HitInfo next = wf_hit_info<false>(sc, nlp, load_shade_rec(sc, leaf), ph, ray);
            const vimg_material& nmat = sc.materials[next.mat];
            const V3 emit = mat_emitted(nmat, rd, next.hit_n_s);
            if (!MIS) {
              finish(thr * emit);
            } else if (f & WF_FIRST) {
              finish(emit);
            } else {
              if (last_pdf != 0) {
                const float light_pdf = emitter_surf_pdf_leaf(nlp, ro, next.hit_p, rd) / sc.n_lights;
                const float G = geometric_term(ro, next.hit_p, next.hit_n_g);
                L = L + thr * balance_heuristic(last_pdf * G, light_pdf) * emit;
              } else {
                L = L + thr * emit;
              }
              finish(L);
            }
          } else if ((lc & 3u) == WF_CLASS_NONE) {
            // base Material: sample_mat -> nullopt.  Under mis the Russian-roulette draw still happens
            // first (mis_integrator.cpp:154-167, 83-88)
            if (!MIS) {
              finish(mk3(0.f));
            } else {
              if (!(f & WF_FIRST) && d > 5) {
                V3 t = thr;
                russian_roulette(rng, eta_scale, t);
              }
              finish(L);
            }
          } else {
            // a surface to shade: the path goes on in this iteration's shade launch
            cls = lc & 3u;
            h0 = make_float4(__int_as_float(leaf), tmax, ph.e0, ph.e1);
            h1 = make_float2(ph.e2, ph.inv_det);
            work = false;
          }
        }
      }
      // ---- D. the slot's path is over (or it never had one): next sample, next pixel, camera ray
      {
        const bool want = work && !(f & WF_ALIVE);
        const uint32_t m = __ballot_sync(0xffffffffu, want);
        if (m) {
          uint32_t pbase = 0;
          if (lane == (uint32_t)(__ffs(m) - 1))
            pbase = exhausted ? args.n_local_units : atomicAdd(st.next_pixel, (uint32_t)__popc(m));
          pbase = __shfl_sync(0xffffffffu, pbase, __ffs(m) - 1);
          exhausted = exhausted || pbase >= args.n_local_units;
          if (want) {
            // tiles overhanging the frame contain pixel numbers without a pixel: take the next
            uint32_t u = pbase + __popc(m & lane_lt), p = 0;
            uint32_t px = 0, py = 0;
            bool got = false;
            while (u < args.n_local_units && !(got = wf_pixel_xy(args, sc.cam, p = u / args.my_streams, px, py)))
              u = atomicAdd(st.next_pixel, 1u);
            if (got) {
              range = args.s_rank + (u - p * args.my_streams) * args.s_n_ranks;
              pcg32_seed(rng, px + ((size_t)(H - 1 - py) * W), range);
              acc = mk3(0.f);
              s = (uint32_t)(((uint64_t)range * args.samples) / args.streams);
              pixel = p;
              f = WF_ALIVE | (range << WF_RANGE_SHIFT);
            } else {
              work = false;  // the frame is handed out: this slot retires
            }
          }
        }
      }
      if (work) {
        // camera ray of the next sample (include/integrators.h:119-123); rand2 gets the first draw
        uint32_t px, py;
        wf_pixel_xy(args, sc.cam, pixel, px, py);
        const V2 off = random_x_y_r2(px + py + s);
        const float rand2 = rand_float(rng);
        const float rand1 = rand_float(rng);
        const Ray cam = generate_ray(sc.cam, (float)px + off.x, (float)py + off.y, rand1, rand2);
        ro = cam.o;
        rd = cam.dir;
        cone = cam.cone;
        f = WF_ALIVE | WF_IN_PATH | WF_FIRST | (range << WF_RANGE_SHIFT);
        thr = mk3(1.f);
        eta_scale = 1.f;
        last_pdf = 0.f;
        L = mk3(0.f);
        d = 0;
        new_ray = true;
      }
      // lanes with a fresh camera ray: trace it now if enough of the warp has one, else next iteration
      if (pass >= QF_MAX_PASSES || __popc(__ballot_sync(0xffffffffu, work)) < QF_REPASS_MIN) break;
    }
    const bool requeue = work;  // camera ray generated, not traced yet

    if (valid) {
      if (new_ray) {
        stg4(&st.slot[i].ray.o_w, f4(ro, cone.cone_width));
        stg4(&st.slot[i].ray.d_s, f4(rd, cone.spread_angle));
        stg4(&st.slot[i].path.thr_eta, f4(thr, eta_scale));
        stg4(&st.slot[i].nee.c_pdf, make_float4(0.f, 0.f, 0.f, last_pdf));
      }
      __stcg(reinterpret_cast<uint4*>(&st.slot[i].path.rng),
             make_uint4((uint32_t)rng.state, (uint32_t)(rng.state >> 32), f, d));
      stg4(&st.slot[i].pix.L_s, f4(L, __uint_as_float(s)));
      stg4(&st.slot[i].pix.acc_pixel, f4(acc, __uint_as_float(pixel)));
      if (cls < WF_NUM_CLASSES) {
        stg4(&st.slot[i].hit.h0, h0);
        st.slot[i].hit.h1 = h1;
      }
    }
    bool want[WF_NUM_CLASSES + 1];
#pragma unroll
    for (uint32_t k = 0; k < WF_NUM_CLASSES; ++k) want[k] = valid && cls == k;
    want[WF_NUM_CLASSES] = requeue;
    QwAppend<WF_NUM_CLASSES + 1> ap;
    ap.count(s_cnt, want, lane, lane_lt);
    __syncthreads();
    if (tid < (int)WF_NUM_CLASSES) {
      if (s_cnt[tid]) s_base[tid] = atomicAdd(&c->n_shade[parity][tid], s_cnt[tid]);
    } else if (tid == (int)WF_NUM_CLASSES) {
      if (s_cnt[tid]) s_base[tid] = atomicAdd(&c->n_closest[parity ^ 1u], s_cnt[tid]);
    }
    __syncthreads();
    if (valid && cls < WF_NUM_CLASSES)
      st.q_shade[cls][s_base[cls] + ap.off[cls]] = make_uint2(i, (uint32_t)__float_as_int(h0.x));
    else if (requeue)
      st.q_closest[parity ^ 1u][s_base[WF_NUM_CLASSES] + ap.off[WF_NUM_CLASSES]] = i;
  }
  // the last block to get here zeroes this launch's own input counter: every block has read it, and the
  // shade launches of the NEXT iteration append to it
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(&c->blocks_done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && tid == 0) {
    c->n_closest[parity] = 0;
    c->blocks_done = 0;
  }
  warp_add_counter(&counters->rays_closest, rays_closest);
  warp_add_counter(&counters->rays_shadow, rays_shadow);
  warp_add_counter(&counters->nan_samples, nan_samples);
  warp_add_counter(&counters->samples, n_samples);
}

inline void qf_fill_kernel_table(QwKernelTable& t) {
  t.trace_flat[0] = (const void*)&qf_trace_flat_kernel<false>;
  t.trace_flat[1] = (const void*)&qf_trace_flat_kernel<true>;
  t.path_flat[0] = (const void*)&qf_path_kernel<VIMG_INTEGRATOR_MATERIAL>;
  t.path_flat[1] = (const void*)&qf_path_kernel<VIMG_INTEGRATOR_MIS>;
  t.build_flat = (const void*)&qf_build_kernel;
}

}  // namespace vimg
