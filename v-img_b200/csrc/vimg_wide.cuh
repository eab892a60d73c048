// vimg_wide.cuh -- tolerance build only (VIMG_FLAG_FAST): a 4-wide BVH built on the device from the reference's
// binary BVH, and the traversal kernel for it (SURVEY.md section 8(f)4 for LARGE scenes; vimg_flat.cuh is the
// re-layout of small ones).
//
// Why: on the 2 M / 10 M-triangle scenes the binary traversal is a chain of ~33 dependent 64-byte record fetches
// per ray from L2 / HBM (profiles/r5_trace_c4_synthetic2m.txt: long_scoreboard 8.2 of 14 cycles per instruction).
// Collapsing every second level halves the chain: a wide node holds the (up to) four grandchildren of a binary
// inner node, one 128-byte, 128-byte-aligned record = one line fetch, boxes as SoA so that a lane tests all four
// with 16-byte loads.  The reference's visiting order is not reproduced -- children are visited nearest first
// and a popped entry whose entry distance lies behind the closest hit so far is dropped without a fetch -- which
// is why this is the tolerance mode's traversal only: the same primitives win except at float-rounding distance
// from a tie or a box face.
//
// Build: qw_wide_build_kernel, thread <-> binary inner node (= its pair record): the node's two children, each
// replaced by ITS two children when it is inner.  Every binary inner node gets a wide record at its own pair index
// -- no frontier, no atomics, one pass over the array; traversal from the root only ever reaches every second
// level, the records of the levels in between are never fetched.
#pragma once
#include "vimg_queues.cuh"

namespace vimg {

#define WIDE_EMPTY 0xffffffffu
#define WIDE_LEAF 0x80000000u  // leaf reference: WIDE_LEAF | (count - 1) << 27 | first leaf record (< 2^27)
#define WIDE_MAX_LEAF_COUNT 16u
#define WIDE_MAX_LEAF_FIRST (1u << 27)

#ifndef WIDE_NODE_BYTES
#define WIDE_NODE_BYTES 128  // 256: the address-range experiment of DESIGN.md 5.1d
#endif
struct __align__(WIDE_NODE_BYTES) DWideNode {
  float4 minx, miny, minz, maxx, maxy, maxz;  // the four children's boxes, one coordinate per vector
  uint4 ref;                                  // WIDE_EMPTY | leaf reference | wide index of an inner child
  uint4 pad_;
};
static_assert(sizeof(DWideNode) == WIDE_NODE_BYTES, "one line per wide node");

#ifdef VIMG_FAST
VD uint32_t wide_ref(uint32_t first, uint32_t count) {
  return count == 0 ? (first - 1u) >> 1 : (WIDE_LEAF | ((count - 1u) << 27) | first);
}

__global__ void qw_wide_build_kernel(const DPairNode* __restrict__ pairs, uint32_t n_pairs, DWideNode* __restrict__ out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_pairs) return;
  const DPairNode p = pairs[r];
  float mn[4][3], mx[4][3];
  uint32_t ref[4];
  int n = 0;
  auto add = [&](const float* bmin, const float* bmax, uint32_t first, uint32_t count) {
    for (int k = 0; k < 3; ++k) {
      mn[n][k] = bmin[k];
      mx[n][k] = bmax[k];
    }
    ref[n++] = wide_ref(first, count);
  };
  auto expand = [&](const float* bmin, const float* bmax, uint32_t first, uint32_t count) {
    if (count != 0) {
      add(bmin, bmax, first, count);
    } else {
      const DPairNode q = pairs[(first - 1u) >> 1];
      add(q.lmin, q.lmax, q.l_first, q.l_count);
      add(q.rmin, q.rmax, q.r_first, q.r_count);
    }
  };
  expand(p.lmin, p.lmax, p.l_first, p.l_count);
  expand(p.rmin, p.rmax, p.r_first, p.r_count);
  for (; n < 4; ++n) {  // an empty child: the traversal looks at the reference, not at this box
    for (int k = 0; k < 3; ++k) mn[n][k] = mx[n][k] = 0.f;
    ref[n] = WIDE_EMPTY;
  }
  DWideNode w;
  w.minx = make_float4(mn[0][0], mn[1][0], mn[2][0], mn[3][0]);
  w.miny = make_float4(mn[0][1], mn[1][1], mn[2][1], mn[3][1]);
  w.minz = make_float4(mn[0][2], mn[1][2], mn[2][2], mn[3][2]);
  w.maxx = make_float4(mx[0][0], mx[1][0], mx[2][0], mx[3][0]);
  w.maxy = make_float4(mx[0][1], mx[1][1], mx[2][1], mx[3][1]);
  w.maxz = make_float4(mx[0][2], mx[1][2], mx[2][2], mx[3][2]);
  w.ref = make_uint4(ref[0], ref[1], ref[2], ref[3]);
  w.pad_ = make_uint4(0, 0, 0, 0);
  out[r] = w;
}

// entry distance of the ray into one child box, +inf if it misses ((b - o) * inv as b * inv - o * inv: one FFMA)
VD float wide_slab(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, V3 inv, V3 oinv,
                   float tmin, float tmax) {
  const float x0 = bminx * inv.x - oinv.x, x1 = bmaxx * inv.x - oinv.x;
  const float y0 = bminy * inv.y - oinv.y, y1 = bmaxy * inv.y - oinv.y;
  const float z0 = bminz * inv.z - oinv.z, z1 = bmaxz * inv.z - oinv.z;
  const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
  const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
  return tn <= tf ? tn : CUDART_INF_F;
}
VD void wide_order(float& ta, uint32_t& ra, float& tb, uint32_t& rb) {  // afterwards ta <= tb
  if (tb < ta) {
    const float t = ta;
    ta = tb;
    tb = t;
    const uint32_t r = ra;
    ra = rb;
    rb = r;
  }
}

// BVH::hit<closest> / BVH::occlude for every queued ray over the wide BVH; same queues in and out as
// qw_trace_kernel (persistent, idle lanes refilled from the warp's chunk), fresh stack per ray.
template <bool COUNT>
__global__ void __launch_bounds__(QW_TRACE_THREADS, QW_TRACE_BLOCKS_PER_SM)
qw_trace_wide_kernel(const __grid_constant__ DScene sc, const __grid_constant__ QwState st, const uint32_t parity,
                     const DWideNode* __restrict__ wide, DCounters* __restrict__ counters) {
  QwCounters* c = st.cnt;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
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
  const uint32_t n_warps = gridDim.x * (QW_TRACE_THREADS / 32);
  uint32_t chunk = (total / n_warps) & ~31u;
  chunk = chunk < 32u ? 32u : (chunk > (uint32_t)QW_CHUNK ? (uint32_t)QW_CHUNK : chunk);
  PathCounters pc = {};
  uint32_t chunk_next = 0, chunk_end = 0;
  bool more = total > 0;
  const uint32_t root = wide_ref(sc.root_first, sc.root_count);  // binary node 0: a leaf, or the wide record at its pair index

  bool have = false, done = false, any_hit = false, occluded = false;
  uint32_t slot = 0, qpos = 0, cur = 0;
  Ray ray;
  RayPre pre;
  PreHit ph;
  V3 inv = mk3(0.f), oinv = mk3(0.f);
  int sp = 0;
  uint2 stack[VIMG_STACK_MAX];  // {reference, entry distance}
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
    if (done) {
      QwHit* h = &st.slot[slot].hit;
      if (any_hit) {
        h->occl = occluded ? 1u : 0u;
      } else {
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
            ph.e0 = ph.e1 = ph.e2 = ph.inv_det = 0.f;
            ph.leaf_slot = -1;
            occluded = false;
            sp = 0;
            float4 o4, d4;
            if (idx < n_c) {
              slot = qc[idx] & ~QW_PAIR_BIT;
              qpos = idx;
              any_hit = false;
              o4 = ldg4(&st.slot[slot].ray.o_w);
              d4 = ldg4(&st.slot[slot].ray.d_s);
              ray.maxT = CUDART_INF_F;
              pc.rays_closest++;
            } else {
              any_hit = true;
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
              // a direction component of 0: a huge finite reciprocal keeps 0 * inf out of the slab arithmetic
              const float big = 1e30f;
              inv = mk3(fabsf(ray.dir.x) > 1e-30f ? pre.inv_dir.x : big, fabsf(ray.dir.y) > 1e-30f ? pre.inv_dir.y : big,
                        fabsf(ray.dir.z) > 1e-30f ? pre.inv_dir.z : big);
              oinv = ray.o * inv;
              const float root_hit = wide_slab(sc.root_min[0], sc.root_min[1], sc.root_min[2], sc.root_max[0], sc.root_max[1],
                                               sc.root_max[2], inv, oinv, ray.minT, ray.maxT);
              if (!is_inf(root_hit)) {
                cur = root;
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

    // wide inner nodes: descend until this lane holds a leaf or its ray is done
    while (have && !(cur & WIDE_LEAF)) {
      if (COUNT) pc.trav.inner++;
      const float4* rec = reinterpret_cast<const float4*>(wide + cur);
      const float4 ax = __ldg(rec), ay = __ldg(rec + 1), az = __ldg(rec + 2);
      const float4 bx = __ldg(rec + 3), by = __ldg(rec + 4), bz = __ldg(rec + 5);
      const uint4 rf = __ldg(reinterpret_cast<const uint4*>(rec + 6));
      float t0 = wide_slab(ax.x, ay.x, az.x, bx.x, by.x, bz.x, inv, oinv, ray.minT, ray.maxT);
      float t1 = wide_slab(ax.y, ay.y, az.y, bx.y, by.y, bz.y, inv, oinv, ray.minT, ray.maxT);
      float t2 = wide_slab(ax.z, ay.z, az.z, bx.z, by.z, bz.z, inv, oinv, ray.minT, ray.maxT);
      float t3 = wide_slab(ax.w, ay.w, az.w, bx.w, by.w, bz.w, inv, oinv, ray.minT, ray.maxT);
      uint32_t r0 = rf.x, r1 = rf.y, r2 = rf.z, r3 = rf.w;
      t2 = r2 == WIDE_EMPTY ? CUDART_INF_F : t2;  // a binary node has at least two children: slots 2 and 3 may be empty
      t3 = r3 == WIDE_EMPTY ? CUDART_INF_F : t3;
      // nearest first (misses carry +inf and sink to the end); shadow rays take them as they come
      if (!any_hit) {
        wide_order(t0, r0, t1, r1);
        wide_order(t2, r2, t3, r3);
        wide_order(t0, r0, t2, r2);
        wide_order(t1, r1, t3, r3);
        wide_order(t1, r1, t2, r2);
      }
      // push the hit children other than the first one, farthest first
      if (any_hit) {
        // unordered: the first hit child is visited, the others pushed
        uint32_t first = WIDE_EMPTY;
        float tf = 0.f;
#define WIDE_TAKE(t, r)                                             \
  if (!is_inf(t)) {                                                 \
    if (first == WIDE_EMPTY) {                                      \
      first = r;                                                    \
      tf = t;                                                       \
    } else {                                                        \
      stack[sp++] = make_uint2(r, __float_as_uint(t));              \
    }                                                               \
  }
        WIDE_TAKE(t0, r0)
        WIDE_TAKE(t1, r1)
        WIDE_TAKE(t2, r2)
        WIDE_TAKE(t3, r3)
#undef WIDE_TAKE
        (void)tf;
        if (first != WIDE_EMPTY) {
          cur = first;
        } else if (sp > 0) {
          cur = stack[--sp].x;
        } else {
          have = false;
          done = true;
        }
      } else {
        if (!is_inf(t3)) stack[sp++] = make_uint2(r3, __float_as_uint(t3));
        if (!is_inf(t2)) stack[sp++] = make_uint2(r2, __float_as_uint(t2));
        if (!is_inf(t1)) stack[sp++] = make_uint2(r1, __float_as_uint(t1));
        if (!is_inf(t0)) {
          cur = r0;
        } else {
          // nothing entered: the next pending entry that can still hold a closer hit
          bool got = false;
          while (sp > 0) {
            const uint2 e = stack[--sp];
            if (__uint_as_float(e.y) <= ray.maxT) {
              cur = e.x;
              got = true;
              break;
            }
          }
          if (!got) {
            have = false;
            done = true;
          }
        }
      }
      if (QW_INNER_MIN_LANES > 0 && __popc(__activemask()) < QW_INNER_MIN_LANES) break;
    }
    if (have && (cur & WIDE_LEAF)) {
      if (COUNT) pc.trav.leaf++;
      const uint32_t first = cur & (WIDE_MAX_LEAF_FIRST - 1u), count = ((cur >> 27) & 15u) + 1u;
      for (uint32_t i = 0; i < count; ++i) {
        const uint32_t ls = first + i;
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
      bool got = false;
      if (!occluded) {
        while (sp > 0) {
          const uint2 e = stack[--sp];
          if (any_hit || __uint_as_float(e.y) <= ray.maxT) {
            cur = e.x;
            got = true;
            break;
          }
        }
      }
      if (!got) {
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

inline void qw_wide_fill_kernel_table(QwKernelTable& t) {
  t.wide_build = (const void*)&qw_wide_build_kernel;
  t.trace_wide[0] = (const void*)&qw_trace_wide_kernel<false>;
  t.trace_wide[1] = (const void*)&qw_trace_wide_kernel<true>;
}
#endif  // VIMG_FAST

}  // namespace vimg
