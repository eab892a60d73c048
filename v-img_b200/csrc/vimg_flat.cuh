// vimg_flat.cuh -- tolerance build only (VIMG_FLAG_FAST): the device-side re-layout of a small BVH.
//
// SURVEY.md section 8(f)4.  For a scene of a few dozen primitives -- every scene the reference ships has 13 - 18 --
// the useful re-layout of the reference's binary BVH is NO tree.  qf_build_kernel turns the leaf array into flat
// lists that a block keeps in shared memory:
//   * parallelograms: two triangles that share an edge and whose fourth corner closes the parallelogram (each
//     "quad" of a scene file) become ONE entry;
//   * the remaining triangles;
//   both as three plane equations -- the supporting plane and the two barycentric planes -- so that t, u, v are
//   three dot products and one division;
//   * spheres as centre + radius.
// A ray tests every entry (qf_trace).  All lanes of a warp test the SAME entry at the same time, the record is one
// broadcast read from shared memory, there is no stack, no box test, no branch: the binary traversal of the same
// scene runs with 16 of 32 lanes per instruction (profiles/r4_trace_kernel_fresh_stack.txt), this loop with 31.
//
// What differs from the reference's traversal, and why this is the tolerance mode only: primitives the box tests
// would have culled are tested too, ties between primitives at the same distance resolve in list order, the
// triangle test is the plane form instead of the watertight one, directions are taken as unit vectors.  Each
// changes a hit only at float-rounding distance from an edge, a tie or the ray's own origin (tests/test_gpu_fast.py
// states what that does to an image).  The kernels that trace through these lists are in vimg_stream.cuh.
#pragma once
#include "vimg_queues.cuh"

namespace vimg {

#ifndef QF_FLAT_MAX
#define QF_FLAT_MAX 64  // leaf records a flat scene may have
#endif

struct QfScene {
  // entries 0 .. n_quad - 1 of `tri` are PARALLELOGRAMS: two triangles that share an edge and whose fourth corner
  // closes the parallelogram (every "quad" of a scene file is one) are one plane test with 0 <= u, v <= 1 instead
  // of two with u + v <= 1; entries n_quad .. n_tri - 1 are single triangles
  uint32_t n_tri, n_sph, n_quad, pad_;
  float4 tri[QF_FLAT_MAX * 3];  // {N0, d0} {N1, d1} {N2, d2}: t = (d0 - N0.o) / (N0.dir); u = N1.P + d1; v = N2.P + d2
  float4 sph[QF_FLAT_MAX];      // centre, radius
  uint8_t tri_leaf[QF_FLAT_MAX], sph_leaf[QF_FLAT_MAX];    // leaf record (index into leaf_prims) of each entry
  uint8_t tri_class[QF_FLAT_MAX], sph_class[QF_FLAT_MAX];  // its leaf_class byte
  // parallelograms: the second triangle (u + v > 1), and for both triangles which corner each of their vertices
  // p0, p1, p2 is -- 2 bits per vertex: 0 root (triangle A) / far corner (B), 1 = first shared vertex, 2 = second
  uint8_t quad_leaf_b[QF_FLAT_MAX], quad_class_b[QF_FLAT_MAX];
  uint16_t quad_roles[QF_FLAT_MAX];  // bits 0-5: triangle A's vertices, bits 6-11: triangle B's
};

// one block; thread <-> leaf record.  Degenerate triangles (triangle.h:89-92) are dropped.
VD bool qf_same(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
VD void qf_put_tri(QfScene* out, uint32_t k, V3 r, V3 s1, V3 s2) {
  const V3 e1 = s1 - r, e2 = s2 - r;
  const V3 n = cross(e1, e2);
  const float inv_n2 = 1.0f / dot(n, n);
  const V3 n1 = cross(e2, n) * inv_n2, n2 = cross(n, e1) * inv_n2;  // n1.e1 = 1, n1.e2 = 0; n2.e1 = 0, n2.e2 = 1
  out->tri[3 * k] = make_float4(n.x, n.y, n.z, dot(n, r));
  out->tri[3 * k + 1] = make_float4(n1.x, n1.y, n1.z, -dot(n1, r));
  out->tri[3 * k + 2] = make_float4(n2.x, n2.y, n2.z, -dot(n2, r));
}
__global__ void qf_build_kernel(const __grid_constant__ DScene sc, QfScene* __restrict__ out) {
  __shared__ V3 s_v[QF_FLAT_MAX][3];
  __shared__ uint8_t s_kind[QF_FLAT_MAX];          // 0 nothing, 1 triangle, 2 sphere
  __shared__ int s_partner[QF_FLAT_MAX];           // triangle: the triangle it forms a parallelogram with, or -1
  __shared__ uint16_t s_roles[QF_FLAT_MAX];        // for the pair (i, partner[i]) with i < partner[i]
  __shared__ unsigned long long s_cand[QF_FLAT_MAX];
  const uint32_t i = threadIdx.x, n = sc.n_leaf_recs;
  DLeafPrim lp;
  if (i < n) {
    lp = sc.leaf_prims[i];
    s_kind[i] = (lp.flags & VIMG_LEAF_SPHERE) ? 2 : ((lp.flags & VIMG_LEAF_DEGENERATE) ? 0 : 1);
    for (int k = 0; k < 3; ++k) s_v[i][k] = mk3(lp.a[3 * k], lp.a[3 * k + 1], lp.a[3 * k + 2]);
  }
  __syncthreads();
  // candidates: j shares exactly two vertices with i and its third vertex closes the parallelogram
  if (i < n) {
    unsigned long long cand = 0;
    if (s_kind[i] == 1)
      for (uint32_t j = 0; j < n; ++j) {
        if (j == i || s_kind[j] != 1) continue;
        int shared_i[3] = {-1, -1, -1}, n_shared = 0;  // shared_i[a] = vertex of j equal to vertex a of i
        for (int a = 0; a < 3; ++a)
          for (int b2 = 0; b2 < 3; ++b2)
            if (shared_i[a] < 0 && qf_same(s_v[i][a], s_v[j][b2])) {
              shared_i[a] = b2;
              ++n_shared;
            }
        if (n_shared != 2) continue;
        int root = shared_i[0] < 0 ? 0 : (shared_i[1] < 0 ? 1 : 2);
        int far_j = 3 - shared_i[(root + 1) % 3] - shared_i[(root + 2) % 3];
        const V3 r = s_v[i][root], a1 = s_v[i][(root + 1) % 3], a2 = s_v[i][(root + 2) % 3], q = s_v[j][far_j];
        const V3 want = a1 + a2 - r, d = q - want;
        const float scale = dot(a1 - r, a1 - r) + dot(a2 - r, a2 - r);
        if (dot(d, d) <= 1e-10f * scale) cand |= 1ull << j;
      }
    s_cand[i] = cand;
    s_partner[i] = -1;
  }
  __syncthreads();
  if (i == 0) {  // greedy pairing in leaf order (n <= 64)
    for (uint32_t a = 0; a < n; ++a) {
      if (s_kind[a] != 1 || s_partner[a] >= 0) continue;
      for (uint32_t b2 = a + 1; b2 < n; ++b2)
        if (((s_cand[a] >> b2) & 1ull) && ((s_cand[b2] >> a) & 1ull) && s_partner[b2] < 0) {
          s_partner[a] = (int)b2;
          s_partner[b2] = (int)a;
          break;
        }
    }
    // list positions: parallelograms first (by their first triangle), then single triangles, spheres apart
    uint32_t nq = 0, nt = 0, ns = 0;
    for (uint32_t a = 0; a < n; ++a)
      if (s_kind[a] == 1 && s_partner[a] > (int)a) ++nq;
    uint32_t kq = 0, kt = nq;
    for (uint32_t a = 0; a < n; ++a) {
      if (s_kind[a] == 2) {
        out->sph[ns] = make_float4(s_v[a][0].x, s_v[a][0].y, s_v[a][0].z, s_v[a][1].x);
        out->sph_leaf[ns] = (uint8_t)a;
        out->sph_class[ns] = sc.leaf_class[a];
        ++ns;
      } else if (s_kind[a] == 1 && s_partner[a] < 0) {
        qf_put_tri(out, kt, s_v[a][0], s_v[a][1], s_v[a][2]);
        out->tri_leaf[kt] = (uint8_t)a;
        out->tri_class[kt] = sc.leaf_class[a];
        out->quad_leaf_b[kt] = (uint8_t)a;
        out->quad_class_b[kt] = sc.leaf_class[a];
        out->quad_roles[kt] = (uint16_t)(0u | (1u << 2) | (2u << 4));  // p0 root, p1 first, p2 second
        ++kt;
        ++nt;
      } else if (s_kind[a] == 1 && s_partner[a] > (int)a) {
        const uint32_t b2 = (uint32_t)s_partner[a];
        int sh[3] = {-1, -1, -1};
        for (int x = 0; x < 3; ++x)
          for (int y = 0; y < 3; ++y)
            if (sh[x] < 0 && qf_same(s_v[a][x], s_v[b2][y])) sh[x] = y;
        const int root = sh[0] < 0 ? 0 : (sh[1] < 0 ? 1 : 2), i1 = (root + 1) % 3, i2 = (root + 2) % 3;
        qf_put_tri(out, kq, s_v[a][root], s_v[a][i1], s_v[a][i2]);
        uint32_t roles = 0;
        roles |= 0u << (2 * root);
        roles |= 1u << (2 * i1);
        roles |= 2u << (2 * i2);
        const int far_b = 3 - sh[i1] - sh[i2];
        roles |= 0u << (6 + 2 * far_b);
        roles |= 1u << (6 + 2 * sh[i1]);
        roles |= 2u << (6 + 2 * sh[i2]);
        out->tri_leaf[kq] = (uint8_t)a;
        out->tri_class[kq] = sc.leaf_class[a];
        out->quad_leaf_b[kq] = (uint8_t)b2;
        out->quad_class_b[kq] = sc.leaf_class[b2];
        out->quad_roles[kq] = (uint16_t)roles;
        ++kq;
      }
    }
    out->n_quad = nq;
    out->n_tri = nq + nt;
    out->n_sph = ns;
  }
}

// ---- branch-free tests over the lists: every lane runs every entry, no lane waits for another ----
struct QfBest {
  float t;     // closest accepted distance so far (starts at the ray's maxT)
  int k;       // entry: triangle k, or n_tri + sphere k; -1 none
  float u, v;  // barycentric weights of p1 and p2 (triangles)
};

template <bool QUAD>
VD void qf_test_tri(const float4 a, const float4 b, const float4 c, V3 o, V3 d, float tmin, int k, QfBest& best) {
  const float det = a.x * d.x + a.y * d.y + a.z * d.z;
  const float t = __fdividef(a.w - (a.x * o.x + a.y * o.y + a.z * o.z), det);  // det = 0: +-inf or NaN, no hit
  const float px = o.x + t * d.x, py = o.y + t * d.y, pz = o.z + t * d.z;
  const float u = b.x * px + b.y * py + b.z * pz + b.w;
  const float v = c.x * px + c.y * py + c.z * pz + c.w;
  const bool inside = QUAD ? ((u <= 1.f) & (v <= 1.f)) : (u + v <= 1.f);
  const bool h = (t > tmin) & (t < best.t) & (u >= 0.f) & (v >= 0.f) & inside;
  best.t = h ? t : best.t;
  best.k = h ? k : best.k;
  best.u = h ? u : best.u;
  best.v = h ? v : best.v;
}

// Sphere::sphere_hit_template + solveQuadratic (sphere.h:13-100) in float, without the early returns, for a UNIT
// direction (every direction the integrators trace is normalised: camera rays, BSDF samples, light samples)
VD void qf_test_sphere(const float4 s, V3 o, V3 d, float tmin, int k, QfBest& best) {
  const float r2 = s.w * s.w;
  const V3 f = mk3(o.x - s.x, o.y - s.y, o.z - s.z);
  const float bp = -dot(f, d);
  const float cc = dot(f, f) - r2;
  const V3 tmp = f + bp * d;
  const float disc = r2 - dot(tmp, tmp);
  const float sq = sqrtf(fmaxf(disc, 0.f));
  const float q = bp + (bp > 0.f ? sq : -sq);
  const float ta = __fdividef(cc, q);
  const float t0 = fminf(ta, q), t1 = fmaxf(ta, q);
  const float t = (t0 >= tmin && t0 <= best.t) ? t0 : t1;
  const bool h = (disc >= 0.f) & (t >= tmin) & (t <= best.t);
  best.t = h ? t : best.t;
  best.k = h ? k : best.k;
}

VD void qf_trace(const QfScene& fs, V3 o, V3 d, float tmin, QfBest& best) {
  const int nq = (int)fs.n_quad, nt = (int)fs.n_tri, ns = (int)fs.n_sph;
#pragma unroll 2
  for (int k = 0; k < nq; ++k) qf_test_tri<true>(fs.tri[3 * k], fs.tri[3 * k + 1], fs.tri[3 * k + 2], o, d, tmin, k, best);
#pragma unroll 2
  for (int k = nq; k < nt; ++k) qf_test_tri<false>(fs.tri[3 * k], fs.tri[3 * k + 1], fs.tri[3 * k + 2], o, d, tmin, k, best);
#pragma unroll 2
  for (int k = 0; k < ns; ++k) qf_test_sphere(fs.sph[k], o, d, tmin, nt + k, best);
}

// the closest ray and the shadow ray of one path segment (same origin) in ONE pass over the lists: an entry's record is
// read once and the two tests are independent instruction streams (the fused tail of the shade kernels, vimg_stream.cuh)
VD void qf_trace2(const QfScene& fs, V3 o, V3 d1, V3 d2, float tmin, QfBest& best1, QfBest& best2) {
  const int nq = (int)fs.n_quad, nt = (int)fs.n_tri, ns = (int)fs.n_sph;
#pragma unroll 2
  for (int k = 0; k < nq; ++k) {
    const float4 a = fs.tri[3 * k], b = fs.tri[3 * k + 1], c = fs.tri[3 * k + 2];
    qf_test_tri<true>(a, b, c, o, d1, tmin, k, best1);
    qf_test_tri<true>(a, b, c, o, d2, tmin, k, best2);
  }
#pragma unroll 2
  for (int k = nq; k < nt; ++k) {
    const float4 a = fs.tri[3 * k], b = fs.tri[3 * k + 1], c = fs.tri[3 * k + 2];
    qf_test_tri<false>(a, b, c, o, d1, tmin, k, best1);
    qf_test_tri<false>(a, b, c, o, d2, tmin, k, best2);
  }
#pragma unroll 2
  for (int k = 0; k < ns; ++k) {
    const float4 sp = fs.sph[k];
    qf_test_sphere(sp, o, d1, tmin, nt + k, best1);
    qf_test_sphere(sp, o, d2, tmin, nt + k, best2);
  }
}

// what entry best.k is: leaf record, class byte and the weights of that leaf's p0, p1, p2 (Triangle::hit_info reads
// them as e_i * inv_det; spheres: zeros)
VD void qf_decode(const QfScene& fs, const QfBest& best, int32_t& leaf, uint32_t& lc, float4& e) {
  const bool is_tri = (uint32_t)best.k < fs.n_tri;
  e = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!is_tri) {
    const uint32_t k = (uint32_t)best.k - fs.n_tri;
    leaf = fs.sph_leaf[k];
    lc = fs.sph_class[k];
    return;
  }
  const uint32_t k = (uint32_t)best.k;
  const bool second = best.u + best.v > 1.f;  // only a parallelogram reaches this half
  leaf = second ? fs.quad_leaf_b[k] : fs.tri_leaf[k];
  lc = second ? fs.quad_class_b[k] : fs.tri_class[k];
  const uint32_t roles = (uint32_t)fs.quad_roles[k] >> (second ? 6 : 0);
  const float w[3] = {second ? best.u + best.v - 1.f : 1.f - best.u - best.v, second ? 1.f - best.v : best.u,
                      second ? 1.f - best.u : best.v};
  const uint32_t r0 = roles & 3u, r1 = (roles >> 2) & 3u, r2 = (roles >> 4) & 3u;
  e.x = r0 == 0 ? w[0] : (r0 == 1 ? w[1] : w[2]);
  e.y = r1 == 0 ? w[0] : (r1 == 1 ? w[1] : w[2]);
  e.z = r2 == 0 ? w[0] : (r2 == 1 ? w[1] : w[2]);
  e.w = 1.f;
}

inline void qf_fill_kernel_table(QwKernelTable& t) { t.build_flat = (const void*)&qf_build_kernel; }

}  // namespace vimg
