// vimg_bvh.cuh -- camera rays, BVH traversal, ray/triangle and ray/sphere tests.
//
// Everything here feeds the BIT-EXACT parity targets (primary hit ids, normal
// images), so each function follows the reference's scalar path operation by
// operation: src/tl_camera.cpp:25-53, include/bvh.h:83-225 (#else branch),
// include/hit_utils.h:134-151, include/geometry/triangle.h:78-180,
// include/geometry/sphere.h:13-100.
#pragma once
#include "vimg_rng.cuh"

namespace vimg {

// ---------------------------------------------------------------- camera ----
// TLCam::generate_ray (src/tl_camera.cpp:25-53) + Ray::xform_ray (ray.h:36-41)
VD Ray generate_ray(const DCamera& cam, float x, float y, float rand1, float rand2) {
  float x_dir = (cam.p_size[0] * (x / cam.res[0])) - (cam.p_size[0] / 2.0f);
  float y_dir = (cam.p_size[1] * (y / cam.res[1])) - (cam.p_size[1] / 2.0f);
  V3 ray_dir = normalize(mk3(x_dir, y_dir, -1.0f));
  Ray r;
  r.o = mk3(0.f);
  r.dir = ray_dir;
  r.minT = 0.0001f;
  r.maxT = CUDART_INF_F;
  if (cam.aperture_radius > 0.f) {
    V2 d = cam.aperture_radius * sample_disk(rand1, rand2);
    V3 ray_origin = mk3(d.x, d.y, 0.f);
    float ft = cam.focal_dist / fabsf(ray_dir.z);
    V3 focal_plane_p = ray_dir * ft;
    r.o = ray_origin;
    r.dir = normalize(focal_plane_p - ray_origin);
  }
  V4 d4 = mat4_mul(cam.m, mk4(r.dir.x, r.dir.y, r.dir.z, 0.0f));
  V4 o4 = mat4_mul(cam.m, mk4(r.o.x, r.o.y, r.o.z, 1.0f));
  r.dir = mk3(d4.x, d4.y, d4.z);
  r.o = mk3(o4.x / o4.w, o4.y / o4.w, o4.z / o4.w);
  r.dir = normalize(r.dir);
  r.cone.cone_width = 0.f;
  r.cone.spread_angle = cam.cone_spread;
  return r;
}

// ------------------------------------------------------------- slab test ----
// Per-ray constants of the traversal: exact reciprocal direction
// (bvh.h:119-122) and the watertight test's permutation/shear, which the
// reference recomputes per triangle from the ray alone (triangle.h:105-125).
struct RayPre {
  V3 inv_dir;
  bool has_inf;  // some inv_dir component is +-inf => 0*inf NaNs possible in the slab test
  int kx, ky, kz;
  float Sx, Sy, Sz;
};

VD int max_component_index(V3 d) {  // triangle.h:23-40 (strict >)
  float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  int index = 0;
  float max_val = ax;
  if (ay > max_val) { index = 1; max_val = ay; }
  if (az > max_val) { index = 2; max_val = az; }
  return index;
}

VD float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

VD RayPre ray_precompute(const Ray& ray) {
  RayPre p;
  p.inv_dir = mk3(1.0f / ray.dir.x, 1.0f / ray.dir.y, 1.0f / ray.dir.z);
  // not finite (inf or NaN): take the exact-ternary slab path
  p.has_inf = !(fabsf(p.inv_dir.x) < CUDART_INF_F) || !(fabsf(p.inv_dir.y) < CUDART_INF_F)
              || !(fabsf(p.inv_dir.z) < CUDART_INF_F);
  p.kz = max_component_index(ray.dir);
  p.kx = p.kz + 1;
  if (p.kx == 3) p.kx = 0;
  p.ky = p.kx + 1;
  if (p.ky == 3) p.ky = 0;
  float dx = comp(ray.dir, p.kx), dy = comp(ray.dir, p.ky), dz = comp(ray.dir, p.kz);
  p.Sx = -dx / dz;
  p.Sy = -dy / dz;
  p.Sz = 1.f / dz;
  return p;
}

// slab_intersect_aabb_array (hit_utils.h:134-151).  EXACT=true writes out the
// glm::min/max and std::max/min ternaries so NaNs (0*inf) propagate exactly as on
// the CPU; EXACT=false uses fminf/fmaxf, identical whenever no
// NaN can occur, i.e. whenever inv_dir is finite.
template <bool EXACT>
VD float slab_test(V3 bmin, V3 bmax, V3 o, V3 inv, float minT, float maxT) {
  float lx = (bmin.x - o.x) * inv.x, ly = (bmin.y - o.y) * inv.y, lz = (bmin.z - o.z) * inv.z;
  float ux = (bmax.x - o.x) * inv.x, uy = (bmax.y - o.y) * inv.y, uz = (bmax.z - o.z) * inv.z;
  float tBoxMin, tBoxMax;
  if (EXACT) {
    float mnx = glm_min(lx, ux), mny = glm_min(ly, uy), mnz = glm_min(lz, uz);
    float mxx = glm_max(lx, ux), mxy = glm_max(ly, uy), mxz = glm_max(lz, uz);
    tBoxMin = std_max(mnx, std_max(mny, std_max(mnz, minT)));
    tBoxMax = std_min(mxx, std_min(mxy, std_min(mxz, maxT)));
  } else {
    tBoxMin = fmaxf(fminf(lx, ux), fmaxf(fminf(ly, uy), fmaxf(fminf(lz, uz), minT)));
    tBoxMax = fminf(fmaxf(lx, ux), fminf(fmaxf(ly, uy), fminf(fmaxf(lz, uz), maxT)));
  }
  return (tBoxMin <= tBoxMax) ? tBoxMin : CUDART_INF_F;
}

VD float slab(const float* bmin, const float* bmax, const Ray& ray, const RayPre& pre) {
  V3 mn = mk3(bmin[0], bmin[1], bmin[2]), mx = mk3(bmax[0], bmax[1], bmax[2]);
  return pre.has_inf ? slab_test<true>(mn, mx, ray.o, pre.inv_dir, ray.minT, ray.maxT)
                     : slab_test<false>(mn, mx, ray.o, pre.inv_dir, ray.minT, ray.maxT);
}

// ----------------------------------------------------------- primitives ----
// Triangle::tri_hit_template (triangle.h:78-180).  Returns true on hit and
// updates ray.maxT; e0..e2/invDet go to `ph` (ForHitInfo).
VD bool tri_hit(const float* a, uint32_t flags, Ray& ray, const RayPre& pre, PreHit& ph) {
  if (flags & VIMG_LEAF_DEGENERATE) return false;  // length2(cross(e2,e1)) == 0, precomputed
  V3 p0 = mk3(a[0], a[1], a[2]), p1 = mk3(a[3], a[4], a[5]), p2 = mk3(a[6], a[7], a[8]);
  V3 p0t = p0 - ray.o, p1t = p1 - ray.o, p2t = p2 - ray.o;
  // permute: (kx, ky, kz) is a cyclic rotation -- kz = 2: (x,y,z), kz = 0: (y,z,x), kz = 1: (z,x,y)
  const bool z0 = pre.kz == 0, z1 = pre.kz == 1;
  float p0x = z0 ? p0t.y : (z1 ? p0t.z : p0t.x), p0y = z0 ? p0t.z : (z1 ? p0t.x : p0t.y),
        p0z = z0 ? p0t.x : (z1 ? p0t.y : p0t.z);
  float p1x = z0 ? p1t.y : (z1 ? p1t.z : p1t.x), p1y = z0 ? p1t.z : (z1 ? p1t.x : p1t.y),
        p1z = z0 ? p1t.x : (z1 ? p1t.y : p1t.z);
  float p2x = z0 ? p2t.y : (z1 ? p2t.z : p2t.x), p2y = z0 ? p2t.z : (z1 ? p2t.x : p2t.y),
        p2z = z0 ? p2t.x : (z1 ? p2t.y : p2t.z);
  // shear
  p0x += pre.Sx * p0z;
  p0y += pre.Sy * p0z;
  p1x += pre.Sx * p1z;
  p1y += pre.Sy * p1z;
  p2x += pre.Sx * p2z;
  p2y += pre.Sy * p2z;
  // difference_of_products (triangle.h:11-15): cd = c*d; fma(a, b, -cd)
  float e0 = fmaf(p1x, p2y, -(p1y * p2x));
  float e1 = fmaf(p2x, p0y, -(p2y * p0x));
  float e2 = fmaf(p0x, p1y, -(p0y * p1x));
  if (e0 == 0.f || e1 == 0.f || e2 == 0.f) {
    // difference_of_products_double (triangle.h:17-21): cd is the FLOAT product widened,
    // a*b is exact in double; x87 fmal rounds to 64 bits, then to double, then to float.
    // The double fma below rounds the same exact value once to 53 bits (differs from the
    // x87 path only in a double-rounding corner).
    e0 = (float)fma((vdbl)p1x, (vdbl)p2y, -(vdbl)(p1y * p2x));
    e1 = (float)fma((vdbl)p2x, (vdbl)p0y, -(vdbl)(p2y * p0x));
    e2 = (float)fma((vdbl)p0x, (vdbl)p1y, -(vdbl)(p0y * p1x));
  }
  if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
  float det = e0 + e1 + e2;
  if (det == 0) return false;
  p0z *= pre.Sz;
  p1z *= pre.Sz;
  p2z *= pre.Sz;
  float tScaled = e0 * p0z + e1 * p1z + e2 * p2z;
  if (det < 0 && (tScaled >= 0 || tScaled < ray.maxT * det || tScaled > ray.minT * det)) {
    return false;
  } else if (det > 0 && (tScaled <= 0 || tScaled > ray.maxT * det || tScaled < ray.minT * det)) {
    return false;
  }
  float invDet = 1.f / det;
  float t = tScaled * invDet;
  ray.maxT = t;
  ph.e0 = e0;
  ph.e1 = e1;
  ph.e2 = e2;
  ph.inv_det = invDet;
  return true;
}

// Sphere::sphere_hit_template + solveQuadratic (sphere.h:13-100).  `q`, x0, x1 run
// in double because the reference's unqualified sqrt() resolves to ::sqrt(vdbl).
VD bool sphere_hit(const float* a, Ray& r) {
  V3 center = mk3(a[0], a[1], a[2]);
  float radius = a[3];
  float t0, t1;
  const float radius_squared = radius * radius;
  V3 f = r.o - center;
  const float aa = dot(r.dir, r.dir);
  const float b_prime = dot(-1.0f * f, r.dir);
  const float c = dot(f, f) - radius_squared;
  const V3 temp = f + (b_prime / aa) * r.dir;
  const float discriminant = radius_squared - dot(temp, temp);
  if (discriminant < 0) return false;
  {
    float sign = (b_prime > 0) ? 1.0f : -1.0f;
    vdbl q = (vdbl)b_prime + (vdbl)sign * sqrt((vdbl)(aa * discriminant));
    if (discriminant == 0) {
      t0 = t1 = (float)((vdbl)c / q);
    } else {
      t0 = (float)((vdbl)c / q);
      t1 = (float)(q / (vdbl)aa);
    }
    if (t0 > t1) {
      float tmp = t0;
      t0 = t1;
      t1 = tmp;
    }
  }
  if (t0 < r.minT || t0 > r.maxT) {
    t0 = t1;
    if (t0 < r.minT || t0 > r.maxT) return false;
  }
  r.maxT = t0;
  return true;
}

// ------------------------------------------------------------ traversal ----
struct TravCounters {
  uint32_t inner, leaf, tri, sph;
};

#ifndef VIMG_STACK_MAX
#define VIMG_STACK_MAX 64  // bvh.max_depth + 2 entries per traversal (integrators.h:98-99), twice with stale entries below
#endif

// BVH::hit<T> (include/bvh.h:83-225).  ANY_HIT=false: closest hit, ordered push (far child
// first; on a tie the FIRST child is pushed first so the SECOND is visited first), last
// successful primitive wins, popped nodes are not re-tested against the shrunken maxT.
// ANY_HIT=true: BVH::occlude, unordered push (left then right), early exit.
//
// The stack belongs to the caller, like the reference's per-thread `thread_stack`, and that matters:
// BVH::hit pushes the root ON TOP of whatever the vector holds (bvh.h:143, no clear()), and
// BVH::occlude returns from the middle of its loop on the first hit (bvh.h:165-168), leaving the
// nodes it had pushed but not visited.  The closest-hit call that follows on the same thread -- the
// next segment of the same path, mis_integrator.cpp:64 -> :119 -- first runs its own traversal and
// then pops those STALE nodes too, testing their primitives without the box test that would have
// culled them.  That changes results where the float slab test is not conservative: a path leaving
// a wall re-hits the same wall at t just above minT (measured on cornell_skylight: 27 of 16384
// pixels at 64 spp differ without this).  So: `sp` entries of `stack` are kept below the root on
// entry, and an any-hit traversal that ends early leaves its unvisited entries there on return.
template <bool ANY_HIT, bool COUNT>
VD bool bvh_traverse(const DScene& sc, Ray& ray, PreHit& ph, TravCounters* cnt, uint2* stack, int& sp) {
  ph.leaf_slot = -1;
  if (sc.n_nodes == 0) return false;
  const RayPre pre = ray_precompute(ray);
  float root_hit = slab(sc.root_min, sc.root_max, ray, pre);
  if (is_inf(root_hit)) return false;
  if (sp + (int)sc.stack_cap > VIMG_STACK_MAX) sp = 0;  // no room for stale entries below a full traversal

  // "while-while" traversal: every lane first descends through inner nodes until it holds a
  // leaf (or runs dry), then the warp tests leaf primitives together.  Per ray the sequence of
  // box tests, pushes and primitive tests is exactly the reference's push-then-pop order:
  // instead of pushing the child that would be popped next, it is kept in `cur`.
  uint2 cur = make_uint2(sc.root_first, sc.root_count);
  bool have = true;
  while (have) {
    while (have && cur.y == 0) {  // inner: test both children of the pair record
      if (COUNT) cnt->inner++;
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
          // closest hit: push the more distant box first (tie: first child pushed first);
          // any hit: unordered, left pushed then right (bvh.h:195-212)
          if (!ANY_HIT && bb_hit2 > bb_hit1) {
            uint2 t = first_child;
            first_child = sec_child;
            sec_child = t;
          }
          stack[sp++] = first_child;
        }
        cur = sec_child;  // would be pushed last and popped next
      } else if (h1) {
        cur = first_child;
      } else {
        have = sp > 0;
        if (have) cur = stack[--sp];
      }
    }
    if (have) {  // leaf
      if (COUNT) cnt->leaf++;
      for (uint32_t i = 0; i < cur.y; ++i) {
        const uint32_t slot = cur.x + i;
        const float4* rec = reinterpret_cast<const float4*>(sc.leaf_prims + slot);
        const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
        const float a[9] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
        const uint32_t flags = __float_as_uint(r2.z);
        bool hit;
        if (flags & VIMG_LEAF_SPHERE) {
          if (COUNT) cnt->sph++;
          hit = sphere_hit(a, ray);
          if (hit && !ANY_HIT) ph.e0 = ph.e1 = ph.e2 = ph.inv_det = 0.f;
        } else {
          if (COUNT) cnt->tri++;
          hit = tri_hit(a, flags, ray, pre, ph);
        }
        if (hit) {
          if (ANY_HIT) return true;
          ph.leaf_slot = (int32_t)slot;
        }
      }
      have = sp > 0;
      if (have) cur = stack[--sp];
    }
  }
  return ph.leaf_slot >= 0;
}

// a traversal that starts from an empty stack (camera rays of the normal integrators, ids, heat map)
template <bool ANY_HIT, bool COUNT>
VD bool bvh_traverse(const DScene& sc, Ray& ray, PreHit& ph, TravCounters* cnt) {
  uint2 stack[VIMG_STACK_MAX];
  int sp = 0;
  return bvh_traverse<ANY_HIT, COUNT>(sc, ray, ph, cnt, stack, sp);
}

}  // namespace vimg
