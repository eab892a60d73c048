/*
 * vimg_oracle.c -- ORACLE / TEST INFRASTRUCTURE.  Never linked into the product.
 *
 * Plain-C restatement of the BIT-EXACT part of v-img's render hot path, working directly on a
 * flattened scene (include/vimg_cuda.h):
 *
 *   per-pixel PCG32 stream + R2 jitter      include/integrators.h:109-123
 *                                           include/rng/pcg_rand.h:15-33, sampling.h:85-105,228-239
 *   thin-lens camera ray                    src/tl_camera.cpp:25-53, include/ray.h:36-48
 *   closest-hit BVH traversal (scalar path) include/bvh.h:83-225, include/hit_utils.h:134-151
 *   watertight ray/triangle                 include/geometry/triangle.h:78-180
 *   ray/sphere                              include/geometry/sphere.h:13-100
 *   geometric / shading normal images       src/integrators/normals.cpp:4-46,
 *                                           src/geometry/triangle.cpp:13-43, sphere.cpp:12-14
 *
 * It is independent of the CUDA code (different language, different data layout: it walks the
 * reference's own BVH arrays, not the GPU pair records) and of the glm shim (no vector
 * library: every expression is written out).  PINNING: tests/test_oracle.py checks it against
 * the compiled reference (oracle/_ref, parity flavour) through tests/golden -- primary-hit ids,
 * hit distances and both normal images must match bit for bit.  The shading-normal image is
 * restated for meshes WITHOUT normal maps (all JSON scenes); material / mis are not restated
 * here -- for those the compiled reference itself is the oracle.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle/vimg_oracle.c -o oracle/_ref/libvimg_oracle.so -lm
 * (no -march: the reference's parity build has no FMA contraction either.)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/vimg_cuda.h"

typedef struct { float x, y, z; } v3;

static v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
/* glm::dot: tmp = a*b; (tmp.x + tmp.y) + tmp.z */
static float dot(v3 a, v3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
/* glm::normalize: v * (1 / sqrt(dot(v, v))) */
static v3 normalize(v3 v) { return muls(v, 1.0f / sqrtf(dot(v, v))); }
static v3 cross(v3 a, v3 b) { return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
static float comp(v3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

/* ---- include/rng/pcg_rand.h:15-33 ---- */
typedef struct { uint64_t state, inc; } pcg32;
static uint32_t pcg32_random(pcg32* r) {
  uint64_t old = r->state;
  r->state = old * 6364136223846793005ULL + r->inc;
  uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  uint32_t rot = (uint32_t)(old >> 59u);
  return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31));
}
static void pcg32_seed(pcg32* r, uint64_t initstate, uint64_t initseq) {
  r->state = 0U;
  r->inc = (initseq << 1u) | 1u;
  pcg32_random(r);
  r->state += initstate;
  pcg32_random(r);
}
/* ---- include/rng/sampling.h:85-105 ---- */
static float rand_float(pcg32* r) {
  uint64_t r1 = pcg32_random(r), r2 = pcg32_random(r);
  uint64_t u = (r1 << 32) | r2;
  uint32_t z = u ? (uint32_t)__builtin_clzll(u) : 64u;
  if (z <= 40) {
    uint32_t bits = ((126 - z) << 23) | (((uint32_t)u) & 0x7fffff);
    float f;
    memcpy(&f, &bits, 4);
    return f;
  }
  return 0x1.0p-64f * (float)((uint32_t)u);
}
/* ---- include/rng/sampling.h:228-239: g is a float, a1/a2 are folded in double then narrowed ---- */
static void r2_offset(uint32_t n, float* ox, float* oy) {
  const float g = 1.32471795724474602596;
  const float a1 = 1.0 - (1.0 / g);
  const float a2 = 1.0 - (1.0 / (g * g));
  float x = a1 * n, y = a2 * n;
  *ox = x - floorf(x);
  *oy = y - floorf(y);
}

typedef struct { v3 o, dir; float minT, maxT; } ray_t;

/* glm mat4 * vec4: (m0*v0 + m1*v1) + (m2*v2 + m3*v3), column-major */
static void mat4_mul(const float* m, const float v[4], float out[4]) {
  for (int r = 0; r < 4; ++r) out[r] = (m[r] * v[0] + m[4 + r] * v[1]) + (m[8 + r] * v[2] + m[12 + r] * v[3]);
}

/* ---- src/tl_camera.cpp:25-53 (sample_disk: sampling.h:15-22) ---- */
static ray_t generate_ray(const vimg_camera* c, float x, float y, float rand1, float rand2) {
  float x_dir = (c->p_size[0] * (x / c->resolution[0])) - (c->p_size[0] / 2.0f);
  float y_dir = (c->p_size[1] * (y / c->resolution[1])) - (c->p_size[1] / 2.0f);
  v3 ray_dir = normalize(V(x_dir, y_dir, -1.0f));
  ray_t r;
  r.o = V(0.f, 0.f, 0.f);
  r.dir = ray_dir;
  r.minT = 0.0001f;
  r.maxT = INFINITY;
  if (c->aperture_radius > 0.f) {
    float rr = sqrtf(rand1);
    float phi = (float)(2.f * 3.141592653589793238462643383279502884 * rand2);
    v3 ray_origin = V(c->aperture_radius * (rr * cosf(phi)), c->aperture_radius * (rr * sinf(phi)), 0.f);
    float ft = c->focal_dist / fabsf(ray_dir.z);
    v3 focal_plane_p = muls(ray_dir, ft);
    r.o = ray_origin;
    r.dir = normalize(sub(focal_plane_p, ray_origin));
  }
  float d4[4] = {r.dir.x, r.dir.y, r.dir.z, 0.0f}, o4[4] = {r.o.x, r.o.y, r.o.z, 1.0f}, d[4], o[4];
  mat4_mul(c->cam_to_world, d4, d);
  mat4_mul(c->cam_to_world, o4, o);
  r.dir = normalize(V(d[0], d[1], d[2]));
  r.o = V(o[0] / o[3], o[1] / o[3], o[2] / o[3]);
  return r;
}

/* ---- include/hit_utils.h:134-151: glm::min/max and std::max/min written out (NaN order) ---- */
static float gmin(float a, float b) { return (b < a) ? b : a; }
static float gmax(float a, float b) { return (a < b) ? b : a; }
static float slab(const ray_t* r, v3 inv, const float* bmin, const float* bmax) {
  float lx = (bmin[0] - r->o.x) * inv.x, ly = (bmin[1] - r->o.y) * inv.y, lz = (bmin[2] - r->o.z) * inv.z;
  float ux = (bmax[0] - r->o.x) * inv.x, uy = (bmax[1] - r->o.y) * inv.y, uz = (bmax[2] - r->o.z) * inv.z;
  float tBoxMin = gmax(gmin(lx, ux), gmax(gmin(ly, uy), gmax(gmin(lz, uz), r->minT)));
  float tBoxMax = gmin(gmax(lx, ux), gmin(gmax(ly, uy), gmin(gmax(lz, uz), r->maxT)));
  return (tBoxMin <= tBoxMax) ? tBoxMin : INFINITY;
}

typedef struct { float e0, e1, e2, inv_det; int64_t prim; } prehit_t;

static void tri_vertices(const vimg_scene_desc* d, const vimg_prim* p, v3* p0, v3* p1, v3* p2) {
  const vimg_mesh* m = &d->meshes[p->a];
  const uint32_t* idx = m->indices + 3 * (size_t)p->b;
  const float* v = m->vertices;
  *p0 = V(v[3 * idx[0]], v[3 * idx[0] + 1], v[3 * idx[0] + 2]);
  *p1 = V(v[3 * idx[1]], v[3 * idx[1] + 1], v[3 * idx[1] + 2]);
  *p2 = V(v[3 * idx[2]], v[3 * idx[2] + 1], v[3 * idx[2] + 2]);
}

/* ---- include/geometry/triangle.h:78-180 ---- */
static int tri_hit(const vimg_scene_desc* d, const vimg_prim* p, ray_t* ray, prehit_t* ph) {
  v3 p0, p1, p2;
  tri_vertices(d, p, &p0, &p1, &p2);
  v3 c = cross(sub(p2, p0), sub(p1, p0));
  if (dot(c, c) == 0.f) return 0;
  v3 p0t = sub(p0, ray->o), p1t = sub(p1, ray->o), p2t = sub(p2, ray->o);
  float ax = fabsf(ray->dir.x), ay = fabsf(ray->dir.y), az = fabsf(ray->dir.z);
  int kz = 0;
  float mx = ax;
  if (ay > mx) { kz = 1; mx = ay; }
  if (az > mx) { kz = 2; mx = az; }
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  float dx = comp(ray->dir, kx), dy = comp(ray->dir, ky), dz = comp(ray->dir, kz);
  float p0x = comp(p0t, kx), p0y = comp(p0t, ky), p0z = comp(p0t, kz);
  float p1x = comp(p1t, kx), p1y = comp(p1t, ky), p1z = comp(p1t, kz);
  float p2x = comp(p2t, kx), p2y = comp(p2t, ky), p2z = comp(p2t, kz);
  float Sx = -dx / dz, Sy = -dy / dz, Sz = 1.f / dz;
  p0x += Sx * p0z; p0y += Sy * p0z;
  p1x += Sx * p1z; p1y += Sy * p1z;
  p2x += Sx * p2z; p2y += Sy * p2z;
  float cd;
  cd = p1y * p2x; float e0 = fmaf(p1x, p2y, -cd);
  cd = p2y * p0x; float e1 = fmaf(p2x, p0y, -cd);
  cd = p0y * p1x; float e2 = fmaf(p0x, p1y, -cd);
  if (e0 == 0.f || e1 == 0.f || e2 == 0.f) {
    /* difference_of_products_double: double cd = (float product); fmal; narrowed to double, float */
    double c0 = p1y * p2x, c1 = p2y * p0x, c2 = p0y * p1x;
    e0 = (float)(double)fmal(p1x, p2y, -c0);
    e1 = (float)(double)fmal(p2x, p0y, -c1);
    e2 = (float)(double)fmal(p0x, p1y, -c2);
  }
  if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return 0;
  float det = e0 + e1 + e2;
  if (det == 0) return 0;
  p0z *= Sz; p1z *= Sz; p2z *= Sz;
  float tScaled = e0 * p0z + e1 * p1z + e2 * p2z;
  if (det < 0 && (tScaled >= 0 || tScaled < ray->maxT * det || tScaled > ray->minT * det)) return 0;
  else if (det > 0 && (tScaled <= 0 || tScaled > ray->maxT * det || tScaled < ray->minT * det)) return 0;
  float invDet = 1.f / det;
  ray->maxT = tScaled * invDet;
  ph->e0 = e0; ph->e1 = e1; ph->e2 = e2; ph->inv_det = invDet;
  return 1;
}

/* ---- include/geometry/sphere.h:13-100: q, x0, x1 in double (unqualified sqrt) ---- */
static int sphere_hit(const vimg_sphere* s, ray_t* r) {
  v3 center = V(s->center[0], s->center[1], s->center[2]);
  float radius_squared = s->radius * s->radius;
  v3 f = sub(r->o, center);
  float a = dot(r->dir, r->dir);
  float b_prime = dot(muls(f, -1.0f), r->dir);
  float c = dot(f, f) - radius_squared;
  v3 temp = add(f, muls(r->dir, b_prime / a));
  float discriminant = radius_squared - dot(temp, temp);
  if (discriminant < 0) return 0;
  float sign = (b_prime > 0) ? 1.0f : -1.0f;
  double q = b_prime + sign * sqrt((double)(a * discriminant));
  float t0, t1;
  if (discriminant == 0) t0 = t1 = (float)(c / q);
  else { t0 = (float)(c / q); t1 = (float)(q / a); }
  if (t0 > t1) { float t = t0; t0 = t1; t1 = t; }
  if (t0 < r->minT || t0 > r->maxT) {
    t0 = t1;
    if (t0 < r->minT || t0 > r->maxT) return 0;
  }
  r->maxT = t0;
  return 1;
}

/* ---- include/bvh.h:83-225, closest hit, scalar branch ---- */
/* `cost` (nullable): what BVH::hit<float> returns for the heat map (bvh.h:128-131,159-163,189-192):
 * traversal_cost 0.5 for the root box, 2 * 0.5 per inner node popped (scalar branch), intersection_cost
 * 1 per primitive test */
static int64_t bvh_closest_cost(const vimg_scene_desc* d, ray_t* ray, prehit_t* ph, float* cost) {
  const vimg_bvh* b = &d->bvh;
  ph->prim = -1;
  if (cost) *cost = 0.f;
  if (b->n_nodes == 0) return -1;
  v3 inv = V(1.0f / ray->dir.x, 1.0f / ray->dir.y, 1.0f / ray->dir.z);
  const float* bb = b->bb_mins_maxes;
  const float root_hit = slab(ray, inv, bb + 0, bb + 6);
  if (cost) *cost += 0.5f;
  if (isinf(root_hit)) return -1;
  uint64_t stack[128];
  int sp = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    uint64_t n = stack[--sp];
    uint32_t first = b->nodes[2 * n], count = b->nodes[2 * n + 1];
    if (count != 0) {
      for (uint32_t i = 0; i < count; ++i) {
        uint64_t prim = b->obj_indices[first + i];
        const vimg_prim* p = &d->prims[prim];
        int hit;
        if (cost) *cost += 1.f;
        if (p->type == VIMG_PRIM_SPHERE) {
          hit = sphere_hit(&d->spheres[p->a], ray);
          if (hit) ph->e0 = ph->e1 = ph->e2 = ph->inv_det = 0.f;
        } else {
          hit = tri_hit(d, p, ray, ph);
        }
        if (hit) ph->prim = (int64_t)prim;  /* last successful primitive wins (bvh.h:155-158) */
      }
    } else {
      uint64_t first_child = first, sec_child = first_child + 1;
      uint64_t lmin = first_child * 2 + 2, lmax = lmin + 2, rmin = lmin + 1, rmax = lmax + 1;
      float h1 = slab(ray, inv, bb + 3 * lmin, bb + 3 * lmax);
      float h2 = slab(ray, inv, bb + 3 * rmin, bb + 3 * rmax);
      if (cost) *cost += 0.5f * 2.f;
      if (!isinf(h2)) {
        if (!isinf(h1)) {
          if (h2 > h1) { uint64_t t = first_child; first_child = sec_child; sec_child = t; }
          stack[sp++] = first_child;
        }
        stack[sp++] = sec_child;
      } else if (!isinf(h1)) {
        stack[sp++] = first_child;
      }
    }
  }
  return ph->prim;
}
static int64_t bvh_closest(const vimg_scene_desc* d, ray_t* ray, prehit_t* ph) {
  return bvh_closest_cost(d, ray, ph, 0);
}

/* hit_n_s / hit_n_g of Triangle::hit_info (triangle.cpp:13-43) and Sphere::hit_info (sphere.cpp:12-14);
 * returns 0 if the mesh's material has a normal map (not restated) */
static int hit_normals(const vimg_scene_desc* d, const ray_t* ray, const prehit_t* ph, v3* n_s, v3* n_g) {
  const vimg_prim* p = &d->prims[ph->prim];
  if (p->type == VIMG_PRIM_SPHERE) {
    const vimg_sphere* s = &d->spheres[p->a];
    v3 hit_p = add(ray->o, muls(ray->dir, ray->maxT));
    *n_s = *n_g = normalize(sub(hit_p, V(s->center[0], s->center[1], s->center[2])));
    return 1;
  }
  const vimg_mesh* m = &d->meshes[p->a];
  if (d->materials[m->material].normal_map >= 0) return 0;
  v3 p0, p1, p2;
  tri_vertices(d, p, &p0, &p1, &p2);
  float u = ph->e0 * ph->inv_det, v = ph->e1 * ph->inv_det, w = ph->e2 * ph->inv_det;
  v3 tri_normal = normalize(cross(sub(p1, p0), sub(p2, p0)));
  *n_g = tri_normal;
  if (m->normals) {
    const uint32_t* idx = m->indices + 3 * (size_t)p->b;
    const float* nn = m->normals;
    v3 n0 = V(nn[3 * idx[0]], nn[3 * idx[0] + 1], nn[3 * idx[0] + 2]);
    v3 n1 = V(nn[3 * idx[1]], nn[3 * idx[1] + 1], nn[3 * idx[1] + 2]);
    v3 n2 = V(nn[3 * idx[2]], nn[3 * idx[2] + 1], nn[3 * idx[2] + 2]);
    *n_s = normalize(add(add(muls(n0, u), muls(n1, v)), muls(n2, w)));
  } else {
    *n_s = tri_normal;
  }
  return 1;
}

/* Primary-hit ids / distances for all samples and (optionally) the two normal images.
 * out_ids, out_t: W*H*spp (index = image_index*spp + s); s_normal, g_normal: W*H*3 or NULL.
 * Returns 0, or 1 if a normal-mapped mesh was hit while normal images were requested. */
int vimg_oracle_primary(const vimg_scene_desc* d, uint32_t spp, int32_t* out_ids, float* out_t,
                        float* s_normal, float* g_normal) {
  const uint32_t W = (uint32_t)d->camera.resolution[0], H = (uint32_t)d->camera.resolution[1];
  int unsupported = 0;
  for (uint32_t y = 0; y < H; ++y)
    for (uint32_t x = 0; x < W; ++x) {
      size_t image_index = x + ((size_t)(H - 1 - y) * W);
      pcg32 st;
      pcg32_seed(&st, image_index, 0);
      v3 acc_s = V(0, 0, 0), acc_g = V(0, 0, 0);
      for (uint32_t s = 0; s < spp; ++s) {
        float ox, oy;
        r2_offset(x + y + s, &ox, &oy);
        float rand2 = rand_float(&st); /* g++ evaluates the call's arguments right to left */
        float rand1 = rand_float(&st);
        ray_t ray = generate_ray(&d->camera, (float)x + ox, (float)y + oy, rand1, rand2);
        prehit_t ph;
        int64_t prim = bvh_closest(d, &ray, &ph);
        if (out_ids) out_ids[image_index * spp + s] = (int32_t)prim;
        if (out_t) out_t[image_index * spp + s] = ray.maxT;
        if (s_normal || g_normal) {
          v3 cs, cg;
          if (prim >= 0) {
            v3 ns, ng;
            if (!hit_normals(d, &ray, &ph, &ns, &ng)) { unsupported = 1; ns = ng = V(0, 0, 0); }
            cs = V((ns.x + 1.0f) / 2.0f, (ns.y + 1.0f) / 2.0f, (ns.z + 1.0f) / 2.0f);
            cg = V((ng.x + 1.0f) / 2.0f, (ng.y + 1.0f) / 2.0f, (ng.z + 1.0f) / 2.0f);
          } else { /* normals.cpp:17-22: a is computed in double */
            v3 ud = normalize(ray.dir);
            float a = (float)(0.5 * (ud.y + 1.0));
            cs = cg = add(muls(V(1.0f, 1.0f, 1.0f), 1.0f - a), muls(V(0.5f, 0.7f, 1.0f), a));
          }
          acc_s = add(acc_s, cs);
          acc_g = add(acc_g, cg);
        }
      }
      if (s_normal) { s_normal[3 * image_index] = acc_s.x / (float)spp; s_normal[3 * image_index + 1] = acc_s.y / (float)spp; s_normal[3 * image_index + 2] = acc_s.z / (float)spp; }
      if (g_normal) { g_normal[3 * image_index] = acc_g.x / (float)spp; g_normal[3 * image_index + 1] = acc_g.y / (float)spp; g_normal[3 * image_index + 2] = acc_g.z / (float)spp; }
    }
  return unsupported;
}

/* ============================================================================================
 * Post-process chain: src/main.cpp:304-356
 *
 *   tonemapper        include/color_utils.h:21-25 (clamp), src/tonemap/aces.cpp:5-31,
 *                     src/tonemap/agx.cpp:6-99, src/tonemap/reinhard.cpp:3-37
 *   sRGB OETF         include/color_utils.h:50-68
 *   8-bit quantise    src/main.cpp:339-356 (NaN pixel -> magenta, int(255.999 * v) clamped)
 *
 * Every glm expression is written out per component in glm's evaluation order (mat3 * vec3 =
 * m[0][r]*v.x + m[1][r]*v.y + m[2][r]*v.z; clamp = min(max(x, lo), hi) with glm's ternaries, which
 * let NaN through).  powf / log2f are the host libm's, i.e. what the reference calls.
 * PINNING: tests/test_oracle.py compares it byte for byte and float for float with the compiled
 * reference's own functions (oracle/_ref, vref_postprocess) on the golden images and on a
 * synthetic ramp with NaN / inf / negative / huge values.
 * ============================================================================================ */
static float glm_min(float a, float b) { return (b < a) ? b : a; }
static float glm_max(float a, float b) { return (a < b) ? b : a; }
static float glm_clamp(float x, float lo, float hi) { return glm_min(glm_max(x, lo), hi); }
/* color_utils.h:9-11 */
static float pp_luminance(v3 c) { return dot(c, V(0.212671f, 0.715160f, 0.072169f)); }
static v3 mat3_mul(const float m[9], v3 v) { /* m column-major: m[3*col + row] */
  return V(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
           m[2] * v.x + m[5] * v.y + m[8] * v.z);
}

/* aces.cpp:5-25 */
static float aces_fit1(float v) {
  float a = v * (v + 0.0245786f) - 0.000090537f;
  float b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
  return a / b;
}
static v3 pp_aces(v3 v) {
  static const float in_m[9] = {0.59719f, 0.07600f, 0.02840f, 0.35458f, 0.90834f, 0.13383f,
                                0.04823f, 0.01566f, 0.83777f};
  static const float out_m[9] = {1.60475f, -0.10208f, -0.00327f, -0.53108f, 1.10813f, -0.07276f,
                                 -0.07367f, -0.00605f, 1.07602f};
  v = mat3_mul(in_m, v);
  v = V(aces_fit1(v.x), aces_fit1(v.y), aces_fit1(v.z));
  return mat3_mul(out_m, v);
}

/* agx.cpp:6-14: (((((15.5*x4)*x2 - (40.14*x4)*x) + 31.96*x4) - (6.868*x2)*x) + 0.4298*x2) + 0.1191*x) - 0.00232 */
static float agx_contrast1(float x) {
  float x2 = x * x, x4 = x2 * x2;
  return 15.5f * x4 * x2 - 40.14f * x4 * x + 31.96f * x4 - 6.868f * x2 * x + 0.4298f * x2 + 0.1191f * x
         - 0.00232f;
}
static v3 pp_agx(v3 val) {
  /* agx.cpp:17-19, 41-44: double literals narrowed to float by glm::mat3's constructor */
  static const float agx_mat[9] = {(float)0.842479062253094, (float)0.0423282422610123, (float)0.0423756549057051,
                                   (float)0.0784335999999992, (float)0.878468636469772, (float)0.0784336,
                                   (float)0.0792237451477643, (float)0.0791661274605434, (float)0.879142973793104};
  static const float agx_mat_inv[9] = {(float)1.19687900512017, (float)-0.0528968517574562, (float)-0.0529716355144438,
                                       (float)-0.0980208811401368, (float)1.15190312990417, (float)-0.0980434501171241,
                                       (float)-0.0990297440797205, (float)-0.0989611768448433, (float)1.15107367264116};
  const float min_ev = -12.47393f, max_ev = 4.026069f;
  /* agx(): inset, log2 encoding, sigmoid (agx.cpp:16-37) */
  val = mat3_mul(agx_mat, val);
  val = V(glm_clamp(log2f(val.x), min_ev, max_ev), glm_clamp(log2f(val.y), min_ev, max_ev),
          glm_clamp(log2f(val.z), min_ev, max_ev));
  val = V((val.x - min_ev) / (max_ev - min_ev), (val.y - min_ev) / (max_ev - min_ev),
          (val.z - min_ev) / (max_ev - min_ev));
  val = V(agx_contrast1(val.x), agx_contrast1(val.y), agx_contrast1(val.z));
  /* agxLook(): slope 1, offset 0, power 1, sat 1 (agx.cpp:69-82) */
  val = V(powf(val.x * 1.0f + 0.0f, 1.0f), powf(val.y * 1.0f + 0.0f, 1.0f), powf(val.z * 1.0f + 0.0f, 1.0f));
  {
    float luma = pp_luminance(val);
    val = V(luma + 1.0f * (val.x - luma), luma + 1.0f * (val.y - luma), luma + 1.0f * (val.z - luma));
  }
  /* agxEotf(): outset, negative -> 0, pow 2.2 (agx.cpp:39-66) */
  val = mat3_mul(agx_mat_inv, val);
  if (val.x < 0.f) val.x = 0.f;
  if (val.y < 0.f) val.y = 0.f;
  if (val.z < 0.f) val.z = 0.f;
  return V(powf(val.x, 2.2f), powf(val.y, 2.2f), powf(val.z, 2.2f));
}

/* color_utils.h:50-68, one channel */
static float srgb_oetf1(float p) {
  p = glm_clamp(p, 0.0f, 1.0f);
  if (p < 0.0031308f) return p * 12.92f;
  return 1.055f * powf(p, 1.0f / 2.4f) - 0.055f;
}

/* tonemapper: 0 clamp, 1 agx, 2 reinhard, 3 aces, else none.  out_srgb / out_rgb8 nullable. */
void vimg_oracle_postprocess(const float* rgb, uint64_t n_pixels, int tonemapper, float* out_srgb,
                             uint8_t* out_rgb8) {
  float largest_L = 0.0f;
  if (tonemapper == 2) { /* reinhard.cpp:3-13 */
    for (uint64_t i = 0; i < n_pixels; ++i) {
      float l = pp_luminance(V(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
      if (l > largest_L) largest_L = l;
    }
  }
  for (uint64_t i = 0; i < n_pixels; ++i) {
    v3 c = V(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    switch (tonemapper) {
      case 0: c = V(glm_clamp(c.x, 0.f, 1.f), glm_clamp(c.y, 0.f, 1.f), glm_clamp(c.z, 0.f, 1.f)); break;
      case 1: c = pp_agx(c); break;
      case 2: { /* reinhard.cpp:16-37 */
        float in_L = pp_luminance(c);
        float numerator = in_L * (1.0f + (in_L / (largest_L * largest_L)));
        float new_L = numerator / (1.0f + in_L);
        float l_in = pp_luminance(c);
        if (l_in > 0.f) c = muls(c, new_L / l_in);
        else c = V(0.f, 0.f, 0.f);
        break;
      }
      case 3: c = pp_aces(c); break;
      default: break;
    }
    c = V(srgb_oetf1(c.x), srgb_oetf1(c.y), srgb_oetf1(c.z));
    if (out_srgb) {
      out_srgb[3 * i] = c.x;
      out_srgb[3 * i + 1] = c.y;
      out_srgb[3 * i + 2] = c.z;
    }
    if (out_rgb8) {
      if (isnan(c.x) || isnan(c.y) || isnan(c.z)) { /* main.cpp:342-346 */
        out_rgb8[3 * i] = 255;
        out_rgb8[3 * i + 1] = 0;
        out_rgb8[3 * i + 2] = 255;
      } else { /* main.cpp:348-351: double product, truncation, std::clamp(int, 0, 255) */
        int r = (int)(255.999 * c.x), g = (int)(255.999 * c.y), b = (int)(255.999 * c.z);
        out_rgb8[3 * i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
        out_rgb8[3 * i + 1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
        out_rgb8[3 * i + 2] = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
      }
    }
  }
}

/* ============================================================================================
 * BVH heat map: src/integrators/heatmap.cpp:20-147
 *   per pixel: the integrators' camera rays (same PCG stream, R2 jitter), BVH::hit<float> cost of
 *   each, float sum, static_cast<uint32_t>(sum / spp) stored as float, turbo_colormap(v / factor).
 * glm::dot(vec4) = (x*x' + y*y') + (z*z' + w*w'), dot(vec2) = x*x' + y*y'; the colour-map constants
 * are double literals narrowed to float by the glm constructors.
 * out_rgb: W*H*3; out_cost (nullable): W*H, the truncated mean cost.
 * ============================================================================================ */
static float dot4(const float a[4], const float b[4]) {
  float tx = a[0] * b[0], ty = a[1] * b[1], tz = a[2] * b[2], tw = a[3] * b[3];
  return (tx + ty) + (tz + tw);
}
static float dot2(const float a[2], const float b[2]) { return a[0] * b[0] + a[1] * b[1]; }
static v3 turbo_colormap(float x) { /* heatmap.cpp:20-37 */
  static const float r4[4] = {(float)0.13572138, (float)4.61539260, (float)-42.66032258, (float)132.13108234};
  static const float g4[4] = {(float)0.09140261, (float)2.19418839, (float)4.84296658, (float)-14.18503333};
  static const float b4[4] = {(float)0.10667330, (float)12.64194608, (float)-60.58204836, (float)110.36276771};
  static const float r2[2] = {(float)-152.94239396, (float)59.28637943};
  static const float g2[2] = {(float)4.27729857, (float)2.82956604};
  static const float b2[2] = {(float)-89.90310912, (float)27.34824973};
  x = (x < 0.f) ? 0.f : ((1.f < x) ? 1.f : x); /* std::clamp */
  const float v4[4] = {1.0f, x, x * x, x * x * x};
  const float v2[2] = {v4[2] * v4[2], v4[3] * v4[2]};
  return V(dot4(v4, r4) + dot2(v2, r2), dot4(v4, g4) + dot2(v2, g2), dot4(v4, b4) + dot2(v2, b2));
}
void vimg_oracle_heatmap(const vimg_scene_desc* d, uint32_t spp, float factor, float* out_rgb, float* out_cost) {
  const uint32_t W = (uint32_t)d->camera.resolution[0], H = (uint32_t)d->camera.resolution[1];
  if (factor <= 0) factor = 20.f; /* heatmap.cpp:135-137 */
  for (uint32_t y = 0; y < H; ++y)
    for (uint32_t x = 0; x < W; ++x) {
      size_t image_index = x + ((size_t)(H - 1 - y) * W);
      pcg32 st;
      pcg32_seed(&st, image_index, 0);
      float acc = 0.f;
      for (uint32_t s = 0; s < spp; ++s) {
        float ox, oy, cost;
        r2_offset(x + y + s, &ox, &oy);
        float rand2 = rand_float(&st);
        float rand1 = rand_float(&st);
        ray_t ray = generate_ray(&d->camera, (float)x + ox, (float)y + oy, rand1, rand2);
        prehit_t ph;
        bvh_closest_cost(d, &ray, &ph, &cost);
        acc += cost;
      }
      const float hits = (float)(uint32_t)(acc / spp);
      if (out_cost) out_cost[image_index] = hits;
      v3 c = turbo_colormap(hits / factor);
      out_rgb[3 * image_index] = c.x;
      out_rgb[3 * image_index + 1] = c.y;
      out_rgb[3 * image_index + 2] = c.z;
    }
}

/* ============================================================================================
 * Scene preparation
 *   mip chain        ImageTexture::ImageTexture, src/image_texture.cpp:59-130 (Wronski 8-tap filter
 *                    over bilinear fetches, col_at_uv_mipmap :132-161, handle_wrapping
 *                    include/texture/texture_common.h:22-53)
 *   env-map CDFs     ArraySampling2D / ArraySampling1D, include/rng/sampling.h:114-135,167-197
 * PINNING: tests/test_prep.py -- the golden .vscene files carry the levels and CDFs the
 * reference's constructors built; level 0 in, everything else out, bit for bit.
 * ============================================================================================ */
static int clampi(int v, int lo, int hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }
static float std_clampf(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }
static float handle_wrapping(float coord, uint32_t mode) { /* texture_common.h:22-53 */
  switch (mode) {
    case VIMG_WRAP_REPEAT: {
      float fraction = coord - (float)(int)coord;
      return signbit(fraction) ? 1.f + fraction : fraction;
    }
    case VIMG_WRAP_MIRRORED_REPEAT: {
      int int_part = (int)coord;
      float fraction = coord - (float)int_part;
      if (signbit(fraction)) {
        if (int_part % 2) return fabsf(fraction);
        return 1.f + fraction;
      }
      return fraction;
    }
    default: return std_clampf(coord, 0.f, 1.f);
  }
}
static v3 mix3(v3 x, v3 y, float a) { /* glm::mix(x, y, a) = x * (1 - a) + y * a */
  return add(muls(x, 1.0f - a), muls(y, a));
}
static v3 texel(const float* lvl, size_t i) { return V(lvl[3 * i], lvl[3 * i + 1], lvl[3 * i + 2]); }
static v3 bilinear_level(const float* lvl, uint32_t mip_w, uint32_t mip_h, uint32_t wrap_u, uint32_t wrap_v,
                         float u, float v) { /* image_texture.cpp:132-161 */
  float pixel_u = handle_wrapping(u, wrap_u) * mip_w;
  float pixel_v = handle_wrapping(v, wrap_v) * mip_h;
  int curr_x = clampi((int)pixel_u, 0, (int)mip_w - 1), curr_y = clampi((int)pixel_v, 0, (int)mip_h - 1);
  int next_x = clampi(curr_x + 1, 0, (int)mip_w - 1), next_y = clampi(curr_y + 1, 0, (int)mip_h - 1);
  float x_fraction = pixel_u - curr_x, y_fraction = pixel_v - curr_y;
  v3 a = mix3(texel(lvl, curr_x + (size_t)curr_y * mip_w), texel(lvl, next_x + (size_t)curr_y * mip_w), x_fraction);
  v3 b = mix3(texel(lvl, curr_x + (size_t)next_y * mip_w), texel(lvl, next_x + (size_t)next_y * mip_w), x_fraction);
  return mix3(a, b, y_fraction);
}
/* number of levels the constructor builds (image_texture.cpp:73-78), level 0 included */
uint32_t vimg_oracle_mip_levels(uint32_t width, uint32_t height) {
  uint32_t size = width < height ? width : height;
  int n = (int)ceilf(log2f((float)size));
  if (n > 15) n = 15;
  return (uint32_t)(n < 1 ? 1 : n);
}
/* levels 1 .. n-1 one after the other into out_levels */
void vimg_oracle_mip_chain(const float* level0, uint32_t width, uint32_t height, uint32_t wrap_u, uint32_t wrap_v,
                           float* out_levels) {
  const uint32_t n = vimg_oracle_mip_levels(width, height);
  const float* prev = level0;
  float* next = out_levels;
  uint32_t pw = width, ph = height;
  const float k = (float)0.75777, f = (float)2.907;
  const float tap[8][2] = {{-k, -k}, {k, -k}, {k, k}, {-k, k}, {-f, 0.f}, {f, 0.f}, {0.f, -f}, {0.f, f}};
  for (uint32_t l = 1; l < n; ++l) {
    const uint32_t nw = pw / 2u > 1u ? pw / 2u : 1u, nh = ph / 2u > 1u ? ph / 2u : 1u;
    for (uint32_t y = 0; y < nh; ++y)
      for (uint32_t x = 0; x < nw; ++x) {
        const float irx = 1.f / pw, iry = 1.f / ph;
        const float cu = (float)(2 * x) * irx, cv = (float)(2 * y) * iry;
        v3 sum = V(0.f, 0.f, 0.f);
        for (int t = 0; t < 8; ++t) {
          const float w = t < 4 ? 0.37487566f : -0.12487566f;
          v3 c = bilinear_level(prev, pw, ph, wrap_u, wrap_v, cu + tap[t][0] * irx, cv + tap[t][1] * iry);
          sum = add(sum, muls(c, w));
        }
        if (sum.x < 0) sum.x = 0.f;
        if (sum.y < 0) sum.y = 0.f;
        if (sum.z < 0) sum.z = 0.f;
        float* o = next + 3 * ((size_t)x + (size_t)y * nw);
        o[0] = sum.x;
        o[1] = sum.y;
        o[2] = sum.z;
      }
    prev = next;
    next += (size_t)nw * nh * 3;
    pw = nw;
    ph = nh;
  }
}
/* ArraySampling1D (sampling.h:114-135) on n values: cdf has n + 1 entries; returns func_int */
static float cdf_1d(const float* f, size_t n, float* cdf) {
  cdf[0] = 0.f;
  for (size_t x = 1; x < n + 1; ++x) cdf[x] = cdf[x - 1] + fabsf(f[x - 1]);
  const float func_int = cdf[n];
  if (func_int == 0)
    for (size_t i = 0; i < n + 1; ++i) cdf[i] = (float)i / (float)n;
  else
    for (size_t i = 0; i < n + 1; ++i) cdf[i] /= func_int;
  return func_int;
}
/* ArraySampling2D (sampling.h:167-197): row_cdf height + 1, col_cdfs height * (width + 1); `scratch`:
 * width + height floats */
void vimg_oracle_envmap_cdf(const float* image, uint32_t width, uint32_t height, float* row_cdf, float* col_cdfs,
                            float* scratch) {
  float* lum = scratch;
  float* row_int = scratch + width;
  for (uint32_t y = 0; y < height; ++y) {
    float v = ((float)y + 0.5f) / (float)height;
    float sin_elevation = (float)sin(3.141592653589793238462643383279502884 * v);
    for (uint32_t x = 0; x < width; ++x)
      lum[x] = pp_luminance(texel(image, (size_t)y * width + x)) * sin_elevation;
    row_int[y] = cdf_1d(lum, width, col_cdfs + (size_t)y * (width + 1));
  }
  cdf_1d(row_int, height, row_cdf);
}

/* ---- glTF texture conversions (src/scene_loading/gltf_loading.cpp:139-162) --------------------
 * make_texture's byte -> float loop (:139-143) followed by
 *   ImageTexture::convert_sRGB_to_linear  src/image_texture.cpp:257-263: image /= 255.f, then
 *       pix_sRGB_to_linear (include/color_utils.h:28-47) per component, std::pow = the host's powf
 *   ImageTexture::convert_RGB_to_normal   src/image_texture.cpp:265-275: image / 127.5f - 1,
 *       x and y times scale, glm::normalize
 * PINNING: tests/golden/prep/convert.npz holds what the compiled reference's own two functions
 * return for every byte value and a random image (tests/golden/make_golden_convert.py). */
void vimg_oracle_srgb8_to_linear(const uint8_t* rgb8, uint64_t n_pixels, float* out) {
  for (uint64_t i = 0; i < 3 * n_pixels; ++i) {
    float p = (float)rgb8[i];
    p = p / 255.f;
    out[i] = p <= 0.04045f ? p / 12.92f : powf((p + 0.055f) / 1.055f, 2.4f);
  }
}

void vimg_oracle_rgb8_to_normal(const uint8_t* rgb8, uint64_t n_pixels, float scale, float* out) {
  for (uint64_t i = 0; i < n_pixels; ++i) {
    v3 n = V((float)rgb8[3 * i] / 127.5f - 1.f, (float)rgb8[3 * i + 1] / 127.5f - 1.f,
             (float)rgb8[3 * i + 2] / 127.5f - 1.f);
    n.x *= scale;
    n.y *= scale;
    n = normalize(n);
    out[3 * i] = n.x;
    out[3 * i + 1] = n.y;
    out[3 * i + 2] = n.z;
  }
}

/* ---- the host libm on arrays (what the reference's std::sin / std::pow / ... on floats call) ----
 * fn numbering of vimg_libm_eval (include/vimg_cuda.h); lets tests/test_libm.py compare the device's
 * restatements with THIS machine's glibc bit for bit. */
void vimg_oracle_libm(uint32_t fn, const float* x, const float* y, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) {
    const float a = x[i], b = y ? y[i] : 0.f;
    switch (fn) {
      case 0: out[i] = sinf(a); break;
      case 1: out[i] = cosf(a); break;
      case 2: out[i] = logf(a); break;
      case 3: out[i] = log2f(a); break;
      case 4: out[i] = powf(a, b); break;
      case 5: out[i] = acosf(a); break;
      case 6: out[i] = atan2f(a, b); break;
      case 7: out[i] = atanf(a); break;
      default: {  /* 8 / 9: low / high word of the double log2 of the widened input */
        double r = log2((double)a);
        uint64_t u;
        uint32_t w;
        memcpy(&u, &r, 8);
        w = fn == 8 ? (uint32_t)u : (uint32_t)(u >> 32);
        memcpy(&out[i], &w, 4);
        break;
      }
    }
  }
}
