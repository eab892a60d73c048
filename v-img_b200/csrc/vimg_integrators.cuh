// vimg_integrators.cuh -- emitters, background and the four integrators.
#pragma once
#include "vimg_disney.cuh"

namespace vimg {

// ----------------------------------------------------- material dispatch ----
VD bool mat_sample(const DScene& sc, const vimg_material& m, V3 wi, const HitInfo& hit, Pcg32& rng,
                   bool regularize, ScatterInfo& out) {
  switch (m.type) {
    case VIMG_MAT_LAMBERTIAN: return lambertian_sample(wi, hit, rng, out);
    case VIMG_MAT_DIELECTRIC: return dielectric_sample(m, wi, hit, rng, out);
    case VIMG_MAT_PRINCIPLED: return principled_sample(sc, m, wi, hit, rng, regularize, out);
    default: return false;  // DiffuseLight / base: std::nullopt (material.h:37-40)
  }
}

// eval_pdf_pair; base default (0, 1) also covers Dielectric and DiffuseLight (material.h:56-60)
VD void mat_eval_pdf_pair(const DScene& sc, const vimg_material& m, V3 wi, V3 wo, const HitInfo& hit,
                          RayCone cone, bool regularize, V3& eval, float& pdf) {
  switch (m.type) {
    case VIMG_MAT_LAMBERTIAN: lambertian_eval_pdf(sc, m, wi, wo, hit, cone, eval, pdf); return;
    case VIMG_MAT_PRINCIPLED: principled_eval_pdf(sc, m, wi, wo, hit, cone, regularize, eval, pdf); return;
    default:
      eval = mk3(0.f);
      pdf = 1.0f;
  }
}

// eval_div_pdf; base default 0 (material.h:51-54)
VD V3 mat_eval_div_pdf(const DScene& sc, const vimg_material& m, V3 wi, V3 wo, const HitInfo& hit,
                       RayCone cone, bool regularize) {
  switch (m.type) {
    case VIMG_MAT_LAMBERTIAN: return col_at_ray_hit(sc, m.texture, wi, cone, hit);
    case VIMG_MAT_DIELECTRIC: return mk3(1.f);
    case VIMG_MAT_PRINCIPLED: {
      V3 e;
      float p;
      principled_eval_pdf(sc, m, wi, wo, hit, cone, regularize, e, p);
      return e / p;
    }
    default: return mk3(0.f);
  }
}

// ------------------------------------------------------------ background ----
VD V3 xform_dir(const float* m, V3 d) {
  V4 r = mat4_mul(m, mk4(d.x, d.y, d.z, 0.0f));
  return mk3(r.x, r.y, r.z);
}

VD void envmap_uv(const DScene& sc, V3 in_dir, float& u, float& v) {
  V3 dir = normalize(xform_dir(sc.bg.world_to_env, in_dir));
  u = (float)((1.f + (vdbl)m_atan2f(-dir.x, dir.z) * VIMG_INV_PI_D) * 0.5f);
  v = (float)((vdbl)m_acosf(dir.y) * VIMG_INV_PI_D);
}

// ConstBackground / EnvMap::background_emit (background.h:34-36, 88-108)
VD V3 background_emit(const DScene& sc, V3 in_dir, RayCone cone) {
  if (VIMG_LEAN || sc.bg.type == VIMG_BG_CONST) return mk3(sc.bg.color[0], sc.bg.color[1], sc.bg.color[2]);
  float u, v;
  envmap_uv(sc, in_dir, u, v);
  const DTexture& img = sc.textures[sc.bg.image];
  // std::log2(vdbl): |spread| * (height / pi)
  float lambda = (float)d_log2((vdbl)fabsf(cone.spread_angle) * ((vdbl)img.height / VIMG_PI_D));
  lambda = is_nan(lambda) ? 0.f : lambda;
  return col_mipmap_interpolate(sc, img, lambda - 2.f, mk2(u, v)) * sc.bg.radiance_scale;
}

// background_pdf (background.h:56, 141-169)
VD float background_pdf(const DScene& sc, V3 in_dir) {
  if (VIMG_LEAN || sc.bg.type == VIMG_BG_CONST) return (float)(1.f / (4 * VIMG_PI_D));
  float u, v;
  envmap_uv(sc, in_dir, u, v);
  const uint32_t W = sc.bg.cdf_w, H = sc.bg.cdf_h;
  const DTexture& img = sc.textures[sc.bg.image];
  int pixel_u = (int)(u * img.width);
  int pixel_v = (int)(v * img.height);
  int column_index = std_clampi(pixel_u, 0, (int)img.width - 1);
  int row_index = std_clampi(pixel_v, 0, (int)img.height - 1);
  float pdf_y = sc.bg.row_cdf[row_index + 1] - sc.bg.row_cdf[row_index];
  const float* row = sc.bg.col_cdfs + (size_t)row_index * (W + 1);
  float pdf_x = row[column_index + 1] - row[column_index];
  float sin_elevation = (float)d_sin(VIMG_PI_D * (vdbl)v);
  (void)H;
  return (float)((vdbl)(pdf_y * pdf_x * img.width * img.height)
                 / (2.f * VIMG_PI_D * VIMG_PI_D * (vdbl)sin_elevation));
}

// ------------------------------------------------------------------ lights ----
// Triangle::sample (triangle.cpp:178-233)
VD V3 triangle_light_sample(const DScene& sc, const DLightRec& lr, V3 look_from, Pcg32& rng, EmitterInfo& info) {
  V3 p0 = mk3(lr.p[0], lr.p[1], lr.p[2]), p1 = mk3(lr.p[3], lr.p[4], lr.p[5]), p2 = mk3(lr.p[6], lr.p[7], lr.p[8]);
  const V3 edge1 = p1 - p0, edge2 = p2 - p0;
  V3 tri_normal = normalize(cross(edge1, edge2));
  V3 n0, n1, n2;
  if (lr.has_normals) {
    n0 = mk3(lr.n[0], lr.n[1], lr.n[2]);
    n1 = mk3(lr.n[3], lr.n[4], lr.n[5]);
    n2 = mk3(lr.n[6], lr.n[7], lr.n[8]);
  } else {
    n0 = n1 = n2 = tri_normal;
  }
  float rand1 = rand_float(rng);
  float rand2 = rand_float(rng);
  float u, v;
  if (rand1 < rand2) {
    u = rand1 / 2.f;
    v = rand2 - u;
  } else {
    v = rand2 / 2.f;
    u = rand1 - v;
  }
  float w = 1.f - u - v;
  const V3 hit_p = p0 * u + p1 * v + p2 * w;
  V3 hit_n = normalize(u * n0 + v * n1 + w * n2);
  V3 dir_vec = hit_p - look_from;
  float dist2 = length2(dir_vec);
  dir_vec = normalize(dir_vec);
  float area = length(cross(edge2, edge1)) / 2.0f;
  float pdf = 1.f / area;
  float cosine = fabsf(dot(hit_n, -dir_vec));
  float G = cosine / dist2;
  info.wi = dir_vec;
  info.pdf = pdf;
  info.dist = sqrtf(dist2);
  info.G = G;
  return mat_emitted(sc.materials[lr.material], info.wi, hit_n);
}

// Triangle::surf_pdf (triangle.cpp:235-248)
VD float triangle_surf_pdf(const DScene& sc, uint32_t prim) {
  const vimg_prim pr = sc.prims[prim];
  const DMesh& mesh = sc.meshes[pr.a];
  const uint32_t* idx = sc.tri_indices + 3 * ((size_t)mesh.tri_base + pr.b);
  V3 p0 = load3(sc.vertices, idx[0] + mesh.vert_base), p1 = load3(sc.vertices, idx[1] + mesh.vert_base),
     p2 = load3(sc.vertices, idx[2] + mesh.vert_base);
  V3 edge1 = p1 - p0, edge2 = p2 - p0;
  float area = length(cross(edge2, edge1)) / 2.0f;
  return 1.f / area;
}

// Sphere::sample (sphere.cpp:58-118)
VD V3 sphere_light_sample(const DScene& sc, const DLightRec& lr, V3 look_from, Pcg32& rng, EmitterInfo& info) {
  const V3 center = mk3(lr.p[0], lr.p[1], lr.p[2]);
  const float radius = lr.p[3];
  float rand1 = rand_float(rng);
  float rand2 = rand_float(rng);
  V3 shading_normal;
  if (length2(look_from - center) <= radius * radius) {
    V3 point_on_unit_sphere = sample_sphere(rand1, rand2);
    V3 point_on_sphere = (point_on_unit_sphere * radius) + center;
    V3 vec_from_lf_to_pos = point_on_sphere - look_from;
    shading_normal = point_on_unit_sphere;
    const float sphere_sa = (float)(4.f * VIMG_PI_D * (vdbl)radius * (vdbl)radius);
    V3 dir_to_surf = normalize(vec_from_lf_to_pos);
    float dist2 = length2(vec_from_lf_to_pos);
    float cosine = fabsf(dot(shading_normal, -dir_to_surf));
    float G = cosine / dist2;
    info.wi = dir_to_surf;
    info.pdf = 1.f / sphere_sa;
    info.dist = sqrtf(dist2);
    info.G = G;
  } else {
    // unqualified sqrt() on a float => double
    float cos_theta_max
        = (float)sqrt((vdbl)(1.0f - ((radius * radius) / length2(look_from - center))));
    V3 dir_center_to_lf = normalize(look_from - center);
    ONB onb = init_onb(dir_center_to_lf);
    V3 sample_z_dir = sample_sphere_cap(rand1, rand2, cos_theta_max);
    V3 sampled_point = normalize(xform_with_onb(onb, sample_z_dir)) * radius + center;
    float dist2 = length2(sampled_point - look_from);
    shading_normal = normalize(sampled_point - center);
    V3 sampled_dir = normalize(sampled_point - look_from);
    float cosine = fabsf(dot(shading_normal, -sampled_dir));
    float G = cosine / dist2;
    float pdf_solid_angle = (float)(1.0f / (2.f * VIMG_PI_D * (vdbl)(1.0f - cos_theta_max)));
    info.wi = sampled_dir;
    info.pdf = pdf_solid_angle * G;
    info.dist = sqrtf(dist2);
    info.G = G;
  }
  return mat_emitted(sc.materials[lr.material], info.wi, shading_normal);
}

// Sphere::surf_pdf (sphere.cpp:120-139)
VD float sphere_surf_pdf(const DScene& sc, uint32_t prim, V3 look_from, V3 point_on_light, V3 dir) {
  const vimg_sphere& s = sc.spheres[sc.prims[prim].a];
  const V3 center = mk3(s.center[0], s.center[1], s.center[2]);
  const float radius = s.radius;
  if (length2(look_from - center) <= radius * radius) {
    const float sphere_sa = (float)(4.f * VIMG_PI_D * (vdbl)radius * (vdbl)radius);
    return 1.f / sphere_sa;
  }
  float cos_theta_max = (float)sqrt((vdbl)(1.0f - ((radius * radius) / length2(look_from - center))));
  float pdf_solid_angle = (float)(1.0f / (2.f * VIMG_PI_D * (vdbl)(1.0f - cos_theta_max)));
  V3 shading_normal = normalize(point_on_light - center);
  float cosine = fabsf(dot(shading_normal, -dir));
  float dist2 = length2(point_on_light - look_from);
  return pdf_solid_angle * cosine / dist2;
}

// ConstBackground::sample (background.h:37-49) / EnvMap::sample (background.h:110-139)
VD V3 background_light_sample(const DScene& sc, Pcg32& rng, EmitterInfo& info) {
  float r1 = rand_float(rng);
  float r2 = rand_float(rng);
  if (VIMG_LEAN || sc.bg.type == VIMG_BG_CONST) {
    info.wi = sample_sphere(r1, r2);
    info.pdf = (float)(1.f / (4 * VIMG_PI_D));
    info.dist = CUDART_INF_F;
    info.G = 1.f;
    return mk3(sc.bg.color[0], sc.bg.color[1], sc.bg.color[2]);
  }
  // ArraySampling2D::sample (sampling.h:205-222)
  const uint32_t W = sc.bg.cdf_w, H = sc.bg.cdf_h;
  uint32_t row_index, column_index;
  float dv, du;
  sample_cdf_1d(sc.bg.row_cdf, H + 1, r1, row_index, dv);
  const float* row = sc.bg.col_cdfs + (size_t)row_index * (W + 1);
  sample_cdf_1d(row, W + 1, r2, column_index, du);
  const float u_env = ((float)column_index + du) / W;
  const float v_env = ((float)row_index + dv) / H;
  float pdf_y = sc.bg.row_cdf[row_index + 1] - sc.bg.row_cdf[row_index];
  float pdf_x = row[column_index + 1] - row[column_index];
  float choose_sample_pdf = pdf_y * pdf_x;

  const DTexture& img = sc.textures[sc.bg.image];
  float elevation = (float)((vdbl)v_env * VIMG_PI_D);
  float y = (float)d_cos((vdbl)v_env * VIMG_PI_D);  // std::cos(vdbl)
  const float azimuth = (float)((vdbl)(u_env * 2.f) * VIMG_PI_D);
  float x = m_sinf(azimuth) * m_sinf(elevation);
  float z = -1 * m_cosf(azimuth) * m_sinf(elevation);
  V3 wi = normalize(xform_dir(sc.bg.env_to_world, mk3(x, y, z)));
  float sin_elevation = m_sinf(elevation);
  float pdf = (float)((vdbl)(choose_sample_pdf * img.width * img.height)
                      / (2.f * VIMG_PI_D * VIMG_PI_D * (vdbl)sin_elevation));
  info.wi = wi;
  info.pdf = pdf;
  info.dist = CUDART_INF_F;
  info.G = 1.f;
  return col_at_uv_mipmap(sc, img, 0, mk2(u_env, v_env)) * sc.bg.radiance_scale;
}

// GroupOfEmitters::sample (emitters.h:39-56)
VD V3 lights_sample(const DScene& sc, V3 look_from, Pcg32& rng, EmitterInfo& info) {
  float rnd = rand_float(rng);
  float sx = rnd * sc.n_lights;
  const int index_obj = std_clampi((int)sx, 0, (int)sc.n_lights - 1);
  const float prob_obj = 1.f / sc.n_lights;
  const DLightRec lr = load_light_rec(sc, index_obj);  // one record: no prim -> mesh -> vertex walk
  V3 col;
  if (lr.type == VIMG_LIGHTREC_BACKGROUND)
    col = background_light_sample(sc, rng, info);
  else if (lr.type == VIMG_LIGHTREC_SPHERE)
    col = sphere_light_sample(sc, lr, look_from, rng, info);
  else
    col = triangle_light_sample(sc, lr, look_from, rng, info);
  info.pdf *= prob_obj;
  return col;
}

// Triangle::surf_pdf / Sphere::surf_pdf of the emitter a ray hit, from its leaf record (the
// same vertex / centre / radius values as the pools hold)
VD float emitter_surf_pdf_leaf(const DLeafPrim& lp, V3 look_from, V3 point_on_light, V3 dir) {
  if (lp.flags & VIMG_LEAF_SPHERE) {
    const V3 center = mk3(lp.a[0], lp.a[1], lp.a[2]);
    const float radius = lp.a[3];
    if (length2(look_from - center) <= radius * radius) {
      const float sphere_sa = (float)(4.f * VIMG_PI_D * (vdbl)radius * (vdbl)radius);
      return 1.f / sphere_sa;
    }
    float cos_theta_max = (float)sqrt((vdbl)(1.0f - ((radius * radius) / length2(look_from - center))));
    float pdf_solid_angle = (float)(1.0f / (2.f * VIMG_PI_D * (vdbl)(1.0f - cos_theta_max)));
    V3 shading_normal = normalize(point_on_light - center);
    float cosine = fabsf(dot(shading_normal, -dir));
    float dist2 = length2(point_on_light - look_from);
    return pdf_solid_angle * cosine / dist2;
  }
  V3 p0 = mk3(lp.a[0], lp.a[1], lp.a[2]), p1 = mk3(lp.a[3], lp.a[4], lp.a[5]), p2 = mk3(lp.a[6], lp.a[7], lp.a[8]);
  V3 edge1 = p1 - p0, edge2 = p2 - p0;
  float area = length(cross(edge2, edge1)) / 2.0f;
  return 1.f / area;
}

VD float emitter_surf_pdf(const DScene& sc, int32_t prim, V3 look_from, V3 look_at, V3 dir) {
  if (sc.prims[prim].type == VIMG_PRIM_SPHERE)
    return sphere_surf_pdf(sc, (uint32_t)prim, look_from, look_at, dir);
  return triangle_surf_pdf(sc, (uint32_t)prim);
}

// -------------------------------------------------------------- integrators ----
struct PathCounters {
  uint32_t rays_closest, rays_shadow;
  TravCounters trav;
};

// shading_normal_integrator / geometric_normal_integrator (normals.cpp:4-46)
template <bool GEOMETRIC, bool COUNT>
VD V3 normal_integrator(const DScene& sc, Ray& ray, PathCounters& pc) {
  HitInfo hit;
  pc.rays_closest++;
  if (scene_hit<COUNT>(sc, ray, hit, &pc.trav)) {
    V3 n = GEOMETRIC ? hit.hit_n_g : hit.hit_n_s;
    return (n + 1.0f) / 2.0f;
  }
  V3 unit_dir = normalize(ray.dir);
  float a = (float)(VR(0.5) * ((vdbl)unit_dir.y + VR(1.0)));  // double arithmetic (normals.cpp:19)
  return (1.0f - a) * mk3(1.0f, 1.0f, 1.0f) + a * mk3(0.5f, 0.7f, 1.0f);
}

VD bool russian_roulette(Pcg32& rng, float eta_scale, V3& throughput) {
  float rnd = (float)pcg32_random(rng) / (float)0xFFFFFFFFu;  // raw draw, can equal 1.0
  V3 rr = (1.f / eta_scale) * throughput;
  float max_val = std_min(std_max(std_max(rr.x, rr.y), rr.z), 0.95f);
  if (rnd > max_val) return false;
  throughput = throughput / max_val;
  return true;
}

VD float balance_heuristic(float pdf1, float pdf2) { return pdf1 / (pdf1 + pdf2); }

VD float geometric_term(V3 look_from, V3 point_on_surface, V3 surface_normal) {
  V3 dir_from_surf = look_from - point_on_surface;
  float distance2 = length2(dir_from_surf);
  dir_from_surf = normalize(dir_from_surf);
  float cosine = fabsf(dot(surface_normal, dir_from_surf));
  return cosine / distance2;
}

}  // namespace vimg
