// vimg_materials.cuh -- Material interface as type-tagged dispatch.
//
// Only the virtuals the integrators call exist here: sample_mat,
// eval_div_pdf, eval_pdf_pair, emitted, is_emissive, is_delta.  Base-class defaults
// (include/material/material.h:25-72) are reproduced, including the quirk that
// Dielectric inherits eval_pdf_pair = (0, 1).
#pragma once
#include "vimg_hitinfo.cuh"

namespace vimg {

VD float raise_to_power_5(float base) { return base * base * base * base * base; }
VD float luminance(V3 v) { return dot(v, mk3(0.212671f, 0.715160f, 0.072169f)); }

namespace MatConst {
constexpr float regularize_min = 0.03f;
constexpr float regularize_max = 0.1f;
constexpr float roughness_threshold = 0.1f;
}  // namespace MatConst

VD bool mat_is_emissive(const vimg_material& m) { return m.type == VIMG_MAT_DIFFUSE_LIGHT; }
VD bool mat_is_delta(const vimg_material& m) {
  // Lambertian, DiffuseLight, Principled override to false; Dielectric and base are true
  return m.type == VIMG_MAT_DIELECTRIC || m.type == VIMG_MAT_BASE;
}

// DiffuseLight::emitted (diffuse_light.h:30-38); every other material returns 0
VD V3 mat_emitted(const vimg_material& m, V3 ray_dir, V3 shading_normal) {
  if (m.type != VIMG_MAT_DIFFUSE_LIGHT) return mk3(0.f);
  bool front_face = dot(shading_normal, ray_dir) < 0;
  return front_face ? mk3(m.emit[0], m.emit[1], m.emit[2]) : mk3(0.f);
}

// ------------------------------------------------------------- Lambertian ----
// Lambertian::sample_mat (lambertian.cpp:5-26): both draws happen before the back-face test
VD bool lambertian_sample(V3 wi, const HitInfo& hit, Pcg32& rng, ScatterInfo& out) {
  float rand1 = rand_float(rng);
  float rand2 = rand_float(rng);
  bool front_face = dot(wi, hit.hit_n_s) < 0;
  V3 shading_normal = front_face ? hit.hit_n_s : -hit.hit_n_s;
  ONB onb = init_onb(shading_normal);
  V3 dir = xform_with_onb(onb, sample_hemisphere_cosine(rand1, rand2));
  if (!front_face) return false;
  out.wo = dir;
  out.eta = 0.f;
  out.is_specular = false;
  return true;
}

// Lambertian::eval_pdf_pair (lambertian.cpp:46-54): max(0, wo.n_s) / pi in double, narrowed
VD void lambertian_eval_pdf(const DScene& sc, const vimg_material& m, V3 wi, V3 wo,
                            const HitInfo& hit, RayCone cone, V3& eval, float& pdf) {
  float dot_product = (float)((vdbl)std_max(0.0f, dot(wo, hit.hit_n_s)) / VIMG_PI_D);
  eval = col_at_ray_hit(sc, m.texture, wi, cone, hit) * dot_product;
  pdf = dot_product;
}

// ------------------------------------------------------------- Dielectric ----
VD V3 reflect_dir(V3 wi, V3 n) { return wi - (2.f * dot(wi, n) * n); }
VD float schlick_apprx(float cosine, float in_ior, float out_ior) {
  float r0 = (in_ior - out_ior) / (in_ior + out_ior);
  r0 = r0 * r0;
  return r0 + (1.f - r0) * raise_to_power_5(1.f - cosine);
}
VD V3 refract_dir(V3 wi, V3 n, float ratio, float cos_thetaI, float sin_thetaT_square) {
  const float normal_mul = (ratio * cos_thetaI) - sqrtf(1.f - sin_thetaT_square);
  return (ratio * wi) + (normal_mul * n);
}

// Dielectric::sample_mat (dielectric.cpp:27-73): eta != 0 even when reflecting
VD bool dielectric_sample(const vimg_material& m, V3 wi, const HitInfo& hit, Pcg32& rng,
                          ScatterInfo& out) {
  const float ior = m.ior;
  V3 wo;
  float eta;
  bool front_face = dot(wi, hit.hit_n_s) < 0;
  V3 shading_normal = front_face ? hit.hit_n_s : -hit.hit_n_s;
  const float cos_thetaI = -1.f * dot(wi, shading_normal);
  float randf = rand_float(rng);
  if (front_face) {
    eta = ior;
    const float schlick = schlick_apprx(cos_thetaI, 1.0f, ior);
    if (schlick > randf) {
      wo = reflect_dir(wi, shading_normal);
    } else {
      const float ratio = 1.f / ior;
      const float sin_thetaT_square = (ratio * ratio) * (1.f - (cos_thetaI * cos_thetaI));
      wo = refract_dir(wi, shading_normal, ratio, cos_thetaI, sin_thetaT_square);
    }
  } else {
    eta = 1.f / ior;
    const float ratio = ior;
    const float sin_thetaT_square = (ratio * ratio) * (1.f - (cos_thetaI * cos_thetaI));
    if ((sin_thetaT_square > 1.f)
        || (schlick_apprx(sqrtf(1.f - sin_thetaT_square), ior, 1.f) > randf)) {
      wo = reflect_dir(wi, shading_normal);
    } else {
      wo = refract_dir(wi, shading_normal, ratio, cos_thetaI, sin_thetaT_square);
    }
  }
  out.wo = wo;
  out.eta = eta;
  out.is_specular = true;
  return true;
}

}  // namespace vimg
