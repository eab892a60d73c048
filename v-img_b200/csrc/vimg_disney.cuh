// vimg_disney.cuh -- Principled (Disney) BSDF: fused eval+pdf and the lobe samplers.
//
// Follows include/material/principled.h:100-205, src/material/principled.cpp:3-58 and
// include/material/disney_helpers/*.h.  float/double promotions of the reference
// (double literals `1.`, `2.`, std::numbers::pi, unqualified cos/sin/sqrt/fabs) are kept
// expression by expression so GPU and CPU stay on the same RNG trajectory.
#pragma once
#include "vimg_materials.cuh"

namespace vimg {

// G_w (disney_common.h:6-14): caret and the result run in double
VD float G_w(V3 w, float alphax, float alphay, const ONB& frame) {
  const V3 wl = project_onto_onb(frame, w);
  float vec_alpha
      = ((wl.x * alphax) * (wl.x * alphax) + (wl.y * alphay) * (wl.y * alphay)) / (wl.z * wl.z);
  float caret = (float)((sqrt(VR(1.) + (vdbl)vec_alpha) - VR(1.)) / VR(2.));
  return (float)(VR(1.) / (VR(1.) + (vdbl)caret));
}

// anisotropic_sample_visible_normals (disney_common.h:16-52)
VD V3 sample_visible_normals(V3 local_dir_in, float alphax, float alphay, Pcg32& rng) {
  float sign = 1.f;
  V3 top = local_dir_in;
  if (local_dir_in.z < 0.f) {
    sign = -1.f;
    top = -top;
  }
  V3 hemi_dir_in = normalize(mk3(alphax * top.x, alphay * top.y, top.z));
  float rand_x = rand_float(rng);
  float rand_y = rand_float(rng);
  float phi = (float)(2 * VIMG_PI_D * rand_x);
  float z = fmaf((1.0f - rand_y), (1.0f + hemi_dir_in.z), -hemi_dir_in.z);
  float sinTheta = sqrtf(std_clamp(1.0f - z * z, 0.0f, 1.0f));
  float x = (float)((vdbl)sinTheta * d_cos((vdbl)phi));
  float y = (float)((vdbl)sinTheta * d_sin((vdbl)phi));
  const V3 hemi_N = mk3(x, y, z) + hemi_dir_in;
  return sign * normalize(mk3(alphax * hemi_N.x, alphay * hemi_N.y, std_max(0.f, hemi_N.z)));
}

// fresnel_dielectric (disney_common.h:55-69)
VD float fresnel_dielectric(float n_dot_i, float eta) {
  float n_dot_t_sq = 1.f - (1.f - n_dot_i * n_dot_i) / (eta * eta);
  if (n_dot_t_sq < 0) return 1;
  float n_dot_t = sqrtf(n_dot_t_sq);
  n_dot_i = fabsf(n_dot_i);
  float rs = (n_dot_i - eta * n_dot_t) / (n_dot_i + eta * n_dot_t);
  float rp = (eta * n_dot_i - n_dot_t) / (eta * n_dot_i + n_dot_t);
  return (rs * rs + rp * rp) / 2;
}

VD float regularize_alpha(float a) {
  return a < MatConst::roughness_threshold
             ? std_clamp(2.f * a, MatConst::regularize_min, MatConst::regularize_max)
             : a;
}

// ------------------------------------------------------------------ diffuse ----
VD float FD(V3 n, V3 w, float FD_90) {
  return 1.f + (FD_90 - 1.f) * raise_to_power_5(1.f - std_max(dot(n, w), 0.f));
}

// eval_pdf_disney_diffuse (disney_diffuse.h:74-104)
VD void eval_pdf_disney_diffuse(V3 dir_in, V3 dir_out, const HitInfo& hit, V3 base_col,
                                float subsurface, float roughness, V3 half_vec, const ONB& frame,
                                V3& eval, float& pdf) {
  if (dot(hit.hit_n_g, dir_in) < 0 || dot(hit.hit_n_g, dir_out) < 0) {
    eval = mk3(0.f);
    pdf = 0.f;
    return;
  }
  float normal_dirout_dot = dot(frame.w, dir_out);
  const float cos_theta_out = std_max(normal_dirout_dot, 0.f);
  const float cos_theta_in = std_max(dot(frame.w, dir_in), 0.f);
  const float dot_h_out = std_max(dot(half_vec, dir_out), 0.f);
  const float FD_90 = (float)(VR(0.5) + VR(2.0) * (vdbl)roughness * (vdbl)dot_h_out * (vdbl)dot_h_out);
  const V3 base_diffuse = base_col * VIMG_INV_PI_F * FD(frame.w, dir_in, FD_90)
                          * FD(frame.w, dir_out, FD_90) * cos_theta_out;
  const float FSS_90 = roughness * dot_h_out * dot_h_out;
  V3 ss_diffuse = base_col * 1.25f * VIMG_INV_PI_F
                  * (FD(frame.w, dir_in, FSS_90) * FD(frame.w, dir_out, FSS_90)
                         * ((1.f / (cos_theta_out + cos_theta_in)) - 0.5f)
                     + 0.5f)
                  * cos_theta_out;
  eval = (1.f - subsurface) * base_diffuse + subsurface * ss_diffuse;
  pdf = std_max(normal_dirout_dot, 0.f) * VIMG_INV_PI_F;
}

// sample_disney_diffuse (disney_diffuse.h:52-72)
VD bool sample_disney_diffuse(V3 dir_in, const HitInfo& hit, const ONB& frame, Pcg32& rng,
                              ScatterInfo& out) {
  if (dot(hit.hit_n_g, dir_in) < 0) return false;
  float rand1 = rand_float(rng);
  float rand2 = rand_float(rng);
  V3 dir_out = xform_with_onb(frame, sample_hemisphere_cosine(rand1, rand2));
  if (dot(hit.hit_n_g, dir_out) <= 0) return false;
  out.wo = dir_out;
  out.eta = 0.f;
  out.is_specular = false;
  return true;
}

// -------------------------------------------------------------------- sheen ----
// eval_disney_sheen (disney_sheen.h:10-27)
VD V3 eval_disney_sheen(V3 dir_in, V3 dir_out, const HitInfo& hit, V3 base_col, float sheen_tint,
                        V3 half_vec, const ONB& frame) {
  if (dot(hit.hit_n_g, dir_in) < 0 || dot(hit.hit_n_g, dir_out) < 0) return mk3(0.f);
  const float base_lum = luminance(base_col);
  V3 C_tint = base_lum > 0 ? base_col / base_lum : mk3(1.f);
  V3 C_sheen = (mk3(1.f) - mk3(sheen_tint)) + sheen_tint * C_tint;
  return C_sheen * raise_to_power_5(1.f - std_max(dot(half_vec, dir_out), 0.f))
         * std_max(dot(frame.w, dir_out), 0.f);
}

// ---------------------------------------------------------------- clearcoat ----
// eval_pdf_disney_clearcoat (disney_clearcoat.h:106-139)
VD void eval_pdf_disney_clearcoat(V3 dir_in, V3 dir_out, const HitInfo& hit, float alpha_g,
                                  V3 half_vec, const ONB& frame, V3& eval, float& pdf) {
  if (dot(hit.hit_n_g, dir_in) < 0 || dot(hit.hit_n_g, dir_out) < 0) {
    eval = mk3(0.f);
    pdf = 0.f;
    return;
  }
  constexpr float R0 = ((1.5f - 1.f) * (1.5f - 1.f)) / ((1.5f + 1.f) * (1.5f + 1.f));
  float h_dirout_dot = fabsf(dot(half_vec, dir_out));
  float Fresenl = (float)((vdbl)R0 + (VR(1.) - (vdbl)R0) * (vdbl)raise_to_power_5(1.f - h_dirout_dot));
  float G = G_w(dir_in, 0.25f, 0.25f, frame) * G_w(dir_out, 0.25f, 0.25f, frame);
  const float alpha_g_square = alpha_g * alpha_g;
  const V3 local_H = project_onto_onb(frame, half_vec);
  float D = (float)((vdbl)(alpha_g_square - 1.f)
                    / (VIMG_PI_D * (vdbl)m_logf(alpha_g_square)
                       * (VR(1.) + ((vdbl)alpha_g_square - VR(1.)) * (vdbl)local_H.z * (vdbl)local_H.z)));
  float clearcoat_eval = (Fresenl * D * G) / (4.f * fabsf(dot(frame.w, dir_in)));
  pdf = (D * fabsf(dot(frame.w, half_vec))) / (4.f * h_dirout_dot);
  eval = mk3(clearcoat_eval);
}

// sample_local_h_clearcoat (disney_clearcoat.h:60-75): std::pow(float, double) => double pow
VD V3 sample_local_h_clearcoat(Pcg32& rng, float alpha) {
  float rand1 = rand_float(rng);
  float rand2 = rand_float(rng);
  float cos_square_elevation
      = (float)((1.f - d_pow((vdbl)(alpha * alpha), VR(1.) - (vdbl)rand1)) / (vdbl)(1.f - (alpha * alpha)));
  float cos_elevation = sqrtf(cos_square_elevation);
  float sin_elevation = sqrtf(1 - cos_square_elevation);
  float h_azimuth = (float)(2.f * VIMG_PI_D * rand2);
  float x = sin_elevation * m_cosf(h_azimuth);
  float y = sin_elevation * m_sinf(h_azimuth);
  return mk3(x, y, cos_elevation);
}

// sample_disney_clearcoat (disney_clearcoat.h:77-104)
VD bool sample_disney_clearcoat(V3 dir_in, const HitInfo& hit, ONB frame, float clearcoat_gloss,
                                Pcg32& rng, bool regularize, ScatterInfo& out) {
  if (dot(hit.hit_n_g, dir_in) < 0) return false;
  float alpha_g = (1.f - clearcoat_gloss) * 0.1f + clearcoat_gloss * 0.001f;
  if (regularize && alpha_g < MatConst::roughness_threshold)
    alpha_g = std_clamp(2.f * alpha_g, MatConst::regularize_min, MatConst::regularize_max);
  V3 local_h = sample_local_h_clearcoat(rng, alpha_g);
  if (dot(frame.w, dir_in) < 0) {
    frame.u = -frame.u;
    frame.v = -frame.v;
    frame.w = -frame.w;
  }
  const V3 H = normalize(xform_with_onb(frame, local_h));
  V3 reflected = normalize(-dir_in + 2 * dot(dir_in, H) * H);
  if (dot(hit.hit_n_g, reflected) <= 0) return false;
  out.wo = reflected;
  out.eta = 0;
  out.is_specular = true;
  return true;
}

// -------------------------------------------------------------------- metal ----
VD float ggx_D(V3 local_H, float alphax, float alphay) {
  float h = (local_H.x * local_H.x) / (alphax * alphax) + (local_H.y * local_H.y) / (alphay * alphay)
            + (local_H.z * local_H.z);
  return (float)(VR(1.) / (VIMG_PI_D * (vdbl)alphax * (vdbl)alphay * (vdbl)(h * h)));
}

// eval_pdf_disney_metal (disney_metal.h:120-152)
VD void eval_pdf_disney_metal(V3 dir_in, V3 dir_out, const HitInfo& hit, V3 base_col,
                              float spec_tint, float specular, float eta, float metallic,
                              V3 half_vec, const ONB& frame, float G, float G_in, float alphax,
                              float alphay, V3& eval, float& pdf) {
  if (dot(hit.hit_n_g, dir_in) < 0 || dot(hit.hit_n_g, dir_out) < 0) {
    eval = mk3(0.f);
    pdf = 0.f;
    return;
  }
  float base_lum = luminance(base_col);
  V3 C_tint = base_lum > 0 ? base_col / base_lum : mk3(1.f);
  V3 K_s = (mk3(1.f) - mk3(spec_tint)) + spec_tint * C_tint;
  float R0 = ((eta - 1.f) * (eta - 1.f)) / ((eta + 1.f) * (eta + 1.f));
  V3 C_0 = (specular * R0 * (1.f - metallic)) * K_s + metallic * base_col;
  V3 Fresnel = C_0 + (mk3(1.f) - C_0) * raise_to_power_5(1.f - dot(half_vec, dir_out));
  const V3 local_H = project_onto_onb(frame, half_vec);
  float D = ggx_D(local_H, alphax, alphay);
  float D_mul_denominator = D / (4.f * fabsf(dot(frame.w, dir_in)));
  eval = Fresnel * G * D_mul_denominator;
  pdf = G_in * D_mul_denominator;
}

// alphas as the SAMPLERS compute them (roughness not clamped for metal, clamped for glass)
VD void sampler_alphas(float roughness, float anisotropic, bool clamp_roughness, bool regularize,
                       float& alphax, float& alphay) {
  constexpr float alpha_min = 0.0001;
  float aspect = sqrtf(1.f - 0.9f * anisotropic);
  if (clamp_roughness) roughness = std_clamp(roughness, 0.01f, 1.f);
  float roughness_square = roughness * roughness;
  alphax = std_max(alpha_min, roughness_square / aspect);
  alphay = std_max(alpha_min, roughness_square * aspect);
  if (regularize) {
    alphax = regularize_alpha(alphax);
    alphay = regularize_alpha(alphay);
  }
}

// sample_disney_metal (disney_metal.h:78-118)
VD bool sample_disney_metal(V3 dir_in, const HitInfo& hit, float roughness, float anisotropic,
                            const ONB& frame, Pcg32& rng, bool regularize, ScatterInfo& out) {
  if (dot(hit.hit_n_g, dir_in) < 0) return false;
  V3 local_dir_in = project_onto_onb(frame, dir_in);
  float alphax, alphay;
  sampler_alphas(roughness, anisotropic, false, regularize, alphax, alphay);
  V3 local_micro_normal = sample_visible_normals(local_dir_in, alphax, alphay, rng);
  V3 half_vector = normalize(xform_with_onb(frame, local_micro_normal));
  V3 reflected = normalize(-dir_in + 2 * dot(dir_in, half_vector) * half_vector);
  if (dot(reflected, hit.hit_n_g) <= 0) return false;
  out.wo = reflected;
  out.eta = 0.f;
  out.is_specular = true;
  return true;
}

// -------------------------------------------------------------------- glass ----
// eval_pdf_disney_rough_glass (disney_glass.h:181-234)
VD void eval_pdf_disney_rough_glass(V3 dir_in, V3 dir_out, const HitInfo& hit, V3 base_col,
                                    float mat_eta, V3 half_vec, const ONB& frame, float G,
                                    float G_in, float alphax, float alphay, V3& eval, float& pdf) {
  float in_geo_dot = dot(dir_in, hit.hit_n_g);
  bool reflect = (in_geo_dot * dot(hit.hit_n_g, dir_out)) >= 0;
  float eta = in_geo_dot >= 0 ? mat_eta : 1.f / mat_eta;
  if (!reflect) half_vec = normalize(dir_in + dir_out * eta);
  float h_dot_in = dot(half_vec, dir_in);
  float F = fresnel_dielectric(h_dot_in, eta);
  const V3 local_H = project_onto_onb(frame, half_vec);
  float D = ggx_D(local_H, alphax, alphay);
  float normal_in_dot = dot(frame.w, dir_in);
  if (reflect) {
    eval = base_col * (F * D * G) / (4.f * fabsf(normal_in_dot));
    pdf = (F * D * G_in) / (4.f * fabsf(normal_in_dot));
  } else {
    float eta_factor = 1.f / (eta * eta);
    float h_dot_out = dot(half_vec, dir_out);
    float sqrt_denom = h_dot_in + eta * h_dot_out;
    eval = mk3(sqrtf(base_col.x), sqrtf(base_col.y), sqrtf(base_col.z))
           * (eta_factor * (1 - F) * D * G * eta * eta * fabsf(h_dot_out * h_dot_in))
           / (fabsf(normal_in_dot) * sqrt_denom * sqrt_denom);
    float dh_dout = eta * eta * h_dot_out / (sqrt_denom * sqrt_denom);
    pdf = (1.f - F) * D * G_in * fabsf(dh_dout * h_dot_in / normal_in_dot);
  }
}

// sample_disney_rough_glass (disney_glass.h:108-179)
VD bool sample_disney_rough_glass(V3 dir_in, const HitInfo& hit, float mat_eta, float anisotropic,
                                  float roughness, const ONB& frame, Pcg32& rng, bool regularize,
                                  ScatterInfo& out) {
  float in_geo_dot = dot(dir_in, hit.hit_n_g);
  float eta = in_geo_dot >= 0 ? mat_eta : 1.f / mat_eta;
  float alphax, alphay;
  sampler_alphas(roughness, anisotropic, true, regularize, alphax, alphay);
  V3 local_dir_in = project_onto_onb(frame, dir_in);
  V3 local_micro_normal = sample_visible_normals(local_dir_in, alphax, alphay, rng);
  V3 half_vec = xform_with_onb(frame, local_micro_normal);
  float h_dot_in = dot(half_vec, dir_in);
  float F = fresnel_dielectric(h_dot_in, eta);
  float rnd = rand_float(rng);
  if (rnd <= F) {
    V3 reflected = normalize(-dir_in + 2 * dot(dir_in, half_vec) * half_vec);
    if (dot(reflected, hit.hit_n_g) * dot(dir_in, hit.hit_n_g) <= 0) return false;
    out.wo = reflected;
    out.eta = 0.f;
    out.is_specular = true;
    return true;
  }
  float h_dot_out_sq = 1 - (1 - h_dot_in * h_dot_in) / (eta * eta);
  if (h_dot_out_sq <= 0) return false;
  if (h_dot_in < 0) half_vec = -half_vec;
  float h_dot_out = (float)sqrt((vdbl)h_dot_out_sq);
  V3 refracted = -dir_in / eta + (fabsf(h_dot_in) / eta - h_dot_out) * half_vec;
  if (dot(refracted, hit.hit_n_g) * dot(dir_in, hit.hit_n_g) >= 0) return false;
  V3 generalized_h = normalize(dir_in + refracted * eta);
  float g_h_dot_in = dot(generalized_h, dir_in);
  if ((1 - (1 - g_h_dot_in * g_h_dot_in) / (eta * eta)) <= 0) return false;
  out.wo = refracted;
  out.eta = eta;
  out.is_specular = true;
  return true;
}

// --------------------------------------------------------------- Principled ----
VD void principled_frame_mr(const DScene& sc, const vimg_material& m, V3 dir_in, const HitInfo& hit,
                            ONB& frame, float& metallic, float& roughness) {
  frame = hit.frame;
  if ((dot(hit.hit_n_s, dir_in) * dot(hit.hit_n_g, dir_in)) < 0) {
    frame.u = -frame.u;
    frame.v = -frame.v;
    frame.w = -frame.w;
  }
  V2 m_r = mk2(1.f, 1.f);
  if (!VIMG_LEAN && m.texture_rg >= 0) m_r = texrg_get_at_uv(sc, sc.textures_rg[m.texture_rg], hit.mr_uv);
  m_r = m_r * mk2(m.metallic_roughness_factor[0], m.metallic_roughness_factor[1]);
  metallic = m_r.x;
  roughness = m_r.y;
}

// Principled::eval_pdf<T> (principled.h:100-205).  Returns (eval, pdf); the ReturnEval
// instantiation (eval_div_pdf) divides at the same two places the reference does.
VD void principled_eval_pdf(const DScene& sc, const vimg_material& m, V3 wi, V3 wo,
                            const HitInfo& hit, RayCone cone, bool regularize, V3& eval,
                            float& pdf) {
  V3 dir_in = -wi;
  ONB frame;
  float metallic, roughness;
  principled_frame_mr(sc, m, dir_in, hit, frame, metallic, roughness);
  V3 base_color = col_at_ray_hit(sc, m.texture, wi, cone, hit);
  V3 half_vector = normalize(dir_in + wo);
#ifdef VIMG_DEBUG_NEE
  printf("D   principled: base=%.9g %.9g %.9g metallic=%.9g rough=%.9g uv=%.9g %.9g mruv=%.9g %.9g\n", base_color.x, base_color.y, base_color.z, metallic, roughness,
         hit.uv.x, hit.uv.y, hit.mr_uv.x, hit.mr_uv.y);
#endif

  constexpr float alpha_min = 0.0001;
  float aspect = sqrtf(1.f - 0.9f * m.anisotropic);
  float roughness_clamp = std_clamp(roughness, 0.01f, 1.f);
  float roughness_square = roughness_clamp * roughness_clamp;
  float alphax = std_max(alpha_min, roughness_square / aspect);
  float alphay = std_max(alpha_min, roughness_square * aspect);
  if (regularize) {
    alphax = regularize_alpha(alphax);
    alphay = regularize_alpha(alphay);
  }
  float G_in = G_w(dir_in, alphax, alphay, frame);
  float G = G_in * G_w(wo, alphax, alphay, frame);

  V3 eval_glass;
  float pdf_glass;
  eval_pdf_disney_rough_glass(dir_in, wo, hit, base_color, m.eta, half_vector, frame, G, G_in,
                              alphax, alphay, eval_glass, pdf_glass);
  const float st = m.specular_transmission;
  if (dot(hit.hit_n_g, dir_in) < 0) {  // only glass if light is under the surface
    eval = (1.f - metallic) * st * eval_glass;
    pdf = pdf_glass;
    return;
  }
  V3 eval_sheen = eval_disney_sheen(dir_in, wo, hit, base_color, m.sheen_tint, half_vector, frame);
  V3 eval_diff;
  float pdf_diff;
  eval_pdf_disney_diffuse(dir_in, wo, hit, base_color, m.subsurface, roughness, half_vector, frame,
                          eval_diff, pdf_diff);
  float alpha_g = (1.f - m.clearcoat_gloss) * 0.1f + m.clearcoat_gloss * 0.001f;
  alpha_g = (regularize && (alpha_g < MatConst::roughness_threshold))
                ? std_clamp(2.f * alpha_g, MatConst::regularize_min, MatConst::regularize_max)
                : alpha_g;
  V3 eval_clearcoat;
  float pdf_clearcoat;
  eval_pdf_disney_clearcoat(dir_in, wo, hit, alpha_g, half_vector, frame, eval_clearcoat,
                            pdf_clearcoat);
  V3 eval_metal;
  float pdf_metal;
  eval_pdf_disney_metal(dir_in, wo, hit, base_color, m.specular_tint, m.specular, m.eta, metallic,
                        half_vector, frame, G, G_in, alphax, alphay, eval_metal, pdf_metal);

  eval = ((1.f - st) * (1.f - metallic) * eval_diff) + ((1.f - metallic) * m.sheen * eval_sheen)
         + (0.25f * m.clearcoat * eval_clearcoat) + ((1.f - st * (1.f - metallic)) * eval_metal)
         + ((1.f - metallic) * st * eval_glass);

  float diffuse_weight = (1.f - metallic) * (1.f - st);
  float clearcoat_weight = 0.25f * m.clearcoat;
  float metal_weight = (1.f - st * (1.f - metallic));
  float glass_weight = (1.f - metallic) * st;
  float total_w = diffuse_weight + clearcoat_weight + metal_weight + glass_weight;
  float choose_diff = diffuse_weight / total_w;
  float choose_clearcoat = clearcoat_weight / total_w;
  float choose_metal = metal_weight / total_w;
  float choose_glass = glass_weight / total_w;
  pdf = choose_diff * pdf_diff + choose_clearcoat * pdf_clearcoat + choose_metal * pdf_metal
        + choose_glass * pdf_glass;
#ifdef VIMG_DEBUG_NEE
  printf("D   parts: diff=%.9g %.9g %.9g sheen=%.9g %.9g %.9g cc=%.9g %.9g %.9g metal=%.9g %.9g %.9g glass=%.9g %.9g %.9g eval=%.9g %.9g %.9g pdf=%.9g\n", eval_diff.x, eval_diff.y, eval_diff.z,
         eval_sheen.x, eval_sheen.y, eval_sheen.z, eval_clearcoat.x, eval_clearcoat.y, eval_clearcoat.z, eval_metal.x, eval_metal.y, eval_metal.z,
         eval_glass.x, eval_glass.y, eval_glass.z, eval.x, eval.y, eval.z, pdf);
#endif
}

// Principled::sample_mat (principled.cpp:3-58)
VD bool principled_sample(const DScene& sc, const vimg_material& m, V3 wi, const HitInfo& hit,
                          Pcg32& rng, bool regularize, ScatterInfo& out) {
  V3 dir_in = -wi;
  ONB frame;
  float metallic, roughness;
  principled_frame_mr(sc, m, dir_in, hit, frame, metallic, roughness);
  if (dot(hit.hit_n_g, dir_in) < 0)
    return sample_disney_rough_glass(dir_in, hit, m.eta, m.anisotropic, roughness, frame, rng,
                                     regularize, out);
  const float st = m.specular_transmission;
  float diffuse_weight = (1.f - metallic) * (1.f - st);
  float clearcoat_weight = 0.25f * m.clearcoat;
  float metal_weight = (1.f - st * (1.f - metallic));
  float glass_weight = (1.f - metallic) * st;
  float total_w = diffuse_weight + clearcoat_weight + metal_weight + glass_weight;
  float choose_diff = diffuse_weight / total_w;
  float choose_clearcoat = clearcoat_weight / total_w;
  float choose_metal = metal_weight / total_w;
  float choose_glass = glass_weight / total_w;
  float rnd = rand_float(rng);
  if (rnd <= choose_diff) {
    return sample_disney_diffuse(dir_in, hit, frame, rng, out);
  } else if (rnd > choose_diff && rnd <= (choose_diff + choose_clearcoat)) {
    return sample_disney_clearcoat(dir_in, hit, frame, m.clearcoat_gloss, rng, regularize, out);
  } else if (rnd > (choose_diff + choose_clearcoat)
             && rnd <= (choose_diff + choose_clearcoat + choose_metal)) {
    return sample_disney_metal(dir_in, hit, roughness, m.anisotropic, frame, rng, regularize, out);
  } else if (rnd > (choose_diff + choose_clearcoat + choose_metal)
             && rnd <= (choose_diff + choose_clearcoat + choose_metal + choose_glass)) {
    return sample_disney_rough_glass(dir_in, hit, m.eta, m.anisotropic, roughness, frame, rng,
                                     regularize, out);
  }
  return false;
}

}  // namespace vimg
