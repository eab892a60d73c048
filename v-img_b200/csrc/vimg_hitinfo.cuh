// vimg_hitinfo.cuh -- ONB helpers, ray cones, textures and hit-info reconstruction.
#pragma once
#include "vimg_bvh.cuh"

namespace vimg {

// ------------------------------------------------ ONB (hit_utils.h:32-59) ----
VD V3 xform_with_onb(const ONB& onb, V3 vec) { return onb.u * vec.x + onb.v * vec.y + onb.w * vec.z; }
VD V3 project_onto_onb(const ONB& onb, V3 vec) {
  return mk3(dot(vec, onb.u), dot(vec, onb.v), dot(vec, onb.w));
}
VD V3 gram_schmidt(V3 v, V3 w) { return v - dot(v, w) * w; }
VD void get_axis(V3 n, V3& a, V3& b) {
  if (n.z < (-0.9999999f)) {
    a = mk3(0, -1, 0);
    b = mk3(-1, 0, 0);
  } else {
    float aa = 1.f / (1.f + n.z);
    float bb = -n.x * n.y * aa;
    a = mk3(1.f - n.x * n.x * aa, bb, -n.x);
    b = mk3(bb, 1 - n.y * n.y * aa, -n.y);
  }
}
VD ONB init_onb(V3 n) {
  ONB o;
  get_axis(n, o.u, o.v);
  o.w = n;
  return o;
}

// ------------------------------------------------- ray cones (ray.h:44-174) ----
VD float float_sign(float in) { return in > 0.f ? 1.f : -1.f; }

// spread_angle_from_curvature (ray.h:52-61): `1.0e-5` literals are double
VD float spread_angle_from_curvature(float mean_curvature, float cone_width, V3 ray_dir, V3 normal) {
  float dn = -dot(ray_dir, normal);
  dn = ((vdbl)fabsf(dn) < VR(1.0e-5)) ? (float)((vdbl)float_sign(dn) * VR(1.0e-5)) : dn;
  return mean_curvature * cone_width / dn;
}

VD RayCone propagate_reflect_cone(RayCone cone, float surface_spread_angle, float hit_dist) {
  RayCone r;
  r.cone_width = fabsf(cone.spread_angle * hit_dist + cone.cone_width);
  r.spread_angle = cone.spread_angle + surface_spread_angle;
  return r;
}

VD bool refract_with_TIR2D(V2 ray_dir, V2 normal, float eta, V2& out) {
  float NdotD = dot(normal, ray_dir);
  float k = 1.0f - eta * eta * (1.0f - NdotD * NdotD);
  if (k < 0.0f) return false;
  out = ray_dir * eta - normal * (eta * NdotD + sqrtf(k));
  return true;
}

// rotate2DPlusMinus (ray.h:81-92): unqualified cos/sin => double, narrowed to float
VD void rotate2DPlusMinus(V2 vec, float angle, V2& plus, V2& minus) {
  float c = (float)d_cos((vdbl)angle);
  float s = (float)d_sin((vdbl)angle);
  float cx = c * vec.x, sy = s * vec.y, sx = s * vec.x, cy = c * vec.y;
  plus = mk2(cx - sy, +sx + cy);
  minus = mk2(cx + sy, -sx + cy);
}

VD V2 orthogonal(V2 v) { return mk2(-v.y, v.x); }

// propagate_refract_cone (ray.h:95-174)
VD RayCone propagate_refract_cone(RayCone rayCone, V3 ray_in_dir, float surface_spread_angle,
                                  float eta, V3 refractedRayDir) {
  V3 tmp = eta * refractedRayDir + ray_in_dir;
  V3 normal = -tmp / length(tmp);
  V3 xAxis = normalize(ray_in_dir - normal * dot(normal, ray_in_dir));
  V3 yAxis = normal;
  V2 refractedDir2D = mk2(dot(refractedRayDir, xAxis), dot(refractedRayDir, yAxis));
  V2 incidentDir2D = mk2(dot(ray_in_dir, xAxis), dot(ray_in_dir, yAxis));
  V2 incidentDirOrtho2D = orthogonal(incidentDir2D);
  float widthSign = rayCone.cone_width > 0.0f ? 1.0f : -1.0f;
  V2 incidentDir2D_u, incidentDir2D_l;
  rotate2DPlusMinus(incidentDir2D, rayCone.spread_angle * widthSign * 0.5f, incidentDir2D_u,
                    incidentDir2D_l);
  V2 tu = incidentDirOrtho2D * rayCone.cone_width * 0.5f;
  V2 tl = -tu;
  float hitPoint_u_x = tu.x + incidentDir2D_u.x * (-tu.y / incidentDir2D_u.y);
  float hitPoint_l_x = tl.x + incidentDir2D_l.x * (-tl.y / incidentDir2D_l.y);
  float normalSign = hitPoint_u_x > hitPoint_l_x ? +1.0f : -1.0f;
  V2 normal2D = mk2(0.0f, 1.0f);
  V2 normal2D_u, normal2D_l;
  rotate2DPlusMinus(normal2D, -surface_spread_angle * normalSign * 0.5f, normal2D_u, normal2D_l);
  V2 ru, rl;
  if (!refract_with_TIR2D(incidentDir2D_u, normal2D_u, eta, ru)) {
    ru = incidentDir2D_u - normal2D_u * dot(normal2D_u, incidentDir2D_u);
    ru = normalize(ru);
  }
  if (!refract_with_TIR2D(incidentDir2D_l, normal2D_l, eta, rl)) {
    rl = incidentDir2D_l - normal2D_l * dot(normal2D_l, incidentDir2D_l);
    rl = normalize(rl);
  }
  float signA = (ru.x * rl.y - ru.y * rl.x) * normalSign < 0.0f ? +1.0f : -1.0f;
  float spreadAngle = m_acosf(dot(ru, rl)) * signA;
  if (is_nan(spreadAngle)) spreadAngle = 0.f;
  V2 refractDirOrtho2D = orthogonal(refractedDir2D);
  float width = (-hitPoint_u_x * ru.y) / dot(refractDirOrtho2D, orthogonal(ru));
  width += (hitPoint_l_x * rl.y) / dot(refractDirOrtho2D, orthogonal(rl));
  RayCone r;
  r.cone_width = width;
  r.spread_angle = spreadAngle;
  return r;
}

// ----------------------------------------------------------------- textures ----
// handle_wrapping (texture_common.h:22-53)
VD float handle_wrapping(float coord, uint32_t mode) {
  switch (mode) {
    case VIMG_WRAP_CLAMP_TO_EDGE:
      return std_clamp(coord, 0.f, 1.f);
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
    default:
      return std_clamp(coord, 0.f, 1.f);
  }
}

VD V3 load3(const float* p, uint32_t i) { return mk3(p[3 * (size_t)i], p[3 * (size_t)i + 1], p[3 * (size_t)i + 2]); }
VD V2 load2(const float* p, uint32_t i) { return mk2(p[2 * (size_t)i], p[2 * (size_t)i + 1]); }

// ImageTexture::col_at_uv_mipmap (image_texture.cpp:132-161): bilinear with truncation addressing
VD V3 col_at_uv_mipmap(const DScene& sc, const DTexture& t, int level, V2 uv) {
  uint32_t mip_w = max(t.width >> level, 1u), mip_h = max(t.height >> level, 1u);
  float pixel_u = handle_wrapping(uv.x, t.wrap_u) * mip_w;
  float pixel_v = handle_wrapping(uv.y, t.wrap_v) * mip_h;
  int curr_x = std_clampi((int)pixel_u, 0, (int)mip_w - 1);
  int curr_y = std_clampi((int)pixel_v, 0, (int)mip_h - 1);
  int next_x = std_clampi(curr_x + 1, 0, (int)mip_w - 1);
  int next_y = std_clampi(curr_y + 1, 0, (int)mip_h - 1);
  float x_fraction = pixel_u - curr_x;
  float y_fraction = pixel_v - curr_y;
  const float* lvl = sc.texels + 3 * (size_t)t.level_off[level];
  V3 x0 = load3(lvl, curr_x + curr_y * mip_w), x1 = load3(lvl, next_x + curr_y * mip_w);
  V3 a = mix(x0, x1, x_fraction);
  V3 y0 = load3(lvl, curr_x + next_y * mip_w), y1 = load3(lvl, next_x + next_y * mip_w);
  V3 b = mix(y0, y1, x_fraction);
  return mix(a, b, y_fraction);
}

// ImageTexture::col_mipmap_interpolate (image_texture.cpp:176-189)
VD V3 col_mipmap_interpolate(const DScene& sc, const DTexture& t, float lambda, V2 uv) {
  lambda = std_clamp(lambda, 0.f, (float)(t.n_levels - 1));
  int l0 = std_clampi((int)floorf(lambda), 0, (int)t.n_levels - 1);
  int l1 = std_clampi(l0 + 1, 0, (int)t.n_levels - 1);
  float fraction = lambda - floorf(lambda);
  V3 c0 = col_at_uv_mipmap(sc, t, l0, uv);
  V3 c1 = col_at_uv_mipmap(sc, t, l1, uv);
  return mix(c0, c1, fraction);
}

// ImageTexture::get_normal (image_texture.cpp:277-279)
VD V3 tex_get_normal(const DScene& sc, const DTexture& t, V2 uv) {
  return normalize(col_at_uv_mipmap(sc, t, 0, uv));
}

// TextureRGB::col_at_ray_hit for the three texture classes (texture_RGB.h:45-149,
// image_texture.cpp:163-174)
VD V3 col_at_ray_hit(const DScene& sc, int32_t tex, V3 ray_in_dir, RayCone cone, const HitInfo& hit) {
  const DTexture& t = sc.textures[tex];
  if (t.type == VIMG_TEX_CONST) return mk3(t.col_a[0], t.col_a[1], t.col_a[2]);
  if (VIMG_LEAN || t.type == VIMG_TEX_CHECKER) {
    // uint32_t u_board = std::floor(uv[0] * width): x86-64 converts float -> uint32 through a
    // 64-bit cvttss2si and truncates (negative values wrap); CUDA's direct cast saturates.
    uint32_t u_board = (uint32_t)(long long)floorf(hit.uv.x * t.width);
    uint32_t v_board = (uint32_t)(long long)floorf(hit.uv.y * t.height);
    if ((u_board + v_board) % 2 == 0) return mk3(t.col_a[0], t.col_a[1], t.col_a[2]);
    return mk3(t.col_b[0], t.col_b[1], t.col_b[2]);
  }
  // compute_texture_LOD (texture_RGB.h:138-149): std::log2(float) twice, then
  // unqualified log2(width*height) on an unsigned => double
  float lambda = 0.5f * m_log2f(hit.tex_coord_area / hit.primitive_area);
  lambda += m_log2f(fabsf(cone.cone_width) / fabsf(dot(ray_in_dir, hit.hit_n_g)));
  lambda = (float)((vdbl)lambda + (vdbl)t.half_log2_wh);
#ifdef VIMG_DEBUG_NEE
  printf("D   lod: tca=%.9g pa=%.9g cw=%.9g dot=%.9g lambda=%.9g\n", hit.tex_coord_area, hit.primitive_area, cone.cone_width, dot(ray_in_dir, hit.hit_n_g), lambda);
#endif
  if (is_nan(lambda)) lambda = 0.f;
  return col_mipmap_interpolate(sc, t, lambda - 2.f, hit.uv);
}

// TextureRG::get_at_uv (texture_RG.h:32-57) INCLUDING its row-stride quirk: the
// `next_x` taps index rows with `* height` instead of `* width`.
VD V2 texrg_get_at_uv(const DScene& sc, const DTextureRG& t, V2 uv) {
  float pixel_u = handle_wrapping(uv.x, t.wrap_u) * t.width;
  float pixel_v = handle_wrapping(uv.y, t.wrap_v) * t.height;
  int curr_x = std_clampi((int)pixel_u, 0, (int)t.width - 1);
  int curr_y = std_clampi((int)pixel_v, 0, (int)t.height - 1);
  int next_x = std_clampi(curr_x + 1, 0, (int)t.width - 1);
  int next_y = std_clampi(curr_y + 1, 0, (int)t.height - 1);
  float x_fraction = pixel_u - curr_x;
  float y_fraction = pixel_v - curr_y;
  const float* d = sc.rg_texels + 2 * (size_t)t.off;
  V2 x0 = load2(d, curr_x + curr_y * t.width), x1 = load2(d, next_x + curr_y * t.height);
  V2 a = mix(x0, x1, x_fraction);
  V2 y0 = load2(d, curr_x + next_y * t.width), y1 = load2(d, next_x + next_y * t.height);
  V2 b = mix(y0, y1, x_fraction);
  return mix(a, b, y_fraction);
}

// ------------------------------------------------------------------ hit info ----
// Triangle::hit_info (src/geometry/triangle.cpp:13-153)
// general case: walks the SoA pools (any combination of uv sets); out of line, rarely taken
__device__ __noinline__ HitInfo triangle_hit_info_general(const DScene& sc, const DLeafPrim& lp, const PreHit& pre) {
  const vimg_prim pr = sc.prims[lp.prim_id];
  const DMesh& mesh = sc.meshes[pr.a];
  const uint32_t* idx = sc.tri_indices + 3 * ((size_t)mesh.tri_base + pr.b);
  const uint32_t i0 = idx[0] + mesh.vert_base, i1 = idx[1] + mesh.vert_base, i2 = idx[2] + mesh.vert_base;
  V3 p0 = mk3(lp.a[0], lp.a[1], lp.a[2]), p1 = mk3(lp.a[3], lp.a[4], lp.a[5]),
     p2 = mk3(lp.a[6], lp.a[7], lp.a[8]);
  V3 edge1 = p1 - p0, edge2 = p2 - p0;
  float u = pre.e0 * pre.inv_det, v = pre.e1 * pre.inv_det, w = pre.e2 * pre.inv_det;
  const V3 tri_normal = normalize(cross(edge1, edge2));
  V3 n0, n1, n2, shading_normal;
  if (mesh.has_normals) {
    n0 = load3(sc.normals, i0);
    n1 = load3(sc.normals, i1);
    n2 = load3(sc.normals, i2);
    shading_normal = normalize(u * n0 + v * n1 + w * n2);
  } else {
    n0 = n1 = n2 = tri_normal;
    shading_normal = tri_normal;
  }
  const V3 hit_p = u * p0 + v * p1 + w * p2;

  V2 uv = mk2(u, v);
  V2 uv0 = mk2(0, 0), uv1 = mk2(1, 0), uv2 = mk2(1, 1);
  if (mesh.uv_color >= 0) {
    const float* tc = sc.uvs + 2 * (size_t)mesh.uv_color;
    uv0 = load2(tc, idx[0]);
    uv1 = load2(tc, idx[1]);
    uv2 = load2(tc, idx[2]);
    uv = u * uv0 + v * uv1 + w * uv2;
  }
  V2 metallic_roughness_uv = uv;
  if (mesh.uv_mr >= 0) {
    const float* tc = sc.uvs + 2 * (size_t)mesh.uv_mr;
    V2 a0 = load2(tc, idx[0]), a1 = load2(tc, idx[1]), a2 = load2(tc, idx[2]);
    metallic_roughness_uv = u * a0 + v * a1 + w * a2;
  }
  V2 duvds = uv2 - uv0;
  V2 duvdt = uv2 - uv1;
  float det = duvds.x * duvdt.y - duvdt.x * duvds.y;
  float dsdu = 0.f, dtdu = 0.f, dsdv = 0.f, dtdv = 0.f;
  V3 dpdu, dpdv;
  if (fabsf(det) > 1e-8f && !is_nan(det)) {
    dsdu = duvdt.y / det;
    dtdu = -duvds.y / det;
    dsdv = duvdt.x / det;
    dtdv = -duvds.x / det;
    V3 dpds = p2 - p0;
    V3 dpdt = p2 - p1;
    dpdu = dpds * dsdu + dpdt * dtdu;
    dpdv = dpds * dsdv + dpdt * dtdv;
  } else {
    get_axis(shading_normal, dpdu, dpdv);
  }

  const vimg_material& mat = sc.materials[mesh.material];
  if (!VIMG_LEAN && mat.normal_map >= 0) {
    V2 n_uv = mk2(u, v);
    if (mesh.uv_normal >= 0) {
      const float* tc = sc.uvs + 2 * (size_t)mesh.uv_normal;
      V2 a0 = load2(tc, idx[0]), a1 = load2(tc, idx[1]), a2 = load2(tc, idx[2]);
      n_uv = u * a0 + v * a1 + w * a2;
    }
    V3 n_tangent_space = tex_get_normal(sc, sc.textures[mat.normal_map], n_uv);
    ONB onb_n_map = init_onb(shading_normal);
    V3 local_space_normal = xform_with_onb(onb_n_map, n_tangent_space);
    float ulen = length(dpdu), vlen = length(dpdv);
    dpdu = normalize(gram_schmidt(dpdu, local_space_normal)) * ulen;
    dpdv = normalize(cross(local_space_normal, dpdu)) * vlen;
    shading_normal = local_space_normal;
  }

  V3 tangent = normalize(dpdu - shading_normal * dot(shading_normal, dpdu));
  V3 dnds = n2 - n0;
  V3 dndt = n2 - n1;
  V3 dndu = dnds * dsdu + dndt * dtdu;
  V3 dndv = dnds * dsdv + dndt * dtdv;
  V3 bitangent = normalize(cross(shading_normal, tangent));
  float mean_curvature = (dot(dndu, tangent) + dot(dndv, bitangent)) / 2.f;
  float twice_tri_area = length(cross(p1 - p0, p2 - p0));
  float uv_area = fabsf((uv1.x - uv0.x) * (uv2.y - uv0.y) - (uv2.x - uv0.x) * (uv1.y - uv0.y));

  HitInfo h;
  h.mat = mesh.material;
  h.prim = (int32_t)lp.prim_id;
  h.hit_p = hit_p;
  h.hit_n_s = shading_normal;
  h.hit_n_g = tri_normal;
  h.uv = uv;
  h.mr_uv = metallic_roughness_uv;
  h.frame.u = tangent;
  h.frame.v = bitangent;
  h.frame.w = shading_normal;
  h.primitive_area = twice_tri_area;
  h.tex_coord_area = uv_area;
  h.mean_curvature = mean_curvature;
  return h;
}

// one shading / light record with 16-byte loads
VD DShadeRec load_shade_rec(const DScene& sc, int32_t leaf_slot) {
  DShadeRec r;
  const uint4* src = reinterpret_cast<const uint4*>(sc.shade_recs + leaf_slot);
  uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int k = 0; k < 5; ++k) dst[k] = __ldg(src + k);
  return r;
}
VD DLightRec load_light_rec(const DScene& sc, int index) {
  DLightRec r;
  const uint4* src = reinterpret_cast<const uint4*>(sc.light_recs + index);
  uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int k = 0; k < 6; ++k) dst[k] = __ldg(src + k);
  return r;
}

// Triangle::hit_info (src/geometry/triangle.cpp:13-153) from the leaf record (vertices) and the
// shade record (vertex normals, colour uv set, material): the same values and the same
// operations as the general walk, without its dependent loads
// NEED_FRAME = false: the caller reads neither the tangent frame nor anything derived from the
// uv derivatives (Lambertian, Dielectric, emitters); they are then computed only if a normal map
// or the ray cones need them.
template <bool NEED_FRAME = true>
VD HitInfo triangle_hit_info(const DScene& sc, const DLeafPrim& lp, const DShadeRec& sr, const PreHit& pre) {
  if (!VIMG_LEAN && (sr.flags & VIMG_SHADE_GENERAL)) return triangle_hit_info_general(sc, lp, pre);
  V3 p0 = mk3(lp.a[0], lp.a[1], lp.a[2]), p1 = mk3(lp.a[3], lp.a[4], lp.a[5]),
     p2 = mk3(lp.a[6], lp.a[7], lp.a[8]);
  V3 edge1 = p1 - p0, edge2 = p2 - p0;
  float u = pre.e0 * pre.inv_det, v = pre.e1 * pre.inv_det, w = pre.e2 * pre.inv_det;
  const V3 tri_normal = normalize(cross(edge1, edge2));
  V3 n0, n1, n2, shading_normal;
  if (sr.flags & VIMG_SHADE_HAS_NORMALS) {
    n0 = mk3(sr.n[0], sr.n[1], sr.n[2]);
    n1 = mk3(sr.n[3], sr.n[4], sr.n[5]);
    n2 = mk3(sr.n[6], sr.n[7], sr.n[8]);
    shading_normal = normalize(u * n0 + v * n1 + w * n2);
  } else {
    n0 = n1 = n2 = tri_normal;
    shading_normal = tri_normal;
  }
  const V3 hit_p = u * p0 + v * p1 + w * p2;

  V2 uv = mk2(u, v);
  V2 uv0 = mk2(0, 0), uv1 = mk2(1, 0), uv2 = mk2(1, 1);
  if (sr.flags & VIMG_SHADE_HAS_UV) {
    uv0 = mk2(sr.uv[0], sr.uv[1]);
    uv1 = mk2(sr.uv[2], sr.uv[3]);
    uv2 = mk2(sr.uv[4], sr.uv[5]);
    uv = u * uv0 + v * uv1 + w * uv2;
  }
  const int32_t normal_map = VIMG_LEAN ? -1 : sc.materials[sr.material].normal_map;
  const bool derivs = NEED_FRAME || (!VIMG_LEAN && sc.uses_cones != 0) || normal_map >= 0;
  V2 duvds = uv2 - uv0;
  V2 duvdt = uv2 - uv1;
  float det = duvds.x * duvdt.y - duvdt.x * duvds.y;
  float dsdu = 0.f, dtdu = 0.f, dsdv = 0.f, dtdv = 0.f;
  V3 dpdu = mk3(0.f), dpdv = mk3(0.f);
  if (!derivs) {
    // nothing reads dpdu / dpdv
  } else if (fabsf(det) > 1e-8f && !is_nan(det)) {
    dsdu = duvdt.y / det;
    dtdu = -duvds.y / det;
    dsdv = duvdt.x / det;
    dtdv = -duvds.x / det;
    V3 dpds = p2 - p0;
    V3 dpdt = p2 - p1;
    dpdu = dpds * dsdu + dpdt * dtdu;
    dpdv = dpds * dsdv + dpdt * dtdv;
  } else {
    get_axis(shading_normal, dpdu, dpdv);
  }

  if (normal_map >= 0) {
    const V2 n_uv = (sr.flags & VIMG_SHADE_NORMAL_UV_SAME) ? uv : mk2(u, v);
    V3 n_tangent_space = tex_get_normal(sc, sc.textures[normal_map], n_uv);
    ONB onb_n_map = init_onb(shading_normal);
    V3 local_space_normal = xform_with_onb(onb_n_map, n_tangent_space);
    float ulen = length(dpdu), vlen = length(dpdv);
    dpdu = normalize(gram_schmidt(dpdu, local_space_normal)) * ulen;
    dpdv = normalize(cross(local_space_normal, dpdu)) * vlen;
    shading_normal = local_space_normal;
  }

  V3 tangent = mk3(0.f), bitangent = mk3(0.f);
  if (NEED_FRAME || (!VIMG_LEAN && sc.uses_cones != 0)) {
    tangent = normalize(dpdu - shading_normal * dot(shading_normal, dpdu));
    bitangent = normalize(cross(shading_normal, tangent));
  }
  // curvature and the two areas only feed ray cones / texture LOD (DScene::uses_cones)
  float mean_curvature = 0.f, twice_tri_area = 1.f, uv_area = 1.f;
  if (!VIMG_LEAN && sc.uses_cones) {
    V3 dnds = n2 - n0;
    V3 dndt = n2 - n1;
    V3 dndu = dnds * dsdu + dndt * dtdu;
    V3 dndv = dnds * dsdv + dndt * dtdv;
    mean_curvature = (dot(dndu, tangent) + dot(dndv, bitangent)) / 2.f;
    twice_tri_area = length(cross(p1 - p0, p2 - p0));
    uv_area = fabsf((uv1.x - uv0.x) * (uv2.y - uv0.y) - (uv2.x - uv0.x) * (uv1.y - uv0.y));
  }

  HitInfo h;
  h.mat = sr.material;
  h.prim = (int32_t)lp.prim_id;
  h.hit_p = hit_p;
  h.hit_n_s = shading_normal;
  h.hit_n_g = tri_normal;
  h.uv = uv;
  h.mr_uv = uv;  // no metallic-roughness set, or the colour set itself (else VIMG_SHADE_GENERAL)
  h.frame.u = tangent;
  h.frame.v = bitangent;
  h.frame.w = shading_normal;
  h.primitive_area = twice_tri_area;
  h.tex_coord_area = uv_area;
  h.mean_curvature = mean_curvature;
  return h;
}

// Sphere::hit_info (src/geometry/sphere.cpp:12-45).  dpdv is computed by the reference
// but never used; uv needs acos/atan2 and double pi arithmetic.
VD HitInfo sphere_hit_info(const DLeafPrim& lp, const DShadeRec& sr, const Ray& r) {
  V3 center = mk3(lp.a[0], lp.a[1], lp.a[2]);
  float radius = lp.a[3];
  const V3 hit_p = r.o + r.dir * r.maxT;
  const V3 normal = normalize(hit_p - center);
  // uv is only read by non-constant textures (col_at_ray_hit) and by the metallic-roughness
  // texture; for every other material the acos/atan2 pair is dead code in the reference too
  const bool needs_uv = (sr.flags & VIMG_SHADE_SPHERE_NEEDS_UV) != 0;
  float u = 0.f, v = 0.f;
  if (needs_uv) {
    float theta = m_acosf(-normal.y);
    float phi = (float)((vdbl)m_atan2f(-normal.z, normal.x) + VIMG_PI_D);
    u = (float)((vdbl)phi / (2.f * VIMG_PI_D));
    v = (float)((vdbl)theta / VIMG_PI_D);
  }
  V3 dpdu = mk3(-radius * normal.y, radius * normal.x, 0.f);
  V3 tangent = normalize(dpdu - normal * dot(normal, dpdu));
  HitInfo h;
  h.mat = sr.material;
  h.prim = (int32_t)lp.prim_id;
  h.hit_p = hit_p;
  h.hit_n_s = normal;
  h.hit_n_g = normal;
  h.uv = mk2(u, v);
  h.mr_uv = mk2(u, v);
  h.frame.u = tangent;
  h.frame.v = normalize(cross(normal, tangent));
  h.frame.w = normal;
  h.primitive_area = 1.f;
  h.tex_coord_area = 0.000001f;
  h.mean_curvature = 1.f / radius;
  return h;
}

// closest-hit query + hit_info: BVH::hit<std::optional<HitInfo>> as the integrators call it
template <bool COUNT>
VD bool scene_hit(const DScene& sc, Ray& ray, HitInfo& hit, TravCounters* cnt) {
  PreHit ph;
  if (!bvh_traverse<false, COUNT>(sc, ray, ph, cnt)) return false;
  const DLeafPrim lp = sc.leaf_prims[ph.leaf_slot];
  const DShadeRec sr = load_shade_rec(sc, ph.leaf_slot);
  if (lp.flags & VIMG_LEAF_SPHERE)
    hit = sphere_hit_info(lp, sr, ray);
  else
    hit = triangle_hit_info<true>(sc, lp, sr, ph);
  return true;
}

}  // namespace vimg
