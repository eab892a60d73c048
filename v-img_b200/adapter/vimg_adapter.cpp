// vimg_adapter.cpp -- see vimg_adapter.h.  Compile with -fno-access-control.
#include "vimg_adapter.h"

#include <geometry/sphere.h>
#include <geometry/triangle.h>
#include <material/dielectric.h>
#include <material/diffuse_light.h>
#include <material/lambertian.h>
#include <material/principled.h>

#include <cstring>
#include <stdexcept>

namespace {

uint32_t wrap_code(TextureWrappingMode m) {
  switch (m) {
    case TextureWrappingMode::ClampToEdge: return VIMG_WRAP_CLAMP_TO_EDGE;
    case TextureWrappingMode::MirroredRepeat: return VIMG_WRAP_MIRRORED_REPEAT;
    case TextureWrappingMode::Repeat: return VIMG_WRAP_REPEAT;
  }
  return VIMG_WRAP_CLAMP_TO_EDGE;
}

void copy_mat4(const glm::mat4& m, float out[16]) {
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) out[c * 4 + r] = m[c][r];
}

// Interns scene objects by address so shared materials/textures/meshes keep
// one index, like the raw aliases the reference holds (Mesh::mat, Sphere::mat,
// Material::normal_map, Principled::tex ...).
struct Interner {
  VimgFlatScene& fs;
  std::unordered_map<const void*, int32_t> tex_ids, texrg_ids, mat_ids, mesh_ids;

  int32_t texture(const TextureRGB* t) {
    if (!t) return VIMG_NO_TEXTURE;
    auto it = tex_ids.find(t);
    if (it != tex_ids.end()) return it->second;
    vimg_texture out{};
    if (auto* c = dynamic_cast<const ConstColor*>(t)) {
      out.type = VIMG_TEX_CONST;
      std::memcpy(out.col_a, &c->albedo, 12);
    } else if (auto* cb = dynamic_cast<const Checkerboard*>(t)) {
      out.type = VIMG_TEX_CHECKER;
      out.width = cb->width;
      out.height = cb->height;
      std::memcpy(out.col_a, &cb->col_a, 12);
      std::memcpy(out.col_b, &cb->col_b, 12);
    } else if (auto* img = dynamic_cast<const ImageTexture*>(t)) {
      out.type = VIMG_TEX_IMAGE;
      out.width = img->width;
      out.height = img->height;
      out.n_levels = static_cast<uint32_t>(img->mipmap.size());
      out.wrap_u = wrap_code(img->u_wrapping_mode);
      out.wrap_v = wrap_code(img->v_wrapping_mode);
      fs.tex_level_ptrs.emplace_back();
      for (const auto& lvl : img->mipmap)
        fs.tex_level_ptrs.back().push_back(reinterpret_cast<const float*>(lvl.data()));
      // `levels` is patched after all vectors stopped growing (finalise()).
      out.levels = reinterpret_cast<const float* const*>(fs.tex_level_ptrs.size());
    } else {
      throw std::runtime_error("vimg adapter: unknown TextureRGB subclass");
    }
    int32_t id = static_cast<int32_t>(fs.textures.size());
    fs.textures.push_back(out);
    tex_ids[t] = id;
    return id;
  }

  int32_t texture_rg(const TextureRG* t) {
    if (!t) return VIMG_NO_TEXTURE;
    auto it = texrg_ids.find(t);
    if (it != texrg_ids.end()) return it->second;
    vimg_texture_rg out{};
    out.width = t->width;
    out.height = t->height;
    out.wrap_u = wrap_code(t->u_wrapping_mode);
    out.wrap_v = wrap_code(t->v_wrapping_mode);
    out.data = reinterpret_cast<const float*>(t->roughness_metallic_vec.data());
    int32_t id = static_cast<int32_t>(fs.textures_rg.size());
    fs.textures_rg.push_back(out);
    texrg_ids[t] = id;
    return id;
  }

  uint32_t material(const Material* m) {
    if (!m) throw std::runtime_error("vimg adapter: null material");
    auto it = mat_ids.find(m);
    if (it != mat_ids.end()) return static_cast<uint32_t>(it->second);
    vimg_material out{};
    out.texture = out.texture_rg = VIMG_NO_TEXTURE;
    out.normal_map = texture(m->normal_map);
    if (auto* l = dynamic_cast<const Lambertian*>(m)) {
      out.type = VIMG_MAT_LAMBERTIAN;
      out.texture = texture(l->tex);
    } else if (auto* d = dynamic_cast<const DiffuseLight*>(m)) {
      out.type = VIMG_MAT_DIFFUSE_LIGHT;
      std::memcpy(out.emit, &d->emit, 12);
    } else if (auto* g = dynamic_cast<const Dielectric*>(m)) {
      out.type = VIMG_MAT_DIELECTRIC;
      out.ior = g->ior;
    } else if (auto* p = dynamic_cast<const Principled*>(m)) {
      out.type = VIMG_MAT_PRINCIPLED;
      out.texture = texture(p->tex);
      out.texture_rg = texture_rg(p->metallic_roughness);
      out.metallic_roughness_factor[0] = p->metallic_roughness_factor[0];
      out.metallic_roughness_factor[1] = p->metallic_roughness_factor[1];
      out.specular_transmission = p->specular_transmission;
      out.subsurface = p->subsurface;
      out.specular = p->specular;
      out.specular_tint = p->specular_tint;
      out.anisotropic = p->anisotropic;
      out.sheen = p->sheen;
      out.sheen_tint = p->sheen_tint;
      out.clearcoat = p->clearcoat;
      out.clearcoat_gloss = p->clearcoat_gloss;
      out.eta = p->eta;
    } else {
      out.type = VIMG_MAT_BASE;  // plain Material: base-class defaults (material.h:25-72)
    }
    int32_t id = static_cast<int32_t>(fs.materials.size());
    fs.materials.push_back(out);
    mat_ids[m] = id;
    return static_cast<uint32_t>(id);
  }

  uint32_t mesh(const Mesh* m) {
    auto it = mesh_ids.find(m);
    if (it != mesh_ids.end()) return static_cast<uint32_t>(it->second);
    vimg_mesh out{};
    out.indices = reinterpret_cast<const uint32_t*>(m->indices.data());
    out.n_tris = m->indices.size();
    out.vertices = reinterpret_cast<const float*>(m->vertices.data());
    out.n_verts = m->vertices.size();
    out.normals = m->normals.empty() ? nullptr : reinterpret_cast<const float*>(m->normals.data());
    fs.mesh_uv_ptrs.emplace_back();
    for (const auto& set : m->texcoords)
      fs.mesh_uv_ptrs.back().push_back(reinterpret_cast<const float*>(set.data()));
    out.n_uv_sets = static_cast<uint32_t>(m->texcoords.size());
    out.texcoords = reinterpret_cast<const float* const*>(fs.mesh_uv_ptrs.size());  // patched later
    out.color_tex_uv = m->color_tex_uv;
    out.normal_tex_uv = m->normal_tex_uv;
    out.metallic_roughness_tex_uv = m->metallic_roughness_tex_uv;
    out.material = material(m->mat);
    int32_t id = static_cast<int32_t>(fs.meshes.size());
    fs.meshes.push_back(out);
    mesh_ids[m] = id;
    return static_cast<uint32_t>(id);
  }

  // replace the 1-based table indices stashed in pointer fields by real pointers
  void finalise() {
    for (auto& t : fs.textures)
      if (t.type == VIMG_TEX_IMAGE)
        t.levels = fs.tex_level_ptrs[reinterpret_cast<size_t>(t.levels) - 1].data();
    for (auto& m : fs.meshes)
      m.texcoords = fs.mesh_uv_ptrs[reinterpret_cast<size_t>(m.texcoords) - 1].data();
  }
};

}  // namespace

std::unique_ptr<VimgFlatScene> vimg_flatten_scene(
    const integrator_data& settings, const BVH& bvh,
    const std::vector<std::unique_ptr<Surface>>& prims, const std::vector<Emitter*>& lights) {
  auto fs = std::make_unique<VimgFlatScene>();
  Interner in{*fs};
  vimg_scene_desc& d = fs->desc;
  d.abi_version = VIMG_ABI_VERSION;

  // camera (TLCam private members, include/tl_camera.h:11-18)
  const TLCam& cam = settings.camera;
  copy_mat4(cam.camToWorld_xform, d.camera.cam_to_world);
  d.camera.p_size[0] = cam.p_size[0];
  d.camera.p_size[1] = cam.p_size[1];
  d.camera.resolution[0] = cam.resolution[0];
  d.camera.resolution[1] = cam.resolution[1];
  d.camera.vfov_deg = cam.vfov;
  d.camera.aperture_radius = cam.aperture_radius;
  d.camera.focal_dist = cam.focal_dist;

  // BVH arrays are public (include/bvh.h:53-57) and already flat
  static_assert(sizeof(BVHNode) == 8 && sizeof(std::array<float, 3>) == 12);
  d.bvh.nodes = reinterpret_cast<const uint32_t*>(bvh.nodes.data());
  d.bvh.n_nodes = bvh.nodes.size();
  d.bvh.bb_mins_maxes = reinterpret_cast<const float*>(bvh.BB_mins_maxes.data());
  d.bvh.n_bb_slots = bvh.BB_mins_maxes.size();
  fs->obj_indices.assign(bvh.obj_indices.begin(), bvh.obj_indices.end());
  d.bvh.obj_indices = fs->obj_indices.data();
  d.bvh.n_obj_indices = fs->obj_indices.size();
  d.bvh.max_depth = bvh.max_depth;

  // primitives in list_objects order (== primitive index)
  fs->prims.reserve(prims.size());
  for (size_t i = 0; i < prims.size(); ++i) {
    Surface* s = prims[i].get();
    vimg_prim p{};
    if (auto* tri = dynamic_cast<Triangle*>(s)) {
      p.type = VIMG_PRIM_TRIANGLE;
      p.a = in.mesh(tri->obj_mesh);
      p.b = tri->tri_index;
      fs->emitter_to_prim[static_cast<const Emitter*>(tri)] = static_cast<int32_t>(i);
    } else if (auto* sph = dynamic_cast<Sphere*>(s)) {
      p.type = VIMG_PRIM_SPHERE;
      p.a = static_cast<uint32_t>(fs->spheres.size());
      vimg_sphere vs{};
      vs.center[0] = sph->center.x;
      vs.center[1] = sph->center.y;
      vs.center[2] = sph->center.z;
      vs.radius = sph->radius;
      vs.material = in.material(sph->mat);
      fs->spheres.push_back(vs);
      fs->emitter_to_prim[static_cast<const Emitter*>(sph)] = static_cast<int32_t>(i);
    } else {
      throw std::runtime_error("vimg adapter: unknown Surface subclass");
    }
    fs->prims.push_back(p);
  }

  // background (include/background.h)
  vimg_background& bg = d.background;
  bg.image = VIMG_NO_TEXTURE;
  const Background* b = settings.background.get();
  if (auto* cb = dynamic_cast<const ConstBackground*>(b)) {
    bg.type = VIMG_BG_CONST;
    std::memcpy(bg.color, &cb->col, 12);
  } else if (auto* env = dynamic_cast<const EnvMap*>(b)) {
    bg.type = VIMG_BG_ENVMAP;
    bg.image = in.texture(&env->image);
    copy_mat4(env->world_to_env, bg.world_to_env);
    copy_mat4(env->env_to_world, bg.env_to_world);
    bg.radiance_scale = env->radiance_scale;
    const ArraySampling2D& s2 = env->image_sampling;
    bg.cdf_width = s2.width;
    bg.cdf_height = s2.height;
    bg.row_cdf = s2.row_probabilities.cdf.data();
    fs->col_cdfs.reserve(size_t(s2.height) * (s2.width + 1));
    for (const auto& row : s2.image_probabilities)
      fs->col_cdfs.insert(fs->col_cdfs.end(), row.cdf.begin(), row.cdf.end());
    bg.col_cdfs = fs->col_cdfs.data();
  } else {
    throw std::runtime_error("vimg adapter: unknown Background subclass");
  }

  // lights in list_lights order (GroupOfEmitters copies that vector, emitters.h:34)
  for (const Emitter* e : lights) {
    if (e->is_background()) {
      fs->lights.push_back(VIMG_LIGHT_BACKGROUND);
    } else {
      auto it = fs->emitter_to_prim.find(e);
      if (it == fs->emitter_to_prim.end())
        throw std::runtime_error("vimg adapter: light is not a scene primitive");
      fs->lights.push_back(it->second);
    }
  }

  in.finalise();
  d.prims = fs->prims.data();
  d.n_prims = fs->prims.size();
  d.spheres = fs->spheres.data();
  d.n_spheres = static_cast<uint32_t>(fs->spheres.size());
  d.meshes = fs->meshes.data();
  d.n_meshes = static_cast<uint32_t>(fs->meshes.size());
  d.materials = fs->materials.data();
  d.n_materials = static_cast<uint32_t>(fs->materials.size());
  d.textures = fs->textures.data();
  d.n_textures = static_cast<uint32_t>(fs->textures.size());
  d.textures_rg = fs->textures_rg.data();
  d.n_textures_rg = static_cast<uint32_t>(fs->textures_rg.size());
  d.lights = fs->lights.data();
  d.n_lights = static_cast<uint32_t>(fs->lights.size());
  return fs;
}

#ifndef VIMG_ADAPTER_FLATTEN_ONLY  // the oracle build only needs the flattener
std::vector<glm::vec3> scene_integrator_cuda(const integrator_data& settings, const BVH& bvh,
                                             const std::vector<std::unique_ptr<Surface>>& prims,
                                             const std::vector<Emitter*>& lights,
                                             integrator_func func, int device_id) {
  auto fs = vimg_flatten_scene(settings, bvh, prims, lights);
  vimg_scene* scene = nullptr;
  if (vimg_scene_create(&fs->desc, device_id, &scene) != VIMG_OK)
    throw std::runtime_error(std::string("vimg_scene_create: ") + vimg_last_error());
  vimg_render_params rp{};
  rp.integrator = static_cast<uint32_t>(func);  // same order as enum class integrator_func
  rp.samples = settings.samples;
  rp.depth = settings.depth;
  rp.n_ranks = 1;
  std::vector<glm::vec3> image(size_t(settings.resolution[0]) * settings.resolution[1]);
  int rc = vimg_render(scene, &rp, reinterpret_cast<float*>(image.data()), nullptr);
  vimg_scene_destroy(scene);
  if (rc != VIMG_OK) throw std::runtime_error(std::string("vimg_render: ") + vimg_last_error());
  return image;
}

std::vector<glm::vec3> scene_integrator_cuda_multi(const integrator_data& settings, const BVH& bvh,
                                                   const std::vector<std::unique_ptr<Surface>>& prims,
                                                   const std::vector<Emitter*>& lights, integrator_func func,
                                                   const std::vector<int>& device_ids) {
  auto fs = vimg_flatten_scene(settings, bvh, prims, lights);
  vimg_render_params rp{};
  rp.integrator = static_cast<uint32_t>(func);
  rp.samples = settings.samples;
  rp.depth = settings.depth;
  std::vector<glm::vec3> image(size_t(settings.resolution[0]) * settings.resolution[1]);
  // 8x8 tiles dealt round-robin, like the reference's own work list (include/integrators.h:57-65,101):
  // the image is the one-device image bit for bit
  int rc = vimg_render_multi(&fs->desc, device_ids.data(), static_cast<uint32_t>(device_ids.size()), &rp,
                             VIMG_PARTITION_TILE, reinterpret_cast<float*>(image.data()), nullptr);
  if (rc != VIMG_OK) throw std::runtime_error(std::string("vimg_render_multi: ") + vimg_last_error());
  return image;
}
#endif  // VIMG_ADAPTER_FLATTEN_ONLY
