// ref_driver.cpp -- ORACLE / TEST INFRASTRUCTURE.  Not part of the product path.
//
// Links the UNMODIFIED reference sources (compiled where they lie under
// /root/reference by oracle/Makefile) behind a small C API so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// can (1) load a scene with the reference's own loaders and BVH builders,
// (2) render it with the reference's own scene_integrator<F>()
// (include/integrators.h:36-153), (3) dump bit-exact parity targets (primary
// hit primitive ids) and (4) hand out the scene flattened through the product's
// adapter (v-img_b200/adapter/vimg_adapter.cpp).
//
// Built in three flavours (oracle/Makefile):
//   parity : no -march, -ffp-contract=off  -> scalar slab path (bvh.h:118-126),
//            exact 1/dir, no FMA contraction.  THE bit-exact oracle.
//   perf   : -O3 -march=native, as the reference's CMake does (CMakeLists.txt:55-57)
//            -> AVX2 sibling-pair path.  THE timing baseline.
//   count  : parity flags + a build-time instrumented bvh.h (thread-local ray /
//            node / primitive-test counters) for the roofline cross-check.
#include <background.h>
#include <bvh.h>
#include <color_utils.h>
#include <geometry/mesh.h>
#include <geometry/sphere.h>
#include <geometry/triangle.h>
#include <integrators.h>
#include <material/dielectric.h>
#include <material/diffuse_light.h>
#include <material/lambertian.h>
#include <material/principled.h>
#include <scene_loading/json_scene.h>
#include <texture/texture_RG.h>
#include <texture/texture_RGB.h>

#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <thread>

#include "../v-img_b200/adapter/vimg_adapter.h"

#ifdef VREF_COUNTERS
// declared by oracle/count_hook.h, bumped by the build-time instrumented bvh.h
// (oracle/Makefile, count flavour)
std::atomic<uint64_t> vref_counters_g[6];
#endif

namespace {

struct RefScene {
  integrator_data settings;
  std::vector<std::unique_ptr<Material>> mat_list;
  std::vector<std::unique_ptr<Surface>> list_objects;
  std::vector<Emitter*> list_lights;
  std::vector<std::unique_ptr<Mesh>> list_meshes;
  std::vector<std::unique_ptr<TextureRGB>> texture_list;
  std::vector<std::unique_ptr<TextureRG>> textureRG_list;
  BVH bvh;
  std::unique_ptr<VimgFlatScene> flat;
  double bvh_build_s = 0;
};

// src/main.cpp:26-36 (setup_for_bvh) + :181-204, restated because main.cpp is
// not part of the oracle build.
void build_bvh(RefScene& sc, int bvh_type) {
  std::vector<AABB> bboxes;
  std::vector<glm::vec3> centers;
  bboxes.reserve(sc.list_objects.size());
  centers.reserve(sc.list_objects.size());
  for (auto const& obj : sc.list_objects) {
    AABB box = obj->bounds();
    bboxes.push_back(box);
    centers.push_back((box.bboxes[1] + box.bboxes[0]) * 0.5f);
  }
  auto t0 = std::chrono::steady_clock::now();
  if (bvh_type == 1) {
    std::vector<size_t> temp_indices(bboxes.size());
    std::iota(temp_indices.begin(), temp_indices.end(), 0);
    sc.bvh = BVH::build_sweep_bvh(bboxes, centers, temp_indices, 8);
  } else {
    sc.bvh = BVH::build_bin_bvh(bboxes, centers, 16);
  }
  sc.bvh_build_s =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

typedef glm::vec3 (*integrator_fn)(Ray&, std::vector<size_t>&, const BVH&,
                                   const std::vector<std::unique_ptr<Surface>>&,
                                   const GroupOfEmitters&, pcg32_random_t&, uint32_t, Background*);

integrator_fn pick(int integrator) {
  switch (integrator) {
    case 0: return shading_normal_integrator;
    case 1: return geometric_normal_integrator;
    case 2: return material_integrator;
    default: return mis_integrator;
  }
}

// ---- synthetic scene of SURVEY.md section 8(d) (configs C4 / C5), built through the reference's
// own C++ classes: height-field mesh + mip-mapped image texture + normal map + HDR env map + DOF.
float height_fn(float x, float z) {
  return 0.6f * std::sin(3.f * x) * std::cos(2.f * z)
         + 0.15f * std::sin(17.f * x + 1.f) * std::sin(13.f * z)
         + 0.03f * std::sin(71.f * x) * std::cos(67.f * z);
}

}  // namespace

extern "C" {

int vref_flavour() {
#if defined(VREF_COUNTERS)
  return 2;
#elif defined(__AVX2__)
  return 1;
#else
  return 0;
#endif
}

int vref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

void* vref_scene_from_json(const char* path, int bvh_type) {
  auto sc = std::make_unique<RefScene>();
  bool ok = set_scene_from_json(path, sc->settings, sc->list_objects, sc->mat_list,
                                sc->list_lights, sc->list_meshes, sc->textureRG_list,
                                sc->texture_list);
  if (!ok) return nullptr;
  sc->settings.num_threads = -1;
  build_bvh(*sc, bvh_type);
  return sc.release();
}

// G x G quads; tex_size^2 colour texture (mips) and normal map; env_w x env_h env map.
// with_rg != 0 adds a metallic-roughness TextureRG (a30) for coverage: 1 square, 2 wide (64 x 48).
void* vref_scene_synthetic(int G, int width, int height, int tex_size, int env_w, int env_h,
                           uint32_t spp, uint32_t depth, int bvh_type, int with_rg,
                           float aperture_radius) {
  auto sc = std::make_unique<RefScene>();
  const int V = G + 1;
  std::vector<glm::vec3> vertices(size_t(V) * V), normals(size_t(V) * V);
  std::vector<glm::vec2> uvs(size_t(V) * V);
  const float eps = 1e-3f;
  for (int j = 0; j < V; ++j)
    for (int i = 0; i < V; ++i) {
      float x = -5.f + 10.f * float(i) / float(G);
      float z = -5.f + 10.f * float(j) / float(G);
      size_t a = size_t(j) * V + i;
      vertices[a] = glm::vec3(x, height_fn(x, z), z);
      float dx = (height_fn(x + eps, z) - height_fn(x - eps, z)) / (2.f * eps);
      float dz = (height_fn(x, z + eps) - height_fn(x, z - eps)) / (2.f * eps);
      normals[a] = glm::normalize(glm::vec3(-dx, 1.f, -dz));
      uvs[a] = glm::vec2(8.f * float(i) / float(G), 8.f * float(j) / float(G));
    }
  std::vector<std::array<uint32_t, 3>> indices;
  indices.reserve(size_t(G) * G * 2);
  for (int j = 0; j < G; ++j)
    for (int i = 0; i < G; ++i) {
      uint32_t a = uint32_t(j) * V + i, b = a + 1, c = a + V, d = c + 1;
      indices.push_back({a, c, b});
      indices.push_back({b, c, d});
    }

  // colour texture with mips (mip-generating ctor, src/image_texture.cpp:60-130)
  std::vector<glm::vec3> img(size_t(tex_size) * tex_size);
  for (int y = 0; y < tex_size; ++y)
    for (int x = 0; x < tex_size; ++x) {
      float u = float(x) / float(tex_size), v = float(y) / float(tex_size);
      float c = 0.5f + 0.5f * std::sin(40.f * u) * std::sin(40.f * v);
      img[size_t(y) * tex_size + x] = glm::vec3(0.2f + 0.6f * c, 0.3f + 0.3f * u, 0.2f + 0.5f * v);
    }
  sc->texture_list.push_back(std::make_unique<ImageTexture>(
      img, tex_size, tex_size, TextureWrappingMode::Repeat, TextureWrappingMode::Repeat));
  TextureRGB* col_tex = sc->texture_list.back().get();

  // single-level normal map
  std::vector<glm::vec3> nimg(size_t(tex_size) * tex_size);
  for (int y = 0; y < tex_size; ++y)
    for (int x = 0; x < tex_size; ++x) {
      float u = float(x) / float(tex_size), v = float(y) / float(tex_size);
      nimg[size_t(y) * tex_size + x]
          = glm::normalize(glm::vec3(0.3f * std::cos(200.f * u), 0.3f * std::cos(200.f * v), 1.f));
    }
  std::vector<std::vector<glm::vec3>> nlevels{nimg};
  sc->texture_list.push_back(std::make_unique<ImageTexture>(
      nlevels, tex_size, tex_size, TextureWrappingMode::Repeat, TextureWrappingMode::Repeat));
  ImageTexture* normal_map = static_cast<ImageTexture*>(sc->texture_list.back().get());

  TextureRG* rg = nullptr;
  if (with_rg) {
    // with_rg = 2: a WIDE texture (64 x 48), where the row-stride slip of TextureRG::get_at_uv
    // (texture_RG.h:47,52: `next_x + y * height`) reads other texels than `next_x + y * width` would;
    // width > height keeps the reference's own indexing inside its vector
    const int R = 64, RH = with_rg == 2 ? 48 : 64;
    std::vector<glm::vec2> rgimg(size_t(R) * RH);
    for (int y = 0; y < RH; ++y)
      for (int x = 0; x < R; ++x)
        rgimg[size_t(y) * R + x] = glm::vec2(0.5f + 0.5f * std::sin(0.37f * x) * std::cos(0.21f * y),
                                             0.6f + 0.4f * std::cos(0.11f * (x + y)));
    sc->textureRG_list.push_back(std::make_unique<TextureRG>(
        R, RH, rgimg, TextureWrappingMode::MirroredRepeat, TextureWrappingMode::Repeat));
    rg = sc->textureRG_list.back().get();
  }

  sc->mat_list.push_back(std::make_unique<Principled>(
      col_tex, rg, glm::vec2(0.2f, 0.4f), /*spec_trans*/ 0.f, /*subsurface*/ 0.f,
      /*specular*/ 0.5f, /*spec_tint*/ 0.f, /*anisotropic*/ 0.f, /*sheen*/ 0.f,
      /*sheen_tint*/ 0.5f, /*clearcoat*/ 0.3f, /*clearcoat_gloss*/ 0.8f, /*eta*/ 1.5f, normal_map));
  Material* mat = sc->mat_list.back().get();

  std::vector<std::vector<glm::vec2>> uvsets{uvs};
  sc->list_meshes.push_back(std::make_unique<Mesh>(indices, vertices, normals, uvsets, mat,
                                                   /*color uv*/ 0, /*normal uv*/ 0,
                                                   with_rg ? 0 : MeshConsts::no_uv));
  std::vector<Emitter*> no_lights;
  add_tri_list_to_scene(sc->list_objects, sc->list_meshes.back().get(), no_lights);

  // env map (lat-long, clamp, identity transforms)
  std::vector<glm::vec3> env(size_t(env_w) * env_h);
  for (int y = 0; y < env_h; ++y)
    for (int x = 0; x < env_w; ++x) {
      float u = (float(x) + 0.5f) / float(env_w), v = (float(y) + 0.5f) / float(env_h);
      glm::vec3 sky = glm::vec3(0.5f, 0.7f, 1.0f) * (0.3f + 0.7f * (1.f - v));
      float du = u - 0.3f, dv = v - 0.25f;
      glm::vec3 sun = glm::vec3(1.f, 0.9f, 0.7f) * (400.f * std::exp(-4000.f * (du * du + dv * dv)));
      env[size_t(y) * env_w + x] = sky + sun;
    }
  ImageTexture env_img(env, env_w, env_h, TextureWrappingMode::ClampToEdge,
                       TextureWrappingMode::ClampToEdge);
  sc->settings.background
      = std::make_unique<EnvMap>(env_img, glm::mat4(1.0f), glm::mat4(1.0f), 1.0f);
  sc->list_lights.push_back(sc->settings.background.get());

  sc->settings.func = integrator_func::mis;
  sc->settings.resolution = glm::ivec2(width, height);
  sc->settings.samples = spp;
  sc->settings.depth = depth;
  sc->settings.camera = TLCam(camToWorld(glm::vec3(0, 3.5f, 9), glm::vec3(0, 0, 0), glm::vec3(0, 1, 0)),
                              glm::ivec2(width, height), 40.f, aperture_radius, 9.5f);
  sc->settings.num_threads = -1;
  build_bvh(*sc, bvh_type);
  return sc.release();
}

// Rebuild the reference's own scene objects from a flattened descriptor (the inverse of the
// adapter).  /root/reference and its scene files do not exist on the GPU box; the golden
// .vscene files do, so this is how the compiled reference renders C1-C3 there (cpu_baseline,
// --impl reference) and how tests prove that a .vscene carries the complete scene.  The BVH
// is taken from the descriptor, not rebuilt.
void* vref_scene_from_desc(const vimg_scene_desc* d, uint32_t spp, uint32_t depth, int integrator) {
  auto sc = std::make_unique<RefScene>();
  auto wrap = [](uint32_t w) {
    return w == VIMG_WRAP_CLAMP_TO_EDGE ? TextureWrappingMode::ClampToEdge
           : w == VIMG_WRAP_MIRRORED_REPEAT ? TextureWrappingMode::MirroredRepeat
                                            : TextureWrappingMode::Repeat;
  };
  auto v3 = [](const float* p) { return glm::vec3(p[0], p[1], p[2]); };
  auto m4 = [](const float* p) {
    glm::mat4 m;
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) m[c][r] = p[c * 4 + r];
    return m;
  };
  auto image_from = [&](const vimg_texture& t) {
    std::vector<std::vector<glm::vec3>> levels(t.n_levels);
    for (uint32_t l = 0; l < t.n_levels; ++l) {
      size_t w = std::max(t.width >> l, 1u), h = std::max(t.height >> l, 1u);
      levels[l].resize(w * h);
      std::memcpy(levels[l].data(), t.levels[l], w * h * 12);
    }
    return ImageTexture(levels, t.width, t.height, wrap(t.wrap_u), wrap(t.wrap_v));
  };
  for (uint32_t i = 0; i < d->n_textures; ++i) {
    const vimg_texture& t = d->textures[i];
    if (t.type == VIMG_TEX_CONST)
      sc->texture_list.push_back(std::make_unique<ConstColor>(v3(t.col_a)));
    else if (t.type == VIMG_TEX_CHECKER)
      sc->texture_list.push_back(std::make_unique<Checkerboard>(t.width, t.height, v3(t.col_a), v3(t.col_b)));
    else
      sc->texture_list.push_back(std::make_unique<ImageTexture>(image_from(t)));
  }
  for (uint32_t i = 0; i < d->n_textures_rg; ++i) {
    const vimg_texture_rg& t = d->textures_rg[i];
    std::vector<glm::vec2> data(size_t(t.width) * t.height);
    std::memcpy(data.data(), t.data, data.size() * 8);
    sc->textureRG_list.push_back(std::make_unique<TextureRG>(t.width, t.height, data, wrap(t.wrap_u), wrap(t.wrap_v)));
  }
  auto tex = [&](int32_t id) -> TextureRGB* { return id < 0 ? nullptr : sc->texture_list[id].get(); };
  for (uint32_t i = 0; i < d->n_materials; ++i) {
    const vimg_material& m = d->materials[i];
    ImageTexture* nmap = m.normal_map < 0 ? nullptr : static_cast<ImageTexture*>(tex(m.normal_map));
    switch (m.type) {
      case VIMG_MAT_LAMBERTIAN: sc->mat_list.push_back(std::make_unique<Lambertian>(tex(m.texture))); break;
      case VIMG_MAT_DIFFUSE_LIGHT: sc->mat_list.push_back(std::make_unique<DiffuseLight>(v3(m.emit))); break;
      case VIMG_MAT_DIELECTRIC: sc->mat_list.push_back(std::make_unique<Dielectric>(m.ior)); break;
      case VIMG_MAT_PRINCIPLED:
        sc->mat_list.push_back(std::make_unique<Principled>(
            tex(m.texture), m.texture_rg < 0 ? nullptr : sc->textureRG_list[m.texture_rg].get(),
            glm::vec2(m.metallic_roughness_factor[0], m.metallic_roughness_factor[1]),
            m.specular_transmission, m.subsurface, m.specular, m.specular_tint, m.anisotropic,
            m.sheen, m.sheen_tint, m.clearcoat, m.clearcoat_gloss, m.eta, nmap));
        break;
      default: sc->mat_list.push_back(std::make_unique<Material>()); break;
    }
    sc->mat_list.back()->normal_map = nmap;
  }
  for (uint32_t i = 0; i < d->n_meshes; ++i) {
    const vimg_mesh& m = d->meshes[i];
    std::vector<std::array<uint32_t, 3>> indices(m.n_tris);
    std::memcpy(indices.data(), m.indices, m.n_tris * 12);
    std::vector<glm::vec3> vertices(m.n_verts), normals;
    std::memcpy(vertices.data(), m.vertices, m.n_verts * 12);
    if (m.normals) {
      normals.resize(m.n_verts);
      std::memcpy(normals.data(), m.normals, m.n_verts * 12);
    }
    std::vector<std::vector<glm::vec2>> uvsets(m.n_uv_sets);
    for (uint32_t k = 0; k < m.n_uv_sets; ++k) {
      uvsets[k].resize(m.n_verts);
      std::memcpy(uvsets[k].data(), m.texcoords[k], m.n_verts * 8);
    }
    sc->list_meshes.push_back(std::make_unique<Mesh>(indices, vertices, normals, uvsets,
                                                     sc->mat_list[m.material].get(), m.color_tex_uv,
                                                     m.normal_tex_uv, m.metallic_roughness_tex_uv));
  }
  for (uint64_t i = 0; i < d->n_prims; ++i) {
    const vimg_prim& p = d->prims[i];
    if (p.type == VIMG_PRIM_SPHERE) {
      const vimg_sphere& s = d->spheres[p.a];
      sc->list_objects.push_back(std::make_unique<Sphere>(v3(s.center), s.radius, sc->mat_list[s.material].get()));
    } else {
      sc->list_objects.push_back(std::make_unique<Triangle>(sc->list_meshes[p.a].get(), p.b));
    }
  }
  const vimg_background& bg = d->background;
  if (bg.type == VIMG_BG_ENVMAP) {
    ImageTexture img = image_from(d->textures[bg.image]);
    sc->settings.background = std::make_unique<EnvMap>(img, m4(bg.world_to_env), m4(bg.env_to_world), bg.radiance_scale);
  } else {
    sc->settings.background = std::make_unique<ConstBackground>(v3(bg.color));
  }
  for (uint32_t i = 0; i < d->n_lights; ++i) {
    if (d->lights[i] == VIMG_LIGHT_BACKGROUND)
      sc->list_lights.push_back(sc->settings.background.get());
    else
      sc->list_lights.push_back(dynamic_cast<Emitter*>(sc->list_objects[d->lights[i]].get()));
  }
  const vimg_camera& c = d->camera;
  sc->settings.camera = TLCam(m4(c.cam_to_world), glm::ivec2(c.resolution[0], c.resolution[1]),
                              c.vfov_deg, c.aperture_radius, c.focal_dist);
  sc->settings.camera.p_size[0] = c.p_size[0];  // keep the flattened values (-fno-access-control)
  sc->settings.camera.p_size[1] = c.p_size[1];
  sc->settings.resolution = glm::ivec2(c.resolution[0], c.resolution[1]);
  sc->settings.samples = spp;
  sc->settings.depth = depth;
  sc->settings.func = static_cast<integrator_func>(integrator);
  sc->settings.num_threads = -1;
  sc->bvh.nodes.resize(d->bvh.n_nodes);
  std::memcpy(sc->bvh.nodes.data(), d->bvh.nodes, d->bvh.n_nodes * 8);
  sc->bvh.BB_mins_maxes.resize(d->bvh.n_bb_slots);
  std::memcpy(sc->bvh.BB_mins_maxes.data(), d->bvh.bb_mins_maxes, d->bvh.n_bb_slots * 12);
  sc->bvh.obj_indices.assign(d->bvh.obj_indices, d->bvh.obj_indices + d->bvh.n_obj_indices);
  sc->bvh.max_depth = d->bvh.max_depth;
  return sc.release();
}

void vref_scene_free(void* h) { delete static_cast<RefScene*>(h); }

// info[0..7] = width, height, spp, depth, integrator, n_prims, n_nodes, max_depth
void vref_scene_info(void* h, uint64_t info[8], double* bvh_build_s) {
  auto* sc = static_cast<RefScene*>(h);
  info[0] = sc->settings.resolution[0];
  info[1] = sc->settings.resolution[1];
  info[2] = sc->settings.samples;
  info[3] = sc->settings.depth;
  info[4] = (uint64_t)sc->settings.func;
  info[5] = sc->list_objects.size();
  info[6] = sc->bvh.nodes.size();
  info[7] = sc->bvh.max_depth;
  if (bvh_build_s) *bvh_build_s = sc->bvh_build_s;
}

// Change a setting the JSON format cannot express (background colour is parsed
// and ignored, json_scene.cpp:202-206; aperture is fixed to 0, :179).
void vref_scene_set_const_background(void* h, float r, float g, float b, int as_light) {
  auto* sc = static_cast<RefScene*>(h);
  sc->settings.background = std::make_unique<ConstBackground>(glm::vec3(r, g, b));
  // drop any previous background light, then optionally add the new one
  std::vector<Emitter*> kept;
  for (Emitter* e : sc->list_lights)
    if (!e->is_background()) kept.push_back(e);
  sc->list_lights = kept;
  if (as_light) sc->list_lights.push_back(sc->settings.background.get());
  sc->flat.reset();
}

void vref_scene_set_camera_dof(void* h, float aperture_radius, float focal_dist) {
  auto* sc = static_cast<RefScene*>(h);
  sc->settings.camera.aperture_radius = aperture_radius;  // -fno-access-control
  sc->settings.camera.focal_dist = focal_dist;
  sc->flat.reset();
}

// The scene flattened by the product's adapter; owned by the RefScene handle.
const vimg_scene_desc* vref_scene_desc(void* h) {
  auto* sc = static_cast<RefScene*>(h);
  try {
    if (!sc->flat)
      sc->flat = vimg_flatten_scene(sc->settings, sc->bvh, sc->list_objects, sc->list_lights);
  } catch (const std::exception& e) {
    fmt::println("vref_scene_desc: {}", e.what());
    return nullptr;
  }
  return &sc->flat->desc;
}

// Render with the reference's own scene_integrator.  spp == 0 / depth == 0 keep the
// scene-file values.  Returns seconds spent inside scene_integrator (the span
// src/main.cpp:211,300-302 prints as "image rendering").
double vref_render(void* h, int integrator, uint32_t spp, uint32_t depth, int threads,
                   float* out_rgb) {
  auto* sc = static_cast<RefScene*>(h);
  integrator_data& s = sc->settings;
  uint32_t keep_spp = s.samples, keep_depth = s.depth;
  if (spp) s.samples = spp;
  if (depth) s.depth = depth;
  s.num_threads = (int16_t)threads;
  GroupOfEmitters lights(sc->list_lights);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<glm::vec3> img
      = scene_integrator(s, sc->bvh, sc->list_objects, lights, pick(integrator));
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  s.samples = keep_spp;
  s.depth = keep_depth;
  if (out_rgb) std::memcpy(out_rgb, img.data(), img.size() * sizeof(glm::vec3));
  return sec;
}

#ifdef VREF_WITH_CUDA_BINDING
// The binding flavour (oracle/Makefile `binding`): the adapter TU compiled WITH scene_integrator_cuda and
// linked against libvimg_cuda.so -- what a v-img maintainer's build contains.  Same
// arguments as vref_render; the image comes from the cuda device(s) through the reference-side call that
// replaces scene_integrator<F>() at the seam of src/main.cpp:219-248.  n_devices > 1: vimg_render_multi.
// Returns 0, or -1 with the message in err (nullable).
int vref_render_cuda(void* h, int integrator, uint32_t spp, uint32_t depth, const int* devices, int n_devices,
                     float* out_rgb, char* err, int err_len) {
  auto* sc = static_cast<RefScene*>(h);
  integrator_data& s = sc->settings;
  uint32_t keep_spp = s.samples, keep_depth = s.depth;
  if (spp) s.samples = spp;
  if (depth) s.depth = depth;
  int rc = 0;
  try {
    std::vector<glm::vec3> img
        = n_devices > 1 ? scene_integrator_cuda_multi(s, sc->bvh, sc->list_objects, sc->list_lights, static_cast<integrator_func>(integrator),
                                                      std::vector<int>(devices, devices + n_devices))
                        : scene_integrator_cuda(s, sc->bvh, sc->list_objects, sc->list_lights, static_cast<integrator_func>(integrator),
                                                n_devices == 1 ? devices[0] : 0);
    if (out_rgb) std::memcpy(out_rgb, img.data(), img.size() * sizeof(glm::vec3));
  } catch (const std::exception& e) {
    if (err && err_len > 0) {
      std::strncpy(err, e.what(), err_len - 1);
      err[err_len - 1] = 0;
    }
    rc = -1;
  }
  s.samples = keep_spp;
  s.depth = keep_depth;
  return rc;
}
#endif

// The post-process chain of src/main.cpp:304-356 on a W*H linear image: the reference's OWN
// tonemapper (color_utils.h:21-25 simple_clamp, src/tonemap/{agx,reinhard,aces}.cpp), its sRGB
// OETF (color_utils.h:50-68), then main.cpp's NaN -> magenta and int(255.999 * v) quantisation
// loop (main.cpp:339-356; it lives inside main(), so it is restated here line by line).
// tonemapper: 0 clamp, 1 agx, 2 reinhard, 3 aces (enum class tonemapper), anything else = none.
// out_srgb (nullable): the float image after gamma; out_rgb8 (nullable): the PNG bytes.
void vref_postprocess(const float* rgb, uint64_t n_pixels, int tonemapper, float* out_srgb,
                      uint8_t* out_rgb8) {
  std::vector<glm::vec3> acc_image(n_pixels);
  std::memcpy(acc_image.data(), rgb, n_pixels * sizeof(glm::vec3));
  switch (tonemapper) {
    case 0: simple_clamp(acc_image); break;
    case 1: agx(acc_image); break;
    case 2: reinhard_lum(acc_image); break;
    case 3: aces(acc_image); break;
    default: break;
  }
  sRGB_gamma_correction(acc_image);
  if (out_srgb) std::memcpy(out_srgb, acc_image.data(), n_pixels * sizeof(glm::vec3));
  if (!out_rgb8) return;
  uint64_t index = 0;
  for (size_t i = 0; i < acc_image.size(); i++) {
    if (std::isnan(acc_image[i][0]) || std::isnan(acc_image[i][1]) || std::isnan(acc_image[i][2])) {
      out_rgb8[index++] = 255;
      out_rgb8[index++] = 0;
      out_rgb8[index++] = 255;
    } else {
      out_rgb8[index++] = std::clamp(static_cast<int>(255.999 * acc_image[i][0]), 0, 255);
      out_rgb8[index++] = std::clamp(static_cast<int>(255.999 * acc_image[i][1]), 0, 255);
      out_rgb8[index++] = std::clamp(static_cast<int>(255.999 * acc_image[i][2]), 0, 255);
    }
  }
}

// The reference's glTF texture conversions: make_texture's byte -> float loop
// (src/scene_loading/gltf_loading.cpp:139-143; the function is static there, so the three lines are
// restated) followed by the reference's OWN ImageTexture::convert_sRGB_to_linear /
// convert_RGB_to_normal (src/image_texture.cpp:257-275).  out: n_pixels * 3 floats.
static std::vector<glm::vec3> bytes_to_image(const uint8_t* char_arr, uint64_t n_pixels) {
  std::vector<glm::vec3> image(n_pixels);
  for (size_t i = 0; i < image.size(); i++) {
    image[i].x = char_arr[i * 3];
    image[i].y = char_arr[i * 3 + 1];
    image[i].z = char_arr[i * 3 + 2];
  }
  return image;
}
void vref_convert_srgb8_to_linear(const uint8_t* rgb8, uint64_t n_pixels, float* out) {
  std::vector<glm::vec3> image = bytes_to_image(rgb8, n_pixels);
  ImageTexture::convert_sRGB_to_linear(image);
  std::memcpy(out, image.data(), n_pixels * sizeof(glm::vec3));
}
void vref_convert_rgb8_to_normal(const uint8_t* rgb8, uint64_t n_pixels, float scale, float* out) {
  std::vector<glm::vec3> image = bytes_to_image(rgb8, n_pixels);
  ImageTexture::convert_RGB_to_normal(image, scale);
  std::memcpy(out, image.data(), n_pixels * sizeof(glm::vec3));
}

// The reference's own heat map (src/integrators/heatmap.cpp:39-147, called from main.cpp:249-255
// with spp fixed to 4).  factor <= 0 takes the reference's default 20.
void vref_heatmap(void* h, uint32_t spp, float factor, float* out_rgb) {
  auto* sc = static_cast<RefScene*>(h);
  integrator_data& s = sc->settings;
  uint32_t keep_spp = s.samples;
  if (spp) s.samples = spp;
  std::vector<glm::vec3> img = heatmap_img(s, sc->bvh, sc->list_objects, factor);
  s.samples = keep_spp;
  std::memcpy(out_rgb, img.data(), img.size() * sizeof(glm::vec3));
}

// trace_pixel (include/integrators.h:181-220)
void vref_trace_pixel(void* h, int integrator, uint32_t spp, uint32_t depth, int x, int y,
                      float out_rgb[3]) {
  auto* sc = static_cast<RefScene*>(h);
  integrator_data& s = sc->settings;
  uint32_t keep_spp = s.samples, keep_depth = s.depth;
  if (spp) s.samples = spp;
  if (depth) s.depth = depth;
  GroupOfEmitters lights(sc->list_lights);
  glm::vec3 c = trace_pixel(s, sc->bvh, sc->list_objects, lights, pick(integrator), x, y);
  s.samples = keep_spp;
  s.depth = keep_depth;
  out_rgb[0] = c.x;
  out_rgb[1] = c.y;
  out_rgb[2] = c.z;
}

// Primary-hit primitive ids for every sample of every pixel, through the reference's own
// camera, RNG, R2 sequence and BVH::hit (the loop of include/integrators.h:109-127 with the
// integrator call replaced by the closest-hit query).  index = image_index * spp + s.
void vref_trace_ids(void* h, uint32_t spp, int32_t* out_ids, float* out_t) {
  auto* sc = static_cast<RefScene*>(h);
  const integrator_data& rd = sc->settings;
  const size_t W = rd.resolution[0], H = rd.resolution[1];
  std::unordered_map<const Emitter*, int32_t> to_prim;
  for (size_t i = 0; i < sc->list_objects.size(); ++i)
    to_prim[dynamic_cast<const Emitter*>(sc->list_objects[i].get())] = (int32_t)i;

#pragma omp parallel
  {
    std::vector<size_t> stack;
    stack.reserve(sc->bvh.max_depth + 2);
    pcg32_random_t st;
#pragma omp for schedule(dynamic, 4)
    for (long long yy = 0; yy < (long long)H; ++yy) {
      size_t y = (size_t)yy;
      for (size_t x = 0; x < W; ++x) {
        size_t image_index = x + ((H - 1 - y) * W);
        size_t image_seq_start = x + y;
        pcg32_srandom_r(&st, image_index, 0);
        for (size_t s = 0; s < spp; ++s) {
          glm::vec2 off = random_x_y_r2(image_seq_start + s);
          Ray ray = rd.camera.generate_ray(x + off.x, y + off.y, rand_float(st), rand_float(st));
          auto hit = sc->bvh.hit<std::optional<HitInfo>>(ray, stack, sc->list_objects);
          int32_t id = -1;
          if (hit.has_value()) id = to_prim.at(hit.value().obj);
          out_ids[image_index * spp + s] = id;
          if (out_t) out_t[image_index * spp + s] = ray.maxT;
        }
      }
    }
  }
}

// counters[0..5] = rays_closest, rays_shadow, inner nodes popped, leaves popped, primitive
// tests, (unused) since the last reset (count flavour only; zeros otherwise).
void vref_counters(uint64_t out[6], int reset) {
#ifdef VREF_COUNTERS
  for (int i = 0; i < 6; ++i) {
    out[i] = vref_counters_g[i].load();
    if (reset) vref_counters_g[i].store(0);
  }
#else
  for (int i = 0; i < 6; ++i) out[i] = 0;
  (void)reset;
#endif
}

}  // extern "C"
