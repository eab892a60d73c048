// vimg_api.cu -- C ABI of libvimg_cuda.so (include/vimg_cuda.h): scene upload, the
// GPU layout build, kernel launches.  No CPU fallback exists: every compute entry point
// needs a CUDA device and fails loudly otherwise.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <numbers>
#include <string>
#include <vector>

#include "vimg_kernels.cuh"

using namespace vimg;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CU_TRY(expr)                                                                       \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return fail(e__ == cudaErrorMemoryAllocation ? VIMG_ERR_OUT_OF_MEMORY : VIMG_ERR_CUDA, \
                  std::string(#expr) + ": " + cudaGetErrorString(e__));                    \
  } while (0)

// host-side mirror of vimg_math.cuh (this TU is built with -ffp-contract=off and no
// -march, so these evaluate exactly like the device versions and like the oracle)
struct H3 { float x, y, z; };
H3 h_sub(H3 a, H3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
H3 h_cross(H3 a, H3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
float h_dot(H3 a, H3 b) {
  float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
  return tx + ty + tz;
}

}  // namespace

struct vimg_scene {
  int device = 0;
  DScene d{};
  std::vector<void*> allocs;
  uint64_t device_bytes = 0;
  int W = 0, H = 0;
  float* d_frame = nullptr;
  DCounters* d_counters = nullptr;
  unsigned int* d_next_pixel = nullptr;
  int wf_blocks[2][2] = {{0, 0}, {0, 0}};  // resident-grid size of wavefront_kernel<integrator, COUNT>
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  template <typename T> int upload(const T* host, size_t n, const T** out) {
    *out = nullptr;
    if (n == 0) return VIMG_OK;
    void* p = nullptr;
    CU_TRY(cudaMalloc(&p, n * sizeof(T)));
    allocs.push_back(p);
    device_bytes += n * sizeof(T);
    CU_TRY(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<const T*>(p);
    return VIMG_OK;
  }
  template <typename T> int upload(const std::vector<T>& v, const T** out) {
    return upload(v.data(), v.size(), out);
  }
};

static int build_scene(const vimg_scene_desc* desc, vimg_scene* s) {
  DScene& d = s->d;
  // ---- camera ------------------------------------------------------------
  const vimg_camera& c = desc->camera;
  std::memcpy(d.cam.m, c.cam_to_world, sizeof(d.cam.m));
  d.cam.p_size[0] = c.p_size[0];
  d.cam.p_size[1] = c.p_size[1];
  d.cam.res[0] = c.resolution[0];
  d.cam.res[1] = c.resolution[1];
  d.cam.aperture_radius = c.aperture_radius;
  d.cam.focal_dist = c.focal_dist;
  {
    // TLCam::generate_ray: raycone_for_primary_ray((vfov * pi) / 180.f, resolution.y)
    // (src/tl_camera.cpp:50-51, include/ray.h:44-48); constant per camera.
    float vfov = static_cast<float>((c.vfov_deg * std::numbers::pi) / 180.f);
    d.cam.cone_spread
        = std::atan(2.f * (std::tan(vfov / 2.f)) / static_cast<float>((uint32_t)c.resolution[1]));
  }
  s->W = c.resolution[0];
  s->H = c.resolution[1];
  if (s->W <= 0 || s->H <= 0) return fail(VIMG_ERR_INVALID_ARGUMENT, "camera resolution must be positive");

  // ---- geometry pools ------------------------------------------------------
  std::vector<DMesh> meshes(desc->n_meshes);
  std::vector<uint32_t> tri_indices;
  std::vector<float> vertices, normals, uvs;
  for (uint32_t m = 0; m < desc->n_meshes; ++m) {
    const vimg_mesh& hm = desc->meshes[m];
    DMesh& dm = meshes[m];
    dm.vert_base = (uint32_t)(vertices.size() / 3);
    dm.tri_base = (uint32_t)(tri_indices.size() / 3);
    dm.has_normals = hm.normals != nullptr;
    dm.material = hm.material;
    if (hm.material >= desc->n_materials) return fail(VIMG_ERR_INVALID_ARGUMENT, "mesh material index out of range");
    tri_indices.insert(tri_indices.end(), hm.indices, hm.indices + 3 * hm.n_tris);
    vertices.insert(vertices.end(), hm.vertices, hm.vertices + 3 * hm.n_verts);
    if (hm.normals)
      normals.insert(normals.end(), hm.normals, hm.normals + 3 * hm.n_verts);
    else
      normals.resize(normals.size() + 3 * hm.n_verts, 0.f);
    std::vector<int32_t> set_off(hm.n_uv_sets);
    for (uint32_t k = 0; k < hm.n_uv_sets; ++k) {
      set_off[k] = (int32_t)(uvs.size() / 2);
      uvs.insert(uvs.end(), hm.texcoords[k], hm.texcoords[k] + 2 * hm.n_verts);
    }
    auto pick = [&](uint8_t sel) -> int32_t {
      if (sel == VIMG_NO_UV) return -1;
      return sel < hm.n_uv_sets ? set_off[sel] : -2;
    };
    dm.uv_color = pick(hm.color_tex_uv);
    dm.uv_normal = pick(hm.normal_tex_uv);
    dm.uv_mr = pick(hm.metallic_roughness_tex_uv);
    if (dm.uv_color == -2 || dm.uv_normal == -2 || dm.uv_mr == -2)
      return fail(VIMG_ERR_INVALID_ARGUMENT, "mesh uv-set selector out of range");
    for (uint64_t t = 0; t < 3 * hm.n_tris; ++t)
      if (hm.indices[t] >= hm.n_verts) return fail(VIMG_ERR_INVALID_ARGUMENT, "triangle index out of range");
  }

  // ---- BVH: pair records + leaf primitives in obj_indices order -------------
  const vimg_bvh& b = desc->bvh;
  d.n_nodes = (uint32_t)b.n_nodes;
  d.stack_cap = b.max_depth + 2;
  if (d.stack_cap > VIMG_STACK_MAX)
    return fail(VIMG_ERR_UNSUPPORTED, "BVH max_depth + 2 exceeds the traversal stack (VIMG_STACK_MAX)");
  std::vector<DPairNode> pairs;
  std::vector<DLeafPrim> leaf(b.n_obj_indices);
  if (b.n_nodes) {
    if (b.n_nodes % 2 == 0) return fail(VIMG_ERR_INVALID_ARGUMENT, "BVH node count must be odd");
    if (b.n_bb_slots < 2 * b.n_nodes + 2) return fail(VIMG_ERR_INVALID_ARGUMENT, "BB_mins_maxes too short");
    pairs.resize((b.n_nodes - 1) / 2);
    const float* bb = b.bb_mins_maxes;
    std::memcpy(d.root_min, bb + 0, 12);
    std::memcpy(d.root_max, bb + 6, 12);
    d.root_first = b.nodes[0];
    d.root_count = b.nodes[1];
    for (uint64_t n = 0; n < b.n_nodes; ++n) {
      const uint32_t first = b.nodes[2 * n], count = b.nodes[2 * n + 1];
      if (count != 0) {
        if ((uint64_t)first + count > b.n_obj_indices) return fail(VIMG_ERR_INVALID_ARGUMENT, "leaf range out of bounds");
        continue;
      }
      // inner: children (first, first+1); pair slots 2c+2 .. 2c+5 = Lmin,Rmin,Lmax,Rmax
      if ((first & 1u) == 0 || (uint64_t)first + 1 >= b.n_nodes)
        return fail(VIMG_ERR_INVALID_ARGUMENT, "inner node with an even or out-of-range first child");
      DPairNode& p = pairs[(first - 1) / 2];
      const float* slot = bb + 3 * ((uint64_t)first * 2 + 2);
      std::memcpy(p.lmin, slot, 48);  // lmin, rmin, lmax, rmax are contiguous in both layouts
      p.l_first = b.nodes[2 * (uint64_t)first];
      p.l_count = b.nodes[2 * (uint64_t)first + 1];
      p.r_first = b.nodes[2 * ((uint64_t)first + 1)];
      p.r_count = b.nodes[2 * ((uint64_t)first + 1) + 1];
    }
  }
  for (uint64_t i = 0; i < b.n_obj_indices; ++i) {
    const uint64_t prim = b.obj_indices[i];
    if (prim >= desc->n_prims) return fail(VIMG_ERR_INVALID_ARGUMENT, "obj_indices entry out of range");
    const vimg_prim& pr = desc->prims[prim];
    DLeafPrim& lp = leaf[i];
    std::memset(&lp, 0, sizeof(lp));
    lp.prim_id = (uint32_t)prim;
    if (pr.type == VIMG_PRIM_SPHERE) {
      if (pr.a >= desc->n_spheres) return fail(VIMG_ERR_INVALID_ARGUMENT, "sphere index out of range");
      const vimg_sphere& sp = desc->spheres[pr.a];
      lp.a[0] = sp.center[0];
      lp.a[1] = sp.center[1];
      lp.a[2] = sp.center[2];
      lp.a[3] = sp.radius;
      lp.flags = VIMG_LEAF_SPHERE;
    } else {
      if (pr.a >= desc->n_meshes || pr.b >= desc->meshes[pr.a].n_tris)
        return fail(VIMG_ERR_INVALID_ARGUMENT, "triangle primitive out of range");
      const vimg_mesh& hm = desc->meshes[pr.a];
      const uint32_t* idx = hm.indices + 3 * (uint64_t)pr.b;
      for (int k = 0; k < 3; ++k) std::memcpy(lp.a + 3 * k, hm.vertices + 3 * (uint64_t)idx[k], 12);
      // Triangle::tri_hit_template's degenerate reject (triangle.h:89-92), ray independent
      H3 p0{lp.a[0], lp.a[1], lp.a[2]}, p1{lp.a[3], lp.a[4], lp.a[5]}, p2{lp.a[6], lp.a[7], lp.a[8]};
      H3 cr = h_cross(h_sub(p2, p0), h_sub(p1, p0));
      if (h_dot(cr, cr) == 0.f) lp.flags |= VIMG_LEAF_DEGENERATE;
    }
  }

  // ---- textures --------------------------------------------------------------
  std::vector<DTexture> textures(desc->n_textures);
  std::vector<float> texels;
  for (uint32_t t = 0; t < desc->n_textures; ++t) {
    const vimg_texture& ht = desc->textures[t];
    DTexture& dt = textures[t];
    std::memset(&dt, 0, sizeof(dt));
    dt.type = ht.type;
    std::memcpy(dt.col_a, ht.col_a, 12);
    std::memcpy(dt.col_b, ht.col_b, 12);
    dt.width = ht.width;
    dt.height = ht.height;
    dt.n_levels = ht.n_levels;
    dt.wrap_u = ht.wrap_u;
    dt.wrap_v = ht.wrap_v;
    if (ht.type == VIMG_TEX_IMAGE) {
      if (ht.n_levels == 0 || ht.n_levels > 16) return fail(VIMG_ERR_INVALID_ARGUMENT, "image texture needs 1..16 mip levels");
      for (uint32_t l = 0; l < ht.n_levels; ++l) {
        const uint64_t w = std::max(ht.width >> l, 1u), h = std::max(ht.height >> l, 1u);
        dt.level_off[l] = (uint32_t)(texels.size() / 3);
        texels.insert(texels.end(), ht.levels[l], ht.levels[l] + 3 * w * h);
      }
    }
  }
  std::vector<DTextureRG> textures_rg(desc->n_textures_rg);
  std::vector<float> rg_texels;
  for (uint32_t t = 0; t < desc->n_textures_rg; ++t) {
    const vimg_texture_rg& ht = desc->textures_rg[t];
    textures_rg[t] = DTextureRG{ht.width, ht.height, ht.wrap_u, ht.wrap_v, (uint32_t)(rg_texels.size() / 2)};
    // the reference's row-stride quirk (texture_RG.h:47,52) can index up to
    // (w-1) + (h-1)*max(w,h); pad so a non-square texture never reads out of bounds
    const uint64_t n = (uint64_t)ht.width * ht.height;
    const uint64_t reach = (uint64_t)(ht.width - 1) + (uint64_t)(ht.height - 1) * std::max(ht.width, ht.height) + 1;
    rg_texels.insert(rg_texels.end(), ht.data, ht.data + 2 * n);
    if (reach > n) rg_texels.resize(rg_texels.size() + 2 * (reach - n), 0.f);
  }
  for (uint32_t m = 0; m < desc->n_materials; ++m) {
    const vimg_material& hm = desc->materials[m];
    auto tex_ok = [&](int32_t id, bool need_image) {
      if (id < 0) return true;
      if ((uint32_t)id >= desc->n_textures) return false;
      return !need_image || desc->textures[id].type == VIMG_TEX_IMAGE;
    };
    if (!tex_ok(hm.texture, false) || !tex_ok(hm.normal_map, true)
        || (hm.texture_rg >= 0 && (uint32_t)hm.texture_rg >= desc->n_textures_rg))
      return fail(VIMG_ERR_INVALID_ARGUMENT, "material texture index out of range");
    if ((hm.type == VIMG_MAT_LAMBERTIAN || hm.type == VIMG_MAT_PRINCIPLED) && hm.texture < 0)
      return fail(VIMG_ERR_INVALID_ARGUMENT, "lambertian/principled material without a texture");
  }
  for (uint32_t i = 0; i < desc->n_spheres; ++i)
    if (desc->spheres[i].material >= desc->n_materials) return fail(VIMG_ERR_INVALID_ARGUMENT, "sphere material index out of range");
  for (uint32_t i = 0; i < desc->n_lights; ++i)
    if (desc->lights[i] != VIMG_LIGHT_BACKGROUND && (desc->lights[i] < 0 || (uint64_t)desc->lights[i] >= desc->n_prims))
      return fail(VIMG_ERR_INVALID_ARGUMENT, "light index out of range");

  // ---- background ------------------------------------------------------------
  const vimg_background& hb = desc->background;
  d.bg.type = hb.type;
  std::memcpy(d.bg.color, hb.color, 12);
  d.bg.image = hb.image;
  std::memcpy(d.bg.world_to_env, hb.world_to_env, 64);
  std::memcpy(d.bg.env_to_world, hb.env_to_world, 64);
  d.bg.radiance_scale = hb.radiance_scale;
  d.bg.cdf_w = hb.cdf_width;
  d.bg.cdf_h = hb.cdf_height;
  if (hb.type == VIMG_BG_ENVMAP) {
    if (hb.image < 0 || (uint32_t)hb.image >= desc->n_textures || desc->textures[hb.image].type != VIMG_TEX_IMAGE)
      return fail(VIMG_ERR_INVALID_ARGUMENT, "env map needs an image texture");
    d.bg.emissive = 1;  // EnvMap::is_emissive() is always true (background.h:176)
  } else {
    // ConstBackground::is_emissive(): col != vec3(0) (background.h:50-55)
    d.bg.emissive = !(hb.color[0] == 0.f && hb.color[1] == 0.f && hb.color[2] == 0.f);
  }

  // ---- upload ------------------------------------------------------------------
  int rc;
#define UP(call) if ((rc = (call)) != VIMG_OK) return rc
  UP(s->upload(pairs, &d.pairs));
  UP(s->upload(leaf, &d.leaf_prims));
  UP(s->upload(desc->prims, desc->n_prims, &d.prims));
  UP(s->upload(desc->spheres, desc->n_spheres, &d.spheres));
  UP(s->upload(meshes, &d.meshes));
  UP(s->upload(tri_indices, &d.tri_indices));
  UP(s->upload(vertices, &d.vertices));
  UP(s->upload(normals, &d.normals));
  UP(s->upload(uvs, &d.uvs));
  UP(s->upload(desc->materials, desc->n_materials, &d.materials));
  UP(s->upload(textures, &d.textures));
  UP(s->upload(texels, &d.texels));
  UP(s->upload(textures_rg, &d.textures_rg));
  UP(s->upload(rg_texels, &d.rg_texels));
  UP(s->upload(desc->lights, desc->n_lights, &d.lights));
  d.n_lights = desc->n_lights;
  if (hb.type == VIMG_BG_ENVMAP) {
    UP(s->upload(hb.row_cdf, (size_t)hb.cdf_height + 1, &d.bg.row_cdf));
    UP(s->upload(hb.col_cdfs, (size_t)hb.cdf_height * (hb.cdf_width + 1), &d.bg.col_cdfs));
  }
#undef UP
  return VIMG_OK;
}

static RenderArgs make_args(const vimg_scene* s, uint32_t samples, uint32_t depth, uint32_t rank, uint32_t n_ranks) {
  RenderArgs a{};
  a.samples = samples;
  a.depth = depth;
  a.n_ranks = n_ranks ? n_ranks : 1;
  a.rank = rank;
  a.tiles_x = (s->W + 7) / 8;
  a.tiles_y = (s->H + 7) / 8;
  const uint32_t n_tiles = a.tiles_x * a.tiles_y;
  a.n_local_tiles = rank < n_tiles ? (n_tiles - rank + a.n_ranks - 1) / a.n_ranks : 0;
  a.single_x = a.single_y = -1;
  return a;
}

// persistent grid of the wavefront kernel: as many blocks as are resident at once
template <int INTEGRATOR, bool COUNT>
static int wavefront_grid(vimg_scene* s, uint32_t n_pixels) {
  int& cached = s->wf_blocks[INTEGRATOR == VIMG_INTEGRATOR_MIS][COUNT];
  if (cached == 0) {
    int per_sm = 0, sms = 0;
    cudaFuncSetAttribute(wavefront_kernel<INTEGRATOR, COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(WfShared));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wavefront_kernel<INTEGRATOR, COUNT>, WF_THREADS,
                                                  sizeof(WfShared));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    cached = per_sm > 0 && sms > 0 ? per_sm * sms : 148;
  }
  const uint32_t needed = (n_pixels + WF_SLOTS - 1) / WF_SLOTS;
  return (int)(needed < (uint32_t)cached ? needed : (uint32_t)cached);
}

template <bool COUNT>
static void launch_render(vimg_scene* s, const RenderArgs& a, uint32_t integrator, float* out, cudaStream_t st) {
  if (integrator == VIMG_INTEGRATOR_MATERIAL || integrator == VIMG_INTEGRATOR_MIS) {
    WfArgs w{};
    w.samples = a.samples;
    w.depth = a.depth;
    w.rank = a.rank;
    w.n_ranks = a.n_ranks;
    w.tiles_x = a.tiles_x;
    w.tiles_y = a.tiles_y;
    w.single_x = a.single_x;
    w.single_y = a.single_y;
    w.n_local_pixels = a.single_x >= 0 ? 1u : a.n_local_tiles * 64u;
    if (w.n_local_pixels == 0) return;
    cudaMemsetAsync(s->d_next_pixel, 0, sizeof(unsigned int), st);
    if (integrator == VIMG_INTEGRATOR_MIS) {
      const int g = wavefront_grid<VIMG_INTEGRATOR_MIS, COUNT>(s, w.n_local_pixels);
      wavefront_kernel<VIMG_INTEGRATOR_MIS, COUNT><<<g, WF_THREADS, sizeof(WfShared), st>>>(s->d, w, out, s->d_counters, s->d_next_pixel);
    } else {
      const int g = wavefront_grid<VIMG_INTEGRATOR_MATERIAL, COUNT>(s, w.n_local_pixels);
      wavefront_kernel<VIMG_INTEGRATOR_MATERIAL, COUNT><<<g, WF_THREADS, sizeof(WfShared), st>>>(s->d, w, out, s->d_counters, s->d_next_pixel);
    }
    return;
  }
  const uint32_t blocks = a.single_x >= 0 ? 1 : (a.n_local_tiles + 1) / 2;
  if (blocks == 0) return;
  if (integrator == VIMG_INTEGRATOR_S_NORMAL)
    render_kernel<VIMG_INTEGRATOR_S_NORMAL, COUNT><<<blocks, 128, 0, st>>>(s->d, a, out, s->d_counters);
  else
    render_kernel<VIMG_INTEGRATOR_G_NORMAL, COUNT><<<blocks, 128, 0, st>>>(s->d, a, out, s->d_counters);
}

static int check_params(const vimg_scene* s, const vimg_render_params* p) {
  if (!s || !p) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument");
  if (p->integrator > VIMG_INTEGRATOR_MIS) return fail(VIMG_ERR_INVALID_ARGUMENT, "unknown integrator");
  if (p->samples == 0) return fail(VIMG_ERR_INVALID_ARGUMENT, "samples must be > 0");
  if (p->n_ranks > 1 && p->rank >= p->n_ranks) return fail(VIMG_ERR_INVALID_ARGUMENT, "rank >= n_ranks");
  if ((p->integrator == VIMG_INTEGRATOR_MIS) && s->d.n_lights == 0) {
    // GroupOfEmitters::sample with an empty list indexes element -1 on the CPU; refuse instead
    return fail(VIMG_ERR_INVALID_ARGUMENT, "mis integrator needs at least one light");
  }
  return VIMG_OK;
}

extern "C" {

int vimg_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* vimg_last_error(void) { return g_last_error.c_str(); }

void vimg_scene_destroy(vimg_scene* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  for (void* p : s->allocs) cudaFree(p);
  if (s->d_frame) cudaFree(s->d_frame);
  if (s->d_counters) cudaFree(s->d_counters);
  if (s->d_next_pixel) cudaFree(s->d_next_pixel);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  delete s;
}

uint64_t vimg_scene_device_bytes(const vimg_scene* s) { return s ? s->device_bytes : 0; }

int vimg_scene_create(const vimg_scene_desc* desc, int device_id, vimg_scene** out_scene) {
  if (!desc || !out_scene) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument");
  *out_scene = nullptr;
  if (desc->abi_version != VIMG_ABI_VERSION) return fail(VIMG_ERR_INVALID_ARGUMENT, "vimg_scene_desc.abi_version mismatch");
  int n = vimg_cuda_device_count();
  if (n == 0) return fail(VIMG_ERR_NO_DEVICE, "no CUDA device visible: the cuda render device has no CPU fallback");
  if (device_id < 0 || device_id >= n) return fail(VIMG_ERR_INVALID_ARGUMENT, "device_id out of range");
  CU_TRY(cudaSetDevice(device_id));
  vimg_scene* s = new vimg_scene();
  s->device = device_id;
  int rc = build_scene(desc, s);
  if (rc == VIMG_OK) {
    cudaError_t e = cudaMalloc(&s->d_frame, (size_t)s->W * s->H * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_counters, sizeof(DCounters));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_next_pixel, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
    if (e != cudaSuccess) rc = fail(VIMG_ERR_CUDA, std::string("scene buffers: ") + cudaGetErrorString(e));
  }
  if (rc != VIMG_OK) {
    std::string keep = g_last_error;
    vimg_scene_destroy(s);
    g_last_error = keep;
    return rc;
  }
  *out_scene = s;
  return VIMG_OK;
}

int vimg_render_device(vimg_scene* s, const vimg_render_params* p, float** out_dev_rgb, void* stream,
                       vimg_render_stats* stats) {
  int rc = check_params(s, p);
  if (rc != VIMG_OK) return rc;
  if (!out_dev_rgb) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU_TRY(cudaSetDevice(s->device));
  const RenderArgs a = make_args(s, p->samples, p->depth, p->n_ranks > 1 ? p->rank : 0, p->n_ranks);
  CU_TRY(cudaMemsetAsync(s->d_counters, 0, sizeof(DCounters), st));
  if (a.n_ranks > 1) CU_TRY(cudaMemsetAsync(s->d_frame, 0, (size_t)s->W * s->H * 3 * sizeof(float), st));
  CU_TRY(cudaEventRecord(s->ev0, st));
  if (p->flags & VIMG_FLAG_COUNTERS)
    launch_render<true>(s, a, p->integrator, s->d_frame, st);
  else
    launch_render<false>(s, a, p->integrator, s->d_frame, st);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaEventRecord(s->ev1, st));
  CU_TRY(cudaStreamSynchronize(st));
  *out_dev_rgb = s->d_frame;
  if (stats) {
    DCounters c;
    CU_TRY(cudaMemcpy(&c, s->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
    std::memset(stats, 0, sizeof(*stats));
    stats->rays_closest = c.rays_closest;
    stats->rays_shadow = c.rays_shadow;
    stats->inner_nodes = c.inner_nodes;
    stats->leaf_nodes = c.leaf_nodes;
    stats->tri_tests = c.tri_tests;
    stats->sphere_tests = c.sphere_tests;
    stats->nan_samples = c.nan_samples;
    stats->samples = c.samples;
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    stats->kernel_ms = ms;
    stats->total_ms = ms;
    stats->kernel_launches = 1;
#ifdef WF_STAGE_CLOCKS
    fprintf(stderr, "[wf stage cycles, summed over warps] shadow %llu regen %llu closest %llu resolve+bucket %llu "
            "barrier1 %llu shade %llu barrier2 %llu\n", c.stage_clk[0], c.stage_clk[1], c.stage_clk[2],
            c.stage_clk[3], c.stage_clk[4], c.stage_clk[5], c.stage_clk[6]);
#endif
  }
  return VIMG_OK;
}

int vimg_render(vimg_scene* s, const vimg_render_params* p, float* out_rgb, vimg_render_stats* stats) {
  if (!out_rgb) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument");
  float* dev = nullptr;
  cudaEvent_t t0, t1;
  int rc = check_params(s, p);
  if (rc != VIMG_OK) return rc;
  CU_TRY(cudaSetDevice(s->device));
  CU_TRY(cudaEventCreate(&t0));
  CU_TRY(cudaEventCreate(&t1));
  cudaEventRecord(t0, 0);
  rc = vimg_render_device(s, p, &dev, nullptr, stats);
  if (rc == VIMG_OK) {
    cudaError_t e = cudaMemcpy(out_rgb, dev, (size_t)s->W * s->H * 3 * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(VIMG_ERR_CUDA, std::string("frame D2H: ") + cudaGetErrorString(e));
  }
  cudaEventRecord(t1, 0);
  cudaEventSynchronize(t1);
  if (rc == VIMG_OK && stats) {
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    stats->total_ms = ms;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  return rc;
}

int vimg_trace_pixel(vimg_scene* s, const vimg_render_params* p, int x, int y, float out_rgb[3]) {
  int rc = check_params(s, p);
  if (rc != VIMG_OK) return rc;
  if (!out_rgb || x < 0 || y < 0 || x >= s->W || y >= s->H) return fail(VIMG_ERR_INVALID_ARGUMENT, "pixel out of range");
  CU_TRY(cudaSetDevice(s->device));
  RenderArgs a = make_args(s, p->samples, p->depth, 0, 1);
  a.single_x = x;
  a.single_y = y;
  float* d_px = nullptr;
  CU_TRY(cudaMalloc(&d_px, 3 * sizeof(float)));
  CU_TRY(cudaMemset(s->d_counters, 0, sizeof(DCounters)));
  launch_render<false>(s, a, p->integrator, d_px, 0);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out_rgb, d_px, 3 * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d_px);
  if (e != cudaSuccess) return fail(VIMG_ERR_CUDA, std::string("trace_pixel: ") + cudaGetErrorString(e));
  return VIMG_OK;
}

int vimg_trace_ids(vimg_scene* s, uint32_t samples, int32_t* out_prim_id, float* out_t) {
  if (!s || !out_prim_id || samples == 0) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument / zero samples");
  CU_TRY(cudaSetDevice(s->device));
  const size_t n = (size_t)s->W * s->H * samples;
  int32_t* d_ids = nullptr;
  float* d_t = nullptr;
  CU_TRY(cudaMalloc(&d_ids, n * sizeof(int32_t)));
  if (out_t) {
    cudaError_t e = cudaMalloc(&d_t, n * sizeof(float));
    if (e != cudaSuccess) {
      cudaFree(d_ids);
      return fail(VIMG_ERR_OUT_OF_MEMORY, "trace_ids: t buffer");
    }
  }
  const RenderArgs a = make_args(s, samples, 0, 0, 1);
  trace_ids_kernel<<<(a.n_local_tiles + 1) / 2, 128>>>(s->d, a, d_ids, d_t);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out_prim_id, d_ids, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && out_t) e = cudaMemcpy(out_t, d_t, n * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d_ids);
  if (d_t) cudaFree(d_t);
  if (e != cudaSuccess) return fail(VIMG_ERR_CUDA, std::string("trace_ids: ") + cudaGetErrorString(e));
  return VIMG_OK;
}

}  // extern "C"
