// vimg_scene_io.cpp -- relocatable single-file form of a vimg_scene_desc.
//
// A flattened scene (what the adapter hands to vimg_scene_create) can be cached on
// disk: a loaded v-img scene + host-built BVH is deterministic, so the render device can be
// fed without re-running loaders.  The repo's golden scenes (tests/golden/scenes/*.vscene)
// are in this format.  Layout: magic, ABI version, then every array as {u64 count, bytes}
// in the fixed order below; all little-endian, no pointers.
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vimg_cuda.h"

namespace {

constexpr char kMagic[8] = {'V', 'I', 'M', 'G', 'S', 'C', 'N', '1'};

struct Writer {
  FILE* f;
  bool ok = true;
  void raw(const void* p, size_t n) {
    if (ok && n && fwrite(p, 1, n, f) != n) ok = false;
  }
  template <typename T> void pod(const T& v) { raw(&v, sizeof(T)); }
  template <typename T> void arr(const T* p, uint64_t n) {
    pod(n);
    raw(p, n * sizeof(T));
  }
};

struct Reader {
  FILE* f;
  bool ok = true;
  void raw(void* p, size_t n) {
    if (ok && n && fread(p, 1, n, f) != n) ok = false;
  }
  template <typename T> void pod(T& v) { raw(&v, sizeof(T)); }
  template <typename T> void arr(std::vector<T>& v) {
    uint64_t n = 0;
    pod(n);
    if (!ok || n > (1ull << 36) / sizeof(T)) {
      ok = false;
      return;
    }
    v.resize(n);
    raw(v.data(), n * sizeof(T));
  }
};

// owns everything a loaded descriptor points to; `desc` must stay the first member
struct LoadedScene {
  vimg_scene_desc desc{};
  std::vector<uint32_t> nodes;
  std::vector<float> bb;
  std::vector<uint64_t> obj_indices;
  std::vector<vimg_prim> prims;
  std::vector<vimg_sphere> spheres;
  std::vector<vimg_mesh> meshes;
  std::vector<std::vector<uint32_t>> mesh_indices;
  std::vector<std::vector<float>> mesh_vertices, mesh_normals;
  std::vector<std::vector<std::vector<float>>> mesh_uvs;
  std::vector<std::vector<const float*>> mesh_uv_ptrs;
  std::vector<vimg_material> materials;
  std::vector<vimg_texture> textures;
  std::vector<std::vector<std::vector<float>>> tex_levels;
  std::vector<std::vector<const float*>> tex_level_ptrs;
  std::vector<vimg_texture_rg> textures_rg;
  std::vector<std::vector<float>> rg_data;
  std::vector<int32_t> lights;
  std::vector<float> row_cdf, col_cdfs;
};

thread_local std::string g_io_error;

}  // namespace

extern "C" {

int vimg_scene_desc_save(const vimg_scene_desc* d, const char* path) {
  if (!d || !path) return VIMG_ERR_INVALID_ARGUMENT;
  // the file format carries host-built mip levels and CDFs: a descriptor that leaves them to the device
  // (NULL levels 1.. / NULL CDFs, see vimg_scene_create) cannot be written
  for (uint32_t t = 0; t < d->n_textures; ++t)
    if (d->textures[t].type == VIMG_TEX_IMAGE)
      for (uint32_t l = 0; l < d->textures[t].n_levels; ++l)
        if (!d->textures[t].levels || !d->textures[t].levels[l]) return VIMG_ERR_INVALID_ARGUMENT;
  if (d->background.type == VIMG_BG_ENVMAP && (!d->background.row_cdf || !d->background.col_cdfs))
    return VIMG_ERR_INVALID_ARGUMENT;
  FILE* f = fopen(path, "wb");
  if (!f) return VIMG_ERR_IO;
  Writer w{f};
  w.raw(kMagic, 8);
  w.pod(d->abi_version);
  w.pod(d->camera);
  w.arr(d->bvh.nodes, d->bvh.n_nodes * 2);
  w.arr(d->bvh.bb_mins_maxes, d->bvh.n_bb_slots * 3);
  w.arr(d->bvh.obj_indices, d->bvh.n_obj_indices);
  w.pod(d->bvh.max_depth);
  w.arr(d->prims, d->n_prims);
  w.arr(d->spheres, (uint64_t)d->n_spheres);
  w.pod(d->n_meshes);
  for (uint32_t m = 0; m < d->n_meshes; ++m) {
    const vimg_mesh& hm = d->meshes[m];
    w.arr(hm.indices, hm.n_tris * 3);
    w.arr(hm.vertices, hm.n_verts * 3);
    w.arr(hm.normals, hm.normals ? hm.n_verts * 3 : 0);
    w.pod(hm.n_uv_sets);
    for (uint32_t k = 0; k < hm.n_uv_sets; ++k) w.arr(hm.texcoords[k], hm.n_verts * 2);
    w.pod(hm.color_tex_uv);
    w.pod(hm.normal_tex_uv);
    w.pod(hm.metallic_roughness_tex_uv);
    w.pod(hm.material);
  }
  w.arr(d->materials, (uint64_t)d->n_materials);
  w.pod(d->n_textures);
  for (uint32_t t = 0; t < d->n_textures; ++t) {
    vimg_texture ht = d->textures[t];
    const float* const* levels = ht.levels;
    ht.levels = nullptr;
    w.pod(ht);
    if (ht.type == VIMG_TEX_IMAGE)
      for (uint32_t l = 0; l < ht.n_levels; ++l) {
        uint64_t lw = (ht.width >> l) ? (ht.width >> l) : 1, lh = (ht.height >> l) ? (ht.height >> l) : 1;
        w.arr(levels[l], lw * lh * 3);
      }
  }
  w.pod(d->n_textures_rg);
  for (uint32_t t = 0; t < d->n_textures_rg; ++t) {
    vimg_texture_rg ht = d->textures_rg[t];
    const float* data = ht.data;
    ht.data = nullptr;
    w.pod(ht);
    w.arr(data, (uint64_t)ht.width * ht.height * 2);
  }
  w.arr(d->lights, (uint64_t)d->n_lights);
  vimg_background bg = d->background;
  const float *row = bg.row_cdf, *col = bg.col_cdfs;
  bg.row_cdf = bg.col_cdfs = nullptr;
  w.pod(bg);
  const bool env = bg.type == VIMG_BG_ENVMAP;
  w.arr(row, env ? (uint64_t)bg.cdf_height + 1 : 0);
  w.arr(col, env ? (uint64_t)bg.cdf_height * (bg.cdf_width + 1) : 0);
  bool ok = w.ok;
  if (fclose(f) != 0) ok = false;
  return ok ? VIMG_OK : VIMG_ERR_IO;
}

int vimg_scene_desc_load(const char* path, vimg_scene_desc** out) {
  if (!path || !out) return VIMG_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return VIMG_ERR_IO;
  auto L = std::make_unique<LoadedScene>();
  vimg_scene_desc& d = L->desc;
  Reader r{f};
  char magic[8];
  r.raw(magic, 8);
  if (!r.ok || std::memcmp(magic, kMagic, 8) != 0) {
    fclose(f);
    return VIMG_ERR_IO;
  }
  r.pod(d.abi_version);
  r.pod(d.camera);
  r.arr(L->nodes);
  r.arr(L->bb);
  r.arr(L->obj_indices);
  r.pod(d.bvh.max_depth);
  r.arr(L->prims);
  r.arr(L->spheres);
  uint32_t n_meshes = 0;
  r.pod(n_meshes);
  if (!r.ok || n_meshes > (1u << 24)) {
    fclose(f);
    return VIMG_ERR_IO;
  }
  L->meshes.resize(n_meshes);
  L->mesh_indices.resize(n_meshes);
  L->mesh_vertices.resize(n_meshes);
  L->mesh_normals.resize(n_meshes);
  L->mesh_uvs.resize(n_meshes);
  L->mesh_uv_ptrs.resize(n_meshes);
  for (uint32_t m = 0; m < n_meshes && r.ok; ++m) {
    vimg_mesh& hm = L->meshes[m];
    std::memset(&hm, 0, sizeof(hm));
    r.arr(L->mesh_indices[m]);
    r.arr(L->mesh_vertices[m]);
    r.arr(L->mesh_normals[m]);
    r.pod(hm.n_uv_sets);
    if (!r.ok || hm.n_uv_sets > 255) {
      r.ok = false;
      break;
    }
    L->mesh_uvs[m].resize(hm.n_uv_sets);
    for (uint32_t k = 0; k < hm.n_uv_sets; ++k) r.arr(L->mesh_uvs[m][k]);
    r.pod(hm.color_tex_uv);
    r.pod(hm.normal_tex_uv);
    r.pod(hm.metallic_roughness_tex_uv);
    r.pod(hm.material);
  }
  r.arr(L->materials);
  uint32_t n_tex = 0;
  r.pod(n_tex);
  if (!r.ok || n_tex > (1u << 24)) {
    fclose(f);
    return VIMG_ERR_IO;
  }
  L->textures.resize(n_tex);
  L->tex_levels.resize(n_tex);
  L->tex_level_ptrs.resize(n_tex);
  for (uint32_t t = 0; t < n_tex && r.ok; ++t) {
    vimg_texture& ht = L->textures[t];
    r.pod(ht);
    ht.levels = nullptr;
    if (r.ok && ht.type == VIMG_TEX_IMAGE) {
      if (ht.n_levels > 16) {
        r.ok = false;
        break;
      }
      L->tex_levels[t].resize(ht.n_levels);
      for (uint32_t l = 0; l < ht.n_levels; ++l) r.arr(L->tex_levels[t][l]);
    }
  }
  uint32_t n_rg = 0;
  r.pod(n_rg);
  if (!r.ok || n_rg > (1u << 24)) {
    fclose(f);
    return VIMG_ERR_IO;
  }
  L->textures_rg.resize(n_rg);
  L->rg_data.resize(n_rg);
  for (uint32_t t = 0; t < n_rg && r.ok; ++t) {
    r.pod(L->textures_rg[t]);
    r.arr(L->rg_data[t]);
  }
  r.arr(L->lights);
  r.pod(d.background);
  r.arr(L->row_cdf);
  r.arr(L->col_cdfs);
  fclose(f);
  if (!r.ok) return VIMG_ERR_IO;

  // A truncated or hand-edited file must not reach vimg_scene_create, which sizes its reads from the
  // dimensions (3*w*h floats per mip level, 2*n_verts per uv set, (h+1) and h*(w+1) CDF entries, ...):
  // every array's stated length has to be what the dimensions next to it imply.
  {
    bool ok = L->nodes.size() % 2 == 0 && L->bb.size() % 3 == 0;
    const uint64_t n_nodes = L->nodes.size() / 2;
    if (n_nodes && L->bb.size() / 3 < 2 * n_nodes + 2) ok = false;
    for (uint32_t m = 0; m < n_meshes && ok; ++m) {
      const size_t nv3 = L->mesh_vertices[m].size();
      ok = L->mesh_indices[m].size() % 3 == 0 && nv3 % 3 == 0
           && (L->mesh_normals[m].empty() || L->mesh_normals[m].size() == nv3);
      for (const auto& set : L->mesh_uvs[m]) ok = ok && set.size() == nv3 / 3 * 2;
    }
    for (uint32_t t = 0; t < n_tex && ok; ++t) {
      const vimg_texture& ht = L->textures[t];
      if (ht.type != VIMG_TEX_IMAGE) continue;
      ok = ht.n_levels >= 1 && ht.width > 0 && ht.height > 0;
      for (uint32_t l = 0; l < ht.n_levels && ok; ++l) {
        const uint64_t lw = (ht.width >> l) ? (ht.width >> l) : 1, lh = (ht.height >> l) ? (ht.height >> l) : 1;
        ok = L->tex_levels[t][l].size() == 3 * lw * lh;
      }
    }
    for (uint32_t t = 0; t < n_rg && ok; ++t) {
      const vimg_texture_rg& ht = L->textures_rg[t];
      ok = ht.width > 0 && ht.height > 0 && L->rg_data[t].size() == 2ull * ht.width * ht.height;
    }
    if (ok && d.background.type == VIMG_BG_ENVMAP) {
      const uint64_t w = d.background.cdf_width, h = d.background.cdf_height;
      ok = (L->row_cdf.empty() && L->col_cdfs.empty()) || (L->row_cdf.size() == h + 1 && L->col_cdfs.size() == h * (w + 1));
    }
    if (!ok) return VIMG_ERR_IO;
  }

  // wire pointers
  d.bvh.nodes = L->nodes.data();
  d.bvh.n_nodes = L->nodes.size() / 2;
  d.bvh.bb_mins_maxes = L->bb.data();
  d.bvh.n_bb_slots = L->bb.size() / 3;
  d.bvh.obj_indices = L->obj_indices.data();
  d.bvh.n_obj_indices = L->obj_indices.size();
  d.prims = L->prims.data();
  d.n_prims = L->prims.size();
  d.spheres = L->spheres.data();
  d.n_spheres = (uint32_t)L->spheres.size();
  for (uint32_t m = 0; m < n_meshes; ++m) {
    vimg_mesh& hm = L->meshes[m];
    hm.indices = L->mesh_indices[m].data();
    hm.n_tris = L->mesh_indices[m].size() / 3;
    hm.vertices = L->mesh_vertices[m].data();
    hm.n_verts = L->mesh_vertices[m].size() / 3;
    hm.normals = L->mesh_normals[m].empty() ? nullptr : L->mesh_normals[m].data();
    for (auto& set : L->mesh_uvs[m]) L->mesh_uv_ptrs[m].push_back(set.data());
    hm.texcoords = L->mesh_uv_ptrs[m].data();
  }
  d.meshes = L->meshes.data();
  d.n_meshes = n_meshes;
  d.materials = L->materials.data();
  d.n_materials = (uint32_t)L->materials.size();
  for (uint32_t t = 0; t < n_tex; ++t) {
    for (auto& lvl : L->tex_levels[t]) L->tex_level_ptrs[t].push_back(lvl.data());
    L->textures[t].levels = L->tex_level_ptrs[t].empty() ? nullptr : L->tex_level_ptrs[t].data();
  }
  d.textures = L->textures.data();
  d.n_textures = n_tex;
  for (uint32_t t = 0; t < n_rg; ++t) L->textures_rg[t].data = L->rg_data[t].data();
  d.textures_rg = L->textures_rg.data();
  d.n_textures_rg = n_rg;
  d.lights = L->lights.data();
  d.n_lights = (uint32_t)L->lights.size();
  d.background.row_cdf = L->row_cdf.empty() ? nullptr : L->row_cdf.data();
  d.background.col_cdfs = L->col_cdfs.empty() ? nullptr : L->col_cdfs.data();
  *out = &L.release()->desc;
  return VIMG_OK;
}

void vimg_scene_desc_free(vimg_scene_desc* desc) {
  if (desc) delete reinterpret_cast<LoadedScene*>(desc);
}

}  // extern "C"
