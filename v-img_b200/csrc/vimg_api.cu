// vimg_api.cu -- C ABI of libvimg_cuda.so (include/vimg_cuda.h): scene upload, the
// GPU layout build, kernel launches.  No CPU fallback exists: every compute entry point
// needs a CUDA device and fails loudly otherwise.
#include <cuda_runtime.h>

#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <numbers>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "vimg_kernels.cuh"
#include "vimg_stream.cuh"
#include "vimg_wide.cuh"
#include "vimg_post.cuh"
#include "vimg_prep.cuh"

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

// One slot pool of the queue wavefront: its state in HBM, its streams and its host-side
// bookkeeping.  Pools are independent pipelines over disjoint slots that pull pixels from one
// shared counter; running two of them lets one pool's TRACE launch (issue bound) overlap the
// other's SHADE / END launches (latency bound).
constexpr int QW_MAX_POOLS = 4;
constexpr uint32_t QF_FLAT_MAX_HOST = 64;  // = QF_FLAT_MAX of the tolerance build (vimg_queues.cuh)
struct QwPool {
  QwState st{};
  void* mem = nullptr;
  uint32_t capacity = 0;
  QwCounters* h_cnt = nullptr;  // pinned: queue-counter snapshots (termination test)
  uint32_t* h_next_pixel = nullptr;  // pinned: shared pixel counter at the same moments
  cudaStream_t main = nullptr, side[2] = {nullptr, nullptr};  // shade launches run beside END
  cudaEvent_t snap[2] = {nullptr, nullptr}, fork = nullptr, join[2] = {nullptr, nullptr}, done = nullptr;
  bool finished = false;
  int trace_grid = 1, item_grid = 1;  // current launch sizes (shrink as the pool drains)
  // one batch of iterations captured as a CUDA graph (launch sizes it was captured with)
  cudaGraphExec_t batch_exec = nullptr;
  int exec_trace_grid = 0, exec_item_grid = 0;
  uint32_t graph_launches = 0;
  void drop_graph() {
    if (batch_exec) cudaGraphExecDestroy(batch_exec);
    batch_exec = nullptr;
    exec_trace_grid = exec_item_grid = 0;
  }
  void destroy() {
    drop_graph();
    if (mem) cudaFree(mem);
    if (h_cnt) cudaFreeHost(h_cnt);
    if (h_next_pixel) cudaFreeHost(h_next_pixel);
    if (main) cudaStreamDestroy(main);
    for (int k = 0; k < 2; ++k) {
      if (side[k]) cudaStreamDestroy(side[k]);
      if (snap[k]) cudaEventDestroy(snap[k]);
      if (join[k]) cudaEventDestroy(join[k]);
    }
    if (fork) cudaEventDestroy(fork);
    if (done) cudaEventDestroy(done);
    *this = QwPool{};
  }
};

// Render workspace of the queue wavefront: slot pools, shared unit counter, partial sums.
// It is independent of the scene (sized by slots and frame), so a scene takes an idle one of
// its device from a process-wide cache at its first render and gives it back when it is
// destroyed: a caller that creates a scene per frame (bench.py's end-to-end step, the
// reference's main.cpp) does not pay cudaMalloc / cudaFree of ~1 GB inside every render.
// slot lines and queues of one pool of the streamed pipeline (vimg_stream.cuh); streams and events are the QwPool's
struct QvPool {
  QvState st{};
  void* mem = nullptr;
  uint32_t capacity = 0;
  QvCounters* h_cnt = nullptr;  // pinned snapshots
  void destroy() {
    if (mem) cudaFree(mem);
    if (h_cnt) cudaFreeHost(h_cnt);
    *this = QvPool{};
  }
};

struct QwWorkspace {
  QwPool qw[QW_MAX_POOLS];
  QvPool qv[QW_MAX_POOLS];
  uint32_t* d_next_pixel = nullptr;
  float* d_partial = nullptr;  // spp_streams > 1: per-range partial sums, W*H*streams*3 floats
  size_t partial_floats = 0;
  void destroy() {
    for (QwPool& p : qw) p.destroy();
    for (QvPool& p : qv) p.destroy();
    if (d_next_pixel) cudaFree(d_next_pixel);
    if (d_partial) cudaFree(d_partial);
    d_next_pixel = nullptr;
    d_partial = nullptr;
    partial_floats = 0;
  }
};
static std::mutex g_ws_mu;
static std::map<int, QwWorkspace*> g_ws_idle;  // device -> one idle workspace

// One idle scene arena per device rides along with the workspace cache: cudaMalloc / cudaFree
// cost milliseconds each on this driver (measured: 4-19 ms to create, 7-12 ms to destroy a 17 MB
// scene), which a scene-per-frame caller would pay on every frame.
struct IdleArena {
  void* ptr = nullptr;
  size_t bytes = 0;
};
static std::map<int, IdleArena> g_arena_idle;

static void* arena_acquire(int device, size_t bytes, size_t* got_bytes) {
  {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    IdleArena& a = g_arena_idle[device];
    if (a.ptr && a.bytes >= bytes && a.bytes <= 2 * bytes + (64u << 20)) {
      void* p = a.ptr;
      *got_bytes = a.bytes;
      a = IdleArena{};
      return p;
    }
  }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  *got_bytes = bytes;
  return p;
}
static void arena_release(int device, void* ptr, size_t bytes) {
  if (!ptr) return;
  void* drop = ptr;
  {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    IdleArena& a = g_arena_idle[device];
    if (!a.ptr || a.bytes < bytes) {  // keep the larger one
      drop = a.ptr;
      a.ptr = ptr;
      a.bytes = bytes;
    }
  }
  if (drop) cudaFree(drop);
}

static QwWorkspace* ws_acquire(int device) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  auto it = g_ws_idle.find(device);
  if (it != g_ws_idle.end() && it->second) {
    QwWorkspace* w = it->second;
    it->second = nullptr;
    return w;
  }
  return new QwWorkspace();
}
static void ws_release(int device, QwWorkspace* w) {
  if (!w) return;
  {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    QwWorkspace*& slot = g_ws_idle[device];
    if (!slot) {
      slot = w;
      return;
    }
  }
  w->destroy();
  delete w;
}

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
  // queue wavefront (vimg_queues.cuh): slot pools, class mask of the scene's materials
  uint8_t* d_rgb8 = nullptr;      // vimg_render_srgb8: quantised frame
  uint32_t* d_post_max = nullptr;  // ... and the reinhard tonemapper's largest luminance (float bits)
  void* d_flat = nullptr;  // tolerance build: flat re-layout of a small scene (vimg_flat.cuh), built at the first fast render
  DWideNode* d_wide = nullptr;  // tolerance build: 4-wide BVH of a large scene (vimg_wide.cuh), built at the first fast render
  bool wide_ok = false;         // every leaf fits a wide leaf reference
  bool lean = false;            // no image texture, normal map, metallic-roughness texture, env map, general uv walk (vimg_lean.cu)
  uint32_t n_pairs = 0;
  QwWorkspace* ws = nullptr;  // taken at the first material / mis render
  uint32_t class_mask = 0;
  int sm_count = 0;
  uint32_t last_launches = 0;
  double stage_ms[VIMG_STAGE_COUNT] = {0};
  uint32_t stage_launches[VIMG_STAGE_COUNT] = {0};

  // Scene arrays live in ONE device allocation (one cudaMalloc at creation, one cudaFree at
  // destruction instead of ~25 each: a caller that builds a scene per frame pays for every one of
  // them).  upload() only records what to copy; commit_uploads() sizes the arena (256-byte
  // aligned regions, the frame and the small per-scene buffers included), allocates and copies.
  struct Pending {
    const void* host;
    size_t bytes;
    const void** out;
  };
  std::vector<Pending> pending;
  size_t arena_bytes = 0;
  template <typename T> int upload(const T* host, size_t n, const T** out) {
    *out = nullptr;
    if (n == 0) return VIMG_OK;
    pending.push_back({host, n * sizeof(T), reinterpret_cast<const void**>(out)});
    return VIMG_OK;
  }
  template <typename T> int upload(const std::vector<T>& v, const T** out) {
    return upload(v.data(), v.size(), out);
  }
  int commit_uploads() {
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t total = 0;
    for (const Pending& p : pending) total += align(p.bytes);
    const size_t o_frame = total;
    total += align((size_t)W * H * 3 * sizeof(float));
    const size_t o_counters = total;
    total += align(sizeof(DCounters));
    const size_t o_next = total;
    total += align(sizeof(unsigned int));
    void* arena = arena_acquire(device, total, &arena_bytes);
    if (!arena) return fail(VIMG_ERR_OUT_OF_MEMORY, "scene arena");
    allocs.push_back(arena);
    char* b = static_cast<char*>(arena);
    size_t off = 0;
    for (const Pending& p : pending) {
      CU_TRY(cudaMemcpyAsync(b + off, p.host, p.bytes, cudaMemcpyHostToDevice, 0));
      *p.out = b + off;
      device_bytes += p.bytes;
      off += align(p.bytes);
    }
    CU_TRY(cudaStreamSynchronize(0));  // the host arrays are the caller's / build_scene's locals
    pending.clear();
    d_frame = reinterpret_cast<float*>(b + o_frame);
    d_counters = reinterpret_cast<DCounters*>(b + o_counters);
    d_next_pixel = reinterpret_cast<unsigned int*>(b + o_next);
    return VIMG_OK;
  }
};

extern "C" {  // defined next to the C ABI entry points that share them
static void launch_mip_levels(float* d_all, uint32_t width, uint32_t height, uint32_t n_levels, uint32_t wrap_u,
                              uint32_t wrap_v);
static cudaError_t launch_envmap_cdf(const float* d_img, uint32_t width, uint32_t height, float* d_col, float* d_row);
}

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
  d.n_leaf_recs = (uint32_t)b.n_obj_indices;
  s->n_pairs = b.n_nodes ? (uint32_t)((b.n_nodes - 1) / 2) : 0;
  s->wide_ok = s->n_pairs > 0 && b.n_obj_indices < WIDE_MAX_LEAF_FIRST;
  for (uint64_t n = 0; n < b.n_nodes && s->wide_ok; ++n)
    if (b.nodes[2 * n + 1] > WIDE_MAX_LEAF_COUNT) s->wide_ok = false;
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
    // The device's traversal stacks hold max_depth + 2 entries per traversal and are not bounds checked:
    // walk the tree once from the root and refuse a descriptor whose nodes are not a tree (a node reached
    // twice: shared child or cycle) or whose real depth exceeds the stated max_depth.
    {
      std::vector<uint8_t> seen(b.n_nodes, 0);
      std::vector<std::pair<uint32_t, uint32_t>> todo;  // node, depth (root = 0)
      todo.emplace_back(0u, 0u);
      seen[0] = 1;
      uint32_t deepest = 0;
      while (!todo.empty()) {
        const auto [n, depth] = todo.back();
        todo.pop_back();
        deepest = std::max(deepest, depth);
        if (b.nodes[2 * (uint64_t)n + 1] != 0) continue;  // leaf
        const uint32_t first = b.nodes[2 * (uint64_t)n];
        for (uint32_t ch = first; ch <= first + 1; ++ch) {
          if (seen[ch]) return fail(VIMG_ERR_INVALID_ARGUMENT, "BVH nodes do not form a tree (node reached twice)");
          seen[ch] = 1;
          todo.emplace_back(ch, depth + 1);
        }
      }
      if (deepest > b.max_depth) return fail(VIMG_ERR_INVALID_ARGUMENT, "BVH is deeper than bvh.max_depth");
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

  // material class of every leaf record (queue wavefront: routing without pointer chasing)
  std::vector<uint8_t> leaf_class(b.n_obj_indices);
  for (uint64_t i = 0; i < b.n_obj_indices; ++i) {
    const vimg_prim& pr = desc->prims[b.obj_indices[i]];
    const uint32_t mat = pr.type == VIMG_PRIM_SPHERE ? desc->spheres[pr.a].material : desc->meshes[pr.a].material;
    if (mat >= desc->n_materials) return fail(VIMG_ERR_INVALID_ARGUMENT, "primitive material index out of range");
    uint8_t cl = WF_CLASS_NONE;
    switch (desc->materials[mat].type) {
      case VIMG_MAT_LAMBERTIAN: cl = WF_CLASS_LAMBERT; break;
      case VIMG_MAT_PRINCIPLED: cl = WF_CLASS_PRINCIPLED; break;
      case VIMG_MAT_DIELECTRIC: cl = WF_CLASS_DIELECTRIC; break;
      case VIMG_MAT_DIFFUSE_LIGHT: cl = WF_CLASS_NONE | VIMG_LEAF_CLASS_EMISSIVE; break;
      default: break;
    }
    leaf_class[i] = cl;
  }

  // ---- textures --------------------------------------------------------------
  std::vector<DTexture> textures(desc->n_textures);
  std::vector<float> texels;
  std::vector<uint32_t> build_mips;  // image textures whose levels 1.. the device builds
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
    dt.half_log2_wh = (double)0.5f * vimg_libm::glibc_log2((double)(uint32_t)(ht.width * ht.height));
    if (ht.type == VIMG_TEX_IMAGE) {
      if (ht.n_levels == 0 || ht.n_levels > 16) return fail(VIMG_ERR_INVALID_ARGUMENT, "image texture needs 1..16 mip levels");
      if (!ht.levels || !ht.levels[0]) return fail(VIMG_ERR_INVALID_ARGUMENT, "image texture without level 0");
      // levels 1.. may be left NULL (all of them): the device builds the chain from level 0 with the
      // kernels of vimg_build_mip_chain once the pool is uploaded (ImageTexture's constructor)
      uint32_t missing = 0;
      for (uint32_t l = 1; l < ht.n_levels; ++l) missing += ht.levels[l] == nullptr;
      if (missing != 0 && missing != ht.n_levels - 1)
        return fail(VIMG_ERR_INVALID_ARGUMENT, "image texture: pass all mip levels or level 0 only");
      if (missing) {
        uint64_t floats = 0;
        if (vimg_mip_chain_levels(ht.width, ht.height, &floats) != ht.n_levels)
          return fail(VIMG_ERR_INVALID_ARGUMENT, "image texture: n_levels is not what ImageTexture's constructor builds");
        build_mips.push_back(t);
      }
      for (uint32_t l = 0; l < ht.n_levels; ++l) {
        const uint64_t w = std::max(ht.width >> l, 1u), h = std::max(ht.height >> l, 1u);
        dt.level_off[l] = (uint32_t)(texels.size() / 3);
        if (ht.levels[l])
          texels.insert(texels.end(), ht.levels[l], ht.levels[l] + 3 * w * h);
        else
          texels.resize(texels.size() + 3 * w * h, 0.f);
      }
    }
  }
  std::vector<DTextureRG> textures_rg(desc->n_textures_rg);
  std::vector<float> rg_texels;
  for (uint32_t t = 0; t < desc->n_textures_rg; ++t) {
    const vimg_texture_rg& ht = desc->textures_rg[t];
    if (ht.width == 0 || ht.height == 0 || !ht.data) return fail(VIMG_ERR_INVALID_ARGUMENT, "RG texture with zero size or no data");
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
  for (uint32_t m = 0; m < desc->n_materials; ++m) {
    switch (desc->materials[m].type) {
      case VIMG_MAT_LAMBERTIAN: s->class_mask |= 1u << WF_CLASS_LAMBERT; break;
      case VIMG_MAT_PRINCIPLED: s->class_mask |= 1u << WF_CLASS_PRINCIPLED; break;
      case VIMG_MAT_DIELECTRIC: s->class_mask |= 1u << WF_CLASS_DIELECTRIC; break;
      default: break;
    }
  }
  for (uint32_t i = 0; i < desc->n_spheres; ++i)
    if (desc->spheres[i].material >= desc->n_materials) return fail(VIMG_ERR_INVALID_ARGUMENT, "sphere material index out of range");
  for (uint32_t i = 0; i < desc->n_lights; ++i)
    if (desc->lights[i] != VIMG_LIGHT_BACKGROUND && (desc->lights[i] < 0 || (uint64_t)desc->lights[i] >= desc->n_prims))
      return fail(VIMG_ERR_INVALID_ARGUMENT, "light index out of range");

  // ---- shade records (per leaf entry) and light records (per light) ---------------
  // the values hit-info and light sampling would gather through prim -> mesh -> indices ->
  // normals / uvs, laid side by side (vimg_types.cuh)
  std::vector<DShadeRec> shade_recs(b.n_obj_indices);
  for (uint64_t i = 0; i < b.n_obj_indices; ++i) {
    const vimg_prim& pr = desc->prims[b.obj_indices[i]];
    DShadeRec& sr = shade_recs[i];
    std::memset(&sr, 0, sizeof(sr));
    if (pr.type == VIMG_PRIM_SPHERE) {
      sr.material = desc->spheres[pr.a].material;
      const vimg_material& m = desc->materials[sr.material];
      if ((m.texture >= 0 && desc->textures[m.texture].type != VIMG_TEX_CONST) || m.texture_rg >= 0)
        sr.flags |= VIMG_SHADE_SPHERE_NEEDS_UV;
      continue;
    }
    const vimg_mesh& hm = desc->meshes[pr.a];
    const DMesh& dm = meshes[pr.a];
    const uint32_t* idx = hm.indices + 3 * (uint64_t)pr.b;
    sr.material = hm.material;
    if (hm.normals) {
      sr.flags |= VIMG_SHADE_HAS_NORMALS;
      for (int k = 0; k < 3; ++k) std::memcpy(sr.n + 3 * k, hm.normals + 3 * (uint64_t)idx[k], 12);
    }
    if (dm.uv_color >= 0) {
      sr.flags |= VIMG_SHADE_HAS_UV;
      const float* tc = hm.texcoords[hm.color_tex_uv];
      for (int k = 0; k < 3; ++k) std::memcpy(sr.uv + 2 * k, tc + 2 * (uint64_t)idx[k], 8);
    }
    // the other two uv selectors: absent, or the colour set itself (interpolating the same
    // arrays gives the same bits); anything else takes the general walk
    if (dm.uv_mr >= 0 && dm.uv_mr != dm.uv_color) sr.flags |= VIMG_SHADE_GENERAL;
    if (dm.uv_normal >= 0) sr.flags |= dm.uv_normal == dm.uv_color ? VIMG_SHADE_NORMAL_UV_SAME : VIMG_SHADE_GENERAL;
  }
  std::vector<DLightRec> light_recs(desc->n_lights);
  for (uint32_t i = 0; i < desc->n_lights; ++i) {
    DLightRec& lr = light_recs[i];
    std::memset(&lr, 0, sizeof(lr));
    if (desc->lights[i] == VIMG_LIGHT_BACKGROUND) {
      lr.type = VIMG_LIGHTREC_BACKGROUND;
      continue;
    }
    const vimg_prim& pr = desc->prims[desc->lights[i]];
    if (pr.type == VIMG_PRIM_SPHERE) {
      if (pr.a >= desc->n_spheres) return fail(VIMG_ERR_INVALID_ARGUMENT, "sphere index out of range");
      const vimg_sphere& sp = desc->spheres[pr.a];
      lr.type = VIMG_LIGHTREC_SPHERE;
      std::memcpy(lr.p, sp.center, 12);
      lr.p[3] = sp.radius;
      lr.material = sp.material;
    } else {
      if (pr.a >= desc->n_meshes || pr.b >= desc->meshes[pr.a].n_tris)
        return fail(VIMG_ERR_INVALID_ARGUMENT, "triangle primitive out of range");
      const vimg_mesh& hm = desc->meshes[pr.a];
      const uint32_t* idx = hm.indices + 3 * (uint64_t)pr.b;
      lr.type = VIMG_LIGHTREC_TRIANGLE;
      lr.material = hm.material;
      lr.has_normals = hm.normals != nullptr;
      for (int k = 0; k < 3; ++k) {
        std::memcpy(lr.p + 3 * k, hm.vertices + 3 * (uint64_t)idx[k], 12);
        if (hm.normals) std::memcpy(lr.n + 3 * k, hm.normals + 3 * (uint64_t)idx[k], 12);
      }
    }
  }

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

  // who reads a ray cone: image colour textures and the env map (vimg_types.cuh)
  d.uses_cones = hb.type == VIMG_BG_ENVMAP ? 1u : 0u;
  for (uint32_t m = 0; m < desc->n_materials; ++m) {
    const int32_t t = desc->materials[m].texture;
    if (t >= 0 && desc->textures[t].type == VIMG_TEX_IMAGE) d.uses_cones = 1u;
  }
  if (const char* e = std::getenv("VIMG_FORCE_CONES")) d.uses_cones = std::atoi(e) ? 1u : d.uses_cones;
  // a scene the lean build of the streamed kernels renders (VIMG_LEAN, vimg_math.cuh): none of the features it leaves out
  s->lean = d.uses_cones == 0 && hb.type == VIMG_BG_CONST && desc->n_textures_rg == 0;
  for (uint32_t t = 0; t < desc->n_textures; ++t)
    if (desc->textures[t].type != VIMG_TEX_CONST && desc->textures[t].type != VIMG_TEX_CHECKER) s->lean = false;
  for (uint32_t m = 0; m < desc->n_materials; ++m)
    if (desc->materials[m].normal_map >= 0 || desc->materials[m].texture_rg >= 0) s->lean = false;
  for (const DShadeRec& sr : shade_recs)
    if (sr.flags & VIMG_SHADE_GENERAL) s->lean = false;

  // ---- upload ------------------------------------------------------------------
  int rc;
#define UP(call) if ((rc = (call)) != VIMG_OK) return rc
  UP(s->upload(pairs, &d.pairs));
  UP(s->upload(leaf, &d.leaf_prims));
  UP(s->upload(leaf_class, &d.leaf_class));
  UP(s->upload(shade_recs, &d.shade_recs));
  UP(s->upload(light_recs, &d.light_recs));
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
  // env-map CDFs left NULL: built on the device from the image's level 0 (ArraySampling2D's constructor)
  const bool build_cdfs = hb.type == VIMG_BG_ENVMAP && (!hb.row_cdf || !hb.col_cdfs);
  std::vector<float> zero_cdfs;
  if (hb.type == VIMG_BG_ENVMAP) {
    if (build_cdfs) {
      const vimg_texture& img = desc->textures[hb.image];
      if ((hb.cdf_width && hb.cdf_width != img.width) || (hb.cdf_height && hb.cdf_height != img.height))
        return fail(VIMG_ERR_INVALID_ARGUMENT, "env map: cdf size differs from the image size");
      d.bg.cdf_w = img.width;
      d.bg.cdf_h = img.height;
      zero_cdfs.assign((size_t)img.height * (img.width + 1), 0.f);
      UP(s->upload(zero_cdfs.data(), (size_t)img.height + 1, &d.bg.row_cdf));
      UP(s->upload(zero_cdfs.data(), zero_cdfs.size(), &d.bg.col_cdfs));
    } else {
      UP(s->upload(hb.row_cdf, (size_t)hb.cdf_height + 1, &d.bg.row_cdf));
      UP(s->upload(hb.col_cdfs, (size_t)hb.cdf_height * (hb.cdf_width + 1), &d.bg.col_cdfs));
    }
  }
  UP(s->commit_uploads());
#undef UP
  // ---- scene preparation on the device (what the reference's constructors do on the host) ----
  for (uint32_t t : build_mips) {
    const DTexture& dt = textures[t];
    launch_mip_levels(const_cast<float*>(d.texels) + 3 * (size_t)dt.level_off[0], dt.width, dt.height, dt.n_levels,
                      dt.wrap_u, dt.wrap_v);
  }
  if (build_cdfs) {
    const DTexture& dt = textures[hb.image];
    const cudaError_t e = launch_envmap_cdf(d.texels + 3 * (size_t)dt.level_off[0], dt.width, dt.height,
                                            const_cast<float*>(d.bg.col_cdfs), const_cast<float*>(d.bg.row_cdf));
    if (e != cudaSuccess) return fail(VIMG_ERR_CUDA, std::string("env-map CDF build: ") + cudaGetErrorString(e));
  }
  if (!build_mips.empty() || build_cdfs) {
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(0));
  }
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


// ---- queue wavefront driver (vimg_queues.cuh) --------------------------------------------
static uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* v = std::getenv(name);
  if (!v || !*v) return dflt;
  return (uint32_t)std::strtoul(v, nullptr, 10);
}

// path-state pool for `n` slots, carved from one allocation (256-byte aligned arrays)
static int qw_reserve(vimg_scene* s, QwPool& pool, uint32_t n) {
  if (!pool.h_cnt) {
    CU_TRY(cudaMallocHost(reinterpret_cast<void**>(&pool.h_cnt), 2 * sizeof(QwCounters)));
    CU_TRY(cudaMallocHost(reinterpret_cast<void**>(&pool.h_next_pixel), 2 * sizeof(uint32_t)));
    CU_TRY(cudaStreamCreateWithFlags(&pool.main, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      CU_TRY(cudaStreamCreateWithFlags(&pool.side[k], cudaStreamNonBlocking));
      CU_TRY(cudaEventCreateWithFlags(&pool.snap[k], cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&pool.join[k], cudaEventDisableTiming));
    }
    CU_TRY(cudaEventCreateWithFlags(&pool.fork, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&pool.done, cudaEventDisableTiming));
  }
  if (!s->ws->d_next_pixel) CU_TRY(cudaMalloc(&s->ws->d_next_pixel, sizeof(uint32_t)));
  pool.st.next_pixel = s->ws->d_next_pixel;
  if (pool.mem && pool.capacity >= n) {
    pool.st.n_slots = n;
    return VIMG_OK;
  }
  if (pool.mem) {
    cudaFree(pool.mem);
    pool.mem = nullptr;
    pool.capacity = 0;
  }
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~(size_t)255;
    return at;
  };
  const size_t N = n;
  const size_t o_slot = carve(N * sizeof(QwSlot));
  size_t o_qc[2], o_qs[2], o_qe[2], o_qsh[WF_NUM_CLASSES];
  for (int k = 0; k < 2; ++k) o_qc[k] = carve(N * 4);
  const size_t o_ql = carve(N * 4);
  for (int k = 0; k < 2; ++k) o_qs[k] = carve(N * 4);
  for (int k = 0; k < 2; ++k) o_qe[k] = carve(N * 4);
  for (uint32_t k = 0; k < WF_NUM_CLASSES; ++k) o_qsh[k] = carve(N * 8);
  const size_t o_cnt = carve(sizeof(QwCounters));
  CU_TRY(cudaMalloc(&pool.mem, off));
  pool.capacity = n;
  char* b = static_cast<char*>(pool.mem);
  QwState& q = pool.st;
  q.n_slots = n;
  q.slot = reinterpret_cast<QwSlot*>(b + o_slot);
  for (int k = 0; k < 2; ++k) q.q_closest[k] = reinterpret_cast<uint32_t*>(b + o_qc[k]);
  q.q_leaf = reinterpret_cast<int32_t*>(b + o_ql);
  for (int k = 0; k < 2; ++k) q.q_shadow[k] = reinterpret_cast<uint32_t*>(b + o_qs[k]);
  for (int k = 0; k < 2; ++k) q.q_end[k] = reinterpret_cast<uint32_t*>(b + o_qe[k]);
  for (uint32_t k = 0; k < WF_NUM_CLASSES; ++k) q.q_shade[k] = reinterpret_cast<uint2*>(b + o_qsh[k]);
  q.cnt = reinterpret_cast<QwCounters*>(b + o_cnt);
  return VIMG_OK;
}

// VIMG_QW_PROFILE=1: one event after every launch, per-kernel totals printed to stderr
struct QwProfile {
  bool on = false, verbose = false;
  cudaStream_t st = nullptr;
  std::vector<cudaEvent_t> ev;
  std::vector<int> kind;
  void mark(int k) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    kind.push_back(k);
  }
  void report(bool print, double* stage_ms, uint32_t* stage_launches) {
    if (!on || ev.empty()) return;
    static const char* names[] = {"start", "trace", "classify", "shade_principled", "shade_lambert",
                                  "shade_dielectric", "end"};
    double tot[7] = {0}, total = 0;
    unsigned cnt[7] = {0};
    cudaEventSynchronize(ev.back());
    for (size_t i = 1; i < ev.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      tot[kind[i]] += ms;
      cnt[kind[i]]++;
      total += ms;
    }
    for (int k = 1; k < 7; ++k) {  // profile kinds 1..6 = vimg_stage 0..5
      stage_ms[k - 1] = tot[k];
      stage_launches[k - 1] = cnt[k];
    }
    if (print) {
      fprintf(stderr, "[qw profile] %.2f ms:", total);
      for (int k = 1; k < 7; ++k)
        if (cnt[k]) fprintf(stderr, " %s %.2f ms (%.0f%%, %u launches, %.1f us avg)", names[k], tot[k], 100 * tot[k] / total, cnt[k], 1e3 * tot[k] / cnt[k]);
      fprintf(stderr, "\n");
    }
    if (print && verbose) {
      // per-launch times of the first iterations (fully loaded) and of a few later ones
      for (size_t i = 1; i < ev.size(); ++i) {
        const size_t iter = i / 5;
        if (!(iter < 6 || (iter % 100) == 0)) continue;
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
        fprintf(stderr, "%s%s %.1f us", kind[i] == 1 ? "\n  it: " : ", ", names[kind[i]], 1e3 * ms);
      }
      fprintf(stderr, "\n");
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
};

// iota kernel: the initial END queue holds every slot (all without a pixel)
__global__ void qw_iota_kernel(uint32_t* q, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) q[i] = i;
}

// spp_streams > 1: thread <-> local pixel; the pixel's range sums are added in range order and
// divided by spp (include/integrators.h:133-137 for one range)
__global__ void qw_combine_kernel(const __grid_constant__ DScene sc, const __grid_constant__ WfArgs args,
                                  float* __restrict__ out) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t px, py;
  if (p >= args.n_local_pixels || !wf_pixel_xy(args, sc.cam, p, px, py)) return;
  const size_t image_index = px + ((size_t)(sc.cam.res[1] - 1 - py) * sc.cam.res[0]);
  const float* part = args.partial + 3 * image_index * args.streams;
  // this rank's ranges (all of them unless the samples are partitioned over ranks), in range order
  uint32_t k = args.s_rank;
  V3 acc = mk3(part[3 * k], part[3 * k + 1], part[3 * k + 2]);
  for (k += args.s_n_ranks; k < args.streams; k += args.s_n_ranks) acc = acc + mk3(part[3 * k], part[3 * k + 1], part[3 * k + 2]);
  acc = acc / (float)args.samples;
  out[3 * image_index] = acc.x;
  out[3 * image_index + 1] = acc.y;
  out[3 * image_index + 2] = acc.z;
}

struct QwLaunchCfg {
  int trace_grid, item_grid;  // full-size grids
  bool concurrent;
  uint32_t class_mask;
  const QwKernelTable* kt;  // bit-exact or tolerance build of the kernels
  int integ;                // 0 material, 1 mis
  bool count;               // traversal counters
  const DWideNode* wide;    // tolerance build, large scene: trace through the 4-wide BVH (else NULL)
};

// the tolerance build of the same kernels (vimg_fast.cu)
extern "C" void vimg_fast_kernel_table(void* out);
extern "C" size_t vimg_fast_flat_scene_bytes(void);
// ... and its streamed shade / END kernels without the code of the features a lean scene does not use (vimg_lean.cu)
extern "C" void vimg_lean_kernel_table(void* out);
// ... and the bit-exact shade / END kernels without that code (vimg_exact_lean.cu)
extern "C" void vimg_exact_lean_kernel_table(void* out);
static const QwKernelTable& kernel_table(bool fast, bool lean = false) {
  static const QwKernelTable exact = qw_kernel_table();
  static const QwKernelTable exact_lean = [] {
    QwKernelTable t = exact;
    vimg_exact_lean_kernel_table(&t);
    return t;
  }();
  static const QwKernelTable fast_t = [] {
    QwKernelTable t{};
    vimg_fast_kernel_table(&t);
    return t;
  }();
  static const QwKernelTable lean_t = [] {
    QwKernelTable t = fast_t;
    vimg_lean_kernel_table(&t);
    return t;
  }();
  return fast ? (lean ? lean_t : fast_t) : (lean ? exact_lean : exact);
}

template <typename... A>
static cudaError_t launch_k(const void* f, int grid, int block, cudaStream_t st, const A&... a) {
  void* args[] = {const_cast<void*>(static_cast<const void*>(&a))...};
  return cudaLaunchKernel(f, dim3((unsigned)grid), dim3((unsigned)block), args, 0, st);
}

// one iteration of one pool: trace(it) classify(it) {shade(it) | end(it)}; iteration 0 is end only
static uint32_t qw_enqueue_iteration(vimg_scene* s, QwPool& pool, const WfArgs& w, uint32_t it, const QwLaunchCfg& cfg,
                                     float* out, QwProfile& prof) {
  const QwState& q = pool.st;
  const uint32_t parity = it & 1u;
  const QwKernelTable& kt = *cfg.kt;
  DCounters* counters = s->d_counters;
  cudaStream_t st = pool.main;
  uint32_t launches = 0;
  if (it > 0) {
    if (cfg.wide)
      launch_k(kt.trace_wide[cfg.count], pool.trace_grid, QW_TRACE_THREADS, st, s->d, q, parity, cfg.wide, counters);
    else
      launch_k(kt.trace[cfg.count], pool.trace_grid, QW_TRACE_THREADS, st, s->d, q, parity, counters);
    prof.mark(1);
    launch_k(kt.classify, pool.item_grid, QW_THREADS, st, s->d, q, parity);
    prof.mark(2);
    // shade launches only feed the next iteration, END only consumes what classify and the
    // previous shade launches queued: fork them onto side streams, join before the next trace
    cudaStream_t sp = st, sl = st;
    if (cfg.concurrent) {
      sp = pool.side[0];
      sl = pool.side[1];
      cudaEventRecord(pool.fork, st);
      cudaStreamWaitEvent(sp, pool.fork, 0);
      cudaStreamWaitEvent(sl, pool.fork, 0);
    }
    if (cfg.class_mask & (1u << WF_CLASS_PRINCIPLED)) {
      launch_k(kt.shade[cfg.integ][WF_CLASS_PRINCIPLED], pool.item_grid, QW_THREADS, sp, s->d, q, w, parity, counters);
      prof.mark(3);
      ++launches;
    }
    if (cfg.class_mask & (1u << WF_CLASS_LAMBERT)) {
      launch_k(kt.shade[cfg.integ][WF_CLASS_LAMBERT], pool.item_grid, QW_THREADS, sl, s->d, q, w, parity, counters);
      prof.mark(4);
      ++launches;
    }
    if (cfg.class_mask & (1u << WF_CLASS_DIELECTRIC)) {
      launch_k(kt.shade[cfg.integ][WF_CLASS_DIELECTRIC], pool.item_grid, QW_THREADS, sl, s->d, q, w, parity, counters);
      prof.mark(5);
      ++launches;
    }
    if (cfg.concurrent) {
      cudaEventRecord(pool.join[0], sp);
      cudaEventRecord(pool.join[1], sl);
    }
    launches += 2;
  }
  launch_k(kt.end[cfg.integ], pool.item_grid, QW_THREADS, st, s->d, q, w, parity, out, counters);
  prof.mark(6);
  ++launches;
  if (cfg.concurrent && it > 0) {
    cudaStreamWaitEvent(st, pool.join[0], 0);
    cudaStreamWaitEvent(st, pool.join[1], 0);
  }
  return launches;
}

// One frame.  Per pool: end(0) | trace(1) classify(1) shade(1) end(1) | ...  until every queue is
// empty.  The host looks at the queue counters once per batch of iterations and keeps one batch
// queued ahead of the check, so the device never waits for the host.
static int qw_render_t(vimg_scene* s, const WfArgs& w_in, int integ, bool count, bool fast, float* out, cudaStream_t caller,
                       bool stage_times) {
  const WfArgs& w = w_in;
  QwProfile prof;
  const bool print_profile = env_u32("VIMG_QW_PROFILE", 0) != 0;
  prof.on = print_profile || stage_times;
  prof.verbose = env_u32("VIMG_QW_PROFILE", 0) > 1;
  const bool timeline = env_u32("VIMG_QW_TIMELINE", 0) != 0;  // ms and live paths per batch, to stderr
  // the per-kernel profile needs serial launches on one stream
  const bool concurrent = !prof.on && env_u32("VIMG_QW_SERIAL", 0) == 0;
  const uint32_t cap = env_u32("VIMG_QW_SLOTS", 1u << 21);
  const uint32_t total_slots = std::max(1u, std::min(w.n_local_units, cap));
  uint32_t n_pools = std::min<uint32_t>(QW_MAX_POOLS, std::max(1u, env_u32("VIMG_QW_POOLS", 2)));
  if (!concurrent || total_slots < 65536u) n_pools = 1;  // small frames: one pool, lowest latency
  const uint32_t per_pool = (total_slots + n_pools - 1) / n_pools;
  if (s->sm_count == 0) cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);

  int rc;
  for (uint32_t p = 0; p < n_pools; ++p)
    if ((rc = qw_reserve(s, s->ws->qw[p], per_pool)) != VIMG_OK) return rc;
  CU_TRY(cudaMemsetAsync(s->ws->d_next_pixel, 0, sizeof(uint32_t), caller));
  QwLaunchCfg cfg{};
  // (the tolerance build's lean kernels are the streamed ones: qv_render)
  cfg.kt = &kernel_table(fast, !fast && s->lean && env_u32("VIMG_LEAN", 1) != 0);
  if (fast && s->wide_ok && env_u32("VIMG_WIDE", 1) != 0) {
    if (!s->d_wide) {
      // device-side re-layout of a large BVH: every binary inner node's grandchildren as one 128-byte record
      CU_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_wide), (size_t)s->n_pairs * sizeof(DWideNode)));
      CU_TRY(launch_k(cfg.kt->wide_build, (int)((s->n_pairs + 255) / 256), 256, caller, s->d.pairs, s->n_pairs, s->d_wide));
    }
    cfg.wide = s->d_wide;
  }
  if (env_u32("VIMG_VERBOSE", 0))
    fprintf(stderr, "[vimg] slot-record pipeline: %s build, %s traversal (%u pair records, wide_ok %d)\n", fast ? "tolerance" : "bit-exact",
            cfg.wide ? "4-wide" : "binary", s->n_pairs, (int)s->wide_ok);
  cudaEvent_t start = s->ws->qw[0].done;  // reused as the fork event of the caller's stream
  const int slot_blocks = (int)((per_pool + QW_THREADS - 1) / QW_THREADS);
  for (uint32_t p = 0; p < n_pools; ++p) {
    QwPool& pool = s->ws->qw[p];
    pool.finished = false;
    pool.trace_grid = 0;  // set below
    if (p == 0) CU_TRY(cudaEventRecord(start, caller));
    CU_TRY(cudaStreamWaitEvent(pool.main, start, 0));
    CU_TRY(cudaMemsetAsync(pool.st.slot, 0, (size_t)per_pool * sizeof(QwSlot), pool.main));  // flags = 0: no pixel yet
    QwCounters init{};
    init.n_end[0] = per_pool;  // every slot starts without a pixel: the first launch (END) hands them out
    *pool.h_cnt = init;
    CU_TRY(cudaMemcpyAsync(pool.st.cnt, pool.h_cnt, sizeof(init), cudaMemcpyHostToDevice, pool.main));
    qw_iota_kernel<<<slot_blocks, QW_THREADS, 0, pool.main>>>(pool.st.q_end[0], per_pool);
  }
  // with several pools in flight the persistent trace grid leaves room for the other pool's
  // shade / end blocks on every SM
  const int trace_per_sm = (int)env_u32("VIMG_QW_TRACE_BLOCKS", QW_TRACE_BLOCKS_PER_SM);
  const int trace_full = s->sm_count * std::max(1, std::min(trace_per_sm, QW_TRACE_BLOCKS_PER_SM));
  const int trace_need = (int)((2ull * per_pool + QW_TRACE_THREADS - 1) / QW_TRACE_THREADS);
  cfg.trace_grid = std::max(1, std::min(trace_full, trace_need));
  cfg.item_grid = std::max(1, std::min(slot_blocks, s->sm_count * (int)env_u32("VIMG_QW_ITEM_BLOCKS", 8) * (256 / QW_THREADS)));
  cfg.concurrent = concurrent;
  cfg.class_mask = s->class_mask;
  cfg.integ = integ;
  cfg.count = count;
  for (uint32_t p = 0; p < n_pools; ++p) {
    s->ws->qw[p].trace_grid = cfg.trace_grid;
    s->ws->qw[p].item_grid = cfg.item_grid;
  }
  const bool shrink = env_u32("VIMG_QW_SHRINK", 1) != 0;
  // VIMG_QW_GRAPH=1: batches after the first are replayed from a CUDA graph per pool (trace /
  // classify / shade on side streams / end, `batch` iterations): one host call per batch
  // instead of ~12 per iteration.  The graph holds the frame's arguments and the pool's launch
  // sizes, so it is captured per frame and again when the launches shrink (sizes are rounded up
  // to powers of two to keep that rare).  Off by default: measured on B200 the replay is 2 %
  // SLOWER than stream launches (disney_spheres 128 spp: 252.6 vs 247.4 ms) -- the host is not
  // the bottleneck, the kernels run back to back either way; it is kept for hosts whose launch
  // thread is contended.
  const bool use_graph = concurrent && env_u32("VIMG_QW_GRAPH", 0) != 0;
  uint32_t batch = std::max(1u, env_u32("VIMG_QW_BATCH", 8));
  if (use_graph) batch = (batch + 1) & ~1u;  // even: every replay sees the same queue parities
  for (uint32_t p = 0; p < n_pools; ++p) s->ws->qw[p].drop_graph();
  auto pow2_at_least = [](uint64_t v) {
    uint64_t r = 1;
    while (r < v) r <<= 1;
    return r;
  };
  uint32_t launches = n_pools, it = 0;
  prof.st = s->ws->qw[0].main;
  prof.mark(0);
  std::vector<cudaEvent_t> tl_ev;
  std::vector<uint32_t> tl_live;
  for (uint32_t b = 0;; ++b) {
    if (!use_graph || b == 0) {
      for (uint32_t k = 0; k < batch; ++k)
        for (uint32_t p = 0; p < n_pools; ++p)
          if (!s->ws->qw[p].finished)
            launches += qw_enqueue_iteration(s, s->ws->qw[p], w, it + k, cfg, out, prof);
    } else {
      for (uint32_t p = 0; p < n_pools; ++p) {
        QwPool& pool = s->ws->qw[p];
        if (pool.finished) continue;
        if (!pool.batch_exec || pool.exec_trace_grid != pool.trace_grid || pool.exec_item_grid != pool.item_grid) {
          pool.drop_graph();
          cudaGraph_t g = nullptr;
          uint32_t n = 0;
          CU_TRY(cudaStreamBeginCapture(pool.main, cudaStreamCaptureModeRelaxed));
          for (uint32_t k = 0; k < batch; ++k)
            n += qw_enqueue_iteration(s, pool, w, batch + k, cfg, out, prof);
          CU_TRY(cudaStreamEndCapture(pool.main, &g));
          cudaError_t e = cudaGraphInstantiate(&pool.batch_exec, g, 0);
          cudaGraphDestroy(g);
          if (e != cudaSuccess) return fail(VIMG_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
          pool.exec_trace_grid = pool.trace_grid;
          pool.exec_item_grid = pool.item_grid;
          pool.graph_launches = n;
        }
        CU_TRY(cudaGraphLaunch(pool.batch_exec, pool.main));
        launches += pool.graph_launches;
      }
    }
    it += batch;
    CU_TRY(cudaGetLastError());
    for (uint32_t p = 0; p < n_pools; ++p) {
      QwPool& pool = s->ws->qw[p];
      if (pool.finished) continue;
      CU_TRY(cudaMemcpyAsync(&pool.h_cnt[b & 1], pool.st.cnt, sizeof(QwCounters), cudaMemcpyDeviceToHost, pool.main));
      CU_TRY(cudaMemcpyAsync(&pool.h_next_pixel[b & 1], s->ws->d_next_pixel, sizeof(uint32_t), cudaMemcpyDeviceToHost, pool.main));
      CU_TRY(cudaEventRecord(pool.snap[b & 1], pool.main));
    }
    if (timeline) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, s->ws->qw[0].main);
      tl_ev.push_back(e);
    }
    if (b > 0) {
      bool all_done = true;
      uint32_t live = 0;
      for (uint32_t p = 0; p < n_pools; ++p) {
        QwPool& pool = s->ws->qw[p];
        if (pool.finished) continue;
        CU_TRY(cudaEventSynchronize(pool.snap[(b - 1) & 1]));
        // queues the iteration after batch b-1 consumes: parity of its first iteration
        const QwCounters& hc = pool.h_cnt[(b - 1) & 1];
        const uint32_t par = (b * batch) & 1u;
        live += hc.n_closest[par];
        if (hc.n_closest[par] == 0 && hc.n_shadow[par] == 0 && hc.n_end[par] == 0)
          pool.finished = true;  // batch b only ran empty iterations for this pool
        else
          all_done = false;
        if (shrink && pool.h_next_pixel[(b - 1) & 1] >= w.n_local_units) {
          // every unit is handed out: the slots alive now bound every later queue of this pool,
          // so the launches shrink with them (kernels stride over their queues, any grid is valid)
          const uint64_t bound = (uint64_t)hc.n_closest[par] + hc.n_end[par];
          uint64_t ig = (bound + QW_THREADS - 1) / QW_THREADS, tg = (2 * bound + QW_TRACE_THREADS - 1) / QW_TRACE_THREADS;
          if (use_graph) {
            ig = pow2_at_least(ig);
            tg = pow2_at_least(tg);
          }
          pool.item_grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(cfg.item_grid, ig));
          pool.trace_grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(cfg.trace_grid, tg));
        }
      }
      if (timeline) tl_live.push_back(live);
      if (all_done) break;
    }
  }
  // the caller's stream continues when every pool is drained
  for (uint32_t p = 0; p < n_pools; ++p) {
    CU_TRY(cudaEventRecord(s->ws->qw[p].done, s->ws->qw[p].main));
    CU_TRY(cudaStreamWaitEvent(caller, s->ws->qw[p].done, 0));
  }
  if (w.streams > 1) {
    // (a rank that renders a single range of a sample-range partition still divides its sum by spp here)
    qw_combine_kernel<<<(w.n_local_pixels + QW_THREADS - 1) / QW_THREADS, QW_THREADS, 0, caller>>>(s->d, w, out);
    ++launches;
  }
  prof.report(print_profile, s->stage_ms, s->stage_launches);
  if (timeline && !tl_ev.empty()) {
    cudaEventSynchronize(tl_ev.back());
    fprintf(stderr, "[qw timeline] %u pool(s), batch of %u iterations: ms | closest rays queued after it\n", n_pools, batch);
    const size_t step = std::max<size_t>(1, tl_ev.size() / 48);
    for (size_t i = step; i < tl_ev.size(); i += step) {
      float ms = 0;
      cudaEventElapsedTime(&ms, tl_ev[i - step], tl_ev[i]);
      fprintf(stderr, "  batches %zu-%zu: %.3f ms/batch, live %u\n", i - step, i, ms / step,
              i - 1 < tl_live.size() ? tl_live[i - 1] : 0u);
    }
    for (cudaEvent_t e : tl_ev) cudaEventDestroy(e);
  }
  s->last_launches = launches;
  return VIMG_OK;
}

// ---- the streamed pipeline (vimg_stream.cuh): tolerance build, flat scenes ---------------------------------
static int qv_reserve(vimg_scene* s, QwPool& streams, QvPool& pool, uint32_t n) {
  if (!streams.main) {  // streams and events of the pool (shared with the slot-record pipeline)
    CU_TRY(cudaMallocHost(reinterpret_cast<void**>(&streams.h_cnt), 2 * sizeof(QwCounters)));
    CU_TRY(cudaMallocHost(reinterpret_cast<void**>(&streams.h_next_pixel), 2 * sizeof(uint32_t)));
    CU_TRY(cudaStreamCreateWithFlags(&streams.main, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      CU_TRY(cudaStreamCreateWithFlags(&streams.side[k], cudaStreamNonBlocking));
      CU_TRY(cudaEventCreateWithFlags(&streams.snap[k], cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&streams.join[k], cudaEventDisableTiming));
    }
    CU_TRY(cudaEventCreateWithFlags(&streams.fork, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&streams.done, cudaEventDisableTiming));
  }
  if (!pool.h_cnt) CU_TRY(cudaMallocHost(reinterpret_cast<void**>(&pool.h_cnt), 2 * sizeof(QvCounters)));
  if (!s->ws->d_next_pixel) CU_TRY(cudaMalloc(&s->ws->d_next_pixel, sizeof(uint32_t)));
  pool.st.next_pixel = s->ws->d_next_pixel;
  if (pool.mem && pool.capacity >= n) {
    pool.st.n_slots = n;
    return VIMG_OK;
  }
  if (pool.mem) {
    cudaFree(pool.mem);
    pool.mem = nullptr;
    pool.capacity = 0;
  }
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~(size_t)255;
    return at;
  };
  const size_t N = n;
  const size_t o_line = carve(N * sizeof(QvLine));
  size_t o_h[2][WF_NUM_CLASSES][2], o_e[2][2];
  for (int k = 0; k < 2; ++k) {
    for (uint32_t cl = 0; cl < WF_NUM_CLASSES; ++cl)
      for (int a = 0; a < 2; ++a) o_h[k][cl][a] = carve(N * 16);
    for (int a = 0; a < 2; ++a) o_e[k][a] = carve(N * 16);
  }
  const size_t o_cnt = carve(sizeof(QvCounters));
  CU_TRY(cudaMalloc(&pool.mem, off));
  pool.capacity = n;
  char* b = static_cast<char*>(pool.mem);
  QvState& q = pool.st;
  q.n_slots = n;
  q.line = reinterpret_cast<QvLine*>(b + o_line);
  for (int k = 0; k < 2; ++k) {
    q.ea[k] = reinterpret_cast<float4*>(b + o_e[k][0]);
    q.eb[k] = reinterpret_cast<float4*>(b + o_e[k][1]);
    for (uint32_t cl = 0; cl < WF_NUM_CLASSES; ++cl) {
      q.ha[k][cl] = reinterpret_cast<float4*>(b + o_h[k][cl][0]);
      q.hb[k][cl] = reinterpret_cast<float4*>(b + o_h[k][cl][1]);
    }
  }
  q.cnt = reinterpret_cast<QvCounters*>(b + o_cnt);
  return VIMG_OK;
}

// One frame: per pool  end(0) | {shade(1) per class, end(1)} | {shade(2), end(2)} ...  until the hit and END queues
// are empty -- the launches of an iteration read the queues of its parity and fill those of the other one (each
// traces the rays it makes itself), so they run side by side; the host clears the drained counters between two
// iterations and looks at them once per batch of iterations, one batch behind what it has queued.
static int qv_render(vimg_scene* s, const WfArgs& w, int integ, bool count, float* out, cudaStream_t caller, bool stage_times) {
  QwProfile prof;
  const bool print_profile = env_u32("VIMG_QW_PROFILE", 0) != 0;
  prof.on = print_profile || stage_times;
  const bool concurrent = !prof.on && env_u32("VIMG_QW_SERIAL", 0) == 0;
  const uint32_t cap = env_u32("VIMG_QW_SLOTS", 1u << 21);
  const uint32_t total_slots = std::max(1u, std::min(w.n_local_units, cap));
  uint32_t n_pools = std::min<uint32_t>(QW_MAX_POOLS, std::max(1u, env_u32("VIMG_QW_POOLS", 2)));
  if (!concurrent || total_slots < 65536u) n_pools = 1;
  const uint32_t per_pool = (total_slots + n_pools - 1) / n_pools;
  if (s->sm_count == 0) cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
  const QwKernelTable& kt = kernel_table(true, s->lean && env_u32("VIMG_LEAN", 1) != 0);
  int rc;
  for (uint32_t p = 0; p < n_pools; ++p)
    if ((rc = qv_reserve(s, s->ws->qw[p], s->ws->qv[p], per_pool)) != VIMG_OK) return rc;
  CU_TRY(cudaMemsetAsync(s->ws->d_next_pixel, 0, sizeof(uint32_t), caller));
  if (!s->d_flat) {
    CU_TRY(cudaMalloc(&s->d_flat, vimg_fast_flat_scene_bytes()));
    CU_TRY(launch_k(kt.build_flat, 1, (int)QF_FLAT_MAX_HOST, caller, s->d, s->d_flat));
  }
  if (env_u32("VIMG_VERBOSE", 0))
    fprintf(stderr, "[vimg] streamed pipeline: %s shade / END kernels, %u pool(s) of %u slots\n",
            &kt == &kernel_table(true, true) ? "lean" : "full", n_pools, per_pool);
  const void* d_flat = s->d_flat;
  DCounters* counters = s->d_counters;
  cudaEvent_t start = s->ws->qw[0].done;
  const int slot_blocks = (int)((per_pool + QW_THREADS - 1) / QW_THREADS);
  // 4 blocks per SM: every thread works through several items, so the next item's line is on its way while this one is shaded
  const int full_grid = std::max(1, std::min(slot_blocks, s->sm_count * (int)env_u32("VIMG_QW_ITEM_BLOCKS", 4)));
  std::vector<int> grid(n_pools, full_grid);
  std::vector<bool> finished(n_pools, false);
  for (uint32_t p = 0; p < n_pools; ++p) {
    QwPool& st = s->ws->qw[p];
    QvPool& pool = s->ws->qv[p];
    if (p == 0) CU_TRY(cudaEventRecord(start, caller));
    CU_TRY(cudaStreamWaitEvent(st.main, start, 0));
    CU_TRY(cudaMemsetAsync(pool.st.line, 0, (size_t)per_pool * sizeof(QvLine), st.main));  // flags = 0: no pixel yet
    QvCounters init{};
    init.q[0].n_end = per_pool;  // every slot starts in the END queue, without a pixel
    pool.h_cnt[0] = init;
    CU_TRY(cudaMemcpyAsync(pool.st.cnt, &pool.h_cnt[0], sizeof(init), cudaMemcpyHostToDevice, st.main));
    CU_TRY(launch_k(kt.sv_init, slot_blocks, QW_THREADS, st.main, pool.st.ea[0], pool.st.eb[0], per_pool));
  }
  auto item_grid = [&](uint32_t p) { return grid[p] * (256 / QV_THREADS); };  // shade / END blocks of QV_THREADS threads
  const uint32_t count_flag = count ? 1u : 0u;
  auto enqueue = [&](uint32_t p, uint32_t it) -> uint32_t {
    QwPool& st = s->ws->qw[p];
    const QvState& q = s->ws->qv[p].st;
    const uint32_t parity = it & 1u;
    uint32_t launches = 0;
    cudaStream_t m = st.main;
    if (it > 0) {
      // the queues this iteration fills were drained by the previous one
      cudaMemsetAsync(&q.cnt->q[parity ^ 1u], 0, sizeof(QvQueueCounters), m);
      cudaStream_t sp = m, sl = m;
      if (concurrent) {
        sp = st.side[0];
        sl = st.side[1];
        cudaEventRecord(st.fork, m);
        cudaStreamWaitEvent(sp, st.fork, 0);
        cudaStreamWaitEvent(sl, st.fork, 0);
      }
      if (s->class_mask & (1u << WF_CLASS_PRINCIPLED)) {
        launch_k(kt.sv_shade[integ][WF_CLASS_PRINCIPLED], item_grid(p), QV_THREADS, sp, s->d, q, w, parity, d_flat, count_flag, counters);
        prof.mark(3);
        ++launches;
      }
      if (s->class_mask & (1u << WF_CLASS_LAMBERT)) {
        launch_k(kt.sv_shade[integ][WF_CLASS_LAMBERT], item_grid(p), QV_THREADS, sl, s->d, q, w, parity, d_flat, count_flag, counters);
        prof.mark(4);
        ++launches;
      }
      if (s->class_mask & (1u << WF_CLASS_DIELECTRIC)) {
        launch_k(kt.sv_shade[integ][WF_CLASS_DIELECTRIC], item_grid(p), QV_THREADS, sl, s->d, q, w, parity, d_flat, count_flag, counters);
        prof.mark(5);
        ++launches;
      }
      if (concurrent) {
        cudaEventRecord(st.join[0], sp);
        cudaEventRecord(st.join[1], sl);
      }
    }
    launch_k(kt.sv_end[integ], item_grid(p), QV_THREADS, m, s->d, q, w, parity, out, d_flat, count_flag, counters);
    prof.mark(6);
    ++launches;
    if (concurrent && it > 0) {
      cudaStreamWaitEvent(m, st.join[0], 0);
      cudaStreamWaitEvent(m, st.join[1], 0);
    }
    return launches;
  };
  const uint32_t batch = std::max(1u, env_u32("VIMG_QW_BATCH", 8));
  const bool shrink = env_u32("VIMG_QW_SHRINK", 1) != 0;
  uint32_t launches = n_pools, it = 0;
  prof.st = s->ws->qw[0].main;
  prof.mark(0);
  for (uint32_t b = 0;; ++b) {
    for (uint32_t k = 0; k < batch; ++k)
      for (uint32_t p = 0; p < n_pools; ++p)
        if (!finished[p]) launches += enqueue(p, it + k);
    it += batch;
    CU_TRY(cudaGetLastError());
    for (uint32_t p = 0; p < n_pools; ++p) {
      if (finished[p]) continue;
      QwPool& st = s->ws->qw[p];
      QvPool& pool = s->ws->qv[p];
      CU_TRY(cudaMemcpyAsync(&pool.h_cnt[b & 1], pool.st.cnt, sizeof(QvCounters), cudaMemcpyDeviceToHost, st.main));
      CU_TRY(cudaMemcpyAsync(&st.h_next_pixel[b & 1], s->ws->d_next_pixel, sizeof(uint32_t), cudaMemcpyDeviceToHost, st.main));
      CU_TRY(cudaEventRecord(st.snap[b & 1], st.main));
    }
    if (b > 0) {
      bool all_done = true;
      for (uint32_t p = 0; p < n_pools; ++p) {
        if (finished[p]) continue;
        QwPool& st = s->ws->qw[p];
        CU_TRY(cudaEventSynchronize(st.snap[(b - 1) & 1]));
        const QvCounters& hc = s->ws->qv[p].h_cnt[(b - 1) & 1];
        const uint32_t par = (b * batch) & 1u;  // parity of the first iteration after batch b - 1
        uint64_t queued = hc.q[par].n_end;
        for (uint32_t cl = 0; cl < WF_NUM_CLASSES; ++cl) queued += hc.q[par].n_shade[cl];
        if (queued == 0)
          finished[p] = true;  // batch b only ran empty iterations for this pool
        else
          all_done = false;
        if (shrink && st.h_next_pixel[(b - 1) & 1] >= w.n_local_units) {
          grid[p] = (int)std::max<uint64_t>(1, std::min<uint64_t>(full_grid, (queued + QW_THREADS - 1) / QW_THREADS));
        }
      }
      if (all_done) break;
    }
  }
  for (uint32_t p = 0; p < n_pools; ++p) {
    CU_TRY(cudaEventRecord(s->ws->qw[p].done, s->ws->qw[p].main));
    CU_TRY(cudaStreamWaitEvent(caller, s->ws->qw[p].done, 0));
  }
  if (w.streams > 1) {
    qw_combine_kernel<<<(w.n_local_pixels + QW_THREADS - 1) / QW_THREADS, QW_THREADS, 0, caller>>>(s->d, w, out);
    ++launches;
  }
  prof.report(print_profile, s->stage_ms, s->stage_launches);
  s->last_launches = launches;
  return VIMG_OK;
}

template <bool COUNT>
static int qw_render(vimg_scene* s, const WfArgs& w, uint32_t integrator, bool fast, float* out, cudaStream_t st,
                     bool stage_times) {
  // tolerance build, a scene of a few dozen primitives: the streamed pipeline over the flat lists (vimg_stream.cuh)
  if (fast && s->d.n_leaf_recs > 0 && s->d.n_leaf_recs <= std::min<uint32_t>(QF_FLAT_MAX_HOST, env_u32("VIMG_FLAT_MAX", 32))
      && env_u32("VIMG_FLAT", 1) != 0)
    return qv_render(s, w, integrator == VIMG_INTEGRATOR_MIS ? 1 : 0, COUNT, out, st, stage_times);
  return qw_render_t(s, w, integrator == VIMG_INTEGRATOR_MIS ? 1 : 0, COUNT, fast, out, st, stage_times);
}

template <bool COUNT>
static int launch_render(vimg_scene* s, const RenderArgs& a, uint32_t integrator, float* out, cudaStream_t st,
                         uint32_t flags = 0, uint32_t spp_streams = 1, uint32_t s_rank = 0, uint32_t s_n_ranks = 1) {
  s->last_launches = 1;
  std::memset(s->stage_ms, 0, sizeof(s->stage_ms));
  std::memset(s->stage_launches, 0, sizeof(s->stage_launches));
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
    if (w.n_local_pixels == 0) return VIMG_OK;
    if (integrator == VIMG_INTEGRATOR_MATERIAL && a.depth == 0) {
      // material_integrator with depth 0 never traces: every sample is vec3(0) (mat_integrator.cpp:17-21).
      // Nothing for the pipelines to do -- a slot of the queue wavefront that never gets a live path would
      // serve one pixel only -- so the frame is written here: zeros for this rank's pixels (the other
      // ranks' pixels are zero by contract), samples counted.
      unsigned long long n_samples = a.samples;
      if (a.single_x < 0) {
        CU_TRY(cudaMemsetAsync(out, 0, (size_t)s->W * s->H * 3 * sizeof(float), st));
        unsigned long long pixels = 0;
        for (uint32_t lt = 0; lt < a.n_local_tiles; ++lt) {
          const uint32_t tile = a.rank + lt * a.n_ranks, tx = tile / a.tiles_y, ty = tile % a.tiles_y;
          const uint32_t wx = std::min<uint32_t>(8, (uint32_t)s->W - tx * 8), wy = std::min<uint32_t>(8, (uint32_t)s->H - ty * 8);
          pixels += (unsigned long long)wx * wy;
        }
        n_samples *= pixels;
      } else {
        CU_TRY(cudaMemsetAsync(out, 0, 3 * sizeof(float), st));
      }
      CU_TRY(cudaMemcpyAsync(&s->d_counters->samples, &n_samples, sizeof(n_samples), cudaMemcpyHostToDevice, st));
      CU_TRY(cudaStreamSynchronize(st));  // n_samples is a local
      return VIMG_OK;
    }
    // sample ranges: at least one sample each; trace_pixel and the no-trace corner (material,
    // depth 0) keep the single stream
    w.streams = std::max(1u, std::min(spp_streams, a.samples));
    if (a.single_x >= 0 || (integrator == VIMG_INTEGRATOR_MATERIAL && a.depth == 0)) w.streams = 1;
    if ((uint64_t)w.n_local_pixels * w.streams > 0xffffffffull || w.streams > 0xffffu)
      return fail(VIMG_ERR_INVALID_ARGUMENT, "spp_streams too large for this frame");
    // sample-range partition over ranks: this rank's share of every pixel's ranges
    w.s_rank = s_rank;
    w.s_n_ranks = std::max(1u, s_n_ranks);
    if (w.s_n_ranks > 1 && w.streams < w.s_n_ranks)
      return fail(VIMG_ERR_INVALID_ARGUMENT, "VIMG_PARTITION_SPP needs spp_streams (and samples) >= n_ranks");
    w.my_streams = (w.streams - w.s_rank + w.s_n_ranks - 1) / w.s_n_ranks;
    w.n_local_units = w.n_local_pixels * w.my_streams;
    w.div_tiles_y = make_fast_div(std::max(1u, w.tiles_y));
    w.div_my_streams = make_fast_div(std::max(1u, w.my_streams));
    w.div_streams = make_fast_div(w.streams);
    w.samples_small = w.samples < 65536u ? 1u : 0u;
    w.partial = nullptr;
    const bool fast = (flags & VIMG_FLAG_FAST) != 0;
    w.fresh_stack = ((flags & VIMG_FLAG_FRESH_STACK) || w.streams > 1 || fast) ? 1u : 0u;
    // pipeline: the queue wavefront unless the caller (or VIMG_PIPELINE=block) asks for the
    // block-local kernel
    bool use_block = (flags & VIMG_FLAG_PIPELINE_BLOCK) != 0;
    if (!(flags & (VIMG_FLAG_PIPELINE_BLOCK | VIMG_FLAG_PIPELINE_QUEUE))) {
      const char* e = std::getenv("VIMG_PIPELINE");
      use_block = e && std::strcmp(e, "block") == 0;
    }
    if (use_block && (w.streams > 1 || fast))
      return fail(VIMG_ERR_UNSUPPORTED, "spp_streams > 1 and VIMG_FLAG_FAST need the queue wavefront pipeline");
    if (!use_block && !s->ws) s->ws = ws_acquire(s->device);
    if (w.streams > 1) {
      const size_t need = (size_t)s->W * s->H * w.streams * 3;
      if (s->ws->partial_floats < need) {
        if (s->ws->d_partial) cudaFree(s->ws->d_partial);
        s->ws->d_partial = nullptr;
        s->ws->partial_floats = 0;
        if (cudaMalloc(&s->ws->d_partial, need * sizeof(float)) != cudaSuccess) {
          cudaGetLastError();
          return fail(VIMG_ERR_OUT_OF_MEMORY, "partial-sum buffer of the sample ranges");
        }
        s->ws->partial_floats = need;
      }
      w.partial = s->ws->d_partial;
    }
    if (!use_block) return qw_render<COUNT>(s, w, integrator, fast, out, st, (flags & VIMG_FLAG_STAGE_TIMES) != 0);
    cudaMemsetAsync(s->d_next_pixel, 0, sizeof(unsigned int), st);
    if (integrator == VIMG_INTEGRATOR_MIS) {
      const int g = wavefront_grid<VIMG_INTEGRATOR_MIS, COUNT>(s, w.n_local_pixels);
      wavefront_kernel<VIMG_INTEGRATOR_MIS, COUNT><<<g, WF_THREADS, sizeof(WfShared), st>>>(s->d, w, out, s->d_counters, s->d_next_pixel);
    } else {
      const int g = wavefront_grid<VIMG_INTEGRATOR_MATERIAL, COUNT>(s, w.n_local_pixels);
      wavefront_kernel<VIMG_INTEGRATOR_MATERIAL, COUNT><<<g, WF_THREADS, sizeof(WfShared), st>>>(s->d, w, out, s->d_counters, s->d_next_pixel);
    }
    return VIMG_OK;
  }
  const uint32_t blocks = a.single_x >= 0 ? 1 : (a.n_local_tiles + 1) / 2;
  if (blocks == 0) return VIMG_OK;
  if (integrator == VIMG_INTEGRATOR_S_NORMAL)
    render_kernel<VIMG_INTEGRATOR_S_NORMAL, COUNT><<<blocks, 128, 0, st>>>(s->d, a, out, s->d_counters);
  else
    render_kernel<VIMG_INTEGRATOR_G_NORMAL, COUNT><<<blocks, 128, 0, st>>>(s->d, a, out, s->d_counters);
  return VIMG_OK;
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

void vimg_release_workspaces(void) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : g_arena_idle) {
    if (!kv.second.ptr) continue;
    cudaSetDevice(kv.first);
    cudaFree(kv.second.ptr);
    kv.second = IdleArena{};
  }
  for (auto& kv : g_ws_idle) {
    if (!kv.second) continue;
    cudaSetDevice(kv.first);
    kv.second->destroy();
    delete kv.second;
    kv.second = nullptr;
  }
  cudaSetDevice(cur);
}

void vimg_scene_destroy(vimg_scene* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  // the arena and the workspace go back to per-device caches that hand them to the next scene without
  // the device-wide synchronisation cudaFree would have implied: kernels of a failed render, or a
  // consumer of vimg_render_device's frame on the caller's stream, may still be in flight
  cudaDeviceSynchronize();
  for (void* p : s->allocs) arena_release(s->device, p, s->arena_bytes);  // scene arrays, frame, counters
  ws_release(s->device, s->ws);
  s->ws = nullptr;
  if (s->d_rgb8) cudaFree(s->d_rgb8);
  if (s->d_post_max) cudaFree(s->d_post_max);
  if (s->d_flat) cudaFree(s->d_flat);
  if (s->d_wide) cudaFree(s->d_wide);
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
  {
    // experiment knob: L2 fetch granularity hint (bytes) for the scattered 32-byte state records
    const uint32_t g = env_u32("VIMG_L2_FETCH", 0);
    if (g) {
      cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
      size_t got = 0;
      cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
      fprintf(stderr, "[vimg] cudaLimitMaxL2FetchGranularity requested %u, now %zu\n", g, got);
      cudaGetLastError();
    }
  }
  vimg_scene* s = new vimg_scene();
  s->device = device_id;
  int rc = build_scene(desc, s);
  if (rc == VIMG_OK) {
    cudaError_t e = cudaEventCreate(&s->ev0);
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

}  // extern "C"

// Render this rank's share of the frame into `frame` (W*H*3 floats in memory the device can write: the scene's own
// frame, or -- vimg_render_multi -- the frame of device 0 mapped through peer access, so that the kernels'
// pixel stores assemble the image over NVLink).  Returns after the stream is synchronised.
static int render_into(vimg_scene* s, const vimg_render_params* p, float* frame, bool zero_fill, cudaStream_t st,
                       vimg_render_stats* stats) {
  int rc = check_params(s, p);
  if (rc != VIMG_OK) return rc;
  CU_TRY(cudaSetDevice(s->device));
  const bool by_spp = p->partition == VIMG_PARTITION_SPP && p->n_ranks > 1;
  if (by_spp && p->integrator != VIMG_INTEGRATOR_MATERIAL && p->integrator != VIMG_INTEGRATOR_MIS)
    return fail(VIMG_ERR_INVALID_ARGUMENT, "VIMG_PARTITION_SPP is for the material / mis integrators");
  // sample-range partition: every rank renders all tiles, a share of every pixel's sample ranges
  const RenderArgs a = make_args(s, p->samples, p->depth, (p->n_ranks > 1 && !by_spp) ? p->rank : 0, by_spp ? 1 : p->n_ranks);
  const uint32_t s_rank = by_spp ? p->rank : 0, s_n = by_spp ? p->n_ranks : 1;
  CU_TRY(cudaMemsetAsync(s->d_counters, 0, sizeof(DCounters), st));
  if (a.n_ranks > 1 && zero_fill) CU_TRY(cudaMemsetAsync(frame, 0, (size_t)s->W * s->H * 3 * sizeof(float), st));
  CU_TRY(cudaEventRecord(s->ev0, st));
  if (p->flags & VIMG_FLAG_COUNTERS)
    rc = launch_render<true>(s, a, p->integrator, frame, st, p->flags, p->spp_streams, s_rank, s_n);
  else
    rc = launch_render<false>(s, a, p->integrator, frame, st, p->flags, p->spp_streams, s_rank, s_n);
  if (rc != VIMG_OK) return rc;
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaEventRecord(s->ev1, st));
  CU_TRY(cudaStreamSynchronize(st));
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
    stats->kernel_launches = s->last_launches;
    std::memcpy(stats->stage_ms, s->stage_ms, sizeof(s->stage_ms));
    std::memcpy(stats->stage_launches, s->stage_launches, sizeof(s->stage_launches));
    for (int k = 0; k < VIMG_STAGE_COUNT; ++k) stats->stage_items[k] = c.stage_items[k];
#ifdef WF_STAGE_CLOCKS
    fprintf(stderr, "[wf stage cycles, summed over warps] shadow %llu regen %llu closest %llu resolve+bucket %llu "
            "barrier1 %llu shade %llu barrier2 %llu\n", c.stage_clk[0], c.stage_clk[1], c.stage_clk[2],
            c.stage_clk[3], c.stage_clk[4], c.stage_clk[5], c.stage_clk[6]);
#endif
  }
  return VIMG_OK;
}

extern "C" {

int vimg_render_device(vimg_scene* s, const vimg_render_params* p, float** out_dev_rgb, void* stream,
                       vimg_render_stats* stats) {
  if (!s || !out_dev_rgb) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument");
  const int rc = render_into(s, p, s->d_frame, true, static_cast<cudaStream_t>(stream), stats);
  if (rc == VIMG_OK) *out_dev_rgb = s->d_frame;
  return rc;
}

}  // extern "C"

// dst += src, element by element (device 0 pulling a peer's frame over NVLink)
__global__ void multi_add_frame_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<float4*>(dst)[i];
  const float4 b = reinterpret_cast<const float4*>(src)[i];
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
  a.w += b.w;
  reinterpret_cast<float4*>(dst)[i] = a;
}

extern "C" {

int vimg_render_multi(const vimg_scene_desc* desc, const int* device_ids, uint32_t n_devices,
                      const vimg_render_params* params, uint32_t partition, float* out_rgb, vimg_render_stats* stats) {
  if (!desc || !device_ids || !params || !out_rgb || n_devices == 0) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument");
  if (partition > VIMG_PARTITION_SPP) return fail(VIMG_ERR_INVALID_ARGUMENT, "unknown partition");
  const int n_visible = vimg_cuda_device_count();
  if (n_visible == 0) return fail(VIMG_ERR_NO_DEVICE, "no CUDA device visible: the cuda render device has no CPU fallback");
  for (uint32_t i = 0; i < n_devices; ++i) {
    if (device_ids[i] < 0 || device_ids[i] >= n_visible) return fail(VIMG_ERR_INVALID_ARGUMENT, "device id out of range");
    for (uint32_t j = 0; j < i; ++j)
      if (device_ids[j] == device_ids[i]) return fail(VIMG_ERR_INVALID_ARGUMENT, "device listed twice");
  }
  const uint32_t n = n_devices;
  const int dev0 = device_ids[0];
  std::vector<vimg_scene*> scenes(n, nullptr);
  std::vector<int> rcs(n, VIMG_OK);
  std::vector<std::string> errs(n);
  // every worker keeps its thread-local error text for the caller
  auto run_all = [&](auto&& fn) {
    std::vector<std::thread> th;
    for (uint32_t i = 1; i < n; ++i)
      th.emplace_back([&, i] {
        rcs[i] = fn(i);
        if (rcs[i] != VIMG_OK) errs[i] = g_last_error;
      });
    rcs[0] = fn(0);
    if (rcs[0] != VIMG_OK) errs[0] = g_last_error;
    for (auto& t : th) t.join();
    for (uint32_t i = 0; i < n; ++i)
      if (rcs[i] != VIMG_OK) return fail(rcs[i], "device " + std::to_string(device_ids[i]) + ": " + errs[i]);
    return (int)VIMG_OK;
  };
  auto destroy_all = [&] {
    run_all([&](uint32_t i) {
      vimg_scene_destroy(scenes[i]);
      scenes[i] = nullptr;
      return (int)VIMG_OK;
    });
  };
  // ---- the scene on every device (one host thread each: flattening to the GPU layout is host work)
  int rc = run_all([&](uint32_t i) { return vimg_scene_create(desc, device_ids[i], &scenes[i]); });
  if (rc != VIMG_OK) {
    const std::string keep = g_last_error;
    destroy_all();
    return fail(rc, keep);
  }
  const size_t n_floats = (size_t)scenes[0]->W * scenes[0]->H * 3;
  // ---- peer access: TILE: device i stores into device 0's frame; SPP: device 0 loads device i's frame
  bool peer = n > 1;
  for (uint32_t i = 1; i < n && peer; ++i) {
    int can = 0;
    const int from = partition == VIMG_PARTITION_TILE ? device_ids[i] : dev0;
    const int to = partition == VIMG_PARTITION_TILE ? dev0 : device_ids[i];
    if (cudaDeviceCanAccessPeer(&can, from, to) != cudaSuccess || !can) {
      cudaGetLastError();
      peer = false;
      break;
    }
    cudaSetDevice(from);
    const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer = false;
    cudaGetLastError();
  }
  if (const char* e = std::getenv("VIMG_MULTI_NO_PEER")) peer = peer && std::atoi(e) == 0;

  std::vector<vimg_render_stats> st(n);
  std::vector<std::vector<float>> staged(peer ? 0 : n);
  const auto t0 = std::chrono::steady_clock::now();
  rc = run_all([&](uint32_t i) -> int {
    vimg_render_params p = *params;
    p.rank = i;
    p.n_ranks = n;
    p.partition = partition;
    const bool to_dev0 = peer && partition == VIMG_PARTITION_TILE;
    float* frame = to_dev0 ? scenes[0]->d_frame : scenes[i]->d_frame;
    // TILE through peer memory: every pixel has exactly one writer, nothing to zero
    int r = render_into(scenes[i], &p, frame, !to_dev0, nullptr, &st[i]);
    if (r != VIMG_OK) return r;
    if (!peer && n > 1) {
      staged[i].resize(n_floats);
      cudaError_t e = cudaMemcpy(staged[i].data(), scenes[i]->d_frame, n_floats * sizeof(float), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) return fail(VIMG_ERR_CUDA, std::string("partial frame D2H: ") + cudaGetErrorString(e));
    }
    return VIMG_OK;
  });
  if (rc == VIMG_OK) {
    cudaSetDevice(dev0);
    if (n == 1 || peer) {
      if (n > 1 && partition == VIMG_PARTITION_SPP) {
        const size_t n4 = n_floats / 4;  // W * H * 3 floats: a multiple of 4 whenever W * H is; the tail below
        for (uint32_t i = 1; i < n; ++i) {
          if (n4) multi_add_frame_kernel<<<(unsigned)((n4 + 255) / 256), 256>>>(scenes[0]->d_frame, scenes[i]->d_frame, n4);
          for (size_t k = 4 * n4; k < n_floats; ++k) {  // at most 3 floats
            float a = 0, b = 0;
            cudaMemcpy(&a, scenes[0]->d_frame + k, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(&b, scenes[i]->d_frame + k, 4, cudaMemcpyDeviceToHost);
            a += b;
            cudaMemcpy(scenes[0]->d_frame + k, &a, 4, cudaMemcpyHostToDevice);
          }
        }
      }
      const cudaError_t e = cudaMemcpy(out_rgb, scenes[0]->d_frame, n_floats * sizeof(float), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) rc = fail(VIMG_ERR_CUDA, std::string("frame D2H: ") + cudaGetErrorString(e));
    } else {
      // no peer access: the partial frames meet in host memory, added in device order
      std::memcpy(out_rgb, staged[0].data(), n_floats * sizeof(float));
      for (uint32_t i = 1; i < n; ++i)
        for (size_t k = 0; k < n_floats; ++k) out_rgb[k] += staged[i][k];
    }
  }
  const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (rc == VIMG_OK && stats) {
    std::memset(stats, 0, sizeof(*stats));
    for (uint32_t i = 0; i < n; ++i) {
      stats->rays_closest += st[i].rays_closest;
      stats->rays_shadow += st[i].rays_shadow;
      stats->inner_nodes += st[i].inner_nodes;
      stats->leaf_nodes += st[i].leaf_nodes;
      stats->tri_tests += st[i].tri_tests;
      stats->sphere_tests += st[i].sphere_tests;
      stats->nan_samples += st[i].nan_samples;
      stats->samples += st[i].samples;
      stats->kernel_launches += st[i].kernel_launches;
      for (int k = 0; k < VIMG_STAGE_COUNT; ++k) stats->stage_items[k] += st[i].stage_items[k];
      stats->kernel_ms = std::max(stats->kernel_ms, st[i].kernel_ms);
    }
    stats->total_ms = wall_ms;
  }
  const std::string keep = g_last_error;
  destroy_all();
  if (rc != VIMG_OK) return fail(rc, keep);
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

uint32_t vimg_mip_chain_levels(uint32_t width, uint32_t height, uint64_t* out_floats) {
  if (out_floats) *out_floats = 0;
  if (width == 0 || height == 0) return 0;
  const uint32_t size = std::min(width, height);
  // image_texture.cpp:73-78: the loop runs for l = 1 .. num_levels - 1; level 0 always exists
  const int num_levels = std::min((int)std::ceil(std::log2(float(size))), 15);
  const uint32_t n = (uint32_t)std::max(num_levels, 1);
  uint64_t floats = 0;
  uint32_t w = width, h = height;
  for (uint32_t l = 1; l < n; ++l) {
    w = std::max(w / 2u, 1u);
    h = std::max(h / 2u, 1u);
    floats += (uint64_t)w * h * 3;
  }
  if (out_floats) *out_floats = floats;
  return n;
}

// device selection + argument checks shared by the scene-preparation entry points
static int prep_device(int device_id, int* sms) {
  const int n = vimg_cuda_device_count();
  if (n == 0) return fail(VIMG_ERR_NO_DEVICE, "no CUDA device visible: the cuda render device has no CPU fallback");
  if (device_id < 0 || device_id >= n) return fail(VIMG_ERR_INVALID_ARGUMENT, "device_id out of range");
  CU_TRY(cudaSetDevice(device_id));
  CU_TRY(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, device_id));
  return VIMG_OK;
}

// levels 1 .. n-1 behind level 0 in d_all, one launch per level (image_texture.cpp:80-125)
static void launch_mip_levels(float* d_all, uint32_t width, uint32_t height, uint32_t n_levels, uint32_t wrap_u,
                              uint32_t wrap_v) {
  const float* prev = d_all;
  float* next = d_all + (size_t)width * height * 3;
  uint32_t pw = width, ph = height;
  for (uint32_t l = 1; l < n_levels; ++l) {
    const uint32_t nw = std::max(pw / 2u, 1u), nh = std::max(ph / 2u, 1u);
    prep_mip_level_kernel<<<dim3((nw + 255) / 256, nh), 256>>>(prev, pw, ph, next, nw, nh, wrap_u, wrap_v);
    prev = next;
    next += (size_t)nw * nh * 3;
    pw = nw;
    ph = nh;
  }
}

// blocks of 256 threads for `items` grid-stride work items: a multiple of the SM count, at most 8 per SM
static unsigned prep_grid(uint64_t items, int sms) {
  const uint64_t want = (items + 255) / 256;
  const uint64_t cap = (uint64_t)sms * 8;
  if (want >= cap) return (unsigned)cap;
  return (unsigned)std::max<uint64_t>(1, want);
}

// shared body of the byte conversions and vimg_build_image_texture: bytes in, floats out
//   mode 0: sRGB -> linear; mode 1: RGB -> normal (scale); mode 2: sRGB -> linear + mip chain
static int convert_bytes(const char* what, int mode, const uint8_t* rgb8, uint32_t width, uint64_t n_pixels,
                         uint32_t height, float scale, uint32_t wrap_u, uint32_t wrap_v, int device_id,
                         float* out_rgb, float* out_levels, double* kernel_ms) {
  if (kernel_ms) *kernel_ms = 0.0;
  if (n_pixels == 0) return VIMG_OK;
  if (!rgb8 || !out_rgb) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument");
  int sms = 0, rc = prep_device(device_id, &sms);
  if (rc != VIMG_OK) return rc;
  uint64_t floats = 0;
  uint32_t n_levels = 1;
  if (mode == 2) {
    n_levels = vimg_mip_chain_levels(width, height, &floats);
    if (n_levels > 1 && !out_levels) return fail(VIMG_ERR_INVALID_ARGUMENT, "null output");
  }
  const size_t l0 = (size_t)n_pixels * 3;
  uint8_t* d_in = nullptr;
  float* d_all = nullptr;  // the converted image (level 0) followed by levels 1 .. n-1
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMalloc(&d_in, l0);
  if (e == cudaSuccess) e = cudaMalloc(&d_all, (l0 + floats) * sizeof(float));
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) e = cudaMemcpy(d_in, rgb8, l0, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    cudaEventRecord(e0, 0);
    if (mode == 1)
      prep_rgb8_to_normal_kernel<<<prep_grid((n_pixels + 3) / 4, sms), 256>>>(d_in, n_pixels, scale, d_all);
    else
      prep_srgb8_to_linear_kernel<<<prep_grid((l0 + 15) / 16, sms), 256>>>(d_in, l0, d_all);
    if (n_levels > 1) launch_mip_levels(d_all, width, height, n_levels, wrap_u, wrap_v);
    cudaEventRecord(e1, 0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out_rgb, d_all, l0 * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && floats) e = cudaMemcpy(out_levels, d_all + l0, floats * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && kernel_ms) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *kernel_ms = ms;
  }
  if (d_in) cudaFree(d_in);
  if (d_all) cudaFree(d_all);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (e != cudaSuccess) return fail(VIMG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return VIMG_OK;
}

int vimg_convert_srgb8_to_linear(const uint8_t* rgb8, uint64_t n_pixels, int device_id, float* out_rgb,
                                 double* kernel_ms) {
  return convert_bytes("vimg_convert_srgb8_to_linear", 0, rgb8, 0, n_pixels, 0, 0.f, 0, 0, device_id, out_rgb, nullptr,
                       kernel_ms);
}

int vimg_convert_rgb8_to_normal(const uint8_t* rgb8, uint64_t n_pixels, float scale, int device_id, float* out_rgb,
                                double* kernel_ms) {
  return convert_bytes("vimg_convert_rgb8_to_normal", 1, rgb8, 0, n_pixels, 0, scale, 0, 0, device_id, out_rgb, nullptr,
                       kernel_ms);
}

int vimg_build_image_texture(const uint8_t* srgb8, uint32_t width, uint32_t height, uint32_t wrap_u, uint32_t wrap_v,
                             int device_id, float* out_level0, float* out_levels, double* kernel_ms) {
  if (width == 0 || height == 0) return fail(VIMG_ERR_INVALID_ARGUMENT, "empty image");
  if (wrap_u > VIMG_WRAP_REPEAT || wrap_v > VIMG_WRAP_REPEAT) return fail(VIMG_ERR_INVALID_ARGUMENT, "unknown wrap mode");
  return convert_bytes("vimg_build_image_texture", 2, srgb8, width, (uint64_t)width * height, height, 0.f, wrap_u, wrap_v,
                       device_id, out_level0, out_levels, kernel_ms);
}

int vimg_libm_eval(uint32_t fn, const float* x, const float* y, uint64_t n, int device_id, float* out) {
  if (fn > 9) return fail(VIMG_ERR_INVALID_ARGUMENT, "unknown function");
  if (n == 0) return VIMG_OK;
  const bool two = fn == 4 || fn == 6;
  if (!x || !out || (two && !y)) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument");
  int sms = 0, rc = prep_device(device_id, &sms);
  if (rc != VIMG_OK) return rc;
  float *d_x = nullptr, *d_y = nullptr, *d_o = nullptr;
  cudaError_t e = cudaMalloc(&d_x, n * sizeof(float));
  if (e == cudaSuccess && two) e = cudaMalloc(&d_y, n * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_o, n * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(d_x, x, n * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && two) e = cudaMemcpy(d_y, y, n * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    libm_eval_kernel<<<prep_grid(n, sms), 256>>>(fn, d_x, d_y, n, d_o);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, d_o, n * sizeof(float), cudaMemcpyDeviceToHost);
  for (float* p : {d_x, d_y, d_o})
    if (p) cudaFree(p);
  if (e != cudaSuccess) return fail(VIMG_ERR_CUDA, std::string("vimg_libm_eval: ") + cudaGetErrorString(e));
  return VIMG_OK;
}

int vimg_build_mip_chain(const float* level0, uint32_t width, uint32_t height, uint32_t wrap_u, uint32_t wrap_v,
                         int device_id, float* out_levels, double* kernel_ms) {
  if (!level0 || width == 0 || height == 0) return fail(VIMG_ERR_INVALID_ARGUMENT, "null / empty image");
  if (wrap_u > VIMG_WRAP_REPEAT || wrap_v > VIMG_WRAP_REPEAT) return fail(VIMG_ERR_INVALID_ARGUMENT, "unknown wrap mode");
  if (kernel_ms) *kernel_ms = 0.0;
  uint64_t floats = 0;
  const uint32_t n_levels = vimg_mip_chain_levels(width, height, &floats);
  if (n_levels <= 1) return VIMG_OK;
  if (!out_levels) return fail(VIMG_ERR_INVALID_ARGUMENT, "null output");
  int sms = 0, rc = prep_device(device_id, &sms);
  if (rc != VIMG_OK) return rc;
  float* d_all = nullptr;  // level 0 followed by levels 1 .. n-1
  const size_t l0 = (size_t)width * height * 3;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMalloc(&d_all, (l0 + floats) * sizeof(float));
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) e = cudaMemcpy(d_all, level0, l0 * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    cudaEventRecord(e0, 0);
    launch_mip_levels(d_all, width, height, n_levels, wrap_u, wrap_v);
    cudaEventRecord(e1, 0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out_levels, d_all + l0, floats * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && kernel_ms) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *kernel_ms = ms;
  }
  if (d_all) cudaFree(d_all);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (e != cudaSuccess) return fail(VIMG_ERR_CUDA, std::string("vimg_build_mip_chain: ") + cudaGetErrorString(e));
  return VIMG_OK;
}

// ArraySampling2D on an image that is already in HBM: the H sin(elevation) factors come from the host libm
// (sampling.h:181-183: v = (y + 0.5f) / height; float sin_elevation = std::sin(std::numbers::pi * v))
static cudaError_t launch_envmap_cdf(const float* d_img, uint32_t width, uint32_t height, float* d_col, float* d_row) {
  std::vector<float> sin_elev(height);
  for (uint32_t y = 0; y < height; ++y) {
    const float v = (static_cast<float>(y) + 0.5f) / static_cast<float>(height);
    sin_elev[y] = (float)std::sin(std::numbers::pi * v);
  }
  float *d_sin = nullptr, *d_int = nullptr;
  cudaError_t e = cudaMalloc(&d_sin, height * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_int, height * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(d_sin, sin_elev.data(), height * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    prep_row_cdf_kernel<<<height, 32>>>(d_img, width, height, d_sin, d_col, d_int);
    prep_row_prob_kernel<<<1, 32>>>(d_int, height, d_row);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(0);
  if (d_sin) cudaFree(d_sin);
  if (d_int) cudaFree(d_int);
  return e;
}

int vimg_build_envmap_cdf(const float* image, uint32_t width, uint32_t height, int device_id, float* out_row_cdf,
                          float* out_col_cdfs, double* kernel_ms) {
  if (!image || !out_row_cdf || !out_col_cdfs || width == 0 || height == 0)
    return fail(VIMG_ERR_INVALID_ARGUMENT, "null / empty argument");
  if (kernel_ms) *kernel_ms = 0.0;
  int sms = 0, rc = prep_device(device_id, &sms);
  if (rc != VIMG_OK) return rc;
  // sampling.h:181-183: v = (y + 0.5f) / height; float sin_elevation = std::sin(std::numbers::pi * v)
  std::vector<float> sin_elev(height);
  for (uint32_t y = 0; y < height; ++y) {
    const float v = (static_cast<float>(y) + 0.5f) / static_cast<float>(height);
    sin_elev[y] = (float)std::sin(std::numbers::pi * v);
  }
  const size_t n_img = (size_t)width * height * 3, n_col = (size_t)height * (width + 1);
  float *d_img = nullptr, *d_sin = nullptr, *d_col = nullptr, *d_int = nullptr, *d_row = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMalloc(&d_img, n_img * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_sin, height * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_col, n_col * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_int, height * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_row, ((size_t)height + 1) * sizeof(float));
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) e = cudaMemcpy(d_img, image, n_img * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_sin, sin_elev.data(), height * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    cudaEventRecord(e0, 0);
    prep_row_cdf_kernel<<<height, 32>>>(d_img, width, height, d_sin, d_col, d_int);
    prep_row_prob_kernel<<<1, 32>>>(d_int, height, d_row);
    cudaEventRecord(e1, 0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out_col_cdfs, d_col, n_col * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_row_cdf, d_row, ((size_t)height + 1) * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && kernel_ms) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *kernel_ms = ms;
  }
  for (float* p : {d_img, d_sin, d_col, d_int, d_row})
    if (p) cudaFree(p);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (e != cudaSuccess) return fail(VIMG_ERR_CUDA, std::string("vimg_build_envmap_cdf: ") + cudaGetErrorString(e));
  return VIMG_OK;
}

int vimg_heatmap(vimg_scene* s, uint32_t samples, float factor, float* out_rgb, float* out_cost) {
  if (!s || !out_rgb || samples == 0) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument / zero samples");
  CU_TRY(cudaSetDevice(s->device));
  if (factor <= 0) factor = 20.f;  // heatmap.cpp:135-137
  const size_t n = (size_t)s->W * s->H;
  float* d_cost = nullptr;
  if (out_cost) CU_TRY(cudaMalloc(&d_cost, n * sizeof(float)));
  const RenderArgs a = make_args(s, samples, 0, 0, 1);
  heatmap_kernel<<<(a.n_local_tiles + 1) / 2, 128>>>(s->d, a, factor, s->d_frame, d_cost);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out_rgb, s->d_frame, n * 3 * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && out_cost) e = cudaMemcpy(out_cost, d_cost, n * sizeof(float), cudaMemcpyDeviceToHost);
  if (d_cost) cudaFree(d_cost);
  if (e != cudaSuccess) return fail(VIMG_ERR_CUDA, std::string("heatmap: ") + cudaGetErrorString(e));
  return VIMG_OK;
}

// the 256 OETF + quantisation thresholds of a device (vimg_post.cuh), built at first use
static std::map<int, uint32_t*> g_post_thresholds;
static const uint32_t* post_thresholds(int device, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  uint32_t*& t = g_post_thresholds[device];
  if (!t) {
    if (cudaMalloc(&t, 256 * sizeof(uint32_t)) != cudaSuccess) {
      cudaGetLastError();
      t = nullptr;
      return nullptr;  // the pow path needs no table
    }
    post_build_thresholds_kernel<<<1, 256, 0, st>>>(t);
    cudaStreamSynchronize(st);  // once per device: later calls may come on any stream
  }
  return t;
}

// tonemap + OETF + quantise of `n_pixels` pixels already in HBM, on `st`
static int post_launch(const float* d_rgb, uint64_t n_pixels, uint32_t tonemapper, float* d_srgb, uint8_t* d_rgb8,
                       uint32_t* d_max_bits, int sm_count, cudaStream_t st) {
  if (n_pixels == 0) return VIMG_OK;
  int device = 0;
  cudaGetDevice(&device);
  // bytes only: threshold search instead of pow (VIMG_POST_POW=1 forces the pow path)
  const uint32_t* d_thr = (d_srgb == nullptr && env_u32("VIMG_POST_POW", 0) == 0) ? post_thresholds(device, st) : nullptr;
  const int threads = 256;
  const uint64_t quads = (n_pixels + 3) / 4;
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((quads + threads - 1) / threads, (uint64_t)sm_count * 16));
  if (tonemapper == VIMG_TONEMAP_REINHARD) {
    CU_TRY(cudaMemsetAsync(d_max_bits, 0, sizeof(uint32_t), st));
    post_max_luminance_kernel<<<sm_count * 8, threads, 0, st>>>(d_rgb, n_pixels, d_max_bits);
  }
  switch (tonemapper) {
    case VIMG_TONEMAP_CLAMP: post_kernel<VIMG_TONEMAP_CLAMP><<<grid, threads, 0, st>>>(d_rgb, n_pixels, d_max_bits, d_thr, d_srgb, d_rgb8); break;
    case VIMG_TONEMAP_AGX: post_kernel<VIMG_TONEMAP_AGX><<<grid, threads, 0, st>>>(d_rgb, n_pixels, d_max_bits, d_thr, d_srgb, d_rgb8); break;
    case VIMG_TONEMAP_REINHARD: post_kernel<VIMG_TONEMAP_REINHARD><<<grid, threads, 0, st>>>(d_rgb, n_pixels, d_max_bits, d_thr, d_srgb, d_rgb8); break;
    case VIMG_TONEMAP_ACES: post_kernel<VIMG_TONEMAP_ACES><<<grid, threads, 0, st>>>(d_rgb, n_pixels, d_max_bits, d_thr, d_srgb, d_rgb8); break;
    default: post_kernel<VIMG_TONEMAP_NONE><<<grid, threads, 0, st>>>(d_rgb, n_pixels, d_max_bits, d_thr, d_srgb, d_rgb8); break;
  }
  CU_TRY(cudaGetLastError());
  return VIMG_OK;
}

int vimg_postprocess(const float* rgb, uint64_t n_pixels, uint32_t tonemapper, int device_id, uint8_t* out_rgb8,
                     float* out_srgb, double* kernel_ms) {
  if (!rgb || (!out_rgb8 && !out_srgb)) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument");
  if (tonemapper > VIMG_TONEMAP_NONE) return fail(VIMG_ERR_INVALID_ARGUMENT, "unknown tonemapper");
  const int n = vimg_cuda_device_count();
  if (n == 0) return fail(VIMG_ERR_NO_DEVICE, "no CUDA device visible: the cuda render device has no CPU fallback");
  if (device_id < 0 || device_id >= n) return fail(VIMG_ERR_INVALID_ARGUMENT, "device_id out of range");
  if (kernel_ms) *kernel_ms = 0.0;
  if (n_pixels == 0) return VIMG_OK;
  CU_TRY(cudaSetDevice(device_id));
  int sms = 0;
  CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_id));
  float *d_in = nullptr, *d_f = nullptr;
  uint8_t* d_b = nullptr;
  uint32_t* d_max = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = VIMG_OK;
  auto cleanup = [&]() {
    if (d_in) cudaFree(d_in);
    if (d_f) cudaFree(d_f);
    if (d_b) cudaFree(d_b);
    if (d_max) cudaFree(d_max);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
#define PP_TRY(call)                                                                   \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      cleanup();                                                                       \
      return fail(VIMG_ERR_CUDA, std::string("vimg_postprocess: ") + cudaGetErrorString(e_)); \
    }                                                                                  \
  } while (0)
  const size_t fbytes = (size_t)n_pixels * 3 * sizeof(float);
  PP_TRY(cudaMalloc(&d_in, fbytes));
  if (out_srgb) PP_TRY(cudaMalloc(&d_f, fbytes));
  if (out_rgb8) PP_TRY(cudaMalloc(&d_b, (size_t)n_pixels * 3 + 16));
  PP_TRY(cudaMalloc(&d_max, sizeof(uint32_t)));
  PP_TRY(cudaEventCreate(&e0));
  PP_TRY(cudaEventCreate(&e1));
  PP_TRY(cudaMemcpy(d_in, rgb, fbytes, cudaMemcpyHostToDevice));
  PP_TRY(cudaEventRecord(e0, 0));
  rc = post_launch(d_in, n_pixels, tonemapper, d_f, d_b, d_max, sms, 0);
  if (rc != VIMG_OK) {
    cleanup();
    return rc;
  }
  PP_TRY(cudaEventRecord(e1, 0));
  if (out_srgb) PP_TRY(cudaMemcpy(out_srgb, d_f, fbytes, cudaMemcpyDeviceToHost));
  if (out_rgb8) PP_TRY(cudaMemcpy(out_rgb8, d_b, (size_t)n_pixels * 3, cudaMemcpyDeviceToHost));
  PP_TRY(cudaEventSynchronize(e1));
  if (kernel_ms) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *kernel_ms = ms;
  }
#undef PP_TRY
  cleanup();
  return VIMG_OK;
}

int vimg_render_srgb8(vimg_scene* s, const vimg_render_params* p, uint32_t tonemapper, uint8_t* out_rgb8,
                      vimg_render_stats* stats) {
  if (!out_rgb8) return fail(VIMG_ERR_INVALID_ARGUMENT, "null argument");
  if (tonemapper > VIMG_TONEMAP_NONE) return fail(VIMG_ERR_INVALID_ARGUMENT, "unknown tonemapper");
  int rc = check_params(s, p);
  if (rc != VIMG_OK) return rc;
  if (p->n_ranks > 1) return fail(VIMG_ERR_UNSUPPORTED, "vimg_render_srgb8 renders the whole frame (n_ranks <= 1)");
  CU_TRY(cudaSetDevice(s->device));
  const size_t n_pixels = (size_t)s->W * s->H;
  if (!s->d_rgb8) {
    CU_TRY(cudaMalloc(&s->d_rgb8, n_pixels * 3 + 16));
    CU_TRY(cudaMalloc(&s->d_post_max, sizeof(uint32_t)));
  }
  if (s->sm_count == 0) cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
  cudaEvent_t t0, t1, t2;
  CU_TRY(cudaEventCreate(&t0));
  CU_TRY(cudaEventCreate(&t1));
  CU_TRY(cudaEventCreate(&t2));
  cudaEventRecord(t0, 0);
  float* dev = nullptr;
  rc = vimg_render_device(s, p, &dev, nullptr, stats);
  if (rc == VIMG_OK) rc = post_launch(dev, n_pixels, tonemapper, nullptr, s->d_rgb8, s->d_post_max, s->sm_count, 0);
  cudaEventRecord(t1, 0);
  if (rc == VIMG_OK) {
    cudaError_t e = cudaMemcpy(out_rgb8, s->d_rgb8, n_pixels * 3, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(VIMG_ERR_CUDA, std::string("rgb8 D2H: ") + cudaGetErrorString(e));
  }
  cudaEventRecord(t2, 0);
  cudaEventSynchronize(t2);
  if (rc == VIMG_OK && stats) {
    float k = 0, t = 0;
    cudaEventElapsedTime(&k, t0, t1);
    cudaEventElapsedTime(&t, t0, t2);
    stats->kernel_ms = k;
    stats->total_ms = t;
    stats->kernel_launches += tonemapper == VIMG_TONEMAP_REINHARD ? 2 : 1;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaEventDestroy(t2);
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
  rc = launch_render<false>(s, a, p->integrator, d_px, 0, p->flags);
  if (rc != VIMG_OK) {
    cudaFree(d_px);
    return rc;
  }
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
