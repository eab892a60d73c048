/*
 * vimg_cuda.h -- C ABI of the B200 `cuda` render device for v-img.
 *
 * This is the drop-in boundary for the reference's render hot path.  The
 * reference (atom501/v-img) has no plugin/FFI layer; the seam is the C++ call
 *
 *     scene_integrator(const integrator_data&, BVH&, prims, lights, F)
 *         -> std::vector<glm::vec3>                  include/integrators.h:36-153
 *     trace_pixel(..., x, y) -> glm::vec3            include/integrators.h:181-220
 *
 * made from src/main.cpp:219-248 / :271-292 after the (unchanged) scene
 * loaders and host BVH builders have run.  A `cuda` device replaces exactly
 * that call: the host adapter flattens the already-built scene objects into
 * the POD `vimg_scene_desc` below (plain pointers + sizes, reference data
 * layouts, no C++ or torch types), uploads it once with vimg_scene_create()
 * and renders with vimg_render().  INTEGRATION.md shows the adapter TU.
 *
 * Conventions
 *   - all functions return 0 (VIMG_OK) on success, a negative vimg_status
 *     otherwise; vimg_last_error() gives a thread-local message.  Nothing
 *     throws across the ABI.  There is NO CPU fallback: without a CUDA
 *     device every compute entry point fails with VIMG_ERR_NO_DEVICE.
 *   - the caller keeps ownership of every host array in the descriptor;
 *     vimg_scene_create() deep-copies to the device(s).
 *   - one render at a time per vimg_scene (thread-compatible, like the
 *     reference's scene containers, src/main.cpp:116-120).
 *   - output images are W*H*3 floats, linear radiance averaged over spp,
 *     row 0 = TOP of the image: index = x + (H-1-y)*W
 *     (include/integrators.h:113).
 */
#ifndef VIMG_CUDA_H
#define VIMG_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIMG_ABI_VERSION 1

typedef enum vimg_status {
  VIMG_OK = 0,
  VIMG_ERR_INVALID_ARGUMENT = -1,
  VIMG_ERR_NO_DEVICE = -2,
  VIMG_ERR_CUDA = -3,
  VIMG_ERR_OUT_OF_MEMORY = -4,
  VIMG_ERR_IO = -5,
  VIMG_ERR_UNSUPPORTED = -6
} vimg_status;

/* integrator names follow src/scene_loading/json_scene.cpp:213-223 and
 * enum class integrator_func (include/integrators.h:24). */
typedef enum vimg_integrator {
  VIMG_INTEGRATOR_S_NORMAL = 0,
  VIMG_INTEGRATOR_G_NORMAL = 1,
  VIMG_INTEGRATOR_MATERIAL = 2,
  VIMG_INTEGRATOR_MIS = 3
} vimg_integrator;

/* ---- camera: TLCam private state (include/tl_camera.h:11-18) ------------ */
typedef struct vimg_camera {
  float cam_to_world[16]; /* glm::mat4, column-major (m[col][row])          */
  float p_size[2];        /* image-plane size from the TLCam ctor            */
  int32_t resolution[2];  /* width, height                                   */
  float vfov_deg;         /* vertical fov, degrees                           */
  float aperture_radius;  /* 0 => pinhole                                    */
  float focal_dist;
} vimg_camera;

/* ---- BVH: the reference's flattened layout (include/bvh.h:22-28,51-57) -- */
typedef struct vimg_bvh {
  const uint32_t* nodes;       /* n_nodes x {first_index, obj_count}; leaf iff obj_count != 0 */
  uint64_t n_nodes;
  const float* bb_mins_maxes;  /* n_bb_slots x 3 floats; root min/max at slots 0/2; the pair of
                                  children (c, c+1) at slots 2c+2..2c+5 = Lmin,Rmin,Lmax,Rmax */
  uint64_t n_bb_slots;
  const uint64_t* obj_indices; /* leaf primitive permutation (std::vector<size_t>)            */
  uint64_t n_obj_indices;
  uint32_t max_depth;
} vimg_bvh;

/* ---- primitives: list_objects order == primitive index ------------------ */
typedef enum vimg_prim_type { VIMG_PRIM_TRIANGLE = 0, VIMG_PRIM_SPHERE = 1 } vimg_prim_type;

typedef struct vimg_prim {
  uint32_t type; /* vimg_prim_type                                          */
  uint32_t a;    /* triangle: mesh index          | sphere: sphere index     */
  uint32_t b;    /* triangle: tri_index in mesh   | sphere: unused           */
} vimg_prim;

typedef struct vimg_sphere {   /* include/geometry/sphere.h:28-31 */
  float center[3];
  float radius;
  uint32_t material;
} vimg_sphere;

#define VIMG_NO_UV 255u        /* MeshConsts::no_uv, include/geometry/mesh.h:10-12 */

typedef struct vimg_mesh {     /* include/geometry/mesh.h:14-57 */
  const uint32_t* indices;     /* n_tris x 3                                */
  uint64_t n_tris;
  const float* vertices;       /* n_verts x 3                               */
  uint64_t n_verts;
  const float* normals;        /* n_verts x 3 or NULL (Mesh::normals empty)  */
  const float* const* texcoords; /* n_uv_sets arrays of n_verts x 2         */
  uint32_t n_uv_sets;
  uint8_t color_tex_uv;        /* index into texcoords or VIMG_NO_UV         */
  uint8_t normal_tex_uv;
  uint8_t metallic_roughness_tex_uv;
  uint8_t pad_;
  uint32_t material;
} vimg_mesh;

/* ---- textures ----------------------------------------------------------- */
typedef enum vimg_wrap { /* TextureWrappingMode, include/texture/texture_common.h:7 */
  VIMG_WRAP_CLAMP_TO_EDGE = 0,
  VIMG_WRAP_MIRRORED_REPEAT = 1,
  VIMG_WRAP_REPEAT = 2
} vimg_wrap;

typedef enum vimg_texture_type {
  VIMG_TEX_CONST = 0,   /* ConstColor    include/texture/texture_RGB.h:45-57 */
  VIMG_TEX_CHECKER = 1, /* Checkerboard  include/texture/texture_RGB.h:59-81 */
  VIMG_TEX_IMAGE = 2    /* ImageTexture  include/texture/texture_RGB.h:83-150 */
} vimg_texture_type;

typedef struct vimg_texture {
  uint32_t type;
  float col_a[3];        /* const: albedo | checker: col_a                   */
  float col_b[3];        /* checker: col_b                                   */
  uint32_t width;        /* checker: width | image: level-0 width            */
  uint32_t height;
  uint32_t n_levels;     /* image: mipmap.size()                             */
  uint32_t wrap_u, wrap_v;
  const float* const* levels; /* image: n_levels arrays, level l is
                                 max(W>>l,1) x max(H>>l,1) x 3 floats, row 0 = top.
                                 levels[1 .. n_levels-1] may ALL be NULL: vimg_scene_create then
                                 builds them on the device from level 0, bit for bit what
                                 ImageTexture's constructor builds (image_texture.cpp:59-130);
                                 n_levels must be vimg_mip_chain_levels(width, height)          */
} vimg_texture;

typedef struct vimg_texture_rg { /* TextureRG include/texture/texture_RG.h:10-57 */
  uint32_t width, height;
  uint32_t wrap_u, wrap_v;
  const float* data;     /* width*height x 2 floats                          */
} vimg_texture_rg;

/* ---- materials ---------------------------------------------------------- */
typedef enum vimg_material_type {
  VIMG_MAT_LAMBERTIAN = 0,    /* src/material/lambertian.cpp                  */
  VIMG_MAT_DIFFUSE_LIGHT = 1, /* include/material/diffuse_light.h             */
  VIMG_MAT_DIELECTRIC = 2,    /* src/material/dielectric.cpp                  */
  VIMG_MAT_PRINCIPLED = 3,    /* include/material/principled.h                */
  VIMG_MAT_BASE = 4           /* Material base defaults, material.h:25-72     */
} vimg_material_type;

#define VIMG_NO_TEXTURE (-1)

typedef struct vimg_material {
  uint32_t type;
  int32_t texture;       /* TextureRGB* -> index into textures, or -1        */
  int32_t texture_rg;    /* Principled::metallic_roughness or -1             */
  int32_t normal_map;    /* Material::normal_map (image texture index) or -1 */
  float emit[3];         /* DiffuseLight::emit                               */
  float ior;             /* Dielectric::ior                                  */
  /* Principled (include/material/principled.h:25-40) */
  float metallic_roughness_factor[2];
  float specular_transmission;
  float subsurface;
  float specular;
  float specular_tint;
  float anisotropic;
  float sheen;
  float sheen_tint;
  float clearcoat;
  float clearcoat_gloss;
  float eta;
} vimg_material;

/* ---- lights and background ---------------------------------------------- */
#define VIMG_LIGHT_BACKGROUND (-1) /* Emitter::is_background(), emitters.h:25 */

typedef enum vimg_background_type {
  VIMG_BG_CONST = 0,  /* ConstBackground include/background.h:25-65          */
  VIMG_BG_ENVMAP = 1  /* EnvMap          include/background.h:68-179         */
} vimg_background_type;

typedef struct vimg_background {
  uint32_t type;
  float color[3];          /* ConstBackground::col                           */
  int32_t image;           /* EnvMap::image -> index into textures           */
  float world_to_env[16];  /* column-major                                   */
  float env_to_world[16];
  float radiance_scale;
  uint32_t cdf_width;      /* ArraySampling2D::width / height                */
  uint32_t cdf_height;
  const float* row_cdf;    /* cdf_height + 1 floats  (row_probabilities.cdf) */
  const float* col_cdfs;   /* cdf_height x (cdf_width + 1) floats.  Both may be NULL: built on
                              the device from level 0 of `image` (ArraySampling2D's constructor,
                              include/rng/sampling.h:167-197); cdf_width / cdf_height 0 or the image's */
} vimg_background;

/* ---- the whole scene ---------------------------------------------------- */
typedef struct vimg_scene_desc {
  uint32_t abi_version;    /* VIMG_ABI_VERSION */
  vimg_camera camera;
  vimg_bvh bvh;
  const vimg_prim* prims;           uint64_t n_prims;
  const vimg_sphere* spheres;       uint32_t n_spheres;
  const vimg_mesh* meshes;          uint32_t n_meshes;
  const vimg_material* materials;   uint32_t n_materials;
  const vimg_texture* textures;     uint32_t n_textures;
  const vimg_texture_rg* textures_rg; uint32_t n_textures_rg;
  const int32_t* lights;            uint32_t n_lights; /* list_lights order: prim index or
                                                          VIMG_LIGHT_BACKGROUND */
  vimg_background background;
} vimg_scene_desc;

/* ---- render call -------------------------------------------------------- */
typedef enum vimg_partition {
  VIMG_PARTITION_TILE = 0, /* interleaved 8x8 tiles: tile i -> rank i % n_ranks
                              (mirrors include/integrators.h:57-65,101).  A rank writes its own pixels
                              and leaves the others 0: the sum over ranks is the one-GPU frame, bit for bit */
  VIMG_PARTITION_SPP = 1   /* sample ranges (material / mis, spp_streams >= n_ranks): every rank renders EVERY
                              pixel, the sample ranges rank, rank + n_ranks, ... of it (see spp_streams), and
                              writes (sum of its ranges) / spp.  The sum over ranks is the spp_streams image up
                              to the association of K float additions (ranks' sums are added instead of ranges) */
} vimg_partition;

typedef struct vimg_render_params {
  uint32_t integrator;   /* vimg_integrator                                  */
  uint32_t samples;      /* spp (integrator_data::samples)                   */
  uint32_t depth;        /* integrator_data::depth; UINT32_MAX for "depth": -1 */
  uint32_t rank;         /* this process renders tiles rank, rank+n_ranks, ... */
  uint32_t n_ranks;      /* 0 or 1 => whole frame                            */
  uint32_t flags;        /* VIMG_FLAG_*                                      */
  uint32_t spp_streams;  /* material / mis, queue wavefront only.  0 or 1: the reference's ONE PCG
                            stream per pixel carried over all its samples (integrators.h:116-131),
                            images bit-identical to the CPU renderer.  K > 1: the pixel's samples are
                            split into K contiguous ranges [k*spp/K, (k+1)*spp/K), range k drawing
                            from its own PCG stream (initstate = image_index, initseq = k; range 0
                            starts exactly like the reference) with the reference's R2 jitter index
                            x+y+s, so K paths per pixel are in flight instead of one.  Same estimator,
                            different random numbers: the image agrees with the reference within the
                            stated relMSE tolerance, not bit for bit.  Partial sums are combined in
                            range order (deterministic, independent of n_ranks).               */
  uint32_t partition;    /* vimg_partition: what rank / n_ranks divide                        */
} vimg_render_params;

#define VIMG_FLAG_COUNTERS 1u /* also gather traversal counters (slower)     */
/* material / mis pipeline choice (default: queue wavefront; the env var
 * VIMG_PIPELINE=block|queue overrides the default when neither flag is set).
 * Both produce the same image bit for bit. */
#define VIMG_FLAG_PIPELINE_BLOCK 2u /* persistent block-local wavefront kernel  */
#define VIMG_FLAG_PIPELINE_QUEUE 4u /* multi-kernel queue wavefront             */
/* queue wavefront only: run the launches one after the other on one stream with a CUDA event
 * after each, and report per-stage device time in vimg_render_stats.stage_ms (measurement
 * mode: slower than the normal, concurrent schedule) */
#define VIMG_FLAG_STAGE_TIMES 8u

/* material / mis: every BVH traversal starts from an empty stack.  The reference does not: its
 * per-thread `thread_stack` is shared by BVH::occlude and BVH::hit, BVH::occlude returns on the first
 * hit without emptying it (include/bvh.h:165-168) and BVH::hit pushes the root on top of what is there
 * (bvh.h:143), so the closest-hit traversal of a path's next segment also visits the nodes its shadow
 * ray left behind (mis_integrator.cpp:64 -> :119) -- without their box test.  By default (one RNG
 * stream per pixel) the device reproduces that, which is what makes mis images bit-identical where the
 * slab test is not conservative (self-intersections at t just above minT); it costs the extra node
 * visits the reference also pays (measured: 8-13 % of the frame).  This flag drops it; spp_streams > 1
 * (not bit-identical by construction) implies it. */
#define VIMG_FLAG_FRESH_STACK 16u

/* material / mis, queue wavefront: the TOLERANCE mode.  The north star's contract for material / mis
 * images is a relative-MSE tolerance against the CPU renderer (bit-exactness is asked of primary-hit
 * ids and the normal integrators, which this flag does not touch).  With this flag the device may
 * (a) evaluate the path arithmetic with the hardware's approximate division / sqrt / sin / cos / log /
 * pow units, FMA contraction, and float where the reference promotes to double -- the kernels of
 * vimg_fast.cu, the same sources compiled with -use_fast_math, as the reference's own CMake build
 * contracts FMAs with -march=native -- and (b) trace through a device-side re-layout of the BVH
 * instead of the reference's node order: scenes of a few dozen primitives are traced over their
 * leaf array by a divergence-free kernel (qf_trace_flat_kernel).  Same estimator, same RNG streams
 * and draw order per path; images agree with the CPU renderer within the stated tolerance
 * (tests/test_gpu_fast.py), not bit for bit.  Implies VIMG_FLAG_FRESH_STACK. */
#define VIMG_FLAG_FAST 32u

/* stages of the queue wavefront (vimg_render_stats.stage_ms / stage_launches) */
typedef enum vimg_stage {
  VIMG_STAGE_TRACE = 0,            /* BVH::hit / BVH::occlude for every queued ray          */
  VIMG_STAGE_CLASSIFY = 1,         /* route each traced path by the material it hit        */
  VIMG_STAGE_SHADE_PRINCIPLED = 2, /* light sampling + BSDF per material class             */
  VIMG_STAGE_SHADE_LAMBERTIAN = 3,
  VIMG_STAGE_SHADE_DIELECTRIC = 4,
  VIMG_STAGE_END = 5,              /* sample accumulation, next pixel, next camera ray     */
  VIMG_STAGE_COUNT = 6
} vimg_stage;

typedef struct vimg_render_stats {
  uint64_t rays_closest;   /* BVH::hit<optional<HitInfo>> calls               */
  uint64_t rays_shadow;    /* BVH::occlude calls                              */
  uint64_t inner_nodes;    /* inner nodes popped   (VIMG_FLAG_COUNTERS only)  */
  uint64_t leaf_nodes;     /* leaves popped        (VIMG_FLAG_COUNTERS only)  */
  uint64_t tri_tests;      /* triangle tests       (VIMG_FLAG_COUNTERS only)  */
  uint64_t sphere_tests;   /* sphere tests         (VIMG_FLAG_COUNTERS only)  */
  uint64_t nan_samples;    /* samples with a NaN channel (integrators.h:128)  */
  uint64_t samples;        /* pixel samples taken by this rank                */
  double kernel_ms;        /* device time of the render kernel(s), CUDA events */
  double total_ms;         /* kernel + D2H of the frame                       */
  uint32_t kernel_launches;
  uint32_t stage_launches[VIMG_STAGE_COUNT]; /* VIMG_FLAG_STAGE_TIMES only                */
  double stage_ms[VIMG_STAGE_COUNT];         /* VIMG_FLAG_STAGE_TIMES only                */
  uint64_t stage_items[VIMG_STAGE_COUNT];    /* queue wavefront: items the stage's launches processed (rays for
                                                TRACE, path segments for SHADE_*, ended paths for END)   */
} vimg_render_stats;

typedef struct vimg_scene vimg_scene; /* opaque device-resident scene */

/* number of CUDA devices visible (0 if none / no driver).  Never fails. */
int vimg_cuda_device_count(void);

/* thread-local message of the last failing call ("" if none). */
const char* vimg_last_error(void);

/* deep-copy `desc` to device `device_id` and build the GPU layout. */
int vimg_scene_create(const vimg_scene_desc* desc, int device_id, vimg_scene** out_scene);
void vimg_scene_destroy(vimg_scene* scene);

/* The material / mis pipeline keeps its render workspace (path-state pools, queues, partial
 * sums: up to ~1 GB, sized by the frame, independent of the scene) in a process-wide cache, one
 * idle workspace per device, so that creating a scene per frame does not allocate inside every
 * render; the device allocation that holds a destroyed scene's arrays and frame is kept the same
 * way (one per device) for the next vimg_scene_create().  This frees what is idle (scenes that
 * are still alive keep theirs). */
void vimg_release_workspaces(void);

/* bytes of device memory the scene occupies (BVH + geometry + textures). */
uint64_t vimg_scene_device_bytes(const vimg_scene* scene);

/* Replaces scene_integrator<F>() (include/integrators.h:36-153).
 * out_rgb: HOST buffer of W*H*3 floats (row 0 = top).  With n_ranks > 1 only
 * this rank's tiles are written, every other pixel is 0 (so a sum over ranks,
 * e.g. an NCCL reduce, assembles the frame exactly).  The call includes the
 * device->host copy of the frame. */
int vimg_render(vimg_scene* scene, const vimg_render_params* params, float* out_rgb,
                vimg_render_stats* stats /* nullable */);

/* The whole frame on several GPUs of one box, behind one call (what a maintainer's `--device cuda:0,1,...`
 * would call from the seam at src/main.cpp:219-248).  The scene is replicated (one host thread per device
 * uploads it), the work divided as `partition` says, and the image assembled in the memory of the FIRST device:
 *   VIMG_PARTITION_TILE: the other devices' kernels store their finished pixels straight into the first
 *     device's frame through peer-mapped memory (NVLink / NVSwitch); no reduce, no staging -- the stores ARE
 *     the gather.  Bit-identical to the one-device frame.
 *   VIMG_PARTITION_SPP: every device renders its sample ranges of every pixel into its own frame; the first
 *     device then adds the peers' frames in device order with one kernel that reads them over NVLink.
 * Without peer access between two of the devices the partial frames go through host memory instead.
 * params->rank / n_ranks / partition are ignored (n_ranks = n_devices).  out_rgb: HOST buffer, W*H*3 floats.
 * stats (nullable): counters summed over devices, kernel_ms = the slowest device, total_ms = wall time of the
 * call from after the scene uploads (render + assembly + D2H). */
int vimg_render_multi(const vimg_scene_desc* desc, const int* device_ids, uint32_t n_devices,
                      const vimg_render_params* params, uint32_t partition, float* out_rgb,
                      vimg_render_stats* stats /* nullable */);

/* Same, but the frame stays in device memory (for NCCL / peer stores and for
 * timing without the D2H copy).  *out_dev_rgb is owned by the scene and valid
 * until the next render or vimg_scene_destroy().  `stream` is a cudaStream_t
 * (NULL = default stream); the call returns after the stream is synchronised. */
int vimg_render_device(vimg_scene* scene, const vimg_render_params* params,
                       float** out_dev_rgb, void* stream, vimg_render_stats* stats);

/* Replaces trace_pixel<F>() (include/integrators.h:181-220): one pixel, image
 * coordinates as in the reference's -d "x y" (y = 0 is the bottom row). */
int vimg_trace_pixel(vimg_scene* scene, const vimg_render_params* params, int x, int y,
                     float out_rgb[3]);

/* Parity hook: primitive index of the closest hit of every camera ray
 * (-1 = miss), for all `samples` samples of every pixel, using the same
 * per-pixel RNG stream and R2 jitter as the normal integrators.
 * out_prim_id: HOST buffer of W*H*samples int32, index = image_index*samples + s.
 * out_t (nullable): same shape, the hit distance ray.maxT (inf on miss). */
int vimg_trace_ids(vimg_scene* scene, uint32_t samples, int32_t* out_prim_id, float* out_t);

/* Replaces heatmap_img() (src/integrators/heatmap.cpp:39-147; main.cpp:249-255 calls it with 4
 * spp for the -m flag): per pixel the cost BVH::hit<float> accumulates over the integrators'
 * camera rays (0.5 root box + 1 per inner node + 1 per primitive test on the reference's scalar
 * path), truncated mean over `samples`, turbo colour map of mean / factor (factor <= 0 => 20).
 * out_rgb: HOST, W*H*3 floats, row 0 = top.  out_cost (nullable): HOST, W*H floats, the
 * truncated mean cost per pixel (what the colour map is applied to). */
int vimg_heatmap(vimg_scene* scene, uint32_t samples, float factor, float* out_rgb, float* out_cost);

/* ---- scene preparation (the step before the path; the reference does it in its constructors) */
/* Number of mip levels ImageTexture's constructor builds (src/image_texture.cpp:73-78):
 * min(ceil(log2(float(min(w, h)))), 15), at least 1 (level 0).  *out_floats (nullable) receives
 * the number of floats of levels 1 .. n-1 together (level l is max(w>>l,1) x max(h>>l,1) x 3). */
uint32_t vimg_mip_chain_levels(uint32_t width, uint32_t height, uint64_t* out_floats);

/* Replaces the mip-chain loop of ImageTexture::ImageTexture (image_texture.cpp:59-130):
 * level0: HOST, width*height*3 floats (linear RGB, row 0 = top); out_levels: HOST, the floats of
 * levels 1 .. n-1 one after the other (size from vimg_mip_chain_levels).  wrap_u / wrap_v:
 * vimg_wrap.  kernel_ms (nullable): device time of the level kernels. */
int vimg_build_mip_chain(const float* level0, uint32_t width, uint32_t height, uint32_t wrap_u,
                         uint32_t wrap_v, int device_id, float* out_levels, double* kernel_ms);

/* Replaces ArraySampling2D's constructor for an env map (include/rng/sampling.h:167-197, called
 * from EnvMap's constructor, background.h:83): image: HOST, width*height*3 floats;
 * out_row_cdf: HOST, height+1 floats; out_col_cdfs: HOST, height*(width+1) floats -- the
 * vimg_background fields of the same names. */
int vimg_build_envmap_cdf(const float* image, uint32_t width, uint32_t height, int device_id,
                          float* out_row_cdf, float* out_col_cdfs, double* kernel_ms);

/* Replaces make_texture's byte -> float loop (src/scene_loading/gltf_loading.cpp:139-143) followed by
 * ImageTexture::convert_sRGB_to_linear (src/image_texture.cpp:257-263; pix_sRGB_to_linear,
 * include/color_utils.h:28-47): rgb8: HOST, n_pixels * 3 bytes (the decoded image);
 * out_rgb: HOST, n_pixels * 3 floats, linear.  3 bytes per pixel cross PCIe instead of 12. */
int vimg_convert_srgb8_to_linear(const uint8_t* rgb8, uint64_t n_pixels, int device_id, float* out_rgb,
                                 double* kernel_ms);

/* Replaces the same loop followed by ImageTexture::convert_RGB_to_normal (src/image_texture.cpp:265-275):
 * p / 127.5f - 1, x and y times `scale` (glTF normalTexture.scale), glm::normalize. */
int vimg_convert_rgb8_to_normal(const uint8_t* rgb8, uint64_t n_pixels, float scale, int device_id,
                                float* out_rgb, double* kernel_ms);

/* make_texture's TextureType::Image case as a whole (gltf_loading.cpp:147-156): decoded sRGB bytes ->
 * linear level 0 -> the constructor's mip chain, without leaving the device.  out_level0: HOST,
 * width*height*3 floats; out_levels: HOST, levels 1 .. n-1 (size from vimg_mip_chain_levels). */
int vimg_build_image_texture(const uint8_t* srgb8, uint32_t width, uint32_t height, uint32_t wrap_u,
                             uint32_t wrap_v, int device_id, float* out_level0, float* out_levels,
                             double* kernel_ms);

/* Parity hook: the device's copies of the glibc float functions the path calls (std::sin, std::cos,
 * std::log, std::log2, std::pow, std::acos, std::atan2 on floats: include/rng/sampling.h, background.h,
 * tl_camera.cpp, color_utils.h, agx.cpp, principled BSDF), evaluated on caller-given inputs so that a
 * test can compare them with the host libm bit for bit.
 * fn: 0 sinf, 1 cosf, 2 logf, 3 log2f, 4 powf(x, y), 5 acosf, 6 atan2f(x = y-argument, y = x-argument),
 * 7 atanf; 8 / 9: the low / high 32 bits of the DOUBLE log2((double)x) (std::log2 on doubles:
 * background.h:100), stored as float bit patterns.  x, y (y only for fn 4 and 6), out: HOST, n floats. */
int vimg_libm_eval(uint32_t fn, const float* x, const float* y, uint64_t n, int device_id, float* out);

/* ---- post-process chain (the step right after the path: src/main.cpp:304-356) ----------- */
/* enum class tonemapper (include/color_utils.h:7) + "no tonemapper" (main.cpp:325-327). */
typedef enum vimg_tonemapper {
  VIMG_TONEMAP_CLAMP = 0,    /* simple_clamp   include/color_utils.h:21-25 */
  VIMG_TONEMAP_AGX = 1,      /* agx            src/tonemap/agx.cpp          */
  VIMG_TONEMAP_REINHARD = 2, /* reinhard_lum   src/tonemap/reinhard.cpp     */
  VIMG_TONEMAP_ACES = 3,     /* aces           src/tonemap/aces.cpp         */
  VIMG_TONEMAP_NONE = 4
} vimg_tonemapper;

/* Tonemapper, sRGB OETF (color_utils.h:50-68) and main.cpp's 8-bit quantisation (NaN pixel ->
 * magenta, int(255.999 * v) clamped; main.cpp:339-356) of a linear image, on device `device_id`.
 * rgb: HOST, n_pixels * 3 floats.  out_rgb8 (nullable): HOST, n_pixels * 3 bytes = the PNG
 * payload.  out_srgb (nullable): HOST, n_pixels * 3 floats, the image after gamma (what
 * main.cpp prints for -d).  kernel_ms (nullable): device time of the kernels. */
int vimg_postprocess(const float* rgb, uint64_t n_pixels, uint32_t tonemapper, int device_id,
                     uint8_t* out_rgb8, float* out_srgb, double* kernel_ms);

/* scene_integrator + the post-process chain without leaving the device: renders like
 * vimg_render() (single rank) and copies back W*H*3 BYTES instead of W*H*12.
 * stats->kernel_ms covers render + post-process kernels, total_ms adds the copy. */
int vimg_render_srgb8(vimg_scene* scene, const vimg_render_params* params, uint32_t tonemapper,
                      uint8_t* out_rgb8, vimg_render_stats* stats /* nullable */);

/* ---- scene blob I/O (flattened scene cache; used for test fixtures) ------ */
/* serialise a descriptor to a single relocatable file and back.  A loaded
 * scene owns its arrays; free with vimg_scene_desc_free(). */
int vimg_scene_desc_save(const vimg_scene_desc* desc, const char* path);
int vimg_scene_desc_load(const char* path, vimg_scene_desc** out_desc);
void vimg_scene_desc_free(vimg_scene_desc* desc);

#ifdef __cplusplus
}
#endif
#endif /* VIMG_CUDA_H */
