// vimg_types.cuh -- device-resident scene layout (HBM) and per-ray state.
//
// Layout decisions (DESIGN.md "Data layout in HBM"):
//  * BVH: one 64-byte, 64-byte-aligned record per INNER node = the reference's
//    sibling-pair slots [Lmin,Rmin,Lmax,Rmax] (include/bvh.h:171-176, 48 B) +
//    both children's {first_index,obj_count} (include/bvh.h:22-28).  One pop =
//    4 x LDG.128 of one record; no separate node fetch, no dependent second load.
//    Node numbering, child order and leaf order are the reference's, unchanged,
//    so traversal visits boxes and primitives in the reference's order (bit-exact
//    ids need that).
//  * Leaf primitives: 48-byte records pre-gathered in obj_indices order (the
//    reference chases obj_indices -> Surface* -> Mesh -> indices -> vertices).
//  * Shading data: one 80-byte DShadeRec per leaf record (vertex normals, colour uv set,
//    material) and one 96-byte DLightRec per light, pre-gathered so that hit-info and light
//    sampling read ONE record with 16-byte loads instead of chasing
//    prim -> mesh -> indices -> normals / uvs (four dependent loads).  The SoA pools stay for
//    the general case (meshes whose textures use different uv sets).
#pragma once
#include <stdint.h>

#include "../../include/vimg_cuda.h"
#include "vimg_math.cuh"

namespace vimg {

struct __align__(16) DPairNode {
  float lmin[3], rmin[3], lmax[3], rmax[3];
  uint32_t l_first, l_count, r_first, r_count;
};
static_assert(sizeof(DPairNode) == 64, "pair record must be 64 bytes");

#define VIMG_LEAF_SPHERE 1u
#define VIMG_LEAF_DEGENERATE 2u /* length2(cross(e2,e1)) == 0 (triangle.h:92) */
#define VIMG_LEAF_CLASS_EMISSIVE 4u /* leaf_class: the primitive's material is a DiffuseLight */

struct __align__(16) DLeafPrim {
  float a[9];        // triangle: p0,p1,p2 | sphere: center.xyz, radius
  uint32_t prim_id;  // index in list_objects
  uint32_t flags;
  uint32_t pad_;
#ifdef LEAF_PRIM_PAD  // the address-range experiment of DESIGN.md 5.1d
  uint32_t pad2_[LEAF_PRIM_PAD / 4];
#endif
};
#ifndef LEAF_PRIM_PAD
static_assert(sizeof(DLeafPrim) == 48, "leaf record must be 48 bytes");
#endif

// shading record of a leaf entry (same index as leaf_prims)
#define VIMG_SHADE_HAS_NORMALS 1u      /* n[] valid (Mesh::normals not empty)                        */
#define VIMG_SHADE_HAS_UV 2u           /* uv[] valid: the mesh has a colour uv set                   */
#define VIMG_SHADE_NORMAL_UV_SAME 4u   /* the normal map reads the colour uv set (else: barycentrics) */
#define VIMG_SHADE_GENERAL 8u          /* metallic-roughness / normal uv sets differ from the colour
                                          set: hit-info walks the SoA pools                          */
#define VIMG_SHADE_SPHERE_NEEDS_UV 16u /* sphere whose material reads uv (non-constant texture)      */
struct __align__(16) DShadeRec {
  float n[9];   // n0, n1, n2
  float uv[6];  // uv0, uv1, uv2 of the colour set
  uint32_t material;
  uint32_t flags;
  uint32_t pad_[3];
};
static_assert(sizeof(DShadeRec) == 80, "shade record must be 80 bytes");

// light record, list_lights order (GroupOfEmitters::sample indexes it, emitters.h:39-56)
#define VIMG_LIGHTREC_TRIANGLE 0u
#define VIMG_LIGHTREC_SPHERE 1u
#define VIMG_LIGHTREC_BACKGROUND 2u
struct __align__(16) DLightRec {
  float p[9];  // triangle: p0, p1, p2 | sphere: center.xyz, radius
  float n[9];  // triangle vertex normals (has_normals)
  uint32_t type;
  uint32_t material;
  uint32_t has_normals;
  uint32_t pad_[3];
};
static_assert(sizeof(DLightRec) == 96, "light record must be 96 bytes");

struct DMesh {
  uint32_t vert_base;    // into vertices / normals pools
  uint32_t tri_base;     // into tri_indices pool (per triangle, 3 x u32 local indices)
  uint32_t has_normals;
  int32_t uv_color;      // float2 offset of the uv set into the uv pool, or -1
  int32_t uv_normal;
  int32_t uv_mr;
  uint32_t material;
};

struct DTexture {
  uint32_t type;
  float col_a[3], col_b[3];
  uint32_t width, height, n_levels, wrap_u, wrap_v;
  uint32_t level_off[16];  // texel (float3) offset of each mip level in the texel pool
  // 0.5f * log2(width * height) of compute_texture_LOD (texture_RGB.h:141): the unqualified log2 on
  // an unsigned is glibc's DOUBLE log2; one value per texture, computed once when the scene is
  // created with the restatement of vimg_libm.cuh (CUDA's log2(4096.0) is not exactly 12, which moved
  // the LOD of a 64x64 texture by one float ulp whenever the sum fell on a rounding tie)
  double half_log2_wh;
};

struct DTextureRG {
  uint32_t width, height, wrap_u, wrap_v;
  uint32_t off;  // float2 offset into the rg pool
};

struct DCamera {
  float m[16];
  float p_size[2];
  int32_t res[2];
  float aperture_radius, focal_dist;
  float cone_spread;  // raycone_for_primary_ray(): constant per camera (ray.h:44-48)
};

struct DBackground {
  uint32_t type;
  float color[3];
  int32_t image;
  float world_to_env[16], env_to_world[16];
  float radiance_scale;
  uint32_t cdf_w, cdf_h;
  const float* row_cdf;
  const float* col_cdfs;
  uint32_t emissive;  // Background::is_emissive()
};

struct DCounters {  // matches the tail of vimg_render_stats
  unsigned long long rays_closest, rays_shadow, inner_nodes, leaf_nodes, tri_tests, sphere_tests,
      nan_samples, samples;
  unsigned long long stage_clk[8];  // WF_STAGE_CLOCKS builds only: per-stage SM cycles summed over warps
  unsigned long long stage_items[8];  // queue wavefront: items each stage's launches processed (index = vimg_stage)
};

struct DScene {
  DCamera cam;
  // BVH
  const DPairNode* pairs;  // (n_nodes - 1) / 2 records, record of children (c, c+1) at (c-1)/2
  const DLeafPrim* leaf_prims;
  const uint8_t* leaf_class;  // per leaf record: material class (bits 0-1) | VIMG_LEAF_CLASS_EMISSIVE
  const DShadeRec* shade_recs;  // per leaf record
  const DLightRec* light_recs;  // per light
  float root_min[3], root_max[3];
  uint32_t root_first, root_count;  // nodes[0]
  uint32_t n_nodes;
  uint32_t stack_cap;
  uint32_t n_leaf_recs;  // entries of leaf_prims / leaf_class / shade_recs (= BVH::obj_indices.size())
  // primitives / geometry pools
  const vimg_prim* prims;
  const vimg_sphere* spheres;
  const DMesh* meshes;
  const uint32_t* tri_indices;
  const float* vertices;  // xyz
  const float* normals;   // xyz
  const float* uvs;       // uv
  // shading
  const vimg_material* materials;
  const DTexture* textures;
  const float* texels;  // rgb
  const DTextureRG* textures_rg;
  const float* rg_texels;
  const int32_t* lights;
  uint32_t n_lights;
  // ray cones (ray.h:44-174) only feed the mip level of image textures (col_at_ray_hit) and of the
  // env map (background_emit); a scene with neither never reads them, so their propagation (two
  // double sin / cos pairs and an acos per refraction) is skipped.  Images are unchanged.
  uint32_t uses_cones;
  DBackground bg;
};

struct RayCone {
  float cone_width, spread_angle;
};

// Ray of include/ray.h:16-42
struct Ray {
  V3 dir, o;
  float minT, maxT;
  RayCone cone;
};

struct ONB {
  V3 u, v, w;
};

// HitInfo of include/hit_utils.h:61-74; `mat`/`obj` pointers become indices
struct HitInfo {
  uint32_t mat;
  int32_t prim;
  V3 hit_p, hit_n_s, hit_n_g;
  V2 uv, mr_uv;
  ONB frame;
  float primitive_area, tex_coord_area, mean_curvature;
};

// ForHitInfo of include/hit_utils.h:76-82 (prim pointer -> leaf record index)
struct PreHit {
  float e0, e1, e2, inv_det;
  int32_t leaf_slot;  // index into leaf_prims, -1 = no hit
};

struct EmitterInfo {
  V3 wi;
  float pdf, dist, G;
};

struct ScatterInfo {
  V3 wo;
  float eta;
  bool is_specular;
};

}  // namespace vimg
