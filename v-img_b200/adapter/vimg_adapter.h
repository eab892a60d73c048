// vimg_adapter.h -- reference-side binding of the `cuda` render device.
//
// This is the TU a v-img maintainer adds next to src/main.cpp (see
// INTEGRATION.md).  It includes the reference's own headers, walks the scene
// containers that the unchanged loaders and BVH builders produced
// (src/main.cpp:48-56,181-204) and flattens them into the POD
// `vimg_scene_desc` of include/vimg_cuda.h.  Nothing here computes pixels.
//
// Most scene classes keep their parameters private (Sphere sphere.h:28-31,
// TLCam tl_camera.h:11-18, Principled principled.h:25-40, ...).  The reference
// stays unmodified, so this TU is compiled with g++ -fno-access-control.
#pragma once

#include <background.h>
#include <bvh.h>
#include <geometry/emitters.h>
#include <geometry/mesh.h>
#include <geometry/surface.h>
#include <integrators.h>
#include <material/material.h>
#include <texture/texture_RG.h>
#include <texture/texture_RGB.h>
#include <vimg_cuda.h>

#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

// Owns the pointer tables a vimg_scene_desc needs; all bulk arrays alias the
// reference's own containers (no copies), so it must not outlive the scene.
struct VimgFlatScene {
  vimg_scene_desc desc{};
  std::vector<vimg_prim> prims;
  std::vector<vimg_sphere> spheres;
  std::vector<vimg_mesh> meshes;
  std::vector<vimg_material> materials;
  std::vector<vimg_texture> textures;
  std::vector<vimg_texture_rg> textures_rg;
  std::vector<int32_t> lights;
  std::vector<std::vector<const float*>> mesh_uv_ptrs;
  std::vector<std::vector<const float*>> tex_level_ptrs;
  std::vector<float> col_cdfs;  // EnvMap per-row CDFs, concatenated
  std::vector<uint64_t> obj_indices;

  // const Emitter* (== HitInfo::obj, triangle.cpp:143 / sphere.cpp:35) -> prim index
  std::unordered_map<const Emitter*, int32_t> emitter_to_prim;
};

// Flatten a loaded scene.  `lights` is main()'s list_lights (src/main.cpp:50).
// Throws std::runtime_error on a scene object the cuda device does not know.
std::unique_ptr<VimgFlatScene> vimg_flatten_scene(
    const integrator_data& settings, const BVH& bvh,
    const std::vector<std::unique_ptr<Surface>>& prims, const std::vector<Emitter*>& lights);

// Drop-in for scene_integrator<F>() (include/integrators.h:36-153): same
// inputs, same std::vector<glm::vec3> result, rendered by libvimg_cuda.so.
// `func` selects the integrator exactly like the switch in src/main.cpp:219-248.
std::vector<glm::vec3> scene_integrator_cuda(const integrator_data& settings, const BVH& bvh,
                                             const std::vector<std::unique_ptr<Surface>>& prims,
                                             const std::vector<Emitter*>& lights,
                                             integrator_func func, int device_id = 0);

// The same on several GPUs of the box (vimg_render_multi, tile partition: bit-identical to one device).
std::vector<glm::vec3> scene_integrator_cuda_multi(const integrator_data& settings, const BVH& bvh,
                                                   const std::vector<std::unique_ptr<Surface>>& prims,
                                                   const std::vector<Emitter*>& lights, integrator_func func,
                                                   const std::vector<int>& device_ids);
