// vimg_lean.cu -- the tolerance build's streamed shade / END kernels once more, for scenes that use none of the
// features VIMG_LEAN compiles out (vimg_math.cuh): no image texture, normal map, metallic-roughness texture, env map,
// ray cones, general uv walk.  Those paths never run on such a scene but they set the kernels' register need;
// without them the Lambertian / Dielectric shade kernels fit 5 blocks of 128 threads per SM (96 registers, no
// spills) instead of 4, Principled 4 instead of 3 and END 8 (64 registers) instead of 6.  Same sources, namespace
// vimg_lean; vimg_api.cu picks this table when vimg_scene::lean is set.
#define VIMG_FAST 1
#define VIMG_FAST_LIBM 1
#define VIMG_LEAN 1
#define QV_SHADE_MIN_BLOCKS_L 5
#define QV_SHADE_MIN_BLOCKS_P 4
#define QV_END_MIN_BLOCKS 8
#define vimg vimg_lean
#include "vimg_kernels.cuh"
#include "vimg_flat.cuh"
#include "vimg_stream.cuh"
#undef vimg

// fills the streamed entries (sv_*) of a table that already holds the tolerance build's other kernels
extern "C" void vimg_lean_kernel_table(void* out) {
  vimg_lean::QwKernelTable& t = *static_cast<vimg_lean::QwKernelTable*>(out);
  vimg_lean::qv_fill_kernel_table(t);
}
