// vimg_exact_lean.cu -- the BIT-EXACT build's shade / END kernels once more with VIMG_LEAN (vimg_math.cuh): the same
// sources and the same arithmetic flags as vimg_api.cu, without the code of image textures, normal maps,
// metallic-roughness textures, the env map, ray cones and the general uv walk, for scenes that have none of them.
// No contraction, no reassociation in this build: removing code that does not run cannot change a bit of the result
// (the full-size digests of BASELINE configs C1 - C3 in tests/test_gpu_parity.py are rendered through these kernels).
#define VIMG_LEAN 1
#define vimg vimg_exact_lean
#include "vimg_kernels.cuh"
#include "vimg_queues.cuh"
#undef vimg

// replaces the shade / END entries of the bit-exact table
extern "C" void vimg_exact_lean_kernel_table(void* out) {
  vimg_exact_lean::QwKernelTable& t = *static_cast<vimg_exact_lean::QwKernelTable*>(out);
  const vimg_exact_lean::QwKernelTable l = vimg_exact_lean::qw_kernel_table();
  for (int i = 0; i < 2; ++i) {
    for (int c = 0; c < 3; ++c) t.shade[i][c] = l.shade[i][c];
    t.end[i] = l.end[i];
  }
}
