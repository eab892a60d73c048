// vimg_fast.cu -- the TOLERANCE build of the queue-wavefront kernels (VIMG_FLAG_FAST).
//
// The same device sources as the bit-exact build in vimg_api.cu, compiled a second time with
// -use_fast_math (approximate division / sqrt / transcendental units, FMA contraction, flush to
// zero) and with VIMG_FAST defined: `vdbl` (the sub-expressions the reference evaluates in double)
// becomes float, the libm restatements become the hardware's intrinsics, and the kernels the
// bit-exact contract rules out (the flat-list trace of vimg_flat.cuh, the 4-wide BVH of vimg_wide.cuh) exist.  Everything lives in namespace
// vimg_fast so that the two builds can sit in one library; vimg_api.cu launches either set through
// the QwKernelTable below.  The argument structs (DScene, QwState, WfArgs) contain no `vdbl`, so
// their layout is the same in both builds.
#define VIMG_FAST 1
#define VIMG_FAST_LIBM 1
#define vimg vimg_fast
#include "vimg_kernels.cuh"
#include "vimg_flat.cuh"
#include "vimg_stream.cuh"
#include "vimg_wide.cuh"
#undef vimg

extern "C" void vimg_fast_kernel_table(void* out) {
  vimg_fast::QwKernelTable t = vimg_fast::qw_kernel_table();
  vimg_fast::qf_fill_kernel_table(t);
  vimg_fast::qv_fill_kernel_table(t);
  vimg_fast::qw_wide_fill_kernel_table(t);
  *static_cast<vimg_fast::QwKernelTable*>(out) = t;
}

extern "C" size_t vimg_fast_flat_scene_bytes(void) { return sizeof(vimg_fast::QfScene); }
