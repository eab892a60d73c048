"""v-img_b200 -- B200-native `cuda` render device for the v-img path tracer.

The product is libvimg_cuda.so (hand-written CUDA for sm_100a behind the C ABI of
include/vimg_cuda.h).  This package is the thin Python host side: it loads the
library with ctypes and mirrors the reference's render entry points
(scene_integrator / trace_pixel, include/integrators.h:36-220) by name.

There is NO CPU fallback: importing works anywhere, but every compute call
raises VimgError when the CUDA extension or a CUDA device is missing.
"""
from .renderer import (INTEGRATORS, TONEMAPPERS, LoadedDesc, Scene, VimgError, device_count,  # noqa: F401
                       lib_path, load_library, postprocess, tile_owner, build_mip_chain, build_envmap_cdf,
                       desc_envmap, desc_image_textures, convert_srgb8_to_linear, convert_rgb8_to_normal,
                       build_image_texture, libm_eval, LIBM_FUNCTIONS, strip_prepared, release_workspaces, render_multi, spp_ranges)
