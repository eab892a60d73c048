"""ctypes host binding of libvimg_cuda.so (see include/vimg_cuda.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# integrator names of the reference's scene files (src/scene_loading/json_scene.cpp:213-223)
INTEGRATORS = {"s_normal": 0, "g_normal": 1, "material": 2, "mis": 3}
UINT32_MAX = 0xFFFFFFFF  # "depth": -1 in a scene file (json_scene.cpp:186,193)
FLAG_COUNTERS = 1
FLAG_PIPELINE_BLOCK = 2   # persistent block-local wavefront kernel
FLAG_PIPELINE_QUEUE = 4   # multi-kernel queue wavefront (default)
FLAG_STAGE_TIMES = 8      # queue wavefront: serial launches, per-stage device times in the stats
FLAG_FRESH_STACK = 16     # do not reproduce the reference's carried-over traversal stack (faster)
FLAG_FAST = 32            # tolerance mode: fast-math kernels + device-side BVH re-layout (relMSE contract)
STAGES = ("trace", "classify", "shade_principled", "shade_lambertian", "shade_dielectric", "end")
PARTITIONS = {"tile": 0, "spp": 1}  # vimg_partition
PIPELINES = {None: 0, "block": FLAG_PIPELINE_BLOCK, "queue": FLAG_PIPELINE_QUEUE}


class VimgError(RuntimeError):
    pass


class _Camera(C.Structure):
    _fields_ = [("cam_to_world", C.c_float * 16), ("p_size", C.c_float * 2),
                ("resolution", C.c_int32 * 2), ("vfov_deg", C.c_float),
                ("aperture_radius", C.c_float), ("focal_dist", C.c_float)]


class _Bvh(C.Structure):  # vimg_bvh
    _fields_ = [("nodes", C.c_void_p), ("n_nodes", C.c_uint64), ("bb_mins_maxes", C.c_void_p),
                ("n_bb_slots", C.c_uint64), ("obj_indices", C.c_void_p), ("n_obj_indices", C.c_uint64),
                ("max_depth", C.c_uint32)]


class _Texture(C.Structure):  # vimg_texture
    _fields_ = [("type", C.c_uint32), ("col_a", C.c_float * 3), ("col_b", C.c_float * 3),
                ("width", C.c_uint32), ("height", C.c_uint32), ("n_levels", C.c_uint32),
                ("wrap_u", C.c_uint32), ("wrap_v", C.c_uint32), ("levels", C.POINTER(C.POINTER(C.c_float)))]


class _Background(C.Structure):  # vimg_background
    _fields_ = [("type", C.c_uint32), ("color", C.c_float * 3), ("image", C.c_int32),
                ("world_to_env", C.c_float * 16), ("env_to_world", C.c_float * 16),
                ("radiance_scale", C.c_float), ("cdf_width", C.c_uint32), ("cdf_height", C.c_uint32),
                ("row_cdf", C.POINTER(C.c_float)), ("col_cdfs", C.POINTER(C.c_float))]


class _Desc(C.Structure):  # vimg_scene_desc (include/vimg_cuda.h), pointers kept opaque where unused
    _fields_ = [("abi_version", C.c_uint32), ("camera", _Camera), ("bvh", _Bvh),
                ("prims", C.c_void_p), ("n_prims", C.c_uint64),
                ("spheres", C.c_void_p), ("n_spheres", C.c_uint32),
                ("meshes", C.c_void_p), ("n_meshes", C.c_uint32),
                ("materials", C.c_void_p), ("n_materials", C.c_uint32),
                ("textures", C.POINTER(_Texture)), ("n_textures", C.c_uint32),
                ("textures_rg", C.c_void_p), ("n_textures_rg", C.c_uint32),
                ("lights", C.c_void_p), ("n_lights", C.c_uint32),
                ("background", _Background)]


class _DescHead(C.Structure):  # leading members of vimg_scene_desc
    _fields_ = [("abi_version", C.c_uint32), ("camera", _Camera)]


# enum class tonemapper (include/color_utils.h:7) + none
TONEMAPPERS = {"clamp": 0, "agx": 1, "reinhard": 2, "aces": 3, "none": 4}


class RenderParams(C.Structure):
    _fields_ = [("integrator", C.c_uint32), ("samples", C.c_uint32), ("depth", C.c_uint32),
                ("rank", C.c_uint32), ("n_ranks", C.c_uint32), ("flags", C.c_uint32),
                ("spp_streams", C.c_uint32), ("partition", C.c_uint32)]


class RenderStats(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("inner_nodes", C.c_uint64), ("leaf_nodes", C.c_uint64),
                ("tri_tests", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("nan_samples", C.c_uint64), ("samples", C.c_uint64),
                ("kernel_ms", C.c_double), ("total_ms", C.c_double),
                ("kernel_launches", C.c_uint32), ("stage_launches", C.c_uint32 * 6),
                ("stage_ms", C.c_double * 6), ("stage_items", C.c_uint64 * 6)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["stage_ms"] = dict(zip(STAGES, self.stage_ms))
        d["stage_launches"] = dict(zip(STAGES, self.stage_launches))
        d["stage_items"] = dict(zip(STAGES, self.stage_items))
        return d


EXPORTS = ("vimg_cuda_device_count", "vimg_last_error", "vimg_scene_create", "vimg_scene_destroy",
           "vimg_scene_device_bytes", "vimg_render", "vimg_render_device", "vimg_trace_pixel",
           "vimg_trace_ids", "vimg_release_workspaces", "vimg_render_multi", "vimg_postprocess", "vimg_render_srgb8", "vimg_heatmap",
           "vimg_mip_chain_levels", "vimg_build_mip_chain", "vimg_build_envmap_cdf",
           "vimg_convert_srgb8_to_linear", "vimg_convert_rgb8_to_normal", "vimg_build_image_texture", "vimg_libm_eval",
           "vimg_scene_desc_save", "vimg_scene_desc_load", "vimg_scene_desc_free")


def lib_path():
    # VIMG_CUDA_LIB selects an alternative build of the same library (kernel tuning experiments)
    return os.environ.get("VIMG_CUDA_LIB") or os.path.join(_HERE, "libvimg_cuda.so")


_lib = None


def load_library():
    """Load libvimg_cuda.so; raises VimgError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise VimgError(f"{path} is missing: build it with `make -C v-img_b200/csrc` "
                        "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    lib = C.CDLL(path)
    lib.vimg_cuda_device_count.restype = C.c_int
    lib.vimg_last_error.restype = C.c_char_p
    lib.vimg_scene_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.vimg_scene_destroy.argtypes = [C.c_void_p]
    lib.vimg_scene_device_bytes.restype = C.c_uint64
    lib.vimg_scene_device_bytes.argtypes = [C.c_void_p]
    lib.vimg_render.argtypes = [C.c_void_p, C.POINTER(RenderParams), C.c_void_p,
                                C.POINTER(RenderStats)]
    lib.vimg_render_device.argtypes = [C.c_void_p, C.POINTER(RenderParams), C.POINTER(C.c_void_p),
                                       C.c_void_p, C.POINTER(RenderStats)]
    lib.vimg_render_multi.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_uint32, C.POINTER(RenderParams), C.c_uint32,
                                      C.c_void_p, C.POINTER(RenderStats)]
    lib.vimg_postprocess.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_double)]
    lib.vimg_render_srgb8.argtypes = [C.c_void_p, C.POINTER(RenderParams), C.c_uint32, C.c_void_p,
                                      C.POINTER(RenderStats)]
    lib.vimg_mip_chain_levels.restype = C.c_uint32
    lib.vimg_mip_chain_levels.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
    lib.vimg_build_mip_chain.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                         C.c_void_p, C.POINTER(C.c_double)]
    lib.vimg_build_envmap_cdf.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_double)]
    lib.vimg_convert_srgb8_to_linear.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
    lib.vimg_convert_rgb8_to_normal.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_int, C.c_void_p,
                                                C.POINTER(C.c_double)]
    lib.vimg_build_image_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                             C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    lib.vimg_libm_eval.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    lib.vimg_heatmap.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p]
    lib.vimg_trace_pixel.argtypes = [C.c_void_p, C.POINTER(RenderParams), C.c_int, C.c_int,
                                     C.c_void_p]
    lib.vimg_trace_ids.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.vimg_scene_desc_save.argtypes = [C.c_void_p, C.c_char_p]
    lib.vimg_scene_desc_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.vimg_scene_desc_free.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def release_workspaces():
    """Free the idle per-device render workspaces and scene arenas (vimg_release_workspaces)."""
    load_library().vimg_release_workspaces()


def device_count():
    return load_library().vimg_cuda_device_count()


def _check(lib, rc, what):
    if rc != 0:
        raise VimgError(f"{what} failed ({rc}): {lib.vimg_last_error().decode()}")


def save_desc(desc_ptr, path):
    lib = load_library()
    rc = lib.vimg_scene_desc_save(desc_ptr, os.fsencode(path))
    if rc != 0:
        raise VimgError(f"vimg_scene_desc_save({path}) failed ({rc})")


def tile_owner(width, height, n_ranks):
    """Rank that renders each pixel: int32 [H, W], row 0 = top of the image.

    Host-side mirror of the device mapping (vimg_wavefront.cuh::wf_pixel_xy): the reference's
    8x8 tiles, numbered x-major like its work_list (include/integrators.h:57-65), are dealt
    round-robin: tile t -> rank t % n_ranks.
    """
    tiles_y = (height + 7) // 8
    y, x = np.mgrid[0:height, 0:width]          # y counts from the bottom row, like the reference
    tile = (x // 8) * tiles_y + (y // 8)
    owner = (tile % max(n_ranks, 1)).astype(np.int32)
    return owner[::-1].copy()                   # image row 0 is y = H - 1


def spp_ranges(samples, spp_streams, rank=0, n_ranks=1):
    """Sample ranges [begin, end) a rank renders of EVERY pixel under VIMG_PARTITION_SPP, in the order its
    combine pass adds them.  Host-side mirror of the device mapping (vimg_wavefront.cuh WfArgs::s_rank /
    s_n_ranks / my_streams; include/vimg_cuda.h vimg_render_params.spp_streams): K = min(spp_streams, samples)
    ranges [k * samples / K, (k + 1) * samples / K), range k belongs to rank k % n_ranks."""
    k_total = max(1, min(int(spp_streams), int(samples)))
    if n_ranks > 1 and k_total < n_ranks:
        raise VimgError("VIMG_PARTITION_SPP needs spp_streams (and samples) >= n_ranks")
    return [(k * samples // k_total, (k + 1) * samples // k_total) for k in range(rank, k_total, max(n_ranks, 1))]


class LoadedDesc:
    """A vimg_scene_desc read from a .vscene file; `.ptr` is its address (host memory)."""

    def __init__(self, path):
        self._lib = load_library()
        d = C.c_void_p()
        rc = self._lib.vimg_scene_desc_load(os.fsencode(path), C.byref(d))
        if rc != 0:
            raise VimgError(f"vimg_scene_desc_load({path}) failed ({rc})")
        self.ptr = d.value

    def image_textures(self):
        return desc_image_textures(self.ptr)

    def envmap(self):
        return desc_envmap(self.ptr)

    def close(self):
        if self.ptr:
            self._lib.vimg_scene_desc_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DescViews:
    """readers of a vimg_scene_desc in host memory (tests of the scene-preparation kernels)"""

    @staticmethod
    def image_textures(ptr):
        """[(index, wrap_u, wrap_v, [level arrays float32 [h, w, 3]])] of the image textures"""
        d = _Desc.from_address(ptr)
        out = []
        for i in range(d.n_textures):
            t = d.textures[i]
            if t.type != 2:  # VIMG_TEX_IMAGE
                continue
            levels = []
            for lv in range(t.n_levels):
                w, h = max(t.width >> lv, 1), max(t.height >> lv, 1)
                levels.append(np.ctypeslib.as_array(t.levels[lv], shape=(h, w, 3)).copy())
            out.append((i, int(t.wrap_u), int(t.wrap_v), levels))
        return out

    @staticmethod
    def envmap(ptr):
        """(image texture index, row_cdf [H+1], col_cdfs [H, W+1]) of an env-map background, or None"""
        b = _Desc.from_address(ptr).background
        if b.type != 1:  # VIMG_BG_ENVMAP
            return None
        h, w = int(b.cdf_height), int(b.cdf_width)
        return (int(b.image), np.ctypeslib.as_array(b.row_cdf, shape=(h + 1,)).copy(),
                np.ctypeslib.as_array(b.col_cdfs, shape=(h, w + 1)).copy())


desc_image_textures = _DescViews.image_textures
desc_envmap = _DescViews.envmap


class Scene:
    """A scene resident in the HBM of one GPU.

    Build it from a flattened-scene file (`Scene.load`) or from the address of a
    `vimg_scene_desc` produced by the reference-side adapter (`Scene.from_desc`).
    """

    def __init__(self, desc_ptr, device=0, _owned_desc=None):
        self._lib = load_library()
        self._owned_desc = _owned_desc
        self._h = None
        head = _DescHead.from_address(desc_ptr)
        self.width, self.height = int(head.camera.resolution[0]), int(head.camera.resolution[1])
        self.device = device
        h = C.c_void_p()
        _check(self._lib, self._lib.vimg_scene_create(desc_ptr, device, C.byref(h)),
               "vimg_scene_create")
        self._h = h

    @classmethod
    def load(cls, path, device=0):
        lib = load_library()
        d = C.c_void_p()
        rc = lib.vimg_scene_desc_load(os.fsencode(path), C.byref(d))
        if rc != 0:
            raise VimgError(f"vimg_scene_desc_load({path}) failed ({rc})")
        try:
            return cls(d.value, device)
        finally:
            lib.vimg_scene_desc_free(d)  # vimg_scene_create deep-copied everything

    @classmethod
    def from_desc(cls, desc_ptr, device=0):
        return cls(desc_ptr, device)

    def close(self):
        if self._h:
            self._lib.vimg_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self):
        return int(self._lib.vimg_scene_device_bytes(self._h))

    def _params(self, func, samples, depth, rank, n_ranks, counters, pipeline=None, stage_times=False,
                spp_streams=1, fresh_stack=False, fast=False, partition="tile"):
        flags = (FLAG_COUNTERS if counters else 0) | PIPELINES[pipeline] | (FLAG_STAGE_TIMES if stage_times else 0)
        flags |= FLAG_FRESH_STACK if fresh_stack else 0
        flags |= FLAG_FAST if fast else 0
        return RenderParams(INTEGRATORS[func], int(samples), int(depth) & UINT32_MAX, rank, n_ranks, flags,
                            int(spp_streams), PARTITIONS[partition])

    def scene_integrator(self, func, samples, depth, rank=0, n_ranks=1, counters=False, out=None,
                         pipeline=None, stage_times=False, spp_streams=1, fresh_stack=False, fast=False, partition="tile"):
        """Mirror of scene_integrator<F>() (include/integrators.h:36-153).

        Returns (image float32 [H, W, 3] in HOST memory, row 0 = top; stats dict).
        `out` may be a (pinned) float32 array to receive the frame.
        `spp_streams` = 1 follows the reference's one RNG stream per pixel (bit-identical images);
        K > 1 splits every pixel's samples into K ranges with their own streams
        (vimg_render_params.spp_streams, include/vimg_cuda.h).
        `fresh_stack` = True drops the reference's carried-over traversal stack (VIMG_FLAG_FRESH_STACK):
        faster, images within the relMSE tolerance instead of bit-identical.
        `fast` = True selects the tolerance mode (VIMG_FLAG_FAST): fast-math build of the kernels and a
        device-side re-layout of the BVH; material / mis images within the stated relMSE tolerance.
        """
        if out is None:
            out = np.empty((self.height, self.width, 3), np.float32)
        assert out.dtype == np.float32 and out.size == self.height * self.width * 3
        st = RenderStats()
        p = self._params(func, samples, depth, rank, n_ranks, counters, pipeline, stage_times, spp_streams, fresh_stack,
                         fast, partition)
        _check(self._lib, self._lib.vimg_render(self._h, C.byref(p), out.ctypes.data, C.byref(st)),
               "vimg_render")
        return out, st.as_dict()

    def render_device(self, func, samples, depth, rank=0, n_ranks=1, counters=False, stream=None,
                      pipeline=None, stage_times=False, spp_streams=1, fresh_stack=False, fast=False, partition="tile"):
        """Render and leave the frame in HBM.  Returns (device pointer, stats dict)."""
        st = RenderStats()
        p = self._params(func, samples, depth, rank, n_ranks, counters, pipeline, stage_times, spp_streams, fresh_stack,
                         fast, partition)
        dev = C.c_void_p()
        _check(self._lib, self._lib.vimg_render_device(self._h, C.byref(p), C.byref(dev), stream,
                                                       C.byref(st)), "vimg_render_device")
        return dev.value, st.as_dict()

    def frame_tensor(self, dev_ptr):
        """Wrap the device frame as a torch tensor [H, W, 3] without copying (for NCCL)."""
        import torch

        class _Holder:
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (self.height, self.width, 3), "typestr": "<f4",
                                      "data": (dev_ptr, False), "version": 3}
        return torch.as_tensor(h, device=f"cuda:{self.device}")

    def render_srgb8(self, func, samples, depth, tonemapper="clamp", spp_streams=1, pipeline=None):
        """scene_integrator + tonemapper + sRGB OETF + 8-bit quantisation (src/main.cpp:219-356)
        without leaving the device.  Returns (uint8 [H, W, 3] = the PNG payload, stats dict)."""
        out = np.empty((self.height, self.width, 3), np.uint8)
        st = RenderStats()
        p = self._params(func, samples, depth, 0, 1, False, pipeline, False, spp_streams)
        _check(self._lib, self._lib.vimg_render_srgb8(self._h, C.byref(p), TONEMAPPERS[tonemapper],
                                                      out.ctypes.data, C.byref(st)), "vimg_render_srgb8")
        return out, st.as_dict()

    def heatmap(self, samples=4, factor=20.0):
        """Mirror of heatmap_img() (src/integrators/heatmap.cpp:39-147; the -m flag).
        Returns (rgb float32 [H, W, 3], truncated mean BVH cost float32 [H, W])."""
        rgb = np.empty((self.height, self.width, 3), np.float32)
        cost = np.empty((self.height, self.width), np.float32)
        _check(self._lib, self._lib.vimg_heatmap(self._h, samples, factor, rgb.ctypes.data, cost.ctypes.data),
               "vimg_heatmap")
        return rgb, cost

    def trace_pixel(self, func, x, y, samples, depth, pipeline=None):
        """Mirror of trace_pixel<F>() (include/integrators.h:181-220); y = 0 is the bottom row."""
        out = np.empty(3, np.float32)
        p = self._params(func, samples, depth, 0, 1, False, pipeline)
        _check(self._lib, self._lib.vimg_trace_pixel(self._h, C.byref(p), x, y, out.ctypes.data),
               "vimg_trace_pixel")
        return out

    def trace_ids(self, samples, want_t=True):
        """Primary-hit primitive ids, int32 [H*W, samples] (-1 = miss), and hit distances."""
        ids = np.empty((self.height * self.width, samples), np.int32)
        t = np.empty((self.height * self.width, samples), np.float32) if want_t else None
        _check(self._lib, self._lib.vimg_trace_ids(self._h, samples, ids.ctypes.data,
                                                   t.ctypes.data if want_t else None),
               "vimg_trace_ids")
        return ids, t


def render_multi(desc_ptr, devices, func, samples, depth, partition="tile", spp_streams=1, fast=False, fresh_stack=False,
                 out=None):
    """vimg_render_multi: the whole frame on several GPUs of this box behind one call.
    desc_ptr: address of a vimg_scene_desc (LoadedDesc.ptr, or the adapter's).  -> (image [H, W, 3], stats dict)"""
    lib = load_library()
    head = _DescHead.from_address(desc_ptr)
    w, h = head.camera.resolution[0], head.camera.resolution[1]
    if out is None:
        out = np.empty((h, w, 3), np.float32)
    flags = (FLAG_FAST if fast else 0) | (FLAG_FRESH_STACK if fresh_stack else 0)
    p = RenderParams(INTEGRATORS[func], int(samples), int(depth) & UINT32_MAX, 0, 1, flags, int(spp_streams), 0)
    ids = (C.c_int * len(devices))(*devices)
    st = RenderStats()
    _check(lib, lib.vimg_render_multi(desc_ptr, ids, len(devices), C.byref(p), PARTITIONS[partition], out.ctypes.data,
                                      C.byref(st)), "vimg_render_multi")
    return out, st.as_dict()


def postprocess(img, tonemapper="clamp", device=0, want_srgb=True):
    """Mirror of src/main.cpp:304-356 on a linear float image [..., 3]: tonemapper, sRGB OETF,
    8-bit quantisation, on the GPU.  Returns (uint8 image, float32 image after gamma or None,
    kernel milliseconds)."""
    lib = load_library()
    img = np.ascontiguousarray(img, np.float32)
    assert img.shape[-1] == 3
    rgb8 = np.empty(img.shape, np.uint8)
    srgb = np.empty_like(img) if want_srgb else None
    ms = C.c_double()
    _check(lib, lib.vimg_postprocess(img.ctypes.data, img.size // 3, TONEMAPPERS[tonemapper], device,
                                     rgb8.ctypes.data, srgb.ctypes.data if want_srgb else None,
                                     C.byref(ms)), "vimg_postprocess")
    return rgb8, srgb, ms.value


def build_mip_chain(level0, wrap_u=0, wrap_v=0, device=0):
    """Mirror of the mip-chain loop of ImageTexture's constructor (src/image_texture.cpp:59-130).
    level0: float32 [H, W, 3].  Returns ([level 1, level 2, ...] float32 arrays, kernel ms)."""
    lib = load_library()
    level0 = np.ascontiguousarray(level0, np.float32)
    h, w = level0.shape[:2]
    floats = C.c_uint64()
    n = lib.vimg_mip_chain_levels(w, h, C.byref(floats))
    buf = np.empty(floats.value, np.float32)
    ms = C.c_double()
    _check(lib, lib.vimg_build_mip_chain(level0.ctypes.data, w, h, wrap_u, wrap_v, device, buf.ctypes.data,
                                         C.byref(ms)), "vimg_build_mip_chain")
    levels, off = [], 0
    for lv in range(1, n):
        w, h = max(w // 2, 1), max(h // 2, 1)
        levels.append(buf[off:off + w * h * 3].reshape(h, w, 3))
        off += w * h * 3
    return levels, ms.value


def build_envmap_cdf(image, device=0):
    """Mirror of ArraySampling2D's constructor (include/rng/sampling.h:167-197).
    image: float32 [H, W, 3].  Returns (row_cdf [H+1], col_cdfs [H, W+1], kernel ms)."""
    lib = load_library()
    image = np.ascontiguousarray(image, np.float32)
    h, w = image.shape[:2]
    row = np.empty(h + 1, np.float32)
    col = np.empty((h, w + 1), np.float32)
    ms = C.c_double()
    _check(lib, lib.vimg_build_envmap_cdf(image.ctypes.data, w, h, device, row.ctypes.data, col.ctypes.data,
                                          C.byref(ms)), "vimg_build_envmap_cdf")
    return row, col, ms.value


def convert_srgb8_to_linear(rgb8, device=0):
    """Mirror of make_texture's byte -> float loop + ImageTexture::convert_sRGB_to_linear
    (src/scene_loading/gltf_loading.cpp:139-150, src/image_texture.cpp:257-263).
    rgb8: uint8 [..., 3].  Returns (float32 array of the same shape, kernel ms)."""
    lib = load_library()
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    out = np.empty(rgb8.shape, np.float32)
    ms = C.c_double()
    _check(lib, lib.vimg_convert_srgb8_to_linear(rgb8.ctypes.data, rgb8.size // 3, device, out.ctypes.data,
                                                 C.byref(ms)), "vimg_convert_srgb8_to_linear")
    return out, ms.value


def convert_rgb8_to_normal(rgb8, scale=1.0, device=0):
    """Mirror of make_texture's byte -> float loop + ImageTexture::convert_RGB_to_normal
    (src/scene_loading/gltf_loading.cpp:139-162, src/image_texture.cpp:265-275)."""
    lib = load_library()
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    out = np.empty(rgb8.shape, np.float32)
    ms = C.c_double()
    _check(lib, lib.vimg_convert_rgb8_to_normal(rgb8.ctypes.data, rgb8.size // 3, scale, device,
                                                out.ctypes.data, C.byref(ms)), "vimg_convert_rgb8_to_normal")
    return out, ms.value


def build_image_texture(srgb8, wrap_u=0, wrap_v=0, device=0):
    """make_texture's TextureType::Image case (gltf_loading.cpp:147-156): decoded sRGB bytes [H, W, 3]
    -> ([level 0, level 1, ...] float32 arrays, kernel ms), converted and filtered on the device."""
    lib = load_library()
    srgb8 = np.ascontiguousarray(srgb8, np.uint8)
    h, w = srgb8.shape[:2]
    floats = C.c_uint64()
    n = lib.vimg_mip_chain_levels(w, h, C.byref(floats))
    level0 = np.empty((h, w, 3), np.float32)
    buf = np.empty(floats.value, np.float32)
    ms = C.c_double()
    _check(lib, lib.vimg_build_image_texture(srgb8.ctypes.data, w, h, wrap_u, wrap_v, device, level0.ctypes.data,
                                             buf.ctypes.data, C.byref(ms)), "vimg_build_image_texture")
    levels, off = [level0], 0
    for lv in range(1, n):
        w, h = max(w // 2, 1), max(h // 2, 1)
        levels.append(buf[off:off + w * h * 3].reshape(h, w, 3))
        off += w * h * 3
    return levels, ms.value


LIBM_FUNCTIONS = {"sinf": 0, "cosf": 1, "logf": 2, "log2f": 3, "powf": 4, "acosf": 5, "atan2f": 6, "atanf": 7,
                  "log2_lo": 8, "log2_hi": 9}  # double log2 of the widened input, result bits in two halves


def libm_eval(fn, x, y=None, device=0):
    """Parity hook: the device's restatement of glibc's float function `fn` on the float32 array x
    (and y for powf / atan2f; atan2f(x, y) takes x as the first argument like the C function)."""
    lib = load_library()
    x = np.ascontiguousarray(x, np.float32)
    y = None if y is None else np.ascontiguousarray(y, np.float32)
    out = np.empty(x.shape, np.float32)
    _check(lib, lib.vimg_libm_eval(LIBM_FUNCTIONS[fn], x.ctypes.data, None if y is None else y.ctypes.data, x.size,
                                   device, out.ctypes.data), "vimg_libm_eval")
    return out


def strip_prepared(desc_ptr):
    """Null the host-built mip levels 1.. and env-map CDFs of a descriptor IN PLACE, so that
    vimg_scene_create builds them on the device (what an adapter that skips ImageTexture's mip loop and
    ArraySampling2D's constructor hands over).  Returns (textures stripped, cdfs stripped)."""
    d = C.cast(desc_ptr, C.POINTER(_Desc)).contents
    n_tex = 0
    for t in range(d.n_textures):
        tex = d.textures[t]
        if tex.type == 2 and tex.n_levels > 1:  # VIMG_TEX_IMAGE
            for lv in range(1, tex.n_levels):
                tex.levels[lv] = C.POINTER(C.c_float)()
            n_tex += 1
    bg = d.background
    cdfs = bg.type == 1  # VIMG_BG_ENVMAP
    if cdfs:
        bg.row_cdf = C.POINTER(C.c_float)()
        bg.col_cdfs = C.POINTER(C.c_float)()
    return n_tex, cdfs
