"""ctypes binding of the compiled reference renderer (oracle/_ref/libvimg_ref_*.so).

ORACLE / TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.  The
product path (v-img_b200/) never does.

The shared objects are built from the unmodified reference sources by
oracle/Makefile (see oracle/ref_driver.cpp for the C API).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FLAVOURS = ("parity", "perf", "count", "binding")
INTEGRATORS = {"s_normal": 0, "g_normal": 1, "material": 2, "mis": 3}


def lib_path(flavour):
    return os.path.join(_HERE, "_ref", f"libvimg_ref_{flavour}.so")


def available(flavour="parity"):
    return os.path.exists(lib_path(flavour))


_libs = {}


def _load(flavour):
    if flavour in _libs:
        return _libs[flavour]
    lib = C.CDLL(lib_path(flavour))
    lib.vref_scene_from_json.restype = C.c_void_p
    lib.vref_scene_from_json.argtypes = [C.c_char_p, C.c_int]
    lib.vref_scene_synthetic.restype = C.c_void_p
    lib.vref_scene_synthetic.argtypes = [C.c_int] * 6 + [C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                                         C.c_float]
    lib.vref_scene_from_desc.restype = C.c_void_p
    lib.vref_scene_from_desc.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
    lib.vref_scene_free.argtypes = [C.c_void_p]
    lib.vref_scene_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    lib.vref_scene_set_const_background.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float,
                                                    C.c_int]
    lib.vref_scene_set_camera_dof.argtypes = [C.c_void_p, C.c_float, C.c_float]
    lib.vref_scene_desc.restype = C.c_void_p
    lib.vref_scene_desc.argtypes = [C.c_void_p]
    lib.vref_render.restype = C.c_double
    lib.vref_render.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    lib.vref_trace_pixel.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                     C.c_void_p]
    lib.vref_trace_ids.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.vref_counters.argtypes = [C.POINTER(C.c_uint64), C.c_int]
    lib.vref_heatmap.restype = None
    lib.vref_heatmap.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_void_p]
    lib.vref_postprocess.restype = None
    lib.vref_postprocess.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    if flavour == "binding":
        lib.vref_render_cuda.restype = C.c_int
        lib.vref_render_cuda.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_int), C.c_int,
                                         C.c_void_p, C.c_char_p, C.c_int]
    lib.vref_flavour.restype = C.c_int
    lib.vref_hardware_threads.restype = C.c_int
    _libs[flavour] = lib
    return lib


class RefScene:
    """A scene loaded and BVH-built by the reference's own host code."""

    def __init__(self, handle, flavour):
        if not handle:
            raise RuntimeError("reference failed to load the scene")
        self._h = handle
        self.flavour = flavour
        self._lib = _load(flavour)
        info = (C.c_uint64 * 8)()
        bs = C.c_double()
        self._lib.vref_scene_info(self._h, info, C.byref(bs))
        (self.width, self.height, self.spp, self.depth, self.integrator, self.n_prims,
         self.n_nodes, self.max_depth) = [int(v) for v in info]
        self.bvh_build_s = bs.value

    @classmethod
    def from_json(cls, path, bvh="binned", flavour="parity"):
        lib = _load(flavour)
        return cls(lib.vref_scene_from_json(os.fsencode(path), 1 if bvh == "sweep" else 0), flavour)

    @classmethod
    def synthetic(cls, G, width, height, tex_size=2048, env_w=2048, env_h=1024, spp=512, depth=64,
                  bvh="sweep", with_rg=False, aperture=0.02, flavour="parity"):
        lib = _load(flavour)
        h = lib.vref_scene_synthetic(G, width, height, tex_size, env_w, env_h, spp, depth,
                                     1 if bvh == "sweep" else 0, int(with_rg), aperture)
        return cls(h, flavour)

    @classmethod
    def from_desc(cls, desc_ptr, spp, depth, integrator="mis", flavour="parity"):
        """Rebuild the reference's scene objects from a flattened vimg_scene_desc (e.g. a loaded
        tests/golden/scenes/*.vscene); the descriptor's BVH is used as is."""
        lib = _load(flavour)
        return cls(lib.vref_scene_from_desc(desc_ptr, spp, depth & 0xFFFFFFFF,
                                            INTEGRATORS[integrator]), flavour)

    def close(self):
        if self._h:
            self._lib.vref_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_const_background(self, rgb, as_light=True):
        self._lib.vref_scene_set_const_background(self._h, *[float(c) for c in rgb], int(as_light))

    def set_camera_dof(self, aperture, focal_dist):
        self._lib.vref_scene_set_camera_dof(self._h, aperture, focal_dist)

    def desc_ptr(self):
        """Address of the vimg_scene_desc flattened by the product's adapter (owned by self)."""
        p = self._lib.vref_scene_desc(self._h)
        if not p:
            raise RuntimeError("adapter could not flatten the scene")
        return p

    def render(self, integrator, spp=0, depth=0, threads=0):
        """-> (image float32 [H, W, 3] row 0 = top, seconds inside scene_integrator)."""
        out = np.empty((self.height, self.width, 3), np.float32)
        sec = self._lib.vref_render(self._h, INTEGRATORS[integrator], spp, depth, threads,
                                    out.ctypes.data)
        return out, sec

    def render_cuda(self, integrator, spp=0, depth=0, devices=(0,)):
        """binding flavour only: the adapter's scene_integrator_cuda (one device) or
        scene_integrator_cuda_multi (several) on this scene's reference objects -> image [H, W, 3]"""
        out = np.empty((self.height, self.width, 3), np.float32)
        err = C.create_string_buffer(512)
        ids = (C.c_int * len(devices))(*devices)
        rc = self._lib.vref_render_cuda(self._h, INTEGRATORS[integrator], spp, depth, ids, len(devices),
                                        out.ctypes.data, err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return out

    def trace_pixel(self, integrator, x, y, spp=0, depth=0):
        out = np.empty(3, np.float32)
        self._lib.vref_trace_pixel(self._h, INTEGRATORS[integrator], spp, depth, x, y,
                                   out.ctypes.data)
        return out

    def heatmap(self, spp=4, factor=20.0):
        """heatmap_img (src/integrators/heatmap.cpp:39-147) -> float32 [H, W, 3]"""
        out = np.empty((self.height, self.width, 3), np.float32)
        self._lib.vref_heatmap(self._h, spp, factor, out.ctypes.data)
        return out

    def trace_ids(self, spp):
        """-> (ids int32 [H*W, spp], t float32 [H*W, spp]); row index = image_index."""
        ids = np.empty((self.height * self.width, spp), np.int32)
        t = np.empty((self.height * self.width, spp), np.float32)
        self._lib.vref_trace_ids(self._h, spp, ids.ctypes.data, t.ctypes.data)
        return ids, t

    def counters(self, reset=True):
        out = (C.c_uint64 * 6)()
        self._lib.vref_counters(out, int(reset))
        return dict(zip(("rays_closest", "rays_shadow", "inner_nodes", "leaf_nodes", "prim_tests",
                         "unused"), [int(v) for v in out]))


def hardware_threads(flavour="perf"):
    return _load(flavour).vref_hardware_threads()


SURVEY_FNV_BASIS = 1469598103934665603  # SURVEY.md's hashes used this (typo'd) offset basis


def fnv1a64(arr, basis=0xCBF29CE484222325):
    """FNV-1a-64 over the raw bytes of an array (sequential; use on small arrays only)."""
    hv = int(basis)
    for b in np.ascontiguousarray(arr).view(np.uint8).ravel().tobytes():
        hv = ((hv ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{hv:016x}"


def digest(arr):
    """64-bit BLAKE2b of the raw bytes of an array: the hash tests/golden/manifest.json records."""
    import hashlib
    return hashlib.blake2b(np.ascontiguousarray(arr).view(np.uint8).ravel().tobytes(),
                           digest_size=8).hexdigest()


# ---- plain-C restatement of the bit-exact core (oracle/vimg_oracle.c) -------------------------
def restatement_path():
    return os.path.join(_HERE, "_ref", "libvimg_oracle.so")


_restatement = None


def restatement_primary(desc_ptr, width, height, spp, normals=True):
    """Run the C restatement on a flattened scene.  -> (ids, t, s_normal, g_normal, unsupported)"""
    global _restatement
    if _restatement is None:
        _restatement = C.CDLL(restatement_path())
        _restatement.vimg_oracle_primary.restype = C.c_int
        _restatement.vimg_oracle_primary.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 4
    ids = np.empty((height * width, spp), np.int32)
    t = np.empty((height * width, spp), np.float32)
    sn = np.empty((height, width, 3), np.float32) if normals else None
    gn = np.empty((height, width, 3), np.float32) if normals else None
    rc = _restatement.vimg_oracle_primary(desc_ptr, spp, ids.ctypes.data, t.ctypes.data,
                                          sn.ctypes.data if normals else None,
                                          gn.ctypes.data if normals else None)
    return ids, t, sn, gn, bool(rc)


# ---- post-process chain (src/main.cpp:304-356) ---------------------------------------------------
TONEMAPPERS = {"clamp": 0, "agx": 1, "reinhard": 2, "aces": 3, "none": 4}


def postprocess(img, tonemapper, flavour="parity"):
    """The reference's own tonemapper + sRGB OETF + main.cpp's 8-bit quantisation.
    -> (float32 image after gamma, uint8 image), both shaped like `img`."""
    img = np.ascontiguousarray(img, np.float32)
    srgb = np.empty_like(img)
    rgb8 = np.empty(img.shape, np.uint8)
    _load(flavour).vref_postprocess(img.ctypes.data, img.size // 3, TONEMAPPERS[tonemapper],
                                    srgb.ctypes.data, rgb8.ctypes.data)
    return srgb, rgb8


def restatement_postprocess(img, tonemapper):
    """Same chain through the plain-C restatement (oracle/vimg_oracle.c)."""
    lib = C.CDLL(restatement_path())
    lib.vimg_oracle_postprocess.restype = None
    lib.vimg_oracle_postprocess.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    img = np.ascontiguousarray(img, np.float32)
    srgb = np.empty_like(img)
    rgb8 = np.empty(img.shape, np.uint8)
    lib.vimg_oracle_postprocess(img.ctypes.data, img.size // 3, TONEMAPPERS[tonemapper],
                                srgb.ctypes.data, rgb8.ctypes.data)
    return srgb, rgb8


def restatement_heatmap(desc_ptr, width, height, spp=4, factor=20.0):
    """BVH heat map through the plain-C restatement -> (rgb float32 [H, W, 3], cost float32 [H, W])"""
    lib = C.CDLL(restatement_path())
    lib.vimg_oracle_heatmap.restype = None
    lib.vimg_oracle_heatmap.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p]
    rgb = np.empty((height, width, 3), np.float32)
    cost = np.empty((height, width), np.float32)
    lib.vimg_oracle_heatmap(desc_ptr, spp, factor, rgb.ctypes.data, cost.ctypes.data)
    return rgb, cost


def restatement_mip_chain(level0, wrap_u, wrap_v):
    """mip levels 1.. of ImageTexture's constructor through the plain-C restatement"""
    lib = C.CDLL(restatement_path())
    lib.vimg_oracle_mip_levels.restype = C.c_uint32
    lib.vimg_oracle_mip_levels.argtypes = [C.c_uint32, C.c_uint32]
    lib.vimg_oracle_mip_chain.restype = None
    lib.vimg_oracle_mip_chain.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    level0 = np.ascontiguousarray(level0, np.float32)
    h, w = level0.shape[:2]
    n = lib.vimg_oracle_mip_levels(w, h)
    shapes = []
    for _ in range(1, n):
        w, h = max(w // 2, 1), max(h // 2, 1)
        shapes.append((h, w, 3))
    buf = np.empty(sum(a * b * c for a, b, c in shapes), np.float32)
    lib.vimg_oracle_mip_chain(level0.ctypes.data, level0.shape[1], level0.shape[0], wrap_u, wrap_v, buf.ctypes.data)
    out, off = [], 0
    for shp in shapes:
        k = shp[0] * shp[1] * 3
        out.append(buf[off:off + k].reshape(shp))
        off += k
    return out


def restatement_envmap_cdf(image):
    """(row_cdf, col_cdfs) of ArraySampling2D through the plain-C restatement"""
    lib = C.CDLL(restatement_path())
    lib.vimg_oracle_envmap_cdf.restype = None
    lib.vimg_oracle_envmap_cdf.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    image = np.ascontiguousarray(image, np.float32)
    h, w = image.shape[:2]
    row = np.empty(h + 1, np.float32)
    col = np.empty((h, w + 1), np.float32)
    scratch = np.empty(w + h, np.float32)
    lib.vimg_oracle_envmap_cdf(image.ctypes.data, w, h, row.ctypes.data, col.ctypes.data, scratch.ctypes.data)
    return row, col


# ---- glTF texture conversions (gltf_loading.cpp:139-162, image_texture.cpp:257-275) --------------
def convert_srgb8_to_linear(rgb8, flavour="parity"):
    """make_texture's byte -> float loop + the reference's own ImageTexture::convert_sRGB_to_linear"""
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    out = np.empty(rgb8.shape, np.float32)
    lib = _load(flavour)
    lib.vref_convert_srgb8_to_linear.restype = None
    lib.vref_convert_srgb8_to_linear.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    lib.vref_convert_srgb8_to_linear(rgb8.ctypes.data, rgb8.size // 3, out.ctypes.data)
    return out


def convert_rgb8_to_normal(rgb8, scale, flavour="parity"):
    """make_texture's byte -> float loop + the reference's own ImageTexture::convert_RGB_to_normal"""
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    out = np.empty(rgb8.shape, np.float32)
    lib = _load(flavour)
    lib.vref_convert_rgb8_to_normal.restype = None
    lib.vref_convert_rgb8_to_normal.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p]
    lib.vref_convert_rgb8_to_normal(rgb8.ctypes.data, rgb8.size // 3, scale, out.ctypes.data)
    return out


def restatement_srgb8_to_linear(rgb8):
    lib = C.CDLL(restatement_path())
    lib.vimg_oracle_srgb8_to_linear.restype = None
    lib.vimg_oracle_srgb8_to_linear.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    out = np.empty(rgb8.shape, np.float32)
    lib.vimg_oracle_srgb8_to_linear(rgb8.ctypes.data, rgb8.size // 3, out.ctypes.data)
    return out


def restatement_rgb8_to_normal(rgb8, scale):
    lib = C.CDLL(restatement_path())
    lib.vimg_oracle_rgb8_to_normal.restype = None
    lib.vimg_oracle_rgb8_to_normal.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p]
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    out = np.empty(rgb8.shape, np.float32)
    lib.vimg_oracle_rgb8_to_normal(rgb8.ctypes.data, rgb8.size // 3, scale, out.ctypes.data)
    return out


def host_libm(fn, x, y=None):
    """glibc's float function number `fn` (numbering of vimg_libm_eval) on float32 arrays"""
    lib = C.CDLL(restatement_path())
    lib.vimg_oracle_libm.restype = None
    lib.vimg_oracle_libm.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    x = np.ascontiguousarray(x, np.float32)
    y = None if y is None else np.ascontiguousarray(y, np.float32)
    out = np.empty(x.shape, np.float32)
    lib.vimg_oracle_libm(fn, x.ctypes.data, None if y is None else y.ctypes.data, x.size, out.ctypes.data)
    return out
