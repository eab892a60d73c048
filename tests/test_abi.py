"""C-ABI surface: the library loads, exports every symbol include/vimg_cuda.h declares, fails
loudly without a GPU, and .vscene files round-trip.  No compute calls (runs without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, golden_scene_path


def declared_functions():
    text = open(os.path.join(ROOT, "include", "vimg_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vimg_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("vimg_scene_create", "vimg_render", "vimg_render_device", "vimg_trace_pixel",
                 "vimg_trace_ids", "vimg_scene_destroy", "vimg_last_error",
                 "vimg_cuda_device_count"):
        assert must in names


def test_library_exports_every_declared_symbol(vimg):
    lib = vimg.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), f"libvimg_cuda.so does not export {name}"
    assert set(vimg.renderer.EXPORTS) == set(declared_functions())


def test_no_device_fails_loudly(vimg):
    if vimg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(vimg.VimgError) as e:
        vimg.Scene.load(golden_scene_path("cornell_small"))
    assert "no CPU fallback" in str(e.value)


def test_create_rejects_bad_arguments(vimg):
    lib = vimg.load_library()
    out = C.c_void_p()
    assert lib.vimg_scene_create(None, 0, C.byref(out)) == -1  # VIMG_ERR_INVALID_ARGUMENT
    assert b"null" in lib.vimg_last_error()
    d = vimg.LoadedDesc(golden_scene_path("cornell_small"))
    head = C.c_uint32.from_address(d.ptr)
    keep = head.value
    head.value = 999  # wrong abi_version
    assert lib.vimg_scene_create(d.ptr, 0, C.byref(out)) == -1
    assert b"abi_version" in lib.vimg_last_error()
    head.value = keep


@pytest.mark.parametrize("name", ["cornell_small", "synthetic_small", "synthetic_rg", "cornell_skylight"])
def test_vscene_round_trip(vimg, tmp_path, name):
    src = golden_scene_path(name)
    d = vimg.LoadedDesc(src)
    dst = str(tmp_path / "copy.vscene")
    vimg.renderer.save_desc(d.ptr, dst)
    assert open(src, "rb").read() == open(dst, "rb").read()


def test_vscene_load_rejects_garbage(vimg, tmp_path):
    bad = tmp_path / "bad.vscene"
    bad.write_bytes(b"not a scene")
    with pytest.raises(vimg.VimgError):
        vimg.LoadedDesc(str(bad))
    trunc = tmp_path / "trunc.vscene"
    trunc.write_bytes(open(golden_scene_path("cornell_small"), "rb").read()[:200])
    with pytest.raises(vimg.VimgError):
        vimg.LoadedDesc(str(trunc))


def test_golden_files_present(man):
    for name, e in man.items():
        assert os.path.exists(os.path.join(GOLDEN, e["scene"])), name
        assert os.path.exists(os.path.join(GOLDEN, "ref", name + ".npz")), name


def test_vscene_load_rejects_arrays_that_contradict_their_dimensions(vimg, tmp_path):
    """A hand-edited file whose array lengths disagree with the dimensions next to them must not reach
    vimg_scene_create, which sizes its reads from the dimensions: here a mesh's normal array loses one
    vertex, everything else stays parseable."""
    import struct
    raw = open(golden_scene_path("synthetic_small"), "rb").read()
    off = 8 + 4 + 92  # magic, abi version, vimg_camera

    def skip_arr(o, elem):
        n, = struct.unpack_from("<Q", raw, o)
        return o + 8 + n * elem, n

    off, _ = skip_arr(off, 4)    # nodes (u32)
    off, _ = skip_arr(off, 4)    # bb_mins_maxes (f32)
    off, _ = skip_arr(off, 8)    # obj_indices (u64)
    off += 4                     # max_depth
    off, _ = skip_arr(off, 12)   # prims
    off, _ = skip_arr(off, 20)   # spheres
    n_meshes, = struct.unpack_from("<I", raw, off)
    assert n_meshes >= 1
    off += 4
    off, _ = skip_arr(off, 4)    # mesh 0: indices
    off, n_vert_floats = skip_arr(off, 4)  # vertices
    n_norm, = struct.unpack_from("<Q", raw, off)
    assert n_norm == n_vert_floats and n_norm >= 6
    bad = raw[:off] + struct.pack("<Q", n_norm - 3) + raw[off + 8:off + 8 + 4 * (n_norm - 3)] + raw[off + 8 + 4 * n_norm:]
    p = tmp_path / "edited.vscene"
    p.write_bytes(bad)
    with pytest.raises(vimg.VimgError):
        vimg.LoadedDesc(str(p))
    ok = tmp_path / "same.vscene"
    ok.write_bytes(raw)
    vimg.LoadedDesc(str(ok)).close()
