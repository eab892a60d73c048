"""Scene-preparation kernels: mip chain of an image texture
(src/image_texture.cpp:59-130) and the env map's sampling CDFs (include/rng/sampling.h:114-197)
built on the device, against what the reference's constructors built.

The golden .vscene files carry the reference-built mip levels and CDFs (the adapter flattens
them), so level 0 goes in and levels 1.. / the CDFs must come out bit for bit.  A larger case
uses the compiled reference itself (oracle/_ref) when it travelled with the repo.
"""
import numpy as np
import pytest

from helpers import golden_scene_path, load_product

SCENES = ["synthetic_small", "synthetic_rg"]


def test_level_count_matches_the_reference_constructor():
    """min(ceil(log2(float(min(w, h)))), 15), at least one level (image_texture.cpp:73-78)"""
    import ctypes as C
    lib = load_product().load_library()
    floats = C.c_uint64()
    for w, h, want in ((128, 128, 7), (128, 64, 6), (2048, 1024, 10), (1, 1, 1), (2, 2, 1), (3, 5, 2),
                       (65536, 65536, 15), (100, 37, 6)):
        assert lib.vimg_mip_chain_levels(w, h, C.byref(floats)) == want, (w, h)
    # the goldens' level counts are the reference's
    vimg = load_product()
    for name in SCENES:
        d = vimg.LoadedDesc(golden_scene_path(name))
        for _, _, _, levels in d.image_textures():
            if len(levels) > 1:
                h, w = levels[0].shape[:2]
                assert lib.vimg_mip_chain_levels(w, h, C.byref(floats)) == len(levels)
                assert floats.value == sum(lv.size for lv in levels[1:])


@pytest.mark.parametrize("name", SCENES)
def test_restatement_matches_the_reference_built_levels_and_cdfs(name):
    """the plain-C restatement (oracle/vimg_oracle.c) against what the reference's constructors built"""
    import os
    from oracle import refcpu
    if not os.path.exists(refcpu.restatement_path()):
        pytest.skip("oracle/_ref/libvimg_oracle.so not built")
    vimg = load_product()
    d = vimg.LoadedDesc(golden_scene_path(name))
    textures = d.image_textures()
    for idx, wrap_u, wrap_v, levels in textures:
        if len(levels) < 2:
            continue
        built = refcpu.restatement_mip_chain(levels[0], wrap_u, wrap_v)
        assert len(built) == len(levels) - 1
        for got, ref in zip(built, levels[1:]):
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"{name} texture {idx}"
    tex, row_ref, col_ref = d.envmap()
    level0 = [lv for i, _, _, lv in textures if i == tex][0][0]
    row, col = refcpu.restatement_envmap_cdf(level0)
    assert np.array_equal(col.view(np.uint32), col_ref.view(np.uint32))
    assert np.array_equal(row.view(np.uint32), row_ref.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_mip_chain_bit_exact(vimg, name):
    d = vimg.LoadedDesc(golden_scene_path(name))
    checked = 0
    for idx, wrap_u, wrap_v, levels in d.image_textures():
        if len(levels) < 2:
            continue
        built, ms = vimg.build_mip_chain(levels[0], wrap_u, wrap_v)
        assert len(built) == len(levels) - 1
        for lv, (got, ref) in enumerate(zip(built, levels[1:]), start=1):
            assert got.shape == ref.shape
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"{name} texture {idx} level {lv}"
        checked += 1
    assert checked >= 2  # the colour texture (repeat) and the env map (clamp)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_envmap_cdf_bit_exact(vimg, name):
    d = vimg.LoadedDesc(golden_scene_path(name))
    tex, row_ref, col_ref = d.envmap()
    level0 = [lv for i, _, _, lv in d.image_textures() if i == tex][0][0]
    row, col, ms = vimg.build_envmap_cdf(level0)
    assert np.array_equal(col.view(np.uint32), col_ref.view(np.uint32))
    assert np.array_equal(row.view(np.uint32), row_ref.view(np.uint32))
    assert row[0] == 0 and row[-1] == 1 and (np.diff(row) >= 0).all()


@pytest.mark.gpu
def test_prep_against_the_live_reference_at_2k(vimg, refcpu):
    """2048^2 colour texture and 2048x1024 env map, built by the compiled reference on this host and
    by the device: levels and CDFs bit for bit; prints the kernel times"""
    if not refcpu.available("parity"):
        pytest.skip("oracle/_ref/libvimg_ref_parity.so not built")
    rs = refcpu.RefScene.synthetic(G=4, width=64, height=36, tex_size=2048, env_w=2048, env_h=1024, spp=1, depth=4,
                                   bvh="sweep", aperture=0.0, flavour="parity")
    ptr = rs.desc_ptr()
    total_ms = 0.0
    for idx, wrap_u, wrap_v, levels in vimg.desc_image_textures(ptr):
        if len(levels) < 2:
            continue
        built, ms = vimg.build_mip_chain(levels[0], wrap_u, wrap_v)
        total_ms += ms
        for got, ref in zip(built, levels[1:]):
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"texture {idx}"
    tex, row_ref, col_ref = vimg.desc_envmap(ptr)
    level0 = [lv for i, _, _, lv in vimg.desc_image_textures(ptr) if i == tex][0][0]
    row, col, ms = vimg.build_envmap_cdf(level0)
    print(f"mip chains {total_ms:.3f} ms, env-map CDFs {ms:.3f} ms on the device")
    assert np.array_equal(col.view(np.uint32), col_ref.view(np.uint32))
    assert np.array_equal(row.view(np.uint32), row_ref.view(np.uint32))


@pytest.mark.gpu
def test_prep_edge_cases(vimg):
    # 1x1 and 2x2 images have no further level; an all-black env map gets the uniform CDFs
    assert vimg.build_mip_chain(np.ones((1, 1, 3), np.float32))[0] == []
    assert vimg.build_mip_chain(np.ones((2, 2, 3), np.float32))[0] == []
    lv, _ = vimg.build_mip_chain(np.full((5, 3, 3), 0.25, np.float32), 2, 1)  # non-square, odd sizes
    assert [a.shape for a in lv] == [(2, 1, 3)]
    row, col, _ = vimg.build_envmap_cdf(np.zeros((4, 8, 3), np.float32))
    assert np.array_equal(row, np.arange(5, dtype=np.float32) / 4)
    assert np.array_equal(col, np.tile(np.arange(9, dtype=np.float32) / 8, (4, 1)))
    with pytest.raises(vimg.VimgError):
        vimg.build_envmap_cdf(np.zeros((4, 8, 3), np.float32), device=99)


@pytest.mark.gpu
@pytest.mark.parametrize("name,key", [("synthetic_small", "mis_64"), ("synthetic_rg", "mis_64")])
def test_scene_create_builds_missing_levels_and_cdfs(vimg, name, key):
    """the adapter may hand over level 0 only (levels 1.. NULL) and no env-map CDFs: vimg_scene_create builds
    them on the device, and the render is the reference's image bit for bit -- as with the host-built ones"""
    from helpers import golden_ref, manifest
    v = manifest()[name]["ref"][key]
    d = vimg.LoadedDesc(golden_scene_path(name))
    n_tex, cdfs = vimg.strip_prepared(d.ptr)
    assert n_tex >= 2 and cdfs
    sc = vimg.Scene(d.ptr)
    img, st = sc.scene_integrator("mis", v["spp"], v["depth"])
    assert np.array_equal(img, golden_ref(name)[key])
    sc.close()


def test_stripped_descriptor_cannot_be_saved(tmp_path):
    vimg = load_product()
    d = vimg.LoadedDesc(golden_scene_path("synthetic_small"))
    vimg.strip_prepared(d.ptr)
    lib = vimg.load_library()
    assert lib.vimg_scene_desc_save(d.ptr, str(tmp_path / "x.vscene").encode()) != 0
