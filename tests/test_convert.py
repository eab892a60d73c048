"""glTF texture conversions on the device: make_texture's byte -> float
loop followed by ImageTexture::convert_sRGB_to_linear / convert_RGB_to_normal
(src/scene_loading/gltf_loading.cpp:133-166, src/image_texture.cpp:257-275), and the Image case as a
whole (bytes -> linear level 0 -> mip chain).

Oracle: tests/golden/prep/convert.npz = what the compiled reference's OWN two functions return
(tests/golden/make_golden_convert.py); the plain-C restatement (oracle/vimg_oracle.c) is pinned on it
in the CPU suite and then serves the larger cases.  Bar: bit-exact (every float).
"""
import os

import numpy as np
import pytest

from helpers import GOLDEN, NORMAL_SCALES, convert_inputs


@pytest.fixture(scope="module")
def golden_convert():
    return np.load(os.path.join(GOLDEN, "prep", "convert.npz"))


@pytest.fixture(scope="module")
def restatement():
    from oracle import refcpu
    if not os.path.exists(refcpu.restatement_path()):
        pytest.skip("oracle/_ref/libvimg_oracle.so not built")
    return refcpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_restatement_matches_the_reference_functions(restatement, golden_convert):
    for name, rgb8 in convert_inputs().items():
        assert np.array_equal(bits(restatement.restatement_srgb8_to_linear(rgb8)), golden_convert[f"{name}|linear_bits"])
        for scale in NORMAL_SCALES:
            got = restatement.restatement_rgb8_to_normal(rgb8, scale)
            assert np.array_equal(bits(got), golden_convert[f"{name}|normal|{scale}_bits"]), (name, scale)


def test_golden_conversions_look_like_the_functions(golden_convert):
    """sanity of the reference-made vectors: monotone transfer curve with the right end points, unit normals"""
    rgb8 = convert_inputs()["all_bytes"]
    lin = golden_convert["all_bytes|linear_bits"].view(np.float32)
    curve = np.empty(256, np.float32)
    curve[rgb8[..., 0].ravel()] = lin[..., 0].ravel()
    assert curve[0] == 0 and curve[255] == 1 and (np.diff(curve) > 0).all()
    assert abs(curve[10] - 10 / 255 / 12.92) < 1e-9 and abs(curve[128] - 0.2158605) < 1e-6
    for scale in NORMAL_SCALES:
        n = golden_convert[f"random|normal|{scale}_bits"].view(np.float32)
        assert np.allclose(np.linalg.norm(n.astype(np.float64), axis=-1), 1, atol=1e-6)


@pytest.mark.gpu
def test_conversions_bit_exact_against_the_reference_functions(vimg, golden_convert):
    for name, rgb8 in convert_inputs().items():
        lin, _ = vimg.convert_srgb8_to_linear(rgb8)
        assert lin.shape == rgb8.shape
        assert np.array_equal(bits(lin), golden_convert[f"{name}|linear_bits"]), name
        for scale in NORMAL_SCALES:
            nrm, _ = vimg.convert_rgb8_to_normal(rgb8, scale)
            assert np.array_equal(bits(nrm), golden_convert[f"{name}|normal|{scale}_bits"]), (name, scale)


@pytest.mark.gpu
def test_conversions_at_2k_and_ragged_sizes(vimg, restatement):
    """2048^2 random texture (both conversions, kernel times printed) and every size around the 16-byte /
    4-pixel vector paths, against the restatement"""
    rng = np.random.default_rng(7)
    big = rng.integers(0, 256, size=(2048, 2048, 3), dtype=np.uint8)
    lin, ms_l = vimg.convert_srgb8_to_linear(big)
    nrm, ms_n = vimg.convert_rgb8_to_normal(big, 0.75)
    assert np.array_equal(bits(lin), bits(restatement.restatement_srgb8_to_linear(big)))
    assert np.array_equal(bits(nrm), bits(restatement.restatement_rgb8_to_normal(big, 0.75)))
    gb = big.size * 5 / 1e9  # 1 byte read + 4 written per channel
    print(f"2048^2: sRGB->linear {ms_l:.3f} ms ({gb / ms_l * 1e3:.0f} GB/s), RGB->normal {ms_n:.3f} ms "
          f"({gb / ms_n * 1e3:.0f} GB/s)")
    for n in list(range(1, 40)) + [255, 256, 257, 1023, 4099]:
        img = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
        assert np.array_equal(bits(vimg.convert_srgb8_to_linear(img)[0]), bits(restatement.restatement_srgb8_to_linear(img))), n
        assert np.array_equal(bits(vimg.convert_rgb8_to_normal(img, 2.0)[0]), bits(restatement.restatement_rgb8_to_normal(img, 2.0))), n


@pytest.mark.gpu
@pytest.mark.parametrize("wrap", [(0, 0), (2, 1)])
def test_image_texture_from_bytes(vimg, restatement, wrap):
    """make_texture's Image case in one call: level 0 = the converted bytes, levels 1.. = the mip chain the
    restatement (pinned on the reference-built levels, tests/test_prep.py) builds from that level 0"""
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, size=(96, 160, 3), dtype=np.uint8)
    levels, ms = vimg.build_image_texture(img, *wrap)
    level0 = restatement.restatement_srgb8_to_linear(img)
    assert np.array_equal(bits(levels[0]), bits(level0))
    ref_levels = restatement.restatement_mip_chain(level0, *wrap)
    assert len(levels) == 1 + len(ref_levels)
    for lv, (got, ref) in enumerate(zip(levels[1:], ref_levels), start=1):
        assert got.shape == ref.shape and np.array_equal(bits(got), bits(ref)), lv
    # the two-call route gives the same bits
    two, _ = vimg.build_mip_chain(vimg.convert_srgb8_to_linear(img)[0], *wrap)
    for a, b in zip(levels[1:], two):
        assert np.array_equal(bits(a), bits(b))


@pytest.mark.gpu
def test_conversion_edge_cases(vimg):
    empty = np.zeros((0, 3), np.uint8)
    assert vimg.convert_srgb8_to_linear(empty)[0].size == 0 and vimg.convert_rgb8_to_normal(empty)[0].size == 0
    one, _ = vimg.build_image_texture(np.full((1, 1, 3), 255, np.uint8))
    assert len(one) == 1 and np.array_equal(one[0], np.ones((1, 1, 3), np.float32))
    flat = vimg.convert_rgb8_to_normal(np.array([[128, 128, 255]], np.uint8), 0.0)[0]
    assert np.array_equal(flat, np.array([[0, 0, 1]], np.float32))  # scale 0 flattens the map
    with pytest.raises(vimg.VimgError):
        vimg.convert_srgb8_to_linear(np.zeros((4, 3), np.uint8), device=99)
    with pytest.raises(vimg.VimgError):
        vimg.build_image_texture(np.zeros((4, 4, 3), np.uint8), 7, 0)
