"""Post-process chain (src/main.cpp:304-356): the oracle side.

CPU tests: the plain-C restatement (oracle/vimg_oracle.c) against the committed golden bytes /
floats that the compiled reference's own functions produced (tests/golden/make_golden_post.py),
and, when oracle/_ref holds the compiled reference, against it live.  Bar: bit for bit.
"""
import os

import numpy as np
import pytest

from helpers import GOLDEN, post_inputs

TONEMAPPERS = ["clamp", "agx", "reinhard", "aces", "none"]


@pytest.fixture(scope="module")
def golden_post():
    return np.load(os.path.join(GOLDEN, "post", "post.npz"))


@pytest.fixture(scope="module")
def refcpu():
    from oracle import refcpu as r
    if not os.path.exists(r.restatement_path()):
        pytest.skip("oracle/_ref/libvimg_oracle.so not built")
    return r


@pytest.mark.parametrize("tm", TONEMAPPERS)
def test_restatement_matches_the_golden_bytes_and_floats(refcpu, golden_post, tm):
    for name, img in post_inputs().items():
        srgb, rgb8 = refcpu.restatement_postprocess(img, tm)
        assert np.array_equal(rgb8, golden_post[f"{name}|{tm}|rgb8"]), (name, tm)
        if name == "ramp":
            assert np.array_equal(srgb.view(np.uint32), golden_post[f"{name}|{tm}|srgb_bits"])


@pytest.mark.parametrize("tm", TONEMAPPERS)
def test_restatement_matches_the_compiled_reference(refcpu, tm):
    if not refcpu.available("parity"):
        pytest.skip("oracle/_ref/libvimg_ref_parity.so not built")
    rng = np.random.default_rng(7)
    img = (rng.random((97, 131, 3), dtype=np.float32) ** 6 * 50).astype(np.float32)
    img[0, 0] = [np.nan, 0, 1]
    img[0, 1] = [-3, np.inf, 0.2]
    a, a8 = refcpu.postprocess(img, tm)
    b, b8 = refcpu.restatement_postprocess(img, tm)
    assert np.array_equal(a8, b8)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_golden_semantics(golden_post):
    """what the goldens pin down: NaN pixel -> magenta, negative -> 0, >= 1 -> 255, black stays 0"""
    g = golden_post["ramp|clamp|rgb8"]
    assert g[0, 0].tolist() == [255, 0, 255] and g[0, 1].tolist() == [255, 0, 255]
    assert g[0, 2].tolist()[0] == 255 and g[0, 3].tolist()[:2] == [0, 0] and g[0, 4].tolist() == [0, 0, 0]
    assert g[0, 8].tolist() == [255, 255, 255]
    assert (np.diff(g[1, :, 0].astype(int)) >= 0).all()  # monotone ramp


def test_empty_image(refcpu):
    srgb, rgb8 = refcpu.restatement_postprocess(np.zeros((0, 3), np.float32), "aces")
    assert srgb.size == 0 and rgb8.size == 0
