"""Post-process chain on the device (vimg_postprocess / vimg_render_srgb8, include/vimg_cuda.h)
against the reference's chain (src/main.cpp:304-356).

Bars
  * bytes (the PNG payload): IDENTICAL to the golden bytes of the compiled reference for every
    tonemapper on every input image;
  * the float image after gamma: +, -, *, / are bit-exact and std::pow / std::log2 on floats (glibc's
    powf / log2f on the CPU) are restated operation by operation on the device (vimg_libm.cuh), so
    every float is IDENTICAL too, for all five modes (agx chains log2 -> polynomial -> pow);
  * the OETF's pow branch is checked over EVERY float in [0.0031308, 1]: the quantised byte is
    and the float after gamma are the same as the restatement's (which calls the host libm) for all
    ~75 M inputs.
"""
import os

import numpy as np
import pytest

from helpers import GOLDEN, post_inputs

pytestmark = pytest.mark.gpu
TONEMAPPERS = ["clamp", "agx", "reinhard", "aces", "none"]


@pytest.fixture(scope="module")
def golden_post():
    return np.load(os.path.join(GOLDEN, "post", "post.npz"))


def ulp_distance(a_bits, b_bits):
    a = a_bits.astype(np.int64)
    b = b_bits.astype(np.int64)
    a = np.where(a < 0x80000000, a, 0x80000000 - a)  # sign-magnitude -> two's complement order
    b = np.where(b < 0x80000000, b, 0x80000000 - b)
    return np.abs(a - b)


@pytest.mark.parametrize("tm", TONEMAPPERS)
def test_bytes_identical_to_the_reference(vimg, golden_post, tm):
    for name, img in post_inputs().items():
        rgb8, srgb, ms = vimg.postprocess(img, tm)
        ref8 = golden_post[f"{name}|{tm}|rgb8"]
        diff = int((rgb8 != ref8).sum())
        assert diff == 0, f"{name}/{tm}: {diff} bytes differ"
        # bytes only: the OETF + quantisation go through the 256-threshold table instead of pow
        only8, none, _ = vimg.postprocess(img, tm, want_srgb=False)
        assert none is None and np.array_equal(only8, ref8), f"{name}/{tm}: threshold path differs"
        if name == "ramp":
            ref_bits = golden_post[f"{name}|{tm}|srgb_bits"]
            got_bits = srgb.view(np.uint32)
            nan_ref = np.isnan(ref_bits.view(np.float32))
            assert np.array_equal(nan_ref, np.isnan(srgb))
            d = ulp_distance(got_bits[~nan_ref], ref_bits[~nan_ref])
            print(f"ramp/{tm}: max ulp distance {int(d.max())}, floats differing {float((d > 0).mean()):.5f}")
            assert d.max() == 0


def test_oetf_pow_branch_every_float(vimg, refcpu):
    """every float p in [0.0031308, 1]: 1.055 * pow(p, 1/2.4) - 0.055 and its byte equal the CPU's"""
    lo = np.array([0.0031308], np.float32).view(np.uint32)[0]
    hi = np.array([1.0], np.float32).view(np.uint32)[0]
    total = differing_floats = 0
    step = 3 * (1 << 22)
    for start in range(int(lo), int(hi) + 1, step):
        bits = np.arange(start, min(start + step, int(hi) + 1), dtype=np.uint32)
        pad = (-len(bits)) % 3
        bits = np.concatenate([bits, np.full(pad, hi, np.uint32)])
        img = bits.view(np.float32).reshape(-1, 3)
        g8, gs, _ = vimg.postprocess(img, "none")
        t8, _, _ = vimg.postprocess(img, "none", want_srgb=False)  # threshold-table path
        cs, c8 = refcpu.restatement_postprocess(img, "none")
        assert np.array_equal(g8, c8), f"bytes differ in [{start:#x}, {start + step:#x})"
        assert np.array_equal(t8, c8), f"threshold-path bytes differ in [{start:#x}, {start + step:#x})"
        differing_floats += int((gs.view(np.uint32) != cs.view(np.uint32)).sum())
        total += len(bits)
    print(f"OETF pow branch: {total} floats, bytes identical, {differing_floats} floats after gamma differ")
    assert differing_floats == 0


@pytest.mark.parametrize("tm", ["clamp", "reinhard"])
def test_render_srgb8_equals_render_then_postprocess(vimg, refcpu, man, tm):
    """scene_integrator + chain on the device == the reference chain applied to the rendered frame"""
    from test_gpu_parity import scene
    sc = scene(vimg, "disney_small")
    img, _ = sc.scene_integrator("mis", 16, 0xFFFFFFFF)
    rgb8, st = sc.render_srgb8("mis", 16, 0xFFFFFFFF, tonemapper=tm)
    _, ref8 = refcpu.restatement_postprocess(img, tm)
    assert np.array_equal(rgb8, ref8)
    assert st["samples"] == 16 * sc.width * sc.height and st["total_ms"] >= st["kernel_ms"] > 0


def test_threshold_path_below_the_pow_branch(vimg, refcpu):
    """linear branch of the OETF, zeros, denormals, negative zero, values above 1: bytes-only path"""
    vals = np.concatenate([np.linspace(0, 0.0031308, 4099, dtype=np.float32), np.array(
        [0.0, -0.0, 1e-45, 1e-39, 0.0031307, 0.0031308, 0.0031309, 1.0, 1.0000001, 2.0, 1e30, np.inf, -np.inf, -1.0],
        np.float32)])
    vals = np.concatenate([vals, np.zeros((-len(vals)) % 3, np.float32)]).reshape(-1, 3)
    for tm in ("none", "clamp"):
        t8, _, _ = vimg.postprocess(vals, tm, want_srgb=False)
        _, c8 = refcpu.restatement_postprocess(vals, tm)
        assert np.array_equal(t8, c8), tm


def test_postprocess_arguments(vimg):
    with pytest.raises(vimg.VimgError):
        vimg.postprocess(np.zeros((4, 3), np.float32), "clamp", device=99)
    rgb8, srgb, _ = vimg.postprocess(np.zeros((0, 3), np.float32), "aces")
    assert rgb8.size == 0
    for n in (1, 2, 3, 5, 1023):  # sizes that are not a multiple of the 4-pixel vector path
        img = np.linspace(0, 2, 3 * n, dtype=np.float32).reshape(n, 3)
        rgb8, srgb, _ = vimg.postprocess(img, "reinhard")
        assert rgb8.shape == (n, 3) and np.isfinite(srgb).all()
