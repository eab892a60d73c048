"""The device's copies of glibc's float functions (v-img_b200/csrc/vimg_libm.cuh) against the host libm,
bit for bit, through the parity hook vimg_libm_eval.

The reference's std::sin / cos / log / log2 / pow / acos / atan2 on floats are glibc's sinf ... atan2f
(include/rng/sampling.h, include/background.h, src/tl_camera.cpp, include/color_utils.h,
src/tonemap/agx.cpp, the Principled BSDF).  The CPU suite checks the SAME header compiled for the host
against the host libm (exhaustively when run by hand: oracle/tools/libm_check.cpp; a sample here);
the GPU suite checks that the device evaluates it to the same bits (IEEE double / float add, mul, fma,
div, sqrt; no contraction, no flush to zero)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT


def float_range(lo, hi, count, rng):
    """`count` floats whose BIT PATTERNS are uniform between those of lo and hi (same sign)"""
    a, b = (np.array([v], np.float32).view(np.uint32)[0] for v in (lo, hi))
    return rng.integers(min(a, b), max(a, b) + 1, size=count, dtype=np.uint32).view(np.float32)


def libm_cases(rng, n):
    """(function, x, y) with the argument ranges the path produces plus the awkward values"""
    special = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 2.0, np.inf, -np.inf, np.nan, 1e-45, 1e-39, 1.1754944e-38,
                        3.4028235e38, 0.04045, 0.0031308, 0.4375, 0.6875, 1.1875, 2.4375, 119.99999], np.float32)
    angle = np.concatenate([float_range(1e-6, 119.9, n, rng), -float_range(1e-6, 119.9, n // 4, rng),
                            special[~(np.abs(special) >= 120) | ~np.isfinite(special)]])  # restated for |x| < 120
    pos = np.concatenate([float_range(1e-45, 3.4e38, n, rng), float_range(0.5, 2.0, n, rng), special])
    unit = np.concatenate([float_range(1e-8, 1.0, n, rng), -float_range(1e-8, 1.0, n, rng),
                           float_range(0.99, 1.0, n // 4, rng), special])
    wide = np.concatenate([rng.integers(0, 2**32, size=n, dtype=np.uint32).view(np.float32), special])
    px = np.concatenate([float_range(1e-6, 1.0, n, rng), float_range(1e-45, 256.0, n, rng), special])
    ex = np.array([1.0 / 2.4, 2.4, 2.2], np.float32)
    py = np.concatenate([ex[rng.integers(0, 3, size=n)], (rng.random(n, dtype=np.float32) - 0.5) * 16, ex[rng.integers(0, 3, size=len(special))]])
    ay = np.concatenate([unit, rng.permutation(wide)])
    ax = np.concatenate([rng.permutation(unit), wide])
    return [("sinf", angle, None), ("cosf", angle, None), ("logf", pos, None), ("log2f", pos, None),
            ("logf", wide, None), ("log2f", wide, None), ("powf", px, py), ("acosf", unit, None),
            ("acosf", wide, None), ("atanf", wide, None), ("atan2f", ay, ax)]


def same_bits(a, b):
    both_nan = np.isnan(a) & np.isnan(b)
    return (a.view(np.uint32) == b.view(np.uint32)) | both_nan


def test_host_build_of_the_header_matches_the_host_libm():
    exe = os.path.join(ROOT, "oracle", "_ref", "libm_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/libm_check not built")
    out = subprocess.run([exe, "quick"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "powf mismatches 0" in out.stdout and "atan2f mismatches 0" in out.stdout and "logf mismatches 0" in out.stdout


@pytest.mark.gpu
def test_device_libm_equals_the_host_libm(vimg):
    from oracle import refcpu
    if not os.path.exists(refcpu.restatement_path()):
        pytest.skip("oracle/_ref/libvimg_oracle.so not built")
    rng = np.random.default_rng(2026)
    for fn, x, y in libm_cases(rng, 1 << 21):
        got = vimg.libm_eval(fn, x, y)
        want = refcpu.host_libm(vimg.LIBM_FUNCTIONS[fn], x, y)
        ok = same_bits(got, want)
        bad = np.flatnonzero(~ok)
        # sinf / cosf are restated for |x| < 120 (the renderer's angles are in [0, 2 pi])
        detail = "" if bad.size == 0 else f"first: x={x[bad[0]]!r} y={None if y is None else y[bad[0]]!r} got={got[bad[0]]!r} want={want[bad[0]]!r}"
        assert bad.size == 0, f"{fn}: {bad.size} of {x.size} differ; {detail}"


@pytest.mark.gpu
def test_device_double_log2_equals_the_host_libm(vimg):
    """std::log2 on doubles (env-map mip level, background.h:100; log2(width * height), texture_RGB.h:141):
    the device's restatement of glibc's double log2 on widened floats, both halves of the result's bits"""
    from oracle import refcpu
    if not os.path.exists(refcpu.restatement_path()):
        pytest.skip("oracle/_ref/libvimg_oracle.so not built")
    rng = np.random.default_rng(7)
    x = np.concatenate([float_range(1e-45, 3.4e38, 1 << 21, rng), float_range(0.9, 1.1, 1 << 21, rng),
                        np.array([1.0, 4096.0, 16384.0, 0.5, 3.0, 1e-40, 2048.0 * 1024.0], np.float32)])
    for fn in ("log2_lo", "log2_hi"):
        got = vimg.libm_eval(fn, x).view(np.uint32)
        want = refcpu.host_libm(vimg.LIBM_FUNCTIONS[fn], x).view(np.uint32)
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, f"{fn}: {bad.size} of {x.size} differ, first x={x[bad[0]]!r}"


@pytest.mark.gpu
def test_libm_eval_arguments(vimg):
    assert vimg.libm_eval("log2f", np.zeros(0, np.float32)).size == 0
    with pytest.raises(vimg.VimgError):
        vimg.libm_eval("powf", np.ones(4, np.float32))  # powf needs y
    with pytest.raises(vimg.VimgError):
        vimg.libm_eval("acosf", np.ones(4, np.float32), device=99)
