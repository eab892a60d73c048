"""BVH heat map (src/integrators/heatmap.cpp:39-147).

CPU: the plain-C restatement (oracle/vimg_oracle.c) against the golden images the compiled
reference's heatmap_img() produced (tests/golden/make_golden_heatmap.py).  GPU: vimg_heatmap
through the C ABI against the same goldens.  Bar: bit for bit (the cost is a sum of 0.5 / 1.0
steps along the bit-exact closest-hit traversal, the colour map is a float polynomial).
"""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, golden_scene_path, load_product, manifest

SMALL = ["cornell_small", "disney_small", "glass_small", "synthetic_small", "synthetic_rg", "cornell_dof",
         "sphere_light_small", "cornell_checker"]
CASES = [(4, 20.0), (1, 7.5)]


@pytest.fixture(scope="module")
def golden_heat():
    with open(os.path.join(GOLDEN, "heatmap", "digests.json")) as f:
        digests = json.load(f)
    return np.load(os.path.join(GOLDEN, "heatmap", "heatmap.npz")), digests


@pytest.mark.parametrize("name", SMALL)
def test_restatement_heatmap_matches_the_reference(golden_heat, name):
    from oracle import refcpu
    if not os.path.exists(refcpu.restatement_path()):
        pytest.skip("oracle/_ref/libvimg_oracle.so not built")
    vimg = load_product()
    e = manifest()[name]
    d = vimg.LoadedDesc(golden_scene_path(name))
    for spp, factor in CASES:
        rgb, cost = refcpu.restatement_heatmap(d.ptr, e["width"], e["height"], spp, factor)
        assert np.array_equal(rgb.view(np.uint32), golden_heat[0][f"{name}|{spp}|{factor}"])
        assert refcpu.digest(rgb) == golden_heat[1][f"{name}|{spp}|{factor}"]
        assert (cost == np.floor(cost)).all() and cost.min() >= 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL)
def test_device_heatmap_bit_exact(vimg, golden_heat, name):
    from test_gpu_parity import scene
    for spp, factor in CASES:
        rgb, cost = scene(vimg, name).heatmap(spp, factor)
        assert np.array_equal(rgb.view(np.uint32), golden_heat[0][f"{name}|{spp}|{factor}"]), (name, spp)
        assert (cost == np.floor(cost)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell_box_spheres", "disney_spheres", "glass_in_box_sweep"])
def test_device_heatmap_full_size_digest(vimg, refcpu, golden_heat, name):
    from test_gpu_parity import scene
    for spp, factor in CASES:
        rgb, _ = scene(vimg, name).heatmap(spp, factor)
        assert refcpu.digest(rgb) == golden_heat[1][f"{name}|{spp}|{factor}"]


@pytest.mark.gpu
def test_heatmap_cost_is_the_counted_traversal(vimg):
    """the heat map doubles as a check of the traversal counters behind the roofline: the mean
    cost per camera ray equals 0.5 + (inner nodes + primitive tests) / rays of a counted render"""
    from test_gpu_parity import scene
    sc = scene(vimg, "synthetic_small")
    _, cost = sc.heatmap(1, 20.0)  # 1 spp: no truncation of a mean, cost = the ray's own cost
    _, st = sc.scene_integrator("s_normal", 1, 64, counters=True)
    expect = 0.5 * st["rays_closest"] + st["inner_nodes"] + st["tri_tests"] + st["sphere_tests"]
    # truncation drops the .5 of every ray
    assert float(cost.sum()) == expect - 0.5 * st["rays_closest"]


@pytest.mark.gpu
def test_heatmap_arguments(vimg):
    from test_gpu_parity import scene
    sc = scene(vimg, "cornell_small")
    a, _ = sc.heatmap(4, 0.0)  # factor <= 0 -> the reference's default 20
    b, _ = sc.heatmap(4, 20.0)
    assert np.array_equal(a, b)
    with pytest.raises(vimg.VimgError):
        sc.heatmap(0, 20.0)
