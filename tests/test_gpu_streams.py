"""Sample-range streams (vimg_render_params.spp_streams, include/vimg_cuda.h).

spp_streams = 1 (default) follows the reference's single PCG stream per pixel
(include/integrators.h:116-131) and is bit-identical to the CPU renderer (test_gpu_parity.py).
spp_streams = K > 1 splits a pixel's samples into K contiguous ranges with their own streams
(initstate = image_index, initseq = range), keeping the reference's R2 jitter index x + y + s:
the same estimator driven by different random numbers.  The bar for it is the north star's
tolerance bar for material / mis images:

    relMSE(GPU_K at N spp, reference at N spp) <= 1.25 * 2 sigma^2 / N

i.e. what two independent N-sample renders of the reference itself differ by, with sigma^2 (the
per-sample relative variance) taken from the reference's own N vs 4N images (prefix trick,
SURVEY.md section 8(d)(i)), and per-channel image means within 1 %.
"""
import numpy as np
import pytest

from helpers import golden_ref, golden_scene_path, rel_mse
from test_gpu_parity import scene, sigma2_from_prefix

pytestmark = pytest.mark.gpu
NO_DEPTH = 0xFFFFFFFF


@pytest.mark.parametrize("name,key", [("cornell_small", "mis_64"), ("disney_small", "mis_64"),
                                      ("glass_small", "material_64")])
def test_zero_and_one_stream_are_the_reference_stream(vimg, man, name, key):
    v = man[name]["ref"][key]
    integ = key.rsplit("_", 1)[0]
    ref = golden_ref(name)[key]
    for k in (0, 1):
        img, _ = scene(vimg, name).scene_integrator(integ, v["spp"], v["depth"], spp_streams=k)
        assert np.array_equal(img, ref)


@pytest.mark.parametrize("k", [2, 4, 16, 64])
@pytest.mark.parametrize("name,key", [("cornell_small", "mis_256"), ("disney_small", "mis_256"),
                                      ("cornell_small", "mis_64"), ("disney_small", "mis_64")])
def test_streams_within_tolerance_of_the_reference(vimg, man, name, key, k):
    v = man[name]["ref"][key]
    ref = golden_ref(name)
    sigma2 = sigma2_from_prefix(ref, "mis_64", "mis_256", 64, 256)
    bound = 1.25 * 2.0 * sigma2 / v["spp"]
    img, st = scene(vimg, name).scene_integrator("mis", v["spp"], v["depth"], spp_streams=k)
    err = rel_mse(img, ref[key])
    print(f"{name}/{key} streams={k}: relMSE {err:.3e}, independent-render level {2 * sigma2 / v['spp']:.3e}")
    assert err <= bound
    gm, rm = img.reshape(-1, 3).mean(0), ref[key].reshape(-1, 3).mean(0)
    assert np.all(np.abs(gm - rm) <= 0.01 * np.maximum(rm, 1e-3))
    assert st["nan_samples"] == 0 and st["samples"] == v["spp"] * man[name]["width"] * man[name]["height"]


@pytest.mark.parametrize("name,key", [("cornell_small", "material_64"), ("sphere_light_small", "material_64"),
                                      ("cornell_skylight", "mis_64"), ("synthetic_small", "mis_64"),
                                      ("glass_small", "mis_64")])
def test_streams_keep_the_image_mean(vimg, man, name, key):
    """scenes without a 4N reference image: energy check (per-channel mean within 1.5 %) and a
    relMSE no worse than 3x the error of the 16-stream image against itself at other K"""
    v = man[name]["ref"][key]
    integ = key.rsplit("_", 1)[0]
    ref = golden_ref(name)[key]
    imgs = {}
    for k in (8, 32):
        imgs[k], st = scene(vimg, name).scene_integrator(integ, v["spp"], v["depth"], spp_streams=k)
        gm, rm = imgs[k].reshape(-1, 3).mean(0), ref.reshape(-1, 3).mean(0)
        assert np.all(np.abs(gm - rm) <= 0.015 * np.maximum(rm, 1e-3)), (k, gm, rm)
        assert st["nan_samples"] == 0
    # three renders of the same estimator with unrelated random numbers: pairwise errors agree
    e_ref8, e_ref32, e_8_32 = rel_mse(imgs[8], ref), rel_mse(imgs[32], ref), rel_mse(imgs[8], imgs[32])
    print(f"{name}/{key}: relMSE K8-ref {e_ref8:.3e} K32-ref {e_ref32:.3e} K8-K32 {e_8_32:.3e}")
    assert max(e_ref8, e_ref32) <= 3.0 * e_8_32 + 1e-6


def test_streams_are_deterministic_and_rank_independent(vimg):
    """range sums are combined in range order on the device: same bits on every run and for every
    tile partition (so an NCCL sum of the ranks' frames equals the one-GPU frame)"""
    sc = scene(vimg, "disney_small")
    a, st = sc.scene_integrator("mis", 24, NO_DEPTH, spp_streams=8)
    b, _ = sc.scene_integrator("mis", 24, NO_DEPTH, spp_streams=8)
    assert np.array_equal(a, b)
    for n_ranks in (2, 5):
        acc = np.zeros_like(a)
        samples = 0
        for r in range(n_ranks):
            part, pst = sc.scene_integrator("mis", 24, NO_DEPTH, rank=r, n_ranks=n_ranks, spp_streams=8)
            assert not (part != 0).any(-1)[vimg.tile_owner(sc.width, sc.height, n_ranks) != r].any()
            acc += part
            samples += pst["samples"]
        assert np.array_equal(acc, a) and samples == st["samples"]


@pytest.mark.parametrize("spp,k", [(7, 3), (4, 16), (1, 8), (5, 5), (33, 32)])
def test_ragged_ranges(vimg, spp, k):
    """ranges are [k*spp/K, (k+1)*spp/K): uneven lengths, K clamped to spp; every sample is taken once"""
    sc = scene(vimg, "cornell_small")
    img, st = sc.scene_integrator("mis", spp, 64, spp_streams=k)
    one, st1 = sc.scene_integrator("mis", spp, 64)
    assert st["samples"] == st1["samples"] == spp * sc.width * sc.height
    assert st["nan_samples"] == 0 and np.isfinite(img).all()
    # depth 0: the path ends at the camera ray's hit, one closest ray per sample whatever the stream
    _, sd = sc.scene_integrator("mis", spp, 0, spp_streams=k)
    assert sd["rays_closest"] == sd["samples"] == spp * sc.width * sc.height
    assert abs(float(img.mean()) - float(one.mean())) <= 0.25 * float(one.mean())
    if spp == 1 or k == 1:
        assert np.array_equal(img, one)  # a single range is the reference stream


def test_streams_with_material_depth0_and_trace_pixel(vimg):
    sc = scene(vimg, "cornell_small")
    img, st = sc.scene_integrator("material", 4, 0, spp_streams=4)
    assert not img.any() and st["samples"] == 4 * sc.width * sc.height


def test_streams_need_the_queue_pipeline(vimg):
    sc = scene(vimg, "cornell_small")
    with pytest.raises(vimg.VimgError):
        sc.scene_integrator("mis", 8, 64, spp_streams=4, pipeline="block")


def test_full_size_c3_streams_against_the_reference_stream(vimg, man):
    """BASELINE config C3 (disney_spheres 1800x800, depth -1): 16 streams at 64 spp against the
    single-stream image (bit-identical to the CPU reference, test_gpu_parity.py) at the level two
    independent 64-spp renders differ by."""
    sc = scene(vimg, "disney_spheres")
    lo, _ = sc.scene_integrator("mis", 16, NO_DEPTH)
    lo = lo.copy()
    hi, st1 = sc.scene_integrator("mis", 64, NO_DEPTH)
    hi = hi.copy()
    sigma2 = rel_mse(lo, hi) / (1.0 / 16 - 1.0 / 64)
    img, st = sc.scene_integrator("mis", 64, NO_DEPTH, spp_streams=16)
    err = rel_mse(img, hi)
    print(f"C3 64 spp, 16 streams: relMSE {err:.3e}, independent-render level {2 * sigma2 / 64:.3e}")
    assert err <= 1.25 * 2.0 * sigma2 / 64
    gm, rm = img.reshape(-1, 3).mean(0), hi.reshape(-1, 3).mean(0)
    assert np.all(np.abs(gm - rm) <= 0.005 * rm)
    rays, rays1 = st["rays_closest"] + st["rays_shadow"], st1["rays_closest"] + st1["rays_shadow"]
    assert abs(rays - rays1) <= 0.002 * rays1 and st["samples"] == st1["samples"]
