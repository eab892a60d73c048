"""vimg_render_multi: several GPUs of one box behind the C ABI (include/vimg_cuda.h).  Needs >= 2 devices
(gpurun --gpus 2); with one device only the n = 1 path runs."""
import os

import numpy as np
import pytest

from helpers import golden_scene_path, rel_mse

pytestmark = pytest.mark.gpu
NO_DEPTH = 0xFFFFFFFF


def _devices(vimg, want):
    n = vimg.device_count()
    if n < want:
        pytest.skip(f"needs {want} CUDA devices, {n} visible")
    return list(range(want))


def test_one_device_is_vimg_render(vimg):
    d = vimg.LoadedDesc(golden_scene_path("disney_small"))
    one, st1 = vimg.Scene(d.ptr).scene_integrator("mis", 16, NO_DEPTH)
    multi, st = vimg.render_multi(d.ptr, [0], "mis", 16, NO_DEPTH)
    assert np.array_equal(one, multi) and st["samples"] == st1["samples"]


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("name,integ,kw", [("disney_small", "mis", {}), ("disney_small", "mis", {"spp_streams": 8}),
                                           ("cornell_small", "material", {}), ("cornell_small", "s_normal", {}),
                                           ("disney_small", "mis", {"spp_streams": 8, "fast": True})])
def test_tile_partition_is_bit_identical_to_one_device(vimg, man, n, name, integ, kw):
    devs = _devices(vimg, n)
    d = vimg.LoadedDesc(golden_scene_path(name))
    sc = vimg.Scene(d.ptr)
    one, st1 = sc.scene_integrator(integ, 16, man[name]["depth"], **kw)
    sc.close()
    for no_peer in ("0", "1"):  # stores into device 0's frame over NVLink / partial frames through host memory
        os.environ["VIMG_MULTI_NO_PEER"] = no_peer
        multi, st = vimg.render_multi(d.ptr, devs, integ, 16, man[name]["depth"], partition="tile", **kw)
        assert np.array_equal(one, multi), (no_peer, int((one != multi).any(-1).sum()))
        assert st["samples"] == st1["samples"] and st["rays_closest"] == st1["rays_closest"]
    del os.environ["VIMG_MULTI_NO_PEER"]


@pytest.mark.parametrize("n", [2, 4])
def test_spp_partition_is_the_streams_image(vimg, man, n):
    """sample ranges dealt to devices: the same range sums as one device with spp_streams = 8, added per device
    first -- equal up to the association of 8 float additions"""
    devs = _devices(vimg, n)
    name = "disney_small"
    d = vimg.LoadedDesc(golden_scene_path(name))
    sc = vimg.Scene(d.ptr)
    one, st1 = sc.scene_integrator("mis", 64, NO_DEPTH, spp_streams=8)
    sc.close()
    multi, st = vimg.render_multi(d.ptr, devs, "mis", 64, NO_DEPTH, partition="spp", spp_streams=8)
    assert st["samples"] == st1["samples"] and st["rays_closest"] == st1["rays_closest"]
    assert np.allclose(one, multi, rtol=2e-6, atol=1e-7), float(np.abs(one - multi).max())
    assert rel_mse(multi, one) < 1e-12


def test_spp_partition_rank_by_rank(vimg, man):
    """one device, the ranks rendered one after the other through vimg_render (what a torchrun rank does)"""
    name = "cornell_small"
    sc = vimg.Scene.load(golden_scene_path(name))
    one, st1 = sc.scene_integrator("mis", 32, 64, spp_streams=6)
    acc = np.zeros_like(one)
    samples = 0
    for r in range(3):
        part, st = sc.scene_integrator("mis", 32, 64, spp_streams=6, rank=r, n_ranks=3, partition="spp")
        acc += part
        samples += st["samples"]
    assert samples == st1["samples"]
    assert np.allclose(acc, one, rtol=2e-6, atol=1e-7)
    with pytest.raises(vimg.VimgError):
        sc.scene_integrator("mis", 32, 64, spp_streams=2, rank=0, n_ranks=3, partition="spp")
    with pytest.raises(vimg.VimgError):
        sc.scene_integrator("s_normal", 4, 64, rank=0, n_ranks=2, partition="spp")


def test_render_multi_rejects_bad_arguments(vimg):
    d = vimg.LoadedDesc(golden_scene_path("cornell_small"))
    with pytest.raises(vimg.VimgError):
        vimg.render_multi(d.ptr, [0, 0], "mis", 4, 64)
    with pytest.raises(vimg.VimgError):
        vimg.render_multi(d.ptr, [vimg.device_count()], "mis", 4, 64)
