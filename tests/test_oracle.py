"""Pins the oracle: the compiled reference (oracle/_ref, parity flavour) must reproduce
(1) the hashes SURVEY.md section 8(c) recorded from the reference, (2) tests/golden, and it must do so
from the flattened .vscene files alone (adapter -> file -> rebuilt reference objects)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, golden_ref, golden_scene_path

# FNV-1a-64 of the sample-0 primary-hit ids, sweep-SAH BVH, parity flavour
SURVEY_HASHES = {"cornell_box_spheres_sweep": ("22079188903141ba", 0.908270),
                 "glass_in_box_sweep": ("92aa7d6c3b75ed46", 1.000000),
                 "disney_spheres": ("5cfddb2d87478f8c", 0.932294)}


@pytest.mark.parametrize("name", sorted(SURVEY_HASHES))
def test_survey_hashes_are_pinned_in_the_manifest(man, name):
    h, frac = SURVEY_HASHES[name]
    assert man[name]["ref"]["ids_sample0_fnv_survey_basis"] == h
    assert abs(man[name]["ref"]["hit_fraction_sample0"] - frac) < 5e-7


def _scene(vimg, refcpu, man, name):
    d = vimg.LoadedDesc(golden_scene_path(name))
    e = man[name]
    return refcpu.RefScene.from_desc(d.ptr, e["spp"], e["depth"]), d


@pytest.mark.parametrize("name", ["cornell_box_spheres_sweep", "glass_in_box_sweep"])
def test_oracle_reproduces_survey_hash_from_vscene(vimg, refcpu, man, name):
    sc, _keep = _scene(vimg, refcpu, man, name)
    ids, _ = sc.trace_ids(1)
    assert refcpu.fnv1a64(ids, refcpu.SURVEY_FNV_BASIS) == SURVEY_HASHES[name][0]


@pytest.mark.parametrize("name", ["cornell_small", "glass_small", "disney_small", "sphere_light_small",
                                  "cornell_checker", "cornell_dof", "synthetic_small", "synthetic_rg"])
def test_oracle_matches_golden_ids_and_normals(vimg, refcpu, man, name):
    sc, _keep = _scene(vimg, refcpu, man, name)
    r = man[name]["ref"]
    ids, t = sc.trace_ids(r["ids_spp"])
    assert refcpu.digest(ids) == r["ids_hash"]
    assert refcpu.digest(t) == r["t_hash"]
    for integ in ("s_normal", "g_normal"):
        img, _ = sc.render(integ, spp=4)
        assert refcpu.digest(img) == r[integ + "_hash"]


@pytest.mark.parametrize("name,key", [("cornell_small", "mis_64"), ("glass_small", "material_64"),
                                      ("disney_small", "mis_64"), ("cornell_skylight", "mis_64"),
                                      ("sphere_light_medium", "mis_64"), ("synthetic_rg", "mis_64")])
def test_oracle_matches_golden_images(vimg, refcpu, man, name, key):
    sc, _keep = _scene(vimg, refcpu, man, name)
    v = man[name]["ref"][key]
    img, _ = sc.render(key.rsplit("_", 1)[0], spp=v["spp"], depth=v["depth"])
    assert refcpu.digest(img) == v["hash"]
    assert np.array_equal(img, golden_ref(name)[key])


def test_oracle_thread_count_does_not_change_the_image(vimg, refcpu, man):
    sc, _keep = _scene(vimg, refcpu, man, "cornell_small")
    a, _ = sc.render("mis", spp=8, threads=1)
    b, _ = sc.render("mis", spp=8, threads=5)
    assert np.array_equal(a, b)


def test_material_and_mis_agree_statistically(man):
    # both are unbiased estimators of the same image on scenes without Dielectric
    ref = golden_ref("cornell_small")
    mis, mat = ref["mis_256"].reshape(-1, 3).mean(0), ref["material_64"].reshape(-1, 3).mean(0)
    assert np.all(np.abs(mis - mat) / mis < 0.02)


def test_dielectric_mis_quirk_is_in_the_golden_images():
    # under `mis` every path through a Dielectric gets throughput 0, so the
    # mis image of glass_in_box is darker than the material image
    ref = golden_ref("glass_small")
    assert ref["mis_64"].mean() < 0.93 * ref["material_64"].mean()


def test_libm_restatement_is_bit_exact_on_this_host():
    exe = os.path.join(ROOT, "oracle", "_ref", "libm_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/libm_check not built")
    out = subprocess.run([exe, "quick"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout


@pytest.mark.parametrize("name", ["cornell_small", "glass_small", "disney_small", "sphere_light_medium",
                                  "cornell_dof", "synthetic_rg"])
def test_c_restatement_matches_the_compiled_reference(vimg, man, name):
    """oracle/vimg_oracle.c (plain C, no glm shim, reference BVH layout) against the golden
    outputs of the compiled reference: ids, hit distances and both normal images, bit for bit."""
    from oracle import refcpu
    if not os.path.exists(refcpu.restatement_path()):
        pytest.skip("oracle/_ref/libvimg_oracle.so not built")
    e = man[name]
    r = e["ref"]
    d = vimg.LoadedDesc(golden_scene_path(name))
    ids, t, sn, gn, unsupported = refcpu.restatement_primary(d.ptr, e["width"], e["height"], r["ids_spp"])
    assert refcpu.digest(ids) == r["ids_hash"]
    assert refcpu.digest(t) == r["t_hash"]
    if r["ids_spp"] == 4:  # the normal integrators always run 4 spp (src/main.cpp:223,232)
        assert refcpu.digest(gn) == r["g_normal_hash"]
        if not unsupported:
            assert refcpu.digest(sn) == r["s_normal_hash"]


def test_c_restatement_full_size_c2(vimg, man):
    from oracle import refcpu
    if not os.path.exists(refcpu.restatement_path()):
        pytest.skip("oracle/_ref/libvimg_oracle.so not built")
    e = man["cornell_box_spheres"]
    d = vimg.LoadedDesc(golden_scene_path("cornell_box_spheres"))
    ids, t, sn, gn, _ = refcpu.restatement_primary(d.ptr, e["width"], e["height"], 4)
    assert refcpu.digest(ids) == e["ref"]["ids_hash"] and refcpu.digest(t) == e["ref"]["t_hash"]
    assert refcpu.digest(sn) == e["ref"]["s_normal_hash"] and refcpu.digest(gn) == e["ref"]["g_normal_hash"]


def test_binding_flavour_links_the_adapter_and_the_product_library():
    """oracle/_ref/libvimg_ref_binding.so = reference sources + the adapter WITH scene_integrator_cuda +
    libvimg_cuda.so (INTEGRATION.md's build): it loads and exports the call; running it needs a GPU
    (tests/test_gpu_binding.py)"""
    from oracle import refcpu
    if not refcpu.available("binding"):
        pytest.skip("binding flavour not built")
    import ctypes
    lib = ctypes.CDLL(refcpu.lib_path("binding"))
    assert hasattr(lib, "vref_render_cuda") and hasattr(lib, "vref_render")
    assert not hasattr(ctypes.CDLL(refcpu.lib_path("parity")), "vref_render_cuda")
