#!/usr/bin/env python
"""Regenerates tests/golden/ from the compiled reference (oracle/_ref, parity flavour).

Run in the build container, where /root/reference exists:

    make -C oracle && make -C v-img_b200/csrc && python tests/golden/make_golden.py

Outputs
  scenes/<name>.vscene   the reference-loaded scene + host-built BVH, flattened by the
                         product's adapter and serialised with vimg_scene_desc_save()
  ref/<name>.npz         what the reference renders for that scene: primary-hit primitive
                         ids (all samples), s_normal / g_normal images (bit-exact targets),
                         material / mis images at the listed spp (tolerance targets)
  manifest.json          per-scene settings, sizes and FNV-1a-64 hashes

Nothing here is read by the product path; tests/ compare the CUDA device against it.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refcpu  # noqa: E402
from helpers import load_product  # noqa: E402

REF_SCENES = "/root/reference/scenes"
OUT = os.path.dirname(os.path.abspath(__file__))


def variant(name, res=None, integrator=None, spp=None, depth=None, edit=None):
    """Temp copy of a reference scene file with some settings changed."""
    with open(os.path.join(REF_SCENES, name)) as f:
        j = json.load(f)
    if res:
        j["camera"]["resolution"] = list(res)
    if integrator:
        j["integrator"] = {"type": integrator}
    if spp:
        j["sampler"]["samples"] = spp
    if depth is not None:
        j["sampler"]["depth"] = depth
    if edit:
        edit(j)
    fd, path = tempfile.mkstemp(suffix=".json")
    with os.fdopen(fd, "w") as f:
        json.dump(j, f)
    return path


def checker_floor(j):
    for m in j["materials"]:
        if m["name"] == "white":
            m.pop("albedo", None)
            m["texture"] = {"type": "checkered", "width": 6, "height": 5,
                            "col1": [0.8, 0.8, 0.2], "col2": [0.1, 0.2, 0.7]}


def main():
    vimg = load_product()
    os.makedirs(os.path.join(OUT, "scenes"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "ref"), exist_ok=True)
    manifest = {}

    def emit(name, sc, ids_spp=0, normals=False, images=(), full_ids=False, stride=1, note=""):
        """images: list of (integrator, spp, depth or 0)"""
        path = os.path.join(OUT, "scenes", name + ".vscene")
        vimg.renderer.save_desc(sc.desc_ptr(), path)
        entry = {"scene": f"scenes/{name}.vscene", "width": sc.width, "height": sc.height,
                 "spp": sc.spp, "depth": sc.depth, "integrator": sc.integrator,
                 "n_prims": sc.n_prims, "n_nodes": sc.n_nodes, "max_depth": sc.max_depth,
                 "note": note, "ref": {}}
        arrays = {}
        if ids_spp:
            ids, t = sc.trace_ids(ids_spp)
            entry["ref"]["ids_spp"] = ids_spp
            entry["ref"]["ids_hash"] = refcpu.digest(ids)
            entry["ref"]["ids_sample0_fnv_survey_basis"] = refcpu.fnv1a64(
                np.ascontiguousarray(ids[:, 0]), refcpu.SURVEY_FNV_BASIS)
            entry["ref"]["t_hash"] = refcpu.digest(t)
            entry["ref"]["hit_fraction_sample0"] = float((ids[:, 0] >= 0).mean())
            if full_ids:
                arrays["ids"] = ids
                arrays["t"] = t
        if normals:
            for integ in ("s_normal", "g_normal"):
                img, _ = sc.render(integ, spp=4)
                entry["ref"][integ + "_hash"] = refcpu.digest(img)
                arrays[integ + "_strided"] = img[::stride, ::stride].copy()
            entry["ref"]["normal_stride"] = stride
        for integ, spp, depth in images:
            img, sec = sc.render(integ, spp=spp, depth=depth)
            key = f"{integ}_{spp}"
            arrays[key] = img
            entry["ref"][key] = {"spp": spp, "depth": depth or sc.depth, "hash": refcpu.digest(img),
                                 "mean": [float(v) for v in img.reshape(-1, 3).mean(0)]}
        np.savez_compressed(os.path.join(OUT, "ref", name + ".npz"), **arrays)
        manifest[name] = entry
        print(name, {k: v for k, v in entry.items() if k != "ref"}, flush=True)

    # ---- BASELINE.json configs C1-C3 at scene-file resolution (bit-exact targets) ---------
    sc = refcpu.RefScene.from_json(f"{REF_SCENES}/cornell_box_spheres.json", bvh="binned")
    emit("cornell_box_spheres", sc, ids_spp=4, normals=True, full_ids=True, stride=5,
         note="C2, binned BVH (CLI default -b 0)")
    sc = refcpu.RefScene.from_json(f"{REF_SCENES}/cornell_box_spheres.json", bvh="sweep")
    emit("cornell_box_spheres_sweep", sc, ids_spp=1, normals=True, stride=8, note="C2, sweep BVH (-b 1)")
    sc = refcpu.RefScene.from_json(f"{REF_SCENES}/glass_in_box.json", bvh="binned")
    emit("glass_in_box", sc, ids_spp=4, normals=True, stride=5, note="C1, binned BVH")
    sc = refcpu.RefScene.from_json(f"{REF_SCENES}/glass_in_box.json", bvh="sweep")
    emit("glass_in_box_sweep", sc, ids_spp=1, note="C1 with sweep BVH (SURVEY hash check)")
    sc = refcpu.RefScene.from_json(f"{REF_SCENES}/disney_spheres.json", bvh="sweep")
    emit("disney_spheres", sc, ids_spp=2, normals=True, stride=10, note="C3, sweep BVH (README timing)")

    # ---- small-resolution variants with material / mis reference images ---------------------
    p = variant("cornell_box_spheres.json", res=(160, 160))
    sc = refcpu.RefScene.from_json(p, bvh="binned")
    emit("cornell_small", sc, ids_spp=4, normals=True, full_ids=True,
         images=[("mis", 64, 0), ("material", 64, 0), ("mis", 256, 0)], note="C2 at 160x160")
    p = variant("glass_in_box.json", res=(160, 120))
    sc = refcpu.RefScene.from_json(p, bvh="binned")
    emit("glass_small", sc, ids_spp=2, normals=True, full_ids=True,
         images=[("mis", 64, 0), ("material", 64, 0)], note="C1 at 160x120 (Dielectric, A.5 quirk)")
    p = variant("disney_spheres.json", res=(225, 100))
    sc = refcpu.RefScene.from_json(p, bvh="sweep")
    emit("disney_small", sc, ids_spp=2, normals=True, full_ids=True,
         images=[("mis", 64, 0), ("material", 32, 0), ("mis", 256, 0)],
         note="C3 at 225x100 (Principled, depth -1, RR)")
    for nm in ("sphere_light_medium_mis", "sphere_light_small_mis"):
        p = variant(f"MIS_light_tests/{nm}.json", res=(128, 128))
        sc = refcpu.RefScene.from_json(p, bvh="binned")
        emit(nm.replace("_mis", ""), sc, ids_spp=1, normals=True, full_ids=True,
             images=[("mis", 64, 0), ("material", 64, 8)], note="sphere area light (Sphere::sample)")
    p = variant("cornell_box_spheres.json", res=(128, 128), edit=checker_floor)
    sc = refcpu.RefScene.from_json(p, bvh="binned")
    emit("cornell_checker", sc, ids_spp=1, normals=True, full_ids=True,
         images=[("mis", 64, 0)], note="Checkerboard texture on the white walls")
    p = variant("cornell_box_spheres.json", res=(128, 128))
    sc = refcpu.RefScene.from_json(p, bvh="binned")
    sc.set_const_background((0.4, 0.5, 0.9), as_light=True)
    emit("cornell_skylight", sc, images=[("mis", 64, 0), ("material", 64, 0)],
         note="emissive ConstBackground added to list_lights (ConstBackground::sample / pdf)")
    p = variant("cornell_box_spheres.json", res=(128, 128))
    sc = refcpu.RefScene.from_json(p, bvh="binned")
    sc.set_camera_dof(12.0, 800.0)
    emit("cornell_dof", sc, ids_spp=4, normals=True, full_ids=True,
         note="thin-lens camera, aperture > 0 (rand2 gets the first draw, A.4)")

    # ---- synthetic C4-style scene, small: mesh normals, mip texture, normal map, env map, DOF --
    sc = refcpu.RefScene.synthetic(G=48, width=192, height=108, tex_size=128, env_w=128, env_h=64,
                                   spp=64, depth=64, bvh="sweep", with_rg=False, aperture=0.02)
    emit("synthetic_small", sc, ids_spp=4, normals=True, full_ids=True,
         images=[("mis", 64, 0), ("material", 32, 0)], note="C4 recipe, G=48, 128^2 textures")
    sc = refcpu.RefScene.synthetic(G=40, width=160, height=90, tex_size=64, env_w=64, env_h=32,
                                   spp=64, depth=64, bvh="binned", with_rg=True, aperture=0.0)
    emit("synthetic_rg", sc, ids_spp=2, normals=True, full_ids=True,
         images=[("mis", 64, 0)], note="C4 recipe + TextureRG (a30), binned BVH, pinhole")

    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
