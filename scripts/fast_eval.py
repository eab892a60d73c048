"""Tolerance mode (VIMG_FLAG_FAST) against the bit-identical mode on one golden scene:
frame time, per-kernel stage times, relMSE and per-channel means.
usage: fast_eval.py [scene] [spp] [streams]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, manifest, golden_scene_path, rel_mse
name = sys.argv[1] if len(sys.argv) > 1 else "disney_spheres"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
streams = int(sys.argv[3]) if len(sys.argv) > 3 else 64
vimg = load_product(); e = manifest()[name]
integ = ["s_normal", "g_normal", "material", "mis"][e["integrator"]]
sc = vimg.Scene.load(golden_scene_path(name))


def run(label, reps=2, **kw):
    best = None
    for _ in range(reps):
        img, st = sc.scene_integrator(integ, spp, e["depth"], **kw)
        if best is None or st["kernel_ms"] < best[1]["kernel_ms"]:
            best = (img.copy(), st)
    img, st = best
    rays = st["rays_closest"] + st["rays_shadow"]
    print(f"{label:34s} {st['kernel_ms']:9.2f} ms  {rays / st['kernel_ms'] / 1e3:8.1f} Mrays/s  rays {rays}  nan {st['nan_samples']}  "
          f"launches {st['kernel_launches']}  mean {img.reshape(-1, 3).mean(0)}", flush=True)
    return img, st


def stages(label, **kw):
    _, st = sc.scene_integrator(integ, spp, e["depth"], stage_times=True, **kw)
    ms, n = st["stage_ms"], st["stage_launches"]
    print(f"{label:34s} stages: " + ", ".join(f"{k} {ms[k]:.1f} ms/{n[k]}" for k in ms if n[k]), flush=True)


exact, _ = run("exact, 1 stream")
fresh, _ = run(f"exact math, {streams} streams", spp_streams=streams)
fast1, _ = run("fast, 1 stream", fast=True)
fastk, _ = run(f"fast, {streams} streams", spp_streams=streams, fast=True)
os.environ["VIMG_FLAT_MAX"] = "0"
fastb, _ = run(f"fast (bvh), {streams} streams", spp_streams=streams, fast=True)
stages(f"fast (bvh), {streams} streams", spp_streams=streams, fast=True)
del os.environ["VIMG_FLAT_MAX"]
stages(f"exact math, {streams} streams", spp_streams=streams)
stages(f"fast, {streams} streams", spp_streams=streams, fast=True)
for label, img in (("exact-math streams", fresh), ("fast 1 stream", fast1), ("fast streams", fastk), ("fast bvh streams", fastb)):
    print(f"relMSE({label} vs exact) = {rel_mse(img, exact):.4e}   mean ratio {img.reshape(-1, 3).mean(0) / exact.reshape(-1, 3).mean(0)}")
