#!/usr/bin/env python
"""Adds to tests/golden/manifest.json the digests of the reference's mis and material images of the BASELINE configs
C1 / C2 / C3 exactly as their scene files state them (full resolution, full spp: glass_in_box 640x480
200 spp, cornell_box_spheres 800x800 100 spp, disney_spheres 1800x800 512 spp depth -1), rendered by the
compiled reference (oracle/_ref, parity flavour) from the flattened scene.  The images themselves are not
stored (17 MB for C3); the GPU test compares digests.  Run it AFTER make_golden.py (which rewrites the
manifest).  Takes a few minutes of CPU:

    python tests/golden/make_golden_full.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refcpu  # noqa: E402
from helpers import golden_scene_path, load_product  # noqa: E402

vimg = load_product()
path = os.path.join(ROOT, "tests", "golden", "manifest.json")
with open(path) as f:
    manifest = json.load(f)
for name in ("glass_in_box", "cornell_box_spheres", "disney_spheres"):
    e = manifest[name]
    d = vimg.LoadedDesc(golden_scene_path(name))
    for integ in ("mis", "material"):
        if f"{integ}_full" in e["ref"] and "--force" not in sys.argv:
            continue
        rs = refcpu.RefScene.from_desc(d.ptr, e["spp"], e["depth"], integ)
        img, sec = rs.render(integ, spp=e["spp"], depth=e["depth"])
        rs.close()
        if np.isnan(img).any():  # disney_spheres under `material` at depth -1: NaN pixels in the reference itself
            print(name, integ, "reference image has NaN pixels: no digest stored", flush=True)
            continue
        e["ref"][f"{integ}_full"] = {"spp": e["spp"], "depth": e["depth"], "hash": refcpu.digest(img),
                                     "mean": [float(v) for v in img.reshape(-1, 3).mean(0)]}
        print(name, integ, e["ref"][f"{integ}_full"], f"{sec:.1f} s", flush=True)
        with open(path, "w") as f:
            json.dump(manifest, f, indent=1, sort_keys=True)
with open(path, "w") as f:
    json.dump(manifest, f, indent=1, sort_keys=True)
