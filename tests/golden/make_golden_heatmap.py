#!/usr/bin/env python
"""Regenerates tests/golden/heatmap/heatmap.npz from the compiled reference (oracle/_ref, parity
flavour): heatmap_img() (src/integrators/heatmap.cpp:39-147) at 4 spp (what main.cpp's -m uses)
and at 1 spp, factor 20 and 7.5, for the small golden scenes; blake2b digests for the full-size
BASELINE scenes.  Run in the build container after `make -C oracle`:

    python tests/golden/make_golden_heatmap.py 2> /dev/null
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refcpu  # noqa: E402
from helpers import golden_scene_path, load_product, manifest  # noqa: E402

SMALL = ["cornell_small", "disney_small", "glass_small", "synthetic_small", "synthetic_rg", "cornell_dof",
         "sphere_light_small", "cornell_checker"]
FULL = ["cornell_box_spheres", "disney_spheres", "glass_in_box_sweep"]
vimg = load_product()
saved = os.dup(1)
os.dup2(2, 1)  # the reference prints its progress with fmt::print
out, digests = {}, {}
for name in SMALL + FULL:
    d = vimg.LoadedDesc(golden_scene_path(name))
    rs = refcpu.RefScene.from_desc(d.ptr, 4, 64)
    for spp, factor in ((4, 20.0), (1, 7.5)):
        img = rs.heatmap(spp, factor)
        key = f"{name}|{spp}|{factor}"
        digests[key] = refcpu.digest(img)
        if name in SMALL:
            out[key] = img.view(np.uint32)
os.dup2(saved, 1)
dst = os.path.join(ROOT, "tests", "golden", "heatmap")
os.makedirs(dst, exist_ok=True)
np.savez_compressed(os.path.join(dst, "heatmap.npz"), **out)
with open(os.path.join(dst, "digests.json"), "w") as f:
    json.dump(digests, f, indent=1, sort_keys=True)
print(f"wrote {len(out)} images, {os.path.getsize(os.path.join(dst, 'heatmap.npz')) / 1e6:.2f} MB; {len(digests)} digests")
