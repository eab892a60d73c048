#!/usr/bin/env python
"""Regenerates tests/golden/post/post.npz from the compiled reference (oracle/_ref, parity
flavour): the bytes main.cpp would write to the PNG (tonemapper + sRGB OETF + 8-bit quantisation,
src/main.cpp:304-356) and the float image after gamma, for every tonemapper, on the inputs of
tests/helpers.post_inputs().  Run in the build container after `make -C oracle`:

    python tests/golden/make_golden_post.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refcpu  # noqa: E402
from helpers import post_inputs  # noqa: E402

out = {}
for name, img in post_inputs().items():
    for tm in refcpu.TONEMAPPERS:
        srgb, rgb8 = refcpu.postprocess(img, tm)
        out[f"{name}|{tm}|rgb8"] = rgb8
        if name == "ramp":  # the float image after gamma is kept for the ramp only (size)
            out[f"{name}|{tm}|srgb_bits"] = srgb.view(np.uint32)
dst = os.path.join(ROOT, "tests", "golden", "post")
os.makedirs(dst, exist_ok=True)
np.savez_compressed(os.path.join(dst, "post.npz"), **out)
print(f"wrote {len(out)} arrays, {os.path.getsize(os.path.join(dst, 'post.npz')) / 1e6:.2f} MB")
