#!/usr/bin/env python
"""Regenerates tests/golden/prep/convert.npz from the compiled reference (oracle/_ref, parity
flavour): what ImageTexture::convert_sRGB_to_linear and ImageTexture::convert_RGB_to_normal
(src/image_texture.cpp:257-275) return after make_texture's byte -> float loop
(src/scene_loading/gltf_loading.cpp:139-143) on the inputs of tests/helpers.convert_inputs(),
and the mip chain ImageTexture's constructor builds from the converted colour image.
Run in the build container after `make -C oracle`:

    python tests/golden/make_golden_convert.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refcpu  # noqa: E402
from helpers import convert_inputs, NORMAL_SCALES  # noqa: E402

out = {}
for name, rgb8 in convert_inputs().items():
    out[f"{name}|linear_bits"] = refcpu.convert_srgb8_to_linear(rgb8).view(np.uint32)
    for scale in NORMAL_SCALES:
        out[f"{name}|normal|{scale}_bits"] = refcpu.convert_rgb8_to_normal(rgb8, scale).view(np.uint32)
dst = os.path.join(ROOT, "tests", "golden", "prep")
os.makedirs(dst, exist_ok=True)
np.savez_compressed(os.path.join(dst, "convert.npz"), **out)
print(f"wrote {len(out)} arrays, {os.path.getsize(os.path.join(dst, 'convert.npz')) / 1e6:.2f} MB")
