"""Shared test helpers (test infrastructure)."""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_product():
    """Import the product package (its directory name has a hyphen: v-img_b200)."""
    if "v_img_b200" in sys.modules:
        return sys.modules["v_img_b200"]
    pkg = os.path.join(ROOT, "v-img_b200")
    spec = importlib.util.spec_from_file_location("v_img_b200", os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["v_img_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def golden_scene_path(name):
    return os.path.join(GOLDEN, manifest()[name]["scene"])


def golden_ref(name):
    return np.load(os.path.join(GOLDEN, "ref", name + ".npz"))


def rel_mse(img, ref, eps=1e-2):
    """relMSE = mean((img - ref)^2 / (ref^2 + eps)) on linear float images."""
    img = np.asarray(img, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.mean((img - ref) ** 2 / (ref ** 2 + eps)))


def post_inputs():
    """Linear images the post-process tests feed to the chain of src/main.cpp:304-356: three
    reference renders from tests/golden/ref and a synthetic ramp with the awkward values
    (NaN, +-inf, negative, zero, denormal, huge, the OETF's branch point, values around 1)."""
    out = {}
    for name, key in (("disney_small", "mis_256"), ("cornell_small", "mis_256"), ("synthetic_small", "mis_64")):
        out[f"{name}/{key}"] = np.ascontiguousarray(golden_ref(name)[key], np.float32)
    rng = np.random.default_rng(20261018)
    ramp = (rng.random((64, 257, 3), dtype=np.float32) ** 4 * 20).astype(np.float32)
    ramp[1] = np.linspace(0, 1.25, 257, dtype=np.float32)[:, None]
    ramp[2] = np.linspace(0.0031, 0.0032, 257, dtype=np.float32)[:, None]
    ramp[3] = np.geomspace(1e-12, 1e6, 257).astype(np.float32)[:, None]
    specials = [[np.nan, 1, 2], [1, np.nan, 0.5], [np.inf, 0.5, 0.1], [-1, -2, 0.3], [0, 0, 0], [1e30, 1e-30, 5],
                [-np.inf, 1, 1], [1e-42, 1e-40, 1e-39], [1, 1, 1], [0.0031308, 0.0031307, 0.0031309],
                [-0.0, 0.0, -0.0], [3.4e38, 3.4e38, 3.4e38], [0.5, 0.25, 0.125]]
    ramp[0, :len(specials)] = np.array(specials, np.float32)
    out["ramp"] = ramp
    return out


NORMAL_SCALES = (1.0, 0.5, 2.0)


def convert_inputs():
    """8-bit RGB images the texture-conversion tests feed to make_texture's Image / Normals cases
    (src/scene_loading/gltf_loading.cpp:133-166): every byte value in every channel (with the
    other channels varying), and a random 96x61 image (odd width: vector-path tails)."""
    out = {}
    b = np.arange(256, dtype=np.uint8)
    out["all_bytes"] = np.stack([b, np.roll(b, 85), b[::-1]], axis=1).reshape(16, 16, 3)
    out["grid"] = np.stack(list(np.meshgrid(b[::5], b[::5], indexing="ij")) + [np.full((52, 52), 255, np.uint8)], axis=2)
    rng = np.random.default_rng(20261019)
    out["random"] = rng.integers(0, 256, size=(61, 97, 3), dtype=np.uint8)
    return out
