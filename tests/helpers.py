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
