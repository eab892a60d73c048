"""Diagnosis (GPU box): trace_pixel of one pixel through a debug variant of the library that prints every
bounce (scripts/build_variant.sh debugnee -DVIMG_DEBUG_NEE).  usage: diag_pixel.py scene x y samples"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("VIMG_CUDA_LIB", os.path.join(ROOT, "v-img_b200", "variants", "libvimg_debugnee.so"))
if not os.path.exists(os.environ["VIMG_CUDA_LIB"]):
    sys.exit("build the debug variant first: bash scripts/build_variant.sh debugnee -DVIMG_DEBUG_NEE")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, golden_scene_path
vimg = load_product()
name, x, y, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
sc = vimg.Scene.load(golden_scene_path(name))
import numpy as np
r = sc.trace_pixel("mis", x, y, k, 64, pipeline=os.environ.get("DIAG_PIPELINE"))
print("RESULT", r, r.view(np.uint32), os.path.basename(os.environ["VIMG_CUDA_LIB"]), os.environ.get("DIAG_PIPELINE"))
