"""Device time of the post-process chain on a 4K frame: bytes only (threshold table) vs with the float image (pow)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product
vimg = load_product()
big = (np.random.default_rng(0).random((2160, 3840, 3), dtype=np.float32) * 4).astype(np.float32)
for tm in ("clamp", "aces", "reinhard", "agx", "none"):
    a = min(vimg.postprocess(big, tm, want_srgb=False)[2] for _ in range(5))
    b = min(vimg.postprocess(big, tm, want_srgb=True)[2] for _ in range(5))
    nbytes = big.size // 3 * 15 + (big.size // 3 * 12 if tm == "reinhard" else 0)
    print(f"{tm}: bytes only {a:.3f} ms ({nbytes / a / 1e6:.0f} GB/s), with the float image {b:.3f} ms")
