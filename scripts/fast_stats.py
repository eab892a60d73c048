"""Systematic part of the tolerance mode's difference from the bit-identical image, measured like the reference's
own build-to-build spread (scripts/ref_build_spread.py): relMSE of the 8x8-block means.  usage: fast_stats.py [scene] [spp]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, manifest, golden_scene_path, rel_mse
name = sys.argv[1] if len(sys.argv) > 1 else "disney_spheres"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 512
vimg = load_product(); e = manifest()[name]
sc = vimg.Scene.load(golden_scene_path(name))


def down(x, k=8):
    h, w, _ = x.shape
    return x[:h // k * k, :w // k * k].reshape(h // k, k, w // k, k, 3).mean((1, 3))


exact, _ = sc.scene_integrator("mis", spp, e["depth"]); exact = exact.copy()
lo, _ = sc.scene_integrator("mis", spp // 4, e["depth"]); lo = lo.copy()
sigma2 = rel_mse(lo, exact) / (4.0 / spp - 1.0 / spp)
print(f"{name} {spp} spp: sigma2 {sigma2:.4f}, independent-render level {2 * sigma2 / spp:.3e}, its 8x8-block level {2 * sigma2 / spp / 64:.3e}")
for label, kw, env in (("64 streams, exact math", {"spp_streams": 64}, None), ("fast bvh, 1 stream", {"fast": True}, "0"),
                       ("fast bvh, 64 streams", {"fast": True, "spp_streams": 64}, "0"), ("fast flat, 1 stream", {"fast": True}, "32"),
                       ("fast flat, 64 streams", {"fast": True, "spp_streams": 64}, "32")):
    if env is not None:
        os.environ["VIMG_FLAT_MAX"] = env
    img, st = sc.scene_integrator("mis", spp, e["depth"], **kw)
    print(f"  {label:26s} relMSE {rel_mse(img, exact):.3e}  8x8-block relMSE {rel_mse(down(img), down(exact)):.3e}  "
          f"mean/exact {img.reshape(-1, 3).mean(0) / exact.reshape(-1, 3).mean(0)}  nan {st['nan_samples']}  {st['kernel_ms']:.1f} ms", flush=True)
