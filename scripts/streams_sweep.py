"""spp_streams sweep: frame time and relMSE against the single-stream (reference-identical) image."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, manifest, golden_scene_path, rel_mse
name = sys.argv[1] if len(sys.argv) > 1 else "disney_spheres"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ks = [int(k) for k in (sys.argv[3] if len(sys.argv) > 3 else "1,4,8,16,32").split(",")]
vimg = load_product(); e = manifest()[name]
integ = ["s_normal", "g_normal", "material", "mis"][e["integrator"]]
sc = vimg.Scene.load(golden_scene_path(name))
sc.scene_integrator(integ, 8, e["depth"])
base = None
for k in ks:
    img, st = sc.scene_integrator(integ, spp, e["depth"], spp_streams=k)
    img = img.copy()
    rays = st["rays_closest"] + st["rays_shadow"]
    if base is None:
        base = img
    print(f"{name} {integ} {spp}spp streams={k}: {st['kernel_ms']:.2f} ms, {rays} rays, "
          f"{rays/st['kernel_ms']/1e3:.1f} Mrays/s, launches {st['kernel_launches']}, "
          f"relMSE vs first {rel_mse(img, base):.3e}, mean {img.mean():.6f}", flush=True)
