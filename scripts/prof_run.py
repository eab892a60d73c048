"""Tiny driver for ncu: render one golden scene once (plus one warm-up launch)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, manifest, golden_scene_path
name = sys.argv[1] if len(sys.argv) > 1 else "disney_spheres"
integ = sys.argv[2] if len(sys.argv) > 2 else "mis"
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
streams = int(sys.argv[5]) if len(sys.argv) > 5 else 1
fast = (sys.argv[6] == "fast") if len(sys.argv) > 6 else False
vimg = load_product(); e = manifest()[name]
sc = vimg.Scene.load(golden_scene_path(name))
for i in range(reps):
    img, st = sc.scene_integrator(integ, spp, e["depth"], spp_streams=streams, fast=fast)
    rays = st["rays_closest"] + st["rays_shadow"]
    print(f"{name} {integ} {spp}spp streams={streams} fast={fast} nan={st['nan_samples']} mean={float(img.mean()):.6f}: {st['kernel_ms']:.2f} ms, {rays} rays, {rays/st['kernel_ms']/1e3:.1f} Mrays/s, {e['width']*e['height']*spp/st['kernel_ms']/1e3:.2f} spp*Mpix/s")
