"""Repeat one render N times: frame-time stability (outliers) of the queue wavefront."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, manifest, golden_scene_path
name = sys.argv[1]; spp = int(sys.argv[2]); k = int(sys.argv[3]); reps = int(sys.argv[4])
vimg = load_product(); e = manifest()[name]
integ = ["s_normal", "g_normal", "material", "mis"][e["integrator"]]
sc = vimg.Scene.load(golden_scene_path(name))
ts = []
for i in range(reps):
    _, st = sc.render_device(integ, spp, e["depth"], spp_streams=k)
    ts.append(st["kernel_ms"])
print(f"{name} {spp}spp streams={k}: " + " ".join(f"{t:.1f}" for t in ts), flush=True)
