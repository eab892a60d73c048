"""Per-rank frame time of an N-rank tile partition, measured on ONE GPU (rank 0's share)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, manifest, golden_scene_path
name = sys.argv[1]; spp = int(sys.argv[2]); n_ranks = int(sys.argv[3]); ks = [int(k) for k in sys.argv[4].split(",")]
vimg = load_product(); e = manifest()[name]
integ = ["s_normal", "g_normal", "material", "mis"][e["integrator"]]
sc = vimg.Scene.load(golden_scene_path(name))
for k in ks:
    ts = []
    for i in range(4):
        _, st = sc.render_device(integ, spp, e["depth"], rank=0, n_ranks=n_ranks, spp_streams=k)
        ts.append(st["kernel_ms"])
    print(f"{name} {spp}spp rank 0 of {n_ranks}, streams={k}: " + " ".join(f"{t:.1f}" for t in ts) + f"  launches {st['kernel_launches']}", flush=True)
