"""New scene + one render per repetition (pools allocated inside the render): looks for slow frames."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, manifest, golden_scene_path
vimg = load_product(); e = manifest()["disney_spheres"]
spp = int(sys.argv[1]); reps = int(sys.argv[2]); k = int(sys.argv[3]) if len(sys.argv) > 3 else 16
desc = vimg.LoadedDesc(golden_scene_path("disney_spheres"))
for rep in range(reps):
    sc = vimg.Scene(desc.ptr, device=0)
    t0 = time.perf_counter()
    _, st = sc.render_device("mis", spp, e["depth"], spp_streams=k)
    t1 = time.perf_counter()
    print(f"rep {rep}: kernel {st['kernel_ms']:.1f} ms wall {1e3*(t1-t0):.1f} ms launches {st['kernel_launches']}", file=sys.stderr, flush=True)
    sc.close()
