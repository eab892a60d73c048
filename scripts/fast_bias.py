"""Where does the tolerance build's image mean move?  Renders one scene with the bit-exact kernels and
with the tolerance build of every variant library given (scripts/build_fast_variant.sh), one RNG stream
per pixel in all of them, BVH and flat trace.  usage: fast_bias.py scene spp lib [lib ...]"""
import os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    from helpers import load_product, manifest, golden_scene_path
    name, spp, out = sys.argv[2], int(sys.argv[3]), sys.argv[4]
    vimg = load_product(); e = manifest()[name]
    integ = ["s_normal", "g_normal", "material", "mis"][e["integrator"]]
    sc = vimg.Scene.load(golden_scene_path(name))
    res = {}
    streams = int(os.environ.get("STREAMS", "1"))
    for label, kw, env, mode in (("exact", {}, None, "2"), ("fast_bvh", {"fast": True}, "0", "2"), ("fast_flat", {"fast": True}, "32", "1"),
                                 ("fast_fused", {"fast": True}, "32", "2"), ("fast_stream", {"fast": True}, "32", "3")):
        if label != "exact":
            kw["spp_streams"] = streams
        os.environ["VIMG_FLAT_MODE"] = mode
        if env is not None:
            os.environ["VIMG_FLAT_MAX"] = env
        best = 1e30
        for _ in range(2):
            img, st = sc.scene_integrator(integ, spp, e["depth"], **kw)
            best = min(best, st["kernel_ms"])
        res[label] = {"ms": best, "mean": img.reshape(-1, 3).astype(np.float64).mean(0).tolist(), "nan": st["nan_samples"],
                      "rays": st["rays_closest"] + st["rays_shadow"]}
        np.save(out + "." + label + ".npy", img)
    json.dump(res, open(out, "w"))
    sys.exit(0)
from helpers import rel_mse
name, spp = sys.argv[1], int(sys.argv[2])
base = None
for lib in sys.argv[3:]:
    tag = os.path.basename(lib)
    out = f"/tmp/bias_{tag}.json"
    env = dict(os.environ, VIMG_CUDA_LIB=os.path.abspath(lib))
    subprocess.run([sys.executable, __file__, "--child", name, str(spp), out], env=env, check=True)
    r = json.load(open(out))
    ex = np.load(out + ".exact.npy")
    if base is None:
        base = ex
    assert np.array_equal(ex, base), "bit-exact image moved between libraries"
    m0 = np.array(r["exact"]["mean"])
    for k in ("fast_bvh", "fast_flat", "fast_fused", "fast_stream"):
        img = np.load(out + "." + k + ".npy")
        print(f"{tag:24s} {k:10s} {r[k]['ms']:8.2f} ms (exact {r['exact']['ms']:.2f})  mean/exact {np.array(r[k]['mean']) / m0}  "
              f"relMSE {rel_mse(img, ex):.3e}  nan {r[k]['nan']}  rays/exact {r[k]['rays'] / r['exact']['rays']:.5f}", flush=True)
