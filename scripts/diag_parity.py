"""Diagnosis (GPU box): which pixels / samples of a golden scene differ between the cuda device and the
reference, found with trace_pixel on both sides (same per-pixel RNG stream, so the value at spp = k is the
mean of the first k samples).  usage: diag_parity.py scene [integrator spp]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, golden_ref, golden_scene_path, manifest
from oracle import refcpu
name = sys.argv[1]
integ = sys.argv[2] if len(sys.argv) > 2 else "mis"
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 64
vimg = load_product()
e = manifest()[name]
v = e["ref"][f"{integ}_{spp}"]
sc = vimg.Scene.load(golden_scene_path(name))
d = vimg.LoadedDesc(golden_scene_path(name))
rs = refcpu.RefScene.from_desc(d.ptr, spp, v["depth"], integ)
img, _ = sc.scene_integrator(integ, v["spp"], v["depth"])
ref = golden_ref(name)[f"{integ}_{spp}"]
bad = np.argwhere((img != ref).any(-1))
H, W = img.shape[:2]
print(f"{name} {integ} {spp}: {len(bad)} of {H * W} pixels differ")
for (row, x) in bad[:int(os.environ.get("DIAG_PIXELS", "6"))]:
    y = H - 1 - row
    first = None
    for k in range(1, spp + 1):
        g = sc.trace_pixel(integ, int(x), int(y), k, v["depth"])
        r = rs.trace_pixel(integ, int(x), int(y), spp=k, depth=v["depth"])
        if not np.array_equal(g, r):
            first = (k, g, r)
            break
    if first is None:
        print(f"  pixel x={x} y={y}: frame differs but trace_pixel agrees up to {spp} spp; frame {img[row, x]} ref {ref[row, x]}")
        continue
    k, g, r = first
    gp = sc.trace_pixel(integ, int(x), int(y), k - 1, v["depth"]) if k > 1 else np.zeros(3, np.float32)
    print(f"  pixel x={x} y={y}: first differing sample index {k - 1}; sample (device) ~ {g * k - gp * (k - 1)}, "
          f"(reference) ~ {r * k - gp * (k - 1)}; bits {g.view(np.uint32)} vs {r.view(np.uint32)}")
