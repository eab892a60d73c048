"""One-off (GPU box): BASELINE configs C1 / C2 / C3 at FULL resolution and FULL spp, mis integrator, cuda device
(one RNG stream per pixel, reference traversal stack) against the compiled reference (parity flavour) rendered
on this host with all its threads.  Prints pixels differing, relMSE, times.  usage: full_size_parity.py [names]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, golden_scene_path, manifest, rel_mse
from oracle import refcpu
vimg = load_product()
man = manifest()
for name in sys.argv[1:] or ["glass_in_box", "cornell_box_spheres", "disney_spheres"]:
    e = man[name]
    spp, depth = e["spp"], e["depth"]
    sc = vimg.Scene.load(golden_scene_path(name))
    t0 = time.time()
    g, st = sc.scene_integrator("mis", spp, depth)
    t1 = time.time()
    d = vimg.LoadedDesc(golden_scene_path(name))
    rs = refcpu.RefScene.from_desc(d.ptr, spp, depth, "mis")
    r, sec = rs.render("mis", spp=spp, depth=depth)
    diff = int((g != r).any(-1).sum())
    print(f"{name} {e['width']}x{e['height']} {spp} spp mis depth {depth}: {diff} of {e['width'] * e['height']} pixels differ, "
          f"relMSE {rel_mse(g, r):.3e}; device {st['kernel_ms']:.0f} ms ({st['rays_closest'] + st['rays_shadow']} rays), "
          f"reference (parity build, {refcpu.hardware_threads() if hasattr(refcpu, 'hardware_threads') else '?'} threads) {sec:.1f} s", flush=True)
    if diff:
        idx = np.argwhere((g != r).any(-1))[:5]
        for (row, x) in idx:
            print(f"   x={x} y={e['height'] - 1 - row}: device {g[row, x]} reference {r[row, x]}")
    rs.close(); sc.close()
