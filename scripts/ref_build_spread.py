"""How far apart are the reference's OWN two builds?  Renders one golden scene with the compiled reference in its
scalar no-contraction build (parity flavour: the bit-exact oracle) and in the -O3 -march AVX2 + FMA build its
CMake produces (perf flavour), same spp, same seeds, and prints the per-channel ratio of the image means and
the relMSE between them.  CPU only (test infrastructure).  usage: ref_build_spread.py [scene] [spp] [threads]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_product, manifest, golden_scene_path, rel_mse
from oracle import refcpu
name = sys.argv[1] if len(sys.argv) > 1 else "disney_spheres"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
vimg = load_product(); e = manifest()[name]
d = vimg.LoadedDesc(golden_scene_path(name))
imgs = {}
for fl in ("parity", "perf"):
    rs = refcpu.RefScene.from_desc(d.ptr, spp, e["depth"], "mis", flavour=fl)
    imgs[fl], sec = rs.render("mis", spp=spp, threads=threads)
    print(f"{name} mis {spp} spp, {fl} build: {sec:.1f} s, image mean {imgs[fl].reshape(-1, 3).astype(np.float64).mean(0)}", file=sys.stderr)
m = {k: v.reshape(-1, 3).astype(np.float64).mean(0) for k, v in imgs.items()}
print(f"{name} mis {spp} spp: perf / parity image mean {m['perf'] / m['parity']}, relMSE(perf, parity) {rel_mse(imgs['perf'], imgs['parity']):.3e}")
