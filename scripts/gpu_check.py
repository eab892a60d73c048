"""Ad-hoc parity sweep on a GPU box: CUDA device vs tests/golden (and the live oracle)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product, manifest, golden_ref, golden_scene_path, rel_mse
from oracle import refcpu
vimg = load_product()
INTEG = ["s_normal", "g_normal", "material", "mis"]
man = manifest()
names = sys.argv[1:] or list(man)
for name in names:
    e = man[name]; ref = golden_ref(name)
    t0 = time.time()
    sc = vimg.Scene.load(golden_scene_path(name))
    line = [f"{name}: upload {1e3*(time.time()-t0):.0f}ms bytes={sc.device_bytes}"]
    r = e["ref"]
    if "ids_spp" in r:
        ids, t = sc.trace_ids(r["ids_spp"])
        ok_i = refcpu.digest(ids) == r["ids_hash"]; ok_t = refcpu.digest(t) == r["t_hash"]
        line.append(f"ids {'OK' if ok_i else 'MISMATCH'} t {'OK' if ok_t else 'MISMATCH'}")
        if "ids" in ref.files and not ok_i:
            d = np.nonzero(ids != ref["ids"]); line.append(f"  ids differ at {len(d[0])} / {ids.size}: first {d[0][:5]} {d[1][:5]} gpu {ids[d][:5]} ref {ref['ids'][d][:5]}")
        if "t" in ref.files and not ok_t:
            rt = ref["t"]; d = np.nonzero(~((t == rt) | (np.isnan(t) & np.isnan(rt)))); line.append(f"  t differ at {len(d[0])}: gpu {t[d][:4]} ref {rt[d][:4]}")
    for integ in ("s_normal", "g_normal"):
        if integ + "_hash" in r:
            img, st = sc.scene_integrator(integ, 4, e["depth"])
            ok = refcpu.digest(img) == r[integ + "_hash"]
            s = r["normal_stride"]; sub = ref[integ + "_strided"]
            nd = int((img[::s, ::s] != sub).any(-1).sum())
            line.append(f"{integ} {'OK' if ok else 'MISMATCH'} (strided diff px {nd}, maxabs {np.abs(img[::s,::s]-sub).max():.3g}) {st['kernel_ms']:.2f}ms")
    for key, v in r.items():
        if isinstance(v, dict):
            integ = key.rsplit("_", 1)[0]
            img, st = sc.scene_integrator(integ, v["spp"], v["depth"])
            rimg = ref[key]
            same = float((img == rimg).all(-1).mean())
            line.append(f"{key}: relMSE {rel_mse(img, rimg):.3e} identical_px {same:.4f} mean gpu {img.reshape(-1,3).mean(0).round(5)} ref {np.array(v['mean']).round(5)} nan {st['nan_samples']} rays {st['rays_closest']+st['rays_shadow']} {st['kernel_ms']:.1f}ms")
    print("\n   ".join(line), flush=True)
