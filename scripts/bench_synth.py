"""Throughput of the cuda device on the large synthetic scenes (C4 / C5): scene built by the
compiled reference on this host (test infrastructure), then timed on the GPU with counters for the
HBM roofline.  usage: bench_synth.py G width height spp [integrator]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import load_product
from oracle import refcpu
G, W, H, spp = [int(a) for a in sys.argv[1:5]]
integ = sys.argv[5] if len(sys.argv) > 5 and not sys.argv[5].startswith("-") else "mis"
streams = int(os.environ.get("SPP_STREAMS", "1"))
fast = os.environ.get("FAST", "0") == "1"
vimg = load_product()
t0 = time.time()
rs = refcpu.RefScene.synthetic(G=G, width=W, height=H, tex_size=2048, env_w=2048, env_h=1024, spp=spp, depth=64,
                               bvh="sweep", aperture=0.02, flavour="perf")
t1 = time.time()
sc = vimg.Scene(rs.desc_ptr())
t2 = time.time()
BY = {"inner_nodes": 56, "leaf_nodes": 8, "tri_tests": 56, "sphere_tests": 24}
out = {"G": G, "tris": rs.n_prims, "nodes": rs.n_nodes, "bvh_depth": rs.max_depth, "res": [W, H], "spp": spp,
       "integrator": integ, "spp_streams": streams, "fast": fast, "host_build_s": t1 - t0, "bvh_build_s": rs.bvh_build_s, "upload_s": t2 - t1,
       "scene_bytes": sc.device_bytes}
for rep in range(3):
    img, st = sc.scene_integrator(integ, spp, 64, spp_streams=streams, fast=fast)
rays = st["rays_closest"] + st["rays_shadow"]
out.update(kernel_ms=st["kernel_ms"], mrays_s=rays / st["kernel_ms"] / 1e3, spp_mpix_s=W * H * spp / st["kernel_ms"] / 1e3,
           rays_per_sample=rays / (W * H * spp))
# counted with the traversal-stack mode of the timed render (streams > 1 => fresh stacks)
_, c = sc.scene_integrator(integ, min(spp, 4), 64, counters=True, fresh_stack=streams > 1)
cr = c["rays_closest"] + c["rays_shadow"]
bpr = sum(BY[k] * c[k] for k in BY) / cr
out.update(bytes_per_ray=bpr, inner_per_ray=c["inner_nodes"] / cr, leaf_per_ray=c["leaf_nodes"] / cr,
           tri_per_ray=c["tri_tests"] / cr, achieved_gbs=bpr * rays / st["kernel_ms"] / 1e6)
if "--cpu" in sys.argv:
    cpu_spp = 1
    _, sec = rs.render(integ, spp=cpu_spp, threads=0)
    _, sec = rs.render(integ, spp=cpu_spp, threads=0)
    out.update(cpu_s=sec, cpu_spp=cpu_spp, cpu_spp_mpix_s=W * H * cpu_spp / sec / 1e6, cpu_threads=refcpu.hardware_threads())
print(json.dumps(out))
