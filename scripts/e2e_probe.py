"""Where the end-to-end step spends its time: scene create / first render / second render / close."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch
from helpers import load_product, manifest, golden_scene_path
vimg = load_product(); e = manifest()["disney_spheres"]
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
desc = vimg.LoadedDesc(golden_scene_path("disney_spheres"))
host = torch.empty((e["height"], e["width"], 3), dtype=torch.float32).pin_memory().numpy()
for rep in range(3):
    t = [time.perf_counter()]
    sc = vimg.Scene(desc.ptr, device=local); t.append(time.perf_counter())
    _, st = sc.scene_integrator("mis", spp, e["depth"], rank=rank, n_ranks=world, out=host, spp_streams=16); t.append(time.perf_counter())
    _, st2 = sc.scene_integrator("mis", spp, e["depth"], rank=rank, n_ranks=world, out=host, spp_streams=16); t.append(time.perf_counter())
    sc.close(); t.append(time.perf_counter())
    d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
    print(f"rank {rank} rep {rep}: create {d[0]:.1f} ms, render1 {d[1]:.1f} (kernel {st['kernel_ms']:.1f}), "
          f"render2 {d[2]:.1f} (kernel {st2['kernel_ms']:.1f}), close {d[3]:.1f}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
