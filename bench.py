#!/usr/bin/env python
"""bench.py -- throughput of the v-img render hot path on B200 (and the reference beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference]
                    [--mode fast|exact] [--scenes default|none|C1,C2,C4,C5] [--spp S]

One "step" = one pass of the hot path over one frame: every pixel of the workload's scene, all its
samples, through the integrator the scene file names (the reference's scene_integrator<F>() call,
include/integrators.h:36-153).

Headline workload (BASELINE.json configs[2]): `scenes/disney_spheres.json` -- 1800x800, 512 spp, `mis`,
"depth": -1, sweep-SAH BVH -- the scene the north star's ">= 100x" target and the README's only
published number are quoted on.  The other configs ride along in the same JSON line as `scenes`:
C1 glass_in_box, C2 cornell_box_spheres (normal + mis), C4 the 2 M-triangle synthetic mesh (1 GPU,
full spp) and C5 the 10 M-triangle 4K 1024 spp scene, tile-partitioned over the N ranks of this run
with the NCCL reduce inside the timed region.  Small scenes come from tests/golden/scenes/*.vscene (the
reference-loaded scene + host-built BVH, flattened by the adapter); C4 / C5 are built on this host by
the compiled reference's own classes (oracle.refcpu.RefScene.synthetic: mesh, textures, env map,
sweep-SAH BVH -- scene construction, outside every timed region); /root/reference is not read.

metric `Mrays/s`: a ray is one BVH::hit<closest> or BVH::occlude call; rays are
counted by the kernels themselves.  `spp_mpix_per_s` (W*H*spp / s / 1e6) is reported beside it.

Modes.  `value` / `e2e` are measured in the mode --mode names, default `fast`: the TOLERANCE mode
(VIMG_FLAG_FAST + sample-range streams) -- the north star's contract for material / mis images is a
relative-MSE tolerance -- with `value_bit_identical` / `e2e_bit_identical` beside them: the same frame
in the library's default mode (one PCG stream per pixel, the reference's traversal stack, image
bit-identical to the CPU renderer in every pixel).  `--mode exact` makes the bit-identical mode the value.

With --gpus N (torchrun, one rank per GPU) the frame's 8x8 tiles are dealt round-robin to the ranks and
the partial frames are summed with an NCCL reduce to rank 0 inside the timed region (strong scaling:
the frame is fixed); `e2e` ends with the whole frame in rank 0's pinned host memory.

`--impl reference` times the compiled, unmodified reference (oracle/_ref, perf flavour = what its CMake
builds: -O3, AVX2 sibling-pair path, FMA contraction) on all host threads, each step a bounded sample
of the same frame (>= 32 spp; throughput is spp-invariant).  The same measurement is attached to
the cuda arm's line as `cpu_baseline`.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden_scene_path, load_product, manifest  # noqa: E402

INTEG_NAMES = ["s_normal", "g_normal", "material", "mis"]
# algorithmic cost model of SURVEY.md section 8(d), in the REFERENCE's data layout
BYTES = {"inner_nodes": 56, "leaf_nodes": 8, "tri_tests": 56, "sphere_tests": 24}
FLOPS = {"inner_nodes": 50, "tri_tests": 67, "sphere_tests": 40}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent
DEPTH_RR = 0xFFFFFFFF

# BASELINE.json configs.  golden: tests/golden/scenes/<name>.vscene; synthetic: oracle.refcpu.RefScene.synthetic
WORKLOADS = {
    "C1": {"golden": "glass_in_box", "config": "configs[0]"},
    "C2": {"golden": "cornell_box_spheres", "config": "configs[1]", "also": ["s_normal"]},
    "C3": {"golden": "disney_spheres", "config": "configs[2]"},
    "C4": {"synthetic": {"G": 1000, "width": 1366, "height": 768}, "spp": 512, "depth": 64, "integrator": "mis",
           "config": "configs[3]"},
    "C5": {"synthetic": {"G": 2236, "width": 3840, "height": 2160}, "spp": 1024, "depth": 64, "integrator": "mis",
           "config": "configs[4]"},
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def kernel_source_hash():
    """Hash of the device sources: an ncu traffic figure is only quoted for the kernels it was captured on."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "v-img_b200", "csrc")
    for n in sorted(os.listdir(d)):
        if n.endswith((".cuh", ".cu")):
            with open(os.path.join(d, n), "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from this round's committed `ncu --set full`
    capture (profiles/ncu_traffic.json), or None when the capture is of other kernel sources."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, "no capture committed"
    with open(p) as f:
        t = json.load(f)
    if t.get("source_hash") != kernel_source_hash():
        return None, f"capture {t.get('capture')} is of kernel sources {t.get('source_hash')}, not of this build"
    e = t.get("kernels", {}).get(f"{workload}:{kernel}")
    if not e:
        return None, "kernel not in the capture"
    return e, t.get("capture")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                smax.append(float(c[1]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None, "samples": len(sm),
                "reasons": sorted(reasons)}


class StdoutToStderr:
    """The reference's scene_integrator prints its progress with std::cout (integrators.h:48,142) and NCCL
    prints a banner; bench.py's stdout carries exactly one JSON line, so fd 1 points at stderr meanwhile."""

    def __enter__(self):
        import ctypes
        self.libc = ctypes.CDLL(None)
        sys.stdout.flush()
        self.libc.fflush(None)
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        self.libc.fflush(None)  # the reference's text sits in C stdio's buffer when fd 1 is a file
        os.dup2(self.saved, 1)
        os.close(self.saved)


# ------------------------------------------------------------------------------------- workloads
class Workload:
    """One BASELINE config: the flattened scene (vimg_scene_desc) and how to get the reference for it."""

    def __init__(self, key, vimg, spp_override=0):
        w = WORKLOADS[key]
        self.key, self.vimg = key, vimg
        self._ref = {}
        if "golden" in w:
            e = manifest()[w["golden"]]
            self.name = w["golden"]
            self.scene_path = golden_scene_path(self.name)
            self.desc = vimg.LoadedDesc(self.scene_path)
            self.desc_ptr = self.desc.ptr
            self.width, self.height, self.spp, self.depth = e["width"], e["height"], e["spp"], e["depth"]
            self.integrator = INTEG_NAMES[e["integrator"]]
            self.n_prims, self.n_nodes = e["n_prims"], e["n_nodes"]
            self.build = "tests/golden/scenes/%s.vscene (reference-loaded scene, host-built BVH)" % self.name
            self.rs = None
        else:
            from oracle import refcpu  # scene construction by the compiled reference's own classes (untimed)
            s = w["synthetic"]
            self.name = f"synthetic_{s['G']}"
            t0 = time.time()
            self.rs = refcpu.RefScene.synthetic(G=s["G"], width=s["width"], height=s["height"], tex_size=2048, env_w=2048,
                                                env_h=1024, spp=w["spp"], depth=w["depth"], bvh="sweep", aperture=0.02,
                                                flavour="perf")
            self.desc_ptr = self.rs.desc_ptr()
            self.width, self.height, self.spp, self.depth = s["width"], s["height"], w["spp"], w["depth"]
            self.integrator = w["integrator"]
            self.n_prims, self.n_nodes = self.rs.n_prims, self.rs.n_nodes
            self.build = (f"procedural height-field mesh G={s['G']} ({self.n_prims} triangles), 2048^2 mip-mapped colour texture + "
                          f"normal map, 2048x1024 HDR env map, thin lens, sweep-SAH BVH ({self.n_nodes} nodes, depth "
                          f"{self.rs.max_depth}) built by the reference's classes on this host in {time.time() - t0:.1f} s")
            self._ref["perf"] = self.rs
        if spp_override:
            self.spp = spp_override
        self.also = w.get("also", [])
        self.config_name = w["config"]

    def label(self, integ=None, spp=None):
        d = "-1 (RR only)" if self.depth == DEPTH_RR else self.depth
        return f"{self.name} {self.width}x{self.height} {spp or self.spp}spp {integ or self.integrator} depth={d}"

    def reference(self, flavour):
        from oracle import refcpu
        if flavour not in self._ref:
            if not refcpu.available(flavour):
                return None
            if self.rs is not None:  # the synthetic scene exists in the perf flavour only (minutes to rebuild)
                return None
            self._ref[flavour] = refcpu.RefScene.from_desc(self.desc_ptr, 1, self.depth, self.integrator, flavour=flavour)
        return self._ref[flavour]

    def close(self):
        for r in self._ref.values():
            r.close()
        self._ref = {}
        self.rs = None


# --------------------------------------------------------------------------- reference (CPU) arm
def time_reference(w, integ, steps, warmup, budget_s, min_spp=32):
    """Times the compiled reference (perf flavour) on all host threads: one warm call, then `steps` timed
    calls of a bounded sample of the frame (>= min_spp spp each unless the frame has fewer)."""
    from oracle import refcpu
    with StdoutToStderr():
        rs = w.reference("perf")
        if rs is None:
            return None
        cores = refcpu.hardware_threads("perf")
        _, t1 = rs.render(integ, spp=1, threads=cores)  # warms the thread pool / caches; also the cost probe
        _, t2 = rs.render(integ, spp=2, threads=cores)
        per_spp = max(t2 - t1, t2 / 2, 1e-4)
        spp = int(max(min(min_spp, w.spp), min(w.spp, budget_s / (steps + warmup) / per_spp)))
        for _ in range(warmup):
            rs.render(integ, spp=spp, threads=cores)
        t_total = 0.0
        for _ in range(steps):
            _, sec = rs.render(integ, spp=spp, threads=cores)
            t_total += sec
        samples = w.width * w.height * spp * steps
        rays_per_sample, rps_src = None, None
        rc = w.reference("count")  # instrumented flavour of the same reference (small scenes)
        if rc is not None:
            rc.counters(reset=True)
            rc.render(integ, spp=2, threads=cores)
            c = rc.counters()
            rays_per_sample = (c["rays_closest"] + c["rays_shadow"]) / (w.width * w.height * 2)
            rps_src = "count flavour of the reference at 2 spp"
    return {"seconds": t_total, "steps": steps, "spp": spp, "samples": samples, "cores": cores,
            "spp_mpix_per_s": samples / t_total / 1e6, "rays_per_sample": rays_per_sample, "rays_per_sample_source": rps_src}


def cpu_baseline_of(w, integ, r, gpu_rays_per_sample=None):
    rps, src = r["rays_per_sample"], r["rays_per_sample_source"]
    if rps is None and gpu_rays_per_sample:
        # the instrumented reference is not built for this scene: rays per sample as the device counts them in the
        # bit-identical mode, which equal the instrumented reference's wherever both exist (tests/test_gpu_parity.py)
        rps, src = gpu_rays_per_sample, "device counters, bit-identical mode"
    return {"value": r["spp_mpix_per_s"] * rps if rps else None, "unit": "Mrays/s", "cores": r["cores"], "kind": "reference",
            "spp_mpix_per_s": r["spp_mpix_per_s"],
            "sample": f"{w.label(integ, r['spp'])}: full frame at {r['spp']} spp (of {w.spp}), {r['steps']} call(s) after one warm "
                      f"call; compiled reference, perf flavour (-O3 -march=x86-64-v3: AVX2 sibling-pair path + FMA, what the "
                      f"reference's CMake builds with -march=native on an AVX2 host); rays/sample {rps and round(rps, 3)} ({src})"}


def reference_line(args):
    vimg = load_product()  # only for vimg_scene_desc_load (host-side file parsing, no GPU work)
    w = Workload(args.workload, vimg, args.spp)
    r = time_reference(w, w.integrator, args.steps, args.warmup, budget_s=args.ref_budget)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libvimg_ref_perf.so not built"}))
        return
    cb = cpu_baseline_of(w, w.integrator, r)
    line = {"metric": "Mrays/s", "value": cb["value"], "unit": "Mrays/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r["seconds"] / r["steps"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "spp_mpix_per_s": r["spp_mpix_per_s"],
            "config": {"workload": w.label(), "scene": w.build, "prims": w.n_prims, "bvh_nodes": w.n_nodes,
                       "partition": f"reference scene_integrator, {r['cores']} host threads", "l2": "n/a (CPU)",
                       "sample_spp_per_step": r["spp"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------- cuda arm
class Gpu:
    """torch plumbing of one rank: device, process group, barrier, max / sum over ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce_stats(self, values):
        """-> (max over ranks, sum over ranks) of a list of floats"""
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world == 1:
            return t.tolist(), t.tolist()
        tmax = t.clone()
        self.dist.all_reduce(tmax, op=self.dist.ReduceOp.MAX)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return tmax.tolist(), t.tolist()


def mode_kwargs(mode, streams):
    """render options of a mode.  fast: tolerance mode; exact: the library default, bit-identical images"""
    if mode == "fast":
        return {"fast": True, "spp_streams": streams}  # flat lists + streamed queues (<= 32 primitives) or the 4-wide BVH
    return {"spp_streams": 1}


def time_frames(gpu, scene, integ, spp, depth, kw, steps, warmup):
    """W untimed + K timed steps of this rank's tiles (+ NCCL reduce to rank 0); frame stays in HBM.
    -> dict(ms_per_step = max over ranks, rays = all ranks, launches, kernel_ms_max)"""
    torch, dist = gpu.torch, gpu.dist

    def step():
        ptr, st = scene.render_device(integ, spp, depth, rank=gpu.rank, n_ranks=gpu.world, **kw)
        if gpu.world > 1:
            dist.reduce(scene.frame_tensor(ptr), dst=0, op=dist.ReduceOp.SUM)
        return st

    for _ in range(warmup):
        step()
        gpu.flush.zero_()
    gpu.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = kernel_ms = 0.0
    launches = 0
    ev0.record()
    for i in range(steps):
        st = step()
        rays += st["rays_closest"] + st["rays_shadow"]
        kernel_ms += st["kernel_ms"]
        launches += st["kernel_launches"]
        if i + 1 < steps:
            gpu.flush.zero_()
    ev1.record()
    gpu.barrier()
    mx, sm = gpu.reduce_stats([ev0.elapsed_time(ev1), rays, launches, kernel_ms])
    return {"ms_per_step": mx[0] / steps, "seconds": mx[0] / 1e3, "rays": sm[1], "launches": int(sm[2]),
            "kernel_ms_max": mx[3], "steps": steps, "nan_samples": st["nan_samples"]}


def time_e2e(gpu, vimg, w, integ, spp, kw, steps):
    """End to end through the public C-ABI calls with HOST buffers, every step: scene upload H2D
    (vimg_scene_create), render, and the WHOLE frame in rank 0's pinned host memory at the end -- N = 1:
    vimg_render's own D2H; N > 1: partial frames reduced over NCCL to rank 0, then D2H there."""
    torch, dist = gpu.torch, gpu.dist
    host = torch.empty((w.height, w.width, 3), dtype=torch.float32).pin_memory()
    host_np = host.numpy()
    gpu.barrier()
    t0 = time.perf_counter()
    rays, h2d = 0, 0
    for _ in range(steps):
        sc = vimg.Scene(w.desc_ptr, device=gpu.local_rank)
        h2d = sc.device_bytes
        if gpu.world == 1:
            _, st = sc.scene_integrator(integ, spp, w.depth, out=host_np, **kw)
        else:
            ptr, st = sc.render_device(integ, spp, w.depth, rank=gpu.rank, n_ranks=gpu.world, **kw)
            frame = sc.frame_tensor(ptr)
            dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)
            if gpu.rank == 0:
                host.copy_(frame)
            torch.cuda.synchronize()
        rays += st["rays_closest"] + st["rays_shadow"]
        sc.close()
    gpu.barrier()
    mx, sm = gpu.reduce_stats([time.perf_counter() - t0, rays])
    return {"seconds": mx[0], "rays": sm[1], "steps": steps, "h2d": int(h2d)}


# ALGORITHMIC bytes per item of the path-state kernels: what an item must move through HBM in
# the layout of its pipeline -- records read + sectors written + queue records -- not counting scene tables
# (leaf / shade / light records), which are cache resident for small scenes.
STATE_BYTES = {
    # streamed pipeline (vimg_stream.cuh): hit record 32 + slot line 128 read; sectors 0, 2, 3 (96; END: all four) + the hit
    # record of the traced segment (32) written -- the kernel traces the rays it makes itself, there is no ray queue
    "stream": {"shade": 32 + 128 + 96 + 32, "end": 32 + 128 + 128 + 32},
    # slot-record pipeline (vimg_queues.cuh): queue entry 8 + 160-byte slot read; 112 + queue entry written (END: + ray 32)
    "slot": {"shade": 8 + 160 + 112 + 8, "end": 4 + 160 + 112 + 32 + 4},
}
STAGE_KERNEL = {"trace": None, "classify": "qw_classify_kernel", "shade_principled": "shade_kernel<mis, Principled>",
                "shade_lambertian": "shade_kernel<mis, Lambertian>", "shade_dielectric": "shade_kernel<mis, Dielectric>",
                "end": "end_kernel<mis>"}


def pattern_peak(slots, items):
    """What the memory system delivers for the streamed shade / END kernels' access pattern with no arithmetic in the
    way (scripts/gather_peak.cu, built by __graft_entry__.build(): per item a 32-byte queue record in order, a random
    128-byte slot line read, 96 bytes of it written back, a 32-byte record appended): measured here, on this GPU,
    outside every timed region.  roofline.frac stays against the copy peak; this says how much of the gap is the
    pattern's."""
    exe = os.path.join(ROOT, "scripts", "gather_peak")
    if not os.path.exists(exe):
        return None
    try:
        # (rank 0 calls this while the other ranks wait: device 0 of the visible set is its GPU)
        out = subprocess.run([exe, "--one", str(int(slots)), str(int(items))], capture_output=True, text=True, timeout=120)
        r = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # a diagnostic: never fails the bench
        return {"unavailable": repr(e)[:200]}
    return {"value": r["pattern_gbs"], "unit": "GB/s", "copy_gbs_same_run": r["copy_gbs"], "slots": r["slots"], "items": r["items"],
            "bytes_per_item": r["bytes_per_item"],
            "what": "scripts/gather_peak.cu --one: the access pattern of the streamed shade / END kernels without arithmetic "
                    "(32-byte record in order, random 128-byte line read, 96 bytes written back, 32-byte record appended), "
                    "slot numbers in random order over the address range of one serial pool"}


def roofline_of(gpu, scene, w, integ, spp, kw, hbm_peak, peak_src):
    """Roofline of the DOMINANT kernel of one workload on this rank's GPU (rank 0 calls this, whole frame): every
    launch of a serial pass in the measured mode is timed with CUDA events on its stream (VIMG_FLAG_STAGE_TIMES),
    the stage with the largest share is the dominant kernel.  TRACE: algorithmic bytes per ray in the REFERENCE's
    layout, per-ray averages from a counted pass of the bit-exact kernels (the reference's
    traversal).  SHADE / END: path-state bytes per item (STATE_BYTES)."""
    flat = bool(kw.get("fast")) and w.n_prims <= 32
    c_spp = min(spp, 4 if w.n_prims > 100000 else 16)
    _, cst = scene.render_device(integ, c_spp, w.depth, counters=True, spp_streams=1,
                                 fresh_stack=bool(kw.get("fast") or kw.get("spp_streams", 1) > 1))
    c_rays = cst["rays_closest"] + cst["rays_shadow"]
    bytes_per_ray = sum(BYTES[k] * cst[k] for k in BYTES) / c_rays
    flops_per_ray = sum(FLOPS[k] * cst[k] for k in FLOPS) / c_rays
    t_spp = min(spp, 8 if w.n_prims > 100000 else 64)
    kw_t = dict(kw)
    kw_t["spp_streams"] = min(kw.get("spp_streams", 1), t_spp)
    gpu.flush.zero_()
    _, tst = scene.render_device(integ, t_spp, w.depth, stage_times=True, **kw_t)
    total = sum(tst["stage_ms"].values())
    shares = {k: v / total for k, v in tst["stage_ms"].items() if tst["stage_launches"][k]}
    top = max(shares, key=shares.get)
    n_launch = max(tst["stage_launches"][top], 1)
    launch_s = tst["stage_ms"][top] / 1e3 / n_launch
    items_per_launch = tst["stage_items"][top] / n_launch
    layout = "stream" if flat else "slot"
    prefix = "qv_" if flat else "qw_"
    trace_kernel = "qw_trace_wide_kernel" if kw.get("fast") else "qw_trace_kernel"
    if top == "trace":
        kernel = trace_kernel
        per_item, unit_note = bytes_per_ray, ("bytes per ray in the reference's layout (56/inner node, 8/leaf, 56/triangle test, "
                                             f"24/sphere test; SURVEY.md 8(d)), per-ray averages of a counted pass at {c_spp} spp")
    else:
        kernel = prefix + STAGE_KERNEL[top].replace("mis", integ)
        per_item = STATE_BYTES[layout]["end" if top == "end" else "shade"]
        unit_note = (f"path-state bytes per item of the {layout} layout (records read + sectors written + queue records; "
                     "DESIGN.md 5.2), scene tables not counted")
    achieved = per_item * items_per_launch / launch_s / 1e9
    cap, traffic_src = ncu_traffic(kernel, w.key)
    traffic = None
    if cap:
        # the captured launch had cap["items"] items (its queue-record loads, counted by ncu); the bench's average launch has
        # items_per_launch: the DRAM bytes are quoted for the launch size `achieved` is quoted for
        traffic = cap["dram_bytes_per_launch"] * (items_per_launch / cap["items"] if cap.get("items") else 1.0)
        traffic_src = {"capture": traffic_src, "dram_bytes_of_the_captured_launch": cap["dram_bytes_per_launch"],
                       "items_of_the_captured_launch": cap.get("items"),
                       "scaled_to_items_per_launch": bool(cap.get("items"))}
    t_rays = tst["rays_closest"] + tst["rays_shadow"]
    if tst["stage_launches"].get("trace", 0):
        trace_launch_s = tst["stage_ms"]["trace"] / 1e3 / tst["stage_launches"]["trace"]
        rays_per_trace_launch = t_rays / tst["stage_launches"]["trace"]
        trace = {"kernel": trace_kernel, "launch_ms": 1e3 * trace_launch_s,
                 "rays_per_launch": rays_per_trace_launch, "algorithmic_bytes_per_ray": bytes_per_ray,
                 "algorithmic_flops_per_ray": flops_per_ray,
                 "reference_layout_gbs": bytes_per_ray * rays_per_trace_launch / trace_launch_s / 1e9,
                 "fp32_tflops_achieved": flops_per_ray * rays_per_trace_launch / trace_launch_s / 1e12,
                 "note": "the reference-layout figure normalises throughput across scenes; node and primitive records come from L2 / HBM"}
    else:
        # streamed pipeline: no trace launch -- shade / END trace the rays they make over the flat lists in shared memory
        step_s = total / 1e3
        trace = {"kernel": "qv_trace_route (inside qv_shade_kernel / qv_end_kernel)", "launch_ms": None,
                 "rays_per_step": t_rays, "algorithmic_bytes_per_ray": bytes_per_ray, "algorithmic_flops_per_ray": flops_per_ray,
                 "reference_layout_gbs": bytes_per_ray * t_rays / step_s / 1e9,
                 "fp32_tflops_achieved": flops_per_ray * t_rays / step_s / 1e12,
                 "note": "this scene (< 8 KiB) is traced from shared memory at the tail of the kernel that made the ray; the "
                         "reference-layout figures are over the whole serial pass (they normalise throughput across scenes)"}
    pp = None
    if flat and top != "trace":
        pp = pattern_peak(min(2 << 20, w.width * w.height * kw_t.get("spp_streams", 1)), items_per_launch)
        if pp and pp.get("value"):
            pp["achieved_over_it"] = achieved / pp["value"]
    return {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "pattern_peak": pp,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": kernel,
            "launch_ms": 1e3 * launch_s, "launches_timed": n_launch, "items_per_launch": items_per_launch,
            "algorithmic_bytes_per_item": per_item,
            "stage_share_of_step": shares,
            "trace": trace,
            "note": f"dominant kernel = largest stage share of a serial pass at {t_spp} spp, every launch timed with CUDA events "
                    f"on its stream (VIMG_FLAG_STAGE_TIMES), average duration; {unit_note}"}


def run_scene(gpu, vimg, key, args, hbm_peak, peak_src, want_cpu):
    """One entry of the line's `scenes` map."""
    w = Workload(key, vimg)
    big = w.n_prims > 100000
    streams = 16 if big else args.spp_streams
    out = {"config": w.config_name, "workload": w.label(), "scene": w.build, "prims": w.n_prims, "bvh_nodes": w.n_nodes,
           "n_gpus": gpu.world}
    scene = vimg.Scene(w.desc_ptr, device=gpu.local_rank)
    out["scene_device_bytes"] = scene.device_bytes
    samples = w.width * w.height * w.spp
    kw = mode_kwargs("fast", min(streams, w.spp))
    warm_spp = max(1, min(w.spp, 16 if big else w.spp))
    time_frames(gpu, scene, w.integrator, warm_spp, w.depth, mode_kwargs("fast", min(streams, warm_spp)), 1, 1)
    n_steps = args.scene_steps if args.scene_steps else (1 if samples > 4e9 else 2)
    r = time_frames(gpu, scene, w.integrator, w.spp, w.depth, kw, n_steps, 0)
    out["fast"] = {"mrays_per_s": r["rays"] / r["seconds"] / 1e6, "spp_mpix_per_s": samples * r["steps"] / r["seconds"] / 1e6,
                   "ms_per_step": r["ms_per_step"], "steps": r["steps"], "spp_streams": kw["spp_streams"],
                   "rays_per_sample": r["rays"] / (samples * r["steps"]), "gpu_launches": r["launches"], "nan_samples": r["nan_samples"]}
    gpu_rps = None
    if not (big and w.spp > 512):
        # bit-identical mode (C5 at 1024 spp is left out: one stream per pixel at 4K is minutes of tail)
        re = time_frames(gpu, scene, w.integrator, w.spp, w.depth, mode_kwargs("exact", 1), 1, 0)
        gpu_rps = re["rays"] / samples
        out["bit_identical"] = {"mrays_per_s": re["rays"] / re["seconds"] / 1e6, "spp_mpix_per_s": samples / re["seconds"] / 1e6,
                                "ms_per_step": re["ms_per_step"], "rays_per_sample": gpu_rps}
    for integ in w.also:  # the normal integrators: bit-exact by contract, one mode
        rn = time_frames(gpu, scene, integ, w.spp, w.depth, {}, 2, 1)
        out[integ] = {"mrays_per_s": rn["rays"] / rn["seconds"] / 1e6, "spp_mpix_per_s": samples * 2 / rn["seconds"] / 1e6,
                      "ms_per_step": rn["ms_per_step"]}
    if gpu.rank == 0:
        out["roofline"] = roofline_of(gpu, scene, w, w.integrator, w.spp, kw, hbm_peak, peak_src)
        if gpu_rps is None:
            _, c1 = scene.render_device(w.integrator, 1, w.depth, spp_streams=1)
            gpu_rps = (c1["rays_closest"] + c1["rays_shadow"]) / (w.width * w.height)
    gpu.barrier()
    scene.close()
    if want_cpu and gpu.rank == 0:
        rr = time_reference(w, w.integrator, 1, 0, budget_s=args.scene_cpu_budget, min_spp=1 if big else 16)
        if rr is not None:
            out["cpu_baseline"] = cpu_baseline_of(w, w.integrator, rr, gpu_rps)
            if out["cpu_baseline"]["value"]:
                out["fast"]["vs_cpu_baseline"] = out["fast"]["mrays_per_s"] / out["cpu_baseline"]["value"]
    gpu.barrier()
    w.close()
    vimg.release_workspaces()
    return out


def cuda_line(args):
    vimg = load_product()
    if vimg.device_count() == 0:
        raise SystemExit("bench.py --impl cuda needs a CUDA device (no CPU fallback)")
    guard = StdoutToStderr().__enter__()  # NCCL's banner, the reference's progress prints
    gpu = Gpu()
    torch = gpu.torch
    hbm_peak, peak_src = measured_peaks()

    w = Workload(args.workload, vimg, args.spp)
    integ, spp, depth = w.integrator, w.spp, w.depth
    streams = max(1, args.spp_streams)
    kw = mode_kwargs(args.mode, streams)
    other = "exact" if args.mode == "fast" else "fast"
    kw_other = mode_kwargs(other, streams)
    samples = w.width * w.height * spp

    # ---- the timed region of record: W warm-up + K steps, frame resident, CUDA events, max over ranks
    scene = vimg.Scene(w.desc_ptr, device=gpu.local_rank)
    sampler = ClockSampler(gpu.local_rank) if gpu.rank == 0 else None
    main = time_frames(gpu, scene, integ, spp, depth, kw, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    # ---- the same frame in the other mode (fewer steps: it is a companion number)
    o_steps = max(1, min(args.steps, 3))
    side = time_frames(gpu, scene, integ, spp, depth, kw_other, o_steps, 1)
    scene.close()  # gives its workspace back to the per-device cache: the e2e scenes take it over
    # ---- end to end with host buffers
    e_steps = max(1, min(args.steps, 3))
    time_e2e(gpu, vimg, w, integ, spp, kw, 1)  # untimed: first-touch of the pinned buffer
    e2e = time_e2e(gpu, vimg, w, integ, spp, kw, e_steps)
    e2e_o = time_e2e(gpu, vimg, w, integ, spp, kw_other, max(1, min(e_steps, 2)))

    roof = None
    scene = vimg.Scene(w.desc_ptr, device=gpu.local_rank)
    if gpu.rank == 0:
        roof = roofline_of(gpu, scene, w, integ, spp, kw, hbm_peak, peak_src)
    gpu.barrier()
    scene.close()

    # ---- the other BASELINE configs
    if args.scenes == "default":
        keys = ["C1", "C2", "C4", "C5"] if gpu.world == 1 else ["C5"]
    elif args.scenes == "none":
        keys = []
    else:
        keys = [k for k in args.scenes.split(",") if k in WORKLOADS and k != args.workload]
    scenes = {}
    for k in keys:
        scenes[k] = run_scene(gpu, vimg, k, args, hbm_peak, peak_src, want_cpu=not args.no_cpu)

    if gpu.rank != 0:
        if gpu.world > 1:
            gpu.dist.destroy_process_group()
        return

    def mode_note(m):
        return ("one PCG stream per pixel and the reference's carried-over traversal stack (library default): image "
                "bit-identical to the CPU renderer in every pixel" if m == "exact" else
                f"tolerance mode (VIMG_FLAG_FAST): fast-math build of the kernels, flat-list trace for scenes of <= 32 primitives, "
                f"{streams} sample ranges per pixel with their own PCG streams; image within the stated relMSE tolerance of the CPU "
                "renderer (tests/test_gpu_fast.py)")

    secs = main["seconds"]
    value = main["rays"] / secs / 1e6
    side_value = side["rays"] / side["seconds"] / 1e6
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": gpu.world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "cuda",
        "spp_mpix_per_s": samples * args.steps / secs / 1e6,
        "rays_per_sample": main["rays"] / (samples * args.steps),
        "value_mode": args.mode,
        ("value_bit_identical" if other == "exact" else "value_fast"): side_value,
        "companion": {"mode": other, "mrays_per_s": side_value, "ms_per_step": side["ms_per_step"], "steps": side["steps"],
                      "spp_mpix_per_s": samples * side["steps"] / side["seconds"] / 1e6, "note": mode_note(other)},
        "config": {"workload": w.label(), "baseline_config": w.config_name, "scene": w.build, "prims": w.n_prims,
                   "bvh_nodes": w.n_nodes, "mode": args.mode, "mode_note": mode_note(args.mode), "spp_streams": kw["spp_streams"],
                   "partition": f"8x8 tiles round-robin over {gpu.world} rank(s)"
                                + (" + NCCL reduce(sum) to rank 0" if gpu.world > 1 else ""),
                   "l2": "flushed between timed steps (256 MiB memset); path state (> 300 MB) streams through L2 every launch"},
        "clocks": clocks,
        "e2e": {"value": e2e["rays"] / e2e["seconds"] / 1e6, "unit": "Mrays/s",
                "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": w.width * w.height * 12,
                "ms_per_step": 1e3 * e2e["seconds"] / e2e["steps"], "steps": e2e["steps"],
                "spp_mpix_per_s": samples * e2e["steps"] / e2e["seconds"] / 1e6,
                "path": ("vimg_scene_create (H2D) + vimg_render into pinned host memory (kernels + D2H)" if gpu.world == 1 else
                         "per rank vimg_scene_create (H2D) + vimg_render_device, NCCL reduce of the partial frames to rank 0, "
                         "D2H of the whole frame into rank 0's pinned host memory")},
        ("e2e_bit_identical" if other == "exact" else "e2e_fast"): e2e_o["rays"] / e2e_o["seconds"] / 1e6,
        "gpu_launches": main["launches"],
        "roofline": roof,
        "scenes": scenes,
    }
    # ---- the step after the path: tonemapper + sRGB OETF + 8-bit quantisation on the
    # device.  HBM bound: 12 B read + 3 B written per pixel; timed on a 4K frame
    rng = np.random.default_rng(0)
    if not args.no_extras:
        post = {}
        big = (rng.random((2160, 3840, 3), dtype=np.float32) * 4).astype(np.float32)
        for tm in ("clamp", "aces", "reinhard", "agx"):
            best = min(vimg.postprocess(big, tm, device=gpu.local_rank, want_srgb=False)[2] for _ in range(5))
            nbytes = big.size // 3 * 15 + (big.size // 3 * 12 if tm == "reinhard" else 0)
            post[tm] = {"kernel_ms": best, "achieved_gbs": nbytes / best / 1e6, "frac_of_hbm": nbytes / best / 1e6 / hbm_peak}
        line["postprocess"] = {"frame": "3840x2160 synthetic, bytes = 15 per pixel (+12 for reinhard's luminance pass)",
                               "per_tonemapper": post}
        # ---- the step before the path: glTF texture conversion, mip chain, env-map CDFs
        tex8 = rng.integers(0, 256, size=(2048, 2048, 3), dtype=np.uint8)
        conv = {}
        for name, fn in (("srgb8_to_linear", lambda: vimg.convert_srgb8_to_linear(tex8, device=gpu.local_rank)[1]),
                         ("rgb8_to_normal", lambda: vimg.convert_rgb8_to_normal(tex8, 1.0, device=gpu.local_rank)[1])):
            best = min(fn() for _ in range(5))
            nbytes = tex8.size * 5
            conv[name] = {"kernel_ms": best, "achieved_gbs": nbytes / best / 1e6, "frac_of_hbm": nbytes / best / 1e6 / hbm_peak}
        tex_ms = min(vimg.build_image_texture(tex8, device=gpu.local_rank)[1] for _ in range(3))
        env = (rng.random((1024, 2048, 3), dtype=np.float32) * 2).astype(np.float32)
        cdf_ms = min(vimg.build_envmap_cdf(env, device=gpu.local_rank)[2] for _ in range(3))
        line["scene_prep"] = {"texture": "2048x2048 8-bit RGB (12.6 MB in, 50.3 MB out), env map 2048x1024",
                              "conversions": conv, "image_texture_ms": tex_ms, "envmap_cdf_ms": cdf_ms}
    if not args.no_cpu and gpu.world == 1:
        r = time_reference(w, integ, 1, 0, budget_s=args.cpu_budget)
        if r is not None:
            line["cpu_baseline"] = cpu_baseline_of(w, integ, r)
    guard.__exit__()
    print(json.dumps(line), flush=True)
    if gpu.world > 1:
        gpu.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS), help="headline workload (BASELINE config)")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"],
                    help="mode of `value` / `e2e`: fast = tolerance mode (VIMG_FLAG_FAST + sample-range streams), "
                         "exact = library default, bit-identical images; the other mode is reported beside it")
    ap.add_argument("--spp", type=int, default=0, help="override the scene's spp (invalidates the number)")
    ap.add_argument("--spp-streams", type=int, default=64,
                    help="fast mode: sample ranges per pixel with their own RNG streams (vimg_render_params.spp_streams)")
    ap.add_argument("--scenes", default="default",
                    help="other BASELINE configs measured into `scenes`: default (N = 1: C1,C2,C4,C5; N > 1: C5), none, or a list")
    ap.add_argument("--scene-steps", type=int, default=0, help="timed steps per entry of `scenes` (0: 2, or 1 for frames of > 4 G samples)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extras", action="store_true", help="skip the postprocess / scene_prep timings")
    ap.add_argument("--cpu-budget", type=float, default=14.0, help="seconds of CPU work for the headline cpu_baseline")
    ap.add_argument("--scene-cpu-budget", type=float, default=5.0, help="seconds of CPU work per `scenes` cpu_baseline")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: seconds for all steps together")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            reference_line(args)
        return
    cuda_line(args)


if __name__ == "__main__":
    main()
