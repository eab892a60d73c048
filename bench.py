#!/usr/bin/env python
"""bench.py -- throughput of the v-img render hot path on B200 (and the reference beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference]
                    [--workload disney_spheres] [--spp S]

One "step" = one pass of the hot path over one frame: every pixel of the workload's scene,
all its samples, through the integrator the scene file names (the reference's
scene_integrator<F>() call, include/integrators.h:36-153).

Workload (BASELINE.json): configs[2] `scenes/disney_spheres.json` -- 1800x800, 512 spp, `mis`,
"depth": -1, sweep-SAH BVH -- the scene the north star's ">= 100x" target and the README's only
published number are quoted on.  The scene comes from tests/golden/scenes/*.vscene (the
reference-loaded scene + host-built BVH, flattened by the adapter); /root/reference is not read.

metric `Mrays/s`: a ray is one BVH::hit<closest> or BVH::occlude call; rays are
counted by the kernel itself.  `spp_mpix_per_s` (W*H*spp / s / 1e6) is reported beside it.
With --gpus N (torchrun, one rank per GPU) the frame's 8x8 tiles are dealt round-robin to the
ranks and the partial frames are summed with an NCCL reduce to rank 0 inside the timed region
(strong scaling: the frame is fixed).

`--impl reference` times the compiled, unmodified reference (oracle/_ref, perf flavour = what
its CMake builds: -O3, AVX2 sibling-pair path) on the host cores, on a bounded sample of the
same workload (fewer spp; throughput is spp-invariant).  The same measurement is attached to
the cuda arm's line as `cpu_baseline`.
"""
import argparse
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
from helpers import golden_scene_path, load_product, manifest, rel_mse  # noqa: E402

INTEG_NAMES = ["s_normal", "g_normal", "material", "mis"]
# algorithmic cost model of SURVEY.md section 8(d), in the REFERENCE's data layout
BYTES = {"inner_nodes": 56, "leaf_nodes": 8, "tri_tests": 56, "sphere_tests": 24}
FLOPS = {"inner_nodes": 50, "tri_tests": 67, "sphere_tests": 40}
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from an `ncu --set full` capture
# (profiles/r4_trace_kernel_fresh_stack.txt: a fully loaded iteration of disney_spheres, 2 M slots, this
# round's final kernel; mostly path state and queues -- the scene itself is L1 resident)
NCU_TRAFFIC = {"qw_trace_kernel": 243.432448e6 + 45.497088e6}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


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


def workload_info(name, spp_override):
    e = manifest()[name]
    spp = spp_override or e["spp"]
    return {"name": name, "scene": golden_scene_path(name), "width": e["width"], "height": e["height"],
            "spp": spp, "depth": e["depth"], "integrator": INTEG_NAMES[e["integrator"]],
            "n_prims": e["n_prims"], "n_nodes": e["n_nodes"]}


def config_of(w, n_gpus, extra=None):
    c = {"workload": f"{w['name']} {w['width']}x{w['height']} {w['spp']}spp {w['integrator']} "
                     f"depth={'-1 (RR only)' if w['depth'] == 0xFFFFFFFF else w['depth']}",
         "scene": os.path.relpath(w["scene"], ROOT), "prims": w["n_prims"], "bvh_nodes": w["n_nodes"],
         "partition": f"8x8 tiles round-robin over {n_gpus} rank(s)" + (" + NCCL reduce(sum) to rank 0" if n_gpus > 1 else ""),
         "l2": "flushed between timed steps (256 MiB memset); scene itself is < 8 KiB"}
    if extra:
        c.update(extra)
    return c


class StdoutToStderr:
    """The reference's scene_integrator prints its progress with std::cout (integrators.h:48,142);
    bench.py's stdout carries exactly one JSON line, so fd 1 points at stderr while it runs."""

    def __enter__(self):
        import ctypes
        self.libc = ctypes.CDLL(None)
        sys.stdout.flush()
        self.libc.fflush(None)
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        self.libc.fflush(None)  # the reference's text sits in C stdio's buffer when fd 1 is a file
        os.dup2(self.saved, 1)
        os.close(self.saved)


# --------------------------------------------------------------------------- reference arm
def time_reference(w, steps, warmup, budget_s=25.0):
    """Times the compiled reference (perf flavour) on the host cores.  Returns a dict."""
    with StdoutToStderr():
        return _time_reference(w, steps, warmup, budget_s)


def _time_reference(w, steps, warmup, budget_s):
    from oracle import refcpu
    vimg = load_product()  # only for vimg_scene_desc_load (host-side file parsing, no GPU work)
    if not refcpu.available("perf"):
        return None
    d = vimg.LoadedDesc(w["scene"])
    cores = refcpu.hardware_threads("perf")
    rs = refcpu.RefScene.from_desc(d.ptr, 1, w["depth"], w["integrator"], flavour="perf")
    # bounded sample: pick spp so that one step takes ~budget/(steps+warmup) seconds
    _, t1 = rs.render(w["integrator"], spp=1, threads=cores)  # also warms the thread pool / caches
    _, t2 = rs.render(w["integrator"], spp=2, threads=cores)
    per_spp = max(t2 - t1, t2 / 2, 1e-3)
    spp = int(max(1, min(w["spp"], budget_s / (steps + warmup) / per_spp)))
    for _ in range(warmup):
        rs.render(w["integrator"], spp=spp, threads=cores)
    t_total = 0.0
    for _ in range(steps):
        _, sec = rs.render(w["integrator"], spp=spp, threads=cores)
        t_total += sec
    samples = w["width"] * w["height"] * spp * steps
    # rays per sample from the instrumented (count) flavour of the same reference, 2 spp
    rays_per_sample = None
    if refcpu.available("count"):
        rc = refcpu.RefScene.from_desc(d.ptr, 2, w["depth"], w["integrator"], flavour="count")
        rc.counters(reset=True)
        rc.render(w["integrator"], spp=2, threads=cores)
        c = rc.counters()
        rays_per_sample = (c["rays_closest"] + c["rays_shadow"]) / (w["width"] * w["height"] * 2)
    return {"seconds": t_total, "steps": steps, "spp": spp, "samples": samples, "cores": cores,
            "spp_mpix_per_s": samples / t_total / 1e6, "rays_per_sample": rays_per_sample,
            "mrays_per_s": (samples * rays_per_sample / t_total / 1e6) if rays_per_sample else None}


def reference_line(w, args):
    r = time_reference(w, args.steps, args.warmup)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libvimg_ref_perf.so not built"}))
        return
    sample = (f"{w['name']} full frame at {r['spp']} spp (of {w['spp']}), {r['steps']} step(s); "
              f"rays/sample {r['rays_per_sample']:.3f} from the count flavour at 2 spp")
    line = {"metric": "Mrays/s", "value": r["mrays_per_s"], "unit": "Mrays/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r["seconds"] / r["steps"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "spp_mpix_per_s": r["spp_mpix_per_s"],
            "config": config_of(w, 1, {"partition": f"reference scene_integrator, {r['cores']} host threads",
                                       "l2": "n/a (CPU)"}),
            "cpu_baseline": {"value": r["mrays_per_s"], "unit": "Mrays/s", "cores": r["cores"],
                             "kind": "reference", "sample": sample,
                             "spp_mpix_per_s": r["spp_mpix_per_s"]},
            "e2e": {"value": r["mrays_per_s"], "unit": "Mrays/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------- cuda arm
def cuda_line(w, args):
    import torch
    import torch.distributed as dist

    vimg = load_product()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if vimg.device_count() == 0:
        raise SystemExit("bench.py --impl cuda needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    guard = None
    if world > 1:
        # NCCL writes its version banner to fd 1 when the first communicator is created; stdout carries
        # exactly one JSON line, so fd 1 points at stderr until that line is printed
        guard = StdoutToStderr()
        guard.__enter__()
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    desc = vimg.LoadedDesc(w["scene"])
    scene = vimg.Scene(desc.ptr, device=local_rank)
    integ, spp, depth = w["integrator"], w["spp"], w["depth"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    streams = 1 if args.pipeline == "block" else max(1, args.spp_streams)

    def step():
        """kernel(s) on this rank's tiles (+ NCCL reduce to rank 0); frame stays in HBM"""
        ptr, st = scene.render_device(integ, spp, depth, rank=rank, n_ranks=world, pipeline=args.pipeline,
                                      spp_streams=streams)
        if world > 1:
            dist.reduce(scene.frame_tensor(ptr), dst=0, op=dist.ReduceOp.SUM)
        return ptr, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
        flush.zero_()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = kernel_ms = 0.0
    launches = 0
    ev0.record()
    for i in range(args.steps):
        ptr, st = step()
        rays += st["rays_closest"] + st["rays_shadow"]
        kernel_ms += st["kernel_ms"]
        launches += st["kernel_launches"]
        if i + 1 < args.steps:
            flush.zero_()
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms, rays, launches, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        elapsed_ms, kernel_ms_max = float(tmax[0]), float(tmax[3])
        rays, launches = float(t[1]), int(t[2])
    else:
        kernel_ms_max = kernel_ms

    # ---- end-to-end through the public C-ABI call with HOST buffers (rank-local tiles) ----------
    # every step: scene upload H2D (vimg_scene_create), render, frame D2H into pinned memory
    # the resident scene gives its render workspace back to the per-device cache, so the scenes the
    # loop creates take it over instead of allocating ~1 GB inside the first timed step
    scene.close()
    host = torch.empty((w["height"], w["width"], 3), dtype=torch.float32).pin_memory()
    host_np = host.numpy()
    e2e_steps = max(1, min(args.steps, 2))
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    e2e_step_ms = []
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        sc2 = vimg.Scene(desc.ptr, device=local_rank)
        _, st2 = sc2.scene_integrator(integ, spp, depth, rank=rank, n_ranks=world, out=host_np,
                                      pipeline=args.pipeline, spp_streams=streams)
        h2d = sc2.device_bytes
        sc2.close()
        e2e_rays += st2["rays_closest"] + st2["rays_shadow"]
        e2e_step_ms.append((1e3 * (time.perf_counter() - ts), st2["kernel_ms"], st2["total_ms"]))
    t_loop = time.perf_counter() - t0
    barrier()
    e2e_s = time.perf_counter() - t0
    print(f"[bench e2e] rank {rank}: steps (wall, kernel, kernel + D2H) {[tuple(round(v, 1) for v in m) for m in e2e_step_ms]} ms, loop {1e3 * t_loop:.1f} ms, "
          f"with barrier {1e3 * e2e_s:.1f} ms", file=sys.stderr)
    te = torch.tensor([e2e_s, e2e_rays], dtype=torch.float64, device=dev)
    if world > 1:
        tem = te.clone()
        dist.all_reduce(tem, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.SUM)
        e2e_s, e2e_rays = float(tem[0]), float(te[1])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    scene = vimg.Scene(desc.ptr, device=local_rank)
    # ---- roofline of the dominant kernel -----------------------------------------------------
    # counted pass (untimed): per-ray algorithmic cost, in the reference's layout
    c_spp = min(spp, 16)
    _, cst = scene.render_device(integ, c_spp, depth, rank=0, n_ranks=1, counters=True, pipeline=args.pipeline,
                                 spp_streams=min(streams, c_spp))
    c_rays = cst["rays_closest"] + cst["rays_shadow"]
    bytes_per_ray = sum(BYTES[k] * cst[k] for k in BYTES) / c_rays
    flops_per_ray = sum(FLOPS[k] * cst[k] for k in FLOPS) / c_rays
    hbm_peak, peak_src = measured_peaks()
    samples_per_step = w["width"] * w["height"] * spp
    secs = elapsed_ms / 1e3
    stages = None
    if args.pipeline != "block" and integ in ("material", "mis"):
        # queue wavefront: the TRACE kernel (BVH::hit / BVH::occlude of every queued ray) is the
        # dominant kernel.  Its launches are timed with one CUDA event after every launch of a
        # serial pass over the same frame (VIMG_FLAG_STAGE_TIMES): in the timed region above the
        # shade / end launches of a pool overlap on side streams, so per-kernel events there would
        # measure the overlap, not the kernel.
        t_spp = min(spp, 64)
        flush.zero_()
        _, tst = scene.render_device(integ, t_spp, depth, rank=0, n_ranks=1, pipeline="queue", stage_times=True,
                                     spp_streams=min(streams, t_spp))
        t_rays = tst["rays_closest"] + tst["rays_shadow"]
        n_launch = max(tst["stage_launches"]["trace"], 1)
        launch_s = tst["stage_ms"]["trace"] / 1e3 / n_launch
        rays_per_launch_rank = t_rays / n_launch
        kernel_name = "qw_trace_kernel"
        stage_total = sum(tst["stage_ms"].values())
        stages = {"spp": t_spp, "ms": tst["stage_ms"], "launches": tst["stage_launches"],
                  "trace_share_of_step": tst["stage_ms"]["trace"] / stage_total,
                  "note": "serial pass, one CUDA event per launch (VIMG_FLAG_STAGE_TIMES)"}
        timing_note = (f"trace launches timed in a serial stage-timed pass at {t_spp} spp "
                       f"({n_launch} launches, average duration)")
    else:
        rays_per_launch_rank = rays / max(launches, 1)
        launch_s = kernel_ms_max / 1e3 / max(args.steps, 1)  # one launch per step per rank
        kernel_name = f"wavefront_kernel<{integ}>"
        timing_note = "one launch per step, CUDA events around it in the timed region"
    achieved = bytes_per_ray * rays_per_launch_rank / launch_s / 1e9

    line = {
        "metric": "Mrays/s", "value": rays / secs / 1e6, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "cuda",
        "spp_mpix_per_s": samples_per_step * args.steps / secs / 1e6,
        "rays_per_sample": rays / (samples_per_step * args.steps),
        "config": config_of(w, world, {
            "pipeline": args.pipeline or "queue (library default)",
            "spp_streams": streams,
            "rng": ("the reference's one PCG stream per pixel: image bit-identical to the CPU renderer"
                    if streams == 1 else
                    f"{streams} sample ranges per pixel, each with its own PCG stream (initseq = range), "
                    "reference R2 jitter index: image within the stated relMSE tolerance of the CPU "
                    "renderer (tests/test_gpu_streams.py); --spp-streams 1 gives the bit-identical image")}),
        "clocks": clocks,
        "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": w["width"] * w["height"] * 12,
                "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
                "spp_mpix_per_s": samples_per_step * e2e_steps / e2e_s / 1e6,
                "path": "vimg_scene_create (H2D) + vimg_render into pinned host memory (kernel + D2H)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": NCU_TRAFFIC.get(kernel_name),
                     "peak_source": peak_src,
                     "kernel": kernel_name, "launch_ms": 1e3 * launch_s,
                     "algorithmic_bytes_per_ray": bytes_per_ray,
                     "algorithmic_flops_per_ray": flops_per_ray,
                     "rays_per_launch": rays_per_launch_rank,
                     "fp32_tflops_achieved": flops_per_ray * rays_per_launch_rank / launch_s / 1e12,
                     "note": "bytes in the reference's layout (56/inner, 8/leaf, 56/tri test, 24/sphere "
                             "test, SURVEY.md section 8(d)); per-ray averages from a counted pass at "
                             f"{c_spp} spp; {timing_note}; traffic = dram bytes of one fully loaded "
                             "launch under ncu --set full (profiles/), mostly path state and queues: "
                             "this scene's BVH (< 8 KiB) is L1 resident"},
    }
    if stages:
        line["stages"] = stages
    # ---- the step after the path: tonemapper + sRGB OETF + 8-bit quantisation
    # on the device.  HBM bound: 12 B read + 3 B written per pixel; timed on a 4K frame (the 1800x800
    # frame of this workload is 22 MB: one launch of ~10 us, launch-latency bound)
    post = {}
    rng = np.random.default_rng(0)
    big = (rng.random((2160, 3840, 3), dtype=np.float32) * 4).astype(np.float32)
    for tm in ("clamp", "aces", "reinhard", "agx"):
        best = min(vimg.postprocess(big, tm, device=local_rank, want_srgb=False)[2] for _ in range(5))
        nbytes = big.size // 3 * 15 + (big.size // 3 * 12 if tm == "reinhard" else 0)
        post[tm] = {"kernel_ms": best, "achieved_gbs": nbytes / best / 1e6, "frac_of_hbm": nbytes / best / 1e6 / hbm_peak}
    line["postprocess"] = {"frame": "3840x2160 synthetic, bytes = 15 per pixel (+12 for reinhard's luminance pass)",
                           "per_tonemapper": post}
    # ---- the step before the path: glTF texture conversion, mip chain, env-map CDFs
    # on the device, kernel time only (CUDA events inside the calls).  The conversions are HBM bound:
    # 3 B read + 12 B written per pixel.
    if rank == 0:
        tex8 = rng.integers(0, 256, size=(2048, 2048, 3), dtype=np.uint8)
        conv = {}
        for name, fn in (("srgb8_to_linear", lambda: vimg.convert_srgb8_to_linear(tex8, device=local_rank)[1]),
                         ("rgb8_to_normal", lambda: vimg.convert_rgb8_to_normal(tex8, 1.0, device=local_rank)[1])):
            best = min(fn() for _ in range(5))
            nbytes = tex8.size * 5
            conv[name] = {"kernel_ms": best, "achieved_gbs": nbytes / best / 1e6, "frac_of_hbm": nbytes / best / 1e6 / hbm_peak}
        tex_ms = min(vimg.build_image_texture(tex8, device=local_rank)[1] for _ in range(3))
        env = (rng.random((1024, 2048, 3), dtype=np.float32) * 2).astype(np.float32)
        cdf_ms = min(vimg.build_envmap_cdf(env, device=local_rank)[2] for _ in range(3))
        line["scene_prep"] = {"texture": "2048x2048 8-bit RGB (12.6 MB in, 50.3 MB out), env map 2048x1024",
                              "conversions": conv,
                              "image_texture_ms": tex_ms, "envmap_cdf_ms": cdf_ms,
                              "note": "image_texture = sRGB8 -> linear + the 10 further mip levels; the CDFs are "
                                      "sequential float sums per row (one thread per row, the reference's order)"}
    if not args.no_cpu and world == 1:
        r = time_reference(w, 1, 1, budget_s=16.0)
        if r is not None:
            line["cpu_baseline"] = {
                "value": r["mrays_per_s"], "unit": "Mrays/s", "cores": r["cores"], "kind": "reference",
                "spp_mpix_per_s": r["spp_mpix_per_s"],
                "sample": f"{w['name']} full frame at {r['spp']} spp (of {w['spp']}), compiled reference "
                          f"(perf flavour, AVX2), rays/sample {r['rays_per_sample']:.3f} (count flavour, 2 spp)"}
    if guard is not None:
        guard.__exit__()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="disney_spheres")
    ap.add_argument("--pipeline", default=None, choices=["queue", "block"],
                    help="material / mis pipeline (default: the library's, the queue wavefront)")
    ap.add_argument("--spp", type=int, default=0, help="override the scene's spp (invalidates the number)")
    ap.add_argument("--spp-streams", type=int, default=64,
                    help="sample ranges per pixel with their own RNG streams (vimg_render_params.spp_streams); "
                         "1 = the reference's single stream per pixel, bit-identical image")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    w = workload_info(args.workload, args.spp)
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            reference_line(w, args)
        return
    cuda_line(w, args)


if __name__ == "__main__":
    main()
