"""profiles/ncu_traffic.json from this round's `ncu --set full` captures: dram__bytes_read.sum + dram__bytes_write.sum of
one launch per kernel, keyed the way bench.py's roofline names kernels, stamped with the hash of the device sources
the captures were taken on (bench.py quotes the figure only for that build).
usage: ncu_traffic.py WORKLOAD=report.ncu-rep [WORKLOAD=report.ncu-rep ...]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
NAMES = {"qw_trace_wide_kernel": "qw_trace_wide_kernel", "qw_trace_kernel": "qw_trace_kernel",
         "qv_shade_kernel<3, 0>": "qv_shade_kernel<mis, Lambertian>", "qv_shade_kernel<3, 1>": "qv_shade_kernel<mis, Principled>",
         "qv_shade_kernel<3, 2>": "qv_shade_kernel<mis, Dielectric>", "qv_end_kernel<3>": "qv_end_kernel<mis>"}
def launch_items(rep):
    """items of each captured launch of a streamed kernel: the predicated-on thread count of the load of its queue
    record (ld.global.cs = LDG.E.EF.128, one per item), from the report's source page"""
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                                                      capture_output=True, text=True).stdout)))
    items, k, hdr = {}, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            k, hdr = r[1].replace("void ", "").split("::", 1)[-1], None
        elif r[0] == "Address":
            hdr = {x: i for i, x in enumerate(r)}
        elif hdr and r[0].startswith("0x") and "LDG.E.EF.128" in r[1] and k not in items:
            items[k] = int(r[hdr["Predicated-On Thread Instructions Executed"]])
    return items


out = {"source_hash": bench.kernel_source_hash(), "capture": [], "kernels": {}}
for arg in sys.argv[1:]:
    wl, rep = arg.split("=", 1)
    out["capture"].append(os.path.basename(rep))
    items = launch_items(rep)
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    h = {n: i for i, n in enumerate(rows[0])}
    units = rows[1]
    for r in rows[2:]:
        name = r[h["Kernel Name"]].replace("void ", "")
        key = next((v for k, v in NAMES.items() if name.startswith(k)), None)
        if key is None:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(r[h[m]].replace(",", ""))
            u = units[h[m]]
            tot += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        e = {"dram_bytes_per_launch": tot, "gpu_time_us": r[h["gpu__time_duration.sum"]] + " " + units[h["gpu__time_duration.sum"]]}
        n_items = next((v for k, v in items.items() if k.replace("(int)", "").replace("(unsigned int)", "").startswith(next(kk for kk in NAMES if name.startswith(kk)))), None)
        if n_items:
            e["items"] = n_items
        out["kernels"][f"{wl}:{key}"] = e
out["capture"] = ", ".join(out["capture"]) + " (gpurun_out/, summaries in profiles/r6_*_kernels*.txt)"
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
