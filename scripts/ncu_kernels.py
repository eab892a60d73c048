"""One line block per captured launch of an ncu report: time, occupancy, issue, lanes, stalls, DRAM."""
import csv, subprocess, sys, io
rep = sys.argv[1]
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
keys = [("us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("regs", "launch__registers_per_thread"),
        ("warp_inst", "smsp__inst_executed.sum"), ("lanes", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("lat/inst", "smsp__average_warp_latency_per_inst_issued.ratio"), ("dram_rd_MB", "dram__bytes_read.sum"),
        ("dram_wr_MB", "dram__bytes_write.sum"), ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2hit%", "lts__t_sector_hit_rate.pct"),
        ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"), ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")]
for r in rows[2:]:
    print("==", r[ix["Kernel Name"]][:70])
    print("   " + "  ".join(f"{k}={r[ix[h]]}" for k, h in keys if h in ix))
    st = []
    for h, i in ix.items():
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v >= 0.25:
                st.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    print("   stalls: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)))
