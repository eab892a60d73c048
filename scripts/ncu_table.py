"""One line per profiled launch of an ncu report: the counters the design notes quote."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__average_warp_latency_per_inst_issued.ratio", "lat/inst"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "wait"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "no_inst"),
        ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%")]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    name = r[col["Kernel Name"]][:60]
    print("== " + name)
    print("   " + "  ".join(f"{lab}={r[col[m]]}" for m, lab in want if m in col))
