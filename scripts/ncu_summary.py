"""Summarise an ncu report: headline metrics + hottest source lines (needs -lineinfo builds)."""
import csv, subprocess, sys, collections, re, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
for i, h in enumerate(hdr):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        v = vals[i]
        try:
            if float(v.replace(",", "")) == 0 and "stalled" in h: continue
        except ValueError:
            pass
        print(f"{h} = {v} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda" if False else "sass"], capture_output=True, text=True).stdout
# per-opcode stall / instruction histogram
lines = list(csv.reader(io.StringIO(src)))
h = lines[1]
ci = {n: i for i, n in enumerate(h)}
ops = collections.Counter(); opsamp = collections.Counter(); noinst = collections.Counter()
tot_inst = 0; tot_samples = 0
for r in lines[2:]:
    if len(r) < len(h): continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ci["Source"]])
    op = m.group(2) if m else "?"
    try:
        n = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0)
    except ValueError:  # a report with several launches repeats the header
        continue
    ops[op] += n; opsamp[op] += s; tot_inst += n; tot_samples += s
print(f"static SASS instructions {len(lines)-2}, executed warp-instructions {tot_inst}, samples {tot_samples}")
print("top opcodes by executed count: " + ", ".join(f"{o}:{100*c/tot_inst:.1f}%" for o, c in ops.most_common(14)))
print("top opcodes by stall samples:  " + ", ".join(f"{o}:{100*c/tot_samples:.1f}%" for o, c in opsamp.most_common(14)))
