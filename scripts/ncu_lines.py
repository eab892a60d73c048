"""Per-source-line view of an ncu report: where warp instructions and stall samples go.
usage: ncu_lines.py report.ncu-rep [topn] [kernel-substring]   (kernel-substring selects by demangled name)"""
import csv, subprocess, sys, io, collections, os
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ksel = sys.argv[3] if len(sys.argv) > 3 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; fn = None; per = collections.defaultdict(list)
for r in rows:
    if not r: continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; ci_inst = r.index("Instructions Executed"); ci_thr = r.index("Thread Instructions Executed"); ci_s = r.index("# Samples"); continue
    if hdr and len(r) > ci_thr and r[2] == "-" and r[0].isdigit():
        try: per[fn].append((cur, int(r[0]), r[1].strip(), int(r[ci_inst] or 0), int(r[ci_thr] or 0), int(r[ci_s] or 0)))
        except ValueError: pass
for fn, lines in per.items():
    if ksel and ksel not in fn: continue
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for f, ln, src, i, t, s in lines:
        a = agg[(f, ln, src)]; a[0] += i; a[1] += t; a[2] += s
    tot_i = sum(a[0] for a in agg.values()); tot_t = sum(a[1] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
    print(f"==== {fn[:110]}\ntotal warp-inst {tot_i:.3e}, thread-inst {tot_t:.3e}, avg lanes {tot_t/max(tot_i,1):.2f}, samples {tot_s}")
    byfile = collections.defaultdict(lambda: [0, 0, 0])
    for (f, ln, src), (i, t, s) in agg.items():
        byfile[f][0] += i; byfile[f][1] += t; byfile[f][2] += s
    for f, (i, t, s) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:28s} inst {100*i/max(tot_i,1):5.1f}%  lanes {t/max(i,1):5.2f}  samples {100*s/max(tot_s,1):5.1f}%")
    key = (lambda kv: -kv[1][2]) if os.environ.get("BY", "inst") == "samples" else (lambda kv: -kv[1][0])
    print("top lines by " + os.environ.get("BY", "inst") + ":")
    for (f, ln, src), (i, t, s) in sorted(agg.items(), key=key)[:topn]:
        print(f"  inst {100*i/max(tot_i,1):5.2f}% lanes {t/max(i,1):5.2f} smp {100*s/max(tot_s,1):5.2f}%  {f}:{ln}  {src[:90]}")
