"""Per-source-line view of an ncu report: where warp instructions and stall samples go."""
import csv, subprocess, sys, io, collections, os
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; ci_inst = r.index("Instructions Executed"); ci_thr = r.index("Thread Instructions Executed"); ci_s = r.index("# Samples"); continue
    if hdr and len(r) > ci_thr and r[2] == "-" and r[0].isdigit():
        try: lines.append((cur, int(r[0]), r[1].strip(), int(r[ci_inst] or 0), int(r[ci_thr] or 0), int(r[ci_s] or 0)))
        except ValueError: pass
tot_i = sum(l[3] for l in lines); tot_t = sum(l[4] for l in lines); tot_s = sum(l[5] for l in lines)
print(f"total warp-inst {tot_i:.3e}, thread-inst {tot_t:.3e}, avg lanes {tot_t/max(tot_i,1):.2f}, samples {tot_s}")
byfile = collections.defaultdict(lambda: [0, 0, 0])
for f, ln, src, i, t, s in lines:
    byfile[f][0] += i; byfile[f][1] += t; byfile[f][2] += s
for f, (i, t, s) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:24s} inst {100*i/tot_i:5.1f}%  lanes {t/max(i,1):5.2f}  samples {100*s/tot_s:5.1f}%")
print("top lines by warp instructions:")
for f, ln, src, i, t, s in sorted(lines, key=lambda l: -l[3])[:topn]:
    print(f"  {100*i/tot_i:5.2f}% lanes {t/max(i,1):5.2f} smp {100*s/tot_s:5.2f}%  {f}:{ln}  {src[:90]}")
