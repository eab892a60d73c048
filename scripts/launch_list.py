"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and share per kernel.
usage: launch_list.py launches.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
mu = h.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("vimg_fast::", "").replace("vimg::", "")
    v = float(r[mv].replace(",", ""))
    unit = r[mu]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    tot[name] += us; cnt[name] += 1
total = sum(tot.values())
print(f"# total {total / 1e3:.2f} ms in {sum(cnt.values())} launches")
print("kernel, launches, total_us, avg_us, share")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k}, {cnt[k]}, {v:.1f}, {v / cnt[k]:.2f}, {v / total:.4f}")
