"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
    python profiles/summarize_launches.py gpurun_out/r1_launches.csv > profiles/r1_launches_summary.csv"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui].startswith("us") else v * 1e3)
    name = re.sub(r"^void ", "", r[ki]).split("(")[0].replace("loner::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "avg_us", "total_us", "share_of_captured_time"])
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, n, round(t / n, 1), round(t, 1), round(t / tot, 4)])
