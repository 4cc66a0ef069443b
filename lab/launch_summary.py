#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python lab/launch_summary.py file.csv [n_steps]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    d = dict(zip(h, r))
    v = float(d["Metric Value"].replace(",", ""))
    v *= {"ms": 1000.0, "us": 1.0, "ns": 0.001, "s": 1e6}.get(d["Metric Unit"], 1.0)
    k = d["Kernel Name"].split("(")[0]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in agg.values())
print(f"{sum(c for c, _ in agg.values())} launches, {tot/1000:.3f} ms of kernel time ({steps} steps incl. warm-up/setup in the command)")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} {c:6d} launches {v/1000:10.3f} ms {v/tot:6.1%}  avg {v/c:9.1f} us")
