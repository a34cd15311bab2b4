#!/usr/bin/env python3
"""Summarise gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum): per-kernel totals and the last decode step."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))


def us(r):
    v = float(r["Metric Value"].replace(",", ""))
    return v / 1000 if r["Metric Unit"] == "ns" else (v * 1000 if r["Metric Unit"] == "ms" else v)


names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "") for r in rows]
vals = [us(r) for r in rows]
agg = collections.OrderedDict()
for n, v in zip(names, vals):
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(vals)
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:60]:60s} n={a[0]:4d} total={a[1]:9.1f}us avg={a[1] / a[0]:8.2f}us share={a[1] / tot * 100:5.1f}%")
print("total us", round(tot, 1))
idx = [i for i, n in enumerate(names) if "step_finish" in n]
if len(idx) >= 2:
    s, e = idx[-2] + 1, idx[-1] + 1
    print("--- last decode step ---")
    for i in range(s, e):
        print(f"{names[i][:50]:50s} {vals[i]:8.2f} grid={rows[i]['Grid Size']}")
    print("step sum us", round(sum(vals[s:e]), 1))
