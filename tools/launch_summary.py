#!/usr/bin/env python3
"""Summarise an ncu gpu__time_duration launch list: per-kernel totals of the LAST prove step."""
import collections
import csv
import sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "k_groth16_combine"
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
# one prove step = the launches between the last two occurrences of the marker kernel (the step's final launch)
ends = [i for i, n in enumerate(names) if marker in n]
lo = ends[-2] + 1 if len(ends) >= 2 else 0
rows = rows[:ends[-1] + 1] if ends else rows
agg = collections.OrderedDict()
tot = 0.0
for r in rows[lo:]:
    k = r["Kernel Name"].split("(")[0].replace("void b200::", "").replace("b200::", "")[:60]
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"{'kernel':62s} {'n':>4s} {'us':>10s} {'share':>6s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} {n:4d} {v:10.1f} {100*v/tot:5.1f}%")
print(f"{'TOTAL':62s} {sum(a[0] for a in agg.values()):4d} {tot:10.1f}")
