#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (markdown)."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
start = next(i for i, r in enumerate(rows) if r[0] == "ID")
hdr, data = rows[start], rows[start + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.OrderedDict(), 0.0
for r in data:
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")[:80]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
mine = sum(t for k, (n, t) in agg.items() if k.startswith(("mvster::", "tc::", "tc2::", "tc3::")))
print(f"{len(data)} launches, {tot:.0f} us of GPU time in one step (serialised, cold-cache: compare shares); "
      f"libmvster_b200 kernels: {mine:.0f} us ({100 * mine / tot:.1f} %)\n")
print("| us | share | launches | kernel |\n|---:|---:|---:|---|")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    if t / tot < 0.002:
        continue
    print(f"| {t:.1f} | {100 * t / tot:.1f} % | {n} | `{k}` |")
