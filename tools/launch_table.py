#!/usr/bin/env python
"""Markdown table (kernel, launches, total ms, share) from an ncu --csv launch list taken with
--metrics gpu__time_duration.sum.   usage: launch_table.py <launches.csv>"""
import csv
import re
import sys
from collections import OrderedDict

rows = OrderedDict()
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("ngsq::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
    n, t = rows.get(name, (0, 0.0))
    rows[name] = (n + 1, t + ms)
total = sum(t for _, t in rows.values())
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, (n, t) in rows.items():
    print(f"| `{k}` | {n} | {t:.3f} | {100 * t / total:.1f} % |")
print(f"| all | {sum(n for n, _ in rows.values())} | {total:.3f} | 100 % |")
