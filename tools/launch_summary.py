#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total time and share per kernel.
usage: launch_summary.py launches.csv > profiles/xxx_summary.txt"""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ci["Metric Value"]].replace(",", ""))
    u = r[ci["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    tot[r[ci["Kernel Name"]]] += v
    cnt[r[ci["Kernel Name"]]] += 1
s = sum(tot.values())
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-78s launches %5d  total %10.1f us  mean %8.1f us  share %5.1f%%" % (k[:78], cnt[k], tot[k], tot[k] / cnt[k], 100 * tot[k] / s))
