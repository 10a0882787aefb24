#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals and shares.
usage: tools/summarise_launches.py gpurun_out/launches_<tag>.csv "<command that was profiled>" > profiles/<tag>_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot, cnt = defaultdict(float), defaultdict(int)
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    m = re.search(r"(\w+)\((?:b200lz4::)?\w*Args\)|::(\w+_kernel\w*)", name)
    short = (m.group(1) or m.group(2)) if m and "at::" not in name else "torch:" + name[:40]
    tot[short] += float(r["Metric Value"].replace(",", "")) / 1e6
    cnt[short] += 1
total = sum(tot.values())
print(f"ncu --metrics gpu__time_duration.sum --clock-control none, {sys.argv[2] if len(sys.argv) > 2 else ''} (cold-cache, serialised: compare SHARES)")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k[:46]:<46} launches {cnt[k]:4d}  total {tot[k]:9.3f} ms  mean {tot[k] / cnt[k]:8.4f} ms  share {100 * tot[k] / total:5.1f}%")
