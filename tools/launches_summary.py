#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launches_summary.py launches.csv "<command line that was profiled>" """
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iN, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}[r[iU]]
    name = re.sub(r"\(.*", "", r[iN]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    tot[name] += v
    cnt[name] += 1
step = sum(v for k, v in tot.items() if "generate" not in k)
print(f"ncu --metrics gpu__time_duration.sum --clock-control none: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("(cold-cache, serialised launches: compare shares; shares exclude the one-off data generator)")
print(f"{'kernel':60s} {'launches':>8s} {'total_ms':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    share = "-" if "generate" in k else f"{v / step:.3f}"
    print(f"{k:60s} {cnt[k]:8d} {v:10.3f} {share:>7s}")
