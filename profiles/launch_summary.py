#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file launches.csv ...).
usage: launch_summary.py launches.csv ["command line of the profiled run"] > launches_summary.txt"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or r[ik] == "Kernel Name":
        continue
    try:
        us = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    except ValueError:
        continue
    tot[r[ik]] += us
    cnt[r[ik]] += 1
total = sum(tot.values())
if len(sys.argv) > 2:
    print("ncu --metrics gpu__time_duration.sum --clock-control none ", sys.argv[2])
print("(cold-cache, serialised launches: compare shares, not absolutes)")
for k, us in tot.most_common(24):
    print(f"{k[:70]:70s} n={cnt[k]:4d} total={us:12.1f} us  share={100 * us / total:5.1f}%")
