"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share, average).
usage: python tools/ncu_launch_summary.py profiles/<launches>.csv > profiles/<launches>_summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v, u = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
    us = v / 1000 if u.startswith("n") else (v if u.startswith("u") else v * 1000)
    a = agg.setdefault(d["Kernel Name"], [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("# serialised, cold-cache launch times under ncu: compare SHARES, not absolutes (peak microbenchmarks of bench.py included)")
for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{name[:110]:110s} n={a[0]:5d} total {a[1] / 1000:9.3f} ms  share {100 * a[1] / tot:5.1f}%  avg {a[1] / a[0]:9.1f} us")
print(f"total {tot / 1000:.3f} ms")
