"""Aggregate an ncu --csv launch list (gpu__time_duration.sum [+ dram bytes]) by kernel name and grid."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0, 0.0])
per = {}
for row in r:
    key = (row["ID"])
    d = per.setdefault(key, {"name": row["Kernel Name"], "grid": row.get("Grid Size", "")})
    d[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    d["unit_" + row["Metric Name"]] = row["Metric Unit"]
for d in per.values():
    t = d.get("gpu__time_duration.sum", 0.0)
    u = d.get("unit_gpu__time_duration.sum", "ns")
    t_us = t / 1e3 if u in ("ns", "nsecond") else (t if u in ("us", "usecond") else t * 1e3)
    by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    k = (d["name"].split("(")[0][-60:], d["grid"])
    agg[k][0] += 1
    agg[k][1] += t_us
    agg[k][2] += by
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':62s} {'grid':>18s} {'n':>6s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s} {'dram_MB':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k[0]:62s} {k[1]:>18s} {v[0]:6d} {v[1]:10.1f} {v[1] / v[0]:8.1f} {100 * v[1] / tot:5.1f}% {v[2] / 1e6:9.1f}")
print("total_us", tot)
