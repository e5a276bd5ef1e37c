#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("bbocr::", "").replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / tot * 100:6.2f}%  {v[1]:10.1f} us  n={v[0]:4d}  avg={v[1] / v[0]:8.1f} us  {k[:80]}")
