#!/usr/bin/env python3
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py file.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value"); ui = H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    agg.setdefault(r[ki].split("(")[0], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:40s} n={len(v):4d} mean={sum(v) / len(v):9.1f} us  total={sum(v):10.1f} us  share={100 * sum(v) / tot:5.1f} %")
