"""Average duration per kernel from an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
d = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > 10:
        d[r[4].split('(')[0]].append(float(r[-1].replace(',', '')))
for k, v in d.items():
    print(f"{k:45s} n={len(v):3d}  avg {sum(v) / len(v) / 1e3:9.1f} us")
