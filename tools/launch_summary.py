"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
agg = collections.defaultdict(list)
for r in rows[1:]:
    d = dict(zip(hdr, r))
    if d.get("Metric Name") == "gpu__time_duration.sum":
        v = float(d["Metric Value"].replace(",", ""))
        agg[d["Kernel Name"].split("(")[0]].append(v * (1e-3 if d["Metric Unit"] == "ns" else 1))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:60]:60s} n={len(v):3d} mean_us={sum(v)/len(v):10.2f} share={100*sum(v)/tot:5.1f}%")
