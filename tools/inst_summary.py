"""Per-kernel time and warp-instruction totals from `ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --csv`."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
t = collections.defaultdict(float); n = collections.defaultdict(int); ins = collections.defaultdict(float)
for r in rows[1:]:
    d = dict(zip(hdr, r))
    k = d["Kernel Name"].split("(")[0]
    v = float(d["Metric Value"].replace(",", ""))
    if d["Metric Name"] == "gpu__time_duration.sum":
        t[k] += v * (1e-3 if d["Metric Unit"] == "ns" else 1); n[k] += 1
    elif d["Metric Name"] == "smsp__inst_executed.sum":
        ins[k] += v
tot_t, tot_i = sum(t.values()), sum(ins.values())
print(f"total time {tot_t:.1f} us, total warp instructions {tot_i/1e6:.1f} M")
for k in sorted(t, key=lambda k: -ins[k]):
    print(f"{k[:44]:44s} n={n[k]:2d} time_us={t[k]:9.1f} Minst={ins[k]/1e6:8.2f} ({100*ins[k]/tot_i:4.1f}%) ipc/SM={ins[k]/(t[k]*1e-6)/148/1.9e9 if t[k] else 0:5.2f}")
