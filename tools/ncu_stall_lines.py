"""Top source-page lines of one stall reason in an ncu report: python tools/ncu_stall_lines.py rep.ncu-rep long_sb [n]"""
import csv, io, subprocess, sys, collections
rep, reason = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for i, r in enumerate(rows):
    if "# Samples" in r:
        h, start = r, i
        break
data = [r for r in rows[start + 1:] if len(r) == len(h)]
isamp, isrc, il = h.index("# Samples"), h.index("Source"), h.index("stall_" + reason)
tot = sum(int(r[il] or 0) for r in data)
print("total", reason, tot, "of", sum(int(r[isamp] or 0) for r in data), "samples;", len(data), "instructions")
top = sorted(range(len(data)), key=lambda k: -int(data[k][il] or 0))[:n]
for k in sorted(top):
    print(f"{k:6d} {data[k][isrc][:80]:80s} {reason} {data[k][il]:>5s} samples {data[k][isamp]:>5s}")
b = collections.Counter(); s = collections.Counter()
for k, r in enumerate(data):
    b[k * 20 // len(data)] += int(r[il] or 0); s[k * 20 // len(data)] += int(r[isamp] or 0)
print(reason, "by 5% position bins", [b[i] for i in range(20)])
print("samples by bins          ", [s[i] for i in range(20)])
