"""Summary of an ncu capture of a whole frame exported as CSV on the GPU box (the .ncu-rep of 35 kernels with sources is
too large to bring back): python tools/ncu_frame_summary.py raw.csv src.csv.gz > profiles/xxx.txt
raw.csv = `ncu -i rep --page raw --csv`, src.csv.gz = `ncu -i rep --page source --csv | gzip`."""
import collections, csv, gzip, sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr = rows[0]


def g(d, k):
    try:
        return float(d[k].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")


print(f"{'kernel':28s} {'us':>8s} {'issue%':>6s} {'occ%':>5s} {'l1tex%':>6s} {'lsu wavefronts%':>15s} {'dram%':>6s} {'Minst':>7s} {'regs':>4s} grid x block")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    us = g(d, "gpu__time_duration.sum")
    us = us * 1000 if us < 10 else us          # (ms in some exports)
    print(f"{d['Kernel Name'].split('(')[0][:28]:28s} {us:8.1f} {g(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} {g(d, 'l1tex__throughput.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g(d, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):15.1f} "
          f"{g(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {g(d, 'smsp__inst_executed.sum') / 1e6:7.2f} "
          f"{d.get('launch__registers_per_thread', ''):>4s} {d['launch__grid_size']} x {d['launch__block_size']}")
print()
srows = list(csv.reader(gzip.open(src, "rt") if src.endswith(".gz") else open(src)))
starts = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"] + [len(srows)]
seen = set()
for k, st in enumerate(starts[:-1]):
    name = srows[st][1]
    h = srows[st + 1]
    data = [r for r in srows[st + 2:starts[k + 1]] if len(r) == len(h)]
    ie = h.index("Instructions Executed")
    tot = sum(int(r[ie] or 0) for r in data)
    key = (name, tot)
    if key in seen:
        continue
    seen.add(key)
    mix = collections.Counter()
    mem = collections.defaultdict(lambda: [0, 0, 0, 0])
    has_mem = "L1 Wavefronts Shared" in h
    iw, ii, it = (h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal"), h.index("L1 Tag Requests Global")) if has_mem else (0, 0, 0)
    for r in data:
        op = [o for o in r[1].split() if not o.startswith("@")][0]
        n = int(r[ie] or 0)
        mix[op.split(".")[0]] += n
        if has_mem and op.split(".")[0] in ("LDS", "STS", "LDG", "STG", "LDGSTS", "LDL", "STL"):
            a = mem[op]
            a[0] += n; a[1] += int(r[iw] or 0); a[2] += int(r[ii] or 0); a[3] += int(r[it] or 0)
    cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    stall = {c: sum(int(r[h.index(c)] or 0) for r in data) for c in cols}
    ts = sum(stall.values()) or 1
    print(f"{name[:100]}\n   SASS instructions {len(data)}, executed {tot / 1e6:.2f} M warp instructions")
    print("   mix:    " + ", ".join(f"{o} {n / max(tot, 1) * 100:.1f}%" for o, n in mix.most_common(12)))
    print("   stalls: " + ", ".join(f"{c[6:]} {v / ts * 100:.1f}%" for c, v in sorted(stall.items(), key=lambda x: -x[1])[:8]))
    for op, a in sorted(mem.items(), key=lambda x: -(x[1][1] + x[1][3]))[:6]:
        print(f"   {op:24s} executed {a[0]:9d}  shared wavefronts {a[1]:9d} (ideal {a[2]:9d})  global tag requests {a[3]:9d}")
