"""Condenses an ncu report (--set full, --import-source on) into the numbers quoted in DESIGN.md / profiles/:
per kernel: duration, DRAM bytes, issue utilisation, occupancy, registers, and the stall-reason mix from the
source page.  Usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    print("=" * 100)
    print(r[hdr.index("Kernel Name")][:110])
    for w in WANT:
        if w in hdr:
            print(f"  {w:70s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:60], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if not b["rows"]:
        continue
    h = b["rows"][0]
    data = [r for r in b["rows"][1:] if len(r) == len(h)]
    if "# Samples" not in h:
        continue
    isamp, iex, isrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    names = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[isamp] or 0) for r in data)
    print("-" * 100)
    print("stall sampling:", b["name"], "samples", tot, "warp instructions", sum(int(r[iex] or 0) for r in data))
    agg = {n: sum(int(r[h.index(n)] or 0) for r in data) for n in names}
    print("  " + ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    op = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        op[m.group(2).split(".")[0] if m else "?"] += int(r[iex] or 0)
    print("  instruction mix: " + ", ".join(f"{o} {c}" for o, c in op.most_common(10)))
