#!/usr/bin/env python
"""Aggregate the warp-stall samples of an `ncu --set full` report (source page) by stall reason and by code region.

    python tools/ncu_stalls.py <report.ncu-rep> [top_n]
"""
import csv, io, subprocess, sys
from collections import Counter

def main(report, top=25):
    raw = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]; data = rows[hi + 1:]
    col = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = Counter(); total = 0
    byop = Counter(); execop = Counter()
    for r in data:
        if len(r) < len(h): continue
        s = int(r[col["# Samples"]] or 0); total += s
        op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
        if op.startswith("@"): op = r[col["Source"]].split()[1]
        byop[op] += s; execop[op] += int(r[col["Instructions Executed"]] or 0)
        for n in stalls:
            tot[n] += int(r[col[n]] or 0)
    print("total samples", total, "instructions", len(data))
    for n, v in tot.most_common(12): print(f"  {n:28s} {v:9d} {100.0*v/max(total,1):5.1f}%")
    print("by opcode (samples, executed warp-instructions):")
    for n, v in byop.most_common(int(top)): print(f"  {n:22s} {v:9d} {100.0*v/max(total,1):5.1f}%   exec {execop[n]:12d}")
    print("hottest instructions:")
    hot = sorted((r for r in data if len(r) >= len(h)), key=lambda r: -int(r[col["# Samples"]] or 0))[:int(top)]
    for r in hot:
        top_stall = max(stalls, key=lambda n: int(r[col[n]] or 0))
        print(f"  {r[col['# Samples']]:>7s} {top_stall:18s} {r[col['Source']].strip()[:90]}")

if __name__ == "__main__":
    main(*sys.argv[1:])
