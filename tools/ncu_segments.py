#!/usr/bin/env python
"""Split the SASS of an ncu report (source page) at BAR.SYNC / EXIT and print samples, executed instructions and the top
stall reasons per segment -- a poor man's per-phase profile of the flow kernel.

    python tools/ncu_segments.py <report.ncu-rep> [min_pct]
"""
import csv, io, subprocess, sys

def main(report, min_pct=0.5):
    raw = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) >= len(h)]
    col = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    seg = []; cur = None
    def new(i): return {"start": i, "samples": 0, "exec": 0, "fp64": 0, "ldg": 0, "lds": 0, "n": 0, "st": {n: 0 for n in stalls}}
    cur = new(0)
    for i, r in enumerate(data):
        src = r[col["Source"]]
        s = int(r[col["# Samples"]] or 0); e = int(r[col["Instructions Executed"]] or 0)
        cur["samples"] += s; cur["exec"] += e; cur["n"] += 1
        if any(k in src for k in ("DFMA", "DMUL", "DADD")): cur["fp64"] += e
        if "LDG" in src: cur["ldg"] += e
        if "LDS" in src or "STS" in src: cur["lds"] += e
        for n in stalls: cur["st"][n] += int(r[col[n]] or 0)
        if "BAR.SYNC" in src or "EXIT" in src:
            cur["end"] = i; seg.append(cur); cur = new(i + 1)
    tot = sum(s["samples"] for s in seg) or 1
    print(f"total samples {tot}, {len(data)} instructions")
    for s in seg:
        if 100.0 * s["samples"] / tot < float(min_pct): continue
        top = sorted(s["st"].items(), key=lambda kv: -kv[1])[:4]
        print(f"{s['start']:6d}-{s['end']:6d} n={s['n']:6d} samples {100*s['samples']/tot:5.1f}%  exec {s['exec']/1e9:7.3f}G fp64 {s['fp64']/1e9:7.3f}G ldg {s['ldg']/1e9:6.3f}G lds {s['lds']/1e9:6.3f}G | "
              + " ".join(f"{n[6:]}={100*v/max(s['samples'],1):.0f}%" for n, v in top))

if __name__ == "__main__":
    main(*sys.argv[1:])
