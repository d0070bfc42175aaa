#!/usr/bin/env python
"""Share of each kernel in the steady-state steps of an `ncu --metrics gpu__time_duration.sum` launch list of bench.py.

    python tools/launch_share.py <launches.csv> [steps]

The list starts with set-up work (autotuning launches of the run-time compiled kernel, state upload); the last `steps`
cutoff steps (default 2) are the timed ones: a step = v2FlowKernel, nodeTableKernel, flow kernel, 2 x eulerKernel, setScalarKernel.
"""
import csv, sys
from collections import OrderedDict

def main(path, steps=2):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 14 and r[0].isdigit()]
    names = [r[4].split("(")[0].replace("void ", "").replace("pffrg::", "") for r in rows]
    ns = [float(r[14]) for r in rows]
    # a step ends with setScalarKernel preceded by eulerKernel
    ends = [i for i, n in enumerate(names) if n.startswith("setScalarKernel") and i > 0 and names[i - 1].startswith("eulerKernel")]
    ends = ends[-int(steps):]
    start = ends[0] - 5
    share = OrderedDict()
    for n, t in zip(names[start:ends[-1] + 1], ns[start:ends[-1] + 1]):
        share[n] = share.get(n, 0.0) + t
    total = sum(share.values())
    print(f"{path}: last {len(ends)} steps, launches {start}..{ends[-1]} of {len(rows)}, {total / len(ends) / 1e6:.3f} ms per step under ncu (serialised, cold caches)")
    for n, t in share.items():
        print(f"  {n:45s} {t / len(ends) / 1e3:10.1f} us/step  {100 * t / total:6.2f} %")

if __name__ == "__main__":
    main(*sys.argv[1:])
