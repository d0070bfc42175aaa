#!/usr/bin/env python
"""Generate the problem tables (frequency mesh, cutoff grid, symmetry-reduced lattice tables, bare couplings) of the
benchmark workloads by running the reference's own TaskFileParser / LatticeModelFactory (oracle/_ref/oracle64) on
bench_data/tasks/*.xml with the resource files in oracle/res. Lattice generation and symmetry reduction are host-side
steps outside the hot path (component 9); their output is the INPUT of the flow kernels, so the
compact tables are committed and the benchmarks never need the reference tree.

    python bench_data/make_tables.py [name ...]
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from spinparser_b200.pfd import read_pfd, write_pfd  # noqa: E402


def main(argv):
    names = argv or sorted(f[:-4] for f in os.listdir(os.path.join(HERE, "tasks")) if f.endswith(".xml"))
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    for name in names:
        with tempfile.TemporaryDirectory() as tmp:
            raw = os.path.join(tmp, "raw.pfd")
            subprocess.run([os.path.join(ROOT, "oracle", "_ref", "oracle64"), "-r", os.path.join(ROOT, "oracle", "res"),
                            os.path.join(HERE, "tasks", name + ".xml"), "--out", raw, "--max-steps", "0", "--no-measure"], check=True)
            d = read_pfd(raw)
        L = int(d["lattice/size"])
        core = bytes(d["core"]).decode()
        out = {k: v for k, v in d.items() if k.startswith("lattice/") or k in ("core", "frequency", "cutoff", "spinLength", "normalization")}
        out.pop("lattice/positions", None); out.pop("lattice/parameters", None)
        n = 16 * L if core == "TRI" else L
        arrays = sorted(k for k in d if k.startswith("initial/v4_"))
        # the initial condition is frequency independent: keep the first row only (src/SU2/SU2EffectiveAction.hpp:46-59)
        out["bare"] = np.stack([d[k][:n] for k in arrays])
        for k in arrays:
            assert np.array_equal(d[k].reshape(-1, n), np.broadcast_to(d[k][:n], (d[k].size // n, n)))
        write_pfd(os.path.join(HERE, name + ".tables.pfd"), out)
        print(name, core, "Nw", len(d["frequency"]), "L", L, "overlapTotal", len(d["lattice/overlap_rid1"]), "inRange", len(d["lattice/range0_ids"]),
              os.path.getsize(os.path.join(HERE, name + ".tables.pfd")) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1:])
