#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

The reference sources under /root/reference/src are compiled (oracle/Makefile -> oracle/_ref/oracle64, the FP64
`#define float double` build, and oracle/_ref/oracle32, the as-shipped FP32 build) and run on the small task files in
tests/golden/tasks/ with the reference's own resource files (/root/reference/res). Each run dumps lattice tables,
meshes, and for selected cutoff steps the vertex state, the vertex flow and the correlation measurement.

Only runs in the build container (needs /root/reference); the resulting .pfd files are committed so that the tests
never read /root/reference at run time.

    python tests/golden/make_fixtures.py [case ...]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PFFRG_REFERENCE", "/root/reference")

# case -> steps whose state/flow are dumped (53 cutoffs: 50 * 0.9^k down to 0.2)
CASES = {
    "su2_square_r3_nw10": [0, 1, 10, 25, 40, 51],
    "su2_kagome_r4_nw8": [0, 12, 30, 51],
    "su2_kagome_r7_nw6": [0, 20, 45],
    "xyz_honeycomb_kitaev_r3_nw10": [0, 10, 25, 40, 51],
    "xyz_kagome_r4_nw8": [0, 30, 51],
    "tri_honeycomb_kg_r3_nw8": [0, 10, 30, 51],
    "tri_kagome_dm_r3_nw6": [0, 25, 51],
}


# End-to-end cases for the C++ drop-in adapter (tests/test_host_adapter.py): the same tasks built from THIS repository's
# resource files (oracle/res, which travel to the GPU box; /root/reference/res does not), full flow with the correlation
# measurement. Only the measurement output and the final state are kept.
E2E_CASES = ["e2e_su2_square_r3_nw10", "e2e_xyz_honeycomb_kitaev_r3_nw10", "e2e_tri_kagome_dm_r3_nw6"]


def make_e2e(cases):
    sys.path.insert(0, ROOT)
    from spinparser_b200.pfd import read_pfd, write_pfd
    for case in cases:
        task = os.path.join(HERE, "tasks", case + ".xml")
        for binary, suffix in (("oracle64", "f64"), ("oracle32", "f32")):
            out = os.path.join(HERE, f"{case}.{suffix}.pfd")
            cmd = [os.path.join(ROOT, "oracle", "_ref", binary), "-r", os.path.join(ROOT, "oracle", "res"), task, "--out", out, "--no-lattice"]
            print(" ".join(cmd))
            subprocess.run(cmd, check=True, cwd=HERE)
            d = read_pfd(out)
            keep = {k: v for k, v in d.items() if k.startswith(("h5/", "final/")) or k in ("core", "cutoff", "frequency", "finalStep")}
            write_pfd(out, keep)
            print(case, suffix, os.path.getsize(out) // 1024, "KiB")


def main(argv):
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    make_e2e([c for c in (argv or E2E_CASES) if c in E2E_CASES])
    cases = [c for c in (argv or list(CASES)) if c in CASES]
    for case in cases:
        task = os.path.join(HERE, "tasks", case + ".xml")
        steps = ",".join(str(s) for s in CASES[case])
        for binary, suffix in (("oracle64", "f64"), ("oracle32", "f32")):
            out = os.path.join(HERE, f"{case}.{suffix}.pfd")
            cmd = [os.path.join(ROOT, "oracle", "_ref", binary), "-r", os.path.join(REF, "res"), task, "--out", out]
            # the FP32 run only pins the measurement output (the bridge to the reference's own .ref goldens): no vertex dumps
            cmd += ["--dump-steps", steps] if suffix == "f64" else ["--no-lattice"]
            print(" ".join(cmd))
            subprocess.run(cmd, check=True, cwd=HERE)
        print(case, os.path.getsize(os.path.join(HERE, f"{case}.f64.pfd")) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1:])
