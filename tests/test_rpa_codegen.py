"""The lattice-specialised RPA code generator (spinparser_b200/csrc/pffrg_jit.cpp), checked on the CPU: the straight-line code it
emits for a lattice is interpreted statement by statement (operand loads through the software pipeline, multiplicity copies,
multiply-adds into the accumulators, output writes) and its result compared with the defining sum over Lattice::getOverlap
(src/Lattice.hpp:46-96; SU2FrgCore.cpp:250-266, XYZFrgCore.cpp:297-322). No GPU: the source comes from the device-less
pffrg_jit_compile_check (PFFRG_JIT_DUMP)."""
import os
import re

import numpy as np
import pytest

from conftest import golden

CASES = ["su2_square_r3_nw10", "su2_kagome_r7_nw6", "xyz_honeycomb_kitaev_r3_nw10", "xyz_kagome_r4_nw8"]
# launch-shape / code-shape knobs that change the generated text (tiles, node groups, accumulators, operand chunks, re-sync calls)
VARIANTS = [{}, {"PFFRG_JIT_ACC": "3", "PFFRG_JIT_CHUNK": "5", "PFFRG_JIT_PREFETCH": "2"}, {"PFFRG_JIT_TILES": "1", "PFFRG_JIT_ACC": "16"},
            {"PFFRG_SUBCTAS": "2", "PFFRG_THREADS": "128", "PFFRG_JIT_NBT": "32", "PFFRG_JIT_NB": "16", "PFFRG_JIT_RESYNC": "1"}]


def _interpret(source, n_operands, rng):
    """Run the generated function for one lane: returns (A, B, out) with out[o] accumulated over all tiles."""
    b_off = int(re.search(r"const unsigned B = A \+ (\d+)u;", source).group(1))
    loads = [int(m) for m in re.findall(r"ldsOrdered\([AB], (\d+)\)", source)]
    stride = np.gcd.reduce([x for x in loads if x > 0]) if any(loads) else 8
    A, B = rng.uniform(-1, 1, n_operands), rng.uniform(-1, 1, n_operands)
    assert b_off % stride == 0 and max(loads) // stride < n_operands
    out, var = {}, {}
    tiles = 0
    for line in source.splitlines():
        line = line.strip()
        if line.startswith("case "):
            tiles += 1
            var = {}
            continue
        for stmt in [s.strip() for s in line.strip("{} ").split(";") if s.strip()]:
            m = re.fullmatch(r"(?:const )?(?:double )?(\w+) = ldsOrdered\(([AB]), (\d+)\)", stmt)
            if m:
                src = A if m.group(2) == "A" else B
                var[m.group(1)] = src[int(m.group(3)) // stride]
                continue
            m = re.fullmatch(r"double (acc\d+) = 0\.0", stmt)
            if m:
                var[m.group(1)] = 0.0
                continue
            m = re.fullmatch(r"const double (a\d+) = (a_\d+) \* (\d+)\.0", stmt)
            if m:
                var[m.group(1)] = var[m.group(2)] * float(m.group(3))
                continue
            m = re.fullmatch(r"(acc\d+) = fma\((\w+), (b\d+), (acc\d+)\)", stmt)
            if m:
                assert m.group(1) == m.group(4)
                var[m.group(1)] = var[m.group(2)] * var[m.group(3)] + var[m.group(1)]
                continue
            m = re.fullmatch(r"if \(writer\) out\[(\d+)\] \+= v", stmt)
            if m:
                o = int(m.group(1))
                assert o not in out, f"output {o} written twice"
                out[o] = var[f"acc{o}"]
    return A, B, out, tiles


@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: ",".join(f"{k}={x}" for k, x in v.items()) or "default")
@pytest.mark.parametrize("case", CASES)
def test_generated_rpa_code_equals_the_overlap_sum(case, variant, monkeypatch, tmp_path):
    from spinparser_b200 import ProblemTables
    from spinparser_b200.frgcore import jit_compile_check
    for k, x in variant.items():
        monkeypatch.setenv(k, x)
    monkeypatch.setenv("PFFRG_JIT_DUMP", str(tmp_path / "rpa"))
    monkeypatch.setenv("PFFRG_RPA", "code")  # (lattices above PFFRG_GRAM_MIN_TERMS merged terms get the Gram form, which has no generated code, by default)
    monkeypatch.setenv("PFFRG_RELABEL", "0")  # the generated code is compared in the reference's site order (the relabelling has its own test)
    d = golden(case)
    core = bytes(d["core"]).decode()
    assert jit_compile_check(core, ProblemTables.from_pfd(d)) > 10000
    source = open(tmp_path / "rpa.cu").read()
    L = int(d["lattice/size"])
    n_out = L if core == "SU2" else 4 * L
    A, B, out, tiles = _interpret(source, n_out, np.random.default_rng(3))
    assert tiles >= 1 and sorted(out) == list(range(n_out)), "every output is produced exactly once"
    off, r1, r2 = d["lattice/overlap_offsets"], d["lattice/overlap_rid1"], d["lattice/overlap_rid2"]
    p1, p2 = d["lattice/overlap_perm1"], d["lattice/overlap_perm2"]
    want = np.zeros(n_out)
    for rid in range(L):
        for i in range(off[rid], off[rid + 1]):
            if core == "SU2":
                want[rid] += A[r1[i]] * B[r2[i]]
            else:
                for c in range(3):
                    want[c * L + rid] += A[int(p1[i][c]) * L + r1[i]] * B[int(p2[i][c]) * L + r2[i]]
                want[3 * L + rid] += A[3 * L + r1[i]] * B[3 * L + r2[i]]
    got = np.array([out[o] for o in range(n_out)])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)
    if "PFFRG_JIT_RESYNC" in variant:
        # every path through the function (each tile, and warps without a tile) makes the same number of rendezvous calls
        blocks = re.split(r"\bcase \d+:", source)
        counts = {b.count("clusterRendezvous<") for b in blocks[1:]}
        assert len(counts) == 1 and counts.pop() == blocks[0].count("clusterRendezvous<") > 0
