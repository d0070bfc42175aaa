"""The TRI core's bilinear term tables, which libpffrg derives from the spin algebra (csrc/pffrg_device.cuh, namespace tri),
term by term against the machine-generated statement lists of the reference (src/TRI/TRIFrgCore.cpp), and -- so that the
check also runs where the reference tree is absent -- against a digest of those lists committed here.
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

REFERENCE_FILE = "/root/reference/src/TRI/TRIFrgCore.cpp"
REGIONS = ["ppLadder", "phLadder", "chalice", "inverseChalice", "RPA"]
# sha256 over the canonical (sorted) rows of all five regions as parsed from the reference file (_canonical_from_reference)
DIGEST = "97b86d26c6569060630e21770a547d14dfd0fcdc84b93e5425b6fe50195de7a1"


def _ours():
    from spinparser_b200 import _capi
    out = {}
    for region, name in enumerate(REGIONS):
        n = _capi.lib.pffrg_tri_terms(region, None, 0)
        assert n == (64 if name == "RPA" else 256)
        buf = np.zeros((n, 4), dtype=np.int32)
        assert _capi.lib.pffrg_tri_terms(region, buf.ctypes.data_as(C.POINTER(C.c_int32)), n) == n
        out[name] = sorted(tuple(int(x) for x in row) for row in buf)
    return out


def _canonical_from_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from gen_tri_terms import parse
    regions = parse(REFERENCE_FILE)

    def symbolic(e):
        const, a, b = e
        if a == -1 and b == -1:
            return const
        return 4 * (a if a >= 0 else 3) + (b if b >= 0 else 3)

    out = {}
    pairs = {"ppLadder": ((0, 1), (2, 3)), "phLadder": ((0, 1), (2, 3)), "chalice": ((0, 4), (2, 5)), "inverseChalice": ((1, 6), (3, 7))}
    for name, (first, second) in pairs.items():
        rows = {pair: sorted((o, s, c1, c2) for o, s, b1, c1, b2, c2 in regions[name] if (b1, b2) == pair) for pair in (first, second)}
        assert rows[first] == rows[second] and len(rows[first]) == 256  # both buffer pairs carry the same table
        out[name] = rows[first]
    rpa = {pair: sorted((o, s, symbolic(e1), symbolic(e2)) for o, s, b1, e1, b2, e2 in regions["RPA"] if (b1, b2) == pair) for pair in ((0, 1), (2, 3))}
    assert rpa[(0, 1)] == rpa[(2, 3)] and len(rpa[(0, 1)]) == 64
    out["RPA"] = rpa[(0, 1)]
    return out


def _digest(tables):
    h = hashlib.sha256()
    for name in REGIONS:
        h.update(name.encode())
        h.update(np.asarray(tables[name], dtype=np.int32).tobytes())
    return h.hexdigest()


def test_tri_tables_match_committed_digest_of_the_reference_lists():
    assert _digest(_ours()) == DIGEST


@pytest.mark.skipif(not os.path.exists(REFERENCE_FILE), reason="reference tree not present (GPU box)")
def test_tri_tables_match_reference_term_by_term():
    ref = _canonical_from_reference()
    assert _digest(ref) == DIGEST, "committed digest is stale"
    ours = _ours()
    for name in ref:
        assert ours[name] == ref[name], name


# ---- egg diagram of the TRI correlation measurement (csrc/pffrg_measure.cuh, tri::egg) against
# src/TRI/TRIMeasurementCorrelation.cpp:204-243: rows {output channel, 4 * coefficient, a, b}
EGG_REFERENCE_FILE = "/root/reference/src/TRI/TRIMeasurementCorrelation.cpp"
EGG_DIGEST = "c4af60e6ceb48742afa72324a32dee8cf6da9a5957d0097e38cc363db192c372"


def _egg_ours():
    from spinparser_b200 import _capi
    n = _capi.lib.pffrg_tri_terms(5, None, 0)
    buf = np.zeros((n, 4), dtype=np.int32)
    assert _capi.lib.pffrg_tri_terms(5, buf.ctypes.data_as(C.POINTER(C.c_int32)), n) == n
    return sorted(tuple(int(x) for x in row) for row in buf)


def _egg_from_reference():
    import re
    idx = {"x": 0, "y": 1, "z": 2, "d": 3}
    rows = []
    for m in re.finditer(r"ret\.bundle\((\d+)\)\[0\] ([+-])= ([0-9.]+)f \* v(\w)(\w);", open(EGG_REFERENCE_FILE).read()):
        sign = 1 if m.group(2) == "+" else -1
        rows.append((int(m.group(1)), int(round(4 * sign * float(m.group(3)))), idx[m.group(4)], idx[m.group(5)]))
    return sorted(rows)


def _rows_digest(rows):
    return hashlib.sha256(np.asarray(rows, dtype=np.int32).tobytes()).hexdigest()


def test_tri_egg_table_matches_committed_digest():
    rows = _egg_ours()
    assert len(rows) == 40
    assert _rows_digest(rows) == EGG_DIGEST


@pytest.mark.skipif(not os.path.exists(EGG_REFERENCE_FILE), reason="reference tree not present (GPU box)")
def test_tri_egg_table_matches_reference_term_by_term():
    ref = _egg_from_reference()
    assert len(ref) == 40 and _rows_digest(ref) == EGG_DIGEST, "committed digest is stale"
    assert _egg_ours() == ref
