"""CPU check of the Gram form of the SU2 RPA lattice sum (rpaGram in spinparser_b200/csrc/pffrg_kernels.cuh).

The kernel replaces, per RPA phase, the reference's per-node sum R[rid] = sum_i A[rid1_i] * B[rid2_i] over Lattice::getOverlap(rid)
(src/SU2/SU2FrgCore.cpp:250-266) by G = sum_nodes A (x) B followed by ONE walk of the overlap list, R[rid] = sum_i G[rid1_i][rid2_i],
rows of G in blocks. Here the term tables the library builds (pffrg_gram_tables) are walked exactly the way the kernel walks
them -- same thread grid of the block update, same clamping of rows / columns past the end, same word decoding -- on random operands,
and compared with the direct overlap sum. No GPU involved.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import golden


def _tables(d, rows_per_block, bits):
    from spinparser_b200 import ProblemTables, _capi
    from spinparser_b200.frgcore import make_descriptor
    t = ProblemTables.from_pfd(d)
    L = t.n_sites
    Lp = (L + 3) // 4 * 4
    blocks = (Lp + rows_per_block - 1) // rows_per_block
    seg = np.zeros(blocks * L + 1, dtype=np.int32)
    desc = make_descriptor("SU2", t)
    n = _capi.check(_capi.lib.pffrg_gram_tables(C.byref(desc), rows_per_block, bits, None, 0, seg.ctypes.data_as(C.POINTER(C.c_int32))))
    terms = np.zeros(n, dtype=np.uint16)
    assert _capi.lib.pffrg_gram_tables(C.byref(desc), rows_per_block, bits, terms.ctypes.data_as(C.POINTER(C.c_uint16)), n, seg.ctypes.data_as(C.POINTER(C.c_int32))) == n
    return t, L, Lp, blocks, terms, seg


def _emulate(L, Lp, threads, tm, nodes, A, B, terms, seg, bits):
    """The block update and the reduction with the kernel's index arithmetic (gramcfg / gramBlock / gramReduce)."""
    NT = threads // 64 * 64
    PT = NT // 16
    TM = min(tm, (Lp + PT - 1) // PT)
    PB = PT * TM
    TN = (Lp + 15) // 16
    blocks = (Lp + PB - 1) // PB
    out = np.zeros(L)
    for blk in range(blocks):
        rows = PB if blk < blocks - 1 else Lp - (blocks - 1) * PB
        tmb = TM if blk < blocks - 1 else (rows + PT - 1) // PT
        Gs = np.full(PB * Lp, np.nan)  # entries the kernel does not store must never be read
        for tid in range(NT):
            warp, lane = tid >> 5, tid & 31
            tp, tq = (warp >> 1) * 4 + (lane >> 3), (warp & 1) * 8 + (lane & 7)
            for i in range(tmb):
                pa = min(blk * PB + tp + PT * i, Lp - 1)
                for j in range(TN):
                    qb = min(tq + 16 * j, Lp - 1)
                    acc = float(np.dot(A[:nodes, pa], B[:nodes, qb]))
                    if tp + PT * i < rows and tq + 16 * j < Lp:
                        Gs[(tp + PT * i) * Lp + tq + 16 * j] = acc
        for rid in range(L):
            w = terms[seg[blk * L + rid]:seg[blk * L + rid + 1]].astype(np.int64)
            out[rid] += float(np.sum((w >> bits) * Gs[w & ((1 << bits) - 1)]))
    return out, PB


@pytest.mark.parametrize("threads,tm", [(256, 4), (128, 4), (512, 2), (256, 1), (96, 4)])
@pytest.mark.parametrize("case", ["su2_square_r3_nw10", "su2_kagome_r4_nw8", "su2_kagome_r7_nw6"])
def test_gram_tables_reproduce_the_overlap_sum(case, threads, tm):
    d = golden(case)
    L0 = int(d["lattice/size"])
    Lp0 = (L0 + 3) // 4 * 4
    PT = (threads // 64 * 64) // 16
    PB = PT * min(tm, (Lp0 + PT - 1) // PT)
    bits = max(1, int(np.ceil(np.log2(PB * Lp0))))
    t, L, Lp, blocks, terms, seg = _tables(d, PB, bits)
    rng = np.random.default_rng(5)
    nodes = 11
    A = np.zeros((16, Lp)); B = np.zeros((16, Lp))
    A[:, :L] = rng.uniform(-1, 1, (16, L)); B[:, :L] = rng.uniform(-1, 1, (16, L))
    got, pb = _emulate(L, Lp, threads, tm, nodes, A, B, terms, seg, bits)
    assert pb == PB
    want = np.zeros(L)
    off, r1, r2 = t.overlap_offsets, t.overlap_rid1, t.overlap_rid2
    for rid in range(L):
        for i in range(off[rid], off[rid + 1]):
            want[rid] += float(np.dot(A[:nodes, r1[i]], B[:nodes, r2[i]]))
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12), np.abs(got - want).max()
    # every overlap entry is represented exactly once (multiplicities add up)
    assert int(np.sum(terms.astype(np.int64) >> bits)) == int(off[-1])


def test_large_multiplicities_are_split():
    """With few multiplicity bits a merged term is emitted as several words."""
    d = golden("su2_kagome_r7_nw6")
    L = int(d["lattice/size"]); Lp = (L + 3) // 4 * 4
    bits = 15
    assert 16 * Lp <= 1 << bits
    t, L, Lp, blocks, terms, seg = _tables(d, 16, bits)
    assert int((terms >> bits).max()) == 1
    assert int(np.sum(terms.astype(np.int64) >> bits)) == int(t.overlap_offsets[-1])
