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


def _tables(d, rows_per_block, warps):
    from spinparser_b200 import ProblemTables, _capi
    from spinparser_b200.frgcore import make_descriptor
    t = ProblemTables.from_pfd(d)
    L = t.n_sites
    Lp = (L + 3) // 4 * 4
    blocks = (Lp + rows_per_block - 1) // rows_per_block
    seg = np.zeros(2 * blocks * warps, dtype=np.int32)
    desc = make_descriptor("SU2", t)
    degree = C.c_double(0.0)
    n = _capi.check(_capi.lib.pffrg_gram_tables(C.byref(desc), rows_per_block, warps, None, 0, seg.ctypes.data_as(C.POINTER(C.c_int32)), None))
    terms = np.zeros(n, dtype=np.uint32)
    assert _capi.lib.pffrg_gram_tables(C.byref(desc), rows_per_block, warps, terms.ctypes.data_as(C.POINTER(C.c_uint32)), n, seg.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(degree)) == n
    return t, L, Lp, blocks, terms, seg.reshape(blocks * warps, 2), degree.value


def _best_row_warps(rt, ct, nw):
    """gramcfg::bestRowWarps: the NW warps as a WP x WQ grid over the RT x CT tiles, fewest tiles on the busiest warp, then fewest loads."""
    best = 1
    for wp in range(1, nw + 1):
        if nw % wp:
            continue
        t = lambda w: -(-rt // w) * -(-ct // (nw // w))
        loads = lambda w: -(-rt // w) + -(-ct // (nw // w))
        if t(wp) < t(best) or (t(wp) == t(best) and loads(wp) < loads(best)):
            best = wp
    return best


def _emulate(L, Lp, threads, PB, nodes, A, B, terms, seg):
    """The block update (tile -> warp map and accumulator fragment layout of gramBlock) and the reduction (gramReduce) with the kernel's
    index arithmetic."""
    NW = threads // 32
    LpG = Lp + 1
    CT = (Lp + 7) // 8
    blocks = (Lp + PB - 1) // PB
    out = np.zeros(L)
    for blk in range(blocks):
        rows = PB if blk < blocks - 1 else Lp - (blocks - 1) * PB
        RT = (rows + 7) // 8
        WP = _best_row_warps(RT, CT, NW); WQ = NW // WP
        MI, NI = -(-RT // WP), -(-CT // WQ)
        assert MI * NI <= 16
        Gs = np.full(PB * LpG, np.nan)  # entries the kernel does not store must never be read with a non-zero multiplicity
        written = 0
        for warp in range(NW):
            wp, wq = warp // WQ, warp % WQ
            for lane in range(32):
                fr, fk = lane >> 2, lane & 3
                for i in range(MI):
                    if wp * MI + i >= RT:
                        continue
                    row = 8 * (wp * MI + i) + fr
                    if blk * PB + row >= Lp:
                        continue
                    for j in range(NI):
                        if wq + WQ * j >= CT:
                            continue
                        for e in range(2):
                            col = 8 * (wq + WQ * j) + 2 * fk + e
                            if col < Lp:
                                assert np.isnan(Gs[row * LpG + col]), "every entry has one owner"
                                Gs[row * LpG + col] = float(np.dot(A[:nodes, blk * PB + row], B[:nodes, col]))
                                written += 1
        assert written == min(rows, Lp - blk * PB) * Lp
        owner = {}
        for warp in range(NW):
            begin, T4 = seg[blk * NW + warp]
            assert begin % 4 == 0 and T4 >= 0
            words = terms[begin:begin + 128 * T4].astype(np.int64).reshape(T4, 32, 4)  # [group][lane][4]
            trailing = []
            for lane in range(32):
                acc, last = 0.0, 0
                for g in range(T4):
                    w = words[g, lane]
                    rid = int((w[3] >> 14) & 255)
                    mult = (w >> 22) & 511
                    assert np.all(((w >> 14) & 255) == rid), "the words of a group belong to one rid"
                    assert not (w[:3] >> 31).any(), "flush flag on the last word of a group only"
                    val = Gs[w & 0x3FFF]
                    assert not np.isnan(val[mult > 0]).any()
                    acc += float(np.sum(np.where(mult > 0, mult * np.nan_to_num(val), 0.0)))
                    last = rid
                    if mult.any() or (w[3] >> 31):
                        assert owner.setdefault(rid, warp) == warp, "single writer per output and block"
                    if w[3] >> 31:
                        out[rid] += acc
                        acc = 0.0
                trailing.append((last, acc))
            # what the lanes still hold: runs of equal rid over consecutive lanes (the segmented scan of gramReduce)
            for lane, (rid, acc) in enumerate(trailing):
                if acc != 0.0:
                    assert owner.get(rid, warp) == warp
                out[rid] += acc
            seen, prev = set(), None
            for r, a in trailing:
                if r != prev:
                    assert r not in seen or a == 0.0, "a rid's unfinished pieces sit in consecutive lanes"
                    seen.add(r); prev = r
    return out


@pytest.mark.parametrize("threads,pb", [(256, 64), (128, 32), (512, 64), (256, 8), (96, 16), (192, 24)])
@pytest.mark.parametrize("case", ["su2_square_r3_nw10", "su2_kagome_r4_nw8", "su2_kagome_r7_nw6"])
def test_gram_tables_reproduce_the_overlap_sum(case, threads, pb):
    d = golden(case)
    L0 = int(d["lattice/size"])
    Lp0 = (L0 + 3) // 4 * 4
    PB = min(pb, (Lp0 + 7) // 8 * 8)
    t, L, Lp, blocks, terms, seg, degree = _tables(d, PB, threads // 32)
    rng = np.random.default_rng(5)
    nodes = 11
    A = np.zeros((16, Lp)); B = np.zeros((16, Lp))
    A[:, :L] = rng.uniform(-1, 1, (16, L)); B[:, :L] = rng.uniform(-1, 1, (16, L))
    got = _emulate(L, Lp, threads, PB, nodes, A, B, terms, seg)
    want = np.zeros(L)
    off, r1, r2 = t.overlap_offsets, t.overlap_rid1, t.overlap_rid2
    for rid in range(L):
        for i in range(off[rid], off[rid + 1]):
            want[rid] += float(np.dot(A[:nodes, r1[i]], B[:nodes, r2[i]]))
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12), np.abs(got - want).max()
    # every overlap entry is represented exactly once (multiplicities add up)
    assert int(np.sum((terms.astype(np.int64) >> 22) & 511)) == int(off[-1])
    assert 1.0 <= degree <= 8.0


def test_term_order_of_the_benchmark_lattice_is_nearly_conflict_free():
    """pyrochlore-r8 (L = 103, 40 355 merged terms): the word order keeps the 8 lanes of a quarter warp on different 16-byte bank groups."""
    import os
    from conftest import ROOT
    from spinparser_b200 import read_pfd
    d = read_pfd(os.path.join(ROOT, "bench_data", "pyrochlore_r8_su2_nw64.tables.pfd"))
    t, L, Lp, blocks, terms, seg, degree = _tables(d, 56, 8)
    assert blocks == 2 and L == 103
    assert int(np.sum((terms.astype(np.int64) >> 22) & 511)) == int(t.overlap_offsets[-1])
    assert len(terms) < 1.15 * 40355, "padding overhead of the term array"
    assert degree < 1.5, degree  # a random order gives ~2.5


@pytest.mark.parametrize("case", ["su2_square_r3_nw10", "su2_kagome_r4_nw8", "su2_kagome_r7_nw6", "xyz_honeycomb_kitaev_r3_nw10", "tri_kagome_dm_r3_nw6", "bench:pyrochlore_r8_su2_nw64"])
def test_device_site_order_pairs_inverted_sites(case):
    """pffrg_site_order: a permutation with site 0 first in which j and getInvertedSites()[j] are neighbours within one 128-byte line
    (8 sites of 16 bytes), so that gathers with the site-exchange flag are as coalesced as plain ones."""
    import os
    from conftest import ROOT
    from spinparser_b200 import ProblemTables, _capi, read_pfd
    from spinparser_b200.frgcore import make_descriptor
    d = read_pfd(os.path.join(ROOT, "bench_data", case[6:] + ".tables.pfd")) if case.startswith("bench:") else golden(case)
    t = ProblemTables.from_pfd(d)
    L = t.n_sites
    order = np.zeros(L, dtype=np.int32)
    core = bytes(d["core"]).decode()
    changed = _capi.check(_capi.lib.pffrg_site_order(C.byref(make_descriptor(core, t)), order.ctypes.data_as(C.POINTER(C.c_int32))))
    assert sorted(order) == list(range(L)) and order[0] == 0
    new_of = np.argsort(order)
    inv = t.inverted_rid
    moved = int((inv != np.arange(L)).sum())
    assert changed == (0 if np.array_equal(order, np.arange(L)) else 1)
    straddling = 0
    for j in range(L):
        a, b = int(new_of[j]), int(new_of[inv[j]])
        assert abs(a - b) <= 1, "pair members are neighbours"
        straddling += a // 8 != b // 8
    assert straddling <= max(2, moved // 8), (straddling, moved)


@pytest.mark.parametrize("case,resident,warps", [("tri_kagome_dm_r3_nw6", 3, 8), ("tri_honeycomb_kg_r3_nw8", 2, 8), ("tri_kagome_dm_r3_nw6", 16, 4)])
def test_tri_gram_tables_reproduce_the_overlap_sum(case, resident, warps):
    """TRI: R^{mu nu}[rid] = sum_i sum_k eta(mu,k,nu) A^{p1 mu,p1 k}[rid1_i] B^{p2 k,p2 nu}[rid2_i] (the loop rpaTri8 walks, eta from the
    spin algebra the kernels use) against the Gram blocks G^{(c1,c2)} = sum_rows A^{c1} (x) B^{c2} reduced with the signed term words of
    pffrg_trigram_tables, walked chunk by chunk and lane by lane like triGramReduce."""
    from spinparser_b200 import ProblemTables, _capi
    from spinparser_b200.frgcore import make_descriptor
    d = golden(case)
    t = ProblemTables.from_pfd(d)
    L = t.n_sites
    LpT = (L + 7) // 8 * 8
    GS, GBLK = LpT + 1, LpT * (LpT + 1)
    desc = make_descriptor("TRI", t)
    rounds = C.c_int32(0)
    n = _capi.check(_capi.lib.pffrg_trigram_tables(C.byref(desc), resident, warps, None, 0, None, 0, None, 0, C.byref(rounds)))
    blocks = np.zeros(rounds.value * resident, dtype=np.uint16)
    terms = np.zeros(n, dtype=np.uint32)
    seg = np.zeros(2 * rounds.value * warps, dtype=np.int32)
    assert _capi.lib.pffrg_trigram_tables(C.byref(desc), resident, warps, blocks.ctypes.data_as(C.POINTER(C.c_uint16)), len(blocks), terms.ctypes.data_as(C.POINTER(C.c_uint32)), n,
                                          seg.ctypes.data_as(C.POINTER(C.c_int32)), len(seg), C.byref(rounds)) == n
    seg = seg.reshape(rounds.value * warps, 2)
    # eta(mu, k, nu) as the kernels derive it (region 4 of pffrg_tri_terms: rows {out, sign, 4 mu + k, 4 k + nu})
    rows = np.zeros((64, 4), dtype=np.int32)
    assert _capi.lib.pffrg_tri_terms(4, rows.ctypes.data_as(C.POINTER(C.c_int32)), 64) == 64
    eta = {(int(r[2]) // 4, int(r[2]) % 4, int(r[3]) % 4): int(r[1]) for r in rows}
    rng = np.random.default_rng(11)
    K = 6
    A = rng.uniform(-1, 1, (K, 16, LpT)); B = rng.uniform(-1, 1, (K, 16, LpT))
    A[:, :, L:] = 0.0; B[:, :, L:] = 0.0
    want = np.zeros(16 * L)
    off, r1, r2, p1, p2 = t.overlap_offsets, t.overlap_rid1, t.overlap_rid2, t.overlap_perm1, t.overlap_perm2
    perm = lambda p, i: 3 if i == 3 else int(p[i])
    for rid in range(L):
        for i in range(off[rid], off[rid + 1]):
            for mu in range(4):
                for k in range(4):
                    for nu in range(4):
                        c1, c2 = 4 * perm(p1[i], mu) + perm(p1[i], k), 4 * perm(p2[i], k) + perm(p2[i], nu)
                        want[(4 * mu + nu) * L + rid] += eta[(mu, k, nu)] * float(np.dot(A[:, c1, r1[i]], B[:, c2, r2[i]]))
    got = np.zeros(16 * L)
    for rnd in range(rounds.value):
        Gs = np.full(resident * GBLK, np.nan)
        for slot in range(resident):
            pair = int(blocks[rnd * resident + slot])
            if pair == 0xFFFF:
                continue
            G = np.einsum("kp,kq->pq", A[:, pair & 15, :], B[:, pair >> 4, :])
            for p in range(LpT):
                Gs[slot * GBLK + p * GS:slot * GBLK + p * GS + LpT] = G[p]
        owner = {}
        for w in range(warps):
            begin, end = seg[rnd * warps + w]
            assert begin % 256 == 0 and (end - begin) % 256 == 0
            for pos in range(begin, end, 8):
                wd = terms[pos:pos + 8].astype(np.int64)
                out = int((wd[0] >> 13) & 1023)
                mult = (wd >> 24) * (1 - 2 * ((wd >> 23) & 1))
                assert np.all(((wd >> 13) & 1023)[mult != 0] == out)
                assert owner.setdefault(out, w) == w or not mult.any()
                g = Gs[wd & 8191]
                assert not np.isnan(g[mult != 0]).any()
                got[out] += float(np.sum(np.where(mult != 0, mult * np.nan_to_num(g), 0.0)))
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12), np.abs(got - want).max()


@pytest.mark.parametrize("warps,pb", [(4, 32), (8, 16), (4, 56)])
@pytest.mark.parametrize("case", ["xyz_honeycomb_kitaev_r3_nw10", "xyz_kagome_r4_nw8"])
def test_xyz_gram_tables_reproduce_the_overlap_sum(case, warps, pb):
    """XYZ: R_c[rid] = sum_i A_{p1_i(c)}[rid1_i] B_{p2_i(c)}[rid2_i] for the spin channels (p1, p2 the overlap's spin permutations) and
    R_d = sum_i A_d B_d (src/XYZ/XYZFrgCore.cpp:297-322). The kernel stages the three spin channels of a site as virtual sites c L + j
    with the density channel in the second component of the c = 0 entries, forms one (3 L)^2 Gram matrix of pairs and walks term words
    whose outputs are c L + rid (first component) and 3 L + rid (second component). Same walk as gramReduce, on random operands."""
    from spinparser_b200 import ProblemTables, _capi
    from spinparser_b200.frgcore import make_descriptor
    d = golden(case)
    t = ProblemTables.from_pfd(d)
    L = t.n_sites
    Lv, Lp, LOUT = 3 * L, (3 * L + 3) // 4 * 4, 4 * L
    PB = min(pb, (Lp + 7) // 8 * 8)
    blocks = (Lp + PB - 1) // PB
    desc = make_descriptor("XYZ", t)
    seg = np.zeros(2 * blocks * warps, dtype=np.int32)
    n = _capi.check(_capi.lib.pffrg_gram_tables(C.byref(desc), PB, warps, None, 0, seg.ctypes.data_as(C.POINTER(C.c_int32)), None))
    terms = np.zeros(n, dtype=np.uint32)
    assert _capi.lib.pffrg_gram_tables(C.byref(desc), PB, warps, terms.ctypes.data_as(C.POINTER(C.c_uint32)), n, seg.ctypes.data_as(C.POINTER(C.c_int32)), None) == n
    seg = seg.reshape(blocks * warps, 2)
    rng = np.random.default_rng(9)
    nodes = 7
    A = rng.uniform(-1, 1, (nodes, 4, L)); B = rng.uniform(-1, 1, (nodes, 4, L))
    # staged operands: [node][virtual site] -> (x, y)
    Ax = np.zeros((nodes, Lp)); Ay = np.zeros((nodes, Lp)); Bx = np.zeros((nodes, Lp)); By = np.zeros((nodes, Lp))
    for c in range(3):
        Ax[:, c * L:(c + 1) * L] = A[:, c]; Bx[:, c * L:(c + 1) * L] = B[:, c]
    Ay[:, :L] = A[:, 3]; By[:, :L] = B[:, 3]
    Gx, Gy = Ax.T @ Bx, Ay.T @ By
    out = np.zeros((2, 2 * LOUT))  # [component][output]
    owner = {}
    for blk in range(blocks):
        for warp in range(warps):
            begin, T4 = seg[blk * warps + warp]
            words = terms[begin:begin + 128 * T4].astype(np.int64).reshape(T4, 32, 4)
            for lane in range(32):
                for g in range(T4):
                    w = words[g, lane]
                    o = int((w[3] >> 14) & 255)
                    mult = (w >> 22) & 511
                    assert np.all(((w >> 14) & 255) == o) and o < LOUT
                    if mult.any():
                        assert owner.setdefault((blk, o), warp) == warp
                    off = w & 0x3FFF
                    p, q = blk * PB + off // (Lp + 1), off % (Lp + 1)
                    assert np.all(q[mult > 0] < Lv) and np.all(p[mult > 0] < Lv)
                    out[0, o] += float(np.sum(mult * Gx[np.minimum(p, Lp - 1), np.minimum(q, Lp - 1)]))
                    out[1, o] += float(np.sum(mult * Gy[np.minimum(p, Lp - 1), np.minimum(q, Lp - 1)]))
    got = np.concatenate([out[0, :3 * L], out[1, 3 * L:4 * L]]).reshape(4, L)
    want = np.zeros((4, L))
    off, r1, r2, p1, p2 = t.overlap_offsets, t.overlap_rid1, t.overlap_rid2, t.overlap_perm1, t.overlap_perm2
    for rid in range(L):
        for i in range(off[rid], off[rid + 1]):
            for c in range(3):
                want[c, rid] += float(np.dot(A[:, int(p1[i][c]), r1[i]], B[:, int(p2[i][c]), r2[i]]))
            want[3, rid] += float(np.dot(A[:, 3, r1[i]], B[:, 3, r2[i]]))
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12), np.abs(got - want).max()
    assert int(np.sum((terms.astype(np.int64) >> 22) & 511)) == 4 * int(off[-1])
