"""Multi-GPU host logic on the CPU: the work-item partition (pffrg_plan_partition) and the exchange pattern of the sharded
step, exercised with world_size-2 `gloo` process groups. The flow values come from the oracle here (this container has no
GPU); the `-m gpu` suite and bench.py run the same pattern with the CUDA kernels and NCCL.
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, golden


def _tables(case):
    from spinparser_b200 import ProblemTables
    d = golden(case)
    return d, ProblemTables.from_pfd(d)


@pytest.mark.parametrize("case,core", [("su2_square_r3_nw10", "SU2"), ("xyz_kagome_r4_nw8", "XYZ"), ("tri_kagome_dm_r3_nw6", "TRI")])
@pytest.mark.parametrize("ranks", [1, 2, 3, 8])
def test_partition_is_contiguous_complete_and_balanced(case, core, ranks):
    from spinparser_b200.frgcore import plan_partition
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_port import OraclePort
    d, tables = _tables(case)
    port = OraclePort(d)
    nw, nf = tables.n_frequencies, tables.n_items
    for cutoff in (float(d["cutoff"][0]), float(d["cutoff"][20]), float(d["cutoff"][-1])):
        bounds = plan_partition(core, tables, cutoff, ranks)
        assert bounds[0] == 0 and bounds[-1] == nf and all(a <= b for a, b in zip(bounds, bounds[1:]))
        # balance against the exact node counts of the oracle (the unit of SURVEY.md 8d)
        n = np.array([port.node_count(cutoff, float(w)) for w in port.mesh], dtype=np.float64)
        so, uo = np.tril_indices(nw)
        order = np.argsort(so * (so + 1) // 2 + uo)
        so, uo = so[order], uo[order]
        per_item = (n[so][:, None] + n[uo][:, None] + 4.0 * n[None, :]).reshape(-1)  # t-channel nodes are the expensive ones
        loads = np.array([per_item[a:b].sum() for a, b in zip(bounds, bounds[1:])])
        if ranks > 1 and nf >= 50 * ranks:
            assert loads.max() <= 1.35 * loads.mean(), (cutoff, bounds, loads)


def _worker(rank, world, port_no, case, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_port import OraclePort
    from spinparser_b200 import ProblemTables
    from spinparser_b200.frgcore import plan_partition
    from spinparser_b200.pfd import read_pfd

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = read_pfd(os.path.join(ROOT, "tests", "golden", case + ".f64.pfd"))
    tables = ProblemTables.from_pfd(d)
    port = OraclePort(d)
    pre = "step10/"
    cutoff, new_cutoff = float(d["cutoff"][10]), float(d["cutoff"][11])
    v2 = np.ascontiguousarray(d[pre + "state/v2"])
    v4 = [np.ascontiguousarray(d[pre + f"state/v4_{c}"]) for c in range(port.n_arrays)]
    # every rank computes the (tiny) self-energy flow redundantly: no exchange for it
    f2 = port.v2_flow(cutoff, v2, v4)
    bounds = plan_partition(port.core, tables, cutoff, world)
    begin, end = bounds[rank], bounds[rank + 1]
    mine = port.v4_flow(cutoff, v2, f2, v4, np.arange(begin, end, dtype=np.int32))
    per_item = port.array_len // port.nf
    new_state = []
    for c in range(port.n_arrays):
        # Euler update of the own slice, then every rank broadcasts its slice (the pattern of pffrg_finalize_step)
        state = torch.from_numpy(v4[c].copy())
        lo, hi = begin * per_item, end * per_item
        state[lo:hi] += (new_cutoff - cutoff) * torch.from_numpy(mine[c][lo:hi])
        for r in range(world):
            a, b = bounds[r] * per_item, bounds[r + 1] * per_item
            if b > a:
                piece = state[a:b].clone()
                dist.broadcast(piece, src=r)
                state[a:b] = piece
        new_state.append(state.numpy())
    diverged = torch.tensor([int(any(np.isnan(a).any() for a in mine))])
    dist.all_reduce(diverged, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.stack(new_state))
    assert int(diverged) == 0
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["su2_square_r3_nw10", "xyz_honeycomb_kitaev_r3_nw10"])
def test_two_rank_sharded_step_reproduces_the_reference_state(case, tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port_no, case, str(tmp_path)), nprocs=2, join=True)
    d = golden(case)
    got = [np.load(tmp_path / f"rank{r}.npy") for r in range(2)]
    assert np.array_equal(got[0], got[1]), "ranks disagree after the exchange"
    n = got[0].shape[0]
    step = float(d["cutoff"][11]) - float(d["cutoff"][10])
    for c in range(n):
        want = d["step10/state/v4_%d" % c] + step * d["step10/flow/v4_%d" % c]
        np.testing.assert_allclose(got[0][c], want, rtol=1e-10, atol=1e-12 * np.abs(want).max())


@pytest.mark.parametrize("ranks", [2, 4, 8])
def test_feedback_partition_converges_on_position_dependent_cost(ranks):
    """pffrg_plan_partition_feedback: when the true cost per item deviates from the model along the item axis (cache locality of
    the gathers does, measured on 4 GPUs), feeding the measured times of one step into the next partition evens the ranks out."""
    from spinparser_b200 import ProblemTables, read_pfd
    from spinparser_b200.frgcore import plan_partition, plan_partition_feedback
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_port import OraclePort
    d = read_pfd(os.path.join(ROOT, "bench_data", "cubic_r7_su2_nw64.tables.pfd"))
    tables = ProblemTables.from_pfd(d)
    port = OraclePort(d)
    nw, nf = tables.n_frequencies, tables.n_items
    cutoff = float(d["cutoff"][214])
    n = np.array([port.node_count(cutoff, float(w)) for w in port.mesh], dtype=np.float64)
    so, uo = np.tril_indices(nw)
    order = np.argsort(so * (so + 1) // 2 + uo)
    so, uo = so[order], uo[order]
    x = np.arange(nf) / nf
    true_cost = (n[so][:, None] + n[uo][:, None] + 3.0 * n[None, :]).reshape(-1) * (0.75 + 0.6 * x ** 2)  # late items are slower than modelled
    prefix = np.concatenate([[0.0], np.cumsum(true_cost)])

    def times(bounds):
        return np.array([prefix[b] - prefix[a] for a, b in zip(bounds, bounds[1:])])

    bounds = plan_partition("SU2", tables, cutoff, ranks)
    first = times(bounds)
    for _ in range(4):
        bounds = plan_partition_feedback("SU2", tables, cutoff, bounds, times(bounds))
        assert bounds[0] == 0 and bounds[-1] == nf and all(a <= b for a, b in zip(bounds, bounds[1:]))
    last = times(bounds)
    assert first.max() / first.mean() > 1.10, "the scenario must start out unbalanced"
    assert last.max() / last.mean() < 1.02, (first.max() / first.mean(), last.max() / last.mean(), bounds)
    # equal times are a fixed point, and the static split is reproduced by uniform feedback
    again = plan_partition_feedback("SU2", tables, cutoff, bounds, times(bounds))
    assert max(abs(a - b) for a, b in zip(again, bounds)) <= nf // 200


def _two_steps(port, d, world, rank, dist, tables):
    """Two cutoff steps of the sharded flow as the library runs them: static split in the first step, measured times of every rank
    shared by a sum over one-hot vectors (pffrg_finalize_step), feedback-balanced split in the second."""
    import time
    import torch
    from spinparser_b200.frgcore import plan_partition, plan_partition_feedback
    cut = [float(x) for x in d["cutoff"][10:13]]
    v2 = np.ascontiguousarray(d["step10/state/v2"]).copy()
    v4 = [np.ascontiguousarray(d[f"step10/state/v4_{c}"]).copy() for c in range(port.n_arrays)]
    per_item = port.array_len // port.nf
    bounds, history = None, []
    for k in range(2):
        if bounds is None:
            bounds = plan_partition(port.core, tables, cut[k], world)
        else:
            bounds = plan_partition_feedback(port.core, tables, cut[k], bounds, times)
        history.append(list(bounds))
        begin, end = bounds[rank], bounds[rank + 1]
        f2 = port.v2_flow(cut[k], v2, v4)
        t0 = time.perf_counter()
        mine = port.v4_flow(cut[k], v2, f2, v4, np.arange(begin, end, dtype=np.int32))
        elapsed = time.perf_counter() - t0
        if dist is not None:
            vec = torch.zeros(world, dtype=torch.float64)
            vec[rank] = elapsed * 1e3
            dist.all_reduce(vec)
            times = vec.tolist()
        else:
            times = [elapsed * 1e3]
        v2 = v2 + (cut[k + 1] - cut[k]) * f2
        for c in range(port.n_arrays):
            state = torch.from_numpy(v4[c])
            lo, hi = begin * per_item, end * per_item
            state[lo:hi] += (cut[k + 1] - cut[k]) * torch.from_numpy(mine[c][lo:hi])
            for r in range(world if dist is not None else 0):
                a, b = bounds[r] * per_item, bounds[r + 1] * per_item
                if b > a:
                    piece = state[a:b].clone()
                    dist.broadcast(piece, src=r)
                    state[a:b] = piece
    return v2, v4, history


def _worker_two_steps(rank, world, port_no, case, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_port import OraclePort
    from spinparser_b200 import ProblemTables
    from spinparser_b200.pfd import read_pfd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = read_pfd(os.path.join(ROOT, "tests", "golden", case + ".f64.pfd"))
    port = OraclePort(d)
    v2, v4, history = _two_steps(port, d, world, rank, dist, ProblemTables.from_pfd(d))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), v2=v2, v4=np.stack(v4), bounds=np.array(history))
    dist.destroy_process_group()


def test_two_rank_feedback_partition_over_two_steps(tmp_path):
    """world_size 2 on gloo: the second step's split comes from the first step's measured times; both ranks derive the same
    boundaries from the shared times and end in the state of an unsharded two-step run, bit for bit."""
    import torch.multiprocessing as mp
    from oracle_port import OraclePort
    from spinparser_b200 import ProblemTables
    case = "su2_square_r3_nw10"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    mp.spawn(_worker_two_steps, args=(2, port_no, case, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    assert np.array_equal(got[0]["bounds"], got[1]["bounds"]), "ranks derived different partitions"
    b = got[0]["bounds"]
    assert b.shape == (2, 3) and (b[:, 0] == 0).all() and (b[:, 2] == b[0, 2]).all() and (np.diff(b, axis=1) >= 0).all()
    assert np.array_equal(got[0]["v4"], got[1]["v4"]) and np.array_equal(got[0]["v2"], got[1]["v2"])
    d = golden(case)
    v2, v4, _ = _two_steps(OraclePort(d), d, 1, 0, None, ProblemTables.from_pfd(d))
    assert np.array_equal(got[0]["v2"], v2)
    assert np.array_equal(got[0]["v4"], np.stack(v4)), "sharded two-step state differs from the unsharded one"
