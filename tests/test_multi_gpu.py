"""Sharded flow on several GPUs of one box (one process per GPU, NCCL) against the single-GPU flow.

Every rank runs the same task: initial condition, K cutoff steps with the work items split over the ranks
(pffrg_plan_partition) and the updated slices exchanged after every Euler update (ncclBroadcast group). Rank 0 repeats
the run on a private single-GPU core. A work item is computed by the same code on whichever GPU owns it, so the
sharded state must equal the single-GPU state BIT FOR BIT, and both must match the reference dump.
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on smaller boxes.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = sys.argv[1]; case = sys.argv[2]; steps = int(sys.argv[3])
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from spinparser_b200 import FrgCoreFactory, ProblemTables
from spinparser_b200.pfd import read_pfd
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = read_pfd(os.path.join(ROOT, "tests", "golden", case + ".f64.pfd"))
name = bytes(d["core"]).decode()
opts = {"spin": str(float(d["spinLength"]))} if name == "SU2" else {}
n = {"SU2": 2, "XYZ": 4, "TRI": 1}[name]
cut = [float(x) for x in d["cutoff"]]
start = int(sys.argv[4])
v2 = np.ascontiguousarray(d[f"step{start}/state/v2"]); v4 = [np.ascontiguousarray(d[f"step{start}/state/v4_{c}"]) for c in range(n)]

def run(core, sharded):
    core.setState(cut[start], v2, v4, sharded=sharded and os.environ.get("TEST_SHARDED_UPLOAD") == "1")
    ranges = []
    for k in range(steps):
        assert not core.computeStep()
        ranges.append(core.itemRange())
        if k == 0:
            flow = core.flow()  # gathers the slices of all ranks
        core.finalizeStep(cut[start + k + 1])
    return core.flowingFunctional(), flow, ranges

# launch shape: rank 0 creates its core first (autotuning when PFFRG_AUTOTUNE=1 is set), the others adopt its shape (as bench.py does)
core = FrgCoreFactory.newFrgCore(name, ProblemTables.from_pfd(d), opts, device=local) if rank == 0 else None
ids = [(core.uniqueId(), core.shapeEnvironment()) if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
os.environ.update(ids[0][1])  # rank 0 too: its single-GPU comparison core below must use the same shape
if rank != 0:
    core = FrgCoreFactory.newFrgCore(name, ProblemTables.from_pfd(d), opts, device=local)
core.initCommunicator(ids[0][0], rank, world)
state, flow, ranges = run(core, True)
# download of a slice only: the rows of this rank's upload share, into a host copy that is otherwise stale (zeros)
b, e = core.uploadSlice()
from spinparser_b200.frgcore import EffectiveAction
part = EffectiveAction(name, len(v2), int(d["lattice/size"]))
core.flowingFunctional(into=part, items=(b, e))
per = len(v4[0]) // (len(v2) * len(v2) * (len(v2) + 1) // 2)
for c in range(n):
    assert np.array_equal(part.v4[c][b * per:e * per], state.v4[c][b * per:e * per]) and not part.v4[c][:b * per].any() and not part.v4[c][e * per:].any()
core.close()
nf = len(v4[0]) // (int(d["lattice/size"]) * (16 if name == "TRI" else 1))
all_ranges = [None] * world
dist.all_gather_object(all_ranges, ranges)
if rank == 0:
    for k in range(steps):  # the ranges of every step tile [0, nf) in rank order
        edges = [r[k] for r in all_ranges]
        assert edges[0][0] == 0 and edges[-1][1] == nf and all(a[1] == b[0] for a, b in zip(edges, edges[1:])), edges
        assert sum(1 for a in edges if a[1] > a[0]) >= min(world, 2), edges
    single = FrgCoreFactory.newFrgCore(name, ProblemTables.from_pfd(d), opts, device=local)
    ref_state, ref_flow, _ = run(single, False)
    single.close()
    assert state.cutoff == ref_state.cutoff
    assert np.array_equal(state.v2, ref_state.v2)
    for c in range(n):
        assert np.array_equal(flow.v4[c], ref_flow.v4[c]), f"flow channel {c} differs between the sharded and the single-GPU run"
        assert np.array_equal(state.v4[c], ref_state.v4[c]), f"state channel {c} differs between the sharded and the single-GPU run"
        want = d[f"step{start}/flow/v4_{c}"]
        err = np.abs(flow.v4[c] - want)
        assert (err <= 1e-10 * np.abs(want) + 1e-12 * np.abs(want).max()).all(), "sharded flow differs from the reference dump"
    print("MULTI_GPU_OK", world, name, [r[0] for r in all_ranges])
# divergence on ONE rank's slice must be reported by ALL ranks (the NaN flag is max-reduced; the reference's ranks all see the
# broadcast flow and stop together, SpinParser.cpp:151-155)
core = FrgCoreFactory.newFrgCore(name, ProblemTables.from_pfd(d), opts, device=local)
ids = [core.uniqueId() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
core.initCommunicator(ids[0], rank, world)
bad = [a.copy() for a in v4]
per = len(v4[0]) // nf
for a in bad:
    a[(nf - 3) * per:] *= 1e160  # the last work items: the last rank's share
core.setState(cut[start], v2, bad)
diverged = core.computeStep()
first, last = core.itemRange()
flags = [None] * world
dist.all_gather_object(flags, (bool(diverged), first, last))
if rank == 0:
    assert all(f[0] for f in flags), flags
    print("MULTI_GPU_DIVERGENCE_OK", flags)
core.close()
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# exchange of the updated slices: fused Euler + peer-memory stores over NVLink (default), the NCCL broadcast group (PFFRG_EXCHANGE=nccl);
# upload of 1/N of the rows per rank + distribution over NVLink; launch shape autotuned on rank 0 and adopted by the other ranks
# su2_kagome_r7_nw6 runs the warp-specialised Gram kernel with persistent CTAs by default ("p2p"); "autotune": its straight-line code kernel, tuned
MODES = {"p2p": {}, "nccl": {"PFFRG_EXCHANGE": "nccl"}, "sharded_upload": {"TEST_SHARDED_UPLOAD": "1"}, "autotune": {"PFFRG_AUTOTUNE": "1", "PFFRG_RPA": "code"},
         "gram": {"PFFRG_RPA": "gram"}}


@pytest.mark.parametrize("case,mode", [("su2_square_r3_nw10", "p2p"), ("xyz_honeycomb_kitaev_r3_nw10", "p2p"), ("tri_honeycomb_kg_r3_nw8", "p2p"),
                                       ("su2_kagome_r4_nw8", "nccl"), ("xyz_kagome_r4_nw8", "sharded_upload"), ("su2_kagome_r7_nw6", "autotune"),
                                       ("su2_kagome_r7_nw6", "gram"), ("su2_kagome_r7_nw6", "p2p")])
def test_sharded_flow_equals_single_gpu_flow(case, mode, tmp_path):
    from spinparser_b200.frgcore import device_count
    world = min(device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    from conftest import dumped_steps, golden
    start = [k for k in dumped_steps(golden(case)) if k > 0][0]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script), ROOT, case, "3", str(start)]
    env = dict(os.environ)
    env.update(MODES[mode])
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "MULTI_GPU_OK" in proc.stdout and "MULTI_GPU_DIVERGENCE_OK" in proc.stdout
