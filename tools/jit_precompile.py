#!/usr/bin/env python
"""Compile the lattice-specialised flow kernel of a bench workload without a GPU and leave the cubin in the kernel cache
(PFFRG_CACHE_DIR, default here: <repo>/.jitcache, which travels to the GPU box with the gpurun snapshot).

    [PFFRG_JIT_ACC=.. PFFRG_CLUSTER=.. ...] python tools/jit_precompile.py <workload> [...]

`torch` is imported first on purpose: bench.py loads libpffrg after torch, whose bundled libnvrtc.so.12 then satisfies the
library's NVRTC dependency; the cache key contains the NVRTC version.
"""
import os
import sys
import time

import torch  # noqa: F401  (load order, see above)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PFFRG_CACHE_DIR", os.path.join(ROOT, ".jitcache"))
from spinparser_b200 import ProblemTables, read_pfd  # noqa: E402
from spinparser_b200.frgcore import jit_compile_check  # noqa: E402

for workload in sys.argv[1:]:
    d = read_pfd(os.path.join(ROOT, "bench_data", workload + ".tables.pfd"))
    t0 = time.time()
    size = jit_compile_check(bytes(d["core"]).decode(), ProblemTables.from_pfd(d))
    print(f"{workload}: {size} bytes of cubin in {time.time() - t0:.1f} s -> {os.environ['PFFRG_CACHE_DIR']}")
