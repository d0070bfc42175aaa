import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def golden(case: str, kind: str = "f64"):
    from spinparser_b200.pfd import read_pfd
    return read_pfd(os.path.join(GOLDEN, f"{case}.{kind}.pfd"))


def dumped_steps(d):
    return sorted({int(k.split("/")[0][4:]) for k in d if k.startswith("step") and k.endswith("/flow/v2")})


def assert_parity(got: np.ndarray, want: np.ndarray, what: str, rel: float = 1e-10, floor: float = 1e-12):
    """Parity criterion of SURVEY.md section 0.6: |d_i| <= rel*|x_i| + floor*max_j|x_j| per channel array.
    (1e-10 relative per vertex entry; the absolute floor is needed because symmetry-forbidden entries exist only as
    round-off in the reference itself.)"""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert np.isnan(got).sum() == np.isnan(want).sum(), f"{what}: NaN pattern differs"
    scale = np.nanmax(np.abs(want)) if want.size else 0.0
    err = np.abs(got - want)
    tol = rel * np.abs(want) + floor * scale
    bad = err > tol
    if bad.any():
        i = int(np.argmax(err - tol))
        raise AssertionError(f"{what}: {int(bad.sum())}/{got.size} entries out of tolerance; worst at {i}: got {got.flat[i]!r} want {want.flat[i]!r} (scale {scale:.3e})")
