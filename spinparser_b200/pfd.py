"""PFD1 container: a flat list of named n-d arrays.

Layout: magic ``PFD1`` then records ``{u32 name_len, name, u8 dtype, u32 ndim, u64 dims[ndim], raw data}``
with dtype codes 0=f32 1=f64 2=i32 3=i64 4=u8.  It carries problem tables (meshes, lattice tables), vertex
states and flows between the C++ host side, the reference harness under ``oracle/`` and Python.
"""
from __future__ import annotations

import struct
from typing import Dict, Mapping

import numpy as np

_DTYPES = {0: np.float32, 1: np.float64, 2: np.int32, 3: np.int64, 4: np.uint8}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def read_pfd(path: str) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != b"PFD1":
        raise ValueError(f"{path}: not a PFD1 file")
    pos = 4
    while pos < len(buf):
        (nl,) = struct.unpack_from("<I", buf, pos); pos += 4
        name = buf[pos:pos + nl].decode(); pos += nl
        code = buf[pos]; pos += 1
        (nd,) = struct.unpack_from("<I", buf, pos); pos += 4
        dims = struct.unpack_from(f"<{nd}Q", buf, pos); pos += 8 * nd
        dt = np.dtype(_DTYPES[code])
        n = int(np.prod(dims, dtype=np.int64)) if nd else 1
        arr = np.frombuffer(buf, dtype=dt, count=n, offset=pos).reshape(dims).copy()
        pos += n * dt.itemsize
        out[name] = arr
    return out


def write_pfd(path: str, arrays: Mapping[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(b"PFD1")
        for name, a in arrays.items():
            a = np.asarray(a, order="C")
            if a.dtype not in _CODES:
                raise TypeError(f"{name}: unsupported dtype {a.dtype}")
            nb = name.encode()
            f.write(struct.pack("<I", len(nb))); f.write(nb)
            f.write(struct.pack("<B", _CODES[a.dtype]))
            f.write(struct.pack("<I", a.ndim))
            f.write(struct.pack(f"<{a.ndim}Q", *a.shape))
            f.write(a.tobytes())


def as_text(a: np.ndarray) -> str:
    return bytes(a.astype(np.uint8)).decode()
