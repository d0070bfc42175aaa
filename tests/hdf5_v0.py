"""Dependency-free reader for the HDF5 files SpinParser writes (TEST INFRASTRUCTURE; h5py is not in the image).

Covers exactly what `.obs` / `.ref` files contain: superblock version 0 with 8-byte offsets and lengths,
old-style groups (object header v1 -> symbol-table message 0x0011 {B-tree address, local heap address}; B-tree `TREE` nodes ->
`SNOD` leaves; names in the local heap's data segment), datasets with a simple dataspace (message 0x0001), a fixed-point-free
datatype (message 0x0003: class 1 = IEEE float, class 10 = array of floats), contiguous layout (message 0x0008 version 3), and
version-1 attribute messages (0x000C) holding the float attribute `cutoff`; header continuation blocks (0x0010) are followed.

    read_hdf5(path) -> {"/group/.../dataset": ndarray, "/group/...@attribute": ndarray}
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _File:
    def __init__(self, data: bytes):
        self.b = data
        if data[:8] != b"\x89HDF\r\n\x1a\n" or data[8] != 0 or data[13] != 8 or data[14] != 8:
            raise ValueError("not an HDF5 file with a version-0 superblock and 8-byte offsets")
        self.root_header = struct.unpack_from("<Q", data, 24 + 4 * 8 + 8)[0]  # root symbol-table entry: link name offset, object header address

    # ---- object headers (version 1) ---------------------------------------------------------------------------------
    def messages(self, addr: int) -> List[Tuple[int, bytes]]:
        version, _, count, _refs, size = struct.unpack_from("<BBHII", self.b, addr)
        if version != 1:
            raise ValueError(f"object header version {version} at {addr}")
        out: List[Tuple[int, bytes]] = []
        blocks = [(addr + 16, size)]  # the first message starts on an 8-byte boundary after the 12-byte prefix
        while blocks and len(out) < count:
            pos, length = blocks.pop(0)
            end = pos + length
            while pos + 8 <= end and len(out) < count:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.b, pos)
                body = self.b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr, clen))
                out.append((mtype, body))
        return out

    # ---- groups ----------------------------------------------------------------------------------------------------
    def group_entries(self, btree: int, heap: int) -> Dict[str, int]:
        if self.b[heap:heap + 4] != b"HEAP":
            raise ValueError("local heap signature")
        heap_data = struct.unpack_from("<Q", self.b, heap + 8 + 16)[0]
        entries: Dict[str, int] = {}

        def name_at(offset: int) -> str:
            start = heap_data + offset
            return self.b[start:self.b.index(b"\0", start)].decode()

        def walk(node: int) -> None:
            sig = self.b[node:node + 4]
            if sig == b"TREE":
                _ntype, level, used = struct.unpack_from("<BBH", self.b, node + 4)
                pos = node + 8 + 16  # left / right sibling addresses
                for i in range(used):
                    child = struct.unpack_from("<Q", self.b, pos + 8 + i * 16)[0]  # key, child, key, child, ..., key
                    walk(child)
            elif sig == b"SNOD":
                _version, _, count = struct.unpack_from("<BBH", self.b, node + 4)
                for i in range(count):
                    name_offset, header = struct.unpack_from("<QQ", self.b, node + 8 + i * 40)
                    entries[name_at(name_offset)] = header
            else:
                raise ValueError(f"unexpected node signature {sig!r}")

        walk(btree)
        return entries

    # ---- datatypes / dataspaces ---------------------------------------------------------------------------------------
    @staticmethod
    def dtype_of(body: bytes):
        cls = body[0] & 0x0F
        size = struct.unpack_from("<I", body, 4)[0]
        if cls == 1:  # IEEE floating point, little endian
            return np.dtype("<f%d" % size), ()
        if cls == 10:  # array: version 2 or 3; dimensionality, (reserved), dimension sizes, (permutation indices), base type
            version = body[0] >> 4
            ndim = body[8]
            pos = 12 if version == 2 else 9
            dims = struct.unpack_from("<%dI" % ndim, body, pos)
            pos += 4 * ndim + (4 * ndim if version == 2 else 0)
            base, _ = _File.dtype_of(body[pos:])
            return base, tuple(dims)
        raise ValueError(f"datatype class {cls}")

    @staticmethod
    def shape_of(body: bytes):
        version, ndim, flags = body[0], body[1], body[2]
        pos = 8 if version == 1 else 4
        return tuple(struct.unpack_from("<%dQ" % ndim, body, pos))

    # ---- traversal ------------------------------------------------------------------------------------------------------
    def read(self, header: int, path: str, out: Dict[str, np.ndarray]) -> None:
        msgs = self.messages(header)
        kinds = {t for t, _ in msgs}
        for mtype, body in msgs:
            if mtype == 0x000C:  # attribute, version 1: name, datatype and dataspace each padded to 8 bytes
                _version, _, name_size, type_size, space_size = struct.unpack_from("<BBHHH", body, 0)
                pad = lambda n: (n + 7) // 8 * 8
                pos = 8
                name = body[pos:pos + name_size].split(b"\0")[0].decode(); pos += pad(name_size)
                dtype, inner = self.dtype_of(body[pos:pos + type_size]); pos += pad(type_size)
                shape = self.shape_of(body[pos:pos + space_size]); pos += pad(space_size)
                n = int(np.prod(shape + inner)) if shape + inner else 1
                out[f"{path}@{name}"] = np.frombuffer(body, dtype=dtype, count=n, offset=pos).reshape(shape + inner).copy()
        if 0x0011 in kinds:  # group
            btree, heap = struct.unpack_from("<QQ", dict(msgs)[0x0011], 0)
            for name, child in sorted(self.group_entries(btree, heap).items()):
                self.read(child, f"{path}/{name}", out)
        elif 0x0008 in kinds:  # dataset
            table = dict(msgs)
            dtype, inner = self.dtype_of(table[0x0003])
            shape = self.shape_of(table[0x0001])
            layout = table[0x0008]
            if layout[0] != 3 or layout[1] != 1:
                raise ValueError(f"{path}: only contiguous version-3 layouts are supported")
            addr, size = struct.unpack_from("<QQ", layout, 2)
            n = int(np.prod(shape + inner)) if shape + inner else 1
            if addr == UNDEF:
                out[path] = np.zeros(shape + inner, dtype=dtype)
            else:
                out[path] = np.frombuffer(self.b, dtype=dtype, count=n, offset=addr).reshape(shape + inner).copy()


def read_hdf5(path: str) -> Dict[str, np.ndarray]:
    with open(path, "rb") as f:
        data = f.read()
    h = _File(data)
    out: Dict[str, np.ndarray] = {}
    h.read(h.root_header, "", out)
    return out
