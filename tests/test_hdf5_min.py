"""`.obs` / `.checkpoint` files WITHOUT libhdf5: spinparser_b200/host/hdf5_min.hpp serves the HDF5 calls of the
reference's writers and readers and stores the files in the on-disk format of the reference's own golden files (superblock
version 0, old-style groups, contiguous float datasets, version-1 attributes).

* The UNMODIFIED reference (oracle/_ref/oracle32_pin, FP32 as shipped) run with H5MIN_DISK=1 writes `<task>.obs` files that the
  independent, dependency-free reader tests/hdf5_v0.py (validated on the golden files, tests/test_reference_goldens.py) parses, and
  every dataset, attribute and meta array equals the golden file test/scripted/assets/test_reference{1,2,3}.ref BIT FOR BIT.
* The object headers it writes carry the same messages with the same bodies as the ones libhdf5 wrote into the golden files
  (addresses and time stamps aside).
* A checkpoint written by one process is read by another one (`--resume`, src/SpinParser.cpp:131-135) and the flow continues to the
  same final state as an uninterrupted run.
"""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

REFERENCE = "/root/reference"
ASSETS = os.path.join(REFERENCE, "test", "scripted", "assets")
PIN = os.path.join(ROOT, "oracle", "_ref", "oracle32_pin")

pytestmark = pytest.mark.skipif(not (os.path.isdir(ASSETS) and os.path.exists(PIN)), reason="needs the reference tree and oracle/_ref/oracle32_pin")


def _run(tmp_path, xml_text, *extra, name="task"):
    task = tmp_path / f"{name}.xml"
    task.write_text(xml_text)
    out = tmp_path / f"{name}.pfd"
    env = dict(os.environ, H5MIN_DISK="1")
    subprocess.run([PIN, "-r", os.path.join(REFERENCE, "res"), str(task), "--out", str(out), "--no-lattice", *extra], check=True, cwd=tmp_path, capture_output=True, env=env)
    return out


def _gen():
    sys.path.insert(0, GOLDEN)
    import make_reference_goldens as gen
    return gen


@pytest.mark.parametrize("run", ["ref1", "ref2", "ref3"])
def test_obs_file_written_without_libhdf5_equals_the_golden_file(run, tmp_path):
    from hdf5_v0 import read_hdf5
    gen = _gen()
    ref_file, lattice, model, couplings, cores = gen.TASKS[run]
    _run(tmp_path, gen.task_xml(lattice, model, couplings, cores[0]))
    obs = tmp_path / "task.obs"
    assert obs.exists() and obs.read_bytes()[:8] == b"\x89HDF\r\n\x1a\n"
    mine = read_hdf5(str(obs))
    golden = read_hdf5(os.path.join(ASSETS, ref_file))
    assert sorted(mine) == sorted(golden)
    assert sum(1 for k in golden if k.endswith("/data") and "/measurement_" in k) in (68, 136, 340)
    for path, values in golden.items():
        assert mine[path].dtype == values.dtype and mine[path].shape == values.shape, path
        assert np.array_equal(mine[path], values), path


def test_object_headers_carry_the_messages_libhdf5_writes(tmp_path):
    from hdf5_v0 import _File
    gen = _gen()
    ref_file, lattice, model, couplings, cores = gen.TASKS["ref1"]
    _run(tmp_path, gen.task_xml(lattice, model, couplings, cores[0]))
    files = [_File(open(p, "rb").read()) for p in (tmp_path / "task.obs", os.path.join(ASSETS, ref_file))]
    # superblock: version 0, 8-byte offsets and lengths, the library's default B-tree ranks, no free-space / driver info
    for f in files:
        assert f.b[8:24] == files[1].b[8:24]
        assert struct.unpack_from("<Q", f.b, 40)[0] == len(f.b)  # end-of-file address

    def header_of(f, path):
        addr = f.root_header
        for name in [p for p in path.split("/") if p]:
            msgs = dict(f.messages(addr))
            addr = f.group_entries(*struct.unpack_from("<QQ", msgs[0x0011], 0))[name]
        return addr

    def messages(f, path):
        return [(t, b) for t, b in f.messages(header_of(f, path)) if t not in (0x0010, 0x0000)]  # continuation blocks / padding are layout, not content

    for path in ("/SU2CorZZ/data/measurement_3/data", "/SU2CorDD/meta/basis", "/SU2CorDD/meta/sites", "/SU2CorZZ/meta/latticeVectors"):
        mine, theirs = messages(files[0], path), messages(files[1], path)
        assert [t for t, _ in mine] == [t for t, _ in theirs] == [0x0001, 0x0003, 0x0005, 0x0008, 0x0012]
        for (t, a), (_, b) in zip(mine, theirs):
            if t == 0x0008:
                assert a[:2] == b[:2] and a[10:18] == b[10:18]  # version 3, contiguous, the same size
            elif t == 0x0012:
                assert a[:4] == b[:4]
            else:
                assert a == b, (path, hex(t))
    # a measurement group: symbol table + the cutoff attribute, attribute message identical
    mine, theirs = messages(files[0], "/SU2CorZZ/data/measurement_3"), messages(files[1], "/SU2CorZZ/data/measurement_3")
    assert sorted(t for t, _ in mine) == sorted(t for t, _ in theirs) == [0x000C, 0x0011]
    assert dict(mine)[0x000C] == dict(theirs)[0x000C]


def test_checkpoint_written_by_one_process_resumes_in_another(tmp_path):
    from spinparser_b200.pfd import read_pfd
    gen = _gen()
    ref_file, lattice, model, couplings, cores = gen.TASKS["ref1"]
    xml = gen.task_xml(lattice, model, couplings, cores[0])
    full = read_pfd(str(_run(tmp_path, xml, "--no-measure", name="full")))
    # first process: 12 steps, a checkpoint after every step; second process: --resume
    _run(tmp_path, xml, "--no-measure", "--max-steps", "12", "--checkpoint-time", "-1", name="part")
    checkpoint = tmp_path / "part.checkpoint"
    assert checkpoint.exists() and checkpoint.read_bytes()[:8] == b"\x89HDF\r\n\x1a\n"
    saved = tmp_path / "saved.checkpoint"  # (a fresh start of the reference's task-file parser removes <task>.checkpoint, src/TaskFileParser.cpp:74-100)
    checkpoint.rename(saved)
    resumed = read_pfd(str(_run(tmp_path, xml, "--no-measure", "--resume", str(saved), name="part")))
    assert int(resumed["resumedFromStep"]) == 12
    assert int(resumed["finalStep"]) == int(full["finalStep"])
    for key in [k for k in full if k.startswith("final/")]:
        assert np.array_equal(resumed[key], full[key]), key


def test_group_btree_keys_follow_the_library_convention(tmp_path):
    """A group with more entries than one symbol-table node holds (34 measurements): as in the golden file, every child of the
    B-tree node is a sorted SNOD of at most 2 K = 8 entries, key i + 1 is the heap offset of the LAST name of child i and key 0 the
    empty name."""
    from hdf5_v0 import _File
    gen = _gen()
    ref_file, lattice, model, couplings, cores = gen.TASKS["ref1"]
    _run(tmp_path, gen.task_xml(lattice, model, couplings, cores[0]))
    for path in (tmp_path / "task.obs", os.path.join(ASSETS, ref_file)):
        f = _File(open(path, "rb").read())
        b = f.b
        addr = f.root_header
        for name in ("SU2CorZZ", "data"):
            addr = f.group_entries(*struct.unpack_from("<QQ", dict(f.messages(addr))[0x0011], 0))[name]
        btree, heap = struct.unpack_from("<QQ", dict(f.messages(addr))[0x0011], 0)
        heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
        name_at = lambda off: b[heap_data + off:b.index(b"\0", heap_data + off)].decode()
        assert b[btree:btree + 4] == b"TREE" and b[btree + 4] == 0 and b[btree + 5] == 0
        used = struct.unpack_from("<H", b, btree + 6)[0]
        assert struct.unpack_from("<QQ", b, btree + 8) == (2 ** 64 - 1, 2 ** 64 - 1)
        total, previous = 0, ""
        for i in range(used):
            key0, child, key1 = struct.unpack_from("<QQQ", b, btree + 24 + 16 * i)
            assert b[child:child + 4] == b"SNOD"
            count = struct.unpack_from("<H", b, child + 6)[0]
            names = [name_at(struct.unpack_from("<Q", b, child + 8 + 40 * k)[0]) for k in range(count)]
            assert 1 <= count <= 8 and names == sorted(names)
            assert name_at(key0) == previous and name_at(key1) == names[-1] and previous < names[0]
            previous = names[-1]
            total += count
        assert total == 34


@pytest.mark.parametrize("real", ["float", "double"])
def test_api_round_trip_without_reference_code(real, tmp_path):
    """The layer on its own (tests/cpp/hdf5_min_roundtrip.cpp): a group of 300 children (38 symbol-table nodes under two levels of B-tree
    nodes), an empty group, an array-typed dataset, attributes -- written, re-read from disk by the layer's own reader, extended,
    written again, and read by the independent Python reader. H5MIN_REAL = float and double (the FP64 test build of the reference)."""
    import shutil
    from hdf5_v0 import read_hdf5
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = tmp_path / "roundtrip"
    subprocess.run([gxx, "-std=c++17", "-O1", f"-DH5MIN_REAL={real}", "-I", os.path.join(ROOT, "spinparser_b200", "host"),
                    os.path.join(ROOT, "tests", "cpp", "hdf5_min_roundtrip.cpp"), "-o", str(exe)], check=True, capture_output=True)
    path = tmp_path / "roundtrip.obs"
    out = subprocess.run([str(exe), str(path)], check=True, capture_output=True, text=True).stdout.split()
    assert [float(x) for x in out[:2]] == [299.0, 299.0 + 0.125 * 9] and abs(float(out[2]) - 50.0 / 18) < 1e-4  # (printed with %g)
    assert out[3:] == ["300", "measurement_0", "0", "1", "0"]
    d = read_hdf5(str(path))
    dtype = np.float32 if real == "float" else np.float64
    assert len([k for k in d if k.endswith("/data")]) == 300 and len([k for k in d if k.endswith("@cutoff")]) == 300
    assert d["/Cor/meta/basis"].dtype == dtype and np.array_equal(d["/Cor/meta/basis"], np.array([[0, 0, 0], [0.5, 0.25, 1.5]], dtype=dtype))
    for k in (0, 7, 150, 299):
        assert np.array_equal(d[f"/Cor/data/measurement_{k}/data"], (k + 0.125 * np.arange(10, dtype=np.float64)).astype(dtype).reshape(2, 5))
        assert d[f"/Cor/data/measurement_{k}@cutoff"][0] == dtype(50) / dtype(k + 1)
    # two levels of B-tree nodes in the big group; the groups added before and after the re-read are there
    from hdf5_v0 import _File
    f = _File(open(path, "rb").read())
    root = f.group_entries(*struct.unpack_from("<QQ", dict(f.messages(f.root_header))[0x0011], 0))
    assert sorted(root) == ["Cor", "empty", "extra"]
    cor = f.group_entries(*struct.unpack_from("<QQ", dict(f.messages(root["Cor"]))[0x0011], 0))
    btree = struct.unpack_from("<Q", dict(f.messages(cor["data"]))[0x0011], 0)[0]
    assert f.b[btree:btree + 4] == b"TREE" and f.b[btree + 5] == 1 and struct.unpack_from("<H", f.b, btree + 6)[0] == 2
