#!/usr/bin/env python
"""Bridge to the reference's OWN golden vectors (test/scripted/assets/test_reference{1,2,3}.ref, tolerance 1e-5 in
test/scripted/assets/test_eval.py:14): parses the three HDF5 files with the dependency-free reader tests/hdf5_v0.py and writes
tests/golden/reference_goldens.pfd = {<run>/ref/<dataset path>: values} plus, per (golden task, core), the problem tables the
flow needs (frequency mesh, cutoff grid, symmetry-reduced lattice tables, bare couplings), produced by running the reference's own
task-file parser and lattice builder (oracle/_ref/oracle64) on the task files of test/scripted/test_reference{1,2,3}.sh with the
reference's resource files where they lie. The fixture travels to the GPU box; /root/reference does not.

    python tests/golden/make_reference_goldens.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hdf5_v0 import read_hdf5  # noqa: E402
from spinparser_b200.pfd import read_pfd, write_pfd  # noqa: E402

REFERENCE = "/root/reference"
FREQUENCIES = ["0.31812", "0.36329", "0.41812", "0.46329", "0.51334", "0.56880", "0.63024", "0.69833", "0.77378", "0.85737", "0.95", "1.0", "3.0", "10.0"]
# (golden file, lattice, model, couplings, cores; the first core is the one the .ref file was written by)
TASKS = {
    "ref1": ("test_reference1.ref", "hyperkagome", "hyperkagome-heisenberg", {"j": "1.0"}, ["SU2", "XYZ", "TRI"]),
    "ref2": ("test_reference2.ref", "hyperhoneycomb", "hyperhoneycomb-kitaev", {"j": "0.1", "k": "-1.0"}, ["XYZ", "TRI"]),
    "ref3": ("test_reference3.ref", "honeycomb", "honeycomb-kitaev-gamma", {"j": "0.2", "k": "1.0", "g": "-0.1"}, ["TRI"]),
}


def task_xml(lattice, model, couplings, core):
    """The task file of test/scripted/test_reference{1,2,3}.sh for one core."""
    values = "\n".join(f"            <value>{v}</value>" for v in FREQUENCIES)
    params = "\n".join(f"            <{k}>{v}</{k}>" for k, v in couplings.items())
    return f"""<?xml version="1.0" encoding="utf-8"?>
<task>
    <parameters>
        <frequency discretization="manual">
{values}
        </frequency>
        <cutoff discretization="exponential">
            <max>10</max>
            <min>0.3</min>
            <step>0.9</step>
        </cutoff>
        <lattice name="{lattice}" range="3"/>
        <model name="{model}" symmetry="{core}">
{params}
        </model>
    </parameters>
    <measurements>
        <measurement name="correlation" />
    </measurements>
</task>
"""


def run_oracle(binary, xml_text, extra=()):
    with tempfile.TemporaryDirectory() as tmp:
        task = os.path.join(tmp, "task.xml")
        with open(task, "w") as f:
            f.write(xml_text)
        out = os.path.join(tmp, "out.pfd")
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", binary), "-r", os.path.join(REFERENCE, "res"), task, "--out", out, *extra], check=True, cwd=tmp, capture_output=True)
        return read_pfd(out)


def main():
    fixture = {}
    for run, (ref_file, lattice, model, couplings, cores) in TASKS.items():
        for path, values in read_hdf5(os.path.join(REFERENCE, "test", "scripted", "assets", ref_file)).items():
            fixture[f"{run}/ref{path}"] = values
        for core in cores:
            d = run_oracle("oracle64", task_xml(lattice, model, couplings, core), ("--max-steps", "0", "--no-measure"))
            L = int(d["lattice/size"])
            n = 16 * L if core == "TRI" else L
            for k, v in d.items():
                if (k.startswith("lattice/") and k not in ("lattice/positions", "lattice/parameters")) or k in ("core", "frequency", "cutoff", "spinLength", "normalization"):
                    fixture[f"{run}/{core}/{k}"] = v
            fixture[f"{run}/{core}/bare"] = np.stack([d[k][:n] for k in sorted(k for k in d if k.startswith("initial/v4_"))])
            print(run, core, "L", L, "cutoffs", len(d["cutoff"]))
    path = os.path.join(HERE, "reference_goldens.pfd")
    write_pfd(path, fixture)
    print(path, os.path.getsize(path) // 1024, "KiB,", len(fixture), "records")


if __name__ == "__main__":
    main()
