#!/usr/bin/env python
"""Turn the kernel headers into string literals (pffrg_embedded.inc) for run-time compilation with NVRTC."""
import sys

def literal(name, path):
    text = open(path).read()
    out = [f"static const char {name}[] ="]
    # split into chunks: some compilers limit the length of a single literal
    chunk = []
    size = 0
    for line in text.splitlines(keepends=True):
        chunk.append(line); size += len(line)
        if size > 8000:
            out.append('R"PFFRGSRC(' + "".join(chunk) + ')PFFRGSRC"'); chunk, size = [], 0
    out.append('R"PFFRGSRC(' + "".join(chunk) + ')PFFRGSRC";')
    return "\n".join(out) + "\n"

with open(sys.argv[1], "w") as f:
    f.write(literal("kDeviceSource", sys.argv[2]))
    f.write(literal("kKernelSource", sys.argv[3]))
