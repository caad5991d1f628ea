#!/usr/bin/env python3
"""Compiles the product's .cu sources with g++ against tests/emu/cuda_emu.h into
tests/emu/_build/libldd_emu.so -- CPU emulation of the kernels for GPU-less unit tests.
Test scaffolding only; never shipped, never loaded by the package."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "lddecode_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libldd_emu.so")


def sources():
    sys.path.insert(0, CSRC)
    import build as prod_build
    return [os.path.join(CSRC, s) for s in prod_build.SOURCES]


def build(force=False):
    srcs = sources() + [os.path.join(HERE, "cuda_emu.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
        [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "ldd_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(HERE, "_build", os.path.basename(s) + ".o")
        cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-DLDD_EMU", "-I", HERE, "-I", CSRC, "-x", "c++", "-c", s, "-o", o,
               "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable"]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("g++ failed on " + s)
        if out.strip():
            sys.stderr.write(out)
    subprocess.check_call(["g++", "-shared", "-o", OUT] + objs)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
