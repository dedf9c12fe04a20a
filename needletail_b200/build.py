#!/usr/bin/env python3
"""Builds needletail_b200/libntgpu.so (sm_100a only) with nvcc — no torch, no setuptools.

    python needletail_b200/build.py [--force] [--verbose]
"""
import os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "ntgpu.cu")
OUT = os.path.join(HERE, "libntgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-ldl", "-lz"]


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(HERE, "..", "include", "ntgpu.h")]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(s) <= t for s in sources())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
