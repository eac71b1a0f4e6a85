#!/usr/bin/env python
"""Builds libnsb200.so (CUDA, sm_100a) in-tree and the CPU oracle (test infrastructure)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "plugin_navierstokes_b200", "csrc")
OUT = os.path.join(ROOT, "plugin_navierstokes_b200", "libnsb200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "nsb200.h"))
    if not force and not _newer(OUT, deps):
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, os.path.join(CSRC, "nsb200.cu")]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


def build_oracle(force=False):
    sys.path.insert(0, ROOT)
    from oracle import oracle
    return oracle.build(force)


if __name__ == "__main__":
    force = "--force" in sys.argv
    build_cuda(force, "-v" in sys.argv)
    build_oracle(force)
    print("ok")
