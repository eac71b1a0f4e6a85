#!/usr/bin/env python
"""Builds libnsb200.so (CUDA, sm_100a) in-tree and the CPU oracle (test infrastructure).

The kernels of each element type live in their own translation unit (csrc/*_inst.cu compiled with
-DNSB_ELEM=e) so that nvcc runs in parallel; objects go to build/ (git-ignored)."""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "plugin_navierstokes_b200", "csrc")
OBJ = os.path.join(ROOT, "build")
OUT = os.path.join(ROOT, "plugin_navierstokes_b200", "libnsb200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]

# (source, extra defines, object name)
UNITS = [("nsb200.cu", [], "nsb200.o")]
UNITS += [("fv1_inst.cu", ["-DNSB_ELEM=%d" % e], "fv1_e%d.o" % e) for e in range(4)]
UNITS += [("fused_inst.cu", ["-DNSB_ELEM=%d" % e], "fused_e%d.o" % e) for e in range(4)]
UNITS += [("tile_inst.cu", ["-DNSB_ELEM=%d" % e], "tile_e%d.o" % e) for e in (2, 3)]
UNITS += [("dense_inst.cu", ["-DNSB_ELEM=%d" % e], "dense_e%d.o" % e) for e in range(5)]
UNITS += [("prism_inst.cu", [], "prism_e4.o")]
UNITS += [("fvcrq_inst.cu", ["-DNSB_ELEM=%d" % e], "fvcrq_e%d.o" % e) for e in (1, 3)]
UNITS += [("fvcr_inst.cu", ["-DNSB_ELEM=%d" % e], "fvcr_e%d.o" % e) for e in (0, 2)]


def _sources_digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "nsb200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_cuda(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _sources_digest()
    if (not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest):
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_one(unit):
        src, defs, obj = unit
        cmd = [nvcc] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return unit, r

    units = [u for u in UNITS if os.path.exists(os.path.join(CSRC, u[0]))]
    with cf.ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        for unit, r in ex.map(compile_one, units):
            if verbose or r.returncode:
                sys.stderr.write("== %s %s\n%s%s" % (unit[0], " ".join(unit[1]), r.stdout, r.stderr))
            if r.returncode:
                raise RuntimeError("nvcc failed on %s" % unit[0])
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + [os.path.join(OBJ, u[2]) for u in units]
    subprocess.check_call(cmd)
    open(stamp, "w").write(digest)
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
