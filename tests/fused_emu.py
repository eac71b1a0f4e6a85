"""ctypes driver of tests/cpp/emu_fused.cpp: the CPU emulation of the fused patch kernel (TEST INFRASTRUCTURE ONLY).

The emulator compiles the NSB_HD lane functions of plugin_navierstokes_b200/csrc/ns_fused.cuh and the host-side patch
builder (ns_patch.h, ns_graph.h) with g++ and runs them thread by thread. Nothing in the package imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "cpp", "emu_fused.cpp")
_LIB = os.path.join(_HERE, "cpp", "libemu_fused.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "plugin_navierstokes_b200", "csrc")

UPW = {None: 0, "none": 0, "no": 1, "full": 2, "skewed": 3, "lps": 4, "positive": 5}
STAB = {"fields": 0, "flow": 1, "none": 2}
DIFF = {"raw": 0, "fivepoint": 1, "cor": 2}
ELEM = {"tri": 0, "quad": 1, "tet": 2, "hex": 3}


class KParams(C.Structure):
    # mirrors nsb::KParams (ns_base.h)
    _fields_ = [("upw_stab", C.c_int), ("upw_conv", C.c_int), ("stab", C.c_int), ("diff_len", C.c_int),
                ("stokes", C.c_int), ("laplace", C.c_int), ("peclet", C.c_int), ("pac", C.c_int), ("time_dep", C.c_int),
                ("has_source", C.c_int), ("what", C.c_int), ("defect_upwind", C.c_int),
                ("exact_jac", C.c_double), ("visc", C.c_double), ("rho", C.c_double), ("inv_rho", C.c_double),
                ("dt", C.c_double), ("scale_a", C.c_double), ("scale_m", C.c_double), ("grad_div", C.c_double),
                ("src", C.c_double * 3)]


def build(force=False):
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".h", ".cuh"))]
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= max(os.path.getmtime(f) for f in deps):
        return _LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", _SRC, "-o", _LIB])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        assert L.emu_kparams_size() == C.sizeof(KParams)
        _lib = L
    return _lib


def assemble(elem, conn, coords, u, what, upwind="full", stab="fields", diff="raw", visc=1e-2, density=1.0, stokes=False,
             laplace=False, peclet=False, source=None, stab_upwind=None, sol0=None, sol1=None, dt=0.0, scale_a=1.0, scale_m=1.0,
             beta=0.0, values=None, defect=None, nnz=None, ray_fast=1, use_geo=1):
    """returns (values, defect, stats) of the emulated fused kernel; stats = dict of patch statistics"""
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
    k = KParams()
    k.upw_conv = UPW[upwind]
    k.upw_stab = UPW[stab_upwind] if stab_upwind else UPW[upwind]
    k.stab, k.diff_len = STAB[stab], DIFF[diff]
    k.stokes, k.laplace, k.peclet = int(stokes), int(laplace), int(peclet)
    k.time_dep = int(sol0 is not None)
    k.what = what
    k.visc, k.rho, k.inv_rho, k.dt, k.scale_a, k.scale_m = visc, density, 1.0 / density, dt, scale_a, scale_m
    if source is not None:
        k.has_source = 1
        for d, v in enumerate(source):
            k.src[d] = v
    n_dof = u.shape[0]
    if values is None:
        values = np.full(nnz, np.nan)
    if defect is None:
        defect = np.full(n_dof, np.nan)
    s0 = None if sol0 is None else np.ascontiguousarray(sol0, dtype=np.float64).reshape(-1)
    s1 = None if sol1 is None else np.ascontiguousarray(sol1, dtype=np.float64).reshape(-1)
    stats = np.zeros(8, dtype=np.int64)
    err = C.create_string_buffer(512)
    dp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    f = lib().emu_fused_assemble
    f.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(KParams), C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
    rc = f(ELEM[elem], conn.shape[0], coords.shape[0], dp(conn), dp(coords), C.byref(k), dp(u), dp(s0), dp(s1), beta,
           dp(values), dp(defect), ray_fast, use_geo, dp(stats), err, 512)
    if rc != 0:
        raise RuntimeError("emu_fused_assemble: %s" % err.value.decode())
    names = ["n_patch", "scvf_evals", "n_scvf", "max_nodes", "max_work", "max_elems", "ray_fast", "n_not_star_shaped"]
    return values, defect, dict(zip(names, stats.tolist()))
