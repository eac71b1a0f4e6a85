"""the C++ shim (include/navier_stokes_b200.hpp) compiles against the C ABI and mirrors the reference's names and
throw conditions; on a GPU box its compat mode (IElemDisc slots in a ugcore-like loop) equals the fast mode."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_shim")


def _build():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    libdir = os.path.join(ROOT, "plugin_navierstokes_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"),
                           "-o", EXE, "-L", libdir, "-l:libnsb200.so", "-Wl,-rpath," + libdir])


def test_shim_names_and_errors():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_shim_compat_mode_equals_fast_mode():
    _build()
    out = subprocess.run([EXE, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
