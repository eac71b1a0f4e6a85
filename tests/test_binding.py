"""The reference-side binding (include/register_navier_stokes_b200.cpp) compiles with -DNSB_WITH_UG4 against the ugcore mock of
tests/cpp/mock_ug, registers the reference's class names / groups / constructors / overloads, and -- on a GPU box -- its
IElemDisc slots, driven through the dispatch ugcore's element loop uses, reproduce the CPU oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_binding")
GOLDEN = os.path.join(ROOT, "tests", "golden", "registry_names.json")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_registry_names as mrn  # noqa: E402

# registered through a macro in the binding (checked through the run-time dump only)
MACRO = ["NavierStokesNoUpwind", "NavierStokesFullUpwind", "NavierStokesSkewedUpwind", "NavierStokesLinearProfileSkewedUpwind",
         "NavierStokesPositiveUpwind", "NavierStokesRegularUpwind", "NavierStokesFIELDSStabilization", "NavierStokesFLOWStabilization",
         "NavierStokesFV1WithoutStabilization"]


def _build():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    libdir = os.path.join(ROOT, "plugin_navierstokes_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-DNSB_WITH_UG4", "-DNSB_UG4_MOCK", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tests", "cpp", "mock_ug"), os.path.join(ROOT, "tests", "cpp", "test_binding.cpp"),
                           "-o", EXE, "-L", libdir, "-l:libnsb200.so", "-Wl,-rpath," + libdir])


def _runtime_registry():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout[-2000:] + out.stderr
    classes, groups = {}, {}
    for line in out.stdout.splitlines():
        if line.startswith("class "):
            f = [x.strip() for x in line[6:].split("|")]
            classes[f[0]] = dict(group=f[1].split()[1], bases=int(f[2].split()[1]), ctors=int(f[3].split()[1]), smart=f[4].split()[1] == "1",
                                 methods=f[5].split()[1:])
        elif line.startswith("group "):
            _, name, grp, tag = line.split()
            groups[name] = (grp, tag)
    return classes, groups


def test_fixture_is_what_the_reference_registers():
    """the committed fixture equals a fresh extraction from the reference tree (skipped where the tree is absent)"""
    if not os.path.isdir(mrn.REF):
        pytest.skip("reference tree not present")
    assert json.load(open(GOLDEN))["classes"] == json.loads(json.dumps(mrn.reference_registry()))


def test_binding_source_registers_the_reference_signatures():
    """same extraction on the binding's own registration code: typedefs, constructor signatures, method names and overload
    signatures, class group and smart-pointer construction are identical, class by class"""
    ref = json.load(open(GOLDEN))["classes"]
    mine = mrn.extract(open(os.path.join(ROOT, "include", "register_navier_stokes_b200.cpp")).read())
    for cls in mrn.CLASSES:
        if cls in MACRO:
            continue
        assert cls in mine, cls
        for key in ("typedefs", "ctors", "methods", "group", "smart"):
            assert mine[cls][key] == ref[cls][key], (cls, key, mine[cls][key], ref[cls][key])


def test_mock_registry_run_time_names_groups_overloads():
    """InitUGPlugin_NavierStokes on the mock registry: every class of the fixture exists for 2d and 3d in the plugin's group, with the
    reference's number of bases / constructors / overloads per method name and its class group (+ dimension tag)"""
    ref = json.load(open(GOLDEN))["classes"]
    classes, groups = _runtime_registry()
    for cls, r in ref.items():
        for sfx in ("2d", "3d"):
            c = classes[cls + sfx]
            assert c["group"] == "/ug4/SpatialDisc/NavierStokes/"
            assert c["bases"] == len(r["typedefs"]) - 1 and c["ctors"] == len(r["ctors"]) and c["smart"] == r["smart"], (cls, c, r)
            # (the UG_FOR_LUA overloads are compiled only into a Lua build of UG4, like in the reference)
            want = [m[0] for m in r["methods"] if m[1] != "void(T::*)(constchar*)"]
            assert c["methods"] == want, (cls, c["methods"], want)
            assert groups[cls + sfx] == (r["group"], "dim=%s;" % sfx)
    assert len(classes) == 2 * len(ref)


@pytest.mark.gpu
def test_binding_slots_through_the_dispatch_match_the_oracle(ora, tmp_path):
    _build()
    outf = str(tmp_path / "binding_out.txt")
    out = subprocess.run([EXE, "gpu", outf], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
    tok = open(outf).read().split()
    n_elem, n_node, ndof, nnz = (int(t) for t in tok[:4])
    p = 4
    conn = np.array(tok[p:p + 4 * n_elem], dtype=np.int32).reshape(n_elem, 4); p += 4 * n_elem
    coords = np.array(tok[p:p + 2 * n_node], dtype=np.float64).reshape(n_node, 2); p += 2 * n_node
    u = np.array(tok[p:p + ndof], dtype=np.float64); p += ndof
    J = np.array(tok[p:p + nnz], dtype=np.float64); p += nnz
    d = np.array(tok[p:p + ndof], dtype=np.float64)
    from tests import parity
    rowptr, colind = ora.fv1_csr(ora.QUAD, conn, n_node)
    assert colind.size == nnz
    prm = ora.make_params(elem="quad", upwind="lps", stab="flow", diff_len="cor", exact_jac=1.0, kin_visc=0.01, source=[0.2, -0.1])
    ov, od = ora.assemble(prm, conn, coords, u, rowptr, colind, ora.JAC_A | ora.DEF_A | ora.RHS)
    eg, ee = parity.entry_errors(J, ov, rowptr)
    assert eg < parity.TOL and ee < parity.TOL, (eg, ee)
    eg, ee = parity.entry_errors(d, od)
    assert eg < parity.TOL and ee < parity.TOL, (eg, ee)
