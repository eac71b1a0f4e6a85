"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/nsb200.h declares; the host mirror keeps the reference's names, defaults and errors."""
import ctypes
import os
import re

import pytest

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "nsb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nsb_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libnsb200.so does not export %s" % n
    assert set(capi.SYMBOLS) == set(names)


def test_params_struct_layout_matches_header():
    p = capi.Params()
    capi.lib().nsb_params_default(ctypes.byref(p))
    # defaults of the reference: no upwind / stabilisation, RAW, density 1, factor 0, defect upwind on
    assert (p.conv_upwind, p.stab, p.stab_upwind, p.diff_length) == (0, -1, 0, 0)
    assert (p.stokes, p.laplace, p.peclet_blend, p.pac_upwind, p.defect_upwind) == (0, 0, 0, 0, 1)
    assert p.density == 1.0 and p.density_set == 1 and p.kin_visc_set == 0 and p.exact_jacobian == 0.0
    assert ctypes.sizeof(capi.Params) == 14 * 4 + 7 * 8


def test_no_cpu_fallback_without_device():
    """without a GPU the product path must fail loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    disc = pkg.NavierStokesFV1("u,v,p", "Inner")
    with pytest.raises(pkg.UGError, match="no CUDA device"):
        disc._context()


def test_host_mirror_names_and_errors():
    # upwind_interface.cpp:46-61
    for name, cls in (("no", pkg.NavierStokesNoUpwind), ("FULL", pkg.NavierStokesFullUpwind),
                      (" skewed ", pkg.NavierStokesSkewedUpwind), ("lps", pkg.NavierStokesLinearProfileSkewedUpwind),
                      ("LinearProfileSkewed", pkg.NavierStokesLinearProfileSkewedUpwind),
                      ("pos", pkg.NavierStokesPositiveUpwind), ("positive", pkg.NavierStokesPositiveUpwind),
                      ("reg", pkg.NavierStokesRegularUpwind)):
        assert isinstance(pkg.CreateNavierStokesUpwind(name), cls)
    with pytest.raises(pkg.UGError):
        pkg.CreateNavierStokesUpwind("central")
    assert isinstance(pkg.CreateNavierStokesStabilization("Fields"), pkg.NavierStokesFIELDSStabilization)
    assert isinstance(pkg.CreateNavierStokesStabilization("flow"), pkg.NavierStokesFLOWStabilization)
    with pytest.raises(pkg.UGError, match="not a valid name"):
        pkg.CreateNavierStokesStabilization("supg")
    s = pkg.CreateNavierStokesStabilization("flow")
    with pytest.raises(pkg.UGError, match="Diffusion Length"):
        s.set_diffusion_length("foo")
    # lua-include.lua:36-47 factory
    assert pkg.NavierStokes("u,v,p", "Inner").disc_type() == "fv1"
    assert pkg.NavierStokes("u,v,p", "Inner", "fvcr").disc_type() == "fvcr"
    with pytest.raises(pkg.UGError, match="no disc type"):
        pkg.NavierStokes("u,v,p", "Inner", "dg")
    # set_upwind / set_stabilization auto-wiring (fv1/navier_stokes_fv1.h:190-215)
    d = pkg.NavierStokesFV1("u,v,p", "Inner")
    d.set_upwind("full")
    d.set_stabilization("fields", "cor")
    assert d.stabilization().upwind() == pkg.NavierStokesFullUpwind()
    p = d._params()
    assert (p.conv_upwind, p.stab, p.stab_upwind, p.diff_length, p.pac_upwind) == (2, 0, 2, 2, 0)
    d.set_pac_upwind(True)
    p = d._params()
    assert (p.conv_upwind, p.pac_upwind) == (2, 1)
    d2 = pkg.NavierStokesFV1("u,v,p", "Inner")
    with pytest.raises(pkg.UGError, match="Upwind must be specified previously"):
        d2.set_pac_upwind(True)
    assert d.requests_local_time_series() is True
    assert pkg.NavierStokesFVCR("u,v,p", "Inner").use_hanging() is True
