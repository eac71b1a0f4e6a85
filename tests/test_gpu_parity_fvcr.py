"""GPU parity of the FVCR (Crouzeix-Raviart) path against the CPU oracle (through the C ABI)."""
import numpy as np
import pytest

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen
from tests import parity
from tests.parity import TOL

pytestmark = pytest.mark.gpu
FCTS = {2: "u,v,p", 3: "u,v,w,p"}
MODES = {"colored": capi.SCATTER_COLORED, "atomic": capi.SCATTER_ATOMIC, "gather": capi.SCATTER_GATHER}


def _case(elem, n, seed=0):
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=seed)
    es, n_side = meshgen.element_sides(elem, conn)
    dim = coords.shape[1]
    rng = np.random.default_rng(seed + 7)
    # side velocities from a smooth field + noise, element pressures
    cen = np.zeros((n_side, dim))
    cnt = np.zeros(n_side)
    sides = meshgen.SIDES[elem]
    for k, s in enumerate(sides):
        np.add.at(cen, es[:, k], coords[conn[:, list(s)]].mean(axis=1))
        np.add.at(cnt, es[:, k], 1)
    cen /= cnt[:, None]
    vel = np.stack([np.sin(2 * cen[:, 0]) + 0.3, np.cos(3 * cen[:, 1]) - 0.2] + ([0.5 * cen[:, 0] * cen[:, 2]] if dim == 3 else []), axis=1)
    vel += 0.05 * rng.uniform(-1, 1, vel.shape)
    u = np.concatenate([vel.ravel(), rng.uniform(-1, 1, conn.shape[0])])
    return coords, conn, es, n_side, u


@pytest.mark.parametrize("elem,n", [("tri", 7), ("tet", 3), ("quad", 6), ("hex", 3)])
@pytest.mark.parametrize("mode", ["colored", "atomic", "gather"])
@pytest.mark.parametrize("upwind", ["no", "full", "skewed", "lps"])
@pytest.mark.parametrize("flags", [dict(), dict(peclet=True, exact=1.0), dict(laplace=True, grad_div=0.3), dict(defect_upwind=False, exact=0.5),
                                   dict(stokes=True)], ids=lambda f: "-".join("%s=%s" % kv for kv in f.items()) or "default")
def test_fvcr_jac_def(ora, elem, n, mode, upwind, flags):
    coords, conn, es, n_side, u = _case(elem, n)
    dim = coords.shape[1]
    E = ora.ELEM[elem]
    disc = pkg.NavierStokesFVCR(FCTS[dim], "Inner")
    disc.set_kinematic_viscosity(0.02)
    disc.set_density(1.3)
    disc.set_upwind(upwind)
    disc.set_peclet_blend(flags.get("peclet", False))
    disc.set_exact_jacobian(flags.get("exact", 0.0))
    disc.set_laplace(flags.get("laplace", False))
    disc.set_grad_div(flags.get("grad_div", 0.0))
    disc.set_defect_upwind(flags.get("defect_upwind", True))
    disc.set_stokes(flags.get("stokes", False))
    disc.set_grid(elem, conn, coords, es, n_side)
    disc.prep_elem_loop()
    rp, ci = disc.csr()
    rowptr, colind = ora.fvcr_csr(E, es, n_side)
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colind)
    p = ora.make_params(disc="fvcr", elem=elem, upwind=upwind, kin_visc=0.02, density=1.3, peclet_blend=flags.get("peclet", False),
                        exact_jac=flags.get("exact", 0.0), laplace=flags.get("laplace", False), grad_div=flags.get("grad_div", 0.0),
                        defect_upwind=flags.get("defect_upwind", True), stokes=flags.get("stokes", False))
    what = capi.JAC_A | capi.DEF_A
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, what, elem_sides=es, n_side=n_side)
    gv, gd = disc.assemble(what, u, scatter_mode=MODES[mode])
    eg, ee = parity.entry_errors(gv, ov, rowptr)
    assert eg < TOL and ee < TOL, ("jacobian", eg, ee)
    eg, ee = parity.entry_errors(gd, od)
    assert eg < TOL and ee < TOL, ("defect", eg, ee)


@pytest.mark.parametrize("elem,n", [("tri", 6), ("tet", 3), ("quad", 5), ("hex", 3)])
def test_fvcr_mass_rhs_and_scales(ora, elem, n):
    """config 4: instationary parts on the side SCVs; rhs without density (fvcr/navier_stokes_fvcr.cpp:757)"""
    coords, conn, es, n_side, u = _case(elem, n, seed=3)
    dim = coords.shape[1]
    E = ora.ELEM[elem]
    src = [0.3, -0.2, 0.1][:dim]
    disc = pkg.NavierStokesFVCR(FCTS[dim], "Inner")
    disc.set_kinematic_viscosity(1e-3)
    disc.set_density(1.2)
    disc.set_upwind("full")
    disc.set_source(src)
    disc.set_grid(elem, conn, coords, es, n_side)
    rowptr, colind = ora.fvcr_csr(E, es, n_side)
    p = ora.make_params(disc="fvcr", elem=elem, upwind="full", kin_visc=1e-3, density=1.2, source=src)
    what = capi.JAC_A | capi.DEF_A | capi.JAC_M | capi.DEF_M | capi.RHS
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, what, elem_sides=es, n_side=n_side, scale_a=0.01, scale_m=1.0)
    for mode in ("colored", "atomic"):
        gv, gd = disc.assemble(what, u, scale_a=0.01, scale_m=1.0, scatter_mode=MODES[mode])
        eg, ee = parity.entry_errors(gv, ov, rowptr)
        assert eg < TOL and ee < TOL
        eg, ee = parity.entry_errors(gd, od)
        assert eg < TOL and ee < TOL


@pytest.mark.parametrize("elem,n", [("tri", 24), ("tet", 8), ("quad", 16), ("hex", 6)])
def test_fvcr_gather_is_bitwise_the_coloured_result(elem, n):
    """NSB_SCATTER_GATHER serves FVCR (beta = 0) with ONE launch of fire-and-forget reductions in element order: a CR entry has at
    most two contributions (a side has two elements), 0 + a + b == 0 + b + a exactly, so the bits equal the coloured sweeps and
    do not change from run to run; beta != 0 takes the coloured sweeps"""
    coords, conn, es, n_side, u = _case(elem, n, seed=11)
    disc = pkg.NavierStokesFVCR(FCTS[coords.shape[1]], "Inner")
    disc.set_kinematic_viscosity(1e-3); disc.set_upwind("lps"); disc.set_exact_jacobian(1.0)
    disc.set_grid(elem, conn, coords, es, n_side)
    what = capi.JAC_A | capi.DEF_A | capi.JAC_M | capi.DEF_M
    cv, cd = disc.assemble(what, u, scale_a=0.7, scatter_mode=capi.SCATTER_COLORED)
    for _ in range(5):
        gv, gd = disc.assemble(what, u, scale_a=0.7, scatter_mode=capi.SCATTER_GATHER)
        assert np.array_equal(gv, cv) and np.array_equal(gd, cd)
    # accumulate (beta = 1): served by the coloured sweeps, same bits as an explicit coloured request
    gv2, gd2 = disc.assemble(what, u, scale_a=0.7, values=cv.copy(), defect=cd.copy(), beta=1.0, scatter_mode=capi.SCATTER_GATHER)
    cv2, cd2 = disc.assemble(what, u, scale_a=0.7, values=cv.copy(), defect=cd.copy(), beta=1.0, scatter_mode=capi.SCATTER_COLORED)
    assert np.array_equal(gv2, cv2) and np.array_equal(gd2, cd2)
    assert np.allclose(gv2, 2.0 * cv, rtol=1e-13, atol=1e-13 * np.abs(cv).max())
    disc.close()


def test_fvcr_errors():
    coords, conn = meshgen.make_mesh("tri", 3)
    d = pkg.NavierStokesFVCR("u,v,p", "Inner")
    d.set_grid("tri", conn, coords)
    d.set_kinematic_viscosity(0.01)
    with pytest.raises(pkg.UGError, match="Upwinding for convective Term"):
        d.prep_elem_loop()
    d.set_upwind("positive")
    with pytest.raises(pkg.UGError, match="No update function registered"):
        d.prep_elem_loop()
    q, cq = meshgen.make_mesh("prism", 2)
    with pytest.raises(pkg.UGError, match="prism and pyramid CR geometries are out of scope"):
        pkg.NavierStokesFVCR("u,v,w,p", "Inner").set_grid("prism", cq, q, np.zeros((cq.shape[0], 5), np.int32), 1)
    # the CR diagnostics and the constraint stay on simplices
    q, cq = meshgen.make_mesh("quad", 3)
    d4 = pkg.NavierStokesFVCR("u,v,p", "Inner")
    d4.set_grid("quad", cq, q)
    out = np.zeros(1)
    u4 = np.zeros(d4.num_dofs)
    assert capi.lib().nsb_diagnostic(d4._context(), 1, u4.ctypes.data, 0.1, out.ctypes.data, capi.HOST) == capi.ERR_UNSUPPORTED
