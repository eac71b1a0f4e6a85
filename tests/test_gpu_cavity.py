"""end-to-end use of the device path: lid-driven cavity solved on the GPU (examples/cavity.py) -- assembly with the resident
Jacobian, Dirichlet post-pass of walls / lid, dense LU + iterative refinement whose residual comes from nsb_apply_jacobian. The nonlinear defect drops by
eight orders of magnitude and the converged state is a root of the ORACLE's defect as well."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


@pytest.mark.parametrize("dim,cells", [(2, 12), (3, 5)])
def test_cavity_converges_and_satisfies_the_oracle_defect(ora, dim, cells):
    import cavity
    disc, coords, conn, u, hist = cavity.solve(dim, cells, re=50.0, verbose=False)
    assert hist[-1] < 1e-8 * hist[0] and len(hist) < 40
    uh = u.cpu().numpy()
    elem = "quad" if dim == 2 else "hex"
    rowptr, colind = ora.fv1_csr(ora.ELEM[elem], conn, coords.shape[0])
    p = ora.make_params(elem=elem, upwind=disc.upwind_name, stab="fields", kin_visc=1.0 / 50.0)
    _, od = ora.assemble(p, conn, coords, uh, rowptr, colind, ora.DEF_A)
    nf = dim + 1
    lo, hi = coords.min(axis=0), coords.max(axis=0)
    on_bnd = np.zeros(coords.shape[0], dtype=bool)
    for d in range(dim):
        on_bnd |= np.isclose(coords[:, d], lo[d]) | np.isclose(coords[:, d], hi[d])
    free = np.ones(uh.size, dtype=bool)
    for d in range(dim):
        free[np.nonzero(on_bnd)[0] * nf + d] = False
    free[nf - 1] = False
    assert np.abs(od[free]).max() < 1e-8 * hist[0]
    lid = np.isclose(coords[:, dim - 1], hi[dim - 1])
    assert np.allclose(uh.reshape(-1, nf)[lid, 0], 1.0) and np.allclose(uh.reshape(-1, nf)[on_bnd & ~lid, :dim], 0.0)
    disc.close()


def test_cavity_re100_reproduces_the_reference_ghia_tables():
    """Solution-level known-answer check against the only golden data the reference carries for this path: the Ghia et al. centre-line
    velocities embedded in DrivenCavityLinesEval (incompressible/navier_stokes_tools.h:578-598). The cavity at Re = 100 is solved on
    the device (FV1, LPS upwind + FIELDS, walls / lid through the Dirichlet post-pass, resident Jacobian) and evaluated with the
    mirror of that function: the differences shrink with the mesh width and reach plotting accuracy at 64^2."""
    import cavity
    from plugin_navierstokes_b200 import tools
    res = {}
    for cells in (32, 64):
        disc, coords, conn, u, hist = cavity.solve(2, cells, re=100.0, verbose=False, upwind="lps")
        assert hist[-1] < 1e-8 * hist[0]
        res[cells] = tools.DrivenCavityLinesEval(u.cpu().numpy(), coords, conn, 100)["Ghia"]
        disc.close()
    v32, v64 = res[32]["vertical"], res[64]["vertical"]
    h32, h64 = res[32]["horizontal"], res[64]["horizontal"]
    # measured on B200 (profiles/r2_cavity_ghia.txt): 32^2 0.0264 / 0.0231, 64^2 0.0119 / 0.0072, 96^2 0.0067 / 0.0020
    assert v64["max_diff"] < 0.015 and v64["average_diff"] < 0.006
    assert h64["max_diff"] < 0.010 and h64["average_diff"] < 0.003
    assert v64["max_diff"] < 0.6 * v32["max_diff"] and h64["max_diff"] < 0.6 * h32["max_diff"]


def test_fvcr_cavity_re100_converges_to_the_reference_ghia_tables():
    """the same known-answer check for NavierStokesFVCR (Crouzeix-Raviart velocities on triangle sides, FullUpwind, Dirichlet values on
    the boundary sides): first-order convergence to the Ghia table of the reference (profiles/r2_cavity_ghia.txt: u on x = 0.5
    max 0.072 / 0.046 / 0.026 at 16^2 / 32^2 / 64^2)"""
    import cavity
    from plugin_navierstokes_b200 import tools
    res = {}
    for cells in (16, 32):
        disc, coords, conn, es, u, hist = cavity.solve_fvcr(cells, re=100.0, verbose=False, upwind="full")
        assert hist[-1] < 1e-7 * hist[0]
        res[cells] = tools.DrivenCavityLinesEval(u.cpu().numpy(), coords, conn, 100, elem_sides=es)["Ghia"]
        disc.close()
    assert res[32]["vertical"]["max_diff"] < 0.055 and res[32]["horizontal"]["max_diff"] < 0.11
    assert res[32]["vertical"]["max_diff"] < 0.75 * res[16]["vertical"]["max_diff"]
    assert res[32]["horizontal"]["max_diff"] < 0.75 * res[16]["horizontal"]["max_diff"]


def test_fvcr_cavity_on_quadrilaterals_converges_to_the_reference_ghia_tables():
    """NavierStokesFVCR on quadrilaterals (rotated bilinear Crouzeix-Raviart velocities): the same known-answer check, pins the CR
    geometry / shapes of the non-affine element types at the level of the discretisation error (measured on B200,
    profiles/r2_cavity_fvcr_quads.txt: u on x = 0.5 max 0.115 / 0.079, v on y = 0.5 max 0.058 / 0.038 at 16^2 / 32^2, FullUpwind)"""
    import cavity
    from plugin_navierstokes_b200 import tools
    res = {}
    for cells in (16, 32):
        disc, coords, conn, es, u, hist = cavity.solve_fvcr(cells, re=100.0, verbose=False, upwind="full", elem="quad")
        assert hist[-1] < 1e-7 * hist[0]
        res[cells] = tools.DrivenCavityLinesEval(u.cpu().numpy(), coords, conn, 100, elem_sides=es)["Ghia"]
        disc.close()
    print("FVCR quads", {c: (res[c]["vertical"]["max_diff"], res[c]["horizontal"]["max_diff"]) for c in res})
    assert res[32]["vertical"]["max_diff"] < 0.09 and res[32]["horizontal"]["max_diff"] < 0.05
    assert res[32]["vertical"]["max_diff"] < 0.8 * res[16]["vertical"]["max_diff"]
    assert res[32]["horizontal"]["max_diff"] < 0.8 * res[16]["horizontal"]["max_diff"]


@pytest.mark.parametrize("elem", ["hex", "tet", "prism"])
def test_extruded_cavity_pins_the_3d_element_types_to_the_ghia_tables(elem):
    """FV1 on hexahedra / tetrahedra solving the 2-D problem (the square extruded by one cell in z, w = 0, zero flux through the
    z faces): converges to the reference's Ghia table like the quadrilateral run (profiles/r2_cavity_ghia.txt: hex 0.046 / 0.027,
    tet 0.027 / 0.015 at 16^2 / 32^2 for u on x = 0.5, LPS upwind)"""
    import cavity
    from plugin_navierstokes_b200 import tools
    res = {}
    for cells in (16, 32):
        disc, c2, q2, u2, hist = cavity.solve_extruded(elem, cells, re=100.0, verbose=False, upwind="lps")
        assert hist[-1] < 1e-7 * hist[0]
        res[cells] = tools.DrivenCavityLinesEval(u2, c2, q2, 100)["Ghia"]
        disc.close()
    print("extruded", elem, {c: (res[c]["vertical"]["max_diff"], res[c]["horizontal"]["max_diff"]) for c in res})
    assert res[32]["vertical"]["max_diff"] < 0.032 and res[32]["horizontal"]["max_diff"] < 0.030
    assert res[32]["vertical"]["max_diff"] < 0.7 * res[16]["vertical"]["max_diff"]
    assert res[32]["horizontal"]["max_diff"] < 0.7 * res[16]["horizontal"]["max_diff"]
