"""Boundary faces of the FV1 geometry (our spec of ugcore's FV1Geometry BF) and the boundary discs built on them in the oracle:
NavierStokesNoNormalStressOutflowFV1 (fv1/bnd/no_normal_stress_outflow_fv1.cpp:192-427) and the continuity term of
NavierStokesInflowFV1 (fv1/bnd/inflow_fv1_impl.h:42-82). CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp

from plugin_navierstokes_b200 import meshgen
from tests.conftest import jittered_ref_element

ELEMS = ["tri", "quad", "tet", "hex", "prism"]


@pytest.mark.parametrize("elem", ELEMS)
def test_bf_close_the_scv_surfaces(ora, elem):
    """with every side taken as boundary, the SCVFs and BFs of a corner form the closed surface of its SCV"""
    e = ora.ELEM[elem]
    x = jittered_ref_element(elem, seed=3, amp=0.12, scale=1.7, shift=0.3)
    g = ora.fv1_geometry(e, x)
    dim, nsh = ora.DIM[e], ora.NSH[e]
    tot = np.zeros((nsh, dim))
    for ip in range(ora.NIP[e]):
        tot[g["frm"][ip]] += g["normal"][ip]
        tot[g["to"][ip]] -= g["normal"][ip]
    area = 0.0
    for s in range(ora.NSIDE[e]):
        cs = ora.side_corners(e, s)
        ssum = np.zeros(dim)
        for j, co in enumerate(cs):
            nid, n, xip, N, G = ora.fv1_bf_geometry(e, x, s, j)
            assert nid == co
            tot[co] += n
            ssum += n
            assert abs(N.sum() - 1.0) < 1e-14 and np.abs(G.sum(axis=0)).max() < 1e-13
            assert np.allclose(xip, N @ x, atol=1e-14)
            area += np.linalg.norm(n)
        # outward: away from the barycentre
        assert ssum @ (x[cs].mean(axis=0) - x.mean(axis=0)) > 0
    assert np.abs(tot).max() < 1e-14 * max(1.0, area)


def test_bf_unit_cube_and_square(ora):
    cube = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    nid, n, xip, N, G = ora.fv1_bf_geometry(ora.HEX, cube, 5, 0)               # top side, corner 4
    assert nid == 4 and np.allclose(n, [0, 0, 0.25]) and np.allclose(xip, [0.25, 0.25, 1.0])
    sq = np.array([[0, 0], [2, 0], [2, 1], [0, 1]], float)
    nid, n, xip, N, G = ora.fv1_bf_geometry(ora.QUAD, sq, 1, 1)                # right edge, corner 2
    assert nid == 2 and np.allclose(n, [0.5, 0.0]) and np.allclose(xip, [2.0, 0.75])


def _problem(ora, elem, n, seed=4):
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=seed)
    dim = coords.shape[1]
    u = (meshgen.state_cavity2d(coords, seed=seed, noise=0.05) if dim == 2 else meshgen.state_vortex3d(coords, seed=seed, noise=0.05)).reshape(-1)
    rowptr, colind = ora.fv1_csr(ora.ELEM[elem], conn, coords.shape[0])
    return coords, conn, u, rowptr, colind


@pytest.mark.parametrize("elem,n", [("tri", 5), ("quad", 5), ("tet", 3), ("hex", 3), ("prism", 3)])
@pytest.mark.parametrize("laplace", [False, True])
def test_outflow_picard_property_and_pattern(ora, elem, n, laplace):
    """J(u) u = d(u) for the fixed-point Jacobian of the outflow disc (diffusive and continuity parts are linear, the convective
    part is flux(u) * StdVel), and the contributions stay inside the CSR pattern of the element coupling"""
    coords, conn, u, rowptr, colind = _problem(ora, elem, n)
    be, bs = meshgen.boundary_sides(elem, conn, coords, where=lambda c: np.isclose(c[:, 0], coords[:, 0].max()))
    assert len(be) > 0
    p = ora.make_params(elem=elem, upwind="full", stab="fields", laplace=laplace, kin_visc=0.05, density=1.3)
    v, d = ora.fv1_boundary(p, ora.BND_OUTFLOW, be, bs, conn, coords, u, rowptr, colind, ora.JAC_A | ora.DEF_A)
    A = sp.csr_matrix((v, colind, rowptr), shape=(u.size, u.size))
    assert np.abs(A @ u - d).max() < 1e-12 * np.abs(d).max()
    # rows of nodes off the outflow boundary are untouched
    nf = coords.shape[1] + 1
    on = np.zeros(coords.shape[0], bool)
    for q in range(len(be)):
        on[conn[be[q]][list(meshgen.SIDES[elem][bs[q]])]] = True
    rows = np.repeat(np.arange(u.size), np.diff(rowptr))
    assert np.all(v[~on[rows // nf]] == 0) and np.all(d.reshape(-1, nf)[~on] == 0)
    # scale_a scales, Stokes drops the convective part
    v2, d2 = ora.fv1_boundary(p, ora.BND_OUTFLOW, be, bs, conn, coords, u, rowptr, colind, ora.JAC_A | ora.DEF_A, scale_a=0.5)
    assert np.allclose(v2, 0.5 * v, rtol=1e-15, atol=0) and np.allclose(d2, 0.5 * d, rtol=1e-15, atol=0)
    ps = ora.make_params(elem=elem, upwind="full", stab="fields", laplace=laplace, kin_visc=0.05, density=1.3, stokes=True)
    vs, ds = ora.fv1_boundary(ps, ora.BND_OUTFLOW, be, bs, conn, coords, u, rowptr, colind, ora.JAC_A | ora.DEF_A)
    u0 = u.copy()
    As = sp.csr_matrix((vs, colind, rowptr), shape=(u.size, u.size))
    assert np.abs(As @ (2 * u0) - 2 * ds).max() < 1e-12 * np.abs(ds).max()       # linear in u


@pytest.mark.parametrize("elem,n", [("quad", 4), ("hex", 3), ("tri", 4), ("tet", 2), ("prism", 2)])
def test_mass_balance_of_a_uniform_flow(ora, elem, n):
    """uniform velocity, no stabilisation: the continuity defect of the element loop plus outflow faces on the WHOLE boundary is
    zero at every node (each control volume is closed by SCVFs and BFs); the inflow term with the same datum has the same
    boundary sum with the opposite role"""
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=2)
    dim = coords.shape[1]
    nf = dim + 1
    vel = np.array([0.7, -0.3, 0.45])[:dim]
    u = np.zeros((coords.shape[0], nf)); u[:, :dim] = vel
    u = u.reshape(-1)
    rowptr, colind = ora.fv1_csr(ora.ELEM[elem], conn, coords.shape[0])
    p = ora.make_params(elem=elem, upwind="full", stab="none", kin_visc=0.01)
    _, d = ora.assemble(p, conn, coords, u, rowptr, colind, ora.DEF_A)
    be, bs = meshgen.boundary_sides(elem, conn)
    _, d = ora.fv1_boundary(p, ora.BND_OUTFLOW, be, bs, conn, coords, u, rowptr, colind, ora.DEF_A, defect=d)
    assert np.abs(d.reshape(-1, nf)[:, dim]).max() < 1e-13
    data = np.broadcast_to(vel, (len(be), 4, dim)).copy()
    _, di = ora.fv1_boundary(p, ora.BND_INFLOW, be, bs, conn, coords, None, rowptr, colind, ora.DEF_A, data=data)
    assert abs(di.sum()) < 1e-13 and np.abs(di).max() > 1e-3                     # closed boundary: net flux of a constant field = 0
    assert np.all(di.reshape(-1, nf)[:, :dim] == 0)
    # the inflow term equals the outflow continuity term of the same velocity field (both are  u . n  per boundary face)
    _, dq = ora.fv1_boundary(p, ora.BND_OUTFLOW, be, bs, conn, coords, u, rowptr, colind, ora.DEF_A)
    assert np.allclose(di.reshape(-1, nf)[:, dim], dq.reshape(-1, nf)[:, dim], atol=1e-15)


def test_bf_ips_of_meshgen_match_the_oracle(ora):
    for elem, n in [("tri", 3), ("quad", 3), ("tet", 2), ("hex", 2), ("prism", 2)]:
        coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=7)
        be, bs = meshgen.boundary_sides(elem, conn)
        xip = meshgen.fv1_bf_ips(elem, conn, coords, be, bs)
        for q in range(0, len(be), 3):
            for j in range(len(meshgen.SIDES[elem][bs[q]])):
                _, _, x, _, _ = ora.fv1_bf_geometry(ora.ELEM[elem], coords[conn[be[q]]], int(bs[q]), j)
                assert np.allclose(x, xip[q, j], atol=1e-14)
