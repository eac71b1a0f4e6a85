"""Self-consistency invariants of the oracle's element routines and global loop.

The reference ships no tests; these invariants are what pins the restatement (parity unpinned).
"""
import itertools

import numpy as np
import pytest
import scipy.sparse as sp

from plugin_navierstokes_b200 import meshgen
from tests.conftest import jittered_ref_element

ELEMS = ["tri", "quad", "tet", "hex", "prism"]
UPWINDS = ["no", "full", "skewed", "lps", "positive"]
STABS = ["fields", "flow", "none"]


def _rand_u(ora, E, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return scale * rng.uniform(-1, 1, (ora.DIM[E] + 1, ora.NSH[E]))


@pytest.mark.parametrize("elem", ELEMS)
@pytest.mark.parametrize("upwind", UPWINDS)
@pytest.mark.parametrize("stab", STABS)
def test_flux_conservation(ora, elem, upwind, stab):
    """every flux is added to `from` and subtracted from `to`: rows of one function sum to zero"""
    E = ora.ELEM[elem]
    nsh, dim = ora.NSH[E], ora.DIM[E]
    x = jittered_ref_element(elem, seed=1, amp=0.1)
    u = _rand_u(ora, E, 2)
    for flags in (dict(), dict(peclet_blend=True, exact_jac=1.0), dict(laplace=True, exact_jac=0.5),
                  dict(pac=True, exact_jac=1.0)):
        p = ora.make_params(elem=elem, upwind=upwind, stab=stab, kin_visc=0.05, density=1.3, **flags)
        J, d = ora.fv1_elem(p, x, u, ora.JAC_A | ora.DEF_A)
        assert np.isfinite(J).all() and np.isfinite(d).all()
        Jr = J.reshape(dim + 1, nsh, dim + 1, nsh)
        scale = np.abs(J).max()
        assert np.abs(Jr.sum(axis=1)).max() < 1e-12 * scale
        assert np.abs(d.reshape(dim + 1, nsh).sum(axis=1)).max() < 1e-12 * np.abs(d).max()


@pytest.mark.parametrize("elem", ELEMS)
@pytest.mark.parametrize("stab", STABS)
@pytest.mark.parametrize("laplace", [False, True])
def test_stokes_is_linear(ora, elem, stab, laplace):
    """Stokes: J*u == d exactly (no convection; FIELDS/FLOW reduce to a linear closure)"""
    E = ora.ELEM[elem]
    x = jittered_ref_element(elem, seed=3, amp=0.1)
    u = _rand_u(ora, E, 4)
    p = ora.make_params(elem=elem, upwind=None, stab=stab, stokes=True, laplace=laplace, kin_visc=0.3, density=0.9)
    J, d = ora.fv1_elem(p, x, u, ora.JAC_A | ora.DEF_A)
    assert np.allclose(J @ u.ravel(), d, rtol=1e-11, atol=1e-12 * np.abs(d).max())


@pytest.mark.parametrize("elem", ELEMS)
def test_exact_jacobian_matches_finite_differences(ora, elem):
    """No upwind + no stabilisation + exact_jacobian=1 gives the true derivative of the defect"""
    E = ora.ELEM[elem]
    x = jittered_ref_element(elem, seed=5, amp=0.1)
    u = _rand_u(ora, E, 6)
    p = ora.make_params(elem=elem, upwind="no", stab="none", exact_jac=1.0, kin_visc=0.1)
    J, d0 = ora.fv1_elem(p, x, u, ora.JAC_A | ora.DEF_A)
    L = u.size
    Jfd = np.zeros((L, L))
    h = 1e-6
    for j in range(L):
        up, um = u.copy().ravel(), u.copy().ravel()
        up[j] += h
        um[j] -= h
        _, dp = ora.fv1_elem(p, x, up.reshape(u.shape), ora.DEF_A)
        _, dm = ora.fv1_elem(p, x, um.reshape(u.shape), ora.DEF_A)
        Jfd[:, j] = (dp - dm) / (2 * h)
    assert np.abs(J - Jfd).max() < 1e-7 * max(1.0, np.abs(J).max())


@pytest.mark.parametrize("elem", ELEMS)
def test_picard_jacobian_reproduces_defect(ora, elem):
    """fix-point linearisation: J(u)*u == d_A(u) whenever the stabilisation has no source/time terms.
    Holds for every upwind (transported velocity is linear in u once the shapes are frozen)."""
    E = ora.ELEM[elem]
    x = jittered_ref_element(elem, seed=7, amp=0.1)
    u = _rand_u(ora, E, 8)
    for upwind, stab, pac in itertools.product(UPWINDS, STABS, [False, True]):
        for peclet in (False, True):
            p = ora.make_params(elem=elem, upwind=upwind, stab=stab, pac=pac, peclet_blend=peclet, kin_visc=0.07)
            J, d = ora.fv1_elem(p, x, u, ora.JAC_A | ora.DEF_A)
            assert np.allclose(J @ u.ravel(), d, rtol=1e-10, atol=1e-11 * np.abs(d).max()), (upwind, stab, pac, peclet)


@pytest.mark.parametrize("elem", ELEMS)
@pytest.mark.parametrize("stab", ["fields", "flow"])
def test_dense_branch_solves_the_ip_system(ora, elem, stab):
    """Positive upwind -> dense branch (stabilization.cpp:244-403 / :590-771). Rebuild the ip system in
    numpy from the oracle's geometry + upwind output and check stab_vel solves it."""
    E = ora.ELEM[elem]
    dim, nsh, nip = ora.DIM[E], ora.NSH[E], ora.NIP[E]
    x = jittered_ref_element(elem, seed=9, amp=0.1)
    u = _rand_u(ora, E, 10)
    uold = _rand_u(ora, E, 11)
    visc, rho, dt = 0.05, 1.2, 0.1
    src = [0.3, -0.2, 0.1][:dim]
    p = ora.make_params(elem=elem, upwind="positive", stab=stab, kin_visc=visc, density=rho, source=src,
                        dt=dt, time_dependent=True)
    sv, shv, shp = ora.fv1_stab(p, x, u, sol0=u, sol1=uold)
    g = ora.fv1_geometry(E, x)
    N, G = g["shape"], g["ggrad"]
    std = N @ u[:dim].T
    up_sh, up_ip, up_len = ora.fv1_upwind(E, "positive", x, std)
    dn_sh, dn_ip, dn_len = ora.fv1_upwind(E, "positive", x, -std)
    nn = (g["normal"] ** 2).sum(axis=1)
    A = (0.5 * (g["vol"][g["frm"]] + g["vol"][g["to"]])) ** 2
    dl = 1 / (0.5 * A / nn + 3 * (nn if dim == 2 else g["c0c2sq"]) / 8)
    a = visc * dl
    nrm = np.linalg.norm(std, axis=1)
    b = nrm / up_len
    c = nrm / (up_len + dn_len)
    M = np.diag(a + b + 1 / dt) - up_ip * b[:, None]
    if stab == "flow":
        M += c[:, None] * (up_ip - dn_ip)
    for d in range(dim):
        rhs = np.full(nip, src[d]) + (N @ uold[d]) / dt
        cv = a[:, None] * N + b[:, None] * up_sh
        if stab == "flow":
            cv = cv + c[:, None] * (dn_sh - up_sh)
            for d2 in range(dim):
                if d2 != d:
                    cv = cv - std[:, d2:d2 + 1] * G[:, :, d2]
                    rhs = rhs + (std[:, d:d + 1] * G[:, :, d2]) @ u[d2]
        rhs = rhs + cv @ u[d] + (-G[:, :, d] / rho) @ u[dim]
        assert np.allclose(M @ sv[:, d], rhs, rtol=1e-10, atol=1e-12)
        # shapes: M * shape_vel(:,d,d,k) = cv(:,k)
        assert np.allclose(M @ shv[:, d, d, :], cv, rtol=1e-10, atol=1e-12)
        assert np.allclose(M @ shp[:, d, :], -G[:, :, d] / rho, rtol=1e-10, atol=1e-12)


def test_fields_diagonal_formula(ora):
    """hand formula of the FIELDS diagonal branch on the unit square, full upwind, stationary"""
    x = jittered_ref_element("quad", amp=0.0)
    u = np.array([[1.0, 1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 0.0], [0.0, 1.0, 1.0, 0.0]])   # u=(1,0), p = x
    visc = 0.5
    p = ora.make_params(elem="quad", upwind="full", stab="fields", kin_visc=visc)
    sv, shv, shp = ora.fv1_stab(p, x, u)
    # ip 0: n=(0.5,0) nn=0.25, A=(0.25)^2, RAW: 1/(0.5*A/nn + 3*nn/8) ; conv length to corner 0
    dl = 1 / (0.5 * 0.0625 / 0.25 + 3 * 0.25 / 8)
    a = visc * dl
    b = 1.0 / np.hypot(0.5, 0.25)
    diag = a + b
    # rhs_x = sum_k (a N_k + b up_k) u_k - dp/dx = a + b - 1
    assert np.isclose(sv[0, 0], (a + b - 1.0) / diag)
    assert np.isclose(sv[0, 1], 0.0)
    assert np.isclose(shv[0, 0, 0, 0], (a * 0.375 + b) / diag)
    assert np.isclose(shp[0, 0, 1], -0.75 / diag)     # dN_1/dx at ip (0.5,0.25) = (1-y) = 0.75


@pytest.mark.parametrize("elem", ELEMS)
def test_translation_and_scaling_invariance(ora, elem):
    E = ora.ELEM[elem]
    x = jittered_ref_element(elem, seed=12, amp=0.1)
    u = _rand_u(ora, E, 13)
    for upwind, stab in (("full", "fields"), ("lps", "flow"), ("positive", "flow"), ("skewed", "fields")):
        p = ora.make_params(elem=elem, upwind=upwind, stab=stab, kin_visc=0.05, exact_jac=1.0)
        J0, d0 = ora.fv1_elem(p, x, u, ora.JAC_A | ora.DEF_A)
        J1, d1 = ora.fv1_elem(p, x + 3.7, u, ora.JAC_A | ora.DEF_A)
        assert np.allclose(J0, J1, rtol=1e-9, atol=1e-11) and np.allclose(d0, d1, rtol=1e-9, atol=1e-11)


def test_mirror_symmetry_quad(ora):
    """mirroring the element and the state in x maps the defect onto the mirrored defect"""
    x = jittered_ref_element("quad", seed=14, amp=0.1)
    u = _rand_u(ora, ora.QUAD, 15)
    perm = [1, 0, 3, 2]                       # mirrored corner order stays counter-clockwise
    xm = x[perm] * np.array([-1.0, 1.0])
    um = u[:, perm] * np.array([[-1.0], [1.0], [1.0]])
    for upwind, stab in (("full", "fields"), ("lps", "flow"), ("positive", "fields")):
        p = ora.make_params(elem="quad", upwind=upwind, stab=stab, kin_visc=0.05)
        _, d = ora.fv1_elem(p, x, u, ora.DEF_A)
        _, dm = ora.fv1_elem(p, xm, um, ora.DEF_A)
        d, dm = d.reshape(3, 4), dm.reshape(3, 4)
        assert np.allclose(dm[0], -d[0][perm], atol=1e-12)
        assert np.allclose(dm[1], d[1][perm], atol=1e-12)
        assert np.allclose(dm[2], d[2][perm], atol=1e-12)


@pytest.mark.parametrize("elem", ELEMS)
def test_mass_and_rhs(ora, elem):
    E = ora.ELEM[elem]
    dim, nsh = ora.DIM[E], ora.NSH[E]
    x = jittered_ref_element(elem, seed=16, amp=0.1)
    u = _rand_u(ora, E, 17)
    src = [0.5, -1.0, 2.0][:dim]
    p = ora.make_params(elem=elem, density=1.7, source=src)
    g = ora.fv1_geometry(E, x)
    J, d = ora.fv1_elem(p, x, u, ora.JAC_M | ora.DEF_M)
    diag = np.concatenate([np.tile(g["vol"] * 1.7, dim), np.zeros(nsh)])
    assert np.allclose(J, np.diag(diag))
    assert np.allclose(d, diag * u.ravel())
    _, r = ora.fv1_elem(p, x, u, ora.RHS)
    assert np.allclose(r.reshape(dim + 1, nsh)[:dim], np.outer(src, g["vol"] * 1.7))
    assert np.allclose(r.reshape(dim + 1, nsh)[dim], 0)


def test_prep_elem_loop_errors(ora):
    """the throws of prep_elem_loop (fv1/navier_stokes_fv1.cpp:147,156) surface as errors"""
    x = jittered_ref_element("quad", amp=0.0)
    u = _rand_u(ora, ora.QUAD, 1)
    with pytest.raises(ora.OracleError, match="Stabilization has not been set"):
        ora.fv1_elem(ora.make_params(elem="quad", stab=None), x, u, ora.JAC_A)
    with pytest.raises(ora.OracleError, match="Upwinding for convective Term"):
        ora.fv1_elem(ora.make_params(elem="quad", upwind=None), x, u, ora.JAC_A)
    # Stokes needs no upwind
    ora.fv1_elem(ora.make_params(elem="quad", upwind=None, stokes=True), x, u, ora.JAC_A)


# ----------------------------------------------------------------------------------------------
# FVCR
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("elem", ["tri", "tet", "quad", "hex"])
@pytest.mark.parametrize("upwind", ["no", "full", "skewed", "lps"])
def test_fvcr_invariants(ora, elem, upwind):
    E = ora.ELEM[elem]
    dim, ns = ora.DIM[E], ora.NSIDE[E]
    x = jittered_ref_element(elem, seed=18, amp=0.1)
    rng = np.random.default_rng(19)
    u = rng.uniform(-1, 1, dim * ns + 1)
    for flags in (dict(), dict(peclet_blend=True), dict(laplace=True, grad_div=0.3)):
        p = ora.make_params(disc="fvcr", elem=elem, upwind=upwind, kin_visc=0.05, **flags)
        J, d = ora.fvcr_elem(p, x, u, ora.JAC_A | ora.DEF_A)
        assert np.isfinite(J).all() and np.isfinite(d).all()
        # momentum rows: flux conservation over sides
        Jm = J[:dim * ns].reshape(dim, ns, -1)
        assert np.abs(Jm.sum(axis=1)).max() < 1e-12 * np.abs(J).max()
        assert np.abs(d[:dim * ns].reshape(dim, ns).sum(axis=1)).max() < 1e-12 * np.abs(d).max()
        # Picard linearisation reproduces the defect
        assert np.allclose(J @ u, d, rtol=1e-10, atol=1e-12)
    # continuity row = divergence: constant velocity field has zero divergence defect
    uc = np.concatenate([np.repeat([0.3, -0.7, 0.2][:dim], ns), [0.0]])
    p = ora.make_params(disc="fvcr", elem=elem, upwind=upwind, kin_visc=0.05)
    _, d = ora.fvcr_elem(p, x, uc, ora.DEF_A)
    assert abs(d[-1]) < 1e-13


def test_fvcr_positive_upwind_is_rejected(ora):
    x = jittered_ref_element("tri", amp=0.0)
    with pytest.raises(ora.OracleError, match="No update function registered"):
        ora.fvcr_elem(ora.make_params(disc="fvcr", elem="tri", upwind="positive"), x, np.ones(7), ora.JAC_A)


def test_fvcr_rhs_has_no_density(ora):
    """fvcr/navier_stokes_fvcr.cpp:757 vs fv1/navier_stokes_fv1.cpp:866"""
    x = jittered_ref_element("tri", seed=1, amp=0.1)
    g = ora.cr_geometry(ora.TRI, x)
    p = ora.make_params(disc="fvcr", elem="tri", density=3.0, source=[1.0, 2.0])
    _, r = ora.fvcr_elem(p, x, np.zeros(7), ora.RHS)
    assert np.allclose(r[:6].reshape(2, 3), np.outer([1.0, 2.0], g["vol"]))


# ----------------------------------------------------------------------------------------------
# global loop
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("elem,n", [("tri", 5), ("quad", 5), ("tet", 3), ("hex", 3), ("prism", 3)])
def test_csr_pattern_is_full_element_coupling(ora, elem, n):
    E = ora.ELEM[elem]
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=1)
    nf = ora.DIM[E] + 1
    rowptr, colind = ora.fv1_csr(E, conn, coords.shape[0])
    # independent construction with scipy
    dofs = (conn[:, :, None] * nf + np.arange(nf)[None, None, :]).reshape(conn.shape[0], -1)
    L = dofs.shape[1]
    rows = np.repeat(dofs, L, axis=1).ravel()
    cols = np.tile(dofs, (1, L)).ravel()
    A = sp.csr_matrix((np.ones(rows.size), (rows, cols)), shape=(coords.shape[0] * nf,) * 2)
    A.sum_duplicates()
    A.sort_indices()
    assert np.array_equal(A.indptr, rowptr) and np.array_equal(A.indices, colind)


@pytest.mark.parametrize("elem,n", [("quad", 4), ("hex", 2), ("tri", 4), ("tet", 2), ("prism", 2)])
def test_global_assembly_is_sum_of_local(ora, elem, n):
    E = ora.ELEM[elem]
    dim, nsh = ora.DIM[E], ora.NSH[E]
    nf = dim + 1
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=2)
    u = meshgen.random_state(coords.shape[0], nf, seed=3)
    p = ora.make_params(elem=elem, upwind="lps", stab="flow", kin_visc=0.05, exact_jac=1.0)
    rowptr, colind = ora.fv1_csr(E, conn, coords.shape[0])
    vals, dfc = ora.assemble(p, conn, coords, u, rowptr, colind, ora.JAC_A | ora.DEF_A)
    A = sp.csr_matrix((vals, colind, rowptr))
    Aref = sp.lil_matrix(A.shape)
    dref = np.zeros(A.shape[0])
    for e in range(conn.shape[0]):
        J, d = ora.fv1_elem(p, coords[conn[e]], u[conn[e]].T, ora.JAC_A | ora.DEF_A)
        gi = (conn[e][None, :] * nf + np.arange(nf)[:, None]).ravel()      # [fct][sh]
        for i in range(J.shape[0]):
            dref[gi[i]] += d[i]
            for j in range(J.shape[1]):
                Aref[gi[i], gi[j]] += J[i, j]
    assert np.allclose(A.toarray(), Aref.toarray(), rtol=1e-13, atol=1e-14)
    assert np.allclose(dfc, dref, rtol=1e-13, atol=1e-14)
    # threaded (coloured) sweep equals the serial one
    v2, d2 = ora.assemble(p, conn, coords, u, rowptr, colind, ora.JAC_A | ora.DEF_A, nthreads=4)
    assert np.allclose(v2, vals, rtol=1e-13, atol=1e-15) and np.allclose(d2, dfc, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("elem,n", [("tri", 4), ("tet", 2), ("quad", 4), ("hex", 2)])
def test_fvcr_global(ora, elem, n):
    E = ora.ELEM[elem]
    dim = ora.DIM[E]
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=4)
    es, n_side = meshgen.element_sides(elem, conn)
    ndof = n_side * dim + conn.shape[0]
    rowptr, colind = ora.fvcr_csr(E, es, n_side)
    assert rowptr[-1] == colind.size and rowptr.size == ndof + 1
    for r in range(ndof):
        row = colind[rowptr[r]:rowptr[r + 1]]
        assert (np.diff(row) > 0).all()
    u = np.random.default_rng(5).uniform(-1, 1, ndof)
    p = ora.make_params(disc="fvcr", elem=elem, upwind="full", kin_visc=0.05)
    vals, dfc = ora.assemble(p, conn, coords, u, rowptr, colind, ora.JAC_A | ora.DEF_A, elem_sides=es, n_side=n_side)
    A = sp.csr_matrix((vals, colind, rowptr))
    assert np.allclose(A @ u, dfc, rtol=1e-10, atol=1e-12)        # Picard property survives the scatter
    # interior sides: momentum defect of a constant state vanishes (fluxes telescope), pressure 0
    uc = np.concatenate([np.tile([0.3, -0.2, 0.5][:dim], n_side), np.zeros(conn.shape[0])])
    _, dc = ora.assemble(p, conn, coords, uc, rowptr, colind, ora.DEF_A, elem_sides=es, n_side=n_side)
    cnt = np.bincount(es.ravel(), minlength=n_side)
    interior = np.repeat(cnt == 2, dim)
    assert np.abs(dc[:n_side * dim][interior]).max() < 1e-13
