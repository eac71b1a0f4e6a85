"""Host mirror of the CRILUT preconditioner (fvcr/cr_ilut.h:87-497) on FVCR Jacobians assembled by the oracle. CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp

from plugin_navierstokes_b200 import cr_reorder, meshgen
from plugin_navierstokes_b200.cr_ilut import CRILUTPreconditioner


def _fvcr_system(ora, n=5, seed=2):
    coords, conn = meshgen.make_mesh("tri", n, jitter=0.15, seed=seed)
    es, n_side = meshgen.element_sides("tri", conn)
    rng = np.random.default_rng(seed)
    u = np.concatenate([0.3 * rng.uniform(-1, 1, n_side * 2) + np.tile([0.5, 0.1], n_side), rng.uniform(-1, 1, conn.shape[0])])
    p = ora.make_params(disc="fvcr", elem="tri", upwind="full", kin_visc=0.05)
    rowptr, colind = ora.fvcr_csr(ora.TRI, es, n_side)
    # instationary Jacobian: mass / dt + stiffness (velocity block with a dominant diagonal, zero pressure diagonal)
    vals, _ = ora.assemble(p, conn, coords, u, rowptr, colind, ora.JAC_A | ora.JAC_M | ora.DEF_A, elem_sides=es, n_side=n_side, scale_a=1.0, scale_m=20.0)
    return rowptr, colind, vals, es, n_side


def test_zero_threshold_is_the_exact_lu(ora):
    rowptr, colind, vals, es, n_side = _fvcr_system(ora)
    n = rowptr.size - 1
    A = sp.csr_matrix((vals, colind, rowptr), shape=(n, n))
    ilu = CRILUTPreconditioner(0.0)
    assert ilu.preprocess(rowptr, colind, vals)
    d = np.random.default_rng(0).uniform(-1, 1, n)
    c = ilu.step(d)
    assert np.abs(A @ c - d).max() < 1e-9 * np.abs(d).max()
    # L is strictly lower with unit diagonal implied, U starts with its diagonal
    assert all(all(k < i for k in ilu.L[i][0]) for i in range(n)) and all(ilu.U[i][0][0] == i for i in range(n))


def test_thresholds_drop_fill_by_block_type_and_still_precondition(ora):
    rowptr, colind, vals, es, n_side = _fvcr_system(ora)
    n = rowptr.size - 1
    A = sp.csr_matrix((vals, colind, rowptr), shape=(n, n))
    exact = CRILUTPreconditioner(0.0); exact.preprocess(rowptr, colind, vals)
    ilu = CRILUTPreconditioner(1e-3); ilu.preprocess(rowptr, colind, vals)
    assert ilu.nnz_LU < exact.nnz_LU
    # a harder velocity-velocity threshold only removes velocity-velocity fill
    vv = CRILUTPreconditioner(1e-1, 1e-3); vv.preprocess(rowptr, colind, vals)
    assert vv.nnz_LU < ilu.nnz_LU and (vv.eps_vv, vv.eps_vp, vv.eps_pv, vv.eps_pp, vv.eps) == (1e-1, 1e-3, 1e-3, 1e-3, 1e-3)
    # preconditioned Richardson iteration converges
    d = np.random.default_rng(1).uniform(-1, 1, n)
    x = np.zeros(n)
    r0 = np.linalg.norm(d)
    for _ in range(12):
        x = x + ilu.step(d - A @ x)
    assert np.linalg.norm(d - A @ x) < 1e-6 * r0


def test_cuthill_mckee_ordering_keeps_the_factorisation_alive_and_smaller(ora):
    """with OrderCRCuthillMcKee every pressure follows its velocities: the pressure pivots exist (Schur complement fill) and the
    exact factors need far less storage than in the layout 'all velocities, then all pressures'"""
    rowptr, colind, vals, es, n_side = _fvcr_system(ora, n=6)
    n = rowptr.size - 1
    new = cr_reorder.OrderCRCuthillMcKee(es, n_side)
    rp, ci, va = cr_reorder.permute_csr(rowptr, colind, vals, new)
    a = CRILUTPreconditioner(0.0); a.preprocess(rowptr, colind, vals)
    b = CRILUTPreconditioner(0.0); b.preprocess(rp, ci, va)
    assert b.nnz_LU < 0.6 * a.nnz_LU
    A = sp.csr_matrix((va, ci, rp), shape=(n, n))
    d = np.random.default_rng(2).uniform(-1, 1, n)
    assert np.abs(A @ b.step(d) - d).max() < 1e-9


def test_singular_last_row_is_skipped():
    """cr_ilut.h:431-448: a (near-)zero last pivot -- the constant-pressure mode -- gives a zero correction there"""
    rowptr, colind, vals = np.array([0, 2, 4]), np.array([0, 1, 0, 1]), np.array([1.0, 1.0, 1.0, 1.0])
    ilu = CRILUTPreconditioner(1e-6)
    ilu.preprocess(rowptr, colind, vals)
    c = ilu.step(np.array([1.0, 1.0]))
    assert c[1] == 0.0 and c[0] == 1.0 and not ilu.warnings
    ilu.step(np.array([1.0, 2.0]))
    assert ilu.warnings and "non-zero rhs" in ilu.warnings[0]
    with pytest.raises(ValueError, match="1, 2 or 4"):
        CRILUTPreconditioner(1.0, 2.0, 3.0)
