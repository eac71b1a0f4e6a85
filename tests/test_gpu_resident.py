"""GPU-resident Jacobian hand-off (nsb_assemble_resident / nsb_apply_jacobian), the theta-scheme combination and the
Dirichlet post-pass of the wall / inflow boundary conditions (nsb_set_dirichlet / nsb_adjust_*), all through the
C ABI, against the CPU oracle + scipy."""
import numpy as np
import pytest
import scipy.sparse as sp

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen
from tests import parity
from tests.parity import TOL

pytestmark = pytest.mark.gpu
JD = capi.JAC_A | capi.DEF_A


def _setup(ora, elem, n, upwind="lps", stab="fields", disc_kind="fv1"):
    coords, conn, u = parity.make_case(elem, n, seed=5)
    dim = coords.shape[1]
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    parity.configure(disc, upwind=upwind, stab=stab)
    disc.set_grid(elem, conn, coords)
    rowptr, colind = ora.fv1_csr(ora.ELEM[elem], conn, coords.shape[0])
    p = ora.make_params(elem=elem, upwind=upwind, stab=stab)
    return disc, coords, conn, u.reshape(-1), rowptr, colind, p


@pytest.mark.parametrize("elem,n", [("hex", 6), ("tet", 4), ("quad", 12), ("tri", 12)])
def test_resident_jacobian_matvec_host_and_device(ora, elem, n):
    import torch
    disc, coords, conn, u, rowptr, colind, p = _setup(ora, elem, n)
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, JD)
    A = sp.csr_matrix((ov, colind, rowptr), shape=(u.size, u.size))
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, u.size)
    y0 = rng.uniform(-1, 1, u.size)
    # host vectors, device-resident matrix
    d = disc.assemble_resident(JD, u)
    eg, ee = parity.entry_errors(d, od)
    assert eg < TOL and ee < TOL
    y = disc.apply_jacobian(x)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    y = disc.apply_jacobian(x, y=y0.copy(), alpha=-1.0, beta=1.0)             # residual form d - J x
    ref2 = y0 - A @ x
    assert np.abs(y - ref2).max() <= 1e-12 * max(np.abs(ref).max(), np.abs(y0).max())
    # device vectors; explicit values tensor = what nsb_assemble returns
    ud = torch.from_numpy(u).cuda()
    vals, dfc = disc.assemble(JD, ud)
    xd = torch.from_numpy(x).cuda()
    yd = disc.apply_jacobian(xd, values=vals)
    yr = disc.apply_jacobian(xd)                                               # resident copy
    torch.cuda.synchronize()
    assert np.abs(yd.cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()
    assert torch.equal(yd, yr)                                                 # same values, deterministic reduction
    assert disc.resident_jacobian_ptr() != 0
    disc.close()


def test_resident_fvcr_matvec(ora):
    coords, conn = meshgen.tet_grid(4, 3, 3, jitter=0.2, seed=4)
    es, n_side = meshgen.element_sides("tet", conn)
    rng = np.random.default_rng(4)
    u = np.concatenate([rng.uniform(-1, 1, n_side * 3), rng.uniform(-1, 1, conn.shape[0])])
    disc = pkg.NavierStokesFVCR("u,v,w,p", "Inner")
    disc.set_kinematic_viscosity(1e-2)
    disc.set_upwind("full")
    disc.set_grid("tet", conn, coords, es, n_side)
    rowptr, colind = ora.fvcr_csr(ora.TET, es, n_side)
    p = ora.make_params(disc="fvcr", elem="tet", upwind="full", kin_visc=1e-2)
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, JD, elem_sides=es, n_side=n_side)
    A = sp.csr_matrix((ov, colind, rowptr), shape=(u.size, u.size))
    disc.assemble_resident(JD, u)
    x = rng.uniform(-1, 1, u.size)
    y = disc.apply_jacobian(x)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    disc.close()


@pytest.mark.parametrize("theta", [1.0, 0.5])
def test_theta_time_step_combination(ora, theta):
    disc, coords, conn, u, rowptr, colind, p = _setup(ora, "hex", 5, upwind="full")
    rng = np.random.default_rng(2)
    u_old = u + 0.05 * rng.uniform(-1, 1, u.size)
    dt = 1e-2
    pt = ora.make_params(elem="hex", upwind="full", stab="fields", dt=dt, time_dependent=True)
    full = capi.JAC_A | capi.JAC_M | capi.DEF_A | capi.DEF_M | capi.RHS
    ov, od = ora.assemble(pt, conn, coords, u, rowptr, colind, full, sol0=u, sol1=u_old, scale_a=theta * dt, scale_m=1.0)
    what_old = capi.DEF_M | ((capi.DEF_A | capi.RHS) if theta < 1.0 else 0)
    ov2, od = ora.assemble(pt, conn, coords, u_old, rowptr, colind, what_old, sol0=u, sol1=u_old, scale_a=(1 - theta) * dt, scale_m=-1.0,
                           values=np.zeros_like(ov), defect=od)
    step = pkg.ThetaTimeStep(disc, theta)
    d = step.assemble(u, u_old, dt)
    eg, ee = parity.entry_errors(d, od)
    assert eg < TOL and ee < 1e-10, (eg, ee)       # (the M parts cancel to round-off of the state: floored statistic only)
    A = sp.csr_matrix((ov, colind, rowptr), shape=(u.size, u.size))
    x = rng.uniform(-1, 1, u.size)
    y = disc.apply_jacobian(x)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    disc.close()


@pytest.mark.parametrize("elem,n", [("hex", 5), ("quad", 10)])
def test_wall_and_inflow_dirichlet_post_pass(ora, elem, n):
    import torch
    disc, coords, conn, u, rowptr, colind, p = _setup(ora, elem, n, upwind="full")
    dim = coords.shape[1]
    nf = dim + 1
    lo, hi = coords.min(axis=0), coords.max(axis=0)
    wall_nodes = np.nonzero(np.isclose(coords[:, 1], lo[1]) | np.isclose(coords[:, 1], hi[1]))[0]
    inflow_nodes = np.nonzero(np.isclose(coords[:, 0], lo[0]))[0]
    wall = pkg.NavierStokesWall(disc)
    wall.add(wall_nodes)
    inflow = pkg.NavierStokesInflowFV1(disc)
    prof = (lambda x, y: (4.0 * y * (1 - y), 0.0)) if dim == 2 else (lambda x, y, z: (4.0 * y * (1 - y), 0.0, 0.0))
    inflow.add(prof, inflow_nodes, coords)
    dw, vw = wall.dirichlet()
    di, vi = inflow.dirichlet()
    # both constraints act on one context: the wall is registered first and wins on shared nodes (ugcore applies them in order)
    dofs = np.concatenate([dw, di])
    vals = np.concatenate([vw, vi])
    dofs, first = np.unique(dofs, return_index=True)
    vals = vals[first]
    assert np.all(dofs % nf < dim)                                             # velocity components only
    disc.set_dirichlet(dofs)
    ud = torch.from_numpy(u).cuda()
    jv, dv = disc.assemble(JD, ud)
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, JD)
    disc.adjust_jacobian(jv)
    disc.adjust_vector(dv)
    disc.adjust_vector(ud, vals)
    torch.cuda.synchronize()
    # oracle-side post-pass (SetDirichletRow / zero defect / set solution)
    for r in dofs:
        seg = slice(rowptr[r], rowptr[r + 1])
        ov[seg] = (colind[seg] == r).astype(float)
    od[dofs] = 0.0
    uo = u.copy()
    uo[dofs] = vals
    eg, ee = parity.entry_errors(jv.cpu().numpy(), ov, rowptr)
    assert eg < TOL and ee < TOL
    assert np.array_equal(jv.cpu().numpy()[rowptr[dofs[0]]:rowptr[dofs[0] + 1]], ov[rowptr[dofs[0]]:rowptr[dofs[0] + 1]])
    eg, ee = parity.entry_errors(dv.cpu().numpy(), od)
    assert eg < TOL and ee < TOL
    assert np.array_equal(ud.cpu().numpy(), uo)
    # resident variant + host vectors
    disc.assemble_resident(JD, u)
    disc.adjust_jacobian()
    x = np.random.default_rng(3).uniform(-1, 1, u.size)
    y = disc.apply_jacobian(x)
    A = sp.csr_matrix((ov, colind, rowptr), shape=(u.size, u.size))
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.allclose(y[dofs], x[dofs], rtol=0, atol=0)
    h = disc.adjust_vector(np.ones(u.size))
    assert np.all(h[dofs] == 0.0) and h.sum() == u.size - dofs.size
    disc.close()


@pytest.mark.parametrize("elem,n", [("hex", 8), ("tri", 16)])
def test_resident_async_host_handoff_equals_synchronous(elem, n):
    """NSB_HOST_ASYNC (copies on the context's copy streams, completed by nsb_synchronize) returns the bits of the NSB_HOST calls,
    also when the staging buffers are reused by back-to-back steps with changing inputs"""
    import torch
    coords, conn, u = parity.make_case(elem, n, seed=7)
    dim = coords.shape[1]
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    parity.configure(disc, upwind="lps", stab="fields")
    disc.set_grid(elem, conn, coords)
    nd = u.size
    rng = np.random.default_rng(3)
    hu = torch.empty(nd, dtype=torch.float64).pin_memory(); hd = torch.empty(nd, dtype=torch.float64).pin_memory()
    hx = torch.empty(nd, dtype=torch.float64).pin_memory(); hy = torch.empty(nd, dtype=torch.float64).pin_memory()
    for step in range(3):
        us = u.reshape(-1) * (1.0 + 0.1 * step)
        xs = rng.uniform(-1, 1, nd)
        d_ref = disc.assemble_resident(JD, us).copy()
        y_ref = disc.apply_jacobian(xs).copy()
        hu.numpy()[:] = us; hx.numpy()[:] = xs; hd.numpy()[:] = np.nan; hy.numpy()[:] = np.nan
        disc.assemble_resident(JD, hu.numpy(), defect=hd.numpy(), asynchronous=True)
        disc.apply_jacobian(hx.numpy(), y=hy.numpy(), asynchronous=True)
        disc.synchronize(); disc.check_errors()
        assert np.array_equal(hd.numpy(), d_ref) and np.array_equal(hy.numpy(), y_ref)
    # a synchronous host-pointer call right behind asynchronous ones (no nsb_synchronize in between) reuses the staging buffers:
    # the context stream is ordered behind the pending copies, both results are right
    ua, ub = u.reshape(-1) * 0.9, u.reshape(-1) * 1.2
    da_ref = disc.assemble_resident(JD, ua).copy(); ya_ref = disc.apply_jacobian(xs).copy()
    db_ref = disc.assemble_resident(JD, ub).copy(); yb_ref = disc.apply_jacobian(xs).copy()
    hu.numpy()[:] = ua; hd.numpy()[:] = np.nan; hy.numpy()[:] = np.nan
    disc.assemble_resident(JD, hu.numpy(), defect=hd.numpy(), asynchronous=True)
    disc.apply_jacobian(hx.numpy(), y=hy.numpy(), asynchronous=True)
    db = disc.assemble_resident(JD, ub)
    yb = disc.apply_jacobian(xs)
    disc.synchronize()
    assert np.array_equal(db, db_ref) and np.array_equal(yb, yb_ref)
    assert np.array_equal(hd.numpy(), da_ref) and np.array_equal(hy.numpy(), ya_ref)
    with pytest.raises(pkg.UGError):
        disc.apply_jacobian(hx.numpy()[::2], y=hy.numpy(), asynchronous=True)
    disc.close()
