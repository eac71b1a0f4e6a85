"""Boundary element discs on the device (nsb_set_boundary_faces / nsb_assemble_boundary) against the oracle:
NavierStokesNoNormalStressOutflowFV1 (fv1/bnd/no_normal_stress_outflow_fv1.cpp:192-427) and the continuity term of
NavierStokesInflowFV1 (fv1/bnd/inflow_fv1_impl.h:42-82), through the C ABI; and a channel solved with inflow / wall / outflow."""
import numpy as np
import pytest

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen
from tests import parity
from tests.parity import TOL

pytestmark = pytest.mark.gpu
JD = capi.JAC_A | capi.DEF_A


def _inflow(*x):
    return (1.0 + 0.3 * x[1], 0.2 * x[0]) if len(x) == 2 else (1.0 + 0.3 * x[1], 0.2 * x[0], -0.1 * x[2])


@pytest.mark.parametrize("elem,n,axis", [("tri", 7, 0), ("quad", 7, 0), ("tet", 3, 0), ("hex", 4, 0),
                                         ("prism", 3, 0), ("prism", 3, 2)])      # prisms: quadrilateral (x) and triangular (z) boundary sides
@pytest.mark.parametrize("flags", [dict(), dict(laplace=True), dict(stokes=True)])
def test_boundary_discs_match_the_oracle(ora, elem, n, axis, flags):
    import torch
    coords, conn, u = parity.make_case(elem, n, seed=9)
    u = u.reshape(-1)
    dim = coords.shape[1]
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    parity.configure(disc, upwind="full", stab="fields", visc=0.05, density=1.3, **flags)
    disc.set_grid(elem, conn, coords)
    xmax, xmin = coords[:, axis].max(), coords[:, axis].min()
    out_e, out_s = meshgen.boundary_sides(elem, conn, coords, where=lambda c: np.isclose(c[:, axis], xmax))
    in_e, in_s = meshgen.boundary_sides(elem, conn, coords, where=lambda c: np.isclose(c[:, axis], xmin))
    outflow = pkg.NavierStokesNoNormalStressOutflow(disc)
    outflow.add(out_e, out_s)
    outflow.apply()
    inflow = pkg.NavierStokesInflowFV1(disc)
    in_nodes = np.nonzero(np.isclose(coords[:, axis], xmin))[0]
    inflow.add(_inflow, in_nodes, coords, sides=(in_e, in_s), conn=conn, elem=elem)
    inflow.apply()
    rowptr, colind = ora.fv1_csr(ora.ELEM[elem], conn, coords.shape[0])
    p = ora.make_params(elem=elem, upwind="full", stab="fields", kin_visc=0.05, density=1.3, **flags)
    # oracle: element loop + both boundary discs, scale_a = 0.7 on the boundary part
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, JD)
    ov0, od0 = ov.copy(), od.copy()
    ora.fv1_boundary(p, ora.BND_OUTFLOW, out_e, out_s, conn, coords, u, rowptr, colind, JD, scale_a=0.7, values=ov, defect=od)
    data = np.array([[_inflow(*x) for x in row] for row in meshgen.fv1_bf_ips(elem, conn, coords, in_e, in_s)])
    ora.fv1_boundary(p, ora.BND_INFLOW, in_e, in_s, conn, coords, u, rowptr, colind, JD, data=data, scale_a=0.7, values=ov, defect=od)
    assert np.abs(ov - ov0).max() > 1e-3 and np.abs(od - od0).max() > 1e-3
    # device tensors, explicit values
    ud = torch.from_numpy(u).cuda()
    vals, dfc = disc.assemble(JD, ud)
    disc.assemble_boundary(JD, ud, values=vals, defect=dfc, scale_a=0.7)
    torch.cuda.synchronize()
    eg, ee = parity.entry_errors(vals.cpu().numpy(), ov, rowptr)
    assert eg < TOL and ee < TOL
    eg, ee = parity.entry_errors(dfc.cpu().numpy(), od)
    assert eg < TOL and ee < TOL
    # the increment alone (cancellation-free check of the boundary part)
    eg, _ = parity.entry_errors(vals.cpu().numpy() - ov0, ov - ov0)
    assert eg < 1e-11
    # host vectors + resident Jacobian; bitwise equal to the device-tensor run (owner-computes, fixed order)
    d2 = disc.assemble_resident(JD, u)
    d2 = disc.assemble_boundary(JD, u, defect=d2, scale_a=0.7)
    assert np.array_equal(d2, dfc.cpu().numpy())
    x = np.random.default_rng(0).uniform(-1, 1, u.size)
    y = disc.apply_jacobian(x)
    yr = disc.apply_jacobian(torch.from_numpy(x).cuda(), values=vals).cpu().numpy()
    assert np.array_equal(y, yr)
    # defect only
    d3 = disc.assemble_resident(capi.DEF_A, u)
    d3 = disc.assemble_boundary(capi.DEF_A, u, defect=d3, scale_a=0.7)
    assert np.array_equal(d3, d2)
    disc.close()


def test_boundary_errors():
    coords, conn = meshgen.make_mesh("quad", 4)
    disc = pkg.NavierStokesFV1("u,v,p", "Inner")
    parity.configure(disc)
    disc.set_grid("quad", conn, coords)
    with pytest.raises(pkg.UGError):
        disc.set_boundary_faces(capi.BND_OUTFLOW, [0, 1], [0])
    with pytest.raises(pkg.UGError):
        disc.set_boundary_faces(capi.BND_OUTFLOW, [0], [7])                     # no such side
    with pytest.raises(pkg.UGError):
        disc.set_boundary_faces(capi.BND_INFLOW, [0], [0])                      # inflow without data
    with pytest.raises(pkg.UGError):
        disc.assemble_boundary(capi.JAC_A, np.zeros(disc.num_dofs))             # no resident Jacobian yet
    disc.close()


def test_channel_with_inflow_walls_and_outflow_converges(ora):
    """2-D channel: parabolic inflow (Dirichlet + continuity term), no-slip walls, zero-normal-stress outflow, no pressure
    pinning needed. Picard iteration on the device; the converged state is a root of the oracle's defect including the
    oracle's boundary terms, and the discharge through the outflow equals the inflow."""
    import torch
    nx, ny, H, L = 24, 8, 1.0, 3.0
    coords, conn = meshgen.quad_grid(nx, ny, hi=(L, H))
    nf = 3
    dev = torch.device("cuda", 0)
    disc = pkg.NavierStokesFV1("u,v,p", "Inner")
    parity.configure(disc, upwind="lps", stab="fields", visc=0.02)
    disc.set_grid("quad", conn, coords)
    prof = lambda x, y: (4.0 * y * (H - y) / H ** 2, 0.0)
    left, right = np.isclose(coords[:, 0], 0.0), np.isclose(coords[:, 0], L)
    wall_nodes = np.nonzero(np.isclose(coords[:, 1], 0.0) | np.isclose(coords[:, 1], H))[0]
    in_e, in_s = meshgen.boundary_sides("quad", conn, coords, where=lambda c: np.isclose(c[:, 0], 0.0))
    out_e, out_s = meshgen.boundary_sides("quad", conn, coords, where=lambda c: np.isclose(c[:, 0], L))
    wall = pkg.NavierStokesWall(disc); wall.add(wall_nodes)
    inflow = pkg.NavierStokesInflowFV1(disc)
    inflow.add(prof, np.nonzero(left)[0], coords, sides=(in_e, in_s), conn=conn, elem="quad")
    outflow = pkg.NavierStokesNoNormalStressOutflow(disc); outflow.add(out_e, out_s); outflow.apply()
    dw, vw = wall.dirichlet(); di, vi = inflow.dirichlet()
    dofs = np.concatenate([dw, di]); vals = np.concatenate([vw, vi])
    dofs, first = np.unique(dofs, return_index=True); vals = vals[first]
    inflow.apply()                                                             # registers the continuity term (and its own dofs) ...
    disc.set_dirichlet(dofs)                                                   # ... replaced by the union with the walls
    u = torch.zeros(disc.num_dofs, dtype=torch.float64, device=dev)
    disc.adjust_vector(u, vals)
    rowptr, colind = disc.csr()
    rp, ci = torch.from_numpy(rowptr).to(dev), torch.from_numpy(colind.astype(np.int64)).to(dev)
    hist = []
    for it in range(40):
        d = disc.assemble_resident(JD, u)
        disc.assemble_boundary(JD, u, defect=d)
        disc.adjust_jacobian(); disc.adjust_vector(d)
        hist.append(float(d.norm()))
        if hist[-1] < 1e-9 * hist[0]:
            break
        ptr = disc.resident_jacobian_ptr()

        class _H:
            pass
        h = _H(); h.__cuda_array_interface__ = {"shape": (disc.nnz,), "typestr": "<f8", "data": (ptr, False), "version": 3}
        jv = torch.as_tensor(h, device=dev)
        A = torch.sparse_csr_tensor(rp, ci, jv, size=(u.numel(), u.numel())).to_dense()
        u = u + torch.linalg.solve(A, -d)
    assert hist[-1] < 1e-9 * hist[0]
    uh = u.cpu().numpy()
    p = ora.make_params(elem="quad", upwind="lps", stab="fields", kin_visc=0.02)
    _, od = ora.assemble(p, conn, coords, uh, rowptr, colind, capi.DEF_A)
    ora.fv1_boundary(p, ora.BND_OUTFLOW, out_e, out_s, conn, coords, uh, rowptr, colind, capi.DEF_A, defect=od)
    data = np.array([[prof(*x) for x in row] for row in meshgen.fv1_bf_ips("quad", conn, coords, in_e, in_s)])
    ora.fv1_boundary(p, ora.BND_INFLOW, in_e, in_s, conn, coords, uh, rowptr, colind, capi.DEF_A, data=data, defect=od)
    od[dofs] = 0.0
    assert np.abs(od).max() < 1e-8 * hist[0]
    # discharge: trapezoid rule of u over the outflow column vs the inflow column
    yy = coords[right, 1]; o = np.argsort(yy)
    q_out = np.trapezoid(uh.reshape(-1, nf)[right, 0][o], yy[o])
    yl = coords[left, 1]; ol = np.argsort(yl)
    q_in = np.trapezoid(uh.reshape(-1, nf)[left, 0][ol], yl[ol])
    assert abs(q_out - q_in) < 0.05 * q_in                                      # nodal trapezoid rule, not the discrete balance
    disc.close()
