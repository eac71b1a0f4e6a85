"""DiscConstraintFVCR (fvcr/disc_constraint_fvcr.h:164-1198), default configuration -- linear-upwind and
linear-pressure correction of the FVCR defect. Oracle checks on the CPU, device parity through the C ABI (marked gpu)."""
import numpy as np
import pytest

from plugin_navierstokes_b200 import meshgen


def _mesh(elem, n, seed=2, jitter=0.15):
    coords, conn = meshgen.make_mesh(elem, n, jitter=jitter, seed=seed)
    es, n_side = meshgen.element_sides(elem, conn)
    dim = coords.shape[1]
    cen = np.zeros((n_side, dim)); cnt = np.zeros(n_side)
    for k, s in enumerate(meshgen.SIDES[elem]):
        np.add.at(cen, es[:, k], coords[conn[:, list(s)]].mean(axis=1)); np.add.at(cnt, es[:, k], 1)
    return coords, conn, es, n_side, cen / cnt[:, None], cnt


@pytest.mark.parametrize("elem,n", [("tri", 6), ("tet", 3)])
def test_linear_upwind_makes_full_upwind_exact_for_a_linear_field(ora, elem, n):
    """u = A x + b: the Crouzeix-Raviart interpolant is exact, the side gradients equal A, so the reconstructed upwind velocity
    u_base + acGrad (x_ip - x_base) is the velocity at the ip: FullUpwind defect + correction = NoUpwind (central) defect"""
    coords, conn, es, n_side, cen, cnt = _mesh(elem, n)
    dim = coords.shape[1]
    A = np.array([[0.3, -0.2, 0.5], [0.1, 0.4, -0.6], [0.7, 0.2, -0.1]])[:dim, :dim]
    vel = cen @ A.T + np.array([0.2, -0.1, 0.3])[:dim]
    u = np.concatenate([vel.ravel(), np.full(conn.shape[0], 0.7)])                 # constant pressure: no pressure correction
    E = ora.ELEM[elem]
    rowptr, colind = ora.fvcr_csr(E, es, n_side)
    pf = ora.make_params(disc="fvcr", elem=elem, upwind="full", kin_visc=0.02)
    pn = ora.make_params(disc="fvcr", elem=elem, upwind="no", kin_visc=0.02)
    _, df = ora.assemble(pf, conn, coords, u, rowptr, colind, ora.DEF_A, elem_sides=es, n_side=n_side)
    _, dn = ora.assemble(pn, conn, coords, u, rowptr, colind, ora.DEF_A, elem_sides=es, n_side=n_side)
    assert np.abs(df - dn).max() > 1e-4
    dc = ora.fvcr_constraint_defect(E, conn, coords, es, n_side, u, defect=df.copy())
    assert np.abs(dc - dn).max() < 1e-13 * np.abs(dn).max() + 1e-15
    # the pressure part alone vanishes for a constant pressure, the pressure rows are never touched
    dp = ora.fvcr_constraint_defect(E, conn, coords, es, n_side, u, lin_upwind=False)
    assert np.abs(dp).max() < 1e-15
    assert np.all((dc - df)[n_side * dim:] == 0)


@pytest.mark.parametrize("elem,n", [("tri", 6), ("tet", 3)])
def test_conservation_scaling_and_zero_gradient_boundaries(ora, elem, n):
    coords, conn, es, n_side, cen, cnt = _mesh(elem, n, seed=5)
    dim = coords.shape[1]
    rng = np.random.default_rng(1)
    u = np.concatenate([(np.sin(2 * cen) + 0.3 * rng.uniform(-1, 1, cen.shape)).ravel(), rng.uniform(-1, 1, conn.shape[0])])
    E = ora.ELEM[elem]
    d = ora.fvcr_constraint_defect(E, conn, coords, es, n_side, u)
    mom = d[:n_side * dim].reshape(-1, dim)
    assert np.abs(mom).max() > 1e-3 and np.abs(mom.sum(axis=0)).max() < 1e-14     # every flux correction enters +from / -to
    # upwind part ~ s_a (flux) and pressure part ~ s_a: the whole correction is linear in s_a
    d2 = ora.fvcr_constraint_defect(E, conn, coords, es, n_side, u, s_a=0.25)
    assert np.allclose(d2, 0.25 * d, rtol=1e-13, atol=1e-16)
    # sum of the two parts
    du = ora.fvcr_constraint_defect(E, conn, coords, es, n_side, u, lin_pressure=False)
    dp = ora.fvcr_constraint_defect(E, conn, coords, es, n_side, u, lin_upwind=False)
    assert np.allclose(du + dp, d, rtol=1e-12, atol=1e-15)
    # zero-gradient boundary: elements with a flagged side contribute nothing
    bsides = np.nonzero(cnt == 1)[0]
    dz = ora.fvcr_constraint_defect(E, conn, coords, es, n_side, u, zero_grad_sides=bsides)
    touched = np.zeros(n_side, bool)
    for e in range(conn.shape[0]):
        if not np.isin(es[e], bsides).any():
            touched[es[e]] = True
    assert np.all(dz[:n_side * dim].reshape(-1, dim)[~touched] == 0) and np.abs(dz - d).max() > 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("elem,n", [("tri", 9), ("tet", 4)])
@pytest.mark.parametrize("up,pr,zero", [(True, True, False), (True, False, False), (False, True, False), (True, True, True)])
def test_device_constraint_matches_the_oracle(ora, elem, n, up, pr, zero):
    import torch
    import plugin_navierstokes_b200 as pkg
    from plugin_navierstokes_b200 import capi
    coords, conn, es, n_side, cen, cnt = _mesh(elem, n, seed=7, jitter=0.2)
    dim = coords.shape[1]
    rng = np.random.default_rng(2)
    u = np.concatenate([(np.cos(3 * cen) + 0.2 * rng.uniform(-1, 1, cen.shape)).ravel(), rng.uniform(-1, 1, conn.shape[0])])
    disc = pkg.NavierStokesFVCR("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    disc.set_kinematic_viscosity(0.02)
    disc.set_upwind("full")
    disc.set_grid(elem, conn, coords, es, n_side)
    zs = np.nonzero(cnt == 1)[0][::2] if zero else None
    con = pkg.DiscConstraintFVCR(disc, up, False, pr, False, False, zero_grad_sides=zs)
    E = ora.ELEM[elem]
    rowptr, colind = ora.fvcr_csr(E, es, n_side)
    p = ora.make_params(disc="fvcr", elem=elem, upwind="full", kin_visc=0.02)
    _, od = ora.assemble(p, conn, coords, u, rowptr, colind, ora.DEF_A, elem_sides=es, n_side=n_side)
    corr = ora.fvcr_constraint_defect(E, conn, coords, es, n_side, u, s_a=0.6, lin_upwind=up, lin_pressure=pr, zero_grad_sides=zs)
    _, d = disc.assemble(capi.DEF_A, u)
    d0 = d.copy()
    con.adjust_defect(d, u, scale_stiff=0.6)
    assert np.abs(corr).max() > 1e-4
    assert np.abs((d - d0) - corr).max() <= 1e-12 * np.abs(corr).max()
    assert np.abs(d - (od + corr)).max() <= 1e-12 * np.abs(od).max()
    # device tensors: bitwise the same correction (owner-computes, fixed order)
    ud = torch.from_numpy(u).cuda()
    dd = torch.from_numpy(d0).cuda()
    con.adjust_defect(dd, ud, scale_stiff=0.6)
    torch.cuda.synchronize()
    assert np.array_equal(dd.cpu().numpy(), d)
    with pytest.raises(pkg.UGError):
        pkg.DiscConstraintFVCR(disc, True, True, True, False, False)               # Jacobian variant: not on the device path
    disc.close()
