"""SURVEY 8f-4 on the device, through the C ABI, against the oracle: FV1SmagorinskyTurbViscData as provider of the per-ip viscosity
import (nsb_turbulent_viscosity) -- nodal eddy viscosity, the import table it fills, and the assembly that uses it -- and the
diagnostics vorticityFV1 / kineticEnergy / cflNumber (nsb_diagnostic)."""
import numpy as np
import pytest

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen
from tests import parity
from tests.parity import TOL

pytestmark = pytest.mark.gpu
JD = capi.JAC_A | capi.DEF_A


@pytest.mark.parametrize("elem,n", [("tri", 8), ("quad", 8), ("tet", 4), ("hex", 5)])
@pytest.mark.parametrize("zero_bnd", [False, True])
def test_smagorinsky_viscosity_and_the_assembly_that_uses_it(ora, elem, n, zero_bnd):
    import torch
    coords, conn, u = parity.make_case(elem, n, seed=11)
    u = u.reshape(-1)
    dim = coords.shape[1]
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    parity.configure(disc, upwind="lps", stab="fields", visc=5e-3)
    disc.set_grid(elem, conn, coords)
    turb = pkg.FV1SmagorinskyTurbViscData(disc, c=0.17)
    be = bs = zn = None
    if zero_bnd:
        # walls: the x = min and y = max boundary; vertices of those sides are "in the subset" except the ones at x = min, y = min
        be, bs = meshgen.boundary_sides(elem, conn, coords, where=lambda c: np.isclose(c[:, 0], coords[:, 0].min()) | np.isclose(c[:, 1], coords[:, 1].max()))
        on = np.zeros(coords.shape[0], bool)
        for q in range(len(be)):
            on[conn[be[q]][list(meshgen.SIDES[elem][bs[q]])]] = True
        on &= ~np.isclose(coords[:, 1], coords[:, 1].min())       # these keep an evaluated nu_t and therefore use the BF closure
        zn = np.nonzero(on)[0]
        turb.set_turbulence_zero_bnd(be, bs, zn)
    nut_ref, ipv_ref = ora.fv1_smagorinsky(ora.ELEM[elem], conn, coords, u, c=0.17, kin_visc=5e-3, belem=be, bside=bs, zero_nodes=zn)
    nut = turb.update(u, want_nodal=True)                                         # host vectors
    assert np.abs(nut - nut_ref).max() <= 1e-13 * np.abs(nut_ref).max()
    if zero_bnd:
        assert np.all(nut[zn] == 0) and np.abs(nut_ref).max() > 0
    # the assembly now runs with nu(ip) = interpolated nu_t + nu: oracle with the same table as per-ip import
    rowptr, colind = ora.fv1_csr(ora.ELEM[elem], conn, coords.shape[0])
    p = ora.make_params(elem=elem, upwind="lps", stab="fields", kin_visc=5e-3)
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, JD, ip_data=dict(visc=ipv_ref))
    ov0, _ = ora.assemble(p, conn, coords, u, rowptr, colind, JD)
    assert np.abs(ov - ov0).max() > 1e-4 * np.abs(ov0).max()                      # the eddy viscosity matters
    for mode in (capi.SCATTER_GATHER, capi.SCATTER_ATOMIC):
        vals, dfc = disc.assemble(JD, u, scatter_mode=mode)
        eg, ee = parity.entry_errors(vals, ov, rowptr)
        assert eg < TOL and ee < TOL
        eg, ee = parity.entry_errors(dfc, od)
        assert eg < TOL and ee < TOL
    # device tensors: same nodal values bit for bit (fixed-order gather), new state -> new table
    ud = torch.from_numpy(u).cuda()
    nd = turb.update(ud, want_nodal=True)
    torch.cuda.synchronize()
    assert np.array_equal(nd.cpu().numpy(), nut)
    u2 = u * 1.5
    turb.update(u2)
    _, ipv2 = ora.fv1_smagorinsky(ora.ELEM[elem], conn, coords, u2, c=0.17, kin_visc=5e-3, belem=be, bside=bs, zero_nodes=zn)
    ov2, _ = ora.assemble(p, conn, coords, u2, rowptr, colind, JD, ip_data=dict(visc=ipv2))
    vals2, _ = disc.assemble(JD, u2)
    eg, ee = parity.entry_errors(vals2, ov2, rowptr)
    assert eg < TOL and ee < TOL
    # switching the model off returns to the constant viscosity (and to the owner-computes split path)
    turb.disable()
    vals3, _ = disc.assemble(JD, u)
    eg, ee = parity.entry_errors(vals3, ov0, rowptr)
    assert eg < TOL and ee < TOL
    disc.close()


@pytest.mark.parametrize("elem,n", [("tri", 9), ("quad", 9), ("tet", 4), ("hex", 5)])
def test_vorticity(ora, elem, n):
    import torch
    coords, conn, u = parity.make_case(elem, n, seed=2)
    u = u.reshape(-1)
    dim = coords.shape[1]
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    parity.configure(disc)
    disc.set_grid(elem, conn, coords)
    ref = ora.fv1_vorticity(ora.ELEM[elem], conn, coords, u)
    w = disc.vorticity(u)
    assert np.abs(w - ref).max() <= 1e-12 * np.abs(ref).max()
    wd = disc.vorticity(torch.from_numpy(u).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(wd.cpu().numpy(), w)
    with pytest.raises(pkg.UGError):
        disc.kinetic_energy(u)                                                     # Crouzeix-Raviart diagnostics need an FVCR grid
    disc.close()


@pytest.mark.parametrize("elem,n", [("tri", 12), ("tet", 5)])
def test_kinetic_energy_and_cfl_number(ora, elem, n):
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=4)
    es, n_side = meshgen.element_sides(elem, conn)
    dim = coords.shape[1]
    rng = np.random.default_rng(3)
    u = np.concatenate([rng.uniform(-1, 1, n_side * dim) + 0.4, rng.uniform(-1, 1, conn.shape[0])])
    disc = pkg.NavierStokesFVCR("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    disc.set_kinematic_viscosity(1e-2)
    disc.set_upwind("full")
    disc.set_grid(elem, conn, coords, es, n_side)
    ke_ref, cfl_ref = ora.fvcr_diagnostics(ora.ELEM[elem], conn, coords, es, u, dt=0.05)
    ke, cfl = disc.kinetic_energy(u), disc.cfl_number(u, 0.05)
    assert abs(ke - ke_ref) <= 1e-13 * ke_ref and abs(cfl - cfl_ref) <= 1e-13 * cfl_ref
    assert disc.kinetic_energy(u) == ke                                            # fixed-order reduction
    with pytest.raises(pkg.UGError):
        disc.vorticity(u)
    disc.close()
