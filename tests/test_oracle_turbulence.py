"""Oracle restatement of FV1SmagorinskyTurbViscData (fv1/turbulent_viscosity_fv1.h:200-383, _impl.h:504-616,755-762,819-852) and of
the diagnostics vorticityFV1 / kineticEnergy / cflNumber (navier_stokes_tools.h:386-525,731-965): analytic checks. CPU only."""
import numpy as np
import pytest

from plugin_navierstokes_b200 import meshgen


def _node_volumes(ora, elem, conn, coords):
    vol = np.zeros(coords.shape[0])
    for e in range(conn.shape[0]):
        g = ora.fv1_geometry(ora.ELEM[elem], coords[conn[e]])
        np.add.at(vol, conn[e], g["vol"])
    return vol


@pytest.mark.parametrize("elem,n", [("quad", 6), ("tri", 6), ("hex", 4), ("tet", 3)])
def test_smagorinsky_of_a_linear_field_on_a_uniform_grid(ora, elem, n):
    """u = A x on an unjittered grid: the midpoint rule over the closed SCV surface is exact, so D = sym(A) at every interior
    vertex and nu_t = c vol^(2/dim) sqrt(2 S:S); with the BF closure the boundary vertices follow to the accuracy of the nodal
    value used on the boundary faces"""
    coords, conn = meshgen.make_mesh(elem, n)
    dim = coords.shape[1]
    A = np.array([[0.3, -0.2, 0.5], [0.1, 0.4, -0.6], [0.7, 0.2, -0.1]])[:dim, :dim]
    u = np.zeros((coords.shape[0], dim + 1)); u[:, :dim] = coords @ A.T
    S = 0.5 * (A + A.T)
    vol = _node_volumes(ora, elem, conn, coords)
    nut, ipv = ora.fv1_smagorinsky(ora.ELEM[elem], conn, coords, u, c=0.1, kin_visc=0.01)
    ref = 0.1 * vol ** (2.0 / dim) * np.sqrt(2.0 * (S ** 2).sum())
    inner = np.all((coords > 1e-9) & (coords < 1 - 1e-9), axis=1)
    assert inner.any() and np.abs(nut[inner] / ref[inner] - 1).max() < 1e-12
    # interpolation to the SCVF ips: partition of unity -> min/max bounds, + laminar viscosity
    nt = nut[conn]
    assert np.all(ipv >= nt.min(axis=1)[:, None] + 0.01 - 1e-15) and np.all(ipv <= nt.max(axis=1)[:, None] + 0.01 + 1e-15)
    # rigid rotation + translation: no deformation, no eddy viscosity (interior)
    W = np.array([[0, -1.0, 0.5], [1.0, 0, -0.3], [-0.5, 0.3, 0]])[:dim, :dim]
    u[:, :dim] = coords @ W.T + 0.7
    nut2, _ = ora.fv1_smagorinsky(ora.ELEM[elem], conn, coords, u, c=0.1)
    assert np.abs(nut2[inner]).max() < 1e-13
    # turbulence-zero boundary: flagged vertices are 0, the others unchanged in the interior
    be, bs = meshgen.boundary_sides(elem, conn)
    bnodes = np.nonzero(~inner)[0]
    u[:, :dim] = coords @ A.T
    nut3, _ = ora.fv1_smagorinsky(ora.ELEM[elem], conn, coords, u, c=0.1, belem=be, bside=bs, zero_nodes=bnodes)
    assert np.all(nut3[bnodes] == 0) and np.array_equal(nut3[inner], nut[inner])


def test_vorticity_of_a_rotation_and_a_shear(ora):
    for elem, n in [("quad", 5), ("tri", 5), ("hex", 3), ("tet", 3)]:
        coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=3)
        dim = coords.shape[1]
        u = np.zeros((coords.shape[0], dim + 1))
        u[:, 0] = -1.5 * coords[:, 1] + 0.2 * coords[:, 0]
        u[:, 1] = 0.5 * coords[:, 0] + 0.1
        w = ora.fv1_vorticity(ora.ELEM[elem], conn, coords, u)
        assert np.abs(w - 2.0).max() < 1e-12                     # d_x v - d_y u = 0.5 + 1.5, linear field: exact on any grid


def test_kinetic_energy_and_cfl_of_a_uniform_flow(ora):
    for elem, n in [("tri", 4), ("tet", 2)]:
        coords, conn = meshgen.make_mesh(elem, n, jitter=0.15, seed=1)
        dim = coords.shape[1]
        es, ns = meshgen.element_sides(elem, conn)
        vel = np.array([0.3, -0.4, 1.2])[:dim]
        u = np.concatenate([np.tile(vel, ns), np.zeros(conn.shape[0])])
        ke, cfl = ora.fvcr_diagnostics(ora.ELEM[elem], conn, coords, es, u, dt=0.1)
        assert abs(ke - vel @ vel) < 1e-13
        # brute force CFL from the side barycentres
        best = 0.0
        sides = meshgen.SIDES[elem]
        for e in range(conn.shape[0]):
            xs = [coords[conn[e][list(s)]].mean(axis=0) for s in sides]
            for i in range(len(xs)):
                for j in range(i + 1, len(xs)):
                    d = xs[i] - xs[j]
                    best = max(best, 0.1 * abs(d @ vel) / (d @ d))
        assert abs(cfl - best) < 1e-12 * best
