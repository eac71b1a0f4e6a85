"""host-side mirror of the solution-level tools (plugin_navierstokes_b200/tools.py; reference: incompressible/navier_stokes_tools.h):
point evaluation of P1 / Q1 grid functions and the driven-cavity line evaluation against the reference's literature tables"""
import numpy as np
import pytest

from plugin_navierstokes_b200 import meshgen, tools


@pytest.mark.parametrize("elem", ["tri", "quad"])
def test_evaluate_global_reproduces_the_interpolated_field(elem):
    coords, conn = meshgen.make_mesh(elem, 9, jitter=0.2, seed=4)
    rng = np.random.default_rng(0)
    pts = rng.uniform(0.02, 0.98, (40, 2))
    # a linear field is reproduced exactly on both element types
    u = np.zeros((coords.shape[0], 3))
    u[:, 0] = 0.3 + 2.0 * coords[:, 0] - 1.5 * coords[:, 1]
    u[:, 1] = -0.7 * coords[:, 0]
    got = tools.evaluate_global(u.reshape(-1), coords, conn, 0, pts)
    assert np.allclose(got, 0.3 + 2.0 * pts[:, 0] - 1.5 * pts[:, 1], atol=1e-12)
    got = tools.evaluate_global(u.reshape(-1), coords, conn, 1, pts)
    assert np.allclose(got, -0.7 * pts[:, 0], atol=1e-12)
    # grid nodes return the nodal values; points on element boundaries are found
    got = tools.evaluate_global(u.reshape(-1), coords, conn, 0, coords[::7])
    assert np.allclose(got, u[::7, 0], atol=1e-12)
    with pytest.raises(ValueError):
        tools.evaluate_global(u.reshape(-1), coords, conn, 0, [(1.5, 0.5)])


def test_bilinear_field_on_an_axis_aligned_quad_grid():
    coords, conn = meshgen.quad_grid(8, 8)
    u = np.zeros((coords.shape[0], 3))
    u[:, 0] = coords[:, 0] * coords[:, 1]                        # in the Q1 space of an axis-aligned grid
    pts = np.random.default_rng(1).uniform(0, 1, (30, 2))
    got = tools.evaluate_global(u.reshape(-1), coords, conn, 0, pts)
    assert np.allclose(got, pts[:, 0] * pts[:, 1], atol=1e-12)


def test_driven_cavity_lines_eval_tables():
    """feeding the tabulated profile back gives zero difference at the sample points that are grid nodes; table shapes, sources and
    the lid / wall end points follow navier_stokes_tools.h:578-617"""
    for Re in (100, 400, 1000):
        assert len(tools._VERT[Re]) == 17 and len(tools._HORIZ[Re]) == 17
        assert tools._VERT[Re][0] == 0.0 and tools._VERT[Re][-1] == 1.0          # wall, lid
        assert tools._HORIZ[Re][0] == 0.0 and tools._HORIZ[Re][-1] == 0.0        # walls
    # a grid whose node lines are exactly the Ghia sample positions: u(0.5, y_i) := table -> max diff 0
    xs = np.array(sorted(set(tools.GHIA_HORIZ_X)))
    ys = np.array(sorted(set(tools.GHIA_VERT_Y)))
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    nx, ny = xs.size, ys.size
    conn = np.array([[j * nx + i, j * nx + i + 1, (j + 1) * nx + i + 1, (j + 1) * nx + i] for j in range(ny - 1) for i in range(nx - 1)], dtype=np.int32)
    u = np.zeros((coords.shape[0], 3))
    u[:, 0] = np.interp(coords[:, 1], tools.GHIA_VERT_Y, tools._VERT[100])
    u[:, 1] = np.interp(coords[:, 0], tools.GHIA_HORIZ_X, tools._HORIZ[100])
    lines = []
    out = tools.DrivenCavityLinesEval(u.reshape(-1), coords, conn, 100, log=lines.append)
    assert set(out) == {"Ghia"}
    assert out["Ghia"]["vertical"]["max_diff"] < 1e-14 and out["Ghia"]["horizontal"]["max_diff"] < 1e-14
    assert any("Max Diff" in l for l in lines) and any("Ghia, Re = 100" in l for l in lines)
    assert set(tools.DrivenCavityLinesEval(u.reshape(-1), coords, conn, 1000)) == {"Ghia", "Botella/Peyret"}
    assert tools.DrivenCavityLinesEval(u.reshape(-1), coords, conn, 123) == {}


@pytest.mark.parametrize("elem,jitter", [("tri", 0.2), ("quad", 0.0)])
def test_evaluate_global_cr_reproduces_linear_fields(elem, jitter):
    """a linear field interpolated at the side midpoints is reproduced exactly by the Crouzeix-Raviart evaluation (quadrilaterals:
    rotated bilinear shapes, exact for fields linear in the local coordinates -> parallelograms)"""
    coords, conn = meshgen.make_mesh(elem, 7, jitter=jitter, seed=2)
    es, n_side = meshgen.element_sides(elem, conn)
    mid = np.zeros((n_side, 2)); cnt = np.zeros(n_side)
    for k, sd in enumerate(meshgen.SIDES[elem]):
        np.add.at(mid, es[:, k], coords[conn[:, list(sd)]].mean(axis=1)); np.add.at(cnt, es[:, k], 1)
    mid /= cnt[:, None]
    u = np.zeros(n_side * 2 + conn.shape[0])
    u[0:n_side * 2:2] = 1.0 + 0.5 * mid[:, 0] - 2.0 * mid[:, 1]
    u[1:n_side * 2:2] = -0.25 * mid[:, 0]
    pts = np.random.default_rng(5).uniform(0.03, 0.97, (25, 2))
    assert np.allclose(tools.evaluate_global_cr(u, coords, conn, es, 0, pts), 1.0 + 0.5 * pts[:, 0] - 2.0 * pts[:, 1], atol=1e-12)
    assert np.allclose(tools.evaluate_global_cr(u, coords, conn, es, 1, pts), -0.25 * pts[:, 0], atol=1e-12)


@pytest.mark.parametrize("elem", ["tri", "tet"])
def test_interpolate_cr_to_lagrange(elem):
    """against a plain loop over elements / corners / sides with the SCV ip built from the SCV corners (navier_stokes_tools.h:149-228)"""
    n = 5 if elem == "tri" else 3
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=6)
    es, n_side = meshgen.element_sides(elem, conn)
    dim = coords.shape[1]
    rng = np.random.default_rng(2)
    u = np.concatenate([rng.uniform(-1, 1, n_side * dim), rng.uniform(-1, 1, conn.shape[0])])
    got = tools.interpolateCRToLagrange(u, coords, conn, es, n_side)
    sides = meshgen.SIDES[elem]
    nco = dim + 1
    ref = np.zeros((coords.shape[0], dim)); vs = np.zeros(coords.shape[0])
    import itertools
    for e in range(conn.shape[0]):
        x = coords[conn[e]]
        vol = abs(np.linalg.det(x[1:] - x[0])) / (2.0 if dim == 2 else 6.0)
        for i in range(nco):
            # SCV corners in barycentric coordinates: the node, the midpoints of its edges, (the centres of its faces,) the barycentre
            pts = [np.eye(nco)[i]]
            for j in range(nco):
                if j != i:
                    b = np.zeros(nco); b[[i, j]] = 0.5; pts.append(b)
            if dim == 3:
                for j, k in itertools.combinations([c for c in range(nco) if c != i], 2):
                    b = np.zeros(nco); b[[i, j, k]] = 1.0 / 3.0; pts.append(b)
            pts.append(np.full(nco, 1.0 / nco))
            lam = np.mean(pts, axis=0)
            val = np.zeros(dim)
            for s, sd in enumerate(sides):
                o = [c for c in range(nco) if c not in sd][0]
                val += (1.0 - dim * lam[o]) * u[es[e, s] * dim:es[e, s] * dim + dim]
            ref[conn[e, i]] += vol / nco * val; vs[conn[e, i]] += vol / nco
    ref /= vs[:, None]
    assert np.allclose(got, ref, atol=1e-13)
    # a constant field is reproduced
    uc = np.concatenate([np.tile([0.7, -0.2, 0.4][:dim], n_side), np.zeros(conn.shape[0])])
    assert np.allclose(tools.interpolateCRToLagrange(uc, coords, conn, es, n_side), [0.7, -0.2, 0.4][:dim], atol=1e-13)


@pytest.mark.parametrize("elem", ["tri", "quad", "tet", "hex", "prism"])
def test_drag_lift_on_analytic_fields(elem):
    """DragLift mirror (navier_stokes_tools.h:981-1228) on fields the P1 / Q1 space holds exactly: linear pressure over the whole
    outer boundary of the unit box, Couette shear on the bottom wall"""
    n = 6 if elem in ("tri", "quad") else 3
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.15, seed=3)
    dim = coords.shape[1]
    nf = dim + 1
    a, b, c, nu, rho = 0.7, -0.4, 1.3, 0.05, 1.2
    last = dim - 1
    u = np.zeros((coords.shape[0], nf))
    u[:, dim] = 2.0 + a * coords[:, 0] + b * coords[:, last]
    be, bs = meshgen.boundary_sides(elem, conn)
    drag, lift = tools.DragLift(u.reshape(-1), coords, conn, elem, be, bs, nu, rho, quad_order=2)
    # - int p n_x with the inner normal: int (p(1, .) - p(0, .)) = a ; lift likewise = b
    assert abs(drag - a) < 1e-12 and abs(lift - b) < 1e-12
    # Couette flow u = (c x_last, 0, ..), constant pressure, bottom wall only: inner normal e_last, t = (1, .., 0):
    # drag = nu rho c * area, lift = -p * area
    u = np.zeros((coords.shape[0], nf))
    u[:, 0] = c * coords[:, last]
    u[:, dim] = 0.3
    be, bs = meshgen.boundary_sides(elem, conn, coords, where=lambda x: np.isclose(x[:, last], 0.0))
    drag, lift = tools.DragLift(u.reshape(-1), coords, conn, elem, be, bs, nu, rho, quad_order=2)
    assert abs(drag - nu * rho * c) < 1e-12 and abs(lift + 0.3) < 1e-12


@pytest.mark.parametrize("elem,n", [("quad", 5), ("hex", 3)])
def test_interpolate_cr_to_lagrange_on_quadrilaterals_and_hexahedra(ora, elem, n):
    """rotated bi- / trilinear CR shapes at the SCV ips: (1) on a parallelepiped grid an affine field given at the side centres is
    returned exactly at the interior nodes (their SCV ips lie symmetrically around the node), a constant everywhere;
    (2) the weights are the FV1 SCV volumes of the oracle's geometry (jittered grid)"""
    coords0, conn = meshgen.make_mesh(elem, n)
    dim = coords0.shape[1]
    A = np.array([[1.0, 0.3, 0.1], [0.2, 0.9, -0.2], [0.0, 0.1, 1.1]])[:dim, :dim]
    coords = coords0 @ A.T
    es, n_side = meshgen.element_sides(elem, conn)

    def side_centres(c):
        mid = np.zeros((n_side, dim)); cnt = np.zeros(n_side)
        for k, sd in enumerate(meshgen.SIDES[elem]):
            np.add.at(mid, es[:, k], c[conn[:, list(sd)]].mean(axis=1)); np.add.at(cnt, es[:, k], 1)
        return mid / cnt[:, None]

    def f(X):
        return np.stack([1 + 0.5 * X[:, 0] - 2 * X[:, 1], -0.25 * X[:, 0] + X[:, dim - 1]] + ([0.7 * X[:, 1] - X[:, 2]] if dim == 3 else []), axis=1)
    u = np.concatenate([f(side_centres(coords)).ravel(), np.zeros(conn.shape[0])])
    got = tools.interpolateCRToLagrange(u, coords, conn, es, n_side)
    interior = np.all((coords0 > 1e-9) & (coords0 < 1 - 1e-9), axis=1)
    assert interior.sum() > 0 and np.abs(got - f(coords))[interior].max() < 1e-13
    one = np.concatenate([np.ones(n_side * dim), np.zeros(conn.shape[0])])
    assert np.abs(tools.interpolateCRToLagrange(one, coords, conn, es, n_side) - 1.0).max() < 1e-14
    # (2) brute force with the oracle's SCV volumes on a jittered grid
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=4)
    es, n_side = meshgen.element_sides(elem, conn)
    rng = np.random.default_rng(1)
    u = np.concatenate([rng.uniform(-1, 1, n_side * dim), np.zeros(conn.shape[0])])
    got = tools.interpolateCRToLagrange(u, coords, conn, es, n_side)
    ref = np.zeros((coords.shape[0], dim)); vs = np.zeros(coords.shape[0])
    rc = np.array(meshgen_ref_corners(elem), float)
    for e in range(conn.shape[0]):
        g = ora.fv1_geometry(ora.ELEM[elem], coords[conn[e]])
        for i in range(conn.shape[1]):
            N = tools._cr_shapes_tensor(dim, 0.25 + 0.5 * rc[i])
            val = sum(N[s] * u[es[e, s] * dim:(es[e, s] + 1) * dim] for s in range(es.shape[1]))
            ref[conn[e, i]] += g["vol"][i] * val; vs[conn[e, i]] += g["vol"][i]
    assert np.allclose(got, ref / vs[:, None], atol=1e-13)


def meshgen_ref_corners(elem):
    return {"quad": [(0, 0), (1, 0), (1, 1), (0, 1)],
            "hex": [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]}[elem]


@pytest.mark.parametrize("elem", ["quad", "hex"])
def test_rotated_cr_shapes_are_nodal_at_the_side_centres(ora, elem):
    """the coefficient table of the rotated bi- / trilinear Crouzeix-Raviart shapes (the same rationals as in csrc/ns_fvcr_q.cuh):
    N_s(centre of side t) = delta_st, partition of unity, span {1, x, y, (z,) x^2 - y^2 (, y^2 - z^2)} -- and the oracle, which obtains
    its coefficients by inverting the Vandermonde matrix numerically, evaluates the same functions at its SCVF ips"""
    rc = np.array(meshgen_ref_corners(elem), float)
    dim = rc.shape[1]
    sides = meshgen.SIDES[elem]
    for t, sd in enumerate(sides):
        N = tools._cr_shapes_tensor(dim, rc[list(sd)].mean(axis=0))
        assert np.allclose(N, np.eye(len(sides))[t], atol=1e-15)
    rng = np.random.default_rng(0)
    for _ in range(5):
        assert abs(tools._cr_shapes_tensor(dim, rng.uniform(0, 1, dim)).sum() - 1.0) < 1e-14
    # x^2 + y^2 (+ z^2) is NOT in the span, x^2 - y^2 is: interpolate it from the side centres and compare at a point
    f = lambda p: p[0] ** 2 - p[1] ** 2
    p = rng.uniform(0, 1, dim)
    vals = np.array([f(rc[list(sd)].mean(axis=0)) for sd in sides])
    assert abs(tools._cr_shapes_tensor(dim, p) @ vals - f(p)) < 1e-14
    g = ora.cr_geometry(ora.ELEM[elem], rc)
    for ip in range(g["nip"]):
        assert np.allclose(g["shape"][ip], tools._cr_shapes_tensor(dim, g["lip"][ip]), atol=1e-14)
