"""Oracle geometry / upwind building blocks against hand-derived fixtures and invariants.

The reference has no tests; these pin the restatement of the ugcore conventions
(App. B) that the oracle -- and through it the CUDA path -- relies on.
"""
import numpy as np
import pytest

from tests.conftest import jittered_ref_element

ELEMS = ["tri", "quad", "tet", "hex", "prism"]


def _elem_volume(elem, x):
    if elem == "tri":
        return 0.5 * abs(np.linalg.det(x[1:] - x[0]))
    if elem == "tet":
        return abs(np.linalg.det(x[1:] - x[0])) / 6
    if elem == "quad":
        a, b = x[2] - x[0], x[3] - x[1]
        return 0.5 * abs(a[0] * b[1] - a[1] * b[0])
    if elem == "prism":
        # 3-point (degree 2) triangle rule x 2-point Gauss along the axis: exact for det J of the prism map
        vol = 0.0
        for (a, b) in ((1 / 6, 1 / 6), (2 / 3, 1 / 6), (1 / 6, 2 / 3)):
            for c in (0.5 - 0.5 / np.sqrt(3), 0.5 + 0.5 / np.sqrt(3)):
                lam, dl = np.array([1 - a - b, a, b]), np.array([[-1, -1], [1, 0], [0, 1]], float)
                dN = np.zeros((6, 3))
                for k in range(6):
                    fz, sz = (1 - c, -1.0) if k < 3 else (c, 1.0)
                    dN[k] = [dl[k % 3, 0] * fz, dl[k % 3, 1] * fz, lam[k % 3] * sz]
                vol += np.linalg.det(dN.T @ x) / 12
        return abs(vol)
    # hex: 2x2x2 Gauss quadrature of det J (exact for a trilinear map)
    gp = np.array([0.5 - 0.5 / np.sqrt(3), 0.5 + 0.5 / np.sqrt(3)])
    ref = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    vol = 0.0
    for a in gp:
        for b in gp:
            for c in gp:
                xi = np.array([a, b, c])
                dN = np.zeros((8, 3))
                for k in range(8):
                    f = np.where(ref[k] > 0.5, xi, 1 - xi)
                    s = np.where(ref[k] > 0.5, 1.0, -1.0)
                    dN[k] = [s[0] * f[1] * f[2], f[0] * s[1] * f[2], f[0] * f[1] * s[2]]
                vol += np.linalg.det(dN.T @ x) / 8
    return abs(vol)


def test_unit_square_fixture(ora):
    """SURVEY App. B-2 hand-checkable fixture for the unit quad"""
    g = ora.fv1_geometry(ora.QUAD, jittered_ref_element("quad", amp=0.0))
    assert np.allclose(g["xip"], [[0.5, 0.25], [0.75, 0.5], [0.5, 0.75], [0.25, 0.5]])
    assert np.allclose(g["normal"], [[0.5, 0], [0, 0.5], [-0.5, 0], [0, -0.5]])
    assert np.allclose(g["shape"][0], [0.375, 0.375, 0.125, 0.125])
    assert np.allclose(g["vol"], 0.25)
    assert list(g["frm"]) == [0, 1, 2, 3] and list(g["to"]) == [1, 2, 3, 0]


def test_unit_cube_fixture(ora):
    g = ora.fv1_geometry(ora.HEX, jittered_ref_element("hex", amp=0.0))
    assert np.allclose(g["xip"][0], [0.5, 0.25, 0.25])
    assert np.allclose(g["normal"][0], [0.25, 0, 0])
    assert np.isclose(g["c0c2sq"][0], 0.5)
    assert np.allclose(g["vol"], 0.125)
    # every SCVF of the unit cube has area 1/4 and is axis aligned
    assert np.allclose(np.linalg.norm(g["normal"], axis=1), 0.25)


def test_unit_prism_fixture(ora):
    """hand-derived: reference prism (volume 1/2). SCVF of edge (0, 1) = rectangle [edge midpoint (1/2, 0, 0), bottom-triangle centre
    (1/3, 1/3, 0), barycentre (1/3, 1/3, 1/2), centre of the quadrilateral y = 0 (1/2, 0, 1/2)]: ip = (5/12, 1/6, 1/4), area vector
    (1/6, 1/12, 0) pointing from corner 0 to corner 1; every SCV holds 1/12; the vertical edge (0, 3) has the SCVF [edge midpoint,
    centres of the quadrilaterals y = 0 and x = 0, barycentre] with normal (0, 0, 1/6)"""
    g = ora.fv1_geometry(ora.PRISM, jittered_ref_element("prism", amp=0.0))
    assert g["nsh"] == 6 and g["nip"] == 9
    assert list(g["frm"]) == [0, 1, 2, 0, 1, 2, 3, 4, 5] and list(g["to"]) == [1, 2, 0, 3, 4, 5, 4, 5, 3]
    assert np.allclose(g["xip"][0], [5 / 12, 1 / 6, 1 / 4]) and np.allclose(g["normal"][0], [1 / 6, 1 / 12, 0])
    assert np.allclose(g["vol"], 1 / 12)
    lam = np.array([5 / 12, 5 / 12, 1 / 6])
    assert np.allclose(g["shape"][0], np.concatenate([0.75 * lam, 0.25 * lam]))
    # edge (0, 3): corners (0, 0, 1/2), (1/2, 0, 1/2), (1/3, 1/3, 1/2), (0, 1/2, 1/2) -> planar, area 1/2 |(1/3, 1/3) x (-1/2, 1/2)| = 1/6
    assert np.allclose(g["xip"][3], [5 / 24, 5 / 24, 1 / 2]) and np.allclose(g["normal"][3], [0, 0, 1 / 6])
    assert np.isclose(g["c0c2sq"][3], 2 / 9)


@pytest.mark.parametrize("elem", ELEMS)
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fv1_geometry_invariants(ora, elem, seed):
    x = jittered_ref_element(elem, seed=seed, amp=0.12, scale=0.7, shift=0.3)
    g = ora.fv1_geometry(ora.ELEM[elem], x)
    dim = g["dim"]
    # SCV volumes tile the element
    assert np.isclose(g["vol"].sum(), _elem_volume(elem, x), rtol=1e-13)
    assert (g["vol"] > 0).all()
    # normals point from `from` to `to`
    for ip in range(g["nip"]):
        assert g["normal"][ip] @ (x[g["to"][ip]] - x[g["frm"][ip]]) > 0
    # shapes: partition of unity, gradients sum to zero and reproduce linear fields
    assert np.allclose(g["shape"].sum(axis=1), 1.0)
    assert np.allclose(g["ggrad"].sum(axis=1), 0.0, atol=1e-13)
    for ip in range(g["nip"]):
        assert np.allclose(g["ggrad"][ip].T @ x, np.eye(dim), atol=1e-12)
    # closure of each SCV: sum of outward SCVF normals == -(boundary part); for the whole element the
    # signed sum over SCVFs of (from:+n, to:-n) vanishes identically -> check per-corner flux of a
    # constant field through SCVFs equals minus the flux through the element-boundary part of the SCV.
    # (2-D check via polygon closure)
    if dim == 2:
        nsh = g["nsh"]
        bary = x.mean(axis=0)
        for c in range(nsh):
            tot = np.zeros(2)
            for ip in range(g["nip"]):
                if g["frm"][ip] == c:
                    tot += g["normal"][ip]
                if g["to"][ip] == c:
                    tot -= g["normal"][ip]
            # boundary part: two half edges at corner c, outward normals
            prv, nxt = x[(c - 1) % nsh], x[(c + 1) % nsh]
            for a, b in ((0.5 * (prv + x[c]), x[c]), (x[c], 0.5 * (x[c] + nxt))):
                e = b - a
                tot += np.array([e[1], -e[0]])
            assert np.allclose(tot, 0.0, atol=1e-13), (c, tot, bary)


@pytest.mark.parametrize("elem", ELEMS)
def test_ip_is_mean_of_scvf_corners(ora, elem):
    """global ip = mean of the global SCVF corners; local ip consistent with shapes"""
    x = jittered_ref_element(elem, seed=5, amp=0.1)
    g = ora.fv1_geometry(ora.ELEM[elem], x)
    if elem in ("tri", "tet"):
        # affine map: shapes at the local ip interpolate the global ip
        assert np.allclose(g["shape"] @ x, g["xip"], atol=1e-14)
    # local ip of edge (a,b): quad (mid+centre)/2, hex (mid+2 faces+centre)/4
    if elem == "hex":
        assert np.allclose(g["lip"][0], [0.5, 0.25, 0.25])
        assert np.allclose(g["lip"][4], [0.25, 0.25, 0.5])
    if elem == "tet":
        # edge (0,1): mean of edge mid, two face barycentres, barycentre
        mid = np.array([0.5, 0, 0]); b = np.array([0.25, 0.25, 0.25])
        f1 = np.array([1, 1, 0]) / 3; f2 = np.array([1, 0, 1]) / 3
        assert np.allclose(g["lip"][0], (mid + b + f1 + f2) / 4)


@pytest.mark.parametrize("elem", ELEMS)
def test_side_ray_intersection(ora, elem):
    rng = np.random.default_rng(7)
    x = jittered_ref_element(elem, seed=3, amp=0.1 if elem in ("tri", "tet") else 0.0)   # planar quadrilateral sides
    E = ora.ELEM[elem]
    g = ora.fv1_geometry(E, x)
    for ip in range(g["nip"]):
        for _ in range(5):
            v = rng.uniform(-1, 1, g["dim"])
            ok, side, gc, lc = ora.side_ray_intersection(E, x, g["xip"][ip], v, positive=False)
            assert ok
            # the cut lies upstream: gc = xip + t v with t <= 0
            t = (gc - g["xip"][ip]) @ v / (v @ v)
            assert t < 0
            assert np.allclose(g["xip"][ip] + t * v, gc, atol=1e-13)
            # local cut maps to the global cut (affine / undistorted elements)
            N = _p1_shapes(elem, lc)
            assert np.allclose(N @ x, gc, atol=1e-12)
            # and lies on the boundary of the reference element
            assert _on_ref_boundary(elem, lc)
            ok2, side2, gc2, _ = ora.side_ray_intersection(E, x, g["xip"][ip], v, positive=True)
            assert ok2 and (gc2 - g["xip"][ip]) @ v > 0


def _p1_shapes(elem, xi):
    if elem == "tri":
        return np.array([1 - xi[0] - xi[1], xi[0], xi[1]])
    if elem == "tet":
        return np.array([1 - xi.sum(), xi[0], xi[1], xi[2]])
    if elem == "quad":
        x, y = xi
        return np.array([(1 - x) * (1 - y), x * (1 - y), x * y, (1 - x) * y])
    if elem == "prism":
        lam = np.array([1 - xi[0] - xi[1], xi[0], xi[1]])
        return np.concatenate([lam * (1 - xi[2]), lam * xi[2]])
    ref = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    return np.array([np.prod(np.where(r > 0.5, xi, 1 - xi)) for r in ref])


def _on_ref_boundary(elem, xi, tol=1e-10):
    if elem in ("quad", "hex"):
        return (np.abs(xi) < tol).any() or (np.abs(xi - 1) < tol).any()
    if elem == "prism":
        return (np.abs(xi) < tol).any() or abs(xi[0] + xi[1] - 1) < tol or abs(xi[2] - 1) < tol
    return (np.abs(xi) < tol).any() or abs(xi.sum() - 1) < tol


@pytest.mark.parametrize("elem", ELEMS)
@pytest.mark.parametrize("upwind", ["full", "skewed", "lps", "positive"])
def test_upwind_partition_of_unity(ora, elem, upwind):
    """sum_sh up_sh + sum_ip up_ip = 1 wherever there is flow"""
    rng = np.random.default_rng(11)
    x = jittered_ref_element(elem, seed=4, amp=0.1)
    E = ora.ELEM[elem]
    for _ in range(4):
        vel = rng.uniform(-1, 1, (ora.NIP[E], ora.DIM[E]))
        sh, ipm, ln = ora.fv1_upwind(E, upwind, x, vel)
        assert np.allclose(sh.sum(axis=1) + ipm.sum(axis=1), 1.0, atol=1e-12)
        assert (ln > 0).all()
        if upwind in ("full", "skewed"):
            assert ((sh == 0) | (sh == 1)).all()
        if upwind == "positive":
            assert (sh >= -1e-14).all() and (ipm >= 0).all()


def test_full_upwind_picks_upstream_corner(ora):
    x = jittered_ref_element("quad", amp=0.0)
    vel = np.tile([1.0, 0.0], (4, 1))        # flow in +x
    sh, _, ln = ora.fv1_upwind(ora.QUAD, "full", x, vel)
    # SCVF 0 (edge 0->1, n=+x): flux>0 -> from=0 ; SCVF 2 (edge 2->3, n=-x): flux<0 -> to=3
    assert sh[0, 0] == 1 and sh[2, 3] == 1
    # SCVF 1 (n=+y): flux == 0 -> else branch -> `to` (upwind.cpp:166-170)
    assert sh[1, 2] == 1
    assert np.isclose(ln[0], np.hypot(0.5, 0.25))


def test_lps_unit_square(ora):
    """ray from ip (0.5,0.25) against u=(1,0) hits the side x=0 at y=0.25: shapes 0.75/0.25 on corners 0/3"""
    x = jittered_ref_element("quad", amp=0.0)
    vel = np.tile([1.0, 0.0], (4, 1))
    sh, _, ln = ora.fv1_upwind(ora.QUAD, "lps", x, vel)
    assert np.allclose(sh[0], [0.75, 0, 0, 0.25])
    assert np.isclose(ln[0], 0.5)
    sk, _, _ = ora.fv1_upwind(ora.QUAD, "skewed", x, vel)
    assert np.allclose(sk[0], [1, 0, 0, 0])


def test_zero_velocity_guards(ora):
    x = jittered_ref_element("hex", amp=0.0)
    vel = np.zeros((12, 3))
    for up in ("skewed", "lps"):
        sh, _, ln = ora.fv1_upwind(ora.HEX, up, x, vel)
        assert (sh == 0).all() and (ln == 1).all()          # upwind.cpp:407-413, 531-537
    sh, ipm, ln = ora.fv1_upwind(ora.HEX, "positive", x, vel)
    assert np.allclose(sh.sum(axis=1), 1.0) and (ipm == 0).all()      # 1/2-1/2 on from/to, :677-684


@pytest.mark.parametrize("elem", ["tri", "tet"])
def test_cr_geometry(ora, elem):
    x = jittered_ref_element(elem, seed=2, amp=0.1)
    g = ora.cr_geometry(ora.ELEM[elem], x)
    dim = g["dim"]
    assert np.isclose(g["vol"].sum(), _elem_volume(elem, x))
    # outward side normals sum to zero (closed surface)
    assert np.allclose(g["scv_normal"].sum(axis=0), 0, atol=1e-14)
    # CR shapes: partition of unity, value 1 at own side barycentre
    assert np.allclose(g["shape"].sum(axis=1), 1.0)
    assert np.allclose(g["ggrad"].sum(axis=1), 0.0, atol=1e-13)
    for ip in range(g["nip"]):
        f, t = g["frm"][ip], g["to"][ip]
        assert g["normal"][ip] @ (g["scv_xip"][t] - g["scv_xip"][f]) > 0
        # gradient reproduces linear fields when evaluated with side-barycentre values
        assert np.allclose(g["ggrad"][ip].T @ g["scv_xip"], np.eye(dim), atol=1e-12)
    # each SCV (cone side <-> barycentre) is closed: outward side normal + signed SCVF normals = 0
    for s in range(g["nsh"]):
        tot = g["scv_normal"][s].copy()
        for ip in range(g["nip"]):
            if g["frm"][ip] == s:
                tot += g["normal"][ip]
            if g["to"][ip] == s:
                tot -= g["normal"][ip]
        assert np.allclose(tot, 0, atol=1e-14)


def test_cr_geometry_unit_square_and_cube_fixtures(ora):
    """hand-derived CRFVGeometry values on the non-simplex reference elements: SCV = triangle / pyramid between a side and the
    barycentre (1/4 of the square, 1/6 of the cube), outward unit side normals; SCVF of corner 0 of the square: segment (0, 0) ->
    (1/2, 1/2) between side 0 (bottom) and side 3 (left), normal (-1/2, 1/2); SCVF of edge (0, 1) of the cube: triangle
    (0,0,0), (1,0,0), (1/2,1/2,1/2) between side 0 (bottom) and side 1 (y = 0), area vector (0, -1/4, 1/4), ip (1/2, 1/6, 1/6)"""
    q = ora.cr_geometry(ora.QUAD, jittered_ref_element("quad", amp=0.0))
    assert q["nsh"] == 4 and q["nip"] == 4 and np.allclose(q["vol"], 0.25)
    assert np.allclose(q["scv_normal"], [[0, -1], [1, 0], [0, 1], [-1, 0]])
    assert (q["frm"][0], q["to"][0]) == (0, 3) and np.allclose(q["normal"][0], [-0.5, 0.5]) and np.allclose(q["xip"][0], [0.25, 0.25])
    h = ora.cr_geometry(ora.HEX, jittered_ref_element("hex", amp=0.0))
    assert h["nsh"] == 6 and h["nip"] == 12 and np.allclose(h["vol"], 1 / 6)
    assert np.allclose(h["scv_normal"], [[0, 0, -1], [0, -1, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, 0, 1]])
    assert (h["frm"][0], h["to"][0]) == (0, 1) and np.allclose(h["normal"][0], [0, -0.25, 0.25]) and np.allclose(h["xip"][0], [0.5, 1 / 6, 1 / 6])
    # the rotated shapes at that ip: N_0 (bottom) = N_1 (y = 0) by symmetry, gradients sum to zero
    assert np.isclose(h["shape"][0][0], h["shape"][0][1]) and np.abs(h["ggrad"][0].sum(axis=0)).max() < 1e-14
