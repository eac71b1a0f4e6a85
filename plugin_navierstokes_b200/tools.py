"""Host-side mirror of the solution-level evaluation tools of the plugin (incompressible/navier_stokes_tools.h).

DrivenCavityLinesEval (navier_stokes_tools.h:571-713, registered in incompressible_navier_stokes_plugin.cpp:202) evaluates the
velocity of a 2-D lid-driven-cavity solution at the sample points of the literature tables the reference carries (Ghia, Ghia & Shin
1982; Botella & Peyret 1998 for Re = 1000) -- u on the vertical line x = 0.5, v on the horizontal line y = 0.5 -- and reports the
maximum and the average difference (DrivenCavityEvalAtPoints, :539-569). These tables are the only known-answer data the reference
holds for the path this package accelerates; tests/test_gpu_cavity.py uses them as the solution-level check of the whole chain
(assembly, boundary conditions, Dirichlet post-pass, resident Jacobian, linear solve).

The grid function is P1 / Q1 on triangles / quadrilaterals in the FV1 dof layout node * 3 + fct (u, v, p); point evaluation
follows ugcore's GlobalGridFunctionNumberData::evaluate_global: find the element containing the point, map to local coordinates,
evaluate the Lagrange shapes."""
import numpy as np

# sample positions of the Ghia tables (navier_stokes_tools.h:578-591)
GHIA_VERT_X = 0.5
GHIA_VERT_Y = (0.0000, 0.0547, 0.0625, 0.0703, 0.1016, 0.1719, 0.2813, 0.4531, 0.5000, 0.6172, 0.7344, 0.8516, 0.9531, 0.9609, 0.9688,
               0.9766, 1.0000)
GHIA_HORIZ_Y = 0.5
GHIA_HORIZ_X = (0.0000, 0.0625, 0.0703, 0.0781, 0.0938, 0.1563, 0.2266, 0.2344, 0.5000, 0.8047, 0.8594, 0.9063, 0.9453, 0.9531, 0.9609,
                0.9688, 1.0000)
# u(x = 0.5, y) and v(x, y = 0.5) per Reynolds number (navier_stokes_tools.h:580-598; Botella :614-617)
_VERT = {
    100: (0., -0.03717, -0.04192, -0.04775, -0.06434, -0.10150, -0.15662, -0.2109, -0.20581, -0.13641, 0.00332, 0.23151, 0.68717, 0.73722,
          0.78871, 0.84123, 1.),
    400: (0., -0.08186, -0.09266, -0.10338, -0.14612, -0.24299, -0.32726, -0.17119, -0.11477, 0.02135, 0.16256, 0.29093, 0.55892, 0.61756,
          0.68439, 0.75837, 1.),
    1000: (0., -0.18109, -0.20196, -0.22220, -0.29730, -0.38289, -0.27805, -0.10648, -0.06080, 0.05702, 0.18719, 0.33304, 0.46604, 0.51117,
           0.57492, 0.65928, 1.),
}
_HORIZ = {
    100: (0.00000, 0.09233, 0.10091, 0.10890, 0.12317, 0.16077, 0.17507, 0.17527, 0.05454, -0.24533, -0.22445, -0.16914, -0.10313, -0.08864,
          -0.07391, -0.05906, 0.00000),
    400: (0.00000, 0.18360, 0.19713, 0.20920, 0.22965, 0.28124, 0.30203, 0.30174, 0.05186, -0.38598, -0.44993, -0.3827, -0.22847, -0.19254,
          -0.15663, -0.12146, 0.00000),
    1000: (0.00000, 0.27485, 0.29012, 0.30353, 0.32627, 0.37095, 0.33075, 0.32235, 0.02526, -0.31966, -0.42665, -0.51550, -0.39188, -0.33714,
           -0.27669, -0.21388, 0.00000),
}
_VERT_BOTELLA_1000 = (0.0000000, -0.1812881, -0.2023300, -0.2228955, -0.3004561, -0.3885691, -0.2803696, -0.1081999, -0.0620561, 0.0570178,
                      0.1886747, 0.3372212, 0.4723329, 0.5169277, 0.5808359, 0.6644227, 1.0000000)
_HORIZ_BOTELLA_1000 = (0.0000000, 0.2807056, 0.2962703, 0.3099097, 0.3330442, 0.3769189, 0.3339924, 0.3253592, 0.0257995, -0.3202137,
                       -0.4264545, -0.5264392, -0.4103754, -0.3553213, -0.2936869, -0.2279225, 0.0000000)


def _local_coordinates(xe, pt):
    """local coordinates of pt in the triangle / quadrilateral with corners xe (reference numbering); None when outside"""
    tol = 1e-10
    if xe.shape[0] == 3:
        T = np.array([xe[1] - xe[0], xe[2] - xe[0]]).T
        xi = np.linalg.solve(T, pt - xe[0])
        return xi if xi.min() >= -tol and xi.sum() <= 1 + tol else None
    xi = np.array([0.5, 0.5])
    for _ in range(30):                                   # Newton on the bilinear map
        x, y = xi
        N = np.array([(1 - x) * (1 - y), x * (1 - y), x * y, (1 - x) * y])
        dN = np.array([[-(1 - y), -(1 - x)], [(1 - y), -x], [y, x], [-y, (1 - x)]])
        r = N @ xe - pt
        if np.abs(r).max() < 1e-14:
            break
        xi = xi - np.linalg.solve((dN.T @ xe).T, r)
    return xi if xi.min() >= -tol and xi.max() <= 1 + tol else None


def evaluate_global(u, coords, conn, fct, points, nf=3):
    """value of component fct of the P1 / Q1 grid function u (layout node * nf + fct) at each point"""
    u = np.asarray(u, dtype=np.float64).reshape(-1, nf)
    xe_all = coords[conn]                                 # [n_elem][nsh][2]
    lo, hi = xe_all.min(axis=1), xe_all.max(axis=1)
    out = np.empty(len(points))
    for i, pt in enumerate(np.asarray(points, dtype=np.float64)):
        cand = np.nonzero(np.all((lo <= pt + 1e-12) & (hi >= pt - 1e-12), axis=1))[0]
        for e in cand:
            xi = _local_coordinates(xe_all[e], pt)
            if xi is None:
                continue
            if conn.shape[1] == 3:
                N = np.array([1 - xi[0] - xi[1], xi[0], xi[1]])
            else:
                x, y = xi
                N = np.array([(1 - x) * (1 - y), x * (1 - y), x * y, (1 - x) * y])
            out[i] = N @ u[conn[e], fct]
            break
        else:
            raise ValueError("evaluate_global: point %s is outside the grid" % (pt,))
    return out


def evaluate_global_cr(u, coords, conn, elem_sides, fct, points, dim=2):
    """value of velocity component fct of a Crouzeix-Raviart grid function (FVCR layout side * dim + fct) at each point. Triangles:
    sum over the element sides of u_side * (1 - 2 lambda_o), lambda_o = barycentric coordinate of the corner opposite the side;
    quadrilaterals: the rotated bilinear shapes nodal at the side midpoints (span {1, x, y, x^2 - y^2})"""
    from . import meshgen
    u = np.asarray(u, dtype=np.float64)
    quad = conn.shape[1] == 4
    sides = meshgen.SIDES["quad" if quad else "tri"]
    opp = None if quad else [[c for c in range(3) if c not in sd][0] for sd in sides]
    xe_all = coords[conn]
    lo, hi = xe_all.min(axis=1), xe_all.max(axis=1)
    out = np.empty(len(points))
    for i, pt in enumerate(np.asarray(points, dtype=np.float64)):
        cand = np.nonzero(np.all((lo <= pt + 1e-12) & (hi >= pt - 1e-12), axis=1))[0]
        vals = []
        for e in cand:
            xi = _local_coordinates(xe_all[e], pt)
            if xi is None:
                continue
            if quad:
                x, y = xi
                q = x * x - y * y
                N = [0.75 + x - 2 * y - q, -0.25 + y + q, -0.25 + x - q, 0.75 - 2 * x + y + q]
            else:
                lam = np.array([1 - xi[0] - xi[1], xi[0], xi[1]])
                N = [1.0 - 2.0 * lam[opp[s]] for s in range(3)]
            vals.append(sum(u[elem_sides[e, s] * dim + fct] * N[s] for s in range(len(sides))))
        if not vals:
            raise ValueError("evaluate_global_cr: point %s is outside the grid" % (pt,))
        out[i] = np.mean(vals)                            # CR functions jump across sides: a point on a side takes the mean of its elements
    return out


def interpolateCRToLagrange(u_cr, coords, conn, elem_sides, n_side):
    """mirror of interpolateCRToLagrange (navier_stokes_tools.h:53-229, registered as "CRToLagrange"): nodal (Lagrange P1) velocities
    from a Crouzeix-Raviart function on triangles / tetrahedra (and, through `_cr_to_lagrange_tensor`, quadrilaterals / hexahedra). Per element and corner i the CR function is evaluated at the local ip of the FV1
    sub-control volume of the corner (:196-201), weighted with the SCV volume (:195, :205) and divided by the summed nodal volume at
    the end (:216-228). u_cr: FVCR layout side * dim + d (pressures behind the velocities are ignored). Returns [n_node][dim]."""
    from . import meshgen
    coords = np.asarray(coords, dtype=np.float64)
    dim = coords.shape[1]
    nco = dim + 1
    if conn.shape[1] == 2 ** dim:
        return _cr_to_lagrange_tensor(u_cr, coords, conn, elem_sides, n_side)
    elem = "tri" if dim == 2 else "tet"
    if conn.shape[1] != nco:
        raise ValueError("interpolateCRToLagrange: triangles / quadrilaterals / tetrahedra / hexahedra")
    sides = meshgen.SIDES[elem]
    opp = [[c for c in range(nco) if c not in sd][0] for sd in sides]
    # barycentric coordinates of the SCV ip of corner i = mean of the SCV corners (node, edge midpoints, (face centres,) barycentre)
    own, other = (7.0 / 12.0, 5.0 / 24.0) if dim == 2 else (15.0 / 32.0, 17.0 / 96.0)
    vel = np.asarray(u_cr, dtype=np.float64)[:n_side * dim].reshape(n_side, dim)
    x = coords[conn]                                       # [ne][nco][dim]
    e = x[:, 1:, :] - x[:, :1, :]
    vol = np.abs(np.linalg.det(e)) / (2.0 if dim == 2 else 6.0)
    scv = vol / nco                                        # simplex: every SCV holds 1 / (dim + 1) of the element
    out = np.zeros((coords.shape[0], dim))
    vsum = np.zeros(coords.shape[0])
    us = vel[elem_sides]                                   # [ne][nside][dim]
    for i in range(nco):
        lam = np.full(nco, other); lam[i] = own
        shape = np.array([1.0 - dim * lam[opp[s]] for s in range(len(sides))])      # CR shapes at the SCV ip
        np.add.at(out, conn[:, i], scv[:, None] * np.einsum("s,esd->ed", shape, us))
        np.add.at(vsum, conn[:, i], scv)
    return out / vsum[:, None]


def _cr_shapes_tensor(dim, xi):
    """rotated bi- / trilinear Crouzeix-Raviart shapes nodal at the side centres (span {1, x, y, (z,) x^2 - y^2 (, y^2 - z^2)})"""
    if dim == 2:
        x, y = xi
        q = x * x - y * y
        return np.array([0.75 + x - 2 * y - q, -0.25 + y + q, -0.25 + x - q, 0.75 - 2 * x + y + q])
    x, y, z = xi
    b = np.array([1.0, x, y, z, x * x - y * y, y * y - z * z])
    C = np.array([[2, 2, 2, -7, -2, -4], [2, 2, -7, 2, -2, 2], [-1, -1, 2, 2, 4, 2], [-1, 2, -1, 2, -2, 2], [2, -7, 2, 2, 4, 2], [-1, 2, 2, -1, -2, -4]]) / 3.0
    return C @ b


def _cr_to_lagrange_tensor(u_cr, coords, conn, elem_sides, n_side):
    """interpolateCRToLagrange on quadrilaterals / hexahedra: the SCV of corner i is the image of the reference quadrant / octant at
    that corner (local ip = its centre), its volume the integral of det J over it (2-point Gauss per direction, exact)"""
    dim = coords.shape[1]
    elem = "quad" if dim == 2 else "hex"
    rc = _REF_CORNERS[elem]
    x = coords[conn]                                       # [ne][nco][dim]
    vel = np.asarray(u_cr, dtype=np.float64)[:n_side * dim].reshape(n_side, dim)
    us = vel[elem_sides]                                   # [ne][nside][dim]
    out = np.zeros((coords.shape[0], dim)); vsum = np.zeros(coords.shape[0])
    g = 0.25 / np.sqrt(3.0)
    import itertools
    for i in range(rc.shape[0]):
        centre = 0.25 + 0.5 * rc[i]                       # centre of the quadrant / octant adjacent to corner i
        vol = np.zeros(conn.shape[0])
        for sg in itertools.product((-g, g), repeat=dim):
            _, dN = _lagrange(elem, centre + np.array(sg))
            J = np.einsum("kd,ekj->edj", dN, x)           # [ne][dim][dim]
            vol += np.abs(np.linalg.det(J)) * 0.25 ** dim
        shape = _cr_shapes_tensor(dim, centre)
        np.add.at(out, conn[:, i], vol[:, None] * np.einsum("s,esd->ed", shape, us))
        np.add.at(vsum, conn[:, i], vol)
    return out / vsum[:, None]


_REF_CORNERS = {
    "tri": np.array([(0, 0), (1, 0), (0, 1)], dtype=np.float64),
    "quad": np.array([(0, 0), (1, 0), (1, 1), (0, 1)], dtype=np.float64),
    "tet": np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], dtype=np.float64),
    "hex": np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)], dtype=np.float64),
    "prism": np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1)], dtype=np.float64),
}


def _lagrange(elem, xi):
    """P1 / Q1 shapes and local gradients at the local point xi"""
    rc = _REF_CORNERS[elem]
    if elem in ("tri", "tet"):
        N = np.concatenate([[1.0 - xi.sum()], xi])
        dN = np.vstack([-np.ones(len(xi)), np.eye(len(xi))])
        return N, dN
    if elem == "prism":                                    # P1 on the triangle x P1 along the axis
        lam, dl, z = np.array([1.0 - xi[0] - xi[1], xi[0], xi[1]]), np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]]), xi[2]
        N = np.concatenate([lam * (1 - z), lam * z])
        dN = np.vstack([np.hstack([dl * (1 - z), -lam[:, None]]), np.hstack([dl * z, lam[:, None]])])
        return N, dN
    f = np.where(rc > 0.5, xi, 1.0 - xi)                   # [nsh][dim] factors
    sg = np.where(rc > 0.5, 1.0, -1.0)
    N = f.prod(axis=1)
    dN = np.empty_like(f)
    for d in range(rc.shape[1]):
        other = np.delete(f, d, axis=1).prod(axis=1)
        dN[:, d] = sg[:, d] * other
    return N, dN


def _side_rule(n_corner, order):
    """quadrature on the reference side: Gauss-Legendre on [0, 1], its tensor product on the unit square, or a symmetric rule on the
    unit triangle (degree 1 or 2), exact for polynomials of the requested order -> (points [n][dim-1], weights [n])"""
    ng = max(1, order // 2 + 1)
    g, w = np.polynomial.legendre.leggauss(ng)
    g, w = 0.5 * (g + 1.0), 0.5 * w
    if n_corner == 2:
        return g[:, None], w
    if n_corner == 4:
        P = np.array([(a, b) for a in g for b in g]); W = np.array([wa * wb for wa in w for wb in w])
        return P, W
    if order <= 1:
        return np.array([[1.0 / 3.0, 1.0 / 3.0]]), np.array([0.5])
    if order <= 2:
        return np.array([[1.0 / 6.0, 1.0 / 6.0], [2.0 / 3.0, 1.0 / 6.0], [1.0 / 6.0, 2.0 / 3.0]]), np.full(3, 1.0 / 6.0)
    raise ValueError("DragLift: triangle sides are integrated with rules up to order 2")


def DragLift(u, coords, conn, elem, belem, bside, kin_visco, density, quad_order=2):
    """mirror of DragLift(u, "u,v(,w),p", BndSubsets, InnerSubsets, kinVisco, density, quadOrder) (navier_stokes_tools.h:981-1228) for
    FV1 grid functions (layout node * (dim + 1) + fct): boundary integral over the given sides (belem, bside) = (element, local side),
    e.g. from meshgen.boundary_sides(where=...), with the INNER normal n (:1122-1124) and the tangent t = (n[dim-1], .., -n[0])
    (:1127-1129):   drag += w det ( nu rho (grad u n) . t  n[dim-1] - p n[0] ),   lift -= w det ( nu rho (grad u n) . t  n[0] + p n[dim-1] )
    (:1200-1206). Returns [drag, lift]."""
    from . import meshgen
    coords = np.asarray(coords, dtype=np.float64)
    dim = coords.shape[1]
    nf = dim + 1
    uu = np.asarray(u, dtype=np.float64).reshape(-1, nf)
    rc = _REF_CORNERS[elem]
    drag = lift = 0.0
    for e, sd in zip(np.asarray(belem), np.asarray(bside)):
        cs = list(meshgen.SIDES[elem][sd])
        xe = coords[conn[e]]                               # element corners
        ue = uu[conn[e]]
        xs, ls = xe[cs], rc[cs]                            # side corners, global and local
        P, W = _side_rule(len(cs), quad_order)
        for q, wq in zip(P, W):
            # point on the reference side -> local element coordinates -> shapes; side Jacobian for the surface measure
            if len(cs) == 2:
                sh = np.array([1.0 - q[0], q[0]]); dsh = np.array([[-1.0], [1.0]])
            elif len(cs) == 3:
                sh = np.array([1.0 - q[0] - q[1], q[0], q[1]]); dsh = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
            else:
                a, b = q
                sh = np.array([(1 - a) * (1 - b), a * (1 - b), a * b, (1 - a) * b])
                dsh = np.array([[-(1 - b), -(1 - a)], [(1 - b), -a], [b, a], [-b, (1 - a)]])
            xi = sh @ ls
            JTs = dsh.T @ xs                               # [dim-1][dim]
            det = np.sqrt(np.linalg.det(JTs @ JTs.T))      # SqrtGramDeterminant
            if dim == 2:
                nrm = np.array([JTs[0, 1], -JTs[0, 0]])
            else:
                nrm = np.cross(JTs[0], JTs[1])
            nrm /= np.linalg.norm(nrm)
            if nrm @ (xs.mean(axis=0) - xe.mean(axis=0)) < 0:      # outer normal of the element at this side ...
                nrm = -nrm
            nrm = -nrm                                     # ... turned into the inner one (:1124)
            tng = np.zeros(dim); tng[0] = nrm[dim - 1]; tng[dim - 1] = -nrm[0]
            N, dN = _lagrange(elem, xi)
            JT = dN.T @ xe                                 # [dim][dim]
            G = dN @ np.linalg.inv(JT).T                   # global gradients [nsh][dim]
            grad_vel = ue[:, :dim].T @ G                   # (d1, d2) = d u_d1 / d x_d2
            pressure = N @ ue[:, dim]
            flux = grad_vel @ nrm
            shear = kin_visco * density * (flux @ tng)
            drag += wq * det * (shear * nrm[dim - 1] - pressure * nrm[0])
            lift -= wq * det * (shear * nrm[0] + pressure * nrm[dim - 1])
    return [drag, lift]


def _eval_at_points(u, coords, conn, fct, points, reference, elem_sides=None):
    """DrivenCavityEvalAtPoints (navier_stokes_tools.h:539-569): measured values, reference values, max and average difference"""
    val = evaluate_global(u, coords, conn, fct, points) if elem_sides is None else evaluate_global_cr(u, coords, conn, elem_sides, fct, points)
    ref = np.asarray(reference, dtype=np.float64)
    diff = np.abs(ref - val)
    return {"positions": np.asarray(points), "measure": val, "reference": ref, "max_diff": float(diff.max()), "average_diff": float(diff.mean())}


def DrivenCavityLinesEval(u, coords, conn, Re, vel_cmp=(0, 1), log=None, elem_sides=None):
    """mirror of DrivenCavityLinesEval(u, {"u","v"}, Re): returns {source: {"vertical": {...}, "horizontal": {...}}} with the tables the
    reference prints (Ghia for Re in 100 / 400 / 1000, Botella & Peyret for Re = 1000); an unknown Re gives an empty dict, as the
    reference prints nothing then. log: optional callable receiving the lines the reference writes with UG_LOG.
    elem_sides: the grid function is Crouzeix-Raviart on triangles (FVCR layout), as the reference's function accepts both spaces."""
    Re = int(Re)
    out = {}
    vert_pts = [(GHIA_VERT_X, y) for y in GHIA_VERT_Y]
    horiz_pts = [(x, GHIA_HORIZ_Y) for x in GHIA_HORIZ_X]
    sources = []
    if Re in _VERT:
        sources.append(("Ghia", _VERT[Re], _HORIZ[Re]))
    if Re == 1000:
        sources.append(("Botella/Peyret", _VERT_BOTELLA_1000, _HORIZ_BOTELLA_1000))
    for name, vert, horiz in sources:
        res = {"vertical": _eval_at_points(u, coords, conn, vel_cmp[0], vert_pts, vert, elem_sides),
               "horizontal": _eval_at_points(u, coords, conn, vel_cmp[1], horiz_pts, horiz, elem_sides)}
        out[name] = res
        if log:
            for line, what in (("vertical", "u values on a vertical line through x = 0.5"), ("horizontal", "v values on a horizontal line through y = 0.5")):
                log("  ------ %s, Re = %d: %s  ------" % (name, Re, what))
                r = res[line]
                for i in range(len(r["measure"])):
                    log("%2d  (%.4f, %.4f)  %.8f  %.7f  %.3e" % (i + 1, r["positions"][i][0], r["positions"][i][1], r["measure"][i], r["reference"][i],
                                                                 abs(r["measure"][i] - r["reference"][i])))
                log("\t     Max Diff: %g" % r["max_diff"])
                log("\t Average Diff: %g" % r["average_diff"])
    return out
