"""Synthetic meshes and states for the five BASELINE.json configurations.

Pure numpy input generators (no arithmetic of the hot path lives here).  Node numbering of the
structured grids is x-fastest; element corner order follows the UG4 reference elements
(SURVEY.md App. B-1): quads counter-clockwise, hexahedra bottom face ccw then top face.
"""
import itertools

import numpy as np

ELEM_NSH = {"tri": 3, "quad": 4, "tet": 4, "hex": 8, "prism": 6}
ELEM_DIM = {"tri": 2, "quad": 2, "tet": 3, "hex": 3, "prism": 3}

# sides of the reference elements (corner lists), SURVEY App. B-1
SIDES = {
    "tri": [(0, 1), (1, 2), (2, 0)],
    "quad": [(0, 1), (1, 2), (2, 3), (3, 0)],
    "tet": [(0, 2, 1), (1, 2, 3), (0, 3, 2), (0, 1, 3)],
    "hex": [(0, 3, 2, 1), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7), (4, 5, 6, 7)],
    "prism": [(0, 2, 1), (0, 1, 4, 3), (1, 2, 5, 4), (2, 0, 3, 5), (3, 4, 5)],
}


def _jitter(coords, shape, h, jitter, seed):
    """displace interior grid nodes by at most jitter*h per coordinate"""
    if not jitter:
        return coords
    rng = np.random.default_rng(seed)
    dim = coords.shape[1]
    idx = np.indices([s for s in shape[::-1]]).reshape(dim, -1)[::-1]  # x fastest
    interior = np.ones(coords.shape[0], dtype=bool)
    for d in range(dim):
        interior &= (idx[d] > 0) & (idx[d] < shape[d] - 1)
    disp = rng.uniform(-1.0, 1.0, size=coords.shape) * (jitter * np.asarray(h))
    coords = coords.copy()
    coords[interior] += disp[interior]
    return coords


def quad_grid(nx, ny, lo=(0.0, 0.0), hi=(1.0, 1.0), jitter=0.0, seed=0):
    xs = np.linspace(lo[0], hi[0], nx + 1)
    ys = np.linspace(lo[1], hi[1], ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    n0 = (i + (nx + 1) * j).ravel()
    conn = np.stack([n0, n0 + 1, n0 + 1 + (nx + 1), n0 + (nx + 1)], axis=1).astype(np.int32)
    h = ((hi[0] - lo[0]) / nx, (hi[1] - lo[1]) / ny)
    return _jitter(coords, (nx + 1, ny + 1), h, jitter, seed), conn


def tri_grid(nx, ny, lo=(0.0, 0.0), hi=(1.0, 1.0), jitter=0.0, seed=0, hole=None):
    """structured quads split into two ccw triangles; optional (cx, cy, r) hole removes triangles
    whose centroid lies inside the disc (config 2)."""
    coords, q = quad_grid(nx, ny, lo, hi, jitter, seed)
    t1 = q[:, [0, 1, 2]]
    t2 = q[:, [0, 2, 3]]
    conn = np.empty((2 * q.shape[0], 3), dtype=np.int32)
    conn[0::2], conn[1::2] = t1, t2
    if hole is not None:
        cx, cy, r = hole
        cen = coords[conn].mean(axis=1)
        keep = (cen[:, 0] - cx) ** 2 + (cen[:, 1] - cy) ** 2 > r * r
        conn = conn[keep]
        used = np.unique(conn)
        remap = -np.ones(coords.shape[0], dtype=np.int64)
        remap[used] = np.arange(used.size)
        coords = coords[used]
        conn = remap[conn].astype(np.int32)
    return coords, conn


def hex_grid(nx, ny, nz, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), jitter=0.0, seed=0):
    xs = np.linspace(lo[0], hi[0], nx + 1)
    ys = np.linspace(lo[1], hi[1], ny + 1)
    zs = np.linspace(lo[2], hi[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    sx, sy = 1, nx + 1
    sz = (nx + 1) * (ny + 1)
    n0 = (i * sx + j * sy + k * sz).ravel().astype(np.int64)
    conn = np.stack([n0, n0 + sx, n0 + sx + sy, n0 + sy,
                     n0 + sz, n0 + sx + sz, n0 + sx + sy + sz, n0 + sy + sz], axis=1).astype(np.int32)
    h = ((hi[0] - lo[0]) / nx, (hi[1] - lo[1]) / ny, (hi[2] - lo[2]) / nz)
    return _jitter(coords, (nx + 1, ny + 1, nz + 1), h, jitter, seed), conn


def _kuhn_tets():
    """6 positively oriented tets of the unit cube (hex corner numbering), all sharing diagonal 0-6"""
    corner = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
    lookup = {tuple(c): i for i, c in enumerate(corner)}
    tets = []
    for perm in itertools.permutations(range(3)):
        p = np.zeros(3, dtype=int)
        ids = [lookup[tuple(p)]]
        for a in perm:
            p = p.copy()
            p[a] = 1
            ids.append(lookup[tuple(p)])
        x = corner[ids].astype(float)
        if np.linalg.det(x[1:] - x[0]) < 0:
            ids[1], ids[2] = ids[2], ids[1]
        tets.append(ids)
    return np.array(tets)


def tet_grid(nx, ny, nz, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), jitter=0.0, seed=0):
    coords, hexes = hex_grid(nx, ny, nz, lo, hi, jitter, seed)
    kt = _kuhn_tets()
    conn = hexes[:, kt].reshape(-1, 4).astype(np.int32)
    return coords, conn


def prism_grid(nx, ny, nz, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), jitter=0.0, seed=0):
    """every cell of the structured grid split into two prisms along the diagonal 0-2 of its bottom face (prism corner order:
    bottom triangle counter-clockwise, then the top triangle above it)"""
    coords, hexes = hex_grid(nx, ny, nz, lo, hi, jitter, seed)
    conn = hexes[:, [[0, 1, 2, 4, 5, 6], [0, 2, 3, 4, 6, 7]]].reshape(-1, 6).astype(np.int32)
    return coords, conn


def make_mesh(elem, n, **kw):
    """convenience: n elements per direction (tri/tet: n cells, split)."""
    if elem == "quad":
        return quad_grid(n, n, **kw)
    if elem == "tri":
        return tri_grid(n, n, **kw)
    if elem == "hex":
        return hex_grid(n, n, n, **kw)
    if elem == "tet":
        return tet_grid(n, n, n, **kw)
    if elem == "prism":
        return prism_grid(n, n, n, **kw)
    raise ValueError(elem)


def element_sides(elem, conn):
    """Unique numbering of element sides (edges in 2-D, faces in 3-D) for the FVCR dof layout.
    Returns (elem_sides [n_elem, nside] int32, n_side). Side k of an element is the reference side k."""
    sides = SIDES[elem]
    ne = conn.shape[0]
    width = max(len(sd) for sd in sides)                    # prisms mix triangles and quadrilaterals: pad the keys with -1
    keys = np.stack([np.sort(np.pad(conn[:, list(sd)], ((0, 0), (width - len(sd), 0)), constant_values=-1), axis=1) for sd in sides], axis=1)  # [ne, nside, nco]
    flat = keys.reshape(ne * len(sides), -1)
    _, inv = np.unique(flat, axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    return inv.reshape(ne, len(sides)).astype(np.int32), int(inv.max()) + 1


# ----------------------------------------------------------------------------------------------
# states (velocity..., pressure) per node, SURVEY §8(d)
# ----------------------------------------------------------------------------------------------
def state_cavity2d(coords, seed=1, noise=0.01):
    rng = np.random.default_rng(seed)
    x, y = coords[:, 0], coords[:, 1]
    xi = 1.0 + noise * rng.uniform(-1, 1, size=(coords.shape[0], 2))
    u = np.sin(np.pi * x) * np.cos(np.pi * y) * xi[:, 0]
    v = -np.cos(np.pi * x) * np.sin(np.pi * y) * xi[:, 1]
    p = np.cos(np.pi * x) * np.cos(np.pi * y)
    return np.stack([u, v, p], axis=1)


def state_channel2d(coords, seed=2, noise=0.01, H=0.41, umax=0.3):
    rng = np.random.default_rng(seed)
    y = coords[:, 1]
    u = 4 * umax * y * (H - y) / H ** 2 + noise * rng.uniform(-1, 1, coords.shape[0])
    v = noise * rng.uniform(-1, 1, coords.shape[0])
    p = 0.1 * (2.2 - coords[:, 0])
    return np.stack([u, v, p], axis=1)


def state_vortex3d(coords, seed=3, noise=0.01):
    rng = np.random.default_rng(seed)
    x, y, z = (np.pi * coords[:, d] for d in range(3))
    xi = 1.0 + noise * rng.uniform(-1, 1, size=(coords.shape[0], 3))
    u = np.sin(x) * np.cos(y) * np.cos(z) * xi[:, 0] + 0.05
    v = -0.5 * np.cos(x) * np.sin(y) * np.cos(z) * xi[:, 1] - 0.03
    w = -0.5 * np.cos(x) * np.cos(y) * np.sin(z) * xi[:, 2] + 0.02
    p = np.cos(x) * np.cos(y) * np.cos(z)
    return np.stack([u, v, w, p], axis=1)


def state_taylor_green(coords, t=0.0, nu=1.0 / 1600, seed=5, noise=0.0):
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    f = np.exp(-3 * nu * t)
    u = np.sin(x) * np.cos(y) * np.cos(z) * f
    v = -np.cos(x) * np.sin(y) * np.cos(z) * f
    w = np.zeros_like(u)
    p = (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2) / 16 * f * f
    s = np.stack([u, v, w, p], axis=1)
    if noise:
        s = s + noise * np.random.default_rng(seed).uniform(-1, 1, s.shape)
    return s


def random_state(n, nf, seed=0, scale=1.0):
    return scale * np.random.default_rng(seed).uniform(-1, 1, size=(n, nf))


# ----------------------------------------------------------------------------------------------------
# integration points of the FV1 geometry: where the reference evaluates its UserData imports
# ----------------------------------------------------------------------------------------------------
_EDGES = {"tri": [(0, 1), (1, 2), (2, 0)], "quad": [(0, 1), (1, 2), (2, 3), (3, 0)],
          "tet": [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)],
          "hex": [(0, 1), (1, 2), (2, 3), (3, 0), (0, 4), (1, 5), (2, 6), (3, 7), (4, 5), (5, 6), (6, 7), (7, 4)],
          "prism": [(0, 1), (1, 2), (2, 0), (0, 3), (1, 4), (2, 5), (3, 4), (4, 5), (5, 3)]}
_FACES = {"tet": [(0, 2, 1), (1, 2, 3), (0, 3, 2), (0, 1, 3)],
          "hex": [(0, 3, 2, 1), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7), (4, 5, 6, 7)],
          "prism": SIDES["prism"]}


def fv1_scvf_ips(elem, conn, coords):
    """global positions of the SCVF integration points, [n_elem][nip][dim]: the centroid of the SCVF (2-D: midpoint of edge
    midpoint and element barycentre; 3-D: mean of edge midpoint, the two adjacent face centres and the barycentre)"""
    x = coords[conn]                                        # [n_elem][nsh][dim]
    cen = x.mean(axis=1)
    out = []
    for (a, b) in _EDGES[elem]:
        mid = 0.5 * (x[:, a] + x[:, b])
        if coords.shape[1] == 2:
            out.append(0.5 * (mid + cen))
        else:
            fs = [f for f in _FACES[elem] if a in f and b in f]
            fc = [x[:, list(f)].mean(axis=1) for f in fs]
            out.append(0.25 * (mid + fc[0] + cen + fc[1]))
    return np.stack(out, axis=1)


def fv1_scv_ips(elem, conn, coords):
    """global positions of the SCV integration points, [n_elem][nsh][dim]: the corners (ugcore FV1Geometry SCV::global_ip)"""
    return coords[conn].copy()


# ----------------------------------------------------------------------------------------------------
# boundary sides and boundary-face integration points (FV1Geometry BF; boundary discs of SURVEY 8f-1)
# ----------------------------------------------------------------------------------------------------
def boundary_sides(elem, conn, coords=None, where=None):
    """(element, local side) pairs of the sides that belong to exactly one element. where(centres [n][dim]) -> bool mask selects
    a part of the boundary by the side centres (the role of ugcore's boundary subsets)."""
    es, n_side = element_sides(elem, conn)
    count = np.bincount(es.reshape(-1), minlength=n_side)
    be, bs = np.nonzero(count[es] == 1)
    if where is not None:
        cen = np.stack([coords[conn[be, k]] for k in range(conn.shape[1])], axis=1)      # [n][nsh][dim]
        sc = np.zeros((be.size, coords.shape[1]))
        for s, corners in enumerate(SIDES[elem]):
            m = bs == s
            sc[m] = cen[m][:, list(corners)].mean(axis=1)
        keep = np.asarray(where(sc), dtype=bool)
        be, bs = be[keep], bs[keep]
    return be.astype(np.int32), bs.astype(np.int32)


def fv1_bf_ips(elem, conn, coords, belem, bside):
    """global positions of the boundary-face ips of the given sides, [n_side][4][dim] (slot j = side corner j; unused slots 0):
    mean of the BF corners -- 2-D [corner, edge midpoint]; 3-D [corner, midpoint to the next side corner, side centre, midpoint to
    the previous side corner]"""
    dim = coords.shape[1]
    out = np.zeros((len(belem), 4, dim))
    x = coords[conn[belem]]                                 # [n][nsh][dim]
    for s, corners in enumerate(SIDES[elem]):
        m = np.asarray(bside) == s
        if not m.any():
            continue
        xs = x[m][:, list(corners)]                         # [n][ns][dim]
        ns = len(corners)
        for j in range(ns):
            if dim == 2:
                out[m, j] = 0.5 * (xs[:, j] + 0.5 * (xs[:, j] + xs[:, 1 - j]))
            else:
                nx, pv = (j + 1) % ns, (j + ns - 1) % ns
                out[m, j] = 0.25 * (xs[:, j] + 0.5 * (xs[:, j] + xs[:, nx]) + xs.mean(axis=1) + 0.5 * (xs[:, j] + xs[:, pv]))
    return out
