"""ctypes binding of the CPU oracle (oracle/ns_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/ns_oracle.h).  Import only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libns_oracle.so")

TRI, QUAD, TET, HEX, PRISM = 0, 1, 2, 3, 4
ELEM = {"tri": TRI, "quad": QUAD, "tet": TET, "hex": HEX, "prism": PRISM}
UPWIND = {None: 0, "none": 0, "no": 1, "full": 2, "skewed": 3, "lps": 4, "linearprofileskewed": 4,
          "positive": 5, "pos": 5}
STAB = {"fields": 0, "flow": 1, "none": 2}
DIFF = {"raw": 0, "fivepoint": 1, "cor": 2}
JAC_A, DEF_A, JAC_M, DEF_M, RHS = 1, 2, 4, 8, 16
NSH = {TRI: 3, QUAD: 4, TET: 4, HEX: 8, PRISM: 6}
NIP = {TRI: 3, QUAD: 4, TET: 6, HEX: 12, PRISM: 9}
DIM = {TRI: 2, QUAD: 2, TET: 3, HEX: 3, PRISM: 3}
NSIDE = {TRI: 3, QUAD: 4, TET: 4, HEX: 6, PRISM: 5}


class Params(C.Structure):
    _fields_ = [("disc", C.c_int32), ("elem", C.c_int32), ("conv_upwind", C.c_int32), ("stab", C.c_int32),
                ("stab_upwind", C.c_int32), ("diff_len", C.c_int32), ("stokes", C.c_int32),
                ("laplace", C.c_int32), ("peclet_blend", C.c_int32), ("pac", C.c_int32),
                ("defect_upwind", C.c_int32), ("time_dependent", C.c_int32), ("has_source", C.c_int32),
                ("pad0", C.c_int32), ("exact_jac", C.c_double), ("grad_div", C.c_double),
                ("kin_visc", C.c_double), ("density", C.c_double), ("source", C.c_double * 3),
                ("dt", C.c_double),
                # per-ip data imports (FV1): global arrays [n_elem][nip] / [n_elem][nsh] (x dim for the sources), NULL = constants
                ("ip_visc", C.c_void_p), ("ip_rho_scvf", C.c_void_p), ("ip_rho_scv", C.c_void_p), ("ip_src_scvf", C.c_void_p),
                ("ip_src_scv", C.c_void_p), ("elem_index", C.c_int64)]


class FV1Geom(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nsh", C.c_int32), ("nip", C.c_int32), ("pad", C.c_int32),
                ("from_", C.c_int32 * 12), ("to", C.c_int32 * 12),
                ("normal", C.c_double * 3 * 12), ("xip", C.c_double * 3 * 12), ("lip", C.c_double * 3 * 12),
                ("shape", C.c_double * 8 * 12), ("ggrad", C.c_double * 3 * 8 * 12),
                ("c0c2sq", C.c_double * 12), ("vol", C.c_double * 8)]


class CRGeom(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nsh", C.c_int32), ("nip", C.c_int32), ("nco", C.c_int32),
                ("from_", C.c_int32 * 12), ("to", C.c_int32 * 12),
                ("normal", C.c_double * 3 * 12), ("xip", C.c_double * 3 * 12), ("lip", C.c_double * 3 * 12),
                ("shape", C.c_double * 6 * 12), ("ggrad", C.c_double * 3 * 6 * 12),
                ("scv_normal", C.c_double * 3 * 6), ("scv_xip", C.c_double * 3 * 6), ("vol", C.c_double * 6)]


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "ns_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src),
                                                   os.path.getmtime(os.path.join(_HERE, "ns_oracle.h")))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libns_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp, ip32, ip64 = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.ora_last_error.restype = C.c_char_p
        L.ora_fv1_geometry.argtypes = [C.c_int, dp, C.POINTER(FV1Geom)]
        L.ora_cr_geometry.argtypes = [C.c_int, dp, C.POINTER(CRGeom)]
        L.ora_fv1_upwind.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp]
        L.ora_side_ray_intersection.argtypes = [C.c_int, dp, dp, dp, C.c_int, C.POINTER(C.c_int), dp, dp]
        L.ora_fv1_elem.argtypes = [C.POINTER(Params), dp, dp, dp, dp, C.c_int, dp, dp]
        L.ora_fvcr_elem.argtypes = [C.POINTER(Params), dp, dp, C.c_int, dp, dp]
        L.ora_fv1_stab.argtypes = [C.POINTER(Params), dp, dp, dp, dp, dp, dp, dp]
        L.ora_fv1_csr.restype = C.c_int64
        L.ora_fv1_csr.argtypes = [C.c_int, C.c_int64, C.c_int64, ip32, ip64, ip32]
        L.ora_fvcr_csr.restype = C.c_int64
        L.ora_fvcr_csr.argtypes = [C.c_int, C.c_int64, C.c_int64, ip32, ip64, ip32]
        L.ora_assemble.argtypes = [C.POINTER(Params), C.c_int64, C.c_int64, ip32, dp, ip32, dp, dp, dp,
                                   ip64, ip32, C.c_int, C.c_double, C.c_double, dp, dp, C.c_int]
        L.ora_side_corners.argtypes = [C.c_int, C.c_int]
        L.ora_side_corner.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ora_fv1_bf_geometry.argtypes = [C.c_int, dp, C.c_int, C.c_int, C.POINTER(C.c_int), dp, dp, dp, dp]
        L.ora_fv1_boundary.argtypes = [C.POINTER(Params), C.c_int, C.c_int64, ip32, ip32, dp, ip32, dp, dp, ip64, ip32,
                                       C.c_int, C.c_double, dp, dp]
        L.ora_fv1_smagorinsky.argtypes = [C.c_int, C.c_int64, C.c_int64, ip32, dp, dp, C.c_double, C.c_double, C.c_int64, ip32, ip32,
                                          C.POINTER(C.c_uint8), dp, dp]
        L.ora_fv1_vorticity.argtypes = [C.c_int, C.c_int64, C.c_int64, ip32, dp, dp, dp]
        L.ora_fvcr_diagnostics.argtypes = [C.c_int, C.c_int64, ip32, dp, ip32, dp, C.c_double, dp]
        L.ora_fvcr_constraint_defect.argtypes = [C.c_int, C.c_int64, C.c_int64, ip32, dp, ip32, dp, C.c_double, C.c_int, C.c_int,
                                                 C.POINTER(C.c_uint8), dp]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _chk(rc):
    if rc < 0:
        raise OracleError(lib().ora_last_error().decode())


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _i32(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _i64(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int64))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def make_params(disc="fv1", elem="hex", upwind="full", stab="fields", stab_upwind="same", diff_len="raw",
                stokes=False, laplace=False, peclet_blend=False, pac=False, defect_upwind=True,
                exact_jac=0.0, grad_div=0.0, kin_visc=1e-2, density=1.0, source=None, dt=0.0,
                time_dependent=False):
    p = Params()
    p.disc = 0 if disc == "fv1" else 1
    p.elem = ELEM[elem] if isinstance(elem, str) else elem
    p.conv_upwind = UPWIND[upwind]
    p.stab = STAB[stab] if stab is not None else -1
    p.stab_upwind = UPWIND[upwind] if stab_upwind == "same" else UPWIND[stab_upwind]
    p.diff_len = DIFF[diff_len]
    p.stokes, p.laplace, p.peclet_blend, p.pac = int(stokes), int(laplace), int(peclet_blend), int(pac)
    p.defect_upwind = int(defect_upwind)
    p.time_dependent = int(time_dependent)
    p.exact_jac, p.grad_div, p.kin_visc, p.density, p.dt = float(exact_jac), grad_div, kin_visc, density, dt
    if source is not None:
        p.has_source = 1
        for d, v in enumerate(source):
            p.source[d] = v
    return p


def fv1_geometry(elem, coords):
    g = FV1Geom()
    coords = _f64(coords)
    _chk(lib().ora_fv1_geometry(elem, _dp(coords), C.byref(g)))
    nip, nsh, dim = g.nip, g.nsh, g.dim
    return dict(dim=dim, nsh=nsh, nip=nip, frm=np.array(g.from_[:nip]), to=np.array(g.to[:nip]),
                normal=np.array(g.normal)[:nip, :dim], xip=np.array(g.xip)[:nip, :dim],
                lip=np.array(g.lip)[:nip, :dim], shape=np.array(g.shape)[:nip, :nsh],
                ggrad=np.array(g.ggrad)[:nip, :nsh, :dim], c0c2sq=np.array(g.c0c2sq)[:nip],
                vol=np.array(g.vol)[:nsh])


def cr_geometry(elem, coords):
    g = CRGeom()
    coords = _f64(coords)
    _chk(lib().ora_cr_geometry(elem, _dp(coords), C.byref(g)))
    nip, nsh, dim = g.nip, g.nsh, g.dim
    return dict(dim=dim, nsh=nsh, nip=nip, frm=np.array(g.from_[:nip]), to=np.array(g.to[:nip]),
                normal=np.array(g.normal)[:nip, :dim], xip=np.array(g.xip)[:nip, :dim],
                lip=np.array(g.lip)[:nip, :dim], shape=np.array(g.shape)[:nip, :nsh],
                ggrad=np.array(g.ggrad)[:nip, :nsh, :dim], scv_normal=np.array(g.scv_normal)[:nsh, :dim],
                scv_xip=np.array(g.scv_xip)[:nsh, :dim], vol=np.array(g.vol)[:nsh])


def fv1_upwind(elem, upwind, coords, ipvel):
    nip, nsh = NIP[elem], NSH[elem]
    coords, ipvel = _f64(coords), _f64(ipvel)
    sh, ipm, ln = np.zeros((nip, nsh)), np.zeros((nip, nip)), np.zeros(nip)
    _chk(lib().ora_fv1_upwind(elem, UPWIND[upwind], _dp(coords), _dp(ipvel), _dp(sh), _dp(ipm), _dp(ln)))
    return sh, ipm, ln


def side_ray_intersection(elem, coords, frm, direction, positive=False):
    dim = DIM[elem]
    coords, frm, direction = _f64(coords), _f64(frm), _f64(direction)
    side = C.c_int(-1)
    g, l = np.zeros(dim), np.zeros(dim)
    ok = lib().ora_side_ray_intersection(elem, _dp(coords), _dp(frm), _dp(direction), int(positive),
                                         C.byref(side), _dp(g), _dp(l))
    return bool(ok == 1), side.value, g, l


def fv1_elem(p, coords, u, what, sol0=None, sol1=None):
    """u: [dim+1][nsh] (LocalVector order). returns (Jloc [L,L] with index fct*nsh+sh, dloc [L])."""
    nsh, dim = NSH[p.elem], DIM[p.elem]
    L = (dim + 1) * nsh
    coords, u, sol0, sol1 = _f64(coords), _f64(u), _f64(sol0), _f64(sol1)
    J, d = np.zeros((L, L)), np.zeros(L)
    _chk(lib().ora_fv1_elem(C.byref(p), _dp(coords), _dp(u), _dp(sol0), _dp(sol1), what, _dp(J), _dp(d)))
    return J, d


def fvcr_elem(p, coords, u, what):
    ns, dim = NSIDE[p.elem], DIM[p.elem]
    L = dim * ns + 1
    coords, u = _f64(coords), _f64(u)
    J, d = np.zeros((L, L)), np.zeros(L)
    _chk(lib().ora_fvcr_elem(C.byref(p), _dp(coords), _dp(u), what, _dp(J), _dp(d)))
    return J, d


def fv1_stab(p, coords, u, sol0=None, sol1=None):
    nsh, dim, nip = NSH[p.elem], DIM[p.elem], NIP[p.elem]
    coords, u, sol0, sol1 = _f64(coords), _f64(u), _f64(sol0), _f64(sol1)
    sv, shv, shp = np.zeros((nip, dim)), np.zeros((nip, dim, dim, nsh)), np.zeros((nip, dim, nsh))
    _chk(lib().ora_fv1_stab(C.byref(p), _dp(coords), _dp(u), _dp(sol0), _dp(sol1), _dp(sv), _dp(shv), _dp(shp)))
    return sv, shv, shp


def fv1_csr(elem, conn, n_node):
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    n_elem = conn.shape[0]
    nf = DIM[elem] + 1
    nnz = lib().ora_fv1_csr(elem, n_elem, n_node, _i32(conn), None, None)
    rowptr = np.zeros(n_node * nf + 1, dtype=np.int64)
    colind = np.zeros(nnz, dtype=np.int32)
    lib().ora_fv1_csr(elem, n_elem, n_node, _i32(conn), _i64(rowptr), _i32(colind))
    return rowptr, colind


def fvcr_csr(elem, elem_sides, n_side):
    es = np.ascontiguousarray(elem_sides, dtype=np.int32)
    n_elem = es.shape[0]
    dim = DIM[elem]
    nnz = lib().ora_fvcr_csr(elem, n_elem, n_side, _i32(es), None, None)
    rowptr = np.zeros(n_side * dim + n_elem + 1, dtype=np.int64)
    colind = np.zeros(nnz, dtype=np.int32)
    lib().ora_fvcr_csr(elem, n_elem, n_side, _i32(es), _i64(rowptr), _i32(colind))
    return rowptr, colind


IP_KINDS = ("visc", "rho_scvf", "rho_scv", "src_scvf", "src_scv")


def assemble(p, conn, coords, u, rowptr, colind, what, sol0=None, sol1=None, elem_sides=None, n_side=0,
             scale_a=1.0, scale_m=1.0, nthreads=1, values=None, defect=None, ip_data=None):
    """Serial (or coloured-threaded) element loop + scatter. returns (values, defect).
    ip_data: {kind: array} per-ip data imports (FV1), kinds = IP_KINDS, arrays [n_elem][nip | nsh]([dim])."""
    keep = {}
    for kind in IP_KINDS:
        a = None if ip_data is None else ip_data.get(kind)
        if a is not None:
            a = keep[kind] = np.ascontiguousarray(a, dtype=np.float64)
        setattr(p, "ip_" + kind, a.ctypes.data if a is not None else None)
    try:
        return _assemble(p, conn, coords, u, rowptr, colind, what, sol0, sol1, elem_sides, n_side, scale_a, scale_m, nthreads, values, defect)
    finally:
        for kind in IP_KINDS:
            setattr(p, "ip_" + kind, None)


def _assemble(p, conn, coords, u, rowptr, colind, what, sol0=None, sol1=None, elem_sides=None, n_side=0,
              scale_a=1.0, scale_m=1.0, nthreads=1, values=None, defect=None):
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    coords = _f64(coords)
    u, sol0, sol1 = _f64(u), _f64(sol0), _f64(sol1)
    es = None if elem_sides is None else np.ascontiguousarray(elem_sides, dtype=np.int32)
    n_ent = coords.shape[0] if p.disc == 0 else n_side
    if values is None:
        values = np.zeros(colind.shape[0])
    if defect is None:
        defect = np.zeros(rowptr.shape[0] - 1)
    _chk(lib().ora_assemble(C.byref(p), conn.shape[0], n_ent, _i32(conn), _dp(coords), _i32(es), _dp(u),
                            _dp(sol0), _dp(sol1), _i64(rowptr), _i32(colind), what, scale_a, scale_m,
                            _dp(values), _dp(defect), nthreads))
    return values, defect


# ---- boundary faces and the boundary discs on them ----
BND_OUTFLOW, BND_INFLOW = 0, 1


def side_corners(elem, side):
    """element corners of a side in reference order"""
    n = lib().ora_side_corners(elem, side)
    return [lib().ora_side_corner(elem, side, j) for j in range(n)]


def fv1_bf_geometry(elem, coords, side, j):
    """boundary face j of a side: (node_id, normal[dim], xip[dim], shape[nsh], ggrad[nsh][dim])"""
    dim, nsh = DIM[elem], NSH[elem]
    coords = _f64(coords)
    nid = C.c_int(0)
    n, x, N, G = np.zeros(dim), np.zeros(dim), np.zeros(nsh), np.zeros((nsh, dim))
    _chk(lib().ora_fv1_bf_geometry(elem, _dp(coords), side, j, C.byref(nid), _dp(n), _dp(x), _dp(N), _dp(G)))
    return nid.value, n, x, N, G


def fv1_boundary(p, kind, belem, bside, conn, coords, u, rowptr, colind, what, data=None, scale_a=1.0, values=None, defect=None):
    """adds the boundary-disc contributions of the (element, side) pairs to values / defect. returns (values, defect)"""
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    belem, bside = np.ascontiguousarray(belem, dtype=np.int32), np.ascontiguousarray(bside, dtype=np.int32)
    coords, u, data = _f64(coords), _f64(u), _f64(data)
    if values is None:
        values = np.zeros(colind.shape[0])
    if defect is None:
        defect = np.zeros(rowptr.shape[0] - 1)
    _chk(lib().ora_fv1_boundary(C.byref(p), kind, belem.shape[0], _i32(belem), _i32(bside), _dp(data), _i32(conn), _dp(coords),
                                _dp(u), _i64(rowptr), _i32(colind), what, scale_a, _dp(values), _dp(defect)))
    return values, defect


# ---- turbulent viscosity and diagnostics ----
def fv1_smagorinsky(elem, conn, coords, u, c=0.05, kin_visc=0.0, belem=None, bside=None, zero_nodes=None):
    """FV1SmagorinskyTurbViscData: returns (nu_t [n_node], ip_visc [n_elem][nip] = interpolated nu_t + kin_visc)"""
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    coords, u = _f64(coords), _f64(u)
    n_node = coords.shape[0]
    nb = 0 if belem is None else len(belem)
    be = None if belem is None else np.ascontiguousarray(belem, dtype=np.int32)
    bs = None if bside is None else np.ascontiguousarray(bside, dtype=np.int32)
    z = None
    if zero_nodes is not None:
        z = np.zeros(n_node, dtype=np.uint8)
        z[np.asarray(zero_nodes, dtype=np.int64)] = 1
    nut, ipv = np.zeros(n_node), np.zeros((conn.shape[0], NIP[elem]))
    _chk(lib().ora_fv1_smagorinsky(elem, conn.shape[0], n_node, _i32(conn), _dp(coords), _dp(u), c, kin_visc, nb, _i32(be), _i32(bs),
                                   None if z is None else z.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(nut), _dp(ipv)))
    return nut, ipv


def fv1_vorticity(elem, conn, coords, u):
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    coords, u = _f64(coords), _f64(u)
    vort = np.zeros(coords.shape[0])
    _chk(lib().ora_fv1_vorticity(elem, conn.shape[0], coords.shape[0], _i32(conn), _dp(coords), _dp(u), _dp(vort)))
    return vort


def fvcr_diagnostics(elem, conn, coords, elem_sides, u, dt=1.0):
    """(kinetic energy, max CFL number) of a Crouzeix-Raviart velocity field"""
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    es = np.ascontiguousarray(elem_sides, dtype=np.int32)
    coords, u = _f64(coords), _f64(u)
    out = np.zeros(2)
    _chk(lib().ora_fvcr_diagnostics(elem, conn.shape[0], _i32(conn), _dp(coords), _i32(es), _dp(u), dt, _dp(out)))
    return out[0], out[1]


def fvcr_constraint_defect(elem, conn, coords, elem_sides, n_side, u, s_a=1.0, lin_upwind=True, lin_pressure=True, zero_grad_sides=None,
                           defect=None):
    """DiscConstraintFVCR (default configuration): adds the linear-upwind / linear-pressure defect correction. returns defect"""
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    es = np.ascontiguousarray(elem_sides, dtype=np.int32)
    coords, u = _f64(coords), _f64(u)
    if defect is None:
        defect = np.zeros(u.shape[0])
    z = None
    if zero_grad_sides is not None:
        z = np.zeros(n_side, dtype=np.uint8)
        z[np.asarray(zero_grad_sides, dtype=np.int64)] = 1
    _chk(lib().ora_fvcr_constraint_defect(elem, conn.shape[0], n_side, _i32(conn), _dp(coords), _i32(es), _dp(u), s_a, int(lin_upwind),
                                          int(lin_pressure), None if z is None else z.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(defect)))
    return defect
