"""Per-ip data imports ("const or per-ip arrays"; nsb_set_ip_data): spatially varying kinematic viscosity, density and
source, evaluated at the integration points the reference evaluates its UserData at (fv1/navier_stokes_fv1.cpp:184-197)."""
import numpy as np
import pytest

from plugin_navierstokes_b200 import meshgen


def nu_fn(*x):
    return 1e-2 * (1.0 + 0.5 * np.sin(3.0 * x[0]) * np.cos(2.0 * x[-1]))


def rho_fn(*x):
    return 1.0 + 0.3 * x[0] + 0.1 * x[-1] ** 2


def src_fn(*x):
    return [0.2 * x[0], -0.1 + 0.3 * x[-1], 0.05][:len(x)]


def ip_arrays(elem, conn, coords):
    xf = meshgen.fv1_scvf_ips(elem, conn, coords)
    xv = meshgen.fv1_scv_ips(elem, conn, coords)
    ev = lambda f, x: np.array([f(*p) for p in x.reshape(-1, x.shape[-1])]).reshape(x.shape[:2] + (-1,)).squeeze(-1) if np.ndim(f(*x[0, 0])) == 0 \
        else np.array([f(*p) for p in x.reshape(-1, x.shape[-1])]).reshape(x.shape[:2] + (-1,))
    return dict(visc=ev(nu_fn, xf), rho_scvf=ev(rho_fn, xf), rho_scv=ev(rho_fn, xv), src_scvf=ev(src_fn, xf), src_scv=ev(src_fn, xv))


@pytest.mark.parametrize("elem", ["tri", "quad", "tet", "hex"])
def test_host_ip_positions_are_the_oracle_geometry(ora, elem):
    """the host evaluates UserData at fv1_scvf_ips / fv1_scv_ips: they are the oracle's (= FV1Geometry's) global ips"""
    coords, conn = meshgen.make_mesh(elem, 3, jitter=0.2, seed=3)
    xf = meshgen.fv1_scvf_ips(elem, conn, coords)
    xv = meshgen.fv1_scv_ips(elem, conn, coords)
    for e in range(min(conn.shape[0], 20)):
        g = ora.fv1_geometry(ora.ELEM[elem], coords[conn[e]])
        assert np.abs(xf[e] - g["xip"]).max() < 1e-14
        assert np.array_equal(xv[e], coords[conn[e]])


def test_oracle_constant_arrays_equal_constants(ora):
    coords, conn = meshgen.hex_grid(3, 3, 2, jitter=0.2, seed=1)
    u = meshgen.state_vortex3d(coords, seed=2, noise=0.05).reshape(-1)
    p = ora.make_params(elem="hex", upwind="lps", stab="flow", kin_visc=0.02, density=1.3, source=[0.1, 0.2, 0.3])
    rp, ci = ora.fv1_csr(ora.HEX, conn, coords.shape[0])
    W = ora.JAC_A | ora.DEF_A | ora.JAC_M | ora.DEF_M | ora.RHS
    v0, d0 = ora.assemble(p, conn, coords, u, rp, ci, W)
    ne = conn.shape[0]
    ipd = dict(visc=np.full((ne, 12), 0.02), rho_scvf=np.full((ne, 12), 1.3), rho_scv=np.full((ne, 8), 1.3),
               src_scvf=np.tile([0.1, 0.2, 0.3], (ne, 12, 1)), src_scv=np.tile([0.1, 0.2, 0.3], (ne, 8, 1)))
    v1, d1 = ora.assemble(p, conn, coords, u, rp, ci, W, ip_data=ipd)
    assert np.array_equal(v0, v1) and np.array_equal(d0, d1)
    ipd["visc"] = ipd["visc"] * (1.0 + np.random.default_rng(0).uniform(0, 1, (ne, 12)))
    v2, _ = ora.assemble(p, conn, coords, u, rp, ci, W, ip_data=ipd)
    assert np.abs(v2 - v0).max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
@pytest.mark.parametrize("elem,n,upwind,stab", [("hex", 5, "lps", "fields"), ("hex", 4, "full", "flow"), ("tet", 4, "skewed", "fields"),
                                                ("quad", 10, "lps", "flow"), ("tri", 10, "full", "fields"), ("quad", 8, "no", "none")])
def test_variable_viscosity_density_source_against_the_oracle(ora, elem, n, upwind, stab, mode):
    import plugin_navierstokes_b200 as pkg
    from plugin_navierstokes_b200 import capi
    from tests import parity
    coords, conn, u = parity.make_case(elem, n, seed=7)
    dim = coords.shape[1]
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    parity.configure(disc, upwind=upwind, stab=stab)
    disc.set_kinematic_viscosity(nu_fn)                     # UserData: evaluated at the SCVF ips by the host mirror
    disc.set_density(rho_fn)
    disc.set_source(src_fn)
    disc.set_grid(elem, conn, coords)
    rowptr, colind = ora.fv1_csr(ora.ELEM[elem], conn, coords.shape[0])
    p = ora.make_params(elem=elem, upwind=upwind, stab=stab)
    what = capi.JAC_A | capi.DEF_A | capi.JAC_M | capi.DEF_M | capi.RHS
    ov, od = ora.assemble(p, conn, coords, u.reshape(-1), rowptr, colind, what, scale_a=0.7, scale_m=1.3, ip_data=ip_arrays(elem, conn, coords))
    gv, gd = disc.assemble(what, u.reshape(-1), scale_a=0.7, scale_m=1.3,
                           scatter_mode={"gather": capi.SCATTER_GATHER, "colored": capi.SCATTER_COLORED, "atomic": capi.SCATTER_ATOMIC}[mode])
    eg, ee = parity.entry_errors(gv, ov, rowptr)
    assert eg < parity.TOL and ee < parity.TOL, (eg, ee)
    eg, ee = parity.entry_errors(gd, od)
    assert eg < parity.TOL and ee < parity.TOL, (eg, ee)
    # back to constants: the arrays are cleared and the owner-computes path serves gather again
    disc.set_kinematic_viscosity(1e-2); disc.set_density(1.0); disc.set_source([0.1, 0.2, 0.3][:dim])
    p2 = ora.make_params(elem=elem, upwind=upwind, stab=stab, source=[0.1, 0.2, 0.3][:dim])
    ov, od = ora.assemble(p2, conn, coords, u.reshape(-1), rowptr, colind, what)
    gv, gd = disc.assemble(what, u.reshape(-1))
    eg, ee = parity.entry_errors(gv, ov, rowptr)
    assert eg < parity.TOL and ee < parity.TOL, (eg, ee)
    disc.close()


@pytest.mark.gpu
def test_ip_data_with_positive_upwind_is_rejected(ora):
    import plugin_navierstokes_b200 as pkg
    from plugin_navierstokes_b200 import capi
    coords, conn = meshgen.quad_grid(4, 4)
    disc = pkg.NavierStokesFV1("u,v,p", "Inner")
    disc.set_kinematic_viscosity(nu_fn)
    disc.set_upwind("pos"); disc.set_stabilization("fields")
    disc.set_grid("quad", conn, coords)
    with pytest.raises(pkg.UGError, match="per-ip data"):
        disc.assemble(capi.JAC_A, np.zeros(coords.shape[0] * 3))
    disc.close()
