"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Tolerance: 1e-12 relative per defect entry / Jacobian nonzero (north_star); identical CSR sparsity."""
import numpy as np
import pytest

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi
from tests import parity
from tests.parity import TOL

pytestmark = pytest.mark.gpu

SIZES = {"tri": 9, "quad": 8, "tet": 4, "hex": 4}
FCTS = {2: "u,v,p", 3: "u,v,w,p"}
MODES = {"gather": capi.SCATTER_GATHER, "colored": capi.SCATTER_COLORED, "atomic": capi.SCATTER_ATOMIC}


def _run_case(ora, elem, mode, upwind="full", stab="fields", diff="raw", what=None, time_dep=False, seed=0, cond_aware=False, **flags):
    """cond_aware: the per-entry bound is max(TOL, 4 x the change of the ORACLE's own result under 1-ulp perturbations of the inputs)
    -- for a case whose entries are ill-conditioned sums no implementation can agree with another one better than that"""
    coords, conn, u = parity.make_case(elem, SIZES[elem], seed=seed)
    dim = coords.shape[1]
    E = ora.ELEM[elem]
    disc = pkg.NavierStokesFV1(FCTS[dim], "Inner")
    parity.configure(disc, upwind=upwind, stab=stab, diff=diff, **flags)
    disc.set_grid(elem, conn, coords)
    disc.prep_elem_loop()
    rp, ci = disc.csr()
    rowptr, colind = ora.fv1_csr(E, conn, coords.shape[0])
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colind)
    what = what if what is not None else (capi.JAC_A | capi.DEF_A)
    ts, s0, s1, dt = None, None, None, 0.0
    if time_dep:
        dt = 0.05
        s0 = u * 1.01 + 0.003
        s1 = u * 0.97 - 0.002
        ts = (s0, s1, dt)
    p = ora.make_params(elem=elem, upwind=upwind, stab=stab, diff_len=diff, kin_visc=flags.get("visc", 1e-2),
                        density=flags.get("density", 1.0), stokes=flags.get("stokes", False),
                        laplace=flags.get("laplace", False), peclet_blend=flags.get("peclet", False),
                        pac=flags.get("pac", False), exact_jac=flags.get("exact", 0.0),
                        source=flags.get("source"), dt=dt, time_dependent=time_dep,
                        stab_upwind=flags.get("stab_upwind") or "same")
    sa, sm = (0.7, 1.3) if (what & (capi.JAC_M | capi.DEF_M)) else (1.0, 1.0)
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, what, sol0=s0, sol1=s1, scale_a=sa, scale_m=sm)
    gv, gd = disc.assemble(what, u, time_series=ts, scale_a=sa, scale_m=sm, scatter_mode=MODES[mode])
    tol_e = TOL
    if cond_aware:
        for s in range(3):
            rng = np.random.default_rng(100 + s)
            c2 = coords * (1 + 2.2e-16 * rng.integers(-1, 2, coords.shape))
            u2 = u * (1 + 2.2e-16 * rng.integers(-1, 2, u.shape))
            pv, _ = ora.assemble(p, conn, c2, u2, rowptr, colind, what, sol0=s0, sol1=s1, scale_a=sa, scale_m=sm)
            tol_e = max(tol_e, 4 * parity.entry_errors(pv, ov, rowptr)[1])
    if what & (capi.JAC_A | capi.JAC_M):
        eg, ee = parity.entry_errors(gv, ov, rowptr)
        assert eg < TOL and ee < tol_e, ("jacobian", elem, mode, upwind, stab, eg, ee)
    if what & (capi.DEF_A | capi.DEF_M | capi.RHS):
        eg, ee = parity.entry_errors(gd, od)
        assert eg < TOL and ee < TOL, ("defect", elem, mode, upwind, stab, eg, ee)
    disc.close()


@pytest.mark.parametrize("elem", ["hex", "tet", "quad", "tri"])
@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
@pytest.mark.parametrize("upwind", ["no", "full", "skewed", "lps"])
@pytest.mark.parametrize("stab", ["fields", "flow", "none"])
def test_stationary_jac_def(ora, elem, mode, upwind, stab):
    _run_case(ora, elem, mode, upwind=upwind, stab=stab)


@pytest.mark.parametrize("elem", ["hex", "tet", "quad", "tri"])
@pytest.mark.parametrize("mode", ["gather", "colored"])
@pytest.mark.parametrize("flags", [
    dict(exact=1.0), dict(peclet=True), dict(peclet=True, exact=0.5), dict(laplace=True), dict(stokes=True, upwind=None),
    dict(pac=True), dict(pac=True, exact=1.0, peclet=True), dict(pac=True, stab="flow", exact=1.0),
    dict(diff="fivepoint"), dict(diff="cor"), dict(stab="flow", diff="cor", upwind="lps"),
    dict(density=1.3, visc=3e-3, source=[0.3, -0.2, 0.1]), dict(upwind="lps", stab_upwind="full"),
], ids=lambda f: "-".join("%s=%s" % kv for kv in f.items()))
def test_flags(ora, elem, mode, flags):
    flags = dict(flags)
    if "source" in flags and elem in ("tri", "quad"):
        flags["source"] = flags["source"][:2]
    _run_case(ora, elem, mode, **flags)


@pytest.mark.parametrize("elem", ["hex", "tet", "quad", "tri"])
@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
@pytest.mark.parametrize("stab", ["fields", "flow"])
def test_instationary_parts(ora, elem, mode, stab):
    """two time points: 1/dt in the ip system, u_old/dt in its rhs, lumped mass + rhs parts, scales"""
    src = [0.3, -0.2, 0.1][: (2 if elem in ("tri", "quad") else 3)]
    what = capi.JAC_A | capi.DEF_A | capi.JAC_M | capi.DEF_M | capi.RHS
    _run_case(ora, elem, mode, upwind="lps", stab=stab, what=what, time_dep=True, source=src, density=1.2)
    _run_case(ora, elem, mode, upwind="full", stab=stab, what=capi.DEF_M | capi.JAC_M)


@pytest.mark.parametrize("elem", ["hex", "tet", "quad", "tri"])
@pytest.mark.parametrize("mode", ["colored", "atomic", "gather"])
@pytest.mark.parametrize("flags", [
    dict(stab="fields"), dict(stab="flow"), dict(stab="none"), dict(stab="flow", exact=1.0, peclet=True),
    dict(stab="fields", pac=True, exact=1.0), dict(stab="flow", pac=True, exact=0.5, peclet=True),
    dict(stab="flow", diff="cor"), dict(stab="fields", upwind="full", stab_upwind="positive"),
    dict(stab="flow", upwind="positive", stab_upwind="lps", exact=1.0), dict(stab="fields", laplace=True, density=1.4),
], ids=lambda f: "-".join("%s=%s" % kv for kv in f.items()))
def test_positive_upwind_dense_branch(ora, elem, mode, flags):
    """PositiveUpwind has ip-shapes: nIp x nIp system per element (stabilization.cpp:244-403, :590-771).
    (the gather mode routes these configurations to the coloured element kernel)"""
    flags = dict(flags)
    flags.setdefault("upwind", "positive")
    _run_case(ora, elem, mode, **flags)


@pytest.mark.parametrize("elem", ["hex", "tet", "quad", "tri"])
@pytest.mark.parametrize("stab", ["fields", "flow"])
def test_positive_upwind_instationary(ora, elem, stab):
    """config 5: FLOW + PositiveUpwind with two time points, mass parts and scales"""
    src = [0.3, -0.2, 0.1][: (2 if elem in ("tri", "quad") else 3)]
    what = capi.JAC_A | capi.DEF_A | capi.JAC_M | capi.DEF_M | capi.RHS
    _run_case(ora, elem, "colored", upwind="positive", stab=stab, what=what, time_dep=True, source=src, density=1.2)
    _run_case(ora, elem, "atomic", upwind="positive", stab=stab, what=capi.JAC_A | capi.DEF_A, time_dep=True, pac=True, exact=1.0)


@pytest.mark.parametrize("elem", ["hex", "tet", "quad", "tri"])
def test_local_contributions_match_oracle(ora, elem):
    """compat mode of the IElemDisc slots: per-element LocalMatrix / LocalVector blocks"""
    coords, conn, u = parity.make_case(elem, 3, seed=5)
    dim = coords.shape[1]
    disc = pkg.NavierStokesFV1(FCTS[dim], "Inner")
    parity.configure(disc, upwind="lps", stab="flow", exact=1.0)
    disc.set_grid(elem, conn, coords)
    J, d = disc.local_contributions(capi.JAC_A | capi.DEF_A, u)
    p = ora.make_params(elem=elem, upwind="lps", stab="flow", exact_jac=1.0)
    for e in range(conn.shape[0]):
        Jo, do = ora.fv1_elem(p, coords[conn[e]], u[conn[e]].T, ora.JAC_A | ora.DEF_A)
        assert np.abs(J[e] - Jo).max() <= TOL * np.abs(Jo).max(), e
        assert np.abs(d[e] - do).max() <= TOL * np.abs(do).max(), e


def test_prep_elem_loop_errors_mirror_the_reference():
    coords, conn, u = parity.make_case("quad", 3)
    d = pkg.NavierStokesFV1("u,v,p", "Inner")
    d.set_grid("quad", conn, coords)
    d.set_kinematic_viscosity(0.01)
    with pytest.raises(pkg.UGError, match="Stabilization has not been set"):
        d.prep_elem_loop()
    d.set_stabilization("fields")
    with pytest.raises(pkg.UGError, match="Upwinding for convective Term"):
        d.prep_elem_loop()
    d.set_stokes(True)
    d.prep_elem_loop()
    d2 = pkg.NavierStokesFV1("u,v,p", "Inner")
    d2.set_grid("quad", conn, coords)
    d2.set_upwind("full")
    d2.set_stabilization("fields")
    with pytest.raises(pkg.UGError, match="Kinematic Viscosity has not been set"):
        d2.prep_elem_loop()
    with pytest.raises(pkg.UGError, match="Wrong number of functions"):
        pkg.NavierStokesFV1("u,v,w,p", "Inner").set_grid("quad", conn, coords)


def test_deterministic_modes_are_bitwise_reproducible():
    coords, conn, u = parity.make_case("hex", 6, seed=9)
    disc = pkg.NavierStokesFV1("u,v,w,p", "Inner")
    parity.configure(disc, upwind="lps", stab="flow")
    disc.set_grid("hex", conn, coords)
    for mode in (capi.SCATTER_GATHER, capi.SCATTER_COLORED):
        a = disc.assemble(capi.JAC_A | capi.DEF_A, u, scatter_mode=mode)
        b = disc.assemble(capi.JAC_A | capi.DEF_A, u, scatter_mode=mode)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_device_pointer_path_and_beta():
    import torch
    coords, conn, u = parity.make_case("hex", 5, seed=3)
    disc = pkg.NavierStokesFV1("u,v,w,p", "Inner")
    parity.configure(disc, upwind="full", stab="fields")
    disc.set_grid("hex", conn, coords)
    hv, hd = disc.assemble(capi.JAC_A | capi.DEF_A, u)
    ud = torch.from_numpy(u.reshape(-1)).cuda()
    disc.use_stream(torch.cuda.current_stream().cuda_stream)
    for mode in (capi.SCATTER_GATHER, capi.SCATTER_COLORED, capi.SCATTER_ATOMIC):
        dv, dd = disc.assemble(capi.JAC_A | capi.DEF_A, ud, scatter_mode=mode)
        disc.check_errors()
        assert np.allclose(dv.cpu().numpy(), hv, rtol=1e-13, atol=1e-13 * np.abs(hv).max())
        # accumulate a second time with beta = 1 -> exactly twice for the deterministic modes
        dv2, dd2 = disc.assemble(capi.JAC_A | capi.DEF_A, ud, values=dv.clone(), defect=dd.clone(), beta=1.0, scatter_mode=mode)
        disc.check_errors()
        assert np.allclose(dv2.cpu().numpy(), 2 * hv, rtol=1e-12, atol=1e-12 * np.abs(hv).max())
        assert np.allclose(dd2.cpu().numpy(), 2 * hd, rtol=1e-12, atol=1e-12 * np.abs(hd).max())


@pytest.mark.parametrize("elem", ["hex", "tri"])
@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
def test_unreferenced_nodes_and_ragged_valence(ora, elem, mode):
    """Nodes no element touches (empty CSR rows, zero defect) in the middle and at the end of the numbering, and a mesh
    with a removed element block (ragged node valence): the owner-computes scheduler must neither skip nor stall."""
    coords, conn, u = parity.make_case(elem, SIZES[elem], seed=5)
    dim = coords.shape[1]
    nf = dim + 1
    keep = np.ones(conn.shape[0], bool)
    keep[3:9] = False                                           # punch a hole
    conn = conn[keep]
    # insert unreferenced nodes: one in the middle of the numbering, two at the end
    mid = coords.shape[0] // 2
    coords2 = np.concatenate([coords[:mid], [[9.0] * dim], coords[mid:], [[8.0] * dim, [7.0] * dim]])
    conn2 = np.where(conn >= mid, conn + 1, conn).astype(np.int32)
    u = u.reshape(-1, nf)
    u2 = np.concatenate([u[:mid], np.full((1, nf), 0.5), u[mid:], np.full((2, nf), -0.25)])
    used = np.zeros(coords2.shape[0], bool)
    used[conn2.ravel()] = True
    assert (~used).sum() >= 3
    disc = pkg.NavierStokesFV1(FCTS[dim], "Inner")
    parity.configure(disc, upwind="lps", stab="fields", diff="raw")
    disc.set_grid(elem, conn2, coords2)
    rp, ci = disc.csr()
    rowptr, colind = ora.fv1_csr(ora.ELEM[elem], conn2, coords2.shape[0])
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colind)
    p = ora.make_params(elem=elem, upwind="lps", stab="fields", diff_len="raw", kin_visc=1e-2, density=1.0)
    what = capi.JAC_A | capi.DEF_A | capi.JAC_M | capi.DEF_M
    ov, od = ora.assemble(p, conn2, coords2, u2.reshape(-1), rowptr, colind, what, scale_a=0.7, scale_m=1.3)
    gv, gd = disc.assemble(what, u2.reshape(-1), scale_a=0.7, scale_m=1.3, scatter_mode=MODES[mode])
    eg, ee = parity.entry_errors(gv, ov, rowptr)
    assert eg < TOL and ee < TOL
    eg, ee = parity.entry_errors(gd, od)
    assert eg < TOL and ee < TOL
    assert np.all(gd.reshape(-1, nf)[~used] == 0.0)
    disc.close()


@pytest.mark.parametrize("elem", ["hex", "tet", "quad", "tri"])
def test_split_path_reuse_across_parameter_changes(ora, elem):
    """one context, many passes: the split path caches the static Jacobian part J0 per mesh. It must follow a
    change of the laplace flag (J0 rebuilt), of viscosity / density (run-time scale of J0), and a detour through the
    general rows kernel (exact Newton, FLOW) and back (record table re-laid out)."""
    coords, conn, u = parity.make_case(elem, SIZES[elem], seed=3)
    dim = coords.shape[1]
    E = ora.ELEM[elem]
    disc = pkg.NavierStokesFV1(FCTS[dim], "Inner")
    parity.configure(disc, upwind="lps", stab="fields")
    disc.set_grid(elem, conn, coords)
    rowptr, colind = ora.fv1_csr(E, conn, coords.shape[0])
    what = capi.JAC_A | capi.DEF_A
    steps = [dict(), dict(laplace=True), dict(laplace=True, visc=3e-3, density=1.7), dict(exact=1.0), dict(),
             dict(stab="flow"), dict(visc=5e-2), dict(laplace=True)]
    for st in steps:
        visc, dens = st.get("visc", 1e-2), st.get("density", 1.0)
        disc.set_kinematic_viscosity(visc)
        disc.set_density(dens)
        disc.set_laplace(bool(st.get("laplace", False)))
        disc.set_exact_jacobian(st.get("exact", 0.0))
        disc.set_stabilization(st.get("stab", "fields"), "raw")
        disc.set_upwind("lps")
        disc.prep_elem_loop()
        p = ora.make_params(elem=elem, upwind="lps", stab=st.get("stab", "fields"), diff_len="raw", kin_visc=visc, density=dens,
                            laplace=st.get("laplace", False), exact_jac=st.get("exact", 0.0))
        ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, what)
        gv, gd = disc.assemble(what, u, scatter_mode=capi.SCATTER_GATHER)
        eg, ee = parity.entry_errors(gv, ov, rowptr)
        assert eg < TOL and ee < TOL, ("jacobian", elem, st, eg, ee)
        eg, ee = parity.entry_errors(gd, od)
        assert eg < TOL and ee < TOL, ("defect", elem, st, eg, ee)
    disc.close()
