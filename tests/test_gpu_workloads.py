"""GPU parity on the ACTUAL BASELINE.json workloads (small sizes) and on the branches the jittered look-alikes never reach:

  (i)   config 5: unjittered hex box [0, 2 pi]^3 + Taylor-Green state (w == 0 exactly, nodes with u == 0), FLOW + PositiveUpwind,
        J_A + D_A + D_M with two time points -> a third of the SCVFs take PositiveUpwind's no-flux 1/2-1/2 branch
        (upwind.cpp:662-701), in all three scatter modes;
  (ii)  exact zero-velocity ips for Skewed / LPS (|u| < 1e-14 guard, upwind.cpp:407-413,531-537; CR: :605);
  (iii) config 3: the unjittered bench mesh + state_vortex3d(seed=3) at 8^3 .. 16^3 (axis-aligned faces, ray hits on face
        diagonals, first-hit tie-break);
  (iv)  config 2: tri_grid(hole=...) + state_channel2d;
  (v)   config 4: Kuhn tets, FVCR, A + M parts in `gather` mode.
Every case reports the STRICT per-entry relative error (no floor) next to the floored statistic the tolerance is applied to
(tests/parity.py): an entry that is a sum of cancelling fluxes is only defined to eps x the size of the fluxes."""
import numpy as np
import pytest

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen
from tests import parity
from tests.parity import TOL

pytestmark = pytest.mark.gpu
MODES = {"gather": capi.SCATTER_GATHER, "colored": capi.SCATTER_COLORED, "atomic": capi.SCATTER_ATOMIC}
JD = capi.JAC_A | capi.DEF_A
REPORT = []


def strict_rel(a, b, rowptr=None):
    """(max |a-b| / |b| over ALL entries with b != 0 (no floor), fraction of them above 1e-12,
        the same maximum over the entries that are not themselves round-off: |b| >= 1e-6 x the largest entry of the row)"""
    a, b = np.asarray(a), np.asarray(b)
    nzm = b != 0
    if not nzm.any():
        return 0.0, 0.0, 0.0
    r = np.zeros(b.shape)
    r[nzm] = np.abs(a - b)[nzm] / np.abs(b)[nzm]
    if rowptr is not None:
        lens = np.diff(np.asarray(rowptr))
        rowmax = np.zeros(lens.shape[0])
        ne = lens > 0
        rowmax[ne] = np.maximum.reduceat(np.abs(b), np.asarray(rowptr)[:-1][ne])
        scale = np.repeat(rowmax, lens)
    else:
        scale = np.full(b.shape, np.abs(b).max())
    sig = np.abs(b) >= 1e-6 * scale
    return float(r.max()), float((r[nzm] > 1e-12).mean()), float(r[sig & nzm].max()) if (sig & nzm).any() else 0.0


def _compare(tag, gv, gd, ov, od, rowptr, what):
    if what & (capi.JAC_A | capi.JAC_M):
        eg, ee = parity.entry_errors(gv, ov, rowptr)
        s, frac, ssig = strict_rel(gv, ov, rowptr)
        REPORT.append("%-70s J: global %.1e floored %.1e | strict, all nonzeros %.1e (%.2g of them > 1e-12) | strict, entries >= 1e-6 x row max %.1e" % (tag, eg, ee, s, frac, ssig))
        assert eg < TOL and ee < TOL, ("jacobian", tag, eg, ee)
    if what & (capi.DEF_A | capi.DEF_M | capi.RHS):
        eg, ee = parity.entry_errors(gd, od)
        s, frac, ssig = strict_rel(gd, od)
        REPORT.append("%-70s d: global %.1e floored %.1e | strict, all nonzeros %.1e (%.2g of them > 1e-12) | strict, entries >= 1e-6 x max %.1e" % (tag, eg, ee, s, frac, ssig))
        assert eg < TOL and ee < TOL, ("defect", tag, eg, ee)


def _fv1(ora, tag, elem, coords, conn, u, upwind, stab, mode, what=JD, visc=1e-2, ts=None, scale_a=1.0, scale_m=1.0):
    dim = coords.shape[1]
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    parity.configure(disc, upwind=upwind, stab=stab, visc=visc)
    disc.set_grid(elem, conn, coords)
    disc.prep_elem_loop()
    rowptr, colind = ora.fv1_csr(ora.ELEM[elem], conn, coords.shape[0])
    rp, ci = disc.csr()
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colind)
    s0 = s1 = None
    dt = 0.0
    if ts is not None:
        s0, s1, dt = ts
    p = ora.make_params(elem=elem, upwind=upwind, stab=stab, kin_visc=visc, dt=dt, time_dependent=ts is not None)
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, what, sol0=s0, sol1=s1, scale_a=scale_a, scale_m=scale_m)
    gv, gd = disc.assemble(what, u, time_series=ts, scale_a=scale_a, scale_m=scale_m, scatter_mode=MODES[mode])
    _compare("%s %s+%s %s" % (tag, upwind, stab, mode), gv, gd, ov, od, rowptr, what)
    disc.close()


@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
@pytest.mark.parametrize("n", [4, 7])
def test_config5_taylor_green_no_flux_branch(ora, mode, n):
    coords, conn = meshgen.hex_grid(n, n, n, lo=(0, 0, 0), hi=(2 * np.pi,) * 3)
    nu, dt = 1.0 / 1600, 1e-2
    u = meshgen.state_taylor_green(coords, t=0.0, nu=nu).reshape(-1)
    uo = meshgen.state_taylor_green(coords, t=-dt, nu=nu).reshape(-1)
    assert np.abs(u.reshape(-1, 4)[:, 2]).max() == 0.0                      # w == 0 exactly
    what = capi.JAC_A | capi.DEF_A | capi.DEF_M
    _fv1(ora, "config5 TG hex %d^3" % n, "hex", coords, conn, u, "positive", "flow", mode, what=what, visc=nu, ts=(u, uo, dt),
         scale_a=dt, scale_m=1.0)


@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
@pytest.mark.parametrize("upwind", ["skewed", "lps"])
@pytest.mark.parametrize("elem", ["hex", "quad", "tet", "tri"])
def test_zero_velocity_ips(ora, elem, upwind, mode):
    n = {"hex": 5, "tet": 4, "quad": 8, "tri": 8}[elem]
    coords, conn = meshgen.make_mesh(elem, n)                               # unjittered
    dim = coords.shape[1]
    u = np.zeros((coords.shape[0], dim + 1))
    upper = coords[:, dim - 1] > 0.5                                          # exactly zero velocity below
    if upwind == "lps":
        u[upper, 0] = 1.0                                                     # exactly axis-aligned: cuts on edges of the side triangulation
    else:
        # Skewed picks the NEAREST corner of the cut side: a discontinuous function of the cut point, ill-defined at exact
        # ties (ray through an edge / equidistant corners). A generic direction keeps the comparison meaningful.
        u[upper, :dim] = np.array([0.83, 0.31, 0.17])[:dim]
    u[:, dim] = coords[:, 0]
    for stab in ("fields", "flow"):
        _fv1(ora, "zero-velocity %s" % elem, elem, coords, conn, u.reshape(-1), upwind, stab, mode)


@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
@pytest.mark.parametrize("n", [8, 12, 16])
def test_config3_bench_input(ora, mode, n):
    coords, conn = meshgen.hex_grid(n, n, n)
    u = meshgen.state_vortex3d(coords, seed=3).reshape(-1)
    _fv1(ora, "config3 bench input hex %d^3" % n, "hex", coords, conn, u, "lps", "fields", mode)


@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
@pytest.mark.parametrize("upwind", ["no", "full", "skewed", "lps", "positive"])
def test_config2_channel_with_cylinder(ora, mode, upwind):
    coords, conn = meshgen.tri_grid(66, 14, lo=(0, 0), hi=(2.2, 0.41), jitter=0.2, seed=2, hole=(0.2, 0.2, 0.05))
    u = meshgen.state_channel2d(coords, seed=2).reshape(-1)
    _fv1(ora, "config2 tri channel+cylinder", "tri", coords, conn, u, upwind, "fields", mode, visc=1e-3)
    if upwind == "lps":
        u0 = meshgen.state_channel2d(coords, seed=2, noise=0.0).reshape(-1)   # v == 0 exactly
        _fv1(ora, "config2 tri channel+cylinder, v == 0", "tri", coords, conn, u0, upwind, "fields", mode, visc=1e-3)


@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
def test_config1_cavity_quads_exact_newton(ora, mode):
    coords, conn = meshgen.quad_grid(24, 24)
    u = meshgen.state_cavity2d(coords, seed=1).reshape(-1)
    disc = pkg.NavierStokesFV1("u,v,p", "Inner")
    parity.configure(disc, upwind="full", stab="fields", exact=1.0)
    disc.set_grid("quad", conn, coords)
    rowptr, colind = ora.fv1_csr(ora.QUAD, conn, coords.shape[0])
    p = ora.make_params(elem="quad", upwind="full", stab="fields", exact_jac=1.0)
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, JD)
    gv, gd = disc.assemble(JD, u, scatter_mode=MODES[mode])
    _compare("config1 quad cavity exact Newton %s" % mode, gv, gd, ov, od, rowptr, JD)
    disc.close()


@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
def test_config4_kuhn_fvcr_instationary(ora, mode):
    n = 4
    coords, conn = meshgen.tet_grid(4 * n, n, n, lo=(0, 0, 0), hi=(2.5, 0.41, 0.41), jitter=0.2, seed=4)
    es, n_side = meshgen.element_sides("tet", conn)
    rng = np.random.default_rng(4)
    u = np.concatenate([0.3 * rng.uniform(-1, 1, n_side * 3) + np.tile([0.3, 0.0, 0.0], n_side), rng.uniform(-1, 1, conn.shape[0])])
    disc = pkg.NavierStokesFVCR("u,v,w,p", "Inner")
    disc.set_kinematic_viscosity(1e-3)
    disc.set_upwind("full")
    disc.set_defect_upwind(True)
    disc.set_grid("tet", conn, coords, es, n_side)
    rowptr, colind = ora.fvcr_csr(ora.TET, es, n_side)
    p = ora.make_params(disc="fvcr", elem="tet", upwind="full", kin_visc=1e-3, defect_upwind=True)
    what = capi.JAC_A | capi.DEF_A | capi.JAC_M | capi.DEF_M
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, what, elem_sides=es, n_side=n_side, scale_a=1e-2, scale_m=1.0)
    gv, gd = disc.assemble(what, u, scale_a=1e-2, scale_m=1.0, scatter_mode=MODES[mode])
    _compare("config4 Kuhn tets FVCR A+M %s" % mode, gv, gd, ov, od, rowptr, what)
    disc.close()


def test_fused_kernel_serves_gather_mode():
    """2-D element types: the fused patch kernel (not the two-kernel split path, not an element kernel) must be what
    NSB_SCATTER_GATHER runs for FIELDS / no stabilisation with the fixed-point Jacobian: one launch per pass, no SCVF record
    table in HBM. (3-D element types take the split path by default, NSB_FUSED=1 selects the fused kernel there.)"""
    coords, conn = meshgen.quad_grid(40, 40)
    u = meshgen.state_cavity2d(coords, seed=1).reshape(-1)
    disc = pkg.NavierStokesFV1("u,v,p", "Inner")
    parity.configure(disc, upwind="lps", stab="fields")
    disc.set_grid("quad", conn, coords)
    assert disc.query(capi.Q_FUSED) == 1.0
    assert disc.query(capi.Q_SCVF_EVALS) / (4 * conn.shape[0]) < 1.25
    disc.assemble(JD, u)                        # builds the static tables once
    l0 = disc.launch_count
    disc.assemble(JD, u)
    assert disc.launch_count - l0 == 1
    disc.close()


def test_zz_report_strict_errors():
    """prints the strict (un-floored) per-entry statistics collected above (visible with -s / in the junit log)"""
    print("\n".join(REPORT))
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "strict_errors.txt"), "w") as f:
            f.write("\n".join(REPORT) + "\n")
