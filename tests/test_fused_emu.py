"""CPU parity of the FUSED PATCH KERNEL's arithmetic and tables (no GPU needed).

tests/cpp/emu_fused.cpp runs the NSB_HD lane functions of plugin_navierstokes_b200/csrc/ns_fused.cuh -- the same source the
CUDA kernel is compiled from -- and the host-side patch builder (ns_patch.h) thread by thread on the CPU. Here its CSR
values / defect are compared with the oracle to the north-star tolerance (<= 1e-12 relative), so a wrong formula, a wrong
patch table or a wrong summation plan is caught before any GPU time is spent. The GPU parity tests (-m gpu) then only have
to prove that the device executes the same functions correctly."""
import numpy as np
import pytest

from plugin_navierstokes_b200 import meshgen
from tests import fused_emu, parity
from tests.parity import TOL

JAC_A, DEF_A, JAC_M, DEF_M, RHS = 1, 2, 4, 8, 16
SIZES = {"tri": 11, "quad": 14, "tet": 4, "hex": 6}


def _check(ora, elem, coords, conn, u, what=JAC_A | DEF_A, upwind="lps", stab="fields", diff="raw", sol0=None, sol1=None,
           dt=0.0, scale_a=1.0, scale_m=1.0, ray_fast=1, use_geo=1, **flags):
    E = ora.ELEM[elem]
    rowptr, colind = ora.fv1_csr(E, conn, coords.shape[0])
    p = ora.make_params(elem=elem, upwind=upwind, stab=stab, diff_len=diff, kin_visc=flags.get("visc", 1e-2),
                        density=flags.get("density", 1.0), stokes=flags.get("stokes", False), laplace=flags.get("laplace", False),
                        peclet_blend=flags.get("peclet", False), source=flags.get("source"), dt=dt, time_dependent=sol0 is not None,
                        stab_upwind=flags.get("stab_upwind") or "same")
    ov, od = ora.assemble(p, conn, coords, u, rowptr, colind, what, sol0=sol0, sol1=sol1, scale_a=scale_a, scale_m=scale_m)
    gv, gd, st = fused_emu.assemble(elem, conn, coords, u, what, upwind=upwind, stab=stab, diff=diff, sol0=sol0, sol1=sol1, dt=dt,
                                    scale_a=scale_a, scale_m=scale_m, nnz=colind.shape[0], ray_fast=ray_fast, use_geo=use_geo, **flags)
    if what & (JAC_A | JAC_M):
        eg, ee = parity.entry_errors(gv, ov, rowptr)
        assert eg < TOL and ee < TOL, ("jacobian", elem, upwind, stab, eg, ee)
    if what & (DEF_A | DEF_M | RHS):
        eg, ee = parity.entry_errors(gd, od)
        assert eg < TOL and ee < TOL, ("defect", elem, upwind, stab, eg, ee)
    return gv, gd, st


@pytest.mark.parametrize("elem", ["hex", "tet", "quad", "tri"])
@pytest.mark.parametrize("upwind", ["no", "full", "skewed", "lps"])
@pytest.mark.parametrize("stab", ["fields", "none"])
@pytest.mark.parametrize("geo", [1, 0], ids=["geotab", "onthefly"])
def test_stationary_jac_def(ora, elem, upwind, stab, geo):
    """geo = 1: static SCVF geometry records (the default of the device path), geo = 0: geometry recomputed per SCVF"""
    coords, conn, u = parity.make_case(elem, SIZES[elem], seed=2)
    _check(ora, elem, coords, conn, u, upwind=upwind, stab=stab, use_geo=geo)


@pytest.mark.parametrize("elem", ["hex", "tet", "quad", "tri"])
@pytest.mark.parametrize("flags", [
    dict(peclet=True), dict(laplace=True), dict(stokes=True), dict(diff="fivepoint"), dict(diff="cor"),
    dict(density=1.3, visc=3e-3, source=[0.3, -0.2, 0.1]), dict(upwind="lps", stab_upwind="full"),
    dict(upwind="skewed", stab_upwind="lps", peclet=True),
], ids=lambda f: "-".join("%s=%s" % kv for kv in f.items()))
def test_flags(ora, elem, flags):
    flags = dict(flags)
    if "source" in flags and elem in ("tri", "quad"):
        flags["source"] = flags["source"][:2]
    coords, conn, u = parity.make_case(elem, SIZES[elem], seed=4)
    _check(ora, elem, coords, conn, u, **flags)
    if "diff" in flags:
        _check(ora, elem, coords, conn, u, use_geo=0, **flags)


@pytest.mark.parametrize("elem", ["hex", "tet", "quad", "tri"])
def test_instationary_parts_and_scales(ora, elem):
    coords, conn, u = parity.make_case(elem, SIZES[elem], seed=6)
    src = [0.3, -0.2, 0.1][: coords.shape[1]]
    s0, s1 = u * 1.01 + 0.003, u * 0.97 - 0.002
    _check(ora, elem, coords, conn, u, what=JAC_A | DEF_A | JAC_M | DEF_M | RHS, sol0=s0, sol1=s1, dt=0.05, scale_a=0.7,
           scale_m=1.3, source=src, density=1.2)
    _check(ora, elem, coords, conn, u, what=JAC_M | DEF_M, upwind="full", scale_a=0.7, scale_m=1.3)
    _check(ora, elem, coords, conn, u, what=DEF_A)
    _check(ora, elem, coords, conn, u, what=JAC_A)


def test_beta_accumulates(ora):
    coords, conn, u = parity.make_case("hex", 5, seed=1)
    gv, gd, _ = _check(ora, "hex", coords, conn, u)
    v2, d2, _ = fused_emu.assemble("hex", conn, coords, u, JAC_A | DEF_A, upwind="lps", stab="fields", beta=1.0,
                                   values=gv.copy(), defect=gd.copy())
    assert np.allclose(v2, 2 * gv, rtol=1e-13, atol=1e-13 * np.abs(gv).max())
    assert np.allclose(d2, 2 * gd, rtol=1e-13, atol=1e-13 * np.abs(gd).max())


@pytest.mark.parametrize("upwind", ["skewed", "lps"])
def test_predicted_ray_search_equals_ordered_search(ora, upwind):
    """hex: the predicted-side ray search (star-shaped elements only) must give what the ordered search gives; on a
    jittered grid a few elements are not star-shaped w.r.t. an ip and keep the ordered search."""
    coords, conn, u = parity.make_case("hex", 7, seed=8, jitter=0.25)
    v1, d1, st1 = _check(ora, "hex", coords, conn, u, upwind=upwind, ray_fast=1)
    v0, d0, st0 = _check(ora, "hex", coords, conn, u, upwind=upwind, ray_fast=0)
    assert st1["ray_fast"] == 1 and st0["ray_fast"] == 0
    assert 0 < st1["n_not_star_shaped"] < conn.shape[0] // 4
    assert np.abs(v1 - v0).max() <= 1e-13 * np.abs(v0).max()
    assert np.abs(d1 - d0).max() <= 1e-13 * np.abs(d0).max()


def test_bench_input_unjittered_hex(ora):
    """the contract bench input (unjittered hex_grid + state_vortex3d(seed=3, 1 % noise)) at a small size: axis-aligned
    faces, ray cuts on face diagonals / edges, every element star-shaped (predicted-side search everywhere)"""
    coords, conn = meshgen.hex_grid(9, 9, 9)
    u = meshgen.state_vortex3d(coords, seed=3)
    gv, gd, st = _check(ora, "hex", coords, conn, u, upwind="lps", stab="fields")
    assert st["n_not_star_shaped"] == 0 and st["max_nodes"] == 32 and st["max_work"] == 512
    # 4 x 4 x 2 node tiles: SCVFs on tile boundaries are evaluated twice (1 + (1/4 + 1/4 + 1/2) / 3 in the interior)
    assert st["scvf_evals"] / st["n_scvf"] < 1.34


def test_zero_velocity_guard_and_axis_aligned_flow(ora):
    """exact zero velocity at some ips (|u| < 1e-14 guard of Skewed / LPS, upwind.cpp:407-413,531-537) and a flow that is
    exactly axis-aligned on an unjittered grid (ray cuts exactly on edges of the reference triangulation)"""
    coords, conn = meshgen.hex_grid(6, 6, 6)
    u = np.zeros((coords.shape[0], 4))
    u[:, 0] = np.where(coords[:, 2] > 0.5, 1.0, 0.0)          # u = (1, 0, 0) in the upper half, exactly 0 below
    u[:, 3] = coords[:, 0]
    for upwind in ("skewed", "lps"):
        _check(ora, "hex", coords, conn, u.reshape(-1), upwind=upwind)


@pytest.mark.parametrize("elem", ["hex", "tri"])
def test_unreferenced_nodes_and_ragged_valence(ora, elem):
    coords, conn, u = parity.make_case(elem, SIZES[elem], seed=5)
    dim = coords.shape[1]
    nf = dim + 1
    keep = np.ones(conn.shape[0], bool)
    keep[3:9] = False
    conn = conn[keep]
    mid = coords.shape[0] // 2
    coords2 = np.concatenate([coords[:mid], [[9.0] * dim], coords[mid:], [[8.0] * dim, [7.0] * dim]])
    conn2 = np.where(conn >= mid, conn + 1, conn).astype(np.int32)
    u = u.reshape(-1, nf)
    u2 = np.concatenate([u[:mid], np.full((1, nf), 0.5), u[mid:], np.full((2, nf), -0.25)]).reshape(-1)
    used = np.zeros(coords2.shape[0], bool)
    used[conn2.ravel()] = True
    gv, gd, st = _check(ora, elem, coords2, conn2, u2, what=JAC_A | DEF_A | JAC_M | DEF_M, scale_a=0.7, scale_m=1.3)
    assert np.all(gd.reshape(-1, nf)[~used] == 0.0)


def test_channel_mesh_with_hole(ora):
    """config 2 look-alike: triangulated channel with a cylinder hole and a parabolic inflow state (v == 0 exactly)"""
    coords, conn = meshgen.tri_grid(44, 10, lo=(0.0, 0.0), hi=(2.2, 0.41), hole=(0.2, 0.2, 0.05), jitter=0.2, seed=2)
    u = meshgen.state_channel2d(coords, seed=2)
    _check(ora, "tri", coords, conn, u, upwind="lps", stab="fields", visc=1e-3)
