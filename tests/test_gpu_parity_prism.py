"""GPU parity on prisms (NavierStokesFV1 is registered for Prism, fv1/navier_stokes_fv1.cpp:1542): the element kernels (coloured,
atomic, local) and the dense-ip-system kernel through the C ABI against the CPU oracle, same tolerance as the other element types
(1e-12 relative per defect entry / Jacobian nonzero), identical CSR sparsity. NSB_SCATTER_GATHER is served by the coloured
element kernel for prisms and must say so."""
import numpy as np
import pytest

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi
from tests import parity
from tests.parity import TOL
from tests.test_gpu_parity_fv1 import SIZES, _run_case

pytestmark = pytest.mark.gpu

SIZES.setdefault("prism", 3)                                  # 3^3 cells x 2 prisms, jittered (non-planar quadrilateral sides)


@pytest.mark.parametrize("mode", ["gather", "colored", "atomic"])
@pytest.mark.parametrize("upwind", ["no", "full", "skewed", "lps"])
@pytest.mark.parametrize("stab", ["fields", "flow", "none"])
def test_prism_stationary_jac_def(ora, mode, upwind, stab):
    _run_case(ora, "prism", mode, upwind=upwind, stab=stab)


@pytest.mark.parametrize("mode", ["colored", "atomic"])
@pytest.mark.parametrize("flags", [
    dict(exact=1.0), dict(peclet=True), dict(peclet=True, exact=0.5), dict(laplace=True), dict(stokes=True, upwind=None),
    dict(pac=True), dict(pac=True, exact=1.0, peclet=True), dict(pac=True, stab="flow", exact=1.0),
    dict(diff="fivepoint"), dict(diff="cor"), dict(stab="flow", diff="cor", upwind="lps"),
    dict(density=1.3, visc=3e-3, source=[0.3, -0.2, 0.1]), dict(upwind="lps", stab_upwind="full"),
], ids=lambda f: "-".join("%s=%s" % kv for kv in f.items()))
def test_prism_flags(ora, mode, flags):
    # PAC + exact Newton + Peclet blend on this jittered prism mesh has entries that are ill-conditioned sums: the oracle's own
    # result moves by 1.3e-12 (floored per-entry statistic; 4e-15 of the matrix scale) under 1-ulp perturbations of its inputs,
    # the device path differs from it by 1.1e-12 / 2.2e-15 -> conditioning-aware bound for this one combination
    cond = flags.get("pac") and flags.get("peclet") and flags.get("exact")
    _run_case(ora, "prism", mode, cond_aware=bool(cond), **dict(flags))


@pytest.mark.parametrize("mode", ["colored", "atomic"])
@pytest.mark.parametrize("stab", ["fields", "flow"])
def test_prism_instationary_parts(ora, mode, stab):
    what = capi.JAC_A | capi.DEF_A | capi.JAC_M | capi.DEF_M | capi.RHS
    _run_case(ora, "prism", mode, upwind="lps", stab=stab, what=what, time_dep=True, source=[0.3, -0.2, 0.1], density=1.2)
    _run_case(ora, "prism", mode, upwind="full", stab=stab, what=capi.DEF_M | capi.JAC_M)


@pytest.mark.parametrize("mode", ["colored", "atomic"])
@pytest.mark.parametrize("flags", [
    dict(stab="fields"), dict(stab="flow"), dict(stab="none"), dict(stab="flow", exact=1.0, peclet=True),
    dict(stab="fields", pac=True, exact=1.0), dict(stab="flow", diff="cor"),
    dict(stab="fields", upwind="full", stab_upwind="positive"), dict(stab="flow", upwind="positive", stab_upwind="lps", exact=1.0),
], ids=lambda f: "-".join("%s=%s" % kv for kv in f.items()))
def test_prism_positive_upwind_dense_branch(ora, mode, flags):
    """PositiveUpwind: 9 x 9 ip system per prism (stabilization.cpp:244-403, :590-771)"""
    flags = dict(flags)
    flags.setdefault("upwind", "positive")
    _run_case(ora, "prism", mode, **flags)


def test_prism_positive_upwind_instationary(ora):
    what = capi.JAC_A | capi.DEF_A | capi.JAC_M | capi.DEF_M | capi.RHS
    _run_case(ora, "prism", "colored", upwind="positive", stab="flow", what=what, time_dep=True, source=[0.3, -0.2, 0.1], density=1.2)


def test_prism_local_contributions_match_oracle(ora):
    """compat mode of the IElemDisc slots: per-element LocalMatrix / LocalVector blocks (24 x 24)"""
    coords, conn, u = parity.make_case("prism", 2, seed=5)
    disc = pkg.NavierStokesFV1("u,v,w,p", "Inner")
    parity.configure(disc, upwind="lps", stab="flow", exact=1.0)
    disc.set_grid("prism", conn, coords)
    J, d = disc.local_contributions(capi.JAC_A | capi.DEF_A, u)
    p = ora.make_params(elem="prism", upwind="lps", stab="flow", exact_jac=1.0)
    for e in range(conn.shape[0]):
        Jo, do = ora.fv1_elem(p, coords[conn[e]], u[conn[e]].T, ora.JAC_A | ora.DEF_A)
        assert np.abs(J[e] - Jo).max() <= TOL * np.abs(Jo).max(), e
        assert np.abs(d[e] - do).max() <= TOL * np.abs(do).max(), e
    disc.close()


def test_prism_gather_is_served_by_the_coloured_kernel_and_is_deterministic():
    coords, conn, u = parity.make_case("prism", 4, seed=9)
    disc = pkg.NavierStokesFV1("u,v,w,p", "Inner")
    parity.configure(disc, upwind="lps", stab="fields")
    disc.set_grid("prism", conn, coords)
    a = disc.assemble(capi.JAC_A | capi.DEF_A, u, scatter_mode=capi.SCATTER_GATHER)
    assert disc.query(capi.Q_LAST_SCATTER) == capi.SCATTER_COLORED
    b = disc.assemble(capi.JAC_A | capi.DEF_A, u, scatter_mode=capi.SCATTER_COLORED)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    disc.close()


def test_prism_unsupported_callers_say_so():
    coords, conn, u = parity.make_case("prism", 2)
    disc = pkg.NavierStokesFV1("u,v,w,p", "Inner")
    parity.configure(disc, upwind="full", stab="fields")
    disc.set_grid("prism", conn, coords)
    out = np.zeros(coords.shape[0])
    rc = capi.lib().nsb_diagnostic(disc._context(), 0, u.ctypes.data, 0.0, out.ctypes.data, capi.HOST)
    assert rc == capi.ERR_UNSUPPORTED
    disc.close()
