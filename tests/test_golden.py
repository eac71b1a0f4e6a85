"""Golden fixtures (tests/golden/): frozen oracle outputs on small seeded cases + hand-derived geometry values.
CPU: the oracle still reproduces them. GPU: the CUDA path reproduces them without the oracle in the loop."""
import os

import numpy as np
import pytest

from tests import parity
from tests.golden.make_golden import CASES as _CASES_R1, CASES_ELEM_TYPES
from tests.parity import TOL

CASES = _CASES_R1 + CASES_ELEM_TYPES
_GDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = {}
for _f in ("fv1_fvcr_golden.npz", "elem_types_golden.npz"):
    with np.load(os.path.join(_GDIR, _f)) as _z:
        GOLD.update({k: _z[k] for k in _z.files})


def _get(name, key):
    return GOLD[name + "/" + key]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_golden(ora, case):
    from tests.golden.make_golden import build
    fresh = build(case)
    for k, v in fresh.items():
        g = GOLD[k]
        if v.dtype.kind == "f":
            assert np.allclose(v, g, rtol=1e-13, atol=1e-13 * max(1.0, np.abs(g).max())), k
        else:
            assert np.array_equal(v, g), k


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_path_reproduces_golden(case):
    import plugin_navierstokes_b200 as pkg
    from plugin_navierstokes_b200 import capi
    name, disc_t, elem, n, upwind, stab, flags = case
    coords, conn, u = _get(name, "coords"), _get(name, "conn"), _get(name, "u")
    dim = coords.shape[1]
    fcts = "u,v,p" if dim == 2 else "u,v,w,p"
    disc = pkg.NavierStokes(fcts, "Inner", disc_t)
    disc.set_kinematic_viscosity(flags.get("kin_visc", 1e-2))
    disc.set_density(flags.get("density", 1.0))
    if disc_t == "fv1":
        disc.set_stabilization(stab)
        disc.set_upwind(upwind)
        disc.set_peclet_blend(flags.get("peclet_blend", False))
        disc.set_exact_jacobian(flags.get("exact_jac", 0.0))
        if flags.get("pac"):
            disc.set_pac_upwind(True)
        disc.set_grid(elem, conn, coords)
    else:
        disc.set_upwind(upwind)
        disc.set_grad_div(flags.get("grad_div", 0.0))
        disc.set_laplace(flags.get("laplace", False))
        disc.set_grid(elem, conn, coords, _get(name, "elem_sides"), int(_get(name, "n_side")))
    rp, ci = disc.csr()
    assert np.array_equal(rp, _get(name, "rowptr")) and np.array_equal(ci, _get(name, "colind"))
    ts = (_get(name, "s0"), _get(name, "s1"), flags["dt"]) if flags.get("time_dependent") else None
    modes = [capi.SCATTER_GATHER, capi.SCATTER_COLORED, capi.SCATTER_ATOMIC]
    for mode in modes:
        vals, dfc = disc.assemble(capi.JAC_A | capi.DEF_A, u, time_series=ts, scatter_mode=mode)
        eg, ee = parity.entry_errors(vals, _get(name, "values"), rp)
        assert eg < TOL and ee < TOL, (name, mode, eg, ee)
        eg, ee = parity.entry_errors(dfc, _get(name, "defect"))
        assert eg < TOL and ee < TOL, (name, mode, eg, ee)
