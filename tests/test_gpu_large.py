"""Parity at (near) full size through size-independent properties (the oracle cannot hold these sizes in seconds):
  * Picard property: the fix-point Jacobian reproduces the stiffness defect, J(u) u = d_A(u), for every upwind /
    stabilisation without source or time terms (tests/test_oracle_invariants.py proves it for the oracle);
  * the three scatter variants (owner-computes, coloured, atomic) agree;
  * flux conservation: the continuity/momentum defect summed over all nodes of a closed... (interior telescoping)."""
import numpy as np
import pytest
import torch

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen

pytestmark = pytest.mark.gpu


def _csr_matvec(rowptr, colind, vals, x):
    A = torch.sparse_csr_tensor(torch.from_numpy(rowptr).cuda(), torch.from_numpy(colind.astype(np.int64)).cuda(), vals,
                                size=(rowptr.size - 1, rowptr.size - 1))
    return A @ x


@pytest.mark.parametrize("elem,n,upwind,stab", [("hex", 64, "lps", "fields"), ("hex", 48, "positive", "flow"),
                                                 ("tet", 24, "skewed", "flow"), ("quad", 512, "full", "fields"),
                                                 ("tri", 384, "lps", "fields")])
def test_picard_property_and_variant_agreement_at_scale(elem, n, upwind, stab):
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.15, seed=7)
    dim = coords.shape[1]
    u = (meshgen.state_vortex3d if dim == 3 else meshgen.state_cavity2d)(coords, seed=8)
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    disc.set_kinematic_viscosity(1e-2)
    disc.set_upwind(upwind)
    disc.set_stabilization(stab)
    disc.set_grid(elem, conn, coords)
    disc.use_stream(torch.cuda.current_stream().cuda_stream)
    ud = torch.from_numpy(np.ascontiguousarray(u.reshape(-1))).cuda()
    rowptr, colind = disc.csr()
    ref = None
    for mode in (capi.SCATTER_GATHER, capi.SCATTER_COLORED, capi.SCATTER_ATOMIC):
        vals, dfc = disc.assemble(capi.JAC_A | capi.DEF_A, ud, scatter_mode=mode)
        disc.check_errors()
        r = _csr_matvec(rowptr, colind, vals, ud) - dfc
        assert float(r.abs().max() / dfc.abs().max()) < 1e-10, (mode, float(r.abs().max()))
        if ref is None:
            ref = (vals.clone(), dfc.clone())
        else:
            assert float((vals - ref[0]).abs().max() / ref[0].abs().max()) < 1e-12
            assert float((dfc - ref[1]).abs().max() / ref[1].abs().max()) < 1e-12
    # every flux enters one node with + and another with -: the defect of each function sums to ~0 over the grid
    d = ref[1].reshape(-1, dim + 1)
    assert float(d.sum(dim=0).abs().max() / d.abs().sum(dim=0).max()) < 1e-12


def test_fvcr_picard_property_at_scale():
    coords, conn = meshgen.make_mesh("tet", 20, jitter=0.15, seed=9)
    disc = pkg.NavierStokesFVCR("u,v,w,p", "Inner")
    disc.set_kinematic_viscosity(1e-3)
    disc.set_upwind("full")
    disc.set_grid("tet", conn, coords)
    disc.use_stream(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(10)
    ud = torch.from_numpy(rng.uniform(-1, 1, disc.num_dofs)).cuda()
    rowptr, colind = disc.csr()
    vals, dfc = disc.assemble(capi.JAC_A | capi.DEF_A, ud, scatter_mode=capi.SCATTER_COLORED)
    r = _csr_matvec(rowptr, colind, vals, ud) - dfc
    assert float(r.abs().max() / dfc.abs().max()) < 1e-10
    v2, d2 = disc.assemble(capi.JAC_A | capi.DEF_A, ud, scatter_mode=capi.SCATTER_ATOMIC)
    assert float((v2 - vals).abs().max() / vals.abs().max()) < 1e-12
