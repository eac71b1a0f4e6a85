"""Phased owner-computes assembly (nsb_set_priority_nodes, NSB_PHASE_PRIORITY / NSB_PHASE_REST): the rows of the priority nodes
first, the rest in a second call -- bitwise the unphased result on every path; paths without a separate rows kernel do the
whole pass in the first phase."""
import numpy as np
import pytest

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi
from tests import parity

pytestmark = pytest.mark.gpu
JD = capi.JAC_A | capi.DEF_A


@pytest.mark.parametrize("elem,n,kw,mode", [
    ("hex", 7, dict(upwind="lps", stab="fields"), capi.SCATTER_GATHER),                 # split path, owner-lane rows kernel
    ("tet", 5, dict(upwind="lps", stab="fields"), capi.SCATTER_GATHER),
    ("hex", 5, dict(upwind="lps", stab="flow"), capi.SCATTER_GATHER),                   # general rows kernel
    ("quad", 12, dict(upwind="full", stab="fields", exact=1.0), capi.SCATTER_GATHER),   # exact Newton: general rows kernel
    ("tri", 12, dict(upwind="lps", stab="fields"), capi.SCATTER_GATHER),                # fused 2-D kernel: one phase
    ("hex", 5, dict(upwind="lps", stab="fields"), capi.SCATTER_COLORED),                # element kernels: one phase
])
def test_two_phases_equal_one_pass(elem, n, kw, mode):
    import torch
    coords, conn, u = parity.make_case(elem, n, seed=3)
    dim = coords.shape[1]
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    parity.configure(disc, **kw)
    disc.set_grid(elem, conn, coords)
    ud = torch.from_numpy(u.reshape(-1)).cuda()
    v0, d0 = disc.assemble(JD, ud, scatter_mode=mode)
    torch.cuda.synchronize()
    lo = coords.min(axis=0)
    prio = np.nonzero(np.isclose(coords[:, 0], lo[0]) | np.isclose(coords[:, 1], lo[1]))[0][::-1]     # any order, any subset
    disc.set_priority_nodes(prio)
    nf = dim + 1
    rowptr, _ = disc.csr()
    nan = float("nan")
    v = torch.full_like(v0, nan); d = torch.full_like(d0, nan)
    disc.assemble(JD | capi.PHASE_PRIORITY, ud, values=v, defect=d, scatter_mode=mode)
    torch.cuda.synchronize()
    vh, dh = v.cpu().numpy(), d.cpu().numpy()
    # after the first phase the priority rows are final
    rows = (prio[:, None] * nf + np.arange(nf)[None, :]).reshape(-1)
    assert np.array_equal(dh[rows], d0.cpu().numpy()[rows])
    for r in rows[:: max(1, rows.size // 50)]:
        assert np.array_equal(vh[rowptr[r]:rowptr[r + 1]], v0.cpu().numpy()[rowptr[r]:rowptr[r + 1]])
    disc.assemble(JD | capi.PHASE_REST, ud, values=v, defect=d, scatter_mode=mode)
    torch.cuda.synchronize()
    assert torch.equal(v, v0) and torch.equal(d, d0)
    # the unphased call walks the same node order in one launch
    v2, d2 = disc.assemble(JD, ud, scatter_mode=mode)
    assert torch.equal(v2, v0) and torch.equal(d2, d0)
    disc.set_priority_nodes([])
    v3, d3 = disc.assemble(JD, ud, scatter_mode=mode)
    assert torch.equal(v3, v0) and torch.equal(d3, d0)
    with pytest.raises(pkg.UGError):
        disc.assemble(JD | capi.PHASE_PRIORITY, u.reshape(-1))                           # host vectors: no phases
    disc.close()


def test_served_scatter_mode_is_reported():
    """a GATHER request that the owner-computes path cannot serve is routed to the coloured element kernels -- and says so"""
    import torch
    from plugin_navierstokes_b200 import meshgen
    coords, conn, u = parity.make_case("hex", 4, seed=1)
    disc = pkg.NavierStokesFV1("u,v,w,p", "Inner")
    parity.configure(disc, upwind="lps", stab="fields")
    disc.set_grid("hex", conn, coords)
    disc.assemble(JD, u.reshape(-1), scatter_mode=capi.SCATTER_GATHER)
    assert disc.query(capi.Q_LAST_SCATTER) == capi.SCATTER_GATHER
    disc.assemble(JD, u.reshape(-1), scatter_mode=capi.SCATTER_ATOMIC)
    assert disc.query(capi.Q_LAST_SCATTER) == capi.SCATTER_ATOMIC
    parity.configure(disc, upwind="positive", stab="flow")
    disc.assemble(JD, u.reshape(-1), scatter_mode=capi.SCATTER_GATHER)
    assert disc.query(capi.Q_LAST_SCATTER) == capi.SCATTER_COLORED                  # dense ip systems
    disc.close()
    coords, conn = meshgen.make_mesh("tet", 3, jitter=0.1, seed=1)
    es, n_side = meshgen.element_sides("tet", conn)
    d2 = pkg.NavierStokesFVCR("u,v,w,p", "Inner")
    d2.set_kinematic_viscosity(1e-2); d2.set_upwind("full")
    d2.set_grid("tet", conn, coords, es, n_side)
    d2.assemble(JD, np.zeros(d2.num_dofs), scatter_mode=capi.SCATTER_GATHER)
    assert d2.query(capi.Q_LAST_SCATTER) == capi.SCATTER_ATOMIC       # CR: every entry has at most two contributions, 0 + a + b is order-free
    v = np.zeros(d2.nnz); d = np.zeros(d2.num_dofs)
    d2.assemble(JD, np.zeros(d2.num_dofs), values=v, defect=d, beta=1.0, scatter_mode=capi.SCATTER_GATHER)
    assert d2.query(capi.Q_LAST_SCATTER) == capi.SCATTER_COLORED      # beta != 0: a third term, coloured sweeps
    d2.close()
