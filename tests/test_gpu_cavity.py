"""end-to-end use of the device path: lid-driven cavity solved on the GPU (examples/cavity.py) -- assembly with the resident
Jacobian, Dirichlet post-pass of walls / lid, dense LU + iterative refinement whose residual comes from nsb_apply_jacobian. The nonlinear defect drops by
eight orders of magnitude and the converged state is a root of the ORACLE's defect as well."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


@pytest.mark.parametrize("dim,cells", [(2, 12), (3, 5)])
def test_cavity_converges_and_satisfies_the_oracle_defect(ora, dim, cells):
    import cavity
    disc, coords, conn, u, hist = cavity.solve(dim, cells, re=50.0, verbose=False)
    assert hist[-1] < 1e-8 * hist[0] and len(hist) < 40
    uh = u.cpu().numpy()
    elem = "quad" if dim == 2 else "hex"
    rowptr, colind = ora.fv1_csr(ora.ELEM[elem], conn, coords.shape[0])
    p = ora.make_params(elem=elem, upwind=disc.upwind_name, stab="fields", kin_visc=1.0 / 50.0)
    _, od = ora.assemble(p, conn, coords, uh, rowptr, colind, ora.DEF_A)
    nf = dim + 1
    lo, hi = coords.min(axis=0), coords.max(axis=0)
    on_bnd = np.zeros(coords.shape[0], dtype=bool)
    for d in range(dim):
        on_bnd |= np.isclose(coords[:, d], lo[d]) | np.isclose(coords[:, d], hi[d])
    free = np.ones(uh.size, dtype=bool)
    for d in range(dim):
        free[np.nonzero(on_bnd)[0] * nf + d] = False
    free[nf - 1] = False
    assert np.abs(od[free]).max() < 1e-8 * hist[0]
    lid = np.isclose(coords[:, dim - 1], hi[dim - 1])
    assert np.allclose(uh.reshape(-1, nf)[lid, 0], 1.0) and np.allclose(uh.reshape(-1, nf)[on_bnd & ~lid, :dim], 0.0)
    disc.close()
