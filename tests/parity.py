"""shared helpers of the GPU parity tests: build the same problem for the oracle and the device path"""
import numpy as np

from plugin_navierstokes_b200 import meshgen

TOL = 1e-12          # north_star: <= 1e-12 relative per defect entry and per Jacobian nonzero


def entry_errors(a, b, rowptr=None):
    """(max |a-b| / max|b|,  max per-entry relative error).

    The per-entry error is measured against max(|b_i|, 1e-2 * s_i), s_i = the largest magnitude in the
    entry's matrix row (or the vector's max): an entry that is the sum of cancelling flux contributions
    is only defined to eps * (size of the contributions), not eps * (its own size); the floor admits an
    absolute error of 1e-14 * (row scale) on such entries."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    gmax = np.abs(b).max()
    if gmax == 0:
        return np.abs(a).max(), np.abs(a).max()
    diff = np.abs(a - b)
    if rowptr is not None:
        rowptr = np.asarray(rowptr)
        lens = np.diff(rowptr)
        rowmax = np.zeros(lens.shape[0])
        ne = lens > 0                                            # reduceat cannot handle empty rows (unreferenced nodes)
        rowmax[ne] = np.maximum.reduceat(np.abs(b), rowptr[:-1][ne])
        s = np.repeat(rowmax, lens)
    else:
        s = np.full(b.shape, gmax)
    denom = np.maximum(np.abs(b), 1e-2 * np.maximum(s, 1e-300))
    return diff.max() / gmax, (diff / denom).max()


def make_case(elem, n, seed=0, jitter=0.2, scale=1.0):
    coords, conn = meshgen.make_mesh(elem, n, jitter=jitter, seed=seed)
    dim = coords.shape[1]
    coords = coords * scale
    if dim == 2:
        u = meshgen.state_cavity2d(coords / scale, seed=seed + 1, noise=0.05)
    else:
        u = meshgen.state_vortex3d(coords / scale, seed=seed + 1, noise=0.05)
    return coords, conn, u


def configure(disc, upwind="full", stab="fields", diff="raw", visc=1e-2, density=1.0, stokes=False, laplace=False,
              peclet=False, pac=False, exact=0.0, source=None, stab_upwind=None):
    disc.set_kinematic_viscosity(visc)
    disc.set_density(density)
    if stab == "none":
        from plugin_navierstokes_b200 import NavierStokesFV1WithoutStabilization
        disc.set_stabilization(NavierStokesFV1WithoutStabilization())
    elif stab is not None:
        disc.set_stabilization(stab, diff)
    if stab_upwind is not None:
        from plugin_navierstokes_b200 import CreateNavierStokesUpwind
        disc.stabilization().set_upwind(CreateNavierStokesUpwind(stab_upwind))
    if upwind is not None:
        disc.set_upwind(upwind)
        if stab == "none" and disc.stabilization().upwind() is None:
            disc.stabilization().set_upwind(disc._conv_upwind)
    disc.set_stokes(stokes)
    disc.set_laplace(laplace)
    disc.set_peclet_blend(peclet)
    disc.set_exact_jacobian(exact)
    if source is not None:
        disc.set_source(source)
    if pac:
        disc.set_pac_upwind(True)
