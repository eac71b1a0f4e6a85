"""a handful of small assemblies through every kernel family, for compute-sanitizer runs:
   compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_cases.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen

JD = capi.JAC_A | capi.DEF_A


def fv1(elem, n, upwind, stab, exact=0.0, mode=capi.SCATTER_GATHER, what=JD, td=False):
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=1)
    dim = coords.shape[1]
    u = meshgen.random_state(coords.shape[0], dim + 1, seed=2)
    d = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    d.set_kinematic_viscosity(0.02); d.set_upwind(upwind)
    if stab:
        d.set_stabilization(stab)
    d.set_exact_jacobian(exact)
    d.set_grid(elem, conn, coords)
    ts = (u.reshape(-1), 0.98 * u.reshape(-1), 1e-2) if td else None
    v, f = d.assemble(what, u.reshape(-1), time_series=ts, scatter_mode=mode)
    assert np.isfinite(v).all() and np.isfinite(f).all()
    if mode == capi.SCATTER_GATHER and not td:
        df = d.assemble_resident(what, u.reshape(-1))
        y = d.apply_jacobian(u.reshape(-1))
        assert np.isfinite(df).all() and np.isfinite(y).all()
    d.close()
    print("ok", elem, upwind, stab, exact, mode, flush=True)


fv1("hex", 5, "lps", "fields")                                   # split path: flux kernel + owner-lane rows kernel, j0 kernel, SpMV
fv1("tet", 4, "lps", "fields")
fv1("hex", 4, "skewed", "flow")                                  # general path
fv1("quad", 9, "full", "fields", exact=1.0)                      # general path, exact Newton
fv1("tri", 12, "lps", "fields")                                  # fused patch kernel (two CTAs per SM)
fv1("quad", 9, "lps", "fields")
fv1("hex", 4, "positive", "flow", mode=capi.SCATTER_COLORED, what=JD | capi.DEF_M, td=True)      # dense ip systems
fv1("hex", 4, "lps", "fields", mode=capi.SCATTER_ATOMIC)
fv1("tet", 4, "full", "fields", mode=capi.SCATTER_COLORED)
coords, conn = meshgen.make_mesh("tet", 3, jitter=0.2, seed=1)
es, ns = meshgen.element_sides("tet", conn)
d = pkg.NavierStokesFVCR("u,v,w,p", "Inner")
d.set_kinematic_viscosity(1e-2); d.set_upwind("lps")
d.set_grid("tet", conn, coords, es, ns)
u = np.random.default_rng(0).uniform(-1, 1, d.num_dofs)
for mode in (capi.SCATTER_GATHER, capi.SCATTER_COLORED):
    v, f = d.assemble(JD | capi.DEF_M, u, scatter_mode=mode)
    assert np.isfinite(v).all()
d.close()
print("ok fvcr", flush=True)
