"""device-resident timing of the five BASELINE.json configurations through the public API; one line per
configuration with elements/s, algorithmic GB/s and the fraction of the measured HBM roofline. Not the contract bench
(bench.py measures config 3 at the contract size); sizes here are the single-GPU points of §8d that fit comfortably.

  python tools/config_bench.py [scale]     # scale 1.0 = sizes below
  NSB_CONFIGS=6,7,8,9 python tools/config_bench.py   # the element types added at the end of round 2 (FV1 prisms, FVCR quad / hex)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen

PEAK = 6540.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


MODES = {"gather": capi.SCATTER_GATHER, "colored": capi.SCATTER_COLORED, "atomic": capi.SCATTER_ATOMIC}


def timeit(disc, what, ud, ts, reps=5):
    if os.environ.get("NSB_SCATTER"):                      # scatter-mode comparison: NSB_SCATTER=gather|colored|atomic
        disc.scatter_mode = MODES[os.environ["NSB_SCATTER"]]
    vals = torch.empty(disc.nnz, dtype=torch.float64, device="cuda")
    dfc = torch.empty(disc.num_dofs, dtype=torch.float64, device="cuda")
    for _ in range(3):
        disc.assemble(what, ud, values=vals, defect=dfc, time_series=ts)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        disc.assemble(what, ud, values=vals, defect=dfc, time_series=ts)
        ev[i + 1].record()
    torch.cuda.synchronize()
    disc.check_errors()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


def report(name, ne, nsh, dim, n_geo, n_dof, nnz, ntp, ms, launches):
    b = (4 * nsh * ne + 8 * dim * n_geo + 8 * n_dof * ntp + 8 * nnz + 8 * n_dof) / ne
    gbs = b * ne / ms / 1e6
    print("%-58s %9d el  %8.3f ms  %7.3f G el/s  %6.0f B/el  %7.1f GB/s  %5.1f %% of HBM  (%d launches/pass)"
          % (name, ne, ms, ne / ms / 1e6, b, gbs, 100 * gbs / PEAK, launches), flush=True)


def fv1(name, elem, coords, conn, u, upwind, stab, what, exact=0.0, td=None, visc=1e-2):
    dim = coords.shape[1]
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    disc.set_kinematic_viscosity(visc); disc.set_upwind(upwind); disc.set_stabilization(stab)
    disc.set_exact_jacobian(exact)
    disc.set_grid(elem, conn, coords)
    disc.use_stream(torch.cuda.current_stream().cuda_stream)
    ud = torch.from_numpy(np.ascontiguousarray(u.reshape(-1))).cuda()
    ts = None
    if td is not None:
        s1, dt = td
        ts = (ud, torch.from_numpy(np.ascontiguousarray(s1.reshape(-1))).cuda(), dt)
    l0 = disc.launch_count
    ms = timeit(disc, what, ud, ts)
    launches = (disc.launch_count - l0) // 8
    report(name, conn.shape[0], conn.shape[1], dim, coords.shape[0], disc.num_dofs, disc.nnz, 1 if td is None else 2, ms, launches)
    disc.close()


def main():
    sc = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    JD = capi.JAC_A | capi.DEF_A
    sel = set(int(x) for x in os.environ.get("NSB_CONFIGS", "1,2,3,4,5").split(","))   # subset of the configurations (knob sweeps)
    # config 1: 2-D cavity, quads, FullUpwind + FIELDS/RAW, exact Jacobian
    n = int(2048 * sc)
    if 1 in sel:
      coords, conn = meshgen.quad_grid(n, n)
      fv1("config1 quad %d^2 FULL+FIELDS exact_jacobian=1" % n, "quad", coords, conn, meshgen.state_cavity2d(coords, seed=1), "full", "fields", JD, exact=1.0)
    # config 2: channel with cylinder, triangles, LPS + FIELDS (headline upwind of the config)
    nx, ny = int(2816 * sc), int(524 * sc)
    if 2 in sel:
      coords, conn = meshgen.tri_grid(nx, ny, lo=(0, 0), hi=(2.2, 0.41), jitter=0.2, seed=2, hole=(0.2, 0.2, 0.05))
      fv1("config2 tri channel+cylinder %dx%d LPS+FIELDS" % (nx, ny), "tri", coords, conn, meshgen.state_channel2d(coords, seed=2), "lps", "fields", JD, visc=1e-3)
    # config 3 (reduced size; contract size in bench.py): hex, LPS + FIELDS
    n = int(128 * sc)
    if 3 in sel:
      coords, conn = meshgen.hex_grid(n, n, n)
      fv1("config3 hex %d^3 LPS+FIELDS" % n, "hex", coords, conn, meshgen.state_vortex3d(coords, seed=3), "lps", "fields", JD)
    # config 4: tets (Kuhn) + jitter, FVCR, Full upwind, time-dependent
    n = int(64 * sc)
    if 4 in sel:
      fvcr4(n)
    # config 5: Taylor-Green, hex, FLOW + PositiveUpwind (dense ip systems), instationary parts
    n = int(96 * sc)
    if 5 in sel:
      coords, conn = meshgen.hex_grid(n, n, n, lo=(0, 0, 0), hi=(2 * np.pi,) * 3)
      u = meshgen.state_taylor_green(coords, t=0.0)
      uo = meshgen.state_taylor_green(coords, t=-1e-2)
      fv1("config5 hex %d^3 FLOW+POSITIVE instationary (A + M defect)" % n, "hex", coords, conn, u, "positive", "flow", JD | capi.DEF_M, td=(uo, 1e-2), visc=1.0 / 1600)
    extra(sc)


def fvcr_generic(name, elem, coords, conn):
    """FVCR on the given grid, FullUpwind, A + M defect (the measured operations of config 4)"""
    JD = capi.JAC_A | capi.DEF_A
    dim = coords.shape[1]
    es, n_side = meshgen.element_sides(elem, conn)
    rng = np.random.default_rng(4)
    mean = [0.3, 0.0, 0.0][:dim]
    u = np.concatenate([0.3 * rng.uniform(-1, 1, n_side * dim) + np.tile(mean, n_side), rng.uniform(-1, 1, conn.shape[0])])
    disc = pkg.NavierStokesFVCR("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    disc.set_kinematic_viscosity(1e-3); disc.set_upwind("full"); disc.set_defect_upwind(True)
    disc.set_grid(elem, conn, coords, es, n_side)
    disc.use_stream(torch.cuda.current_stream().cuda_stream)
    ud = torch.from_numpy(u).cuda()
    l0 = disc.launch_count
    ms = timeit(disc, JD | capi.DEF_M, ud, None)
    report(name, conn.shape[0], conn.shape[1], dim, coords.shape[0], disc.num_dofs, disc.nnz, 1, ms, (disc.launch_count - l0) // 8)
    disc.close()


def extra(sc):
    """the element types added at the end of round 2 (NSB_CONFIGS=6,7,8,9): not BASELINE configurations, untuned kernels"""
    JD = capi.JAC_A | capi.DEF_A
    sel = set(int(x) for x in os.environ.get("NSB_CONFIGS", "").split(",") if x)
    if 6 in sel:
        n = int(96 * sc)
        coords, conn = meshgen.prism_grid(n, n, n)
        fv1("extra6 prism %d^3 x2 LPS+FIELDS (element kernel, coloured)" % n, "prism", coords, conn, meshgen.state_vortex3d(coords, seed=3), "lps", "fields", JD)
    if 7 in sel:
        n = int(96 * sc)
        coords, conn = meshgen.prism_grid(n, n, n, lo=(0, 0, 0), hi=(2 * np.pi,) * 3)
        u, uo = meshgen.state_taylor_green(coords, t=0.0), meshgen.state_taylor_green(coords, t=-1e-2)
        fv1("extra7 prism %d^3 x2 FLOW+POSITIVE instationary (A + M defect)" % n, "prism", coords, conn, u, "positive", "flow", JD | capi.DEF_M, td=(uo, 1e-2), visc=1.0 / 1600)
    if 8 in sel:
        n = int(2048 * sc)
        coords, conn = meshgen.quad_grid(n, n, jitter=0.2, seed=4)
        fvcr_generic("extra8 quad %d^2 FVCR FULL (A + M defect)" % n, "quad", coords, conn)
    if 9 in sel:
        n = int(128 * sc)
        coords, conn = meshgen.hex_grid(n, n, n, jitter=0.2, seed=4)
        fvcr_generic("extra9 hex %d^3 FVCR FULL (A + M defect)" % n, "hex", coords, conn)


def fvcr4(n):
    JD = capi.JAC_A | capi.DEF_A
    coords, conn = meshgen.tet_grid(4 * n, n, n, lo=(0, 0, 0), hi=(2.5, 0.41, 0.41), jitter=0.2, seed=4)
    es, n_side = meshgen.element_sides("tet", conn)
    rng = np.random.default_rng(4)
    u = np.concatenate([0.3 * rng.uniform(-1, 1, n_side * 3) + np.tile([0.3, 0.0, 0.0], n_side), rng.uniform(-1, 1, conn.shape[0])])
    disc = pkg.NavierStokesFVCR("u,v,w,p", "Inner")
    disc.set_kinematic_viscosity(1e-3); disc.set_upwind("full"); disc.set_defect_upwind(True)
    disc.set_grid("tet", conn, coords, es, n_side)
    disc.use_stream(torch.cuda.current_stream().cuda_stream)
    ud = torch.from_numpy(u).cuda()
    l0 = disc.launch_count
    ms = timeit(disc, JD | capi.DEF_M, ud, None)
    report("config4 tet %dx%dx%d x6 FVCR FULL (A + M defect)" % (4 * n, n, n), conn.shape[0], 4, 3, coords.shape[0], disc.num_dofs, disc.nnz, 1, ms, (disc.launch_count - l0) // 8)
    disc.close()


if __name__ == "__main__":
    main()
