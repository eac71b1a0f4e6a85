"""development timing of the scatter modes (not the contract bench; see bench.py)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen

def run(elem, n, upwind, stab, modes, what=capi.JAC_A | capi.DEF_A, reps=5):
    t0 = time.time()
    coords, conn = meshgen.make_mesh(elem, n)
    dim = coords.shape[1]
    u = (meshgen.state_vortex3d if dim == 3 else meshgen.state_cavity2d)(coords)
    disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
    disc.set_kinematic_viscosity(1e-2); disc.set_upwind(upwind); disc.set_stabilization(stab)
    disc.set_grid(elem, conn, coords)
    t1 = time.time()
    ud = torch.from_numpy(u.reshape(-1)).cuda()
    disc.use_stream(torch.cuda.current_stream().cuda_stream)
    vals = torch.empty(disc.nnz, dtype=torch.float64, device="cuda"); dfc = torch.empty(disc.num_dofs, dtype=torch.float64, device="cuda")
    ne = conn.shape[0]
    bpe = (4 * conn.shape[1] * ne + 8 * dim * coords.shape[0] + 8 * disc.num_dofs + 8 * disc.nnz + 8 * disc.num_dofs) / ne
    print(f"{elem} n={n} elems={ne} nnz={disc.nnz} colors={disc.num_colors} setup={t1-t0:.1f}s B/elem={bpe:.0f}", flush=True)
    for name, mode in modes:
        for _ in range(2):
            disc.assemble(what, ud, values=vals, defect=dfc, scatter_mode=mode)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            disc.assemble(what, ud, values=vals, defect=dfc, scatter_mode=mode)
            ev[i + 1].record()
        torch.cuda.synchronize()
        disc.check_errors()
        ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        print(f"  {upwind:7s} {stab:6s} {name:8s} {ms:9.3f} ms  {ne/ms/1e6:8.3f} Gelem/s  {bpe*ne/ms/1e6:8.1f} GB/s algorithmic", flush=True)
    disc.close()

if __name__ == "__main__":
    modes = [("gather", capi.SCATTER_GATHER), ("colored", capi.SCATTER_COLORED), ("atomic", capi.SCATTER_ATOMIC)]
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    for upwind, stab in (("full", "fields"), ("lps", "fields"), ("lps", "flow")):
        run("hex", n, upwind, stab, modes)
    run("tet", max(8, n // 2), "full", "fields", modes)
    run("quad", 1024, "full", "fields", modes)
    run("tri", 1024, "lps", "fields", modes)
