"""one FVCR pass (config 4 flavour) for ncu"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mode = {"colored": capi.SCATTER_COLORED, "atomic": capi.SCATTER_ATOMIC}[sys.argv[2] if len(sys.argv) > 2 else "colored"]
coords, conn = meshgen.tet_grid(4 * n, n, n, lo=(0, 0, 0), hi=(2.5, 0.41, 0.41), jitter=0.2, seed=4)
es, n_side = meshgen.element_sides("tet", conn)
rng = np.random.default_rng(4)
u = np.concatenate([0.3 * rng.uniform(-1, 1, n_side * 3) + np.tile([0.3, 0.0, 0.0], n_side), rng.uniform(-1, 1, conn.shape[0])])
disc = pkg.NavierStokesFVCR("u,v,w,p", "Inner")
disc.set_kinematic_viscosity(1e-3); disc.set_upwind("full"); disc.set_defect_upwind(True)
disc.set_grid("tet", conn, coords, es, n_side)
ud = torch.from_numpy(u).cuda()
vals = torch.empty(disc.nnz, dtype=torch.float64, device="cuda"); dfc = torch.empty(disc.num_dofs, dtype=torch.float64, device="cuda")
what = capi.JAC_A | capi.DEF_A | capi.DEF_M
for _ in range(3):
    disc.assemble(what, ud, values=vals, defect=dfc, scatter_mode=mode)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for i in range(3):
    ev[i].record(); disc.assemble(what, ud, values=vals, defect=dfc, scatter_mode=mode)
ev[3].record(); torch.cuda.synchronize()
print("fvcr tets", conn.shape[0], "colors", disc.num_colors, "ms", min(ev[i].elapsed_time(ev[i + 1]) for i in range(3)))
