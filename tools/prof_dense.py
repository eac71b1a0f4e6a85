"""single-configuration driver for ncu captures of the dense (PositiveUpwind) branch: python tools/prof_dense.py 48 [mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen
n = int(sys.argv[1]); mode = sys.argv[2] if len(sys.argv) > 2 else "gather"
coords, conn = meshgen.hex_grid(n, n, n, lo=(0, 0, 0), hi=(2 * np.pi,) * 3)
u = meshgen.state_taylor_green(coords, t=0.0); uo = meshgen.state_taylor_green(coords, t=-1e-2)
disc = pkg.NavierStokesFV1("u,v,w,p", "Inner")
disc.set_kinematic_viscosity(1.0 / 1600); disc.set_upwind("positive"); disc.set_stabilization("flow")
disc.set_grid("hex", conn, coords)
disc.use_stream(torch.cuda.current_stream().cuda_stream)
ud = torch.from_numpy(u.reshape(-1).copy()).cuda(); uod = torch.from_numpy(uo.reshape(-1).copy()).cuda()
vals = torch.empty(disc.nnz, dtype=torch.float64, device="cuda"); dfc = torch.empty(disc.num_dofs, dtype=torch.float64, device="cuda")
m = {"gather": capi.SCATTER_GATHER, "colored": capi.SCATTER_COLORED, "atomic": capi.SCATTER_ATOMIC}[mode]
import time
for _ in range(3):
    disc.assemble(capi.JAC_A | capi.DEF_A | capi.DEF_M, ud, values=vals, defect=dfc, time_series=(ud, uod, 1e-2), scatter_mode=m)
torch.cuda.synchronize(); t0 = time.perf_counter()
disc.assemble(capi.JAC_A | capi.DEF_A | capi.DEF_M, ud, values=vals, defect=dfc, time_series=(ud, uod, 1e-2), scatter_mode=m)
torch.cuda.synchronize(); disc.check_errors()
print("done", conn.shape[0], "%.3f ms" % ((time.perf_counter() - t0) * 1e3))
