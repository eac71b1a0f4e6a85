"""single-configuration driver for ncu captures: python tools/prof_one.py hex 64 lps fields gather [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen

elem, n, upwind, stab, mode = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
coords, conn = meshgen.make_mesh(elem, n)
dim = coords.shape[1]
u = (meshgen.state_vortex3d if dim == 3 else meshgen.state_cavity2d)(coords)
disc = pkg.NavierStokesFV1("u,v,w,p" if dim == 3 else "u,v,p", "Inner")
disc.set_kinematic_viscosity(1e-2); disc.set_upwind(upwind); disc.set_stabilization(stab)
disc.set_grid(elem, conn, coords)
ud = torch.from_numpy(u.reshape(-1)).cuda()
disc.use_stream(torch.cuda.current_stream().cuda_stream)
vals = torch.empty(disc.nnz, dtype=torch.float64, device="cuda"); dfc = torch.empty(disc.num_dofs, dtype=torch.float64, device="cuda")
m = {"gather": capi.SCATTER_GATHER, "colored": capi.SCATTER_COLORED, "atomic": capi.SCATTER_ATOMIC}[mode]
for _ in range(reps):
    disc.assemble(capi.JAC_A | capi.DEF_A, ud, values=vals, defect=dfc, scatter_mode=m)
torch.cuda.synchronize(); disc.check_errors()
print("done", conn.shape[0])
