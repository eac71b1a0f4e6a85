"""saves J / d of hex n^3 LPS+FIELDS (J+d, owner-computes) to compare two settings of the experiment knobs (README) bit by bit:
   python tools/env_compare.py 48 a.npz ; NSB_NOSPLIT=1 python tools/env_compare.py 48 b.npz ; python tools/env_compare.py cmp a.npz b.npz
(split vs general rows kernel: 6e-19 absolute on a 1e-3 matrix scale; ticket group / Z-curve / hints: bitwise equal)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if sys.argv[1] == "cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in ("vals", "dfc"):
        d = np.abs(a[k] - b[k]).max()
        print(k, "bitwise equal" if np.array_equal(a[k], b[k]) else "max abs diff %.3e (scale %.3e)" % (d, np.abs(a[k]).max()))
    sys.exit(0)
import torch
import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen
n = int(sys.argv[1])
coords, conn = meshgen.hex_grid(n, n, n, jitter=0.15, seed=5)
u = meshgen.state_vortex3d(coords, seed=3)
disc = pkg.NavierStokesFV1("u,v,w,p", "Inner")
disc.set_kinematic_viscosity(1e-2); disc.set_upwind("lps"); disc.set_stabilization("fields")
disc.set_grid("hex", conn, coords)
ud = torch.from_numpy(u.reshape(-1)).cuda()
disc.use_stream(torch.cuda.current_stream().cuda_stream)
vals = torch.empty(disc.nnz, dtype=torch.float64, device="cuda"); dfc = torch.empty(disc.num_dofs, dtype=torch.float64, device="cuda")
for _ in range(3):
    vals.fill_(float("nan")); dfc.fill_(float("nan"))
    disc.assemble(capi.JAC_A | capi.DEF_A, ud, values=vals, defect=dfc, scatter_mode=capi.SCATTER_GATHER)
torch.cuda.synchronize(); disc.check_errors()
np.savez(sys.argv[2], vals=vals.cpu().numpy(), dfc=dfc.cpu().numpy())
print("saved", sys.argv[2], "launches", disc.launch_count, "nan:", int(torch.isnan(vals).sum()), int(torch.isnan(dfc).sum()))
