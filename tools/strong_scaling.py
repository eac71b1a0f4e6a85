"""strong-scaling point of config 5 (SURVEY 8e / north_star item 4): Taylor-Green box [0, 2 pi]^3, G^3 hexahedra FIXED (default
256^3 = 16.8 M elements), FLOW + PositiveUpwind (dense ip systems), instationary Jacobian + A/M defect, split into px x py x pz
blocks over the ranks; one pass = assembly + interface summation (NCCL). One JSON line (rank 0).

  python tools/strong_scaling.py --cells 256                                             # 1 GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/strong_scaling.py --cells 256"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen, partition


def block(G, rank, world):
    px, py, pz = partition.block_dims(world)
    bx, by, bz = rank % px, (rank // px) % py, rank // (px * py)
    nx, ny, nz = G // px, G // py, G // pz
    L = 2 * np.pi
    lo = (L * bx / px, L * by / py, L * bz / pz)
    hi = (L * (bx + 1) / px, L * (by + 1) / py, L * (bz + 1) / pz)
    coords, conn = meshgen.hex_grid(nx, ny, nz, lo=lo, hi=hi)
    K, J, I = np.meshgrid(np.arange(nz + 1) + bz * nz, np.arange(ny + 1) + by * ny, np.arange(nx + 1) + bx * nx, indexing="ij")
    l2g = (I + (G + 1) * (J + (G + 1) * K)).ravel().astype(np.int64)
    Kl, Jl, Il = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    onface = ((Il == 0) | (Il == nx) | (Jl == 0) | (Jl == ny) | (Kl == 0) | (Kl == nz)).ravel()
    return coords, conn, dict(l2g=l2g, boundary=np.nonzero(onface)[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--mode", default="colored", choices=["colored", "atomic"])
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.time()
    coords, conn, iface = block(a.cells, rank, world)
    nu, dt = 1.0 / 1600, 1e-2
    u = meshgen.state_taylor_green(coords, t=0.0, nu=nu).reshape(-1)
    uo = meshgen.state_taylor_green(coords, t=-dt, nu=nu).reshape(-1)
    disc = pkg.NavierStokesFV1("u,v,w,p", "Inner", device=local)
    disc.set_kinematic_viscosity(nu); disc.set_upwind("positive"); disc.set_stabilization("flow")
    disc.set_grid("hex", conn, coords)
    ud, uod = torch.from_numpy(u).to(dev), torch.from_numpy(uo).to(dev)
    vals = torch.empty(disc.nnz, dtype=torch.float64, device=dev)
    dfc = torch.empty(disc.num_dofs, dtype=torch.float64, device=dev)
    exch = partition.InterfaceExchange(disc, iface, dev) if world > 1 else None
    setup = time.time() - t0
    what = capi.JAC_A | capi.DEF_A | capi.DEF_M
    mode = capi.SCATTER_COLORED if a.mode == "colored" else capi.SCATTER_ATOMIC

    def step():
        disc.assemble(what, ud, values=vals, defect=dfc, time_series=(ud, uod, dt), scale_a=dt, scale_m=1.0, scatter_mode=mode)
        if exch is not None:
            exch.sum_to_owner(vals, dfc)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    disc.check_errors()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    ne = torch.tensor([float(conn.shape[0])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ne, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"workload": "config5: hex %d^3 FIXED (%d elements), FLOW + PositiveUpwind, instationary J_A + d_A + d_M, %s scatter" % (a.cells, int(ne.item()), a.mode),
                          "n_gpus": world, "ms_per_pass": float(ms.item()), "elements_per_s": float(ne.item()) / (float(ms.item()) * 1e-3),
                          "scaling": "strong", "setup_s": setup, "exchange_bytes_rank0": exch.bytes_per_exchange() if exch else 0}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
