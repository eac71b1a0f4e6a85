"""2-GPU test (skipped on single-GPU boxes): CUDA assembly per rank + NCCL interface summation vs the
single-domain CPU oracle."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import plugin_navierstokes_b200 as pkg
        from plugin_navierstokes_b200 import capi, meshgen, partition
        coords, conn = meshgen.hex_grid(6, 5, 4, jitter=0.2, seed=1)
        u = meshgen.state_vortex3d(coords, seed=2, noise=0.05)
        part = partition.rcb_partition(coords[conn].mean(axis=1), world)
        lconn, lcoords, l2g = partition.local_mesh(conn, coords, part, rank)
        disc = pkg.NavierStokesFV1("u,v,w,p", "Inner", device=rank)
        disc.set_kinematic_viscosity(1e-2)
        disc.set_upwind("lps")
        disc.set_stabilization("fields")
        disc.set_grid("hex", lconn, lcoords)
        disc.use_stream(torch.cuda.current_stream().cuda_stream)
        ud = torch.from_numpy(np.ascontiguousarray(u[l2g].reshape(-1))).cuda()
        vals, dfc = disc.assemble(capi.JAC_A | capi.DEF_A, ud)
        ex = partition.InterfaceExchange(disc, dict(l2g=l2g, boundary=None), torch.device("cuda", rank))
        ex.sum_to_owner(vals, dfc)
        # overlapped variant: interface rows first, exchange behind the interior rows -- bitwise the same sums
        ex.enable_overlap()
        what = capi.JAC_A | capi.DEF_A
        v2, d2 = torch.empty_like(vals), torch.empty_like(dfc)
        disc.assemble(what | capi.PHASE_PRIORITY, ud, values=v2, defect=d2)
        works = ex.start_sum_to_owner(v2, d2)
        disc.assemble(what | capi.PHASE_REST, ud, values=v2, defect=d2)
        ex.finish_sum_to_owner(works, v2, d2)
        torch.cuda.synchronize()
        assert torch.equal(v2, vals) and torch.equal(d2, dfc)
        cons = dfc.clone()
        ex.copy_from_owner(cons)                                 # unique -> consistent over NCCL
        torch.cuda.synchronize()
        disc.check_errors()
        rowptr, colind = disc.csr()
        q.put((rank, l2g, dfc.cpu().numpy(), ex.owner.copy(), ex.launches, cons.cpu().numpy(), vals.cpu().numpy(), rowptr, colind))
    finally:
        dist.destroy_process_group()


def test_two_gpu_defect_matches_single_domain(ora):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from plugin_navierstokes_b200 import meshgen
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    coords, conn = meshgen.hex_grid(6, 5, 4, jitter=0.2, seed=1)
    u = meshgen.state_vortex3d(coords, seed=2, noise=0.05)
    prm = ora.make_params(elem="hex", upwind="lps", stab="fields", kin_visc=1e-2)
    rowptr, colind = ora.fv1_csr(ora.HEX, conn, coords.shape[0])
    gv, gd = ora.assemble(prm, conn, coords, u, rowptr, colind, ora.JAC_A | ora.DEF_A)
    from plugin_navierstokes_b200 import partition
    for rank, l2g, d, own, launches, cons, lv, lrp, lci in res:
        # the summed MATRIX rows of every node the rank owns (interface rows included), at the columns the rank holds
        em, ed = partition.owner_rows_error(lrp, lci, lv, d, l2g, own, rank, rowptr, colind, gv, gd, 4)
        assert em < 1e-12 and ed < 1e-12, (rank, em, ed)
    gd = gd.reshape(-1, 4)
    for rank, l2g, d, own, launches, cons, lv, lrp, lci in res:
        mine = own == rank
        assert np.abs(d.reshape(-1, 4)[mine] - gd[l2g[mine]]).max() < 1e-12 * np.abs(gd).max()
        assert np.abs(cons.reshape(-1, 4) - gd[l2g]).max() < 1e-12 * np.abs(gd).max()     # every copy, after copy_from_owner
    assert sum(r[4] for r in res) > 0
