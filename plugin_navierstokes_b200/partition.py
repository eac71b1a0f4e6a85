"""Element partition + interface summation for the multi-GPU path.

One process per GPU.  Elements are partitioned (recursive coordinate bisection / analytic blocks); every rank
keeps a local copy of every node its elements touch.  Nodes on partition boundaries are duplicated; the lowest
rank holding a node owns it (UG4: master), the others are slaves.  Each rank assembles its own elements into an
ADDITIVE local matrix / defect (exactly UG4's PST_ADDITIVE result, no communication), then
`InterfaceExchange.sum_to_owner` performs the interface step that replaces pcl on this path:

  * defect:  slave entries of shared nodes are sent to the owner and added (additive -> unique);
  * matrix:  slave rows of shared nodes are sent to the owner and added for the columns both ranks hold
             (the analogue of MatAddSlaveRowsToMasterRowOverlap0, fvcr/pcr_ilut.h:189).

Transport: torch.distributed point-to-point (NCCL over NVLink on GPUs, gloo in the CPU tests); pack / unpack-add
run as CUDA kernels of libnsb200 (nsb_pack / nsb_unpack_add) on device tensors, numpy indexing on CPU tensors.
"""
import numpy as np


# ----------------------------------------------------------------------------------------------------
# partitioning (host, numpy)
# ----------------------------------------------------------------------------------------------------
def rcb_partition(centroids, n_parts):
    """recursive coordinate bisection of element centroids -> part id per element (n_parts = power of two or any)"""
    part = np.zeros(centroids.shape[0], dtype=np.int32)

    def split(idx, lo, n):
        if n == 1:
            part[idx] = lo
            return
        c = centroids[idx]
        d = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        nl = n // 2
        order = np.argsort(c[:, d], kind="stable")
        cut = int(round(idx.size * nl / n))
        split(idx[order[:cut]], lo, nl)
        split(idx[order[cut:]], lo + nl, n - nl)

    split(np.arange(centroids.shape[0]), 0, n_parts)
    return part


def dual_graph(conn, n_node, min_shared):
    """element dual graph (scipy CSR, symmetric): elements are adjacent when they share at least `min_shared` nodes
    (= a face for min_shared = dim)"""
    import scipy.sparse as sp
    ne, nsh = conn.shape
    A = sp.csr_matrix((np.ones(ne * nsh, dtype=np.int32), (np.repeat(np.arange(ne), nsh), conn.reshape(-1))), shape=(ne, n_node))
    G = (A @ A.T).tocsr()
    G.setdiag(0)
    G.data = (G.data >= min_shared).astype(np.int32)
    G.eliminate_zeros()
    return G


def _bisect_graph(G, idx, frac, passes=8):
    """splits the sub-graph of elements `idx` into (left, right) with |left| ~ frac * |idx|: breadth-first level-set growing from
    a pseudo-peripheral element (the METIS "graph growing" initial partition), then Fiduccia-Mattheyses-style boundary
    refinement: elements with positive gain (more neighbours on the other side) change sides, largest gain first, the
    balance kept within one per cent"""
    from scipy.sparse.csgraph import breadth_first_order
    sub = G[idx][:, idx].tocsr()
    n = idx.size
    target = int(round(n * frac))
    # pseudo-peripheral start: two BFS sweeps (disconnected remainders are appended in index order)
    start = 0
    for _ in range(2):
        order = breadth_first_order(sub, start, directed=False, return_predecessors=False)
        start = int(order[-1])
    order = breadth_first_order(sub, start, directed=False, return_predecessors=False)
    if order.size < n:
        rest = np.setdiff1d(np.arange(n), order, assume_unique=False)
        order = np.concatenate([order, rest])
    side = np.ones(n, dtype=np.int8)
    side[order[:target]] = 0
    indptr, indices = sub.indptr, sub.indices
    deg = np.diff(indptr)
    tol = max(1, n // 100)
    for _ in range(passes):
        nb_side = side[indices]
        ext = np.add.reduceat(np.where(nb_side != np.repeat(side, deg), 1, 0), indptr[:-1]) if indices.size else np.zeros(n, int)
        ext[deg == 0] = 0
        gain = 2 * ext - deg                                   # cut edges removed minus cut edges created
        cand = np.nonzero(gain > 0)[0]
        if cand.size == 0:
            break
        cand = cand[np.argsort(-gain[cand], kind="stable")]
        n0 = int((side == 0).sum())
        moved = 0
        locked = np.zeros(n, dtype=bool)
        for v in cand:
            if locked[v]:
                continue
            s_ = side[v]
            new_n0 = n0 + (1 if s_ == 1 else -1)
            if abs(new_n0 - target) > tol:
                continue
            nb = indices[indptr[v]:indptr[v + 1]]
            e_now = int((side[nb] != s_).sum())
            if 2 * e_now - nb.size <= 0:                       # the gain changed through earlier moves of this pass
                continue
            side[v] = 1 - s_
            n0 = new_n0
            locked[nb] = True                                   # neighbours wait for the next pass (their gains are stale)
            moved += 1
        if moved == 0:
            break
    return idx[side == 0], idx[side == 1]


def graph_partition(conn, n_node, n_parts, dim=None):
    """graph partitioner for unstructured grids (configs 2 and 4: channel with cylinder, tetrahedra): recursive bisection of the
    element dual graph (face neighbours) by level-set growing + boundary refinement -- the METIS recipe without the multilevel
    coarsening. Returns the part id per element. Compared with rcb_partition it needs no coordinates and follows the mesh
    topology (holes, graded regions)."""
    ne, nsh = conn.shape
    G = dual_graph(conn, n_node, face_nodes(nsh, dim))
    part = np.zeros(ne, dtype=np.int32)

    def split(idx, lo, n):
        if n == 1 or idx.size == 0:
            part[idx] = lo
            return
        nl = n // 2
        a, b = _bisect_graph(G, idx, nl / n)
        split(a, lo, nl)
        split(b, lo + nl, n - nl)

    split(np.arange(ne), 0, n_parts)
    return part


def face_nodes(nsh, dim):
    """nodes two face-neighbours share: tri / quad 2, tet 3, hex 4, prism 3 (its triangular sides; the quadrilateral ones share 4).
    nsh = 4 is a quad in 2-D and a tet in 3-D"""
    return 2 if nsh == 3 or (nsh == 4 and dim == 2) else (3 if nsh in (4, 6) else 4)


def best_partition(conn, coords, n_parts):
    """the partition with the smaller interface (edge cut of the dual graph) of graph_partition and rcb_partition"""
    dim = coords.shape[1]
    ms = face_nodes(conn.shape[1], dim)
    pg = graph_partition(conn, coords.shape[0], n_parts, dim=dim)
    pr = rcb_partition(coords[conn].mean(axis=1), n_parts)
    cg, cr = edge_cut(conn, coords.shape[0], pg, ms), edge_cut(conn, coords.shape[0], pr, ms)
    return (pg, "graph", cg) if cg <= cr else (pr, "rcb", cr)


def edge_cut(conn, n_node, part, min_shared):
    """number of dual-graph edges between different parts (interface faces)"""
    G = dual_graph(conn, n_node, min_shared).tocoo()
    return int((part[G.row] != part[G.col]).sum() // 2)


def local_mesh(conn, coords, part, rank):
    """the rank's elements with local node numbering. returns (conn_local, coords_local, l2g)"""
    mine = conn[part == rank]
    l2g = np.unique(mine)
    g2l = -np.ones(coords.shape[0], dtype=np.int64)
    g2l[l2g] = np.arange(l2g.size)
    return g2l[mine].astype(np.int32), coords[l2g], l2g


def block_dims(world):
    """2x2x2-style block grid for `world` ranks (powers of two)"""
    dims = [1, 1, 1]
    d = 0
    w = world
    while w > 1:
        if w % 2:
            raise ValueError("world size must be a power of two")
        dims[d % 3] *= 2
        w //= 2
        d += 1
    return dims


def gid_noise(gid, seed, k):
    """deterministic per-node noise in [-1, 1) from the GLOBAL node id (identical on every rank)"""
    x = (gid.astype(np.uint64) * np.uint64(6364136223846793005) + np.uint64(1442695040888963407 + 7919 * seed + 104729 * k))
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xff51afd7ed558ccd)
    x ^= x >> np.uint64(33)
    return (x >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2.0 - 1.0


def block_problem(n, rank, world, noise=0.01):
    """config 3 on `world` GPUs: the unit cube meshed with (px*n) x (py*n) x (pz*n) hexahedra; rank owns one n^3 block.
    returns dict(coords, conn, u, iface) with iface = dict(l2g, boundary) for InterfaceExchange."""
    from . import meshgen
    px, py, pz = block_dims(world)
    bx, by, bz = rank % px, (rank // px) % py, rank // (px * py)
    lo = (bx / px, by / py, bz / pz)
    hi = ((bx + 1) / px, (by + 1) / py, (bz + 1) / pz)
    coords, conn = meshgen.hex_grid(n, n, n, lo=lo, hi=hi)
    # global node ids on the (px*n+1) x (py*n+1) x (pz*n+1) lattice
    i = np.arange(n + 1)
    K, J, I = np.meshgrid(i + bz * n, i + by * n, i + bx * n, indexing="ij")
    gx, gy = px * n + 1, py * n + 1
    l2g = (I + gx * (J + gy * K)).ravel().astype(np.int64)
    u = block_state(coords, l2g, noise)
    # candidates for shared nodes: the block faces
    li = np.arange(n + 1)
    Kl, Jl, Il = np.meshgrid(li, li, li, indexing="ij")
    onface = ((Il == 0) | (Il == n) | (Jl == 0) | (Jl == n) | (Kl == 0) | (Kl == n)).ravel()
    return dict(coords=coords, conn=conn, u=u, iface=dict(l2g=l2g, boundary=np.nonzero(onface)[0]))


def block_state(coords, gid, noise=0.01):
    """the state of block_problem as a function of the coordinates and the GLOBAL node ids"""
    x, y, z = (np.pi * coords[:, d] for d in range(3))
    xi = [1.0 + noise * gid_noise(gid, 3, k) for k in range(3)]
    return np.stack([np.sin(x) * np.cos(y) * np.cos(z) * xi[0] + 0.05,
                     -0.5 * np.cos(x) * np.sin(y) * np.cos(z) * xi[1] - 0.03,
                     -0.5 * np.cos(x) * np.cos(y) * np.sin(z) * xi[2] + 0.02,
                     np.cos(x) * np.cos(y) * np.cos(z)], axis=1)


def block_problem_global(n, world, noise=0.01):
    """the single-domain version of block_problem: (coords, conn, u) of the whole (px n) x (py n) x (pz n) mesh, node ids =
    the global ids of block_problem's l2g"""
    from . import meshgen
    px, py, pz = block_dims(world)
    coords, conn = meshgen.hex_grid(px * n, py * n, pz * n)
    return coords, conn, block_state(coords, np.arange(coords.shape[0], dtype=np.int64), noise)


def owner_rows_error(rowptr, colind, vals, dfc, l2g, owner, rank, g_rowptr, g_colind, g_vals, g_dfc, nf, nodes=None):
    """largest deviation of the OWNER's rows (matrix entries at the columns the rank holds, and defect entries) from the
    single-domain result, relative to the largest magnitude of the global matrix / defect. nodes: local nodes to check
    (default: every node the rank owns); rows of shared nodes are complete only after InterfaceExchange.sum_to_owner."""
    import scipy.sparse as sp
    n_g = g_rowptr.size - 1
    G = sp.csr_matrix((g_vals, g_colind, g_rowptr), shape=(n_g, n_g))
    sel = np.nonzero(owner == rank)[0] if nodes is None else np.asarray(nodes)
    sel = sel[owner[sel] == rank]
    emax, dmax = 0.0, 0.0
    gscale, dscale = np.abs(g_vals).max(), np.abs(g_dfc).max()
    rows_l = (sel[:, None] * nf + np.arange(nf)[None, :]).reshape(-1)
    rows_g = (l2g[sel][:, None] * nf + np.arange(nf)[None, :]).reshape(-1)
    for rl, rg in zip(rows_l, rows_g):
        seg = slice(rowptr[rl], rowptr[rl + 1])
        cl = colind[seg]
        cg = l2g[cl // nf] * nf + cl % nf
        ref = np.asarray(G[rg, cg].todense()).reshape(-1)
        emax = max(emax, float(np.abs(vals[seg] - ref).max()) if cl.size else 0.0)
    if dfc is not None and rows_l.size:
        dmax = float(np.abs(dfc[rows_l] - g_dfc[rows_g]).max())
    return emax / gscale, dmax / dscale


# ----------------------------------------------------------------------------------------------------
# interface summation
# ----------------------------------------------------------------------------------------------------
def shared_nodes(l2g_all, cand_all, rank):
    """for `rank`: {q: (local indices on rank, global ids)} of nodes also present on rank q, sorted by global id"""
    mine = l2g_all[rank][cand_all[rank]]
    order = np.argsort(mine)
    mine_sorted, loc_sorted = mine[order], cand_all[rank][order]
    out = {}
    for q in range(len(l2g_all)):
        if q == rank:
            continue
        other = l2g_all[q][cand_all[q]]
        common = np.intersect1d(mine_sorted, other, assume_unique=True)
        if common.size:
            pos = np.searchsorted(mine_sorted, common)
            out[q] = (loc_sorted[pos], common)
    return out


def owner_of(shared, rank, n_local):
    """owner rank per local node (lowest rank holding it)"""
    own = np.full(n_local, rank, dtype=np.int32)
    for q, (loc, _) in shared.items():
        if q < rank:
            own[loc] = np.minimum(own[loc], q)
    return own


def row_pairs(rowptr, colind, nf, l2g, rows_loc, cols_loc):
    """block pairs (a, b) of the local pattern with a in rows_loc and b in cols_loc: global (ga, gb) keys, the index of
    the block's first value entry (row fct 0, col fct 0) and the row stride"""
    inset = np.zeros(l2g.size, dtype=bool)
    inset[cols_loc] = True
    keys, first, stride = [], [], []
    for a in rows_loc:
        r0, r1 = rowptr[a * nf], rowptr[a * nf + 1]
        cols = colind[r0:r1:nf] // nf                       # neighbour nodes of a (block columns)
        sel = np.nonzero(inset[cols])[0]
        keys.append(np.stack([np.full(sel.size, l2g[a]), l2g[cols[sel]]], axis=1))
        first.append(r0 + sel * nf)
        stride.append(np.full(sel.size, r1 - r0))
    if not keys:
        return np.zeros((0, 2), np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(keys), np.concatenate(first), np.concatenate(stride)


def _pair_key(k, base):
    return k[:, 0].astype(np.int64) * base + k[:, 1].astype(np.int64)


class InterfaceExchange:
    """additive -> owner summation of the defect and of the matrix rows of shared nodes.

    iface: dict(l2g [n_local] global node ids, boundary = local candidates for shared nodes (None = all)).
    csr:   (rowptr, colind) of the local scalar CSR pattern (FV1 layout, dof = node*nf + fct).
    The plan is built once (host, numpy + object collectives); sum_to_owner moves only packed values."""

    def __init__(self, disc, iface, device=None, nf=None, csr=None, group_rank_world=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if group_rank_world is None else group_rank_world
        self.disc, self.device = disc, device
        self.launches = 0
        l2g = np.asarray(iface["l2g"], dtype=np.int64)
        cand = np.asarray(iface["boundary"] if iface.get("boundary") is not None else np.arange(l2g.size), dtype=np.int64)
        rowptr, colind = csr if csr is not None else disc.csr()
        nf = nf if nf is not None else (rowptr.size - 1) // l2g.size
        self.nf = nf
        # 1. who shares what (global ids of the candidates travel once)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, l2g[cand])
        mine = l2g[cand]
        order = np.argsort(mine)
        mine_sorted, loc_sorted = mine[order], cand[order]
        shared = {}
        for q in range(self.world):
            if q == self.rank:
                continue
            common = np.intersect1d(mine_sorted, gathered[q], assume_unique=True)
            if common.size:
                shared[q] = (loc_sorted[np.searchsorted(mine_sorted, common)], common)
        self.shared = shared
        self.owner = owner_of(shared, self.rank, l2g.size)
        # a node's owner must be one of the ranks we pair with; nodes shared by >2 ranks are sent by every slave
        # to the lowest rank (which all of them see in `shared`)
        # 2. per neighbour: defect dof lists and matrix block lists (agreed through sorted global keys)
        base = int(max(int(g.max()) if g.size else 0 for g in gathered)) + 1
        my_pairs = {}
        for q, (loc, gid) in shared.items():
            lo = min(self.rank, q)
            rows_sel = self.owner[loc] == lo          # rows of this pair: shared nodes whose (global) owner is the lower rank
            k, first, stride = row_pairs(rowptr, colind, nf, l2g, loc[rows_sel], loc)
            my_pairs[q] = (k, first, stride, loc[rows_sel], gid[rows_sel])
        # exchange the pair keys so both sides keep exactly the common ones
        send_obj = {q: _pair_key(v[0], base) for q, v in my_pairs.items()}
        all_objs = [None] * self.world
        dist.all_gather_object(all_objs, send_obj)
        self.plans = []
        for q in sorted(shared):
            k, first, stride, rloc, rgid = my_pairs[q]
            mykey = _pair_key(k, base)
            theirs = all_objs[q].get(self.rank, np.zeros(0, np.int64))
            common = np.intersect1d(mykey, theirs)
            o = np.argsort(mykey)
            pos = o[np.searchsorted(mykey[o], common)]
            f, st = first[pos], stride[pos]
            # value indices of the nf x nf entries of every common block, ordered (pair, rf, cf)
            rf = np.arange(nf)[None, :, None]
            cf = np.arange(nf)[None, None, :]
            midx = (f[:, None, None] + rf * st[:, None, None] + cf).reshape(-1)
            # defect dofs of the rows (ordered by global id)
            ro = np.argsort(rgid)
            didx = (rloc[ro][:, None] * nf + np.arange(nf)[None, :]).reshape(-1)
            role = "recv" if self.rank < q else "send"                   # lower rank owns the rows of this pair
            self.plans.append(dict(peer=q, role=role, midx=midx.astype(np.int64), didx=didx.astype(np.int64)))
        self._to_device()

    def _to_device(self):
        t = self.torch
        for p in self.plans:
            for key in ("midx", "didx"):
                p[key + "_t"] = t.from_numpy(p[key]).to(self.device) if self.device is not None else t.from_numpy(p[key])
            n = p["midx"].size + p["didx"].size
            p["buf"] = t.empty(n, dtype=t.float64, device=self.device)

    def _on_current_stream(self, tensor):
        """device kernels of the exchange run on torch's current stream: NCCL p2p ops are ordered against that stream"""
        import torch
        self.disc.use_stream(torch.cuda.current_stream(tensor.device).cuda_stream)

    def _pack(self, idx_t, src, out):
        if src.is_cuda:
            import ctypes as C
            from . import _capi as capi
            self._on_current_stream(src)
            rc = capi.lib().nsb_pack(self.disc._ctx, idx_t.numel(), C.c_void_p(idx_t.data_ptr()), C.c_void_p(src.data_ptr()),
                                     C.c_void_p(out.data_ptr()))
            assert rc == 0
            self.launches += 1 if idx_t.numel() else 0
        else:
            out.copy_(src[idx_t])

    def _unpack_add(self, idx_t, buf, dst):
        if dst.is_cuda:
            import ctypes as C
            from . import _capi as capi
            self._on_current_stream(dst)
            rc = capi.lib().nsb_unpack_add(self.disc._ctx, idx_t.numel(), C.c_void_p(idx_t.data_ptr()), C.c_void_p(buf.data_ptr()),
                                           C.c_void_p(dst.data_ptr()))
            assert rc == 0
            self.launches += 1 if idx_t.numel() else 0
        else:
            dst.index_add_(0, idx_t, buf)

    def enable_overlap(self):
        """interface-first assembly: the shared nodes become the priority nodes of the disc, so that
        assemble(what | PHASE_PRIORITY) -> start_sum_to_owner -> assemble(what | PHASE_REST) -> finish_sum_to_owner
        hides the exchange behind the assembly of the interior rows"""
        nodes = np.unique(np.concatenate([loc for loc, _ in self.shared.values()])) if self.shared else np.zeros(0, dtype=np.int64)
        self.disc.set_priority_nodes(nodes)
        return nodes

    def start_sum_to_owner(self, vals, dfc):
        """packs the slave rows and posts the sends / receives (NCCL orders them behind the pack kernels on the current stream
        and runs them on its own stream); returns the pending work handles for finish_sum_to_owner"""
        dist = self.dist
        ops = []
        for p in self.plans:
            nm = p["midx"].size
            if p["role"] == "send":
                if vals is not None:
                    self._pack(p["midx_t"], vals, p["buf"][:nm])
                if dfc is not None:
                    self._pack(p["didx_t"], dfc, p["buf"][nm:])
                ops.append(dist.P2POp(dist.isend, p["buf"], p["peer"]))
            else:
                ops.append(dist.P2POp(dist.irecv, p["buf"], p["peer"]))
        return dist.batch_isend_irecv(ops) if ops else []

    def finish_sum_to_owner(self, works, vals, dfc, zero_slaves=False):
        for w in works:
            w.wait()                                                     # the current stream waits for the transfers
        for p in self.plans:
            nm = p["midx"].size
            if p["role"] == "recv":
                if vals is not None:
                    self._unpack_add(p["midx_t"], p["buf"][:nm], vals)
                if dfc is not None:
                    self._unpack_add(p["didx_t"], p["buf"][nm:], dfc)
            elif zero_slaves:
                if vals is not None:
                    vals[p["midx_t"]] = 0.0
                if dfc is not None:
                    dfc[p["didx_t"]] = 0.0

    def sum_to_owner(self, vals, dfc, zero_slaves=False):
        """vals / dfc: the rank's additive CSR values / defect (torch tensors). After the call the owner's entries of
        shared rows hold the sum over all ranks (for the columns the pair shares); with zero_slaves the slave copies
        are cleared (PST_UNIQUE), otherwise they keep their partial values."""
        self.finish_sum_to_owner(self.start_sum_to_owner(vals, dfc), vals, dfc, zero_slaves)

    def copy_from_owner(self, vec):
        """owner -> copies over the same dof lists in the reverse direction (additive/unique -> consistent): after the
        call every rank's copies of a shared node hold the owner's `nf` values. Used to make the input state `u`
        consistent before an assembly and to turn the summed defect into a consistent vector."""
        dist = self.dist
        ops, nbuf = [], []
        for p in self.plans:
            nm, nd = p["midx"].size, p["didx"].size
            buf = p["buf"][nm:nm + nd]
            if p["role"] == "recv":                                       # the lower rank owns the rows of this pair
                self._pack(p["didx_t"], vec, buf)
                ops.append(dist.P2POp(dist.isend, buf, p["peer"]))
            else:
                ops.append(dist.P2POp(dist.irecv, buf, p["peer"]))
            nbuf.append(buf)
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for p, buf in zip(self.plans, nbuf):
            if p["role"] == "send" and p["didx"].size:
                vec[p["didx_t"]] = buf

    def bytes_per_exchange(self):
        return int(sum(8 * (p["midx"].size + p["didx"].size) for p in self.plans if p["role"] == "send"))
