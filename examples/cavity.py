"""Lid-driven cavity (configs 1 / 3 of BASELINE.json at toy size) solved end to end on the GPU with the pieces of this library:

  Picard / Newton iteration:  J(u) du = -d(u),  u += du
    d, J      : nsb_assemble_resident  (FV1, FIELDS + upwind; the Jacobian never leaves the device)
    walls/lid : NavierStokesWall / NavierStokesInflowFV1 -> nsb_set_dirichlet, nsb_adjust_jacobian / nsb_adjust_vector
    linear    : dense LU of the resident matrix on the device + iterative refinement whose residual b - J x comes from
                nsb_apply_jacobian (the GPU-resident consumer of J); --linear bicgstab: Jacobi-preconditioned BiCGStab with
                nsb_apply_jacobian as the operator (no robust preconditioner for the saddle-point system is part of this library)

Everything the solver touches per iteration is a device pointer; only scalars (norms) reach the host.

  python examples/cavity.py [--dim 2|3] [--cells 16] [--re 100]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import plugin_navierstokes_b200 as pkg
from plugin_navierstokes_b200 import capi, meshgen


def bicgstab(apply_A, b, precond, tol=1e-10, maxit=2000):
    """right-preconditioned BiCGStab on CUDA tensors; apply_A(x) -> A x, precond(x) -> M^-1 x"""
    x = torch.zeros_like(b)
    r = b.clone()
    r0 = r.clone()
    rho = alpha = omega = torch.tensor(1.0, dtype=b.dtype, device=b.device)
    v = torch.zeros_like(b)
    p = torch.zeros_like(b)
    bn = float(b.norm())
    if bn == 0.0:
        return x, 0, 0.0
    for it in range(1, maxit + 1):
        rho_new = torch.dot(r0, r)
        beta = (rho_new / rho) * (alpha / omega)
        p = r + beta * (p - omega * v)
        ph = precond(p)
        v = apply_A(ph)
        alpha = rho_new / torch.dot(r0, v)
        s = r - alpha * v
        if float(s.norm()) <= tol * bn:
            x = x + alpha * ph
            return x, it, float(s.norm()) / bn
        sh = precond(s)
        t = apply_A(sh)
        omega = torch.dot(t, s) / torch.dot(t, t)
        x = x + alpha * ph + omega * sh
        r = s - omega * t
        rho = rho_new
        if float(r.norm()) <= tol * bn:
            return x, it, float(r.norm()) / bn
    return x, maxit, float(r.norm()) / bn


def direct_with_refinement(disc, jv, rowptr_t, colind_t, b, steps=2):
    """small problems: dense LU of the resident matrix on the device (torch.linalg.solve) + iterative refinement whose residual
    r = b - J x is formed by nsb_apply_jacobian, i.e. by the GPU-resident consumer of J"""
    n = b.numel()
    A = torch.sparse_csr_tensor(rowptr_t, colind_t, jv, size=(n, n)).to_dense()
    x = torch.linalg.solve(A, b)
    for _ in range(steps):
        r = disc.apply_jacobian(x, y=b.clone(), alpha=-1.0, beta=1.0)          # r = b - J x
        x = x + torch.linalg.solve(A, r)
    r = disc.apply_jacobian(x, y=b.clone(), alpha=-1.0, beta=1.0)
    return x, float(r.norm()) / max(float(b.norm()), 1e-300)


def solve(dim=2, cells=16, re=100.0, picard_tol=1e-8, verbose=True, linear="direct", upwind=None, elem=None, jitter=0.0, stab="fields"):
    """upwind: default FullUpwind in 2-D (config 1), LinearProfileSkewedUpwind in 3-D (config 3; on coarse 3-D grids the fixed-point
    iteration with FullUpwind ends in a 2-cycle when a face flux changes sign -- the upwind corner jumps, the LPS cut point moves
    continuously)"""
    upwind = upwind or ("full" if dim == 2 else "lps")
    dev = torch.device("cuda", 0)
    if dim == 2 and elem == "tri":                              # unstructured variant: jittered triangles
        coords, conn = meshgen.tri_grid(cells, cells, jitter=jitter, seed=1)
        fcts = "u,v,p"
    elif dim == 2:
        coords, conn = meshgen.quad_grid(cells, cells, jitter=jitter, seed=1)
        elem, fcts = "quad", "u,v,p"
    else:
        coords, conn = meshgen.hex_grid(cells, cells, cells)
        elem, fcts = "hex", "u,v,w,p"
    nf = dim + 1
    disc = pkg.NavierStokesFV1(fcts, "Inner")
    disc.set_kinematic_viscosity(1.0 / re)
    disc.set_upwind(upwind)
    disc.set_stabilization(stab)
    disc.set_grid(elem, conn, coords)
    # boundary conditions: no-slip walls, moving lid (top), one pressure dof pinned (all-Dirichlet velocity problem)
    lo, hi = coords.min(axis=0), coords.max(axis=0)
    on_bnd = np.zeros(coords.shape[0], dtype=bool)
    for d in range(dim):
        on_bnd |= np.isclose(coords[:, d], lo[d]) | np.isclose(coords[:, d], hi[d])
    lid = np.isclose(coords[:, dim - 1], hi[dim - 1])
    wall = pkg.NavierStokesWall(disc)
    wall.add(np.nonzero(on_bnd & ~lid)[0])
    inflow = pkg.NavierStokesInflowFV1(disc)
    inflow.add(lambda *x: (1.0,) + (0.0,) * (dim - 1), np.nonzero(lid)[0], coords)
    dw, vw = wall.dirichlet()
    di, vi = inflow.dirichlet()
    dofs = np.concatenate([dw, di, [nf - 1]])                   # + pressure of node 0
    vals_bc = np.concatenate([vw, vi, [0.0]])
    dofs, first = np.unique(dofs, return_index=True)
    vals_bc = vals_bc[first]
    disc.set_dirichlet(dofs)

    u = torch.zeros(disc.num_dofs, dtype=torch.float64, device=dev)
    disc.adjust_vector(u, vals_bc)                              # adjust_solution
    what = capi.JAC_A | capi.DEF_A
    hist = []
    rowptr, colind = disc.csr()
    rowptr_t, colind_t = torch.from_numpy(rowptr).to(dev), torch.from_numpy(colind.astype(np.int64)).to(dev)
    for it in range(40):
        d = disc.assemble_resident(what, u)                     # J stays on the device
        disc.adjust_jacobian()                                  # Dirichlet rows := unit rows
        disc.adjust_vector(d)                                   # adjust_defect
        dn = float(d.norm())
        hist.append(dn)
        if verbose:
            print("iteration %2d   |defect| = %.3e" % (it, dn))
        if dn < picard_tol * max(hist[0], 1e-300) or dn < 1e-13:
            break
        jv = _wrap(disc.resident_jacobian_ptr(), disc.nnz, dev)     # torch view of the resident values (no copy)
        if linear == "bicgstab":
            dinv = 1.0 / jv[_diag_index(rowptr, colind, dev)]
            du, nit, res = bicgstab(lambda x: disc.apply_jacobian(x), -d, lambda x: dinv * x, tol=1e-10, maxit=4000)
            msg = "BiCGStab: %d iterations" % nit
        else:
            du, res = direct_with_refinement(disc, jv, rowptr_t, colind_t, -d)
            msg = "dense LU + refinement through nsb_apply_jacobian"
        if verbose:
            print("               %s, relative residual %.1e" % (msg, res))
        u = u + du
    disc.upwind_name = upwind
    return disc, coords, conn, u, hist


def solve_fvcr(cells=16, re=100.0, picard_tol=1e-8, verbose=True, upwind="full", jitter=0.0, elem="tri"):
    """the same cavity with NavierStokesFVCR on triangles or quadrilaterals (Crouzeix-Raviart velocities on the sides, piecewise constant pressure; no
    stabilisation needed): Dirichlet values on the boundary SIDES (lid = sides with midpoint on y = 1), pressure of element 0 pinned.
    Returns (disc, coords, conn, elem_sides, u, history)."""
    dev = torch.device("cuda", 0)
    coords, conn = (meshgen.tri_grid if elem == "tri" else meshgen.quad_grid)(cells, cells, jitter=jitter, seed=1)
    es, n_side = meshgen.element_sides(elem, conn)
    disc = pkg.NavierStokesFVCR("u,v,p", "Inner")
    disc.set_kinematic_viscosity(1.0 / re)
    disc.set_upwind(upwind)
    disc.set_grid(elem, conn, coords, es, n_side)
    mid = np.zeros((n_side, 2))
    cnt = np.zeros(n_side)
    for k, sd in enumerate(meshgen.SIDES[elem]):
        np.add.at(mid, es[:, k], coords[conn[:, list(sd)]].mean(axis=1))
        np.add.at(cnt, es[:, k], 1)
    mid /= cnt[:, None]
    bnd = np.nonzero(cnt == 1)[0]                               # a boundary side has one element
    lid = np.isclose(mid[bnd, 1], coords[:, 1].max())
    dofs = np.concatenate([bnd * 2, bnd * 2 + 1, [n_side * 2]])            # u, v on the boundary sides + pressure of element 0
    vals_bc = np.concatenate([lid.astype(np.float64), np.zeros(bnd.size), [0.0]])
    order = np.argsort(dofs)
    dofs, vals_bc = dofs[order], vals_bc[order]
    disc.set_dirichlet(dofs)
    u = torch.zeros(disc.num_dofs, dtype=torch.float64, device=dev)
    disc.adjust_vector(u, vals_bc)
    what = capi.JAC_A | capi.DEF_A
    rowptr, colind = disc.csr()
    rowptr_t, colind_t = torch.from_numpy(rowptr).to(dev), torch.from_numpy(colind.astype(np.int64)).to(dev)
    hist = []
    for it in range(40):
        d = disc.assemble_resident(what, u)
        disc.adjust_jacobian()
        disc.adjust_vector(d)
        dn = float(d.norm())
        hist.append(dn)
        if verbose:
            print("iteration %2d   |defect| = %.3e" % (it, dn))
        if dn < picard_tol * max(hist[0], 1e-300) or dn < 1e-13:
            break
        jv = _wrap(disc.resident_jacobian_ptr(), disc.nnz, dev)
        du, res = direct_with_refinement(disc, jv, rowptr_t, colind_t, -d)
        u = u + du
    return disc, coords, conn, es, u, hist


def solve_extruded(elem="hex", cells=16, re=100.0, picard_tol=1e-8, verbose=True, upwind="lps"):
    """the 2-D cavity on the 3-D element types: the unit square extruded by ONE cell in z (hexahedra, their Kuhn split into six
    tetrahedra, or their split into two prisms), w = 0 at every node, no boundary disc on the two z faces (zero flux through them): the solution is the 2-D one,
    constant in z, so the literature tables of DrivenCavityLinesEval apply to FV1 on hexahedra / tetrahedra as well.
    Returns (disc, coords2d, conn2d, u2d, history): the z-averaged solution on the quadrilateral grid of the bottom layer, FV1 2-D
    layout node * 3 + (u, v, p)."""
    dev = torch.device("cuda", 0)
    h = 1.0 / cells
    gen = {"hex": meshgen.hex_grid, "tet": meshgen.tet_grid, "prism": meshgen.prism_grid}[elem]   # prisms: every cell split into two
    coords, conn = gen(cells, cells, 1, hi=(1.0, 1.0, h))
    nf = 4
    disc = pkg.NavierStokesFV1("u,v,w,p", "Inner")
    disc.set_kinematic_viscosity(1.0 / re)
    disc.set_upwind(upwind)
    disc.set_stabilization("fields")
    disc.set_grid(elem, conn, coords)
    x, y = coords[:, 0], coords[:, 1]
    wall = np.isclose(x, 0.0) | np.isclose(x, 1.0) | np.isclose(y, 0.0) | np.isclose(y, 1.0)
    lid = np.isclose(y, 1.0)
    wn = np.nonzero(wall)[0]
    dofs = np.concatenate([wn * nf, wn * nf + 1, np.arange(coords.shape[0]) * nf + 2, [nf - 1]])
    vals_bc = np.concatenate([lid[wn].astype(np.float64), np.zeros(wn.size), np.zeros(coords.shape[0]), [0.0]])
    order = np.argsort(dofs)
    dofs, vals_bc = dofs[order], vals_bc[order]
    disc.set_dirichlet(dofs)
    u = torch.zeros(disc.num_dofs, dtype=torch.float64, device=dev)
    disc.adjust_vector(u, vals_bc)
    what = capi.JAC_A | capi.DEF_A
    rowptr, colind = disc.csr()
    rowptr_t, colind_t = torch.from_numpy(rowptr).to(dev), torch.from_numpy(colind.astype(np.int64)).to(dev)
    hist = []
    for it in range(40):
        d = disc.assemble_resident(what, u)
        disc.adjust_jacobian()
        disc.adjust_vector(d)
        dn = float(d.norm())
        hist.append(dn)
        if verbose:
            print("iteration %2d   |defect| = %.3e" % (it, dn))
        if dn < picard_tol * max(hist[0], 1e-300) or dn < 1e-13:
            break
        jv = _wrap(disc.resident_jacobian_ptr(), disc.nnz, dev)
        du, res = direct_with_refinement(disc, jv, rowptr_t, colind_t, -d)
        u = u + du
    n2 = (cells + 1) ** 2                                         # both generators number x fastest, then y, then z
    uh = u.cpu().numpy().reshape(-1, nf)
    u2d = 0.5 * (uh[:n2] + uh[n2:2 * n2])[:, [0, 1, 3]]
    coords2d, conn2d = meshgen.quad_grid(cells, cells)
    return disc, coords2d, conn2d, np.ascontiguousarray(u2d).reshape(-1), hist


def _diag_index(rowptr, colind, dev):
    rows = np.repeat(np.arange(rowptr.size - 1), np.diff(rowptr))
    return torch.from_numpy(np.nonzero(colind == rows)[0]).to(dev)


def _wrap(ptr, n, dev):
    """torch view of a device pointer owned by the nsb context (no copy)"""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
    return torch.as_tensor(h, device=dev)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--cells", type=int, default=16)
    ap.add_argument("--re", type=float, default=100.0)
    ap.add_argument("--linear", default="direct", choices=["direct", "bicgstab"])
    ap.add_argument("--upwind", default=None, choices=["no", "full", "skewed", "lps"])
    a = ap.parse_args()
    disc, coords, conn, u, hist = solve(a.dim, a.cells, a.re, linear=a.linear, upwind=a.upwind)
    print("defect reduced by %.1e in %d iterations" % (hist[-1] / hist[0], len(hist) - 1))
