"""Host mirror of the reference's Crouzeix-Raviart Cuthill-McKee ordering (solver-side preprocessing of FVCR systems):
`OrderCRCuthillMcKee` (fvcr/cr_reorder.h:352-393, registered in fvcr/register_fvcr.cpp:100) = `cr_get_connections`
(fvcr/cr_reorder.h:67-132) + `ComputeCRCuthillMcKeeOrder` (fvcr/cr_reorder.cpp:510-600, "original CR cuthill McKee without boost").

The dof layout is the one of the FVCR path here and of the reference's loops (`for i < minpind; i += 2`): velocity dofs side * dim + d
first, the element pressures behind them (minpind = n_side * dim). Pure graph work on the host, done once per grid; the permutation is
applied to the CSR pattern / vectors with `permute_csr` / `permute_vector`. The boost-based variants (CROrderCuthillMcKee / Sloan / King /
MinimumDegree, cr_reorder.cpp:195-505) depend on boost::graph's traversal details and are not mirrored.

Where the reference leaves the order open -- `std::sort` with the degree comparison is not stable -- this mirror uses a stable sort
(ties keep their insertion order). The quirks of the reference are kept and marked."""
import numpy as np


def cr_get_connections(elem_sides, n_side, dim):
    """adjacency lists of the CR dofs, fvcr/cr_reorder.h:67-132: every pair of dofs of an element is connected (both directions,
    insertion order = element order, then local order), finally every list is sorted by `computeDegree`. Returns (vvConnection, minpind)."""
    elem_sides = np.asarray(elem_sides)
    n_elem, ns = elem_sides.shape
    minpind = n_side * dim
    n = minpind + n_elem
    conn = [[] for _ in range(n)]
    seen = [set() for _ in range(n)]
    for e in range(n_elem):
        glob = [int(elem_sides[e, s]) * dim + d for s in range(ns) for d in range(dim)] + [minpind + e]
        for i in glob:
            for j in glob:
                if j not in seen[i]:
                    seen[i].add(j); conn[i].append(j)
                if i not in seen[j]:
                    seen[j].add(i); conn[j].append(i)
    degree = compute_degree(conn, minpind)
    for i in range(n):
        conn[i] = sorted(conn[i], key=lambda k: degree[k])          # :128-130 (std::sort -> stable here)
    return conn, minpind


def compute_degree(conn, minpind):
    """computeDegree, fvcr/cr_reorder.cpp:180-193. Velocity dof: number of connections. Pressure dof: the reference sums `degree[j]`
    over the POSITION j in its connection list (not over the connected index) -- kept as it is."""
    n = len(conn)
    degree = [0] * n
    for i in range(minpind):
        degree[i] = len(conn[i])
    for i in range(minpind, n):
        degree[i] = sum(degree[j] for j in range(len(conn[i])))
    return degree


def ComputeCRCuthillMcKeeOrder(conn, minpind, bReverse=False):
    """fvcr/cr_reorder.cpp:510-600 (2-D: two velocity components per side). Breadth-first over the PRESSURE graph (two pressures are
    adjacent when they share a side), starting at the pressure of minimum degree; visiting a pressure numbers its not yet numbered
    connected dofs (its own velocities, then itself -- sorted by degree, and a velocity's degree is below its pressure's) consecutively.
    bReverse is accepted and, as in the reference, not used. Returns newIndex [n] (old index -> new index)."""
    n = len(conn)
    conn = [list(c) for c in conn]
    num_p = n - minpind
    new_index = [n] * n
    degree = [0] * n
    pconn = [[] for _ in range(num_p)]
    for i in range(0, minpind, 2):                                   # :524 (i, i+1 = the two components of a side)
        ln = len(conn[i])
        degree[i] = degree[i + 1] = ln
        assoc = [k - minpind for k in conn[i] if k >= minpind]
        if len(assoc) > 2:
            raise ValueError("a side with more than two elements")
        if len(assoc) > 1:
            pi, pj = assoc[0], assoc[1]
            # :540-545: the reference searches for pj / pi but stores pj + minpind / pi + minpind, so the search never hits;
            # a pair arises from exactly one side, hence no duplicates either way
            pconn[pi].append(pj + minpind)
            pconn[pj].append(pi + minpind)
    minpdeg, minpdegind = n, minpind
    for i in range(minpind, n):                                      # :551-562
        degree[i] = sum(degree[k] for k in conn[i] if k != i)
        if degree[i] < minpdeg:
            minpdeg, minpdegind = degree[i], i
    plist = [minpdegind]
    count = 0
    for j in range(num_p):                                           # :570-599
        if j >= len(plist):
            raise ValueError("ComputeCRCuthillMcKeeOrder: the pressure graph is not connected (the reference reads past the end of its list here)")
        i = plist[j]
        conn[i] = sorted(conn[i], key=lambda k: degree[k])
        for k in conn[i]:
            if new_index[k] < n:
                continue
            new_index[k] = count
            count += 1
        pind = i - minpind
        pconn[pind] = sorted(pconn[pind], key=lambda k: degree[k])
        for k in pconn[pind]:
            if new_index[k] < n:
                continue
            if k not in plist[j:]:
                plist.append(k)
    return np.asarray(new_index, dtype=np.int64)


def OrderCRCuthillMcKee(elem_sides, n_side, dim=2, bReverse=False):
    """fvcr/cr_reorder.h:352-393 for one grid: newIndex of the FVCR dofs (apply with permute_csr / permute_vector, the role of
    DoFDistribution::permute_indices). The checks of the function-space types (:366-380) are the FVCR layout itself here."""
    if dim != 2:
        raise ValueError("OrderCRCuthillMcKee: the reference loops over velocity pairs (i, i + 1), i.e. two space dimensions")
    conn, minpind = cr_get_connections(elem_sides, n_side, dim)
    return ComputeCRCuthillMcKeeOrder(conn, minpind, bReverse)


def permute_vector(vec, new_index):
    """out[new_index[i]] = vec[i]"""
    out = np.empty_like(np.asarray(vec))
    out[np.asarray(new_index)] = vec
    return out


def permute_csr(rowptr, colind, values, new_index):
    """the matrix with rows and columns renumbered by new_index (P A P^T), columns sorted within a row; returns (rowptr, colind, values)"""
    import scipy.sparse as sp
    n = len(rowptr) - 1
    new_index = np.asarray(new_index)
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    A = sp.csr_matrix((np.asarray(values), (new_index[rows], new_index[np.asarray(colind)])), shape=(n, n))
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data


def bandwidth(rowptr, colind):
    rows = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    return int(np.abs(rows - np.asarray(colind)).max())
