"""Host mirror of OrderCRCuthillMcKee (fvcr/cr_reorder.h:352-393, fvcr/cr_reorder.cpp:510-600). CPU only."""
import numpy as np
import pytest

from plugin_navierstokes_b200 import cr_reorder, meshgen


def test_two_triangles_by_hand():
    """two triangles sharing side 2. Degrees: the dofs of the four outer sides have 7 connections, the two of the shared side 12,
    both pressures 4 * 7 + 2 * 12 = 52 -> start at the first one (strict <): its velocities in list order (0..3, then the shared
    side 4, 5), itself (6); then the neighbour pressure: its remaining velocities (7..10), itself (11)."""
    es = np.array([[0, 1, 2], [2, 3, 4]], dtype=np.int32)
    conn, minpind = cr_reorder.cr_get_connections(es, 5, 2)
    assert minpind == 10 and len(conn) == 12
    assert sorted(conn[0]) == [0, 1, 2, 3, 4, 5, 10] and len(conn[4]) == 12 and sorted(conn[11]) == [4, 5, 6, 7, 8, 9, 11]
    new = cr_reorder.ComputeCRCuthillMcKeeOrder(conn, minpind, False)
    assert list(new) == [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 6, 11]
    assert list(cr_reorder.OrderCRCuthillMcKee(es, 5)) == list(new)


@pytest.mark.parametrize("elem,n", [("tri", 6), ("quad", 6)])
def test_order_is_a_permutation_with_the_cuthill_mckee_structure(ora, elem, n):
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=3)
    es, n_side = meshgen.element_sides(elem, conn)
    new = cr_reorder.OrderCRCuthillMcKee(es, n_side)
    ndof = n_side * 2 + conn.shape[0]
    assert sorted(new) == list(range(ndof))
    # a pressure is numbered behind every velocity dof of its element (velocities have the lower degree, :577-578)
    for e in range(conn.shape[0]):
        vel = np.concatenate([es[e] * 2, es[e] * 2 + 1])
        assert new[n_side * 2 + e] > new[vel].max()
    # the two components of a side stay neighbours in the new numbering or are separated only by dofs of the same visit
    # the permuted FVCR pattern has a far smaller bandwidth than the layout "all velocities, then all pressures"
    rowptr, colind = ora.fvcr_csr(ora.ELEM[elem], es, n_side)
    vals = np.ones(colind.size)
    rp2, ci2, v2 = cr_reorder.permute_csr(rowptr, colind, vals, new)
    assert rp2[-1] == rowptr[-1] and v2.sum() == vals.sum()
    bw0, bw1 = cr_reorder.bandwidth(rowptr, colind), cr_reorder.bandwidth(rp2, ci2)
    assert bw1 < 0.35 * bw0, (bw0, bw1)
    # P A P^T applied to the permuted vector = permuted (A x)
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, colind.size); x = rng.uniform(-1, 1, ndof)
    A = sp.csr_matrix((a, colind, rowptr), shape=(ndof, ndof))
    rp3, ci3, a3 = cr_reorder.permute_csr(rowptr, colind, a, new)
    B = sp.csr_matrix((a3, ci3, rp3), shape=(ndof, ndof))
    assert np.allclose(B @ cr_reorder.permute_vector(x, new), cr_reorder.permute_vector(A @ x, new), atol=1e-13)


def test_errors():
    with pytest.raises(ValueError, match="two space dimensions"):
        cr_reorder.OrderCRCuthillMcKee(np.zeros((1, 4), np.int32), 4, dim=3)
    # two elements without a common side: the pressure graph is not connected
    with pytest.raises(ValueError, match="not connected"):
        cr_reorder.OrderCRCuthillMcKee(np.array([[0, 1, 2], [3, 4, 5]], np.int32), 6)
