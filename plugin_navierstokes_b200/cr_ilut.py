"""Host mirror of the reference's `CRILUT` preconditioner (fvcr/cr_ilut.h:87-497, registered in fvcr/register_fvcr.cpp:225-245):
threshold ILU for Crouzeix-Raviart saddle-point systems with one drop threshold per block type -- a fill-in entry of row i in
column k is kept if |value| > dmax(row i) * eps_{type(i) type(k)}, type = velocity when A(i, i) != 0, pressure when the diagonal
entry is zero (:236, :305).

Solver-side reference implementation on the host (scalar algebra, Python loops over CSR rows): it documents the algorithm next to
the assembly path and serves the tests of `cr_reorder` (the ordering decides whether a pressure pivot exists); it is not a device
preconditioner and no timed path uses it. Quirks of the reference that are kept: dmax is the row maximum of A before the
elimination (:241-251); the fill thresholds compare against min(eps) first (:302, :353); the last row of U may carry a
(near-)zero pivot of the singular pressure mode and is then skipped (:431-448, m_small_lower = 1e-9, m_small_upper = 1e-6)."""
import numpy as np

_SMALL_LOWER, _SMALL_UPPER = 1e-9, 1e-6
_VEL, _PRE = 0, 1


class CRILUTPreconditioner:
    def __init__(self, *thresh, info=False):
        """CRILUT(eps) | CRILUT(threshvv, thresh_vp_pv_pp) | CRILUT(threshvv, threshvp, threshpv, threshpp), cr_ilut.h:109-142"""
        if len(thresh) == 0:
            thresh = (1e-6,)
        self.set_threshold(*thresh)
        self.info = bool(info)
        self.L = self.U = None
        self.warnings = []

    def set_threshold(self, *t):                                     # :165-191
        if len(t) == 1:
            vv = vp = pv = pp = float(t[0])
        elif len(t) == 2:
            vv, vp, pv, pp = float(t[0]), float(t[1]), float(t[1]), float(t[1])
        elif len(t) == 4:
            vv, vp, pv, pp = (float(x) for x in t)
        else:
            raise ValueError("set_threshold: 1, 2 or 4 thresholds")
        self.eps_vv, self.eps_vp, self.eps_pv, self.eps_pp = vv, vp, pv, pp
        self.eps = min(vv, vp, pv, pp)

    def set_info(self, b):
        self.info = bool(b)

    def _keep(self, itype, ktype, val, dmax):
        if not abs(val) > dmax * self.eps:
            return False
        e = (self.eps_vv, self.eps_vp, self.eps_pv, self.eps_pp)[2 * itype + ktype]
        return abs(val) > dmax * e

    def preprocess(self, rowptr, colind, values):
        """factorisation A ~ L U (unit lower L), rows sorted by column (:205-402). Returns True."""
        rowptr, colind, values = np.asarray(rowptr), np.asarray(colind), np.asarray(values, dtype=np.float64)
        n = rowptr.size - 1
        diag = np.zeros(n)
        for i in range(n):
            a, b = rowptr[i], rowptr[i + 1]
            if np.any(np.diff(colind[a:b]) <= 0):
                raise ValueError("CRILUT: the matrix rows must be sorted")
            hit = np.nonzero(colind[a:b] == i)[0]
            if hit.size:
                diag[i] = values[a + hit[0]]
        typ = np.where(diag != 0.0, _VEL, _PRE)
        L = [None] * n
        U = [None] * n
        L[0] = ([], [])
        U[0] = (list(map(int, colind[rowptr[0]:rowptr[1]])), list(map(float, values[rowptr[0]:rowptr[1]])))
        for i in range(1, n):
            a, b = rowptr[i], rowptr[i + 1]
            idx = list(map(int, colind[a:b])); val = list(map(float, values[a:b]))
            dmax = max((abs(v) for v in val), default=0.0)
            itype = int(typ[i])
            u_part = len(idx)
            q = 0
            while q < len(idx):
                k = idx[q]
                if k >= i:
                    u_part = q
                    break
                if val[q] == 0.0:
                    q += 1
                    continue
                uk_i, uk_v = U[k]
                if not uk_i or uk_i[0] != k:
                    raise ZeroDivisionError("CRILUT: row %d has no pivot" % k)
                d = val[q] = val[q] / uk_v[0]                          # L(i, k)
                j, t = q + 1, 1
                while t < len(uk_i) and j < len(idx):                  # merge of the sorted lists (:286-340)
                    if uk_i[t] == idx[j]:
                        val[j] -= uk_v[t] * d; t += 1; j += 1
                    elif uk_i[t] < idx[j]:
                        c = -uk_v[t] * d
                        if self._keep(itype, int(typ[uk_i[t]]), c, dmax):
                            idx.insert(j, uk_i[t]); val.insert(j, c); j += 1
                        t += 1
                    else:
                        j += 1
                while t < len(uk_i):                                   # behind the last connection of row i (:342-381)
                    c = -uk_v[t] * d
                    if self._keep(itype, int(typ[uk_i[t]]), c, dmax):
                        idx.append(uk_i[t]); val.append(c)
                    t += 1
                q += 1
            L[i] = (idx[:u_part], val[:u_part])
            U[i] = (idx[u_part:], val[u_part:])
        self.L, self.U, self.n = L, U, n
        self.nnz_A = int(rowptr[-1])
        self.nnz_LU = sum(len(r[0]) for r in L) + sum(len(r[0]) for r in U)
        if self.info:
            print("CRILUT storage information: A %d, L+U %d connections, increase factor %.3f" % (self.nnz_A, self.nnz_LU, self.nnz_LU / self.nnz_A))
        return True

    def step(self, d):
        """c = (L U)^-1 d (:405-467)"""
        d = np.asarray(d, dtype=np.float64)
        n = self.n
        c = np.empty(n)
        for i in range(n):
            li, lv = self.L[i]
            s = d[i]
            for k, v in zip(li, lv):
                s -= v * c[k]
            c[i] = s
        ui, uv = self.U[n - 1]
        if not ui or ui[0] != n - 1:
            raise ZeroDivisionError("CRILUT: the last row has no diagonal entry")
        if abs(uv[0]) < _SMALL_LOWER:                               # singular pressure mode in the last row
            if abs(c[n - 1]) > _SMALL_UPPER:
                self.warnings.append("zero entry in last row of U with corresponding non-zero rhs entry (%g)" % abs(c[n - 1]))
            c[n - 1] = 0.0
        else:
            c[n - 1] = c[n - 1] / uv[0]
        for i in range(n - 2, -1, -1):
            ui, uv = self.U[i]
            if not ui or ui[0] != i:
                raise ZeroDivisionError("CRILUT: row %d has no diagonal entry" % i)
            s = c[i]
            for k, v in zip(ui[1:], uv[1:]):
                s -= v * c[k]
            c[i] = s / uv[0]
        return c
