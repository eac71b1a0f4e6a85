// ns_fvcr_q.cuh -- FVCR on quadrilaterals and hexahedra (fvcr/navier_stokes_fvcr.cpp:792,811 register the disc for them):
// same element routine as ns_fvcr.cuh (add_jac_A_elem / add_def_A_elem / add_jac_M_elem / add_def_M_elem / add_rhs_elem,
// fvcr/navier_stokes_fvcr.cpp:244-759) on the non-affine CR geometry: rotated bi-/trilinear Crouzeix-Raviart shapes nodal at the
// side centres (span {1, x, y, x^2 - y^2} / {1, x, y, z, x^2 - y^2, y^2 - z^2}), shape gradients through JTInv of the element map
// AT each SCVF ip (they vary over the element), one SCV volume per side (triangle / pyramid between the side and the barycentre).
//
// Mapping: 3 quadrilaterals (L = 9) or 1 hexahedron (L = 19) per warp. Element-level geometry: lane = side. Per-ip phase:
// lane = SCVF. Column phase: lane = one column (side s, component d2 | pressure) of the local Jacobian in registers.
#pragma once
#include "ns_fvcr.cuh"

namespace nsb {

template <> struct CRT<E_QUAD> { static constexpr int DIM = 2, NCO = 4, NS = 4, NIP = 4; };
template <> struct CRT<E_HEX>  { static constexpr int DIM = 3, NCO = 8, NS = 6, NIP = 12; };

// SCVF -> (from side, to side): the two sides sharing the corner (2-D) / edge (3-D), lower index = from
template <int E> __host__ __device__ constexpr int crq_ft(int ip, int j)
{
    if (E == E_QUAD) { constexpr int T[4][2] = {{0, 3}, {0, 1}, {1, 2}, {2, 3}}; return T[ip][j]; }
    constexpr int T[12][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {1, 4}, {1, 2}, {2, 3}, {3, 4}, {1, 5}, {2, 5}, {3, 5}, {4, 5}};
    return T[ip][j];
}

template <int E> struct CRQWS {
    static constexpr int DIM = CRT<E>::DIM, NCO = CRT<E>::NCO, NS = CRT<E>::NS, NIP = CRT<E>::NIP, L = NS * DIM + 1;
    double x[NCO * DIM];
    double u[L];
    double scvn[NS][DIM], scvx[NS][DIM], vol[NS];
    double bary[DIM];
    int64_t rowbase[NS];
    int32_t rowlen[NS], scnt[NS], side[NS];
    CRRec<E> rec[NIP];
};

// rotated bi-/trilinear Crouzeix-Raviart shapes and local gradients: N_s = sum_j C[s][j] b_j, b = {1, x, y, (z,) x^2-y^2 (, y^2-z^2)};
// C = inverse of the generalised Vandermonde matrix at the side centres (exact rationals)
template <int E> NSB_DEV void crq_shapes(const double* xi, double* N, double (*dN)[CRT<E>::DIM])
{
    if constexpr (E == E_QUAD) {
        const double x = xi[0], y = xi[1], q = x * x - y * y;
        constexpr double C[4][4] = {{0.75, 1, -2, -1}, {-0.25, 0, 1, 1}, {-0.25, 1, 0, -1}, {0.75, -2, 1, 1}};
#pragma unroll
        for (int s = 0; s < 4; s++) {
            N[s] = C[s][0] + C[s][1] * x + C[s][2] * y + C[s][3] * q;
            if (dN) { dN[s][0] = C[s][1] + C[s][3] * 2 * x; dN[s][1] = C[s][2] - C[s][3] * 2 * y; }
        }
    } else {
        const double x = xi[0], y = xi[1], z = xi[2], q1 = x * x - y * y, q2 = y * y - z * z;
        constexpr double T = 1.0 / 3.0;
        constexpr double C[6][6] = {{2 * T, 2 * T, 2 * T, -7 * T, -2 * T, -4 * T}, {2 * T, 2 * T, -7 * T, 2 * T, -2 * T, 2 * T},
                                    {-T, -T, 2 * T, 2 * T, 4 * T, 2 * T},          {-T, 2 * T, -T, 2 * T, -2 * T, 2 * T},
                                    {2 * T, -7 * T, 2 * T, 2 * T, 4 * T, 2 * T},   {-T, 2 * T, 2 * T, -T, -2 * T, -4 * T}};
#pragma unroll
        for (int s = 0; s < 6; s++) {
            N[s] = C[s][0] + C[s][1] * x + C[s][2] * y + C[s][3] * z + C[s][4] * q1 + C[s][5] * q2;
            if (dN) { dN[s][0] = C[s][1] + C[s][4] * 2 * x; dN[s][1] = C[s][2] - C[s][4] * 2 * y + C[s][5] * 2 * y; dN[s][2] = C[s][3] - C[s][5] * 2 * z; }
        }
    }
}

// CR upwind shapes of one ip: upwind.cpp:82-104 (No), :174-213 (Full), :432-499 (Skewed), :577-636 (LPS)
template <int E> NSB_DEV bool crq_upwind_ip(int type, const double* x, const double* n, const double* xip, const double* N,
                                            int from, int to, const double* vel, double* up)
{
    constexpr int DIM = CRT<E>::DIM, NS = CRT<E>::NS;
    if (type == UPW_NO) {
#pragma unroll
        for (int s = 0; s < NS; s++) up[s] = N[s];
        return true;
    }
#pragma unroll
    for (int s = 0; s < NS; s++) up[s] = 0.0;
    if (type == UPW_FULL) {
        const double flux = dotv<DIM>(n, vel);
        const int sd = flux > 0.0 ? from : to;
#pragma unroll
        for (int s = 0; s < NS; s++) up[s] = (s == sd) ? 1.0 : 0.0;
        return true;
    }
    const double nrm = sqrt(dotv<DIM>(vel, vel));
    if (type == UPW_SKEWED ? (nrm < 1e-14) : (nrm == 0.0)) return true;
    int side = 0; double gc[DIM], lc[DIM], Nc[NS];
    if (!side_ray_cut<E>(x, xip, vel, side, gc, lc)) return false;
    crq_shapes<E>(lc, Nc, (double (*)[DIM]) nullptr);
    if (type == UPW_SKEWED) {
        double mx = -1000.0; int best = 0;
#pragma unroll
        for (int s = 0; s < NS; s++) if (Nc[s] > mx) { mx = Nc[s]; best = s; }
#pragma unroll
        for (int s = 0; s < NS; s++) up[s] = (s == best) ? 1.0 : 0.0;
    } else {
#pragma unroll
        for (int s = 0; s < NS; s++) up[s] = Nc[s];
    }
    return true;
}

template <int E, int SC, int MINB = 4>
__global__ void __launch_bounds__(128, MINB) fvcrq_elem_kernel(KParams p, FvcrDev m, const int32_t* __restrict__ elem_list,
                                                        int64_t n_list, const double* __restrict__ u,
                                                        double* __restrict__ val, double* __restrict__ def,
                                                        int* __restrict__ errflag)
{
    constexpr int DIM = CRT<E>::DIM, NCO = CRT<E>::NCO, NS = CRT<E>::NS, NIP = CRT<E>::NIP, L = NS * DIM + 1, PI = NS * DIM;
    constexpr int EPW = 32 / L;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CRQWS<E>* wsall = reinterpret_cast<CRQWS<E>*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / L, col = lane - sub * L;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t li = gw * EPW + sub;
    const bool active = sub < EPW && li < n_list;
    CRQWS<E>& ws = wsall[warp * EPW + (sub < EPW ? sub : 0)];
    const int64_t e = active ? (elem_list ? (int64_t)elem_list[li] : li) : 0;
    const int64_t pbase = m.n_side * DIM;
    const double nurho = p.visc * p.rho;

    // ---- stage ----
    if (active) {
        if (col < NS) {
            const int sd = m.esides[e * NS + col];
            ws.side[col] = sd;
            const int nadj = (int)(m.sadj_ptr[sd + 1] - m.sadj_ptr[sd]);
            ws.rowbase[col] = m.srow[sd]; ws.scnt[col] = m.scnt[sd]; ws.rowlen[col] = m.scnt[sd] * DIM + nadj;
        }
        for (int i = col; i < NCO * DIM; i += L) {
            const int kk = i / DIM, dd = i - kk * DIM;
            ws.x[i] = m.coords[(int64_t)m.conn[e * NCO + kk] * DIM + dd];
        }
    }
    __syncwarp();
    if (active) {
        // local dofs: (side s, comp d) at s*DIM+d, pressure last
        if (col < PI) { const int s = col / DIM, d = col - s * DIM; ws.u[col] = u[(int64_t)ws.side[s] * DIM + d]; }
        else ws.u[PI] = u[pbase + e];
        // ---- CRFVGeometry (App. B-3), element-level part: lane s < NS owns the SCV of side s ----
        if (col < NS) {
            const int s = col;
            constexpr int NSC = DIM == 2 ? 2 : 4;
            double bary[DIM], xb[DIM], nn[DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) { double t = 0; for (int k = 0; k < NCO; k++) t += ws.x[k * DIM + d]; bary[d] = t / NCO; xb[d] = 0.0; }
            for (int q = 0; q < NSC; q++) for (int d = 0; d < DIM; d++) xb[d] += ws.x[t_side<E>(s, q) * DIM + d];
#pragma unroll
            for (int d = 0; d < DIM; d++) xb[d] /= NSC;
            double vol;
            if constexpr (DIM == 2) {
                const double* a = ws.x + t_side<E>(s, 0) * 2; const double* b = ws.x + t_side<E>(s, 1) * 2;
                nn[0] = b[1] - a[1]; nn[1] = -(b[0] - a[0]);
                vol = 0.5 * fabs((b[0] - a[0]) * (bary[1] - a[1]) - (b[1] - a[1]) * (bary[0] - a[0]));          // triangle (side, barycentre)
            } else {
                // quadrilateral side: area vector 0.5 (c2-c0) x (c3-c1); SCV = pyramid (side, barycentre) as the two tetrahedra
                // of the side split along its diagonal 0-2
                const double* c0 = ws.x + t_side<E>(s, 0) * 3; const double* c1 = ws.x + t_side<E>(s, 1) * 3;
                const double* c2 = ws.x + t_side<E>(s, 2) * 3; const double* c3 = ws.x + t_side<E>(s, 3) * 3;
                double d1[3], d2[3], a1[3], a2[3], a3[3], a4[3], t[3];
                for (int d = 0; d < 3; d++) { d1[d] = c2[d] - c0[d]; d2[d] = c3[d] - c1[d]; a1[d] = c1[d] - c0[d]; a2[d] = c2[d] - c0[d]; a3[d] = bary[d] - c0[d]; a4[d] = c3[d] - c0[d]; }
                cross3(t, d1, d2);
                for (int d = 0; d < 3; d++) nn[d] = 0.5 * t[d];
                cross3(t, a1, a2); vol = fabs(dotv<3>(t, a3));
                cross3(t, a2, a4); vol += fabs(dotv<3>(t, a3));
                vol = vol / 6.0;
            }
            double outw = 0;
#pragma unroll
            for (int d = 0; d < DIM; d++) outw += nn[d] * (xb[d] - bary[d]);
            const double sg = outw < 0 ? -1.0 : 1.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) { ws.scvn[s][d] = sg * nn[d]; ws.scvx[s][d] = xb[d]; }
            ws.vol[s] = vol;
            if (s == 0) {
#pragma unroll
                for (int d = 0; d < DIM; d++) ws.bary[d] = bary[d];
            }
        }
    }
    __syncwarp();
    // ---- per-ip phase: lane = SCVF ----
    if (active && col < NIP && (p.what & (W_JAC_A | W_DEF_A))) {
        const int ip = col;
        const int from = crq_ft<E>(ip, 0), to = crq_ft<E>(ip, 1);
        CRRec<E>& r = ws.rec[ip];
        // SCVF spanned by the (dim-2)-object and the barycentre; normal oriented from -> to
        double n[DIM], xip[DIM], lip[DIM], N[NS], G[NS][DIM];
        {
            double lb[DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) lb[d] = 0.5;
            if constexpr (DIM == 2) {
                const double* a = ws.x + ip * 2;
                n[0] = ws.bary[1] - a[1]; n[1] = -(ws.bary[0] - a[0]);
                for (int d = 0; d < 2; d++) { xip[d] = 0.5 * (a[d] + ws.bary[d]); lip[d] = 0.5 * (t_corner<E>(ip, d) + lb[d]); }
            } else {
                const int c0 = t_edge<E>(ip, 0), c1 = t_edge<E>(ip, 1);
                double e1[3], e2[3], c[3];
                for (int d = 0; d < 3; d++) { e1[d] = ws.x[c1 * 3 + d] - ws.x[c0 * 3 + d]; e2[d] = ws.bary[d] - ws.x[c0 * 3 + d]; }
                cross3(c, e1, e2);
                for (int d = 0; d < 3; d++) {
                    n[d] = 0.5 * c[d];
                    xip[d] = (ws.x[c0 * 3 + d] + ws.x[c1 * 3 + d] + ws.bary[d]) / 3.0;
                    lip[d] = (t_corner<E>(c0, d) + t_corner<E>(c1, d) + lb[d]) / 3.0;
                }
            }
            double ft = 0;
#pragma unroll
            for (int d = 0; d < DIM; d++) ft += n[d] * (ws.scvx[to][d] - ws.scvx[from][d]);
            if (ft < 0) {
#pragma unroll
                for (int d = 0; d < DIM; d++) n[d] = -n[d];
            }
            // shapes and global gradients at the local ip: JTInv of the bi-/trilinear element map at that point
            double dNc[NS][DIM], dNl[NCO][DIM], JT[DIM][DIM], JI[DIM][DIM];
            crq_shapes<E>(lip, N, dNc);
            lagrange_grad<E>(lip, dNl);
#pragma unroll
            for (int i = 0; i < DIM; i++)
#pragma unroll
                for (int j = 0; j < DIM; j++) { double s = 0; for (int k = 0; k < NCO; k++) s += dNl[k][i] * ws.x[k * DIM + j]; JT[i][j] = s; }
            inv_mat<DIM>(JT, JI);
#pragma unroll
            for (int k = 0; k < NS; k++)
#pragma unroll
                for (int j = 0; j < DIM; j++) { double s = 0; for (int i = 0; i < DIM; i++) s += JI[j][i] * dNc[k][i]; G[k][j] = s; }
        }
        double std[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) { double s = 0; for (int k = 0; k < NS; k++) s += ws.u[k * DIM + d] * N[k]; std[d] = s; }
        const double prod = dotv<DIM>(std, n) * p.rho;
        // Peclet weight (fvcr/navier_stokes_fvcr.cpp:244-265)
        double w = 1.0;
        if (!p.stokes && p.peclet) {
            double dd = 0;
#pragma unroll
            for (int d = 0; d < DIM; d++) { const double t = ws.scvx[to][d] - ws.scvx[from][d]; dd += t * t; }
            const double Pe = dotv<DIM>(std, n) / dotv<DIM>(n, n) * sqrt(dd) / p.visc;
            const double Pe2 = Pe * Pe;
            w = Pe2 / (5.0 + Pe2);
        }
        double up[NS], U[DIM];
        bool ok = true;
        if (!p.stokes) {
            ok = crq_upwind_ip<E>(p.upw_conv, ws.x, n, xip, N, from, to, std, up);
            if (!ok) atomicExch(errflag, 1);
#pragma unroll
            for (int d = 0; d < DIM; d++) { double s = 0; for (int k = 0; k < NS; k++) s += up[k] * ws.u[k * DIM + d]; U[d] = s; }
            if (p.peclet) {
#pragma unroll
                for (int d = 0; d < DIM; d++) U[d] = w * U[d] + (1.0 - w) * std[d];
            }
        } else {
#pragma unroll
            for (int k = 0; k < NS; k++) up[k] = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) U[d] = 0.0;
        }
        if (p.what & W_JAC_A) {                                   // :299-476
#pragma unroll
            for (int d = 0; d < DIM; d++) r.n[d] = n[d];
#pragma unroll
            for (int k = 0; k < NS; k++) {
                double D = -1.0 * nurho * dotv<DIM>(G[k], n);
                if (!p.stokes) { D += up[k] * prod * w; if (p.peclet) D += prod * (1.0 - w) * N[k]; }
                r.D[k] = D;
#pragma unroll
                for (int d1 = 0; d1 < DIM; d1++) {
                    double A = p.laplace ? 0.0 : -1.0 * nurho * G[k][d1];
                    if (!p.stokes && p.exact_jac != 0.0) {
                        A += p.exact_jac * p.rho * std[d1] * N[k];                          // :425-426 (StdVel, not the upwind velocity)
                        if (p.peclet) A += U[d1] * (1.0 - w) * N[k] * p.rho * p.exact_jac;  // :451-454
                    }
                    r.A[k][d1] = A;
                    r.gd[k][d1] = p.grad_div > 0 ? p.grad_div * G[k][d1] : 0.0;          // :341-348
                }
            }
        }
        if (p.what & W_DEF_A) {                                   // :538-641
            double gv[DIM][DIM];
#pragma unroll
            for (int d1 = 0; d1 < DIM; d1++)
#pragma unroll
                for (int d2 = 0; d2 < DIM; d2++) { double s = 0; for (int k = 0; k < NS; k++) s += G[k][d2] * ws.u[k * DIM + d1]; gv[d1][d2] = s; }
            double divu = 0;
#pragma unroll
            for (int d = 0; d < DIM; d++) divu += gv[d][d];
#pragma unroll
            for (int d1 = 0; d1 < DIM; d1++) {
                double df = 0;
#pragma unroll
                for (int d2 = 0; d2 < DIM; d2++) df += gv[d1][d2] * n[d2];
                if (!p.laplace) {
#pragma unroll
                    for (int d2 = 0; d2 < DIM; d2++) df += gv[d2][d1] * n[d2];
                }
                double f = df * (-1.0) * nurho;
                if (p.grad_div > 0) f -= p.grad_div * divu * n[d1];                          // :588-596
                if (!p.stokes) f += (p.defect_upwind ? U[d1] : std[d1]) * prod;              // :602-630
                f += ws.u[PI] * n[d1];
                r.F[d1] = f;
            }
        }
    }
    __syncwarp();
    if (!active) return;
    // ---- column phase ----
    if (p.what & (W_JAC_A | W_JAC_M)) {
        double acc[L];
#pragma unroll
        for (int i = 0; i < L; i++) acc[i] = 0.0;
        const int s = col < PI ? col / DIM : 0, d2 = col < PI ? col - s * DIM : 0;
        if (p.what & W_JAC_A) {
            static_for<NIP>([&](auto ipc) {
                constexpr int ip = decltype(ipc)::value;
                constexpr int f = crq_ft<E>(ip, 0), t = crq_ft<E>(ip, 1);
                const CRRec<E>& r = ws.rec[ip];
#pragma unroll
                for (int d1 = 0; d1 < DIM; d1++) {
                    double v;
                    if (col < PI) { v = r.A[s][d1] * r.n[d2] - r.gd[s][d2] * r.n[d1]; if (d1 == d2) v += r.D[s]; }
                    else v = r.n[d1];                                                  // :470-475
                    acc[f * DIM + d1] += v; acc[t * DIM + d1] -= v;
                }
            });
            if (col < PI) acc[PI] = ws.scvn[s][d2];                                     // :483-489
#pragma unroll
            for (int i = 0; i < L; i++) acc[i] *= p.scale_a;
        }
        if ((p.what & W_JAC_M) && col < PI) {                                           // :672-699
#pragma unroll
            for (int i = 0; i < PI; i++) if (i == col) acc[i] += p.scale_m * ws.vol[s] * p.rho;
        }
        // scatter: velocity rows (side a, d1), then the pressure row of the element
        const int nsl = col < PI ? s : 0;
#pragma unroll
        for (int a = 0; a < NS; a++) {
            int64_t off;
            if (col < PI) off = (int64_t)m.emap[e * (NS * NS) + a * NS + nsl] * DIM + d2;
            else off = (int64_t)ws.scnt[a] * DIM + m.pslot[e * NS + a];
#pragma unroll
            for (int d1 = 0; d1 < DIM; d1++) {
                double* q = val + ws.rowbase[a] + (int64_t)d1 * ws.rowlen[a] + off;
                atomicAdd(q, acc[a * DIM + d1]);
            }
        }
        {
            const int64_t off = col < PI ? (int64_t)m.psort[e * NS + nsl] * DIM + d2 : (int64_t)NS * DIM;
            double* q = val + m.prow0 + e * (int64_t)L + off;
            atomicAdd(q, acc[PI]);
        }
    }
    if (p.what & (W_DEF_A | W_DEF_M | W_RHS)) {
        double d = 0.0;
        if (col < PI) {
            const int s = col / DIM, d1 = col - s * DIM;
            if (p.what & W_DEF_A) {
#pragma unroll
                for (int ip = 0; ip < NIP; ip++) {
                    if (crq_ft<E>(ip, 0) == s) d += ws.rec[ip].F[d1];
                    if (crq_ft<E>(ip, 1) == s) d -= ws.rec[ip].F[d1];
                }
            }
            if ((p.what & W_RHS) && p.has_source) d -= p.src[d1] * ws.vol[s];               // no density factor, :757
            d *= p.scale_a;
            if (p.what & W_DEF_M) d += p.scale_m * ws.u[col] * ws.vol[s] * p.rho;           // :702-729
            double* q = def + (int64_t)ws.side[s] * DIM + d1;
            atomicAdd(q, d);
        } else {
            if (p.what & W_DEF_A) {                                                      // :648-654
                for (int sd = 0; sd < NS; sd++)
#pragma unroll
                    for (int d1 = 0; d1 < DIM; d1++) d += ws.scvn[sd][d1] * ws.u[sd * DIM + d1];
            }
            d *= p.scale_a;
            double* q = def + pbase + e;
            atomicAdd(q, d);
        }
    }
}

}  // namespace nsb
