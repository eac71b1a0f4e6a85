// ns_dense.cuh -- FV1 element kernel for upwinds WITH ip-shapes (NavierStokesPositiveUpwind):
// the Schneider-Raw closure couples all ips of an element, so one warp assembles and solves the
// nIp x nIp system cooperatively in shared memory (reference: stabilization.cpp:244-403 FIELDS,
// :590-771 FLOW; upwind.cpp:643-786 PositiveUpwind).
//
// Phases of one warp (= one element):
//   G  lane = ip      geometry, StdVel                                  -> shared
//   U  lane = ip / corner / ip   Positive upwind (up and, for FLOW, down) or per-ip upwinds
//   S  lane = ip      a, b, c, matrix rows ; all lanes: LU with partial pivoting ; lane = rhs id: solves
//   R  lane = ip      transported velocity, Peclet blend, defect fluxes  -> IpRec
//   C  lane = (k,cf)  Jacobian column in registers (jac_col with the dense shape accessor), scatter
#pragma once
#include "ns_kernels.cuh"

namespace nsb {

template <int E, bool PAC> struct DenseWS {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP, NF = DIM + 1;
    double x[NSH * DIM], u[NSH * NF], s0[NSH * NF], s1[NSH * NF], vol[NSH];
    int64_t rowbase[NSH];
    int32_t cnt[NSH], node[NSH];
    // geometry
    double n[NIP][DIM], xip[NIP][DIM], N[NIP][NSH], G[NIP][NSH][DIM], ds[NIP], nn[NIP];
    double std[NIP][DIM];
    double sv[NIP][DIM][DIM][NSH], sp[NIP][DIM][NSH], svel[NIP][DIM];
    // The upwind shapes and the ip system are dead once phase R has read its upwind rows: the per-ip records of
    // phases R / C live in the same storage (7.6 KB of the 30 KB workspace for hex -> 10 instead of 7 warps per SM).
    struct Closure {
        // upwinds: [0] = upwind of the stabilisation, [1] = its downwind (FLOW), [2] = convective upwind
        double ush[3][NIP][NSH], uip[3][NIP][NIP], ulen[3][NIP];
        double flux[NIP];
        // ip system
        double a[NIP], b[NIP], c[NIP];
        double M[NIP][NIP];
        double dinv[NIP];          // reciprocals of the LU diagonal (one division per row instead of one per row and right-hand side)
        int32_t has[NIP];
        int32_t perm[NIP];
    };
    union {
        Closure cl;
        IpRec<E, PAC> rec[NIP];
    };
};

// NavierStokesPositiveUpwind::compute (upwind.cpp:643-786) for the ip velocities sgn*std; all 32 lanes call.
template <int E, class WS> NSB_DEV void positive_upwind(WS& ws, int lane, double sgn, int slot)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP, NINC = ET<E>::NINC;
    const double eps = 2.220446049250313e-16 * 10;
    if (lane < NIP) {
        const int ip = lane;
        for (int k = 0; k < NSH; k++) ws.cl.ush[slot][ip][k] = 0.0;
        for (int j = 0; j < NIP; j++) ws.cl.uip[slot][ip][j] = 0.0;
        double v[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) v[d] = sgn * ws.std[ip][d];
        const double normsq = dotv<DIM>(v, v);
        const int f = t_edge<E>(ip, 0), t = t_edge<E>(ip, 1);
        int has = 1; double fl = 0.0;
        if (fabs(normsq) <= eps) has = 0;
        else {
            fl = dotv<DIM>(v, ws.n[ip]);
            const double vel = sqrt(normsq), len = sqrt(ws.nn[ip]);
            if (fabs(fl / sqrt(vel * len)) <= eps) has = 0;
        }
        if (!has) { ws.cl.ush[slot][ip][f] = 0.5; ws.cl.ush[slot][ip][t] = 0.5; }
        ws.cl.flux[ip] = fl; ws.cl.has[ip] = has;
    }
    __syncwarp();
    const unsigned any = __ballot_sync(0xffffffffu, lane < NIP && ws.cl.has[lane < NIP ? lane : 0]);
    if (any != 0u && lane < NSH) {
        const int sh = lane;
        int ips[NINC]; double fl[NINC]; int cnt = 0;
        double m_in = 0.0, m_out = 0.0;
#pragma unroll
        for (int q = 0; q < NINC; q++) {
            const int ip = t_inc<E>(sh, q);
            if (!ws.cl.has[ip]) continue;
            const double f = (double)t_inc_sign<E>(sh, q) * ws.cl.flux[ip];
            ips[cnt] = ip; fl[cnt] = f; cnt++;
            m_in += -1.0 * fmin(f, 0.0); m_out += fmax(f, 0.0);
        }
        const double F = fmax(m_in, m_out);
        for (int i = 0; i < cnt; i++) if (fl[i] > 0) {
            double sum = 0.0;
            for (int j = 0; j < cnt; j++) if (fl[j] < 0) { const double s = -1.0 * fl[j] / F; ws.cl.uip[slot][ips[i]][ips[j]] = s; sum += s; }
            ws.cl.ush[slot][ips[i]][sh] = 1.0 - sum;
        }
    }
    __syncwarp();
    if (lane < NIP) {
        const int ip = lane;
        double up[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) up[d] = 0.0;
        for (int k = 0; k < NSH; k++)
#pragma unroll
            for (int d = 0; d < DIM; d++) up[d] += ws.cl.ush[slot][ip][k] * ws.x[k * DIM + d];
        for (int j = 0; j < NIP; j++)
#pragma unroll
            for (int d = 0; d < DIM; d++) up[d] += ws.cl.uip[slot][ip][j] * ws.xip[j][d];
        ws.cl.ulen[slot][ip] = sqrt(dist2<DIM>(ws.xip[ip], up));
    }
    __syncwarp();
}

// per-ip upwinds (No/Full/Skewed/LPS) into the same shared layout; ip shapes are zero
template <int E, class WS> NSB_DEV bool simple_upwind(WS& ws, int lane, int type, double sgn, int slot)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP;
    bool ok = true;
    if (lane < NIP) {
        const int ip = lane;
        IpGeo<E> g;
        g.from = t_edge<E>(ip, 0); g.to = t_edge<E>(ip, 1); g.ds = ws.ds[ip];
#pragma unroll
        for (int d = 0; d < DIM; d++) { g.n[d] = ws.n[ip][d]; g.xip[d] = ws.xip[ip][d]; }
#pragma unroll
        for (int k = 0; k < NSH; k++) g.N[k] = ws.N[ip][k];
        double v[DIM], up[NSH], len;
#pragma unroll
        for (int d = 0; d < DIM; d++) v[d] = sgn * ws.std[ip][d];
        ok = upwind_ip<E>(type, ws.x, g, v, up, len);
#pragma unroll
        for (int k = 0; k < NSH; k++) ws.cl.ush[slot][ip][k] = up[k];
        for (int j = 0; j < NIP; j++) ws.cl.uip[slot][ip][j] = 0.0;
        ws.cl.ulen[slot][ip] = len;
    }
    __syncwarp();
    return ok;
}

template <int E, int SC, bool PAC>
__global__ void __launch_bounds__(128) fv1_dense_kernel(KParams p, MeshDev m, const int32_t* __restrict__ elem_list,
                                                        int64_t n_list, const double* __restrict__ u,
                                                        const double* __restrict__ s0, const double* __restrict__ s1,
                                                        double* __restrict__ val, double* __restrict__ def,
                                                        double* __restrict__ Jloc, double* __restrict__ dloc,
                                                        int* __restrict__ errflag)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP, NF = DIM + 1, L = NSH * NF, P = DIM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DenseWS<E, PAC>& ws = reinterpret_cast<DenseWS<E, PAC>*>(smem_raw)[warp];
    const int64_t li = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (li >= n_list) return;                               // whole warp leaves together
    const int64_t e = elem_list ? (int64_t)elem_list[li] : li;

    // ---- stage ----
    if (lane < NSH) {
        const int nd = m.conn[e * NSH + lane];
        ws.node[lane] = nd; ws.vol[lane] = m.scvvol[e * NSH + lane];
        const int64_t b0 = m.brow[nd], b1 = m.brow[nd + 1];
        ws.rowbase[lane] = b0 * (NF * NF); ws.cnt[lane] = (int32_t)(b1 - b0);
    }
    __syncwarp();
    for (int i = lane; i < NSH * NF; i += 32) {
        const int kk = i / NF, ff = i - kk * NF;
        const int64_t gi = (int64_t)ws.node[kk] * NF + ff;
        ws.u[i] = u[gi];
        ws.s0[i] = p.time_dep ? s0[gi] : ws.u[i];
        ws.s1[i] = p.time_dep ? s1[gi] : 0.0;
    }
    for (int i = lane; i < NSH * DIM; i += 32) {
        const int kk = i / DIM, dd = i - kk * DIM;
        ws.x[i] = m.coords[(int64_t)ws.node[kk] * DIM + dd];
    }
    __syncwarp();
    // ---- G: geometry + StdVel ----
    if (lane < NIP) {
        const int ip = lane;
        IpGeo<E> g; ip_geometry<E>(ws.x, ip, g);
        ws.ds[ip] = g.ds; ws.nn[ip] = dotv<DIM>(g.n, g.n);
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            ws.n[ip][d] = g.n[d]; ws.xip[ip][d] = g.xip[d];
            double s = 0;
#pragma unroll
            for (int k = 0; k < NSH; k++) s += ws.u[k * NF + d] * g.N[k];
            ws.std[ip][d] = s;
        }
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            ws.N[ip][k] = g.N[k];
#pragma unroll
            for (int d = 0; d < DIM; d++) ws.G[ip][k][d] = g.G[k][d];
        }
    }
    __syncwarp();
    // ---- U: upwinds ----
    bool ok = true;
    const bool stab_pos = (p.upw_stab == UPW_POSITIVE);
    const bool conv_pos = (!p.pac && p.upw_conv == UPW_POSITIVE);
    if (!p.stokes) {
        if (stab_pos) positive_upwind<E>(ws, lane, 1.0, 0); else ok &= simple_upwind<E>(ws, lane, p.upw_stab, 1.0, 0);
        if (p.stab == STAB_FLOW) { if (stab_pos) positive_upwind<E>(ws, lane, -1.0, 1); else ok &= simple_upwind<E>(ws, lane, p.upw_stab, -1.0, 1); }
        if (!p.pac && p.upw_conv != p.upw_stab) { if (conv_pos) positive_upwind<E>(ws, lane, 1.0, 2); else ok &= simple_upwind<E>(ws, lane, p.upw_conv, 1.0, 2); }
    }
    if (!ok) atomicExch(errflag, 1);
    const int cslot = (p.upw_conv != p.upw_stab) ? 2 : 0;    // where the convective upwind lives
    // ---- S: Schneider-Raw closure ----
    if (p.stab == STAB_NONE) {                               // stabilization.cpp:805-850
        if (lane < NIP) {
            const int ip = lane;
            for (int d = 0; d < DIM; d++) {
                ws.svel[ip][d] = ws.std[ip][d];
                for (int k = 0; k < NSH; k++) {
                    ws.sp[ip][d][k] = 0.0;
                    for (int d2 = 0; d2 < DIM; d2++) ws.sv[ip][d][d2][k] = (d == d2) ? ws.N[ip][k] : 0.0;
                }
            }
        }
    } else {
        double cmn = 0, cav = 0, cmd = 0;
        if (p.diff_len == DIFF_COR) cor_stats<E>(ws.nn, ws.ds, cmn, cav, cmd);
        if (lane < NIP) {
            const int ip = lane;
            const int f = t_edge<E>(ip, 0), t = t_edge<E>(ip, 1);
            ws.cl.a[ip] = p.visc * diff_len_sq_inv<DIM>(p.diff_len, ws.nn[ip], ws.vol[f], ws.vol[t], ws.ds[ip], cmn, cav, cmd);
            double b = 0.0, c = 0.0;
            if (!p.stokes) {
                const double nrm = sqrt(dotv<DIM>(ws.std[ip], ws.std[ip]));
                b = nrm / ws.cl.ulen[0][ip];
                if (p.stab == STAB_FLOW) c = nrm / (ws.cl.ulen[1][ip] + ws.cl.ulen[0][ip]);
            }
            ws.cl.b[ip] = b; ws.cl.c[ip] = c;
        }
        __syncwarp();
        const bool dense = !p.stokes && stab_pos;
        const bool flow = (p.stab == STAB_FLOW);
        if (!dense) {
            // diagonal branch written into the dense layout (stabilization.cpp:166-241 / :489-587)
            if (lane < NIP) {
                const int ip = lane;
                double diag = ws.cl.a[ip];
                if (p.time_dep) diag += 1.0 / p.dt;
                if (!p.stokes) diag += ws.cl.b[ip];
                for (int d = 0; d < DIM; d++) {
                    double rhs = p.has_source ? p.src[d] : 0.0;
                    if (p.time_dep) { double o = 0.0; for (int k = 0; k < NSH; k++) o += ws.N[ip][k] * ws.s1[k * NF + d]; rhs += o / p.dt; }
                    for (int k = 0; k < NSH; k++) {
                        double sumVel = ws.cl.a[ip] * ws.N[ip][k];
                        if (!p.stokes) {
                            sumVel += ws.cl.b[ip] * ws.cl.ush[0][ip][k];
                            if (flow) sumVel += ws.cl.c[ip] * (ws.cl.ush[1][ip][k] - ws.cl.ush[0][ip][k]);
                        }
                        if (flow) for (int d2 = 0; d2 < DIM; d2++) if (d2 != d) sumVel -= ws.std[ip][d2] * ws.G[ip][k][d2];
                        rhs += sumVel * ws.s0[k * NF + d];
                        ws.sv[ip][d][d][k] = sumVel / diag;
                        for (int d2 = 0; d2 < DIM; d2++) if (d2 != d) {
                            if (flow) { const double s2 = ws.std[ip][d] * ws.G[ip][k][d2]; rhs += s2 * ws.s0[k * NF + d2]; ws.sv[ip][d][d2][k] = s2 / diag; }
                            else ws.sv[ip][d][d2][k] = 0.0;
                        }
                        const double sumP = -1.0 * ws.G[ip][k][d] * p.inv_rho;
                        rhs += sumP * ws.s0[k * NF + P];
                        ws.sp[ip][d][k] = sumP / diag;
                    }
                    ws.svel[ip][d] = rhs / diag;
                }
            }
        } else {
            // matrix rows (stabilization.cpp:267-285 / :613-634)
            if (lane < NIP) {
                const int ip = lane;
                for (int j = 0; j < NIP; j++) {
                    double v = 0.0;
                    if (j == ip) { if (p.time_dep) v += 1.0 / p.dt; v += ws.cl.a[ip]; v += ws.cl.b[ip]; }
                    v -= ws.cl.uip[0][ip][j] * ws.cl.b[ip];
                    if (flow) v += ws.cl.c[ip] * (ws.cl.uip[0][ip][j] - ws.cl.uip[1][ip][j]);
                    ws.cl.M[ip][j] = v;
                }
                ws.cl.perm[ip] = ip;
            }
            __syncwarp();
            // LU with partial pivoting (GetInverse; App. B-5). Every lane takes the same decisions.
            for (int kk = 0; kk < NIP; kk++) {
                int pv = kk; double best = fabs(ws.cl.M[kk][kk]);
                for (int i = kk + 1; i < NIP; i++) { const double v = fabs(ws.cl.M[i][kk]); if (v > best) { best = v; pv = i; } }
                if (!(best > 0.0)) { if (lane == 0) atomicExch(errflag, 2); break; }
                __syncwarp();
                if (pv != kk) {
                    if (lane < NIP) { const double t = ws.cl.M[kk][lane]; ws.cl.M[kk][lane] = ws.cl.M[pv][lane]; ws.cl.M[pv][lane] = t; }
                    if (lane == NIP) { const int t = ws.cl.perm[kk]; ws.cl.perm[kk] = ws.cl.perm[pv]; ws.cl.perm[pv] = t; }
                }
                __syncwarp();
                if (lane > kk && lane < NIP) {
                    const double l = ws.cl.M[lane][kk] / ws.cl.M[kk][kk];
                    ws.cl.M[lane][kk] = l;
                    for (int j = kk + 1; j < NIP; j++) ws.cl.M[lane][j] -= l * ws.cl.M[kk][j];
                }
                __syncwarp();
            }
            if (lane < NIP) ws.cl.dinv[lane] = 1.0 / ws.cl.M[lane][lane];
            __syncwarp();
            // right-hand sides: ids [0, NV) velocity shapes, [NV, NV+NPR) pressure shapes, then DIM stab_vel rhs
            const int NV = flow ? DIM * DIM * NSH : NSH, NPR = DIM * NSH, NR = NV + NPR + DIM;
            for (int r = lane; r < NR; r += 32) {
                int kind, d = 0, d2 = 0, k = 0;
                if (r < NV) { kind = 0; if (flow) { d = r / (DIM * NSH); d2 = (r / NSH) % DIM; k = r % NSH; } else k = r; }
                else if (r < NV + NPR) { kind = 1; d = (r - NV) / NSH; k = (r - NV) % NSH; }
                else { kind = 2; d = r - NV - NPR; }
                double bvec[NIP];
#pragma unroll
                for (int i = 0; i < NIP; i++) {
                    const int ip = ws.cl.perm[i];
                    double v;
                    if (kind == 0) {
                        if (!flow || d2 == d) {
                            v = ws.cl.a[ip] * ws.N[ip][k] + ws.cl.b[ip] * ws.cl.ush[0][ip][k];
                            if (flow) {
                                v += ws.cl.c[ip] * (ws.cl.ush[1][ip][k] - ws.cl.ush[0][ip][k]);
                                for (int q = 0; q < DIM; q++) if (q != d) v -= ws.std[ip][q] * ws.G[ip][k][q];
                            }
                        } else v = ws.std[ip][d] * ws.G[ip][k][d2];
                    } else if (kind == 1) v = -1.0 * ws.G[ip][k][d] * p.inv_rho;
                    else {
                        // stab_vel: the solve is linear, so only the state-independent part r0 = src + old / dt goes through the
                        // system here; the shape part  sum_k sv(d,d2,k) s_k,d2 + sp(d,k) p_k  is added from the solved shapes below
                        v = p.has_source ? p.src[d] : 0.0;
                        if (p.time_dep) { double o = 0.0; for (int q = 0; q < NSH; q++) o += ws.N[ip][q] * ws.s1[q * NF + d]; v += o / p.dt; }
                    }
                    bvec[i] = v;
                }
#pragma unroll
                for (int i = 0; i < NIP; i++)
#pragma unroll
                    for (int j = 0; j < i; j++) bvec[i] -= ws.cl.M[i][j] * bvec[j];
#pragma unroll
                for (int i = NIP - 1; i >= 0; i--) {
#pragma unroll
                    for (int j = i + 1; j < NIP; j++) bvec[i] -= ws.cl.M[i][j] * bvec[j];
                    bvec[i] *= ws.cl.dinv[i];
                }
#pragma unroll
                for (int i = 0; i < NIP; i++) {
                    if (kind == 0) {
                        if (flow) ws.sv[i][d][d2][k] = bvec[i];
                        else for (int q = 0; q < DIM; q++) for (int q2 = 0; q2 < DIM; q2++) ws.sv[i][q][q2][k] = (q == q2) ? bvec[i] : 0.0;
                    } else if (kind == 1) ws.sp[i][d][k] = bvec[i];
                    else ws.svel[i][d] = bvec[i];
                }
            }
            __syncwarp();
            // stab_vel(ip, d) = M^-1 r0 + sum_k [ sum_d2 sv(ip,d,d2,k) s0_k,d2 + sp(ip,d,k) s0_k,P ]   (stabilization.cpp:288-292, 637-641)
            for (int t = lane; t < NIP * DIM; t += 32) {
                const int i = t / DIM, d = t - i * DIM;
                double v = ws.svel[i][d];
                for (int q = 0; q < NSH; q++) {
#pragma unroll
                    for (int d2 = 0; d2 < DIM; d2++) v += ws.sv[i][d][d2][q] * ws.s0[q * NF + d2];
                    v += ws.sp[i][d][q] * ws.s0[q * NF + P];
                }
                ws.svel[i][d] = v;
            }
        }
    }
    __syncwarp();
    // ---- R: transported velocity, blend, defect fluxes, factored flux derivative -> records ----
    // (the records alias the closure storage: every lane reads its upwind rows first, then the warp synchronises)
    IpGeo<E> g;
    double std[DIM], U[DIM], w = 1.0, up[NSH], cvx[NSH];
    if (lane < NIP) {
        const int ip = lane;
        g.from = t_edge<E>(ip, 0); g.to = t_edge<E>(ip, 1);
#pragma unroll
        for (int d = 0; d < DIM; d++) { g.n[d] = ws.n[ip][d]; std[d] = ws.std[ip][d]; U[d] = 0.0; }
#pragma unroll
        for (int k = 0; k < NSH; k++) { up[k] = 0.0; cvx[k] = 0.0; g.N[k] = ws.N[ip][k]; for (int d = 0; d < DIM; d++) g.G[k][d] = ws.G[ip][k][d]; }
        if (!p.stokes) {
            if constexpr (PAC) { for (int d = 0; d < DIM; d++) U[d] = ws.svel[ip][d]; }
            else {
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    const double s = ws.cl.ush[cslot][ip][k];
                    up[k] = s;
                    for (int d = 0; d < DIM; d++) U[d] += s * ws.u[k * NF + d];
                }
                if (conv_pos) {                              // upwind_vel with ip shapes, upwind_interface.h:351-356
                    for (int j = 0; j < NIP; j++) {
                        const double s = ws.cl.uip[cslot][ip][j];
                        for (int d = 0; d < DIM; d++) U[d] += s * ws.std[j][d];
#pragma unroll
                        for (int k = 0; k < NSH; k++) cvx[k] += ws.N[j][k] * s;      // fv1/navier_stokes_fv1.cpp:441-448
                    }
                }
            }
            if (p.peclet) w = peclet_blend<E>(U, g, ws.x, std, p.visc);
        }
    }
    __syncwarp();
    if (lane < NIP) {
        const int ip = lane;
        IpRec<E, PAC>& r = ws.rec[ip];
        const double prod = dotv<DIM>(std, g.n) * p.rho;
        if (p.what & W_DEF_A) {
            double gv[DIM][DIM];
            for (int d1 = 0; d1 < DIM; d1++) for (int d2 = 0; d2 < DIM; d2++) {
                double s = 0; for (int k = 0; k < NSH; k++) s += g.G[k][d2] * ws.u[k * NF + d1];
                gv[d1][d2] = s;
            }
            double pr = 0; for (int k = 0; k < NSH; k++) pr += g.N[k] * ws.u[k * NF + P];
            for (int d1 = 0; d1 < DIM; d1++) {
                double df = 0;
                for (int d2 = 0; d2 < DIM; d2++) df += gv[d1][d2] * g.n[d2];
                if (!p.laplace) for (int d2 = 0; d2 < DIM; d2++) df += gv[d2][d1] * g.n[d2];
                double f = df * (-1.0) * p.visc * p.rho;
                if (!p.stokes) f += U[d1] * prod;
                f += pr * g.n[d1];
                r.F[d1] = f;
            }
            r.F[P] = dotv<DIM>(ws.svel[ip], g.n) * p.rho;
        }
        if (p.what & W_JAC_A) {
            StabDense<E> S{&ws.sv[ip][0][0][0], &ws.sp[ip][0][0]};
            ip_coeffs<E, PAC>(p, g.n, g.N, g.G, up, cvx, U, w, prod, S, p.stab == STAB_FLOW, r);
        }
    }
    __syncwarp();
    // ---- C: column phase ----
    const int k = lane / NF, cf = lane - k * NF;
    if (lane >= L) return;
    if (p.what & (W_JAC_A | W_JAC_M)) {
        double acc[L];
#pragma unroll
        for (int i = 0; i < L; i++) acc[i] = 0.0;
        if (p.what & W_JAC_A) {
            static_for<NIP>([&](auto ipc) {
                constexpr int ip = decltype(ipc)::value;
                constexpr int f = edge_corner<E>(ip, 0), t = edge_corner<E>(ip, 1);
                double v[NF];
                jac_col<E, PAC>(ws.rec[ip], k, cf, v);
#pragma unroll
                for (int rf = 0; rf < NF; rf++) { acc[f * NF + rf] += v[rf]; acc[t * NF + rf] -= v[rf]; }
            });
#pragma unroll
            for (int i = 0; i < L; i++) acc[i] *= p.scale_a;
        }
        if ((p.what & W_JAC_M) && cf < DIM) {
            const double mv = p.scale_m * ws.vol[k] * p.rho;
#pragma unroll
            for (int a = 0; a < NSH; a++)
#pragma unroll
                for (int rf = 0; rf < DIM; rf++) if (a == k && rf == cf) acc[a * NF + rf] += mv;
        }
        if (SC == SC_LOCAL) {
            double* J = Jloc + e * (int64_t)(L * L);
#pragma unroll
            for (int a = 0; a < NSH; a++)
#pragma unroll
                for (int rf = 0; rf < NF; rf++) J[(rf * NSH + a) * L + (cf * NSH + k)] = acc[a * NF + rf];
        } else {
            const uint8_t* em = m.emap + e * (int64_t)(NSH * NSH);
#pragma unroll
            for (int a = 0; a < NSH; a++) {
                const int slot = em[a * NSH + k];
                const int64_t base = ws.rowbase[a] + (int64_t)slot * NF + cf;
                const int64_t rstride = (int64_t)ws.cnt[a] * NF;
#pragma unroll
                for (int rf = 0; rf < NF; rf++) {
                    double* q = val + base + rf * rstride;
                    atomicAdd(q, acc[a * NF + rf]);
                }
            }
        }
    }
    if (p.what & (W_DEF_A | W_DEF_M | W_RHS)) {
        double d = 0.0;
        if (p.what & W_DEF_A) {
#pragma unroll
            for (int t = 0; t < ET<E>::NINC; t++) d += (double)t_inc_sign<E>(k, t) * ws.rec[t_inc<E>(k, t)].F[cf];
        }
        if ((p.what & W_RHS) && p.has_source && cf < DIM) d -= p.src[cf] * ws.vol[k] * p.rho;
        d *= p.scale_a;
        if ((p.what & W_DEF_M) && cf < DIM) d += p.scale_m * ws.u[k * NF + cf] * ws.vol[k] * p.rho;
        if (SC == SC_LOCAL) dloc[e * (int64_t)L + cf * NSH + k] = d;
        else {
            double* q = def + (int64_t)ws.node[k] * NF + cf;
            atomicAdd(q, d);
        }
    }
}

}  // namespace nsb
