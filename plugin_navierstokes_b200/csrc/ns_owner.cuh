// ns_owner.cuh -- owner-computes FV1 assembly for sm_100a (diagonal stabilisation branch), two kernels:
//
//  (A) fv1_flux_kernel : one THREAD per element, warp-uniform control flow. For each SCVF (ip) of the element,
//      in reference order: StdVel -> upwind (No/Full/Skewed/LPS, ray search over constant-memory side tables)
//      -> diffusion length -> FIELDS/FLOW/none closure -> defect fluxes. Every SCVF is evaluated ONCE.
//      Nodal unknowns live in registers, corner coordinates / SCV volumes in thread-private shared columns
//      (conflict-free). The SCVF geometry (normal, ip, global gradients) comes from the table precomputed at
//      upload. Output: one compact FluxRec per (element, ip) = the state-dependent coefficients only.
//  (B) fv1_rows_kernel : one WARP per grid node (= NF consecutive CSR rows). Stages the FluxRecs and the
//      geometry of the <= CH*NINC SCVFs incident to the node with coalesced 128-bit loads, then lane =
//      (corner k, function cf) accumulates its column of every incident SCVF (fixed order -> bitwise
//      deterministic) into the node's rows in shared memory, and the finished rows are streamed to HBM exactly
//      once (st.global.cs): no atomics, no colouring, no read-modify-write, no zero-fill of the matrix.
//
// Arithmetic restated from fv1/navier_stokes_fv1.cpp:250-778, fv1/stabilization.cpp:122-241,436-587,805-850,
// upwind.cpp:52-80,133-172,381-430,505-575, fv1/diffusion_length.h:47-198.
#pragma once
#include <type_traits>
#include "ns_kernels.cuh"

namespace nsb {

// ---- precomputed SCVF geometry table -------------------------------------------------------------
// record of one (element, ip): [ n[DIM], xip[DIM], ds, pad... | G[d][k] d-major, k padded to even ]
template <int E> struct GeoRec {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    static constexpr int HEAD = (DIM == 3) ? 8 : 4;
    static constexpr int NSHP = (NSH + 1) & ~1;
    static constexpr int SZ = HEAD + DIM * NSHP;          // doubles, even -> 16-byte aligned records
};

template <int E>
__global__ void geom_kernel(int64_t n_elem, const int32_t* __restrict__ conn, const double* __restrict__ coords,
                            double* __restrict__ geo)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP;
    using R = GeoRec<E>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elem * NIP) return;
    const int64_t e = i / NIP; const int ip = (int)(i - e * NIP);
    double x[NSH * DIM];
#pragma unroll
    for (int k = 0; k < NSH; k++) {
        const int64_t nd = conn[e * NSH + k];
#pragma unroll
        for (int d = 0; d < DIM; d++) x[k * DIM + d] = coords[nd * DIM + d];
    }
    IpGeo<E> g;
    ip_geometry<E>(x, ip, g);
    double* r = geo + i * R::SZ;
#pragma unroll
    for (int d = 0; d < DIM; d++) { r[d] = g.n[d]; r[DIM + d] = g.xip[d]; }
    if (DIM == 3) { r[6] = g.ds; r[7] = 0.0; }
#pragma unroll
    for (int d = 0; d < DIM; d++)
#pragma unroll
        for (int k = 0; k < R::NSHP; k++) r[R::HEAD + d * R::NSHP + k] = (k < NSH) ? g.G[k < NSH ? k : 0][d] : 0.0;
}

// ---- compact per-(element, ip) record written by (A), read by (B) ------------------------------------
//   F[NF]   defect fluxes (momentum d, continuity)                       add_def_A_elem :686-776
//   inv     1/diag of the ip system (0 for no stabilisation)
//   sn      StdVel . n ; std[DIM] StdVel                                 (FLOW continuity coefficients)
//   cK[k]   (a N_k + b up_k + c (down_k - up_k)) * inv * rho  |  N_k rho (no stabilisation)
//   dK[k]   up_k * prod * w + prod (1-w) N_k                             convective diagonal, :430-468
//   EXACT:  eK[k] = rho (w up_k + (1-w) N_k [peclet]) , U[DIM]           exact-Newton extras, :521-549
template <int E, bool FLOWREC, bool EXACT> struct FluxRec {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1;
    static constexpr int O_F = 0, O_INV = NF, O_SN = NF + 1, O_STD = NF + 2;           // sn / std only when FLOWREC
    static constexpr int HEADRAW = NF + 1 + (FLOWREC ? 1 + DIM : 0);
    static constexpr int O_CK = (HEADRAW + 1) & ~1, O_DK = O_CK + NSH, O_EK = O_DK + NSH, O_U = O_EK + NSH;
    static constexpr int RAW = EXACT ? O_U + DIM : O_EK;
    static constexpr int SZ = (RAW + 3) & ~3;              // multiple of 32 bytes: sector-aligned records (hex FIELDS: 192 B)
};

NSB_DEV double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// thread-private column in shared memory: element i of the calling thread
#define NSB_COL(base, i) (base)[(i) * BS + tid]

// ray / side intersection with warp-uniform side tables (constant memory); corner coordinates in the
// thread's shared column xs. Same tests as side_ray_cut (ns_fv1.cuh), first hit in reference order wins.
template <int E, int BS>
NSB_DEV bool ray_cut_uniform(const double* __restrict__ xs, int tid, const double* from, const double* dir,
                             int& side_out, double* gcut, double* lcut)
{
    constexpr int DIM = ET<E>::DIM, NSIDE = ET<E>::NSIDE;
    constexpr double S = NSB_RAY_SMALL;
    bool found = false;
    int best = 0;
    double tn = 0.0, n1 = 0.0, n2 = 0.0, bdet = 1.0;
    if constexpr (DIM == 2) {
        const double dn2 = dir[0] * dir[0] + dir[1] * dir[1];
        for (int s = 0; s < NSIDE; s++) {
            const int p0 = tab::C_SIDE[E][s][0], p1 = tab::C_SIDE[E][s][1];
            const double x0 = NSB_COL(xs, p0 * 2), y0 = NSB_COL(xs, p0 * 2 + 1);
            const double ex = NSB_COL(xs, p1 * 2) - x0, ey = NSB_COL(xs, p1 * 2 + 1) - y0;
            const double det = dir[0] * (-ey) + dir[1] * ex;
            const double rx = x0 - from[0], ry = y0 - from[1];
            const double t_n = rx * (-ey) + ry * ex, b_n = dir[0] * ry - dir[1] * rx;
            const double sg = det > 0.0 ? 1.0 : -1.0, ad = fabs(det);
            const bool hit = !found && det * det > (S * S) * dn2 * (ex * ex + ey * ey) &&
                             b_n * sg >= -S * ad && b_n * sg <= (1.0 + S) * ad && t_n * sg <= 0.0;
            if (hit) { found = true; best = s; tn = t_n; n1 = b_n; bdet = det; }
        }
        if (!found) return false;
        const double t = tn / bdet, bc = n1 / bdet;
        const int p0 = tab::SIDE[E][best][0], p1 = tab::SIDE[E][best][1];
#pragma unroll
        for (int d = 0; d < 2; d++) {
            gcut[d] = from[d] + t * dir[d];
            lcut[d] = (1 - bc) * tab::CORNER[E][p0][d] + bc * tab::CORNER[E][p1][d];
        }
        side_out = best;
        return true;
    } else {
        const double dn2 = dotv<3>(dir, dir);
        constexpr int TPS = (E == E_HEX) ? 2 : 1;
        for (int i = 0; i < NSIDE * TPS; i++) {
            const int s = i / TPS, kk = i - s * TPS;
            const int p0 = tab::C_SIDE[E][s][0], p1 = tab::C_SIDE[E][s][1 + kk], p2 = tab::C_SIDE[E][s][2 + kk];
            double e1[3], e2[3], r[3], nrm[3], q[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double x0 = NSB_COL(xs, p0 * 3 + d);
                e1[d] = NSB_COL(xs, p1 * 3 + d) - x0; e2[d] = NSB_COL(xs, p2 * 3 + d) - x0; r[d] = from[d] - x0;
            }
            cross3(nrm, e1, e2);
            const double det = -dotv<3>(dir, nrm);
            const double t_n = dotv<3>(r, nrm);
            cross3(q, r, dir);
            const double b1n = dotv<3>(e2, q), b2n = -dotv<3>(e1, q);
            const double sg = det > 0.0 ? 1.0 : -1.0, ad = fabs(det);
            const bool hit = !found && det * det > (S * S) * dn2 * dotv<3>(nrm, nrm) &&
                             b1n * sg >= -S * ad && b2n * sg >= -S * ad && (b1n + b2n) * sg <= (1.0 + S) * ad && t_n * sg <= 0.0;
            if (hit) { found = true; best = i; tn = t_n; n1 = b1n; n2 = b2n; bdet = det; }
        }
        if (!found) return false;
        const double t = tn / bdet, b1 = n1 / bdet, b2 = n2 / bdet;
        const int s = best / TPS, kk = best - s * TPS;
        const int p0 = tab::SIDE[E][s][0], p1 = tab::SIDE[E][s][1 + kk], p2 = tab::SIDE[E][s][2 + kk];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            gcut[d] = from[d] + t * dir[d];
            lcut[d] = (1 - b1 - b2) * tab::CORNER[E][p0][d] + b1 * tab::CORNER[E][p1][d] + b2 * tab::CORNER[E][p2][d];
        }
        side_out = s;
        return true;
    }
}

// upwind shapes of the current ip (warp-uniform `type`, `from`, `to`); see upwind_ip in ns_fv1.cuh
template <int E, int BS>
NSB_DEV bool upwind_uniform(int type, const double* __restrict__ xs, int tid, const double* n, const double* xip,
                            const double* N, int from, int to, const double* vel, double* up, double& len)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    if (type == UPW_NO) {
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = N[k];
        len = 1.0;
        return true;
    }
    if (type == UPW_FULL) {                                      // upwind.cpp:150-171
        const double flux = dotv<DIM>(n, vel);
        const int co = flux > 0.0 ? from : to;
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = (k == co) ? 1.0 : 0.0;
        double s = 0;
#pragma unroll
        for (int d = 0; d < DIM; d++) { const double t = xip[d] - NSB_COL(xs, co * DIM + d); s += t * t; }
        len = sqrt(s);
        return true;
    }
#pragma unroll
    for (int k = 0; k < NSH; k++) up[k] = 0.0;
    if (sqrt(dotv<DIM>(vel, vel)) < 1e-14) { len = 1.0; return true; }      // upwind.cpp:407-413, 531-537
    int side = 0; double gc[DIM], lc[DIM];
    if (!ray_cut_uniform<E, BS>(xs, tid, xip, vel, side, gc, lc)) { len = 1.0; return false; }
    constexpr int NSC = (DIM == 2) ? 2 : (E == E_TET ? 3 : 4);
    if (type == UPW_SKEWED) {                                    // GetNodeNextToCut, upwind.cpp:337-379
        double mn = 1.79769313486231570e308; int bestc = 0;
#pragma unroll
        for (int i = 0; i < NSC; i++) {
            const int co = tab::SIDE[E][side][i];
            double dd = 0;
#pragma unroll
            for (int d = 0; d < DIM; d++) { const double t = gc[d] - NSB_COL(xs, co * DIM + d); dd += t * t; }
            if (dd < mn) { mn = dd; bestc = co; }
        }
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = (k == bestc) ? 1.0 : 0.0;
        double s = 0;
#pragma unroll
        for (int d = 0; d < DIM; d++) { const double t = xip[d] - NSB_COL(xs, bestc * DIM + d); s += t * t; }
        len = sqrt(s);
    } else {                                                     // LPS, upwind.cpp:562-573
        double Nc[NSH];
        lagrange<E>(lc, Nc);
        int mask = 0;
#pragma unroll
        for (int i = 0; i < NSC; i++) mask |= 1 << tab::SIDE[E][side][i];
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = ((mask >> k) & 1) ? Nc[k] : 0.0;
        len = sqrt(dist2<DIM>(xip, gc));
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// (A) flux kernel
// ------------------------------------------------------------------------------------------------
// FV1Geometry of SCVF `ip` (warp-uniform) from the thread's shared corner column; see ip_geometry (ns_fv1.cuh).
// cen = element barycentre (hoisted). dnt = local shape gradients at the ips [NIP][NSH][DIM] (shared memory).
// JI (inverse transposed Jacobian at the ip) is only computed when wantJ: global_grad(k) = JI * dnt[ip][k].
template <int E, int BS>
NSB_DEV void ip_geometry_col(const double* __restrict__ xs, int tid, int ip, const double* cen, const double* __restrict__ dnt,
                             double* n, double* xip, double& ds, bool wantJ, double (*JI)[ET<E>::DIM])
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    const int f = tab::C_EDGE[E][ip][0], t = tab::C_EDGE[E][ip][1];
    double c0[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) c0[d] = 0.5 * (NSB_COL(xs, f * DIM + d) + NSB_COL(xs, t * DIM + d));
    if constexpr (DIM == 2) {
        n[0] = cen[1] - c0[1]; n[1] = -(cen[0] - c0[0]);
        xip[0] = 0.5 * (c0[0] + cen[0]); xip[1] = 0.5 * (c0[1] + cen[1]);
        ds = 0.0;
    } else {
        const int fa = tab::C_FA[E][ip], fb = tab::C_FB[E][ip];
        constexpr int NFC = (E == E_TET) ? 3 : 4;
        double c1[3] = {0, 0, 0}, c3[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < NFC; q++) {
            const int ka = tab::C_SIDE[E][fa][q], kb = tab::C_SIDE[E][fb][q];
#pragma unroll
            for (int d = 0; d < 3; d++) { c1[d] += NSB_COL(xs, ka * 3 + d); c3[d] += NSB_COL(xs, kb * 3 + d); }
        }
        double a[3], b[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            c1[d] *= (1.0 / NFC); c3[d] *= (1.0 / NFC);
            a[d] = cen[d] - c0[d]; b[d] = c3[d] - c1[d];
            xip[d] = 0.25 * (c0[d] + c1[d] + cen[d] + c3[d]);
        }
        cross3(n, a, b);
#pragma unroll
        for (int d = 0; d < 3; d++) n[d] *= 0.5;
        ds = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    }
    if (wantJ) {
        double JT[DIM][DIM];
#pragma unroll
        for (int i = 0; i < DIM; i++)
#pragma unroll
            for (int j = 0; j < DIM; j++) JT[i][j] = 0.0;
        const double* dn = dnt + ip * (NSH * DIM);
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            double dk[DIM];
#pragma unroll
            for (int i = 0; i < DIM; i++) dk[i] = dn[k * DIM + i];
#pragma unroll
            for (int j = 0; j < DIM; j++) {
                const double xkj = NSB_COL(xs, k * DIM + j);
#pragma unroll
                for (int i = 0; i < DIM; i++) JT[i][j] += dk[i] * xkj;
            }
        }
        inv_mat<DIM>(JT, JI);
    }
}

template <int E, int STAB, bool EXACT, int BS, int MINB = 3>
__global__ void __launch_bounds__(BS, MINB) fv1_flux_kernel(KParams p, MeshDev m, const double* __restrict__ geo,
                                                      const double* __restrict__ u, const double* __restrict__ s0,
                                                      const double* __restrict__ s1, double* __restrict__ flux,
                                                      int* __restrict__ errflag)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP, NF = DIM + 1, P = DIM;
    constexpr bool FLOW = (STAB == STAB_FLOW);
    using R = GeoRec<E>;
    using FR = FluxRec<E, FLOW, EXACT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* xs = reinterpret_cast<double*>(smem_raw);            // [NSH*DIM][BS]
    double* vs = xs + NSH * DIM * BS;                            // [NSH][BS]
    double* us = vs + NSH * BS;                                  // [NSH*NF][BS] nodal unknowns (the `u` argument)
    double* dnt = us + NSH * NF * BS;                            // [NIP][NSH][DIM] local shape gradients at the ips
    double* Nt = dnt + NIP * NSH * DIM;                          // [NIP][NSH]      shape values at the ips
    const int tid = threadIdx.x;
    for (int i = tid; i < NIP * NSH * DIM; i += BS) dnt[i] = tab::C_DNIP[E][i / (NSH * DIM)][(i / DIM) % NSH][i % DIM];
    for (int i = tid; i < NIP * NSH; i += BS) Nt[i] = tab::NIPSH[E][i / NSH][i % NSH];
    __syncthreads();
    const int64_t e = (int64_t)blockIdx.x * BS + tid;
    if (e >= m.n_elem) return;                                   // no block-wide barriers below
    const bool td = p.time_dep;
    // ---- element data: unknowns in registers, coordinates / volumes in the thread's shared column ----
    int nd[NSH];
#pragma unroll
    for (int k = 0; k < NSH; k++) nd[k] = m.conn[e * NSH + k];
#pragma unroll
    for (int k = 0; k < NSH; k++) {
        if (NF == 4) {
            const double2 a = ldg2(u + (int64_t)nd[k] * 4), b = ldg2(u + (int64_t)nd[k] * 4 + 2);
            NSB_COL(us, k * NF + 0) = a.x; NSB_COL(us, k * NF + 1) = a.y; NSB_COL(us, k * NF + 2) = b.x; NSB_COL(us, k * NF + NF - 1) = b.y;
        } else {
#pragma unroll
            for (int f = 0; f < NF; f++) NSB_COL(us, k * NF + f) = u[(int64_t)nd[k] * NF + f];
        }
#pragma unroll
        for (int d = 0; d < DIM; d++) NSB_COL(xs, k * DIM + d) = m.coords[(int64_t)nd[k] * DIM + d];
        NSB_COL(vs, k) = m.scvvol[e * NSH + k];
    }
    const double nurho = p.visc * p.rho;
    const bool want_def = p.what & W_DEF_A, want_jac = p.what & W_JAC_A;
    bool ok = true;
    double cen[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < NSH; k++) s += NSB_COL(xs, k * DIM + d);
        cen[d] = s * (1.0 / NSH);
    }
    // COR diffusion length: element-wide statistics of the SCVF normals (diffusion_length.h:139-172)
    double cmn = 0.0, cav = 0.0, cmd = 0.0;
    if (STAB != STAB_NONE && p.diff_len == DIFF_COR) {
        cmn = 1.79769313486231570e308; cmd = 1.79769313486231570e308;
        for (int i = 0; i < NIP; i++) {
            double nn_[DIM], xx_[DIM], dsi;
            ip_geometry_col<E, BS>(xs, tid, i, cen, dnt, nn_, xx_, dsi, false, nullptr);
            const double q = dotv<DIM>(nn_, nn_);
            if (q < cmn) cmn = q;
            cav += q;
            if (DIM == 3 && dsi < cmd) cmd = dsi;
        }
        cav /= NIP;
    }

    for (int ip = 0; ip < NIP; ip++) {
        double* fr = flux + (e * NIP + ip) * FR::SZ;
        const int from = tab::C_EDGE[E][ip][0], to = tab::C_EDGE[E][ip][1];
        double n[DIM], xip[DIM], ds = 0.0, JI[DIM][DIM];
        ip_geometry_col<E, BS>(xs, tid, ip, cen, dnt, n, xip, ds, want_def, JI);
        double N[NSH];
#pragma unroll
        for (int k = 0; k < NSH; k++) N[k] = Nt[ip * NSH + k];
        // ---- StdVel from the `u` argument (:282-293) ----
        double std[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < NSH; k++) s += NSB_COL(us, k * NF + d) * N[k];
            std[d] = s;
        }
        const double sn = dotv<DIM>(std, n);
        const double prod = sn * p.rho;
        // ---- upwinds ----
        double up[NSH], dnm[NSH], uplen = 1.0, dnlen = 1.0;
#pragma unroll
        for (int k = 0; k < NSH; k++) { up[k] = 0.0; dnm[k] = 0.0; }
        if (!p.stokes) {
            ok &= upwind_uniform<E, BS>(p.upw_stab, xs, tid, n, xip, N, from, to, std, up, uplen);
            if (FLOW) {                                          // update_downwind, upwind_interface.h:157-165
                double neg[DIM], dn[NSH];
#pragma unroll
                for (int d = 0; d < DIM; d++) neg[d] = -1.0 * std[d];
                ok &= upwind_uniform<E, BS>(p.upw_stab, xs, tid, n, xip, N, from, to, neg, dn, dnlen);
#pragma unroll
                for (int k = 0; k < NSH; k++) dnm[k] = dn[k] - up[k];
            }
        }
        // ---- diagonal of the ip system and numerators sb_k (stabilization.cpp:166-236 / :489-582) ----
        double inv = 0.0, sb[NSH];
        if (STAB != STAB_NONE) {
            const double nn = dotv<DIM>(n, n);
            const double a = p.visc * diff_len_sq_inv<DIM>(p.diff_len, nn, NSB_COL(vs, from), NSB_COL(vs, to), ds, cmn, cav, cmd);
            double b = 0.0, c = 0.0;
            if (!p.stokes) {
                const double nrm = sqrt(dotv<DIM>(std, std));
                b = nrm / uplen;
                if (FLOW) c = nrm / (dnlen + uplen);
            }
            double diag = a;
            if (td) diag += 1.0 / p.dt;
            if (!p.stokes) diag += b;
            inv = 1.0 / diag;
#pragma unroll
            for (int k = 0; k < NSH; k++) {
                double s = a * N[k];
                if (!p.stokes) { s += b * up[k]; if (FLOW) s += c * dnm[k]; }
                sb[k] = s;
            }
        } else {
#pragma unroll
            for (int k = 0; k < NSH; k++) sb[k] = 0.0;
        }
        // ---- convective upwind, transported velocity, Peclet blend ----
        double U[DIM], w = 1.0;
#pragma unroll
        for (int d = 0; d < DIM; d++) U[d] = 0.0;
        if (!p.stokes) {
            if (p.upw_conv != p.upw_stab) { double l2; ok &= upwind_uniform<E, BS>(p.upw_conv, xs, tid, n, xip, N, from, to, std, up, l2); }
#pragma unroll
            for (int k = 0; k < NSH; k++)
#pragma unroll
                for (int d = 0; d < DIM; d++) U[d] += up[k] * NSB_COL(us, k * NF + d);          // upwind_vel, upwind_interface.h:334-358
            if (p.peclet) {                                       // peclet_blend :871-892
                double dd = 0;
#pragma unroll
                for (int d = 0; d < DIM; d++) { const double t = NSB_COL(xs, to * DIM + d) - NSB_COL(xs, from * DIM + d); dd += t * t; }
                const double Pe = sn / dotv<DIM>(n, n) * sqrt(dd) / p.visc;
                const double Pe2 = Pe * Pe;
                w = Pe2 / (5.0 + Pe2);
#pragma unroll
                for (int d = 0; d < DIM; d++) U[d] = w * U[d] + (1.0 - w) * std[d];
            }
        }
        // ---- Jacobian coefficients ----
        if (want_jac) {
            const double cw = prod * w, cpe = prod * (1.0 - w);
            double ck[NSH], dk[NSH];
#pragma unroll
            for (int k = 0; k < NSH; k++) {
                ck[k] = (STAB == STAB_NONE) ? N[k] * p.rho : sb[k] * inv * p.rho;
                double D = 0.0;
                if (!p.stokes) { D = up[k] * cw; if (p.peclet) D += cpe * N[k]; }
                dk[k] = D;
            }
            if constexpr (NSH % 2 == 0) {
#pragma unroll
                for (int k = 0; k < NSH; k += 2) {
                    *reinterpret_cast<double2*>(fr + FR::O_CK + k) = make_double2(ck[k], ck[k + 1]);
                    *reinterpret_cast<double2*>(fr + FR::O_DK + k) = make_double2(dk[k], dk[k + 1]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < NSH; k++) { fr[FR::O_CK + k] = ck[k]; fr[FR::O_DK + k] = dk[k]; }
            }
            if constexpr (EXACT) {
                const bool exact = !p.stokes && p.exact_jac != 0.0;
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    double ev = 0.0;
                    if (exact) { ev = w * up[k] * p.rho; if (p.peclet) ev += (1.0 - w) * N[k] * p.rho; }   // quirks :528-529,:542-545
                    fr[FR::O_EK + k] = ev;
                }
#pragma unroll
                for (int d = 0; d < DIM; d++) fr[FR::O_U + d] = U[d];
            }
        }
        fr[FR::O_INV] = inv;
        if constexpr (FLOW) {
            fr[FR::O_SN] = sn;
#pragma unroll
            for (int d = 0; d < DIM; d++) fr[FR::O_STD + d] = std[d];
        }
        // ---- defect fluxes (:686-776): stream the global gradients (d-major) ----
        if (want_def) {
            double gv[DIM][DIM], gp[DIM], gv0[DIM][DIM], gp0[DIM], sG[NSH];
#pragma unroll
            for (int k = 0; k < NSH; k++) sG[k] = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                double Gd[NSH];                                  // global_grad(k)[d] = sum_i JI[d][i] * local_grad(k)[i]
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    double g = 0.0;
#pragma unroll
                    for (int i = 0; i < DIM; i++) g += JI[d][i] * dnt[(ip * NSH + k) * DIM + i];
                    Gd[k] = g;
                }
                double sp = 0.0, sv[DIM];
#pragma unroll
                for (int q = 0; q < DIM; q++) sv[q] = 0.0;
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    if (FLOW) sG[k] += Gd[k] * std[d];
                    sp += Gd[k] * NSB_COL(us, k * NF + P);
#pragma unroll
                    for (int q = 0; q < DIM; q++) sv[q] += Gd[k] * NSB_COL(us, k * NF + q);
                }
                gp[d] = sp;
#pragma unroll
                for (int q = 0; q < DIM; q++) gv[q][d] = sv[q];
                if (td) {                                        // the closure uses solution(0) (:296, :646)
                    double sp0 = 0.0, sv0[DIM];
#pragma unroll
                    for (int q = 0; q < DIM; q++) sv0[q] = 0.0;
#pragma unroll
                    for (int k = 0; k < NSH; k++) {
                        sp0 += Gd[k] * s0[(int64_t)nd[k] * NF + P];
#pragma unroll
                        for (int q = 0; q < DIM; q++) sv0[q] += Gd[k] * s0[(int64_t)nd[k] * NF + q];
                    }
                    gp0[d] = sp0;
#pragma unroll
                    for (int q = 0; q < DIM; q++) gv0[q][d] = sv0[q];
                }
            }
            double pr = 0.0;
#pragma unroll
            for (int k = 0; k < NSH; k++) pr += N[k] * NSB_COL(us, k * NF + P);
            double F[NF];
#pragma unroll
            for (int d1 = 0; d1 < DIM; d1++) {
                double df = 0.0;
#pragma unroll
                for (int d2 = 0; d2 < DIM; d2++) df += gv[d1][d2] * n[d2];
                if (!p.laplace) {
#pragma unroll
                    for (int d2 = 0; d2 < DIM; d2++) df += gv[d2][d1] * n[d2];
                }
                double f = df * (-1.0) * nurho;
                if (!p.stokes) f += U[d1] * prod;
                f += pr * n[d1];
                F[d1] = f;
            }
            double cont;
            if (STAB == STAB_NONE) cont = sn * p.rho;
            else {
                // (stab_vel . n) rho with rhs_d = src_d + old_d/dt + sum_k [sv(d,d,k) s_dk + sum_{q!=d} sv(d,q,k) s_qk] - G_kd/rho p_k
                //  = n.src + n.old/dt + sum_k (sb_k [- std.G_k]) (s_k.n) [+ (std.n) div s] - (grad p . n)/rho
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    double sk = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) sk += (td ? s0[(int64_t)nd[k] * NF + d] : NSB_COL(us, k * NF + d)) * n[d];
                    acc += (FLOW ? sb[k] - sG[k] : sb[k]) * sk;
                }
                double gpn = 0.0, div = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) { gpn += (td ? gp0[d] : gp[d]) * n[d]; div += td ? gv0[d][d] : gv[d][d]; }
                acc -= gpn * p.inv_rho;
                if (FLOW) acc += sn * div;
                if (p.has_source) {
#pragma unroll
                    for (int d = 0; d < DIM; d++) acc += p.src[d] * n[d];
                }
                if (td) {
                    double o = 0.0;
#pragma unroll
                    for (int k = 0; k < NSH; k++) {
                        double sk = 0.0;
#pragma unroll
                        for (int d = 0; d < DIM; d++) sk += s1[(int64_t)nd[k] * NF + d] * n[d];
                        o += N[k] * sk;
                    }
                    acc += o / p.dt;
                }
                cont = acc * inv * p.rho;
            }
            F[P] = cont;
#pragma unroll
            for (int f = 0; f < NF; f++) fr[FR::O_F + f] = F[f];
        }
    }
    if (!ok) atomicExch(errflag, 1);
}

// ------------------------------------------------------------------------------------------------
// (B) rows kernel
// ------------------------------------------------------------------------------------------------
template <int E> struct RowCfg {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1, NINC = ET<E>::NINC, NIP = ET<E>::NIP;
    static constexpr int CH = (DIM == 3) ? 8 : 16;          // adjacent elements per round
    static constexpr int NREC = CH * NINC;
};
template <int E, bool FLOWREC, bool EXACT> struct RowWS {
    using C = RowCfg<E>;
    // staged geometry = [n, (pad)] (NH doubles, the first NH of the record) + G[d][k]
    static constexpr int NH = (C::DIM + 1) & ~1, GS = NH + C::DIM * GeoRec<E>::NSHP;
    double geo[C::NREC][GS];
    double flx[C::NREC][FluxRec<E, FLOWREC, EXACT>::SZ];
    double vol[C::CH];
    int32_t ipx[C::NREC];           // ip | (256 if the node is the `to` corner of the SCVF)
    uint8_t slot[C::CH][8];
};

NSB_DEV void cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
NSB_DEV void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

template <int E, int STAB, bool EXACT>
__global__ void __launch_bounds__(96, 5) fv1_rows_kernel(KParams p, MeshDev m, const double* __restrict__ geo,
                                                          const double* __restrict__ flux, const double* __restrict__ u,
                                                          double beta, double* __restrict__ val, double* __restrict__ def,
                                                          unsigned long long* __restrict__ work_counter)
{
    using C = RowCfg<E>;
    constexpr int DIM = C::DIM, NSH = C::NSH, NF = C::NF, NINC = C::NINC, CH = C::CH, L = NSH * NF, NIP = C::NIP;
    constexpr bool FLOW = (STAB == STAB_FLOW);
    using R = GeoRec<E>;
    using FR = FluxRec<E, FLOW, EXACT>;
    using WS = RowWS<E, FLOW, EXACT>;
    constexpr int NH = WS::NH, GS = WS::GS;
    constexpr int GV = GS / 2, FV = FR::SZ / 2, HV = NH / 2;     // 16-byte chunks per staged record
    static_assert(GV <= 32 && FV <= 32, "record wider than a warp");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    // block layout: [Ntab NIP*NSH doubles][inc table NSH*NINC ints][per warp: WS | rowacc NF*NF*max_cnt doubles]
    double* Ntab = reinterpret_cast<double*>(smem_raw);
    int32_t* inctab = reinterpret_cast<int32_t*>(smem_raw + sizeof(double) * NIP * NSH);
    const size_t tab_bytes = (sizeof(double) * NIP * NSH + sizeof(int32_t) * NSH * NINC + 15) & ~(size_t)15;
    const size_t per_warp = (sizeof(WS) + sizeof(double) * NF * NF * m.max_cnt + 15) & ~(size_t)15;
    WS& ws = *reinterpret_cast<WS*>(smem_raw + tab_bytes + warp * per_warp);
    double* rowacc = reinterpret_cast<double*>(smem_raw + tab_bytes + warp * per_warp + sizeof(WS));
    for (int i = threadIdx.x; i < NIP * NSH; i += blockDim.x) Ntab[i] = tab::NIPSH[E][i / NSH][i % NSH];
    for (int i = threadIdx.x; i < NSH * NINC; i += blockDim.x)
        inctab[i] = tab::INC[E][i / NINC][i % NINC] | (tab::INC_SIGN[E][i / NINC][i % NINC] < 0 ? 256 : 0);
    __syncthreads();
    const bool want_jac = p.what & (W_JAC_A | W_JAC_M), want_def = p.what & (W_DEF_A | W_DEF_M | W_RHS);
    const bool jac_a = p.what & W_JAC_A, def_a = p.what & W_DEF_A;
    // lane = (corner k, function cf) of the element block; per-lane selectors make the accumulate branch-free
    const int k = (lane < L) ? lane / NF : 0, cf = (lane < L) ? lane - (lane / NF) * NF : 0;
    const bool isv = cf < DIM;                                   // velocity column / pressure column
    const int cfv = isv ? cf : 0;
    const double wv = isv ? 1.0 : 0.0, wp = isv ? 0.0 : 1.0;
    double msk[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) msk[d] = (isv && d == cf) ? 1.0 : 0.0;
    const double nurho_a = p.laplace ? 0.0 : -1.0 * p.visc * p.rho;   // -nu rho G_kd1 n_d2 vanishes for laplace (:346-356)
    const double nurho_d = -1.0 * p.visc * p.rho;
    const double rho_f = FLOW ? p.rho : 0.0;

    // dynamic work distribution: every warp atomically takes the next node of the traversal order. A static
    // grid-stride assignment lets the persistent warps drift apart, which destroys the L2 reuse of the SCVF
    // records shared by neighbouring nodes (measured: 1.75x table re-reads at 128^3 with the static loop).
    (void)nwarp;
    for (;;) {
        unsigned long long ai_u = 0;
        if (lane == 0) ai_u = atomicAdd(work_counter, 1ULL);
        const int64_t ai = (int64_t)__shfl_sync(0xffffffffu, ai_u, 0);
        if (ai >= m.n_node) break;
        const int64_t a = m.node_order ? (int64_t)m.node_order[ai] : ai;
        const int64_t q0 = m.adj_ptr[a], q1 = m.adj_ptr[a + 1];
        const int64_t b0 = m.brow[a];
        const int cnt = (int)(m.brow[a + 1] - b0);
        const int rowlen = cnt * NF;
        if (want_jac) for (int i = lane; i < NF * rowlen; i += 32) rowacc[i] = 0.0;
        double dsum = 0.0, volsum = 0.0;
        int self_slot = 0;
        for (int64_t qb = q0; qb < q1; qb += CH) {
            const int nj = (int)((q1 - qb) < CH ? (q1 - qb) : CH);
            const int nrec = nj * NINC;
            __syncwarp();
            // ---- adjacency of this round: lane j < nj holds (element, local corner) ----
            const int32_t ad = (lane < nj) ? m.adj[qb + lane] : 0;
            const int e_l = ad / NSH, la_l = ad - e_l * NSH;
            // record handled by this lane (r = lane < nrec): SCVF t of adjacent element j
            const int rj = lane / NINC, rt = lane - rj * NINC;
            const int e_r = __shfl_sync(0xffffffffu, e_l, rj < CH ? rj : 0);
            const int la_r = __shfl_sync(0xffffffffu, la_l, rj < CH ? rj : 0);
            const int ipx_r = inctab[la_r * NINC + rt];
            const int64_t gi_r = (int64_t)e_r * NIP + (ipx_r & 255);
            if (lane < nrec) ws.ipx[lane] = ipx_r;
            // ---- asynchronous staging: every 16-byte chunk of every incident record in flight at once ----
            for (int r = 0; r < nrec; r++) {
                const int64_t gi = __shfl_sync(0xffffffffu, gi_r, r);
                if (jac_a && lane < GV)       // chunks [0, HV) = normal, then the gradients (skipping xip / ds)
                    cp_async16(&ws.geo[r][2 * lane], geo + gi * R::SZ + (lane < HV ? 2 * lane : R::HEAD - NH + 2 * lane));
                if (lane < FV) cp_async16(&ws.flx[r][2 * lane], flux + gi * FR::SZ + 2 * lane);
            }
            // scatter slots + the node's SCV volume in the adjacent elements (plain loads, overlapped with the copies)
            if (lane < nj) {
                const uint8_t* em = m.emap + (int64_t)ad * NSH;
                if (NSH == 8) *reinterpret_cast<uint2*>(ws.slot[lane]) = __ldg(reinterpret_cast<const uint2*>(em));
                else if (NSH == 4) *reinterpret_cast<uint32_t*>(ws.slot[lane]) = __ldg(reinterpret_cast<const uint32_t*>(em));
                else { for (int q = 0; q < NSH; q++) ws.slot[lane][q] = em[q]; }
                ws.vol[lane] = m.scvvol[ad];
            }
            const int sslot = __shfl_sync(0xffffffffu, la_l, 0);
            cp_async_wait_all();
            __syncwarp();
            if (qb == q0) self_slot = ws.slot[0][sslot];
            // ---- accumulate: lane = (k, cf); fixed order j, t  (add_jac_A_elem :317-594) ----
            if (want_jac && lane < L) {
                for (int j = 0; j < nj; j++) {
                    double acc[NF];
#pragma unroll
                    for (int rf = 0; rf < NF; rf++) acc[rf] = 0.0;
                    if (jac_a) {
#pragma unroll
                        for (int t = 0; t < NINC; t++) {
                            const int r = j * NINC + t;
                            const double* gr = ws.geo[r];
                            const double* fl = ws.flx[r];
                            const int ipx = ws.ipx[r];
                            const double sg = (ipx & 256) ? -1.0 : 1.0;
                            double n[DIM], Gk[DIM];
#pragma unroll
                            for (int d = 0; d < DIM; d++) { n[d] = gr[d]; Gk[d] = gr[NH + d * R::NSHP + k]; }
                            const double gn = dotv<DIM>(Gk, n);
                            const double inv = fl[FR::O_INV];
                            const double ncf = gr[cfv];
                            const double Nk = Ntab[(ipx & 255) * NSH + k];
                            // velocity column: X = -nu rho G_k (+ e_k U), Y = n_cf ; pressure column: X = n, Y = N_k (:363-368)
                            const double Y = wv * ncf + wp * Nk;
                            const double D = nurho_d * gn + fl[FR::O_DK + k];
                            double ek = 0.0;
                            if constexpr (EXACT) ek = fl[FR::O_EK + k];
#pragma unroll
                            for (int d1 = 0; d1 < DIM; d1++) {
                                double X = wv * (nurho_a * Gk[d1]) + wp * n[d1];
                                if constexpr (EXACT) X += wv * ek * fl[FR::O_U + d1];
                                acc[d1] += sg * (X * Y + msk[d1] * D);
                            }
                            // continuity row: velocity column (:561-584), pressure column (:586-592, rho cancels)
                            double cv = fl[FR::O_CK + k] * ncf;
                            if constexpr (FLOW) {
                                // sum_q sv(q,d2,k) n_q rho = ((sb_k - std.G_k) n_d2 + G_k[d2] (std.n)) inv rho
                                double sG = 0.0;
#pragma unroll
                                for (int d = 0; d < DIM; d++) sG += fl[FR::O_STD + d] * Gk[d];
                                cv += (gr[NH + cfv * R::NSHP + k] * fl[FR::O_SN] - sG * ncf) * inv * rho_f;
                            }
                            const double cpv = (STAB == STAB_NONE) ? 0.0 : -1.0 * gn * inv;
                            acc[DIM] += sg * (wv * cv + wp * cpv);
                        }
#pragma unroll
                        for (int rf = 0; rf < NF; rf++) acc[rf] *= p.scale_a;
                    }
                    const int slot = ws.slot[j][k];
#pragma unroll
                    for (int rf = 0; rf < NF; rf++) rowacc[rf * rowlen + slot * NF + cf] += acc[rf];
                }
            }
            // ---- defect: lane r < nrec contributes its signed fluxes; deterministic butterfly reduction ----
            if (def_a) {
                double f[NF];
#pragma unroll
                for (int q = 0; q < NF; q++) f[q] = (lane < nrec) ? ((ipx_r & 256) ? -1.0 : 1.0) * ws.flx[lane][FR::O_F + q] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int q = 0; q < NF; q++) f[q] += __shfl_xor_sync(0xffffffffu, f[q], o);
#pragma unroll
                for (int q = 0; q < NF; q++) if (lane == q) dsum += f[q];
            }
            if (lane < NF) for (int j = 0; j < nj; j++) volsum += ws.vol[j];
        }
        __syncwarp();
        if (want_jac) {
            if ((p.what & W_JAC_M) && lane < DIM) rowacc[lane * rowlen + self_slot * NF + lane] += p.scale_m * volsum * p.rho;
            __syncwarp();
            double* out = val + b0 * (NF * NF);
            const int tot = NF * rowlen;
            if (beta == 0.0) {
                if ((NF * NF) % 2 == 0) {
                    for (int i = 2 * lane; i < tot; i += 64) __stcs(reinterpret_cast<double2*>(out + i), make_double2(rowacc[i], rowacc[i + 1]));
                } else for (int i = lane; i < tot; i += 32) __stcs(out + i, rowacc[i]);
            } else for (int i = lane; i < tot; i += 32) out[i] = beta * out[i] + rowacc[i];
        }
        if (want_def && lane < NF) {
            double d = def_a ? dsum : 0.0;
            if ((p.what & W_RHS) && p.has_source && lane < DIM) d -= p.src[lane] * volsum * p.rho;
            d *= p.scale_a;
            if ((p.what & W_DEF_M) && lane < DIM) d += p.scale_m * u[a * NF + lane] * volsum * p.rho;
            double* q = def + a * NF + lane;
            *q = (beta == 0.0) ? d : beta * (*q) + d;
        }
    }
}

}  // namespace nsb
