// ns_gather.cuh -- owner-computes FV1 assembly ("gather") for sm_100a, diagonal stabilisation branch.
//
// One warp owns one grid node = NF consecutive rows of the global CSR matrix (and NF defect entries).
//   stage   : the <= CH elements adjacent to the node: nodal unknowns / corner coordinates / SCV volumes /
//             scatter slots -> padded (bank-conflict-free) shared memory, vectorised coalesced loads
//   phase 1 : lane = (adjacent element j, SCVF t incident to the node). The SCVF geometry (normal, ip,
//             global shape gradients) comes from the table precomputed at upload (geom_kernel), streamed
//             with 128-bit read-only loads; StdVel -> upwind -> diffusion length -> FIELDS/FLOW/none
//             closure -> defect fluxes -> factored flux derivative (FRec) in shared memory
//   phase 2 : lane = (corner k, function cf) = one column of the element block; the node's rows are summed
//             in shared memory in a fixed order (element-ascending) -> bitwise deterministic
//   store   : the finished NF rows (contiguous in CSR) are streamed to HBM exactly once (st.global.cs);
//             no atomics, no colouring, no read-modify-write, no zero-fill of the matrix.
// Arithmetic restated from fv1/navier_stokes_fv1.cpp:250-778, fv1/stabilization.cpp:122-241,436-587,805-850,
// upwind.cpp:52-80,133-172,381-430,505-575, fv1/diffusion_length.h (see ns_fv1.cuh for the building blocks).
#pragma once
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

// ---- shared-memory record of one ip (factored flux derivative, see IpRec in ns_fv1.cuh) --------------
// FULLC (FLOW): continuity-row coefficients C[k][d2] are full; otherwise C[k][d2] = c[k]*n[d2].
// The sign of the SCVF w.r.t. the owning node (+1: node is `from`, -1: `to`) is folded into A, D, C, CP, F, sn.
template <int E, bool FULLC> struct FRec {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1;
    static constexpr int NC = FULLC ? NSH * DIM : NSH;
    static constexpr int O_N = 0, O_SN = DIM, O_F = O_SN + DIM, O_IP = O_F + NF, O_A = O_IP + 1, O_D = O_A + NSH * DIM,
                         O_C = O_D + NSH, O_CP = O_C + NC, RAW = O_CP + NSH;
    static constexpr int SZ = RAW | 1;          // odd number of doubles -> conflict-free lane-strided access
    double v[SZ];
    NSB_DEV double& n(int d) { return v[O_N + d]; }
    NSB_DEV double& sn(int d) { return v[O_SN + d]; }
    NSB_DEV double sn(int d) const { return v[O_SN + d]; }
    NSB_DEV double& F(int f) { return v[O_F + f]; }
    NSB_DEV double& ipd() { return v[O_IP]; }
    NSB_DEV double& A(int k, int d) { return v[O_A + k * DIM + d]; }
    NSB_DEV double& D(int k) { return v[O_D + k]; }
    NSB_DEV double& C(int i) { return v[O_C + i]; }
    NSB_DEV double& CP(int k) { return v[O_CP + k]; }
    NSB_DEV double n(int d) const { return v[O_N + d]; }
    NSB_DEV double F(int f) const { return v[O_F + f]; }
    NSB_DEV double ipd() const { return v[O_IP]; }
    NSB_DEV double A(int k, int d) const { return v[O_A + k * DIM + d]; }
    NSB_DEV double D(int k) const { return v[O_D + k]; }
    NSB_DEV double C(int i) const { return v[O_C + i]; }
    NSB_DEV double CP(int k) const { return v[O_CP + k]; }
};

template <int E> struct GCfg {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1, NINC = ET<E>::NINC, NIP = ET<E>::NIP;
    static constexpr int CH = (DIM == 3) ? 8 : 16;          // adjacent elements per round
    static constexpr int NREC = CH * NINC;
    static constexpr int US = NSH * NF + 1;                 // padded strides (odd)
    static constexpr int XS = (NSH * DIM) | 1;
    static constexpr int VS = NSH | 1;
};

template <int E, bool FULLC> struct GWS {
    using C = GCfg<E>;
    double u[C::CH * C::US];
    double s0[C::CH * C::US];
    double s1[C::CH * C::US];
    double x[C::CH * C::XS];
    double vol[C::CH * C::VS];
    FRec<E, FULLC> rec[C::NREC];
    int32_t elem[C::CH];
    int32_t la[C::CH];
    uint8_t slot[C::CH][C::NSH];
};
// stationary calls do not need s0/s1: the layout below drops them
template <int E, bool FULLC> struct GWS_stat {
    using C = GCfg<E>;
    double u[C::CH * C::US];
    double x[C::CH * C::XS];
    double vol[C::CH * C::VS];
    FRec<E, FULLC> rec[C::NREC];
    int32_t elem[C::CH];
    int32_t la[C::CH];
    uint8_t slot[C::CH][C::NSH];
};

NSB_DEV double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// phase 1 of one (element, ip): everything of SCVF `ip` of the staged element.
//   us : the `u` argument [k][f]; ss0/ss1: local time series (TD) ; xs corner coords ; vs SCV volumes
//   g  : geometry record (global, read-only) ; Nt : shape values at this ip (shared table)
template <int E, int STAB, bool TD>
NSB_DEV bool ip_fast(const KParams& p, const double* __restrict__ us, const double* __restrict__ ss0,
                     const double* __restrict__ ss1, const double* __restrict__ xs, const double* __restrict__ vs,
                     const double* __restrict__ g, const double* __restrict__ gelem, const double* __restrict__ Nt, int ip,
                     double sg, FRec<E, STAB == STAB_FLOW>& r)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1, P = DIM;
    using R = GeoRec<E>;
    constexpr bool FLOW = (STAB == STAB_FLOW);
    bool ok = true;
    // ---- geometry head ----
    IpGeo<E> gg;                       // only n, xip, N, from, to, ds are filled (G is streamed below)
    {
        double h[R::HEAD];
#pragma unroll
        for (int i = 0; i < R::HEAD; i += 2) { const double2 v = ldg2(g + i); h[i] = v.x; h[i + 1] = v.y; }
#pragma unroll
        for (int d = 0; d < DIM; d++) { gg.n[d] = h[d]; gg.xip[d] = h[DIM + d]; }
        gg.ds = (DIM == 3) ? h[R::HEAD - 2] : 0.0;
    }
    gg.from = tab::EDGE[E][ip][0]; gg.to = tab::EDGE[E][ip][1];
    double N[NSH];
#pragma unroll
    for (int k = 0; k < NSH; k++) { N[k] = Nt[k]; gg.N[k] = N[k]; }
    const double* n = gg.n;
    // ---- StdVel (from the `u` argument, :282-293) ----
    double std[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < NSH; k++) s += us[k * NF + d] * N[k];
        std[d] = s;
    }
    const double sn = dotv<DIM>(std, n);
    const double prod = sn * p.rho;
    // ---- upwinds ----
    double up[NSH], dnm[NSH], uplen = 1.0, dnlen = 1.0;      // dnm = down - up shapes (FLOW)
#pragma unroll
    for (int k = 0; k < NSH; k++) { up[k] = 0.0; dnm[k] = 0.0; }
    if (!p.stokes) {
        ok &= upwind_ip<E>(p.upw_stab, xs, gg, std, up, uplen);
        if (FLOW) {
            double neg[DIM], dn[NSH];
#pragma unroll
            for (int d = 0; d < DIM; d++) neg[d] = -1.0 * std[d];
            ok &= upwind_ip<E>(p.upw_stab, xs, gg, neg, dn, dnlen);
#pragma unroll
            for (int k = 0; k < NSH; k++) dnm[k] = dn[k] - up[k];
        }
    }
    // ---- diagonal of the ip system, numerators sb_k ----
    double invdiag = 0.0, sb[NSH];
    if (STAB != STAB_NONE) {
        const double nn = dotv<DIM>(n, n);
        double cmn = 0.0, cav = 0.0, cmd = 0.0;
        if (p.diff_len == DIFF_COR) {                    // element-wide min/avg (diffusion_length.h:139-172)
            cmn = 1.79769313486231570e308; cmd = 1.79769313486231570e308;
            for (int i = 0; i < ET<E>::NIP; i++) {
                const double* h = gelem + i * R::SZ;
                double q = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) { const double t = __ldg(h + d); q += t * t; }
                if (q < cmn) cmn = q;
                cav += q;
                if (DIM == 3) { const double t = __ldg(h + R::HEAD - 2); if (t < cmd) cmd = t; }
            }
            cav /= ET<E>::NIP;
        }
        const double a = p.visc * diff_len_sq_inv<DIM>(p.diff_len, nn, vs[gg.from], vs[gg.to], gg.ds, cmn, cav, cmd);
        double b = 0.0, c = 0.0;
        if (!p.stokes) {
            const double nrm = sqrt(dotv<DIM>(std, std));
            b = nrm / uplen;
            if (FLOW) c = nrm / (dnlen + uplen);
        }
        double diag = a;
        if (TD) diag += 1.0 / p.dt;
        if (!p.stokes) diag += b;
        invdiag = 1.0 / diag;
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
        if (p.upw_conv != p.upw_stab) { double l2; ok &= upwind_ip<E>(p.upw_conv, xs, gg, std, up, l2); }   // sb is final
#pragma unroll
        for (int k = 0; k < NSH; k++)
#pragma unroll
            for (int d = 0; d < DIM; d++) U[d] += up[k] * us[k * NF + d];
        if (p.peclet) w = peclet_blend<E>(U, gg, xs, std, p.visc);
    }
    // ---- stream the global gradients (d-major): gn_k = G_k.n, sG_k = std.G_k, A, defect sums ----
    const bool want_def = p.what & W_DEF_A, want_jac = p.what & W_JAC_A;
    const double nurho = p.visc * p.rho;
    double gn[NSH], sG[NSH];
    double gv[DIM][DIM], gp[DIM], gv0[DIM][DIM], gp0[DIM];     // grad u, grad p (of `u`; of sol0 when TD)
#pragma unroll
    for (int k = 0; k < NSH; k++) { gn[k] = 0.0; sG[k] = 0.0; }
    double ex[NSH];                                             // exact-Newton factor e_k (quirks :528-529,:542-545)
    const bool exact = !p.stokes && p.exact_jac != 0.0;
#pragma unroll
    for (int k = 0; k < NSH; k++) {
        double e = 0.0;
        if (exact) { e = w * up[k] * p.rho; if (p.peclet) e += (1.0 - w) * N[k] * p.rho; }
        ex[k] = e;
    }
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        double Gd[R::NSHP];
#pragma unroll
        for (int k = 0; k < R::NSHP; k += 2) { const double2 v = ldg2(g + R::HEAD + d * R::NSHP + k); Gd[k] = v.x; Gd[k + 1] = v.y; }
        double sp = 0.0, sp0 = 0.0, sv[DIM], sv0[DIM];
#pragma unroll
        for (int q = 0; q < DIM; q++) { sv[q] = 0.0; sv0[q] = 0.0; }
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            gn[k] += Gd[k] * n[d];
            if (FLOW) sG[k] += Gd[k] * std[d];
            if (want_jac) r.A(k, d) = sg * ((p.laplace ? 0.0 : -1.0 * nurho * Gd[k]) + ex[k] * U[d]);
            if (want_def) {
                sp += Gd[k] * us[k * NF + P];
#pragma unroll
                for (int q = 0; q < DIM; q++) sv[q] += Gd[k] * us[k * NF + q];
                if (TD) {
                    sp0 += Gd[k] * ss0[k * NF + P];
#pragma unroll
                    for (int q = 0; q < DIM; q++) sv0[q] += Gd[k] * ss0[k * NF + q];
                }
            }
        }
        gp[d] = sp; gp0[d] = sp0;
#pragma unroll
        for (int q = 0; q < DIM; q++) { gv[q][d] = sv[q]; gv0[q][d] = sv0[q]; }
    }
    // ---- Jacobian coefficients ----
    if (want_jac) {
#pragma unroll
        for (int d = 0; d < DIM; d++) { r.n(d) = n[d]; r.sn(d) = sg * n[d]; }
        r.ipd() = (double)ip;
        const double cw = prod * w, cpe = prod * (1.0 - w);
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            double D = -1.0 * nurho * gn[k];
            if (!p.stokes) { D += up[k] * cw; if (p.peclet) D += cpe * N[k]; }
            r.D(k) = sg * D;
            if constexpr (STAB == STAB_NONE) { r.C(k) = sg * N[k] * p.rho; r.CP(k) = 0.0; }
            else {
                r.CP(k) = -sg * gn[k] * invdiag;                       // sum_q sp(q,k) n_q rho, rho cancels
                if constexpr (!FLOW) r.C(k) = sg * sb[k] * invdiag * p.rho;
                else {
                    // sum_q sv(q,d2,k) n_q rho = ((sb_k - std.G_k) n_d2 + G_k[d2] (std.n)) invdiag rho
                    const double c0 = sg * (sb[k] - sG[k]) * invdiag * p.rho, c1 = sg * sn * invdiag * p.rho;
#pragma unroll
                    for (int d2 = 0; d2 < DIM; d2++) {
                        const double Gkd = __ldg(g + R::HEAD + d2 * R::NSHP + k);
                        r.C(k * DIM + d2) = c0 * n[d2] + Gkd * c1;
                    }
                }
            }
        }
    }
    // ---- defect fluxes (:686-776) ----
    if (want_def) {
        double pr = 0.0;
#pragma unroll
        for (int k = 0; k < NSH; k++) pr += N[k] * us[k * NF + P];
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
            r.F(d1) = sg * f;
        }
        // continuity: stab_vel . n * rho
        double cont;
        if (STAB == STAB_NONE) cont = sn * p.rho;
        else {
            // n . rhs with rhs_d = src_d + old_d/dt + sum_k sv(d,d,k) s_dk + sum_{q!=d} sv(d,q,k) s_qk - G_kd/rho p_k
            const double* s = TD ? ss0 : us;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < NSH; k++) {
                double sk = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) sk += s[k * NF + d] * n[d];
                acc += (FLOW ? sb[k] - sG[k] : sb[k]) * sk;
            }
            double gpn = 0.0, div = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) { gpn += (TD ? gp0[d] : gp[d]) * n[d]; div += TD ? gv0[d][d] : gv[d][d]; }
            acc -= gpn * p.inv_rho;
            if (FLOW) acc += sn * div;                 // sum_d n_d std_d sum_k sum_q G_kq s_qk
            if (p.has_source) {
#pragma unroll
                for (int d = 0; d < DIM; d++) acc += p.src[d] * n[d];
            }
            if (TD) {
                double o = 0.0;
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    double sk = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) sk += ss1[k * NF + d] * n[d];
                    o += N[k] * sk;
                }
                acc += o / p.dt;
            }
            cont = acc * invdiag * p.rho;
        }
        r.F(P) = sg * cont;
    }
    return ok;
}

template <int E, int STAB, bool TD>
__global__ void __launch_bounds__(128, 3) fv1_gather2_kernel(KParams p, MeshDev m, const double* __restrict__ geo,
                                                             const double* __restrict__ u, const double* __restrict__ s0,
                                                             const double* __restrict__ s1, double beta,
                                                             double* __restrict__ val, double* __restrict__ def,
                                                             int* __restrict__ errflag)
{
    using C = GCfg<E>;
    constexpr int DIM = C::DIM, NSH = C::NSH, NF = C::NF, NINC = C::NINC, CH = C::CH, L = NSH * NF, NIP = C::NIP;
    constexpr bool FULLC = (STAB == STAB_FLOW);
    using WS = typename std::conditional<TD, GWS<E, FULLC>, GWS_stat<E, FULLC>>::type;
    using R = GeoRec<E>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    // block layout: [Ntab NIP*NSH doubles][per warp: WS | rowacc NF*NF*max_cnt doubles]
    double* Ntab = reinterpret_cast<double*>(smem_raw);
    const size_t tab_bytes = (sizeof(double) * NIP * NSH + 15) & ~(size_t)15;
    const size_t per_warp = (sizeof(WS) + sizeof(double) * NF * NF * m.max_cnt + 15) & ~(size_t)15;
    WS& ws = *reinterpret_cast<WS*>(smem_raw + tab_bytes + warp * per_warp);
    double* rowacc = reinterpret_cast<double*>(smem_raw + tab_bytes + warp * per_warp + sizeof(WS));
    for (int i = threadIdx.x; i < NIP * NSH; i += blockDim.x) {
        const int ip = i / NSH, k = i - ip * NSH;
        double xi[DIM], Nv[NSH];
#pragma unroll
        for (int d = 0; d < DIM; d++) xi[d] = tab::LIP[E][ip][d];
        lagrange<E>(xi, Nv);
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < NSH; q++) if (q == k) v = Nv[q];
        Ntab[i] = v;
    }
    __syncthreads();
    const bool want_jac = p.what & (W_JAC_A | W_JAC_M), want_def = p.what & (W_DEF_A | W_DEF_M | W_RHS);
    const int k = lane / NF, cf = lane - k * NF;

    for (int64_t a = (int64_t)blockIdx.x * nwarp + warp; a < m.n_node; a += (int64_t)gridDim.x * nwarp) {
        const int64_t q0 = m.adj_ptr[a], q1 = m.adj_ptr[a + 1];
        const int64_t b0 = m.brow[a];
        const int cnt = (int)(m.brow[a + 1] - b0);
        const int rowlen = cnt * NF;
        if (want_jac) for (int i = lane; i < NF * rowlen; i += 32) rowacc[i] = 0.0;
        double dsum = 0.0, volsum = 0.0;
        int self_slot = 0;
        for (int64_t qb = q0; qb < q1; qb += CH) {
            const int nj = (int)((q1 - qb) < CH ? (q1 - qb) : CH);
            __syncwarp();
            if (lane < nj) {
                const int32_t ad = m.adj[qb + lane];
                const int e = ad / NSH;
                ws.elem[lane] = e; ws.la[lane] = ad - e * NSH;
            }
            __syncwarp();
            for (int i = lane; i < nj * NSH; i += 32) {
                const int j = i / NSH, kk = i - j * NSH;
                const int64_t e = ws.elem[j];
                const int64_t nd = m.conn[e * NSH + kk];
                ws.vol[j * C::VS + kk] = m.scvvol[e * NSH + kk];
                ws.slot[j][kk] = m.emap[e * (NSH * NSH) + ws.la[j] * NSH + kk];
#pragma unroll
                for (int d = 0; d < DIM; d++) ws.x[j * C::XS + kk * DIM + d] = m.coords[nd * DIM + d];
                if (NF == 4) {
                    const double2 v0 = ldg2(u + nd * NF), v1 = ldg2(u + nd * NF + 2);
                    double* q = ws.u + j * C::US + kk * NF;
                    q[0] = v0.x; q[1] = v0.y; q[2] = v1.x; q[3] = v1.y;
                } else {
#pragma unroll
                    for (int f = 0; f < NF; f++) ws.u[j * C::US + kk * NF + f] = u[nd * NF + f];
                }
                if constexpr (TD) {
#pragma unroll
                    for (int f = 0; f < NF; f++) { ws.s0[j * C::US + kk * NF + f] = s0[nd * NF + f]; ws.s1[j * C::US + kk * NF + f] = s1[nd * NF + f]; }
                }
            }
            __syncwarp();
            // ---- phase 1: lane = (j, t) ----
            if (p.what & (W_JAC_A | W_DEF_A)) {
                const int j = lane / NINC, t = lane - j * NINC;
                if (j < nj) {
                    const int ip = tab::INC[E][ws.la[j]][t];
                    const double* ge = geo + (int64_t)ws.elem[j] * NIP * R::SZ;
                    const double* gp = ge + ip * R::SZ;
                    const double* ps0 = nullptr; const double* ps1 = nullptr;
                    if constexpr (TD) { ps0 = ws.s0 + j * C::US; ps1 = ws.s1 + j * C::US; }
                    const double sg = (double)tab::INC_SIGN[E][ws.la[j]][t];
                    const bool ok = ip_fast<E, STAB, TD>(p, ws.u + j * C::US, ps0, ps1, ws.x + j * C::XS, ws.vol + j * C::VS,
                                                         gp, ge, Ntab + ip * NSH, ip, sg, ws.rec[lane]);
                    if (!ok) atomicExch(errflag, 1);
                }
            }
            __syncwarp();
            // ---- phase 2: lane = (k, cf); fixed order j, t (signs are already folded into the records) ----
            if (want_jac && lane < L) {
                for (int j = 0; j < nj; j++) {
                    double acc[NF];
#pragma unroll
                    for (int rf = 0; rf < NF; rf++) acc[rf] = 0.0;
                    if (p.what & W_JAC_A) {
#pragma unroll
                        for (int t = 0; t < NINC; t++) {
                            const FRec<E, FULLC>& r = ws.rec[j * NINC + t];
                            if (cf < DIM) {
                                const double ncf = r.n(cf);
#pragma unroll
                                for (int d1 = 0; d1 < DIM; d1++) acc[d1] += r.A(k, d1) * ncf;
#pragma unroll
                                for (int d1 = 0; d1 < DIM; d1++) if (d1 == cf) acc[d1] += r.D(k);
                                if constexpr (FULLC) acc[DIM] += r.C(k * DIM + cf);
                                else acc[DIM] += r.C(k) * ncf;
                            } else {
                                const double Nk = Ntab[(int)r.ipd() * NSH + k];
#pragma unroll
                                for (int d1 = 0; d1 < DIM; d1++) acc[d1] += Nk * r.sn(d1);
                                acc[DIM] += r.CP(k);
                            }
                        }
#pragma unroll
                        for (int rf = 0; rf < NF; rf++) acc[rf] *= p.scale_a;
                    }
                    const int slot = ws.slot[j][k];
#pragma unroll
                    for (int rf = 0; rf < NF; rf++) rowacc[rf * rowlen + slot * NF + cf] += acc[rf];
                }
            }
            if (lane < NF) {
                for (int j = 0; j < nj; j++) {
                    const int la = ws.la[j];
                    if (p.what & W_DEF_A) {
#pragma unroll
                        for (int t = 0; t < NINC; t++) dsum += ws.rec[j * NINC + t].F(lane);
                    }
                    volsum += ws.vol[j * C::VS + la];
                }
                if (qb == q0) self_slot = ws.slot[0][ws.la[0]];
            }
        }
        __syncwarp();
        if (want_jac) {
            if ((p.what & W_JAC_M) && lane < DIM) rowacc[lane * rowlen + self_slot * NF + lane] += p.scale_m * volsum * p.rho;
            __syncwarp();
            double* out = val + b0 * (NF * NF);
            const int tot = NF * rowlen;                 // multiple of NF*NF
            if (beta == 0.0) {
                if ((NF * NF) % 2 == 0) {
                    for (int i = 2 * lane; i < tot; i += 64) __stcs(reinterpret_cast<double2*>(out + i), make_double2(rowacc[i], rowacc[i + 1]));
                } else for (int i = lane; i < tot; i += 32) __stcs(out + i, rowacc[i]);
            } else for (int i = lane; i < tot; i += 32) out[i] = beta * out[i] + rowacc[i];
        }
        if (want_def && lane < NF) {
            double d = (p.what & W_DEF_A) ? dsum : 0.0;
            if ((p.what & W_RHS) && p.has_source && lane < DIM) d -= p.src[lane] * volsum * p.rho;
            d *= p.scale_a;
            if ((p.what & W_DEF_M) && lane < DIM) d += p.scale_m * u[a * NF + lane] * volsum * p.rho;
            double* q = def + a * NF + lane;
            *q = (beta == 0.0) ? d : beta * (*q) + d;
        }
    }
}

}  // namespace nsb
