// ns_owner.cuh -- owner-computes FV1 assembly for sm_100a (diagonal stabilisation branch), two kernels:
//
//  (A) fv1_flux_kernel : one THREAD per element, warp-uniform control flow. For each SCVF (ip) of the element,
//      in reference order: StdVel -> upwind (No/Full/Skewed/LPS, ray search over constant-memory side tables)
//      -> diffusion length -> FIELDS/FLOW/none closure -> defect fluxes. Every SCVF is evaluated ONCE.
//      Nodal unknowns, corner coordinates and SCV volumes live in thread-private shared columns (conflict-free).
//      Output: the flux half of one combined SCVF record per (element, ip) = the state-dependent coefficients
//      only; the static half (normal, G_k.n, global gradients) is written once per mesh by geom_kernel.
//  (B) fv1_rows_kernel : one WARP per grid node (= NF consecutive CSR rows), nodes handed out by an atomic
//      counter. Every incident SCVF record is fetched with ONE TMA bulk copy (cp.async.bulk -> UBLKCP)
//      completing on a warp-private mbarrier. lane = (element jj of JP in parallel, corner k) builds the NF x NF
//      block d r(node) / d u(corner k) of its element in registers (fixed SCVF order -> bitwise deterministic);
//      the blocks are parked in the consumed record area and added by all 32 lanes, one element at a time, into
//      a bank-rotated block-major row accumulator in shared memory. Finished rows are streamed to HBM exactly
//      once (st.global.cs): no atomics, no colouring, no global read-modify-write, no zero-fill of the matrix.
//
// Arithmetic restated from fv1/navier_stokes_fv1.cpp:250-778, fv1/stabilization.cpp:122-241,436-587,805-850,
// upwind.cpp:52-80,133-172,381-430,505-575, fv1/diffusion_length.h:47-198.
#pragma once
#include <type_traits>
#include "ns_kernels.cuh"

namespace nsb {

// ---- precomputed SCVF geometry table -------------------------------------------------------------
// record of one (element, ip): [ n[DIM], 0.. (HEAD doubles; slot HEAD-1 is always 0) | gn[k] = G_k . n | G[d][k] d-major ],
// k padded to even so that every part is 16-byte aligned. Static per mesh: built once in nsb_upload_mesh.
template <int E> struct GeoRec {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    static constexpr int HEAD = 4, O_ZERO = 3;
    static constexpr int NSHP = (NSH + 1) & ~1;
    static constexpr int O_GN = HEAD, O_G = HEAD + NSHP;
    static constexpr int SZ = O_G + DIM * NSHP;           // hex 36, tet 20, quad / tri 16 doubles
};

template <int E>
__global__ void geom_kernel(int64_t n_elem, const int32_t* __restrict__ conn, const double* __restrict__ coords,
                            double* __restrict__ rec, int stride)
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
    double* r = rec + i * stride;                          // geometry part of the combined SCVF record
#pragma unroll
    for (int d = 0; d < R::HEAD; d++) r[d] = (d < DIM) ? g.n[d < DIM ? d : 0] : 0.0;
#pragma unroll
    for (int k = 0; k < R::NSHP; k++) r[R::O_GN + k] = (k < NSH) ? dotv<DIM>(g.G[k < NSH ? k : 0], g.n) : 0.0;
#pragma unroll
    for (int d = 0; d < DIM; d++)
#pragma unroll
        for (int k = 0; k < R::NSHP; k++) r[R::O_G + d * R::NSHP + k] = (k < NSH) ? g.G[k < NSH ? k : 0][d] : 0.0;
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
NSB_DEV unsigned smem_u32(const void* q) { return (unsigned)__cvta_generic_to_shared(q); }
// stride (doubles) of the staged records of the flux kernel: an odd number of 16-byte units
__host__ __device__ constexpr int flux_stage_stride(int rsz) { return ((rsz / 2) % 2 == 0) ? rsz + 2 : rsz; }

// thread-private column in shared memory: element i of the calling thread
// element-major rows with an odd stride: the lanes that share an element (LPE > 1) read different corners from different banks,
// lanes of different elements read the same corner from different banks (measured: the [i * BS + tid] layout gave 4-way conflicts)
#define NSB_CSTR(E_) ((ET<E_>::NSH * (2 * ET<E_>::DIM + 2)) | 1)
#define NSB_COL(base, i) (base)[(i) + tid * NSB_CSTR(E)]

// ray / side intersection with warp-uniform side tables (constant memory); corner coordinates in the
// thread's shared column xs. Same tests as side_ray_cut (ns_fv1.cuh), first hit in reference order wins.
template <int E, int BS>
NSB_DEV bool ray_cut_uniform(const double* __restrict__ xs, int tid, const double* from, const double* dir,
                             int& side_out, double* gcut, double* lcut, const int* __restrict__ sidetab, const double* __restrict__ cortab,
                             int pred_side = -1)
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
        const double ibd = 1.0 / bdet;
        const double t = tn * ibd, bc = n1 * ibd;
        const int p0 = sidetab[best * 4], p1 = sidetab[best * 4 + 1];
#pragma unroll
        for (int d = 0; d < 2; d++) {
            gcut[d] = from[d] + t * dir[d];
            lcut[d] = (1 - bc) * cortab[p0 * 3 + d] + bc * cortab[p1 * 3 + d];
        }
        side_out = best;
        return true;
    } else {
        const double dn2 = dotv<3>(dir, dir);
        constexpr int TPS = (E == E_HEX) ? 2 : 1;
        const unsigned amask = __activemask();
        // hex, element star-shaped w.r.t. its ips (ns_fused.cuh, fused_star_shaped): the side predicted from the ray direction in
        // reference coordinates is confirmed with the exact tests first; exactly one boundary triangle is hit in that case, so the
        // ordered search below (run for the lanes whose confirmation failed) would return the same side
        const bool use_pred = __any_sync(amask, pred_side >= 0);
        for (int s = use_pred ? -1 : 0; s < NSIDE; s++) {
            if (s >= 0 && __all_sync(amask, found)) break;
            // a quadrilateral side is cut as the triangles (p0,p1,p2), (p0,p2,p3): both share r = from - x(p0), q = r x dir
            // and the diagonal edge; evaluating them together halves the dependent chains (same operations per value)
            const int sx = s < 0 ? (pred_side >= 0 ? pred_side : 0) : s;       // lane-dependent in the prediction round
            const bool lane_try = s >= 0 || pred_side >= 0;
            const int p0 = s < 0 ? sidetab[sx * 4] : tab::C_SIDE[E][s][0];
            double ed[TPS + 1][3], r[3], q[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double x0 = NSB_COL(xs, p0 * 3 + d);
                r[d] = from[d] - x0;
#pragma unroll
                for (int j = 0; j <= TPS; j++) ed[j][d] = NSB_COL(xs, (s < 0 ? sidetab[sx * 4 + 1 + j] : tab::C_SIDE[E][s][1 + j]) * 3 + d) - x0;
            }
            cross3(q, r, dir);
            double eq[TPS + 1];
#pragma unroll
            for (int j = 0; j <= TPS; j++) eq[j] = dotv<3>(ed[j], q);
#pragma unroll
            for (int kk = 0; kk < TPS; kk++) {
                double nrm[3];
                cross3(nrm, ed[kk], ed[kk + 1]);
                const double det = -dotv<3>(dir, nrm);
                const double t_n = dotv<3>(r, nrm);
                const double b1n = eq[kk + 1], b2n = -eq[kk];
                const double sg = det > 0.0 ? 1.0 : -1.0, ad = fabs(det);
                const bool hit = lane_try && !found && det * det > (S * S) * dn2 * dotv<3>(nrm, nrm) &&
                                 b1n * sg >= -S * ad && b2n * sg >= -S * ad && (b1n + b2n) * sg <= (1.0 + S) * ad && t_n * sg <= 0.0;
                if (hit) { found = true; best = sx * TPS + kk; tn = t_n; n1 = b1n; n2 = b2n; bdet = det; }
            }
            // the first hit wins (sides in reference order): the loop head stops as soon as every element of the warp has one
        }
        if (!found) return false;
        const double ibd = 1.0 / bdet;                            // one reciprocal instead of three dependent divisions
        const double t = tn * ibd, b1 = n1 * ibd, b2 = n2 * ibd;
        const int s = best / TPS, kk = best - s * TPS;
        const int p0 = sidetab[s * 4], p1 = sidetab[s * 4 + 1 + kk], p2 = sidetab[s * 4 + 2 + kk];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            gcut[d] = from[d] + t * dir[d];
            lcut[d] = (1 - b1 - b2) * cortab[p0 * 3 + d] + b1 * cortab[p1 * 3 + d] + b2 * cortab[p2 * 3 + d];
        }
        side_out = s;
        return true;
    }
}

// upwind shapes of the current ip (warp-uniform `type`, `from`, `to`); see upwind_ip in ns_fv1.cuh
template <int E, int BS>
NSB_DEV bool upwind_uniform(int type, const double* __restrict__ xs, int tid, const double* n, const double* xip,
                            const double* N, int from, int to, const double* vel, double* up, double& len,
                            const int* __restrict__ sidetab, const double* __restrict__ cortab, int pred_side = -1)
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
    if (!ray_cut_uniform<E, BS>(xs, tid, xip, vel, side, gc, lc, sidetab, cortab, pred_side)) { len = 1.0; return false; }
    constexpr int NSC = (DIM == 2) ? 2 : (E == E_TET ? 3 : 4);
    if (type == UPW_SKEWED) {                                    // GetNodeNextToCut, upwind.cpp:337-379
        double mn = 1.79769313486231570e308; int bestc = 0;
#pragma unroll
        for (int i = 0; i < NSC; i++) {
            const int co = sidetab[side * 4 + i];
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
        for (int i = 0; i < NSC; i++) mask |= 1 << sidetab[side * 4 + i];
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
// iptab (shared memory, [NIP][12] ints: from, to, corners of face A, corners of face B) replaces the constant-memory
// tables when the lanes of a warp work on different ips (LPE > 1): constant loads with diverging addresses serialise.
template <int E, int BS, int DSTR = ET<E>::NSH * ET<E>::DIM, bool SMTAB = false>
NSB_DEV void ip_geometry_col(const double* __restrict__ xs, int tid, int ip, const double* cen, const double* __restrict__ dnt,
                             double* n, double* xip, double& ds, bool wantJ, double (*JI)[ET<E>::DIM],
                             const int* __restrict__ iptab = nullptr)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    const int f = SMTAB ? iptab[ip * 12] : tab::C_EDGE[E][ip][0], t = SMTAB ? iptab[ip * 12 + 1] : tab::C_EDGE[E][ip][1];
    double c0[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) c0[d] = 0.5 * (NSB_COL(xs, f * DIM + d) + NSB_COL(xs, t * DIM + d));
    if constexpr (DIM == 2) {
        n[0] = cen[1] - c0[1]; n[1] = -(cen[0] - c0[0]);
        xip[0] = 0.5 * (c0[0] + cen[0]); xip[1] = 0.5 * (c0[1] + cen[1]);
        ds = 0.0;
    } else {
        const int fa = SMTAB ? 0 : tab::C_FA[E][ip], fb = SMTAB ? 0 : tab::C_FB[E][ip];
        constexpr int NFC = (E == E_TET) ? 3 : 4;
        double c1[3] = {0, 0, 0}, c3[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < NFC; q++) {
            const int ka = SMTAB ? iptab[ip * 12 + 2 + q] : tab::C_SIDE[E][fa][q], kb = SMTAB ? iptab[ip * 12 + 6 + q] : tab::C_SIDE[E][fb][q];
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
        const double* dn = dnt + ip * DSTR;
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

// lean SCVF record of the split path (ns_split.cuh): [F | n | cK | dK | pK = -G_k.n / diag]

// hex: side of the reference element the ray from the ip along -dir leaves through, predicted in reference coordinates
// (s = J^-1 dir; going upstream the first plane xi_i in {0, 1} reached); see fused_ray_cut in ns_fused.cuh
NSB_DEV int predict_hex_side(const double (*JI)[3], const double* dir, int ip)
{
    float best = -3.0e38f; int bs = -1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float si = (float)(JI[0][i] * dir[0] + JI[1][i] * dir[1] + JI[2][i] * dir[2]);
        const float xi = (float)tab::LIP[E_HEX][ip][i];
        if (si != 0.0f) {
            const float t = si > 0.0f ? -xi / si : (1.0f - xi) / si;
            const int sd = i == 0 ? (si > 0.0f ? 4 : 2) : (i == 1 ? (si > 0.0f ? 1 : 3) : (si > 0.0f ? 0 : 5));
            if (t > best) { best = t; bs = sd; }
        }
    }
    return bs;
}

// LPE = lanes per element: the SCVFs of an element are dealt to LPE adjacent lanes (ip = ii * LPE + sub), the element's
// unknowns / coordinates live once in a shared column used by all of them. LPE = 1 is the thread-per-element layout.
template <int E, int STAB, bool EXACT, int NT, int MINB = 3, bool LEAN = false, int LPE = 1, bool TD = true>
__global__ void __launch_bounds__(NT, MINB) fv1_flux_kernel(KParams p, MeshDev m,
                                                      const double* __restrict__ u, const double* __restrict__ s0,
                                                      const double* __restrict__ s1, double* __restrict__ rec,
                                                      int* __restrict__ errflag)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP, NF = DIM + 1, P = DIM;
    constexpr bool FLOW = (STAB == STAB_FLOW);
    constexpr int BS = NT / LPE;                                 // elements (shared columns) per block
    constexpr bool SMT = LPE > 1;
    constexpr int DSTR = NSH * DIM + (SMT ? 1 : 0), NSTR = NSH + (SMT ? 1 : 0);   // odd strides: the LPE ips of a warp hit disjoint banks
    static_assert(NT % LPE == 0 && 32 % LPE == 0 && NIP % LPE == 0 && NSH % LPE == 0, "bad LPE");
    using R = GeoRec<E>;
    using FR = FluxRec<E, FLOW, EXACT>;
    static_assert(!LEAN || (!FLOW && !EXACT), "the split path covers FIELDS / no stabilisation with the fixed-point Jacobian");
    constexpr int NSHP = (NSH + 1) & ~1;
    constexpr int L_N = NF, L_CK = (NF + DIM + 1) & ~1, L_DK = L_CK + NSHP, L_PK = L_DK + NSHP;   // = LeanRec<E> offsets
    constexpr bool CREC = LEAN && DIM == 3;                       // compressed record (CompRec<E>, ns_base.h)
    constexpr int LRSZ = CREC ? 14 + NSH : (L_PK + NSHP + 3) & ~3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* xs = reinterpret_cast<double*>(smem_raw);            // [BS][NSB_CSTR]: corner coordinates | SCV volumes | nodal unknowns (the `u` argument)
    double* vs = xs + NSH * DIM;
    double* us = vs + NSH;
    double* dnt = xs + NSB_CSTR(E) * BS;                         // [NIP][DSTR] local shape gradients at the ips
    double* Nt = dnt + NIP * DSTR;                               // [NIP][NSTR] shape values at the ips
    double* cortab = Nt + NIP * NSTR;                            // [8][3]      reference corners (tab::CORNER)
    int* sidetab = reinterpret_cast<int*>(cortab + 24);          // [6][4]      corners of the sides (tab::SIDE)
    int* iptab = sidetab + 24;                                   // [NIP][12]   from, to, face corners (LPE > 1)
    // STAGED: the record (lean / compressed record of the split path, flux half of the combined record otherwise) is assembled
    // in a lane-private shared-memory slot (an odd number of 16-byte units apart: the 16-byte stores of a quarter-warp hit
    // disjoint banks) and leaves the SM as ONE bulk store (cp.async.bulk shared -> global, SASS UBLKCP). Written straight to
    // global memory the 16-byte stores of a warp touch 32 different sectors each: the LSU data pipe was 84 % busy with them
    // (ncu, profiles/r2_ncu_summary.md).
    constexpr bool STAGED = true;
    constexpr int RSZ = LEAN ? LRSZ : FR::SZ;                     // doubles per staged record (a multiple of 2)
    constexpr int SSTR = flux_stage_stride(RSZ);
    double* stg = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(iptab + NIP * 12) + 15) & ~(uintptr_t)15);
    for (int i = threadIdx.x; i < 24; i += NT) {
        cortab[i] = tab::CORNER[E][i / 3][i % 3];
        const int v = tab::SIDE[E][i / 4][i % 4];
        sidetab[i] = v < 0 ? 0 : v;
    }
    for (int i = threadIdx.x; i < NIP * NSH * DIM; i += NT) dnt[(i / (NSH * DIM)) * DSTR + i % (NSH * DIM)] = tab::C_DNIP[E][i / (NSH * DIM)][(i / DIM) % NSH][i % DIM];
    for (int i = threadIdx.x; i < NIP * NSH; i += NT) Nt[(i / NSH) * NSTR + i % NSH] = tab::NIPSH[E][i / NSH][i % NSH];
    if constexpr (SMT) {
        for (int i = threadIdx.x; i < NIP * 12; i += NT) {
            const int ip = i / 12, j = i - ip * 12;
            int v = 0;
            if (j < 2) v = tab::EDGE[E][ip][j];
            else if (DIM == 3 && j < 10) v = (j < 6) ? tab::SIDE[E][tab::SCVF_FA[E][ip]][j - 2] : tab::SIDE[E][tab::SCVF_FB[E][ip]][j - 6];   // slots 10, 11 are padding
            iptab[i] = v < 0 ? 0 : v;
        }
    }
    __syncthreads();
    const int tid = threadIdx.x / LPE, sub = threadIdx.x - tid * LPE;   // element column of this lane, its ip group
    int64_t e = (int64_t)blockIdx.x * BS + tid;
    if constexpr (LPE == 1) { if (e >= m.n_elem) return; }       // no block-wide barriers below
    else if (e >= m.n_elem) e = m.n_elem - 1;                    // surplus lanes redo the last element (identical values)
    const bool td = TD && p.time_dep;                            // TD = false: the time-dependent closure is compiled out
    // ---- element data: unknowns, coordinates and volumes in the element's shared column ----
    const int32_t* nd = m.conn + e * NSH;                        // re-read where needed (time-dependent closure only)
#pragma unroll
    for (int kq = 0; kq < NSH / LPE; kq++) {
        const int k = kq * LPE + sub;
        const int64_t ndk = nd[k];
        if (NF == 4) {
            const double2 a = ldg2(u + ndk * 4), b = ldg2(u + ndk * 4 + 2);
            NSB_COL(us, k * NF + 0) = a.x; NSB_COL(us, k * NF + 1) = a.y; NSB_COL(us, k * NF + 2) = b.x; NSB_COL(us, k * NF + NF - 1) = b.y;
        } else {
#pragma unroll
            for (int f = 0; f < NF; f++) NSB_COL(us, k * NF + f) = u[ndk * NF + f];
        }
#pragma unroll
        for (int d = 0; d < DIM; d++) NSB_COL(xs, k * DIM + d) = m.coords[ndk * DIM + d];
        NSB_COL(vs, k) = m.scvvol[e * NSH + k];
    }
    if constexpr (SMT) __syncwarp();
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
            ip_geometry_col<E, BS, DSTR, SMT>(xs, tid, i, cen, dnt, nn_, xx_, dsi, false, nullptr, iptab);
            const double q = dotv<DIM>(nn_, nn_);
            if (q < cmn) cmn = q;
            cav += q;
            if (DIM == 3 && dsi < cmd) cmd = dsi;
        }
        cav /= NIP;
    }

    for (int ii = 0; ii < NIP / LPE; ii++) {
        const int ip = ii * LPE + sub;
        double* const frg = LEAN ? rec + (e * NIP + ip) * LRSZ                     // lean record of the split path
                                 : rec + (e * NIP + ip) * (R::SZ + FR::SZ) + R::SZ;      // flux part of the combined SCVF record
        double* fr = frg;
        if constexpr (STAGED) {
            if (ii > 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // the previous record has left the slot
            fr = stg + threadIdx.x * SSTR;
        }
        const int from = SMT ? iptab[ip * 12] : tab::C_EDGE[E][ip][0], to = SMT ? iptab[ip * 12 + 1] : tab::C_EDGE[E][ip][1];
        double n[DIM], xip[DIM], ds = 0.0, JI[DIM][DIM];
        ip_geometry_col<E, BS, DSTR, SMT>(xs, tid, ip, cen, dnt, n, xip, ds, want_def || (LEAN && want_jac), JI, iptab);
        const double* N = Nt + ip * NSTR;                        // shape values at the ip, re-read from shared memory at each use
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
        // predicted cut side of the upwind ray (hex elements that are star-shaped w.r.t. their ips, flagged once per mesh)
        int pred_up = -1, pred_dn = -1;
        if constexpr (E == E_HEX) {
            if ((want_def || (LEAN && want_jac)) && !p.stokes && m.elem_fast && m.elem_fast[e] && (p.upw_stab >= UPW_SKEWED || p.upw_conv >= UPW_SKEWED)) {
                pred_up = predict_hex_side(JI, std, ip);
                if (FLOW) { double ng[3] = {-std[0], -std[1], -std[2]}; pred_dn = predict_hex_side(JI, ng, ip); }
            }
        }
        if (!p.stokes) {
            ok &= upwind_uniform<E, BS>(p.upw_stab, xs, tid, n, xip, N, from, to, std, up, uplen, sidetab, cortab, pred_up);
            if (FLOW) {                                          // update_downwind, upwind_interface.h:157-165
                double neg[DIM], dn[NSH];
#pragma unroll
                for (int d = 0; d < DIM; d++) neg[d] = -1.0 * std[d];
                ok &= upwind_uniform<E, BS>(p.upw_stab, xs, tid, n, xip, N, from, to, neg, dn, dnlen, sidetab, cortab, pred_dn);
#pragma unroll
                for (int k = 0; k < NSH; k++) dnm[k] = dn[k] - up[k];
            }
        }
        // ---- diagonal of the ip system and numerators sb_k (stabilization.cpp:166-236 / :489-582) ----
        double inv = 0.0, qa = 0.0, qb = 0.0, qc = 0.0;          // sb_k = qa N_k + qb up_k + qc (down_k - up_k), formed where it is used
        if (STAB != STAB_NONE) {
            const double nn = dotv<DIM>(n, n);
            qa = p.visc * diff_len_sq_inv<DIM>(p.diff_len, nn, NSB_COL(vs, from), NSB_COL(vs, to), ds, cmn, cav, cmd);
            if (!p.stokes) {
                const double nrm = sqrt(dotv<DIM>(std, std));
                qb = nrm / uplen;
                if (FLOW) qc = nrm / (dnlen + uplen);
            }
            double diag = qa;
            if (td) diag += 1.0 / p.dt;
            if (!p.stokes) diag += qb;
            inv = 1.0 / diag;
        }
        auto sbk = [&](int k) -> double {
            if (STAB == STAB_NONE) return 0.0;
            double s = qa * N[k];
            if (!p.stokes) { s += qb * up[k]; if (FLOW) s += qc * dnm[k]; }
            return s;
        };
        // everything that uses the STABILISATION's upwind shapes happens here: the convective upwind below may overwrite `up`
        constexpr int W_CK = LEAN ? L_CK : FR::O_CK, W_DK = LEAN ? L_DK : FR::O_DK;
        if (want_jac && !CREC) {                                 // continuity-row coefficients (:561-584)
            if constexpr (NSH % 2 == 0) {
#pragma unroll
                for (int k = 0; k < NSH; k += 2) {
                    const double c0 = (STAB == STAB_NONE) ? N[k] * p.rho : sbk(k) * inv * p.rho;
                    const double c1 = (STAB == STAB_NONE) ? N[k + 1] * p.rho : sbk(k + 1) * inv * p.rho;
                    *reinterpret_cast<double2*>(fr + W_CK + k) = make_double2(c0, c1);
                }
            } else {
#pragma unroll
                for (int k = 0; k < NSH; k++) fr[W_CK + k] = (STAB == STAB_NONE) ? N[k] * p.rho : sbk(k) * inv * p.rho;
            }
        }
        // transported velocity of the stabilisation's upwind: Us = sum_k up_k u_k (upwind_vel, upwind_interface.h:334-358)
        double Us[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) Us[d] = 0.0;
        if (!p.stokes) {
#pragma unroll
            for (int k = 0; k < NSH; k++)
#pragma unroll
                for (int d = 0; d < DIM; d++) Us[d] += up[k] * NSB_COL(us, k * NF + d);
        }
        double acc = 0.0;                                        // closure sum  sum_k sb_k (s_k . n)
        if (STAB != STAB_NONE && want_def) {
            if (!FLOW && !td) {
                // stationary FIELDS: sum_k (qa N_k + qb up_k) (u_k . n) = (qa StdVel + qb Us) . n
#pragma unroll
                for (int d = 0; d < DIM; d++) acc += (qa * std[d] + qb * Us[d]) * n[d];
            } else {
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    double sk = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) sk += (td ? s0[(int64_t)nd[k] * NF + d] : NSB_COL(us, k * NF + d)) * n[d];
                    acc += sbk(k) * sk;
                }
            }
        }
        // ---- convective upwind, transported velocity, Peclet blend ----
        double U[DIM], w = 1.0;
#pragma unroll
        for (int d = 0; d < DIM; d++) U[d] = Us[d];
        if (!p.stokes) {
            if (p.upw_conv != p.upw_stab) {
                double l2;
                ok &= upwind_uniform<E, BS>(p.upw_conv, xs, tid, n, xip, N, from, to, std, up, l2, sidetab, cortab, pred_up);
#pragma unroll
                for (int d = 0; d < DIM; d++) U[d] = 0.0;
#pragma unroll
                for (int k = 0; k < NSH; k++)
#pragma unroll
                    for (int d = 0; d < DIM; d++) U[d] += up[k] * NSB_COL(us, k * NF + d);
            }
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
        if constexpr (CREC) {
            if (want_jac) {
                // compressed record: the consumer forms cK_k = alpha N_k + beta up_k, dK_k = cw up_k + cpe N_k, pK_k = dN_k . mv
                const double ci = inv * p.rho;
                const double alpha = (STAB == STAB_NONE) ? p.rho : qa * ci, beta = (STAB == STAB_NONE || p.stokes) ? 0.0 : qb * ci;
                const double cw = p.stokes ? 0.0 : prod * w, cpe = (p.stokes || !p.peclet) ? 0.0 : prod * (1.0 - w);
                double mv[DIM];
#pragma unroll
                for (int i = 0; i < DIM; i++) {
                    double s = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) s += JI[d][i] * n[d];
                    mv[i] = s * (-1.0 * inv);
                }
                *reinterpret_cast<double2*>(fr + 4) = make_double2(n[0], n[1]);
                *reinterpret_cast<double2*>(fr + 6) = make_double2(n[DIM - 1], alpha);
                *reinterpret_cast<double2*>(fr + 8) = make_double2(beta, cw);
                *reinterpret_cast<double2*>(fr + 10) = make_double2(cpe, mv[0]);
                *reinterpret_cast<double2*>(fr + 12) = make_double2(mv[1], mv[DIM - 1]);
#pragma unroll
                for (int k = 0; k < NSH; k += 2) *reinterpret_cast<double2*>(fr + 14 + k) = make_double2(p.stokes ? 0.0 : up[k], p.stokes ? 0.0 : up[k + 1]);
            }
        }
        if (want_jac && !CREC) {
            const double cw = prod * w, cpe = prod * (1.0 - w);
            if constexpr (NSH % 2 == 0) {
#pragma unroll
                for (int k = 0; k < NSH; k += 2) {
                    double D0 = 0.0, D1 = 0.0;
                    if (!p.stokes) { D0 = up[k] * cw; D1 = up[k + 1] * cw; if (p.peclet) { D0 += cpe * N[k]; D1 += cpe * N[k + 1]; } }
                    *reinterpret_cast<double2*>(fr + W_DK + k) = make_double2(D0, D1);
                }
            } else {
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    double D = 0.0;
                    if (!p.stokes) { D = up[k] * cw; if (p.peclet) D += cpe * N[k]; }
                    fr[W_DK + k] = D;
                }
            }
            if constexpr (LEAN) {
                // pressure column of the continuity row (:586-592): -G_k.n / diag, with G_k.n = dnt_k . (JI^T n)
                double mv[DIM], pk[NSH];
#pragma unroll
                for (int i = 0; i < DIM; i++) {
                    double s = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) s += JI[d][i] * n[d];
                    mv[i] = s * (-1.0 * inv);
                }
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    double s = 0.0;
#pragma unroll
                    for (int i = 0; i < DIM; i++) s += dnt[ip * DSTR + k * DIM + i] * mv[i];
                    pk[k] = s;
                }
                if constexpr (NSH % 2 == 0) {
#pragma unroll
                    for (int k = 0; k < NSH; k += 2) *reinterpret_cast<double2*>(fr + L_PK + k) = make_double2(pk[k], pk[k + 1]);
                } else {
#pragma unroll
                    for (int k = 0; k < NSH; k++) fr[L_PK + k] = pk[k];
                }
#pragma unroll
                for (int d = 0; d < DIM; d++) fr[L_N + d] = n[d];
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
        if constexpr (!LEAN) fr[FR::O_INV] = inv;
        if constexpr (FLOW) {
            fr[FR::O_SN] = sn;
#pragma unroll
            for (int d = 0; d < DIM; d++) fr[FR::O_STD + d] = std[d];
        }
        // ---- defect fluxes (:686-776): the local gradient tensor Lg[i][q] = sum_k dN_k/dxi_i u_k,q is summed first and
        // mapped to global gradients once (grad = J^-T Lg) instead of forming global_grad(k) per corner ----
        if (want_def) {
            double Lg[DIM][NF], L0[DIM][NF], ms[DIM];
#pragma unroll
            for (int i = 0; i < DIM; i++) {
#pragma unroll
                for (int q = 0; q < NF; q++) { Lg[i][q] = 0.0; L0[i][q] = 0.0; }
                double s = 0.0;                                  // FLOW: std . G_k = dN_k . (J^-1 std)
#pragma unroll
                for (int d = 0; d < DIM; d++) s += JI[d][i] * std[d];
                ms[i] = s;
            }
            double pr = 0.0, oacc = 0.0;
#pragma unroll
            for (int k = 0; k < NSH; k++) {
                double dl[DIM], uk[NF];
#pragma unroll
                for (int i = 0; i < DIM; i++) dl[i] = dnt[ip * DSTR + k * DIM + i];
#pragma unroll
                for (int q = 0; q < NF; q++) uk[q] = NSB_COL(us, k * NF + q);
#pragma unroll
                for (int i = 0; i < DIM; i++)
#pragma unroll
                    for (int q = 0; q < NF; q++) Lg[i][q] += dl[i] * uk[q];
                pr += N[k] * uk[P];
                if (STAB != STAB_NONE) {
                    // (stab_vel . n) rho with rhs_d = src_d + old_d/dt + sum_k [sv(d,d,k) s_dk + sum_{q!=d} sv(d,q,k) s_qk] - G_kd/rho p_k
                    //  = n.src + n.old/dt + sum_k (sb_k [- std.G_k]) (s_k.n) [+ (std.n) div s] - (grad p . n)/rho
                    double sk = 0.0;
                    if (td) {                                    // the closure uses solution(0) (:296, :646)
                        double s0k[NF], o = 0.0;
#pragma unroll
                        for (int q = 0; q < NF; q++) s0k[q] = s0[(int64_t)nd[k] * NF + q];
#pragma unroll
                        for (int i = 0; i < DIM; i++)
#pragma unroll
                            for (int q = 0; q < NF; q++) L0[i][q] += dl[i] * s0k[q];
#pragma unroll
                        for (int d = 0; d < DIM; d++) { sk += s0k[d] * n[d]; o += s1[(int64_t)nd[k] * NF + d] * n[d]; }
                        oacc += N[k] * o;
                    } else if (FLOW) {
#pragma unroll
                        for (int d = 0; d < DIM; d++) sk += uk[d] * n[d];
                    }
                    if (FLOW) acc -= dotv<DIM>(dl, ms) * sk;
                }
            }
            double gv[DIM][DIM], gp[DIM], gv0[DIM][DIM], gp0[DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                double sp_ = 0.0, sp0_ = 0.0;
#pragma unroll
                for (int i = 0; i < DIM; i++) { sp_ += JI[d][i] * Lg[i][P]; sp0_ += JI[d][i] * L0[i][P]; }
                gp[d] = sp_; gp0[d] = sp0_;
#pragma unroll
                for (int q = 0; q < DIM; q++) {
                    double sv_ = 0.0, sv0_ = 0.0;
#pragma unroll
                    for (int i = 0; i < DIM; i++) { sv_ += JI[d][i] * Lg[i][q]; sv0_ += JI[d][i] * L0[i][q]; }
                    gv[q][d] = sv_; gv0[q][d] = sv0_;
                }
            }
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
                double gpn = 0.0, div = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) { gpn += (td ? gp0[d] : gp[d]) * n[d]; div += td ? gv0[d][d] : gv[d][d]; }
                acc -= gpn * p.inv_rho;
                if (FLOW) acc += sn * div;
                if (p.has_source) {
#pragma unroll
                    for (int d = 0; d < DIM; d++) acc += p.src[d] * n[d];
                }
                if (td) acc += oacc / p.dt;
                cont = acc * inv * p.rho;
            }
            F[P] = cont;
#pragma unroll
            for (int f = 0; f < NF; f++) fr[FR::O_F + f] = F[f];
        }
        if constexpr (STAGED) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(frg), "r"(smem_u32(fr)), "r"((unsigned)(RSZ * sizeof(double))) : "memory");
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
    }
    if constexpr (STAGED) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    if (!ok) atomicExch(errflag, 1);
}

// ------------------------------------------------------------------------------------------------
// (B) rows kernel
// ------------------------------------------------------------------------------------------------
template <int E, int CHP = 0> struct RowCfg {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1, NINC = ET<E>::NINC, NIP = ET<E>::NIP;
    static constexpr int CH = CHP ? CHP : ((DIM == 3) ? 8 : 16);   // adjacent elements per round
    static constexpr int NREC = CH * NINC;
};
// shared-memory stride of a staged record: the lanes of a half-warp read the same field of JP different elements'
// records (NINC records apart); the stride is padded so that those reads fall into disjoint banks.
template <int E> constexpr int rows_smem_stride(int rs)
{
    constexpr int NSH = ET<E>::NSH, NINC = ET<E>::NINC;
    if (NSH == 3) return rs;
    for (int s = rs; s < rs + 32; s += 2) {
        const int off = (NINC * s * 8) % 128;             // byte offset (mod one bank sweep) between consecutive elements
        if (NSH == 8 && off == 64) return s;              // 2 elements x 8 corners x 8 B per half-warp
        if (NSH == 4 && (off == 32 || off == 96)) return s;   // 4 elements x 4 corners x 8 B per half-warp
    }
    return rs;
}
template <int E, bool FLOWREC, bool EXACT, int CHP = 0> struct RowWS {
    using C = RowCfg<E, CHP>;
    // staged SCVF record = [geometry record (GS doubles) | flux record], RS doubles in HBM, SS apart in shared memory
    static constexpr int GS = GeoRec<E>::SZ, RS = GS + FluxRec<E, FLOWREC, EXACT>::SZ, SS = rows_smem_stride<E>(RS);
    alignas(16) double rec[C::NREC][SS];
    unsigned long long bar;         // mbarrier the bulk copies of one round complete on
    int32_t ipx[C::NREC];           // ip | (256 if the node is the `to` corner of the SCVF)
    uint8_t slot[C::CH][8];
};
// block-level tables in front of the per-warp work spaces: shape values at the ips, incidence table, slot rotation
template <int E> __host__ __device__ constexpr size_t rows_tab_bytes(int max_cnt)
{
    return (sizeof(double) * ET<E>::NIP * ET<E>::NSH + sizeof(int32_t) * ET<E>::NSH * ET<E>::NINC + (size_t)max_cnt + 15) & ~(size_t)15;
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on a warp-private mbarrier ----
NSB_DEV void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
NSB_DEV void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
NSB_DEV void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
NSB_DEV void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// bulk copy with an L2 eviction-priority hint (createpolicy): evict_last for data with a second reader, evict_first for streams
NSB_DEV unsigned long long l2_policy_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
NSB_DEV unsigned long long l2_policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
NSB_DEV void bulk_g2s_hint(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar, unsigned long long policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

template <int E, int STAB, bool EXACT, int CHP = 0, int MINB = 5>
__global__ void __launch_bounds__(96, MINB) fv1_rows_kernel(KParams p, MeshDev m,
                                                          const double* __restrict__ rec, const double* __restrict__ u,
                                                          double beta, double* __restrict__ val, double* __restrict__ def,
                                                          unsigned long long* __restrict__ work_counter)
{
    using C = RowCfg<E, CHP>;
    constexpr int DIM = C::DIM, NSH = C::NSH, NF = C::NF, NINC = C::NINC, CH = C::CH, L = NSH * NF, NIP = C::NIP;
    constexpr bool FLOW = (STAB == STAB_FLOW);
    using R = GeoRec<E>;
    using FR = FluxRec<E, FLOW, EXACT>;
    using WS = RowWS<E, FLOW, EXACT, CHP>;
    constexpr int GS = WS::GS, RS = WS::RS;
    constexpr bool BLK = (NF == 4);                              // block-major rotated row accumulator (3-D)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    // block layout: [Ntab NIP*NSH doubles][inc table NSH*NINC ints][rottab max_cnt bytes][per warp: WS | rowacc NF*NF*max_cnt doubles]
    double* Ntab = reinterpret_cast<double*>(smem_raw);
    int32_t* inctab = reinterpret_cast<int32_t*>(smem_raw + sizeof(double) * NIP * NSH);
    uint8_t* rottab = reinterpret_cast<uint8_t*>(inctab + NSH * NINC);
    const size_t tab_bytes = rows_tab_bytes<E>(m.max_cnt);
    const size_t per_warp = (sizeof(WS) + sizeof(double) * NF * NF * m.max_cnt + 15) & ~(size_t)15;
    WS& ws = *reinterpret_cast<WS*>(smem_raw + tab_bytes + warp * per_warp);
    double* rowacc = reinterpret_cast<double*>(smem_raw + tab_bytes + warp * per_warp + sizeof(WS));
    for (int i = threadIdx.x; i < NIP * NSH; i += blockDim.x) Ntab[i] = tab::NIPSH[E][i / NSH][i % NSH];
    for (int i = threadIdx.x; i < NSH * NINC; i += blockDim.x)
        inctab[i] = tab::INC[E][i / NINC][i % NINC] | (tab::INC_SIGN[E][i / NINC][i % NINC] < 0 ? 256 : 0);
    // 3-D row accumulator: one 128-byte block (NF x NF doubles) per column slot, its eight 16-byte chunks rotated by
    // rot(slot) so that the corners of one element (slots s0 + {0,1,3,4,9,10,12,13} on a structured hex mesh) hit
    // eight different bank groups in the read-modify-write below; any mesh is handled, only the conflict rate varies.
    for (int i = threadIdx.x; i < m.max_cnt; i += blockDim.x) rottab[i] = (uint8_t)((i % 3 + 2 * ((i / 3) % 3) + 4 * (i / 9)) & 7);
    if (lane == 0) mbar_init(&ws.bar, 1);
    __syncthreads();
    unsigned phase = 0;
    const bool want_jac = p.what & (W_JAC_A | W_JAC_M), want_def = p.what & (W_DEF_A | W_DEF_M | W_RHS);
    const bool jac_a = p.what & W_JAC_A, def_a = p.what & W_DEF_A;
    // a defect-only pass needs the flux part of the records only
    const int cp_off = jac_a ? 0 : GS;
    const unsigned cp_bytes = (unsigned)sizeof(double) * (jac_a ? RS : RS - GS);
    // Jacobian accumulate: lane = (jj, k) = corner k of the jj-th of JP adjacent elements processed in parallel; the lane
    // builds the whole NF x NF block d r(node) / d u(corner k) of its element in registers. Every staged operand is
    // read by exactly one lane per use (no 4-fold broadcast of the k-indexed coefficients over the column index): the
    // shared-memory data pipe was the limiter of the lane = (k, cf) mapping (ncu: 91 % of peak wavefronts).
    constexpr int JP = 32 / NSH;
    const int jj = lane / NSH, k = lane - jj * NSH;
    const double xscale = p.laplace ? 0.0 : -1.0 * p.visc * p.rho;    // -nu rho G_kd1 n_d2 vanishes for laplace (:346-356)
    const double nurho_d = -1.0 * p.visc * p.rho;
    const double rho_f = FLOW ? p.rho : 0.0;
    (void)L; (void)rho_f;

    // dynamic work distribution: every warp atomically takes the next node of the traversal order. A static
    // grid-stride assignment lets the persistent warps drift apart, which destroys the L2 reuse of the SCVF
    // records shared by neighbouring nodes (measured: 1.75x table re-reads at 128^3 with the static loop).
    // Tickets are taken m.ticket_group nodes at a time (one atomic per group).
    (void)nwarp;
    const int TG = m.ticket_group > 0 ? m.ticket_group : 1;
    int64_t tk_base = 0;
    int tk_off = TG;
    for (;;) {
        if (tk_off == TG) {
            unsigned long long ai_u = 0;
            if (lane == 0) ai_u = atomicAdd(work_counter, (unsigned long long)TG);
            tk_base = (int64_t)__shfl_sync(0xffffffffu, ai_u, 0);
            tk_off = 0;
        }
        const int64_t ai = tk_base + tk_off++;
        if (ai >= m.n_node) break;
        const int64_t a = m.node_order ? (int64_t)m.node_order[ai] : ai;
        const int64_t q0 = m.adj_ptr[a], q1 = m.adj_ptr[a + 1];
        const int64_t b0 = m.brow[a];
        const int cnt = (int)(m.brow[a + 1] - b0);
        const int rowlen = cnt * NF;
        __syncwarp();                                            // every lane has read the previous node's rows out of rowacc (compute-sanitizer racecheck)
        if (want_jac) for (int i = lane; i < NF * rowlen; i += 32) rowacc[i] = 0.0;
        double fsum[NF], vsum = 0.0;                             // per-lane partial defect fluxes / SCV volumes
#pragma unroll
        for (int q = 0; q < NF; q++) fsum[q] = 0.0;
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
            // ---- asynchronous staging: one TMA bulk copy per incident SCVF record, all in flight at once ----
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the record area doubles as block staging (generic writes)
            if (lane == 0) mbar_arrive_expect_tx(&ws.bar, cp_bytes * (unsigned)nrec);
            __syncwarp();
            if (lane < nrec) bulk_g2s(&ws.rec[lane][cp_off], rec + gi_r * RS + cp_off, cp_bytes, &ws.bar);
            // scatter slots + the node's SCV volume in the adjacent elements (plain loads, overlapped with the copies)
            if (lane < nj) {
                const uint8_t* em = m.emap + (int64_t)ad * NSH;
                if (NSH == 8) *reinterpret_cast<uint2*>(ws.slot[lane]) = __ldg(reinterpret_cast<const uint2*>(em));
                else if (NSH == 4) *reinterpret_cast<uint32_t*>(ws.slot[lane]) = __ldg(reinterpret_cast<const uint32_t*>(em));
                else { for (int q = 0; q < NSH; q++) ws.slot[lane][q] = em[q]; }
                vsum += m.scvvol[ad];
            }
            const int sslot = __shfl_sync(0xffffffffu, la_l, 0);
            __syncwarp();
            mbar_wait(&ws.bar, phase);
            phase ^= 1u;
            if (qb == q0) self_slot = ws.slot[0][sslot];
            // ---- defect: lane r < nrec adds the signed fluxes of its record (reduced over the warp once per node) ----
            if (def_a && lane < nrec) {
                const double sgr = (ipx_r & 256) ? -1.0 : 1.0;
#pragma unroll
                for (int q = 0; q < NF; q++) fsum[q] += sgr * ws.rec[lane][GS + FR::O_F + q];
            }
            // ---- accumulate (add_jac_A_elem :317-594): block of (element j, corner k) in registers, fixed order t ----
            if (want_jac && jac_a) {
                for (int jb = 0; jb < nj; jb += JP) {
                    const int j = jb + jj;
                    const bool act = (jj < JP) && (j < nj);
                    double B[NF][NF];
#pragma unroll
                    for (int rf = 0; rf < NF; rf++)
#pragma unroll
                        for (int cf = 0; cf < NF; cf++) B[rf][cf] = 0.0;
                    if (act) {
                        double accD = 0.0;
#pragma unroll
                        for (int t = 0; t < NINC; t++) {
                            const int r = j * NINC + t;
                            const double* rc = ws.rec[r];
                            const int ipx = ws.ipx[r];
                            const double sg = (ipx & 256) ? -p.scale_a : p.scale_a;      // orientation x scaling of the A part
                            double Yn[DIM], G[DIM], A[DIM];
#pragma unroll
                            for (int d = 0; d < DIM; d++) { Yn[d] = sg * rc[d]; G[d] = rc[R::O_G + d * R::NSHP + k]; A[d] = xscale * G[d]; }
                            if constexpr (EXACT) {               // + e_k U[d1] n[d2]  (:521-549)
                                const double ek = rc[GS + FR::O_EK + k];
#pragma unroll
                                for (int d = 0; d < DIM; d++) A[d] += ek * rc[GS + FR::O_U + d];
                            }
                            const double gn = rc[R::O_GN + k];
                            const double inv = rc[GS + FR::O_INV];           // 0 without stabilisation
                            const double cK = rc[GS + FR::O_CK + k];
                            const double Nk = Ntab[(ipx & 255) * NSH + k];
                            accD += sg * (nurho_d * gn + rc[GS + FR::O_DK + k]);
#pragma unroll
                            for (int d1 = 0; d1 < DIM; d1++) {
#pragma unroll
                                for (int d2 = 0; d2 < DIM; d2++) B[d1][d2] += A[d1] * Yn[d2];
                                B[d1][DIM] += Nk * Yn[d1];                   // pressure column (:363-368)
                            }
                            // continuity row: velocity columns (:561-584), pressure column (:586-592, rho cancels)
                            if constexpr (!FLOW) {
#pragma unroll
                                for (int d2 = 0; d2 < DIM; d2++) B[DIM][d2] += cK * Yn[d2];
                            } else {
                                // sum_q sv(q,d2,k) n_q rho = ((sb_k - std.G_k) n_d2 + G_k[d2] (std.n)) inv rho
                                double sG = 0.0;
#pragma unroll
                                for (int d = 0; d < DIM; d++) sG += rc[GS + FR::O_STD + d] * G[d];
                                const double ir = inv * rho_f;
                                const double c0 = cK - sG * ir, c1 = sg * (rc[GS + FR::O_SN] * ir);
#pragma unroll
                                for (int d2 = 0; d2 < DIM; d2++) B[DIM][d2] += c0 * Yn[d2] + c1 * G[d2];
                            }
                            B[DIM][DIM] += gn * (-1.0 * (inv * sg));
                        }
#pragma unroll
                        for (int d = 0; d < DIM; d++) B[d][d] += accD;
                    }
                    // scatter into the row accumulator. The same neighbour may be a corner of several of the JP elements, so
                    // they take turns; the corners of one element are distinct nodes.
                    const int ns = (nj - jb) < JP ? (nj - jb) : JP;
                    if constexpr (BLK) {
                        // 3-D: every lane parks its block in the (consumed) record area of this group, chunk c of lane l at
                        // position (c + l) & 7 of a 128-byte line; then all 32 lanes add one element's NSH blocks at a time
                        constexpr int LPB = 32 / NSH, CPL = 8 / LPB;         // lanes per block, 16-byte chunks per lane
                        double* stage = ws.rec[jb * NINC];
                        __syncwarp();
                        if (act) {
#pragma unroll
                            for (int c = 0; c < 8; c++)
                                *reinterpret_cast<double2*>(stage + lane * 16 + (((c + lane) & 7) << 1)) =
                                    make_double2(B[c >> 1][(c & 1) * 2], B[c >> 1][(c & 1) * 2 + 1]);
                        }
                        __syncwarp();
                        const int b = lane / LPB, c0 = (lane - b * LPB) * CPL;
                        for (int s = 0; s < ns; s++) {
                            const int slot = ws.slot[jb + s][b];
                            const int ls = s * NSH + b;                      // lane that produced block b of element s
                            const double* src = stage + ls * 16;
                            double* dst = rowacc + slot * 16;
                            const int rot = rottab[slot];
#pragma unroll
                            for (int cc = 0; cc < CPL; cc++) {
                                const double2 v = *reinterpret_cast<const double2*>(src + (((c0 + cc + ls) & 7) << 1));
                                double2* q = reinterpret_cast<double2*>(dst + (((c0 + cc + rot) & 7) << 1));
                                double2 w = *q;
                                w.x += v.x; w.y += v.y;
                                *q = w;
                            }
                            __syncwarp();
                        }
                    } else {
                        for (int s = 0; s < ns; s++) {
                            if (act && jj == s) {
                                double* ra = rowacc + ws.slot[j][k] * NF;
#pragma unroll
                                for (int rf = 0; rf < NF; rf++)
#pragma unroll
                                    for (int cf = 0; cf < NF; cf++) ra[rf * rowlen + cf] += B[rf][cf];
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
        __syncwarp();
        // deterministic butterfly reductions: SCV volume of the node, defect fluxes
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        const double volsum = vsum;
        double dsum = 0.0;
        if (def_a) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int q = 0; q < NF; q++) fsum[q] += __shfl_xor_sync(0xffffffffu, fsum[q], o);
#pragma unroll
            for (int q = 0; q < NF; q++) if (lane == q) dsum = fsum[q];
        }
        if (want_jac) {
            if ((p.what & W_JAC_M) && lane < DIM) {
                const int pos = BLK ? self_slot * (NF * NF) + ((((lane * 2 + (lane >> 1)) + rottab[self_slot]) & 7) << 1) + (lane & 1)
                                    : lane * rowlen + self_slot * NF + lane;
                rowacc[pos] += p.scale_m * volsum * p.rho;
            }
            __syncwarp();
            double* out = val + b0 * (NF * NF);
            if constexpr (BLK) {
                // un-rotate: CSR row rf of the node = [slot][cf]; a lane moves the 16-byte chunk (slot, rf, column pair cp)
#pragma unroll
                for (int rf = 0; rf < NF; rf++) {
                    double2* orow = reinterpret_cast<double2*>(out + rf * rowlen);
                    for (int i = lane; i < 2 * cnt; i += 32) {
                        const int slot = i >> 1, cp = i & 1;
                        const double2 v = *reinterpret_cast<const double2*>(rowacc + slot * (NF * NF) + (((rf * 2 + cp + rottab[slot]) & 7) << 1));
                        if (beta == 0.0) __stcs(orow + i, v);
                        else { double2 o = orow[i]; o.x = beta * o.x + v.x; o.y = beta * o.y + v.y; orow[i] = o; }
                    }
                }
            } else {
                const int tot = NF * rowlen;
                if (beta == 0.0) for (int i = lane; i < tot; i += 32) __stcs(out + i, rowacc[i]);
                else for (int i = lane; i < tot; i += 32) out[i] = beta * out[i] + rowacc[i];
            }
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
