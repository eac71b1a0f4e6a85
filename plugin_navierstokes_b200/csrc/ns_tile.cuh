// ns_tile.cuh -- fused owner-computes FV1 assembly for the 3-D element types (hex, tet), diagonal stabilisation branch with
// the fixed-point Jacobian (FIELDS / no stabilisation): TWO co-resident CTAs per SM, each assembling the CSR rows of one
// TILE of grid nodes at a time (host tables: ns_patch.h). The SCVF records never leave the SM.
//
//   load  : the tile's local nodes (tile nodes + halo: every corner of an element that touches a tile node) are gathered
//           ONCE into shared memory: x y z u v w p per node (7 doubles, 96 nodes -> 5 KB instead of one 65-double column per
//           element), together with the 8-bit element -> local node table;
//   flux  : one lane per SCVF that touches a tile node (work items sorted by ip, so that a warp reads the reference tables at
//           one or two addresses): geometry, StdVel and the local gradient tensor in ONE pass over the corners, upwind (ray
//           search), diffusion length, FIELDS closure, defect fluxes. Result: a COMPRESSED 22-double record
//               [ F[4] | n[3] | alpha beta cw cpe | mv[3] | up[NSH] ]
//           from which the consumer forms   cK_k = alpha N_k + beta up_k   (continuity row, :561-584)
//                                           dK_k = cw up_k + cpe N_k      (convective diagonal, :430-468)
//                                           pK_k = dN_k . mv              (-G_k.n / diag, :586-592)
//           with the constant tables N_k(ip), dN_k(ip). 176 B instead of the 256-B lean record: a 4x4x2-node hex tile (512
//           SCVF evaluations) fits twice into the 228 KB of an SM;
//   rows  : one warp per tile node. lane = (adjacent element jj of JP in parallel, corner k) sums its NINC incident
//           records into 5 values (D, C[3], PP) and parks them in a per-warp staging row; then lane = COLUMN SLOT b of
//           the node's block row picks up the values whose scatter slot is b (byte compare on the element's slot map) and
//           accumulates them IN REGISTERS in fixed order (bitwise deterministic, no shared-memory accumulators, no merge /
//           zero-fill passes). The lane then owns the 4x4 block (node, b): out = {nu rho, 1} scale_a J0 + state part, written
//           once with 16-byte streaming stores; J0 = static Jacobian part cached per mesh (ns_split.cuh).
//
// The two CTAs of an SM drift apart, so the FP64-latency-bound flux phase of one overlaps the memory-bound rows phase of
// the other. Arithmetic restated from fv1/navier_stokes_fv1.cpp:250-778, fv1/stabilization.cpp:122-241,805-850,
// upwind.cpp:52-80,133-172,381-430,505-575, fv1/diffusion_length.h:47-198 (same formulas as ns_owner.cuh / ns_fused.cuh).
#pragma once
#include "ns_base.h"
#include "ns_patch.h"

namespace nsb {

template <int E> struct TileCfg {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP, NF = DIM + 1, NINC = ET<E>::NINC, NSIDE = ET<E>::NSIDE;
    static_assert(DIM == 3, "the tile kernel serves the 3-D element types");
    static constexpr int NT = 256, NWARP = NT / 32;
    // compressed SCVF record (doubles); every pair starts 16-byte aligned
    static constexpr int O_F = 0, O_N = 4, O_AL = 7, O_BE = 8, O_CW = 9, O_CPE = 10, O_MV = 11, O_UP = 14, RS = 14 + NSH;
    static constexpr int NDS = 7;                                   // doubles per local node: x y z u v w p
    static constexpr int TSTR = NSH * 4 + 2;                        // doubles per ip in the table [ip][k] -> (dN0, dN1, dN2, N)
    static constexpr int JP = 32 / NSH;                             // adjacent elements handled in parallel in the rows phase
    static constexpr int NV = DIM + 2;                              // D, C[DIM], PP
    static constexpr int MAXW = 512;
    static constexpr int MAXE = E == E_HEX ? 80 : 200;
    static constexpr int MAXN = E == E_HEX ? 32 : 20;
    static constexpr int MAXA = E == E_HEX ? 256 : 480;
    static constexpr int MAXLN = E == E_HEX ? 96 : 96;
    static constexpr int MAXCNT = 32;                               // one lane per column slot of a block row
    static PatchCaps caps()
    {
        PatchCaps c;
        c.max_work = MAXW; c.max_elem = MAXE; c.max_node = MAXN; c.max_adj = MAXA; c.max_lnode = MAXLN;
        if (E == E_HEX) { c.tile[0] = 4; c.tile[1] = 4; c.tile[2] = 2; }
        else { c.tile[0] = 2; c.tile[1] = 2; c.tile[2] = 2; }
        return c;
    }
};

template <int E> struct TileLayout {
    using C = TileCfg<E>;
    size_t o_rec, o_nd, o_lnode, o_ecor, o_work, o_efast, o_elid, o_stage, o_idx, o_ctr, o_adj, o_nodes, o_tab4, o_lip, o_cor, o_side, o_iptab, o_inc, total;
    __host__ __device__ TileLayout()
    {
        size_t o = 0;
        auto take = [&](size_t bytes) { const size_t at = o; o = (o + bytes + 15) & ~(size_t)15; return at; };
        o_rec = take(sizeof(double) * C::MAXW * C::RS);
        // region A (load + flux phase) ...
        const size_t a0 = o;
        o_nd = take(sizeof(double) * C::MAXLN * C::NDS);
        o_lnode = take(sizeof(int32_t) * C::MAXLN);
        o_ecor = take((size_t)C::MAXE * C::NSH);
        o_work = take(sizeof(uint32_t) * C::MAXW);
        o_efast = take(C::MAXE);
        o_elid = take(sizeof(int32_t) * C::MAXE);
        const size_t a1 = o;
        // ... aliased by region B (rows phase): per-warp staging rows [NV][32]
        o_stage = a0;
        const size_t b1 = a0 + sizeof(double) * C::NWARP * C::NV * 32;
        o = a1 > b1 ? a1 : b1;
        o = (o + 15) & ~(size_t)15;
        o_idx = take((size_t)C::NWARP * C::JP * 32);                 // per warp: slot -> corner of the JP elements of a round
        o_ctr = take(16);                                            // work counters of the flux / rows phase
        o_adj = take(sizeof(PatchAdj) * C::MAXA);
        o_nodes = take(sizeof(PatchNode) * C::MAXN);
        o_tab4 = take(sizeof(double) * C::NIP * C::TSTR);
        o_lip = take(sizeof(double) * C::NIP * 3);
        o_cor = take(sizeof(double) * 24);
        o_side = take(sizeof(int) * 24);
        o_iptab = take(sizeof(int) * C::NIP * 12);
        o_inc = take(sizeof(int) * C::NSH * C::NINC);
        total = o;
    }
};

template <int E> struct TileSmem {
    double *rec, *nd, *stage, *tab4, *lip, *cortab;
    int32_t *lnode, *elid; uint8_t *ecor, *efast, *idx; uint32_t* work; int* ctr;
    PatchAdj* adj; PatchNode* nodes;
    int *sidetab, *iptab, *inctab;
    __device__ TileSmem(unsigned char* base, const TileLayout<E>& L)
    {
        rec = reinterpret_cast<double*>(base + L.o_rec);
        nd = reinterpret_cast<double*>(base + L.o_nd);
        lnode = reinterpret_cast<int32_t*>(base + L.o_lnode);
        ecor = base + L.o_ecor;
        work = reinterpret_cast<uint32_t*>(base + L.o_work);
        efast = base + L.o_efast;
        elid = reinterpret_cast<int32_t*>(base + L.o_elid);
        stage = reinterpret_cast<double*>(base + L.o_stage);
        idx = base + L.o_idx;
        ctr = reinterpret_cast<int*>(base + L.o_ctr);
        adj = reinterpret_cast<PatchAdj*>(base + L.o_adj);
        nodes = reinterpret_cast<PatchNode*>(base + L.o_nodes);
        tab4 = reinterpret_cast<double*>(base + L.o_tab4);
        lip = reinterpret_cast<double*>(base + L.o_lip);
        cortab = reinterpret_cast<double*>(base + L.o_cor);
        sidetab = reinterpret_cast<int*>(base + L.o_side);
        iptab = reinterpret_cast<int*>(base + L.o_iptab);
        inctab = reinterpret_cast<int*>(base + L.o_inc);
    }
};

struct TileArgs {
    KParams p;
    int32_t n_tile;
    const PatchHdr* hdr; const PatchNode* nodes; const int32_t* elems; const int32_t* lnodes; const uint8_t* ecorner; const uint32_t* work; const PatchAdj* adj;
    const double* coords; const double* scvvol; const double* nodevol;
    const double* u; const double* s0; const double* s1; const double* j0;
    double beta; double* val; double* def;
    int* errflag;
    const uint8_t* elem_fast;      // hex: 1 = element is star-shaped w.r.t. its ips -> predicted-side ray search allowed; null = never
};

#ifdef __CUDACC__

template <int E> __device__ __forceinline__ void tile_stage_tables(const TileSmem<E>& S, int tid, int nthreads)
{
    using C = TileCfg<E>;
    constexpr int NSH = C::NSH, NIP = C::NIP, NINC = C::NINC;
    for (int i = tid; i < 24; i += nthreads) {
        S.cortab[i] = tab::CORNER[E][i / 3][i % 3];
        const int v = tab::SIDE[E][i / 4][i % 4];
        S.sidetab[i] = v < 0 ? 0 : v;
    }
    for (int i = tid; i < NIP * NSH * 4; i += nthreads) {
        const int ip = i / (NSH * 4), k = (i >> 2) % NSH, c = i & 3;
        S.tab4[ip * C::TSTR + k * 4 + c] = c < 3 ? tab::C_DNIP[E][ip][k][c] : tab::NIPSH[E][ip][k];
    }
    for (int i = tid; i < NIP * 3; i += nthreads) S.lip[i] = tab::LIP[E][i / 3][i % 3];
    for (int i = tid; i < NIP * 12; i += nthreads) {
        const int ip = i / 12, j = i - ip * 12;
        int v = 0;
        if (j < 2) v = tab::EDGE[E][ip][j];
        else if (j < 10) v = (j < 6) ? tab::SIDE[E][tab::SCVF_FA[E][ip]][j - 2] : tab::SIDE[E][tab::SCVF_FB[E][ip]][j - 6];   // slots 10, 11 are padding
        S.iptab[i] = v < 0 ? 0 : v;
    }
    for (int i = tid; i < NSH * NINC; i += nthreads)
        S.inctab[i] = tab::INC[E][i / NINC][i % NINC] | (tab::INC_SIGN[E][i / NINC][i % NINC] < 0 ? 256 : 0);
}

// coordinates of corner c (run-time index) of local element el
#define NSB_TX(c, d) S.nd[(int)S.ecor[el * C::NSH + (c)] * C::NDS + (d)]

// SCVF normal, ip position and |barycentre - edge midpoint|^2 (FV1Geometry; see ip_geometry in ns_fv1.cuh)
template <int E>
__device__ __forceinline__ void tile_scvf_frame(const TileSmem<E>& S, int el, int ip, const double* cen, double* n, double* xip, double& ds)
{
    using C = TileCfg<E>;
    const int* it = S.iptab + ip * 12;
    double c0[3];
#pragma unroll
    for (int d = 0; d < 3; d++) c0[d] = 0.5 * (NSB_TX(it[0], d) + NSB_TX(it[1], d));
    constexpr int NFC = (E == E_TET) ? 3 : 4;
    double c1[3] = {0, 0, 0}, c3[3] = {0, 0, 0};
#pragma unroll
    for (int q = 0; q < NFC; q++) {
        const int oa = (int)S.ecor[el * C::NSH + it[2 + q]] * C::NDS, ob = (int)S.ecor[el * C::NSH + it[6 + q]] * C::NDS;
#pragma unroll
        for (int d = 0; d < 3; d++) { c1[d] += S.nd[oa + d]; c3[d] += S.nd[ob + d]; }
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

// One side of the ray search (ElementSideRayIntersection): the triangle(s) of side `s` in reference order;
// the tests are those of side_ray_cut (ns_fv1.cuh). Returns the index of the hit triangle or -1 and leaves the un-divided
// Cramer numerators in tn / n1 / n2 / bdet.
template <int E>
__device__ __forceinline__ int tile_ray_side(const TileSmem<E>& S, int el, int s, const double* from, const double* dir, double dn2,
                                             double& tn, double& n1, double& n2, double& bdet)
{
    using C = TileCfg<E>;
    constexpr double SM = 1e-12;                                     // NSB_RAY_SMALL
    constexpr int TPS = (E == E_HEX) ? 2 : 1;
    const int o0 = (int)S.ecor[el * C::NSH + S.sidetab[s * 4]] * C::NDS;
    double ed[TPS + 1][3], r[3], q[3];
#pragma unroll
    for (int j = 0; j <= TPS; j++) {
        const int oj = (int)S.ecor[el * C::NSH + S.sidetab[s * 4 + 1 + j]] * C::NDS;
#pragma unroll
        for (int d = 0; d < 3; d++) ed[j][d] = S.nd[oj + d] - S.nd[o0 + d];
    }
#pragma unroll
    for (int d = 0; d < 3; d++) r[d] = from[d] - S.nd[o0 + d];
    cross3(q, r, dir);
    double eq[TPS + 1];
#pragma unroll
    for (int j = 0; j <= TPS; j++) eq[j] = dotv<3>(ed[j], q);
    int res = -1;
#pragma unroll
    for (int kk = 0; kk < TPS; kk++) {
        double nrm[3];
        cross3(nrm, ed[kk], ed[kk + 1]);
        const double det = -dotv<3>(dir, nrm);
        const double t_n = dotv<3>(r, nrm);
        const double b1n = eq[kk + 1], b2n = -eq[kk];
        const double sg = det > 0.0 ? 1.0 : -1.0, ad = fabs(det);
        const bool hit = res < 0 && det * det > (SM * SM) * dn2 * dotv<3>(nrm, nrm) &&
                         b1n * sg >= -SM * ad && b2n * sg >= -SM * ad && (b1n + b2n) * sg <= (1.0 + SM) * ad && t_n * sg <= 0.0;
        if (hit) { res = kk; tn = t_n; n1 = b1n; n2 = b2n; bdet = det; }
    }
    return res;
}

// Ray / element-boundary intersection. Hex with `fast`: the cut side is PREDICTED from the ray direction in reference
// coordinates (sref = J^-1 dir; going upstream the first plane xi_i in {0, 1} reached) and confirmed with the exact tests of the
// reference routine; only if the confirmation fails are the sides searched in reference order (see fused_ray_cut, ns_fused.cuh).
template <int E>
__device__ __forceinline__ bool tile_ray_cut(const TileSmem<E>& S, int el, int ip, const double* from, const double* dir, bool fast,
                                             const double* sref, int& side_out, double* gcut, double* lcut)
{
    using C = TileCfg<E>;
    constexpr int NSIDE = C::NSIDE;
    constexpr int TPS = (E == E_HEX) ? 2 : 1;
    double tn = 0.0, n1 = 0.0, n2 = 0.0, bdet = 1.0;
    const double dn2 = dotv<3>(dir, dir);
    int side = -1, tri = -1;
    if constexpr (E == E_HEX) {
        if (fast) {
            float best = -3.0e38f; int bs = -1;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const float si = (float)sref[i];
                const float xi = (float)S.lip[ip * 3 + i];
                if (si != 0.0f) {
                    const float t = si > 0.0f ? -xi / si : (1.0f - xi) / si;
                    // reference sides: zeta=0 -> 0, eta=0 -> 1, xi=1 -> 2, eta=1 -> 3, xi=0 -> 4, zeta=1 -> 5
                    const int sd = i == 0 ? (si > 0.0f ? 4 : 2) : (i == 1 ? (si > 0.0f ? 1 : 3) : (si > 0.0f ? 0 : 5));
                    if (t > best) { best = t; bs = sd; }
                }
            }
            if (bs >= 0) {
                tri = tile_ray_side<E>(S, el, bs, from, dir, dn2, tn, n1, n2, bdet);
                if (tri >= 0) side = bs;
            }
        }
    }
    if (side < 0) {
        for (int s = 0; s < NSIDE; s++) {
            tri = tile_ray_side<E>(S, el, s, from, dir, dn2, tn, n1, n2, bdet);
            if (tri >= 0) { side = s; break; }
        }
        if (side < 0) return false;
    }
    const double ibd = 1.0 / bdet;
    const double t = tn * ibd, b1 = n1 * ibd, b2 = n2 * ibd;
    const int kk = TPS == 2 ? tri : 0;
    const int p0 = S.sidetab[side * 4], p1 = S.sidetab[side * 4 + 1 + kk], p2 = S.sidetab[side * 4 + 2 + kk];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        gcut[d] = from[d] + t * dir[d];
        lcut[d] = (1 - b1 - b2) * S.cortab[p0 * 3 + d] + b1 * S.cortab[p1 * 3 + d] + b2 * S.cortab[p2 * 3 + d];
    }
    side_out = side;
    return true;
}

// upwind shapes of one ip (No / Full / Skewed / LPS) into up[NSH] (shared memory: the record's up slots); see upwind_uniform
// (ns_owner.cuh), upwind.cpp:52-80,133-172,381-430,505-575. tb = table row of the ip: tb[4 k + 3] = N_k.
template <int E>
__device__ __forceinline__ bool tile_upwind(const TileSmem<E>& S, int el, int ip, int type, bool fast, const double* sref,
                                            const double* n, const double* xip, const double* tb, const double* vel,
                                            double* up, double& len)
{
    using C = TileCfg<E>;
    constexpr int NSH = C::NSH;
    if (type == UPW_NO) {
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = tb[4 * k + 3];
        len = 1.0;
        return true;
    }
    if (type == UPW_FULL) {
        const double flux = dotv<3>(n, vel);
        const int co = flux > 0.0 ? S.iptab[ip * 12] : S.iptab[ip * 12 + 1];
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = (k == co) ? 1.0 : 0.0;
        double s = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) { const double t = xip[d] - NSB_TX(co, d); s += t * t; }
        len = sqrt(s);
        return true;
    }
#pragma unroll
    for (int k = 0; k < NSH; k++) up[k] = 0.0;
    if (sqrt(dotv<3>(vel, vel)) < 1e-14) { len = 1.0; return true; }      // upwind.cpp:407-413, 531-537
    int side = 0; double gc[3], lc[3];
    if (!tile_ray_cut<E>(S, el, ip, xip, vel, fast, sref, side, gc, lc)) { len = 1.0; return false; }
    constexpr int NSC = (E == E_TET ? 3 : 4);
    if (type == UPW_SKEWED) {                                    // GetNodeNextToCut, upwind.cpp:337-379
        double mn = 1.79769313486231570e308; int bestc = 0;
#pragma unroll
        for (int i = 0; i < NSC; i++) {
            const int co = S.sidetab[side * 4 + i];
            double dd = 0;
#pragma unroll
            for (int d = 0; d < 3; d++) { const double t = gc[d] - NSB_TX(co, d); dd += t * t; }
            if (dd < mn) { mn = dd; bestc = co; }
        }
        up[bestc] = 1.0;
        double s = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) { const double t = xip[d] - NSB_TX(bestc, d); s += t * t; }
        len = sqrt(s);
    } else {                                                     // LPS, upwind.cpp:562-573
        double Nc[NSH];
        lagrange<E>(lc, Nc);
        int mask = 0;
#pragma unroll
        for (int i = 0; i < NSC; i++) mask |= 1 << S.sidetab[side * 4 + i];
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = ((mask >> k) & 1) ? Nc[k] : 0.0;
        len = sqrt(dist2<3>(xip, gc));
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// flux phase: one SCVF (local element el, ip) -> compressed record `fr` (shared memory)
// ------------------------------------------------------------------------------------------------
template <int E, int STAB, bool TD>
__device__ __forceinline__ bool tile_scvf(const TileArgs& A, const TileSmem<E>& S, int el, int ip, double* __restrict__ fr)
{
    using C = TileCfg<E>;
    constexpr int DIM = 3, NSH = C::NSH, NIP = C::NIP, NF = 4, P = 3;
    const KParams& p = A.p;
    const bool td = TD && p.time_dep;
    const double nurho = p.visc * p.rho;
    const bool want_def = p.what & W_DEF_A, want_jac = p.what & W_JAC_A;
    bool ok = true;
    // local node offsets of the element's corners (8-bit ids, packed)
    int off[NSH];
    {
        const uint32_t* ec = reinterpret_cast<const uint32_t*>(S.ecor + el * NSH);
#pragma unroll
        for (int k = 0; k < NSH; k++) off[k] = (int)((ec[k >> 2] >> (8 * (k & 3))) & 255u) * C::NDS;
    }
    const int from = S.iptab[ip * 12], to = S.iptab[ip * 12 + 1];
    const double* tb = S.tab4 + ip * C::TSTR;
    const int64_t eg = S.elid[el];
    // SCV volumes of the two corners of the SCVF (diffusion length): issued now, used after the ray search
    double volf = 0.0, volt = 0.0;
    if (STAB != STAB_NONE) { volf = __ldg(A.scvvol + eg * NSH + from); volt = __ldg(A.scvvol + eg * NSH + to); }
    double n[DIM], xip[DIM], std[DIM], sref[DIM], mvn[DIM], dlinv = 0.0, sn, oacc = 0.0, gpn = 0.0, ds = 0.0;
    double cen[DIM];
    {
        // ---- one pass over the corners: barycentre, J^T = sum dN_k (x) x_k, StdVel (:282-293), local gradient tensor, pressure ----
        double JT[DIM][DIM], Lg[DIM][NF], pr = 0.0;
#pragma unroll
        for (int i = 0; i < DIM; i++) {
            cen[i] = 0.0; std[i] = 0.0;
#pragma unroll
            for (int j = 0; j < DIM; j++) JT[i][j] = 0.0;
#pragma unroll
            for (int q = 0; q < NF; q++) Lg[i][q] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            const double2 t0 = *reinterpret_cast<const double2*>(tb + 4 * k), t1 = *reinterpret_cast<const double2*>(tb + 4 * k + 2);
            const double dl[DIM] = {t0.x, t0.y, t1.x};
            const double Nk = t1.y;
            double xk[DIM], uk[NF];
#pragma unroll
            for (int d = 0; d < DIM; d++) xk[d] = S.nd[off[k] + d];
#pragma unroll
            for (int q = 0; q < NF; q++) uk[q] = S.nd[off[k] + 3 + q];
#pragma unroll
            for (int d = 0; d < DIM; d++) { cen[d] += xk[d]; std[d] += uk[d] * Nk; }
#pragma unroll
            for (int i = 0; i < DIM; i++) {
#pragma unroll
                for (int j = 0; j < DIM; j++) JT[i][j] += dl[i] * xk[j];
#pragma unroll
                for (int q = 0; q < NF; q++) Lg[i][q] += dl[i] * uk[q];
            }
            pr += Nk * uk[P];
        }
#pragma unroll
        for (int d = 0; d < DIM; d++) cen[d] *= (1.0 / NSH);
        double JI[DIM][DIM];
        inv_mat<DIM>(JT, JI);
        tile_scvf_frame<E>(S, el, ip, cen, n, xip, ds);
        sn = dotv<DIM>(std, n);
        // reference-space direction of the upwind ray, J^-1 StdVel (predicted-side search), and J^-T n (pressure column)
#pragma unroll
        for (int i = 0; i < DIM; i++) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) { a += JI[d][i] * std[d]; b += JI[d][i] * n[d]; }
            sref[i] = a; mvn[i] = b;
        }
        // defect (:686-776), part 1: diffusive + pressure flux and grad p . n (the local gradient tensor is mapped once by J^-T)
        if (want_def) {
            double L0[DIM] = {0.0, 0.0, 0.0};
            if (STAB != STAB_NONE && td) {                       // the closure uses solution(0) (:296, :646) and the old solution
#pragma unroll
                for (int k = 0; k < NSH; k++) {
                    const int64_t ndk = S.lnode[off[k] / C::NDS];
                    const double p0k = A.s0[ndk * NF + P];
                    double o = 0.0;
#pragma unroll
                    for (int i = 0; i < DIM; i++) L0[i] += tb[4 * k + i] * p0k;
#pragma unroll
                    for (int d = 0; d < DIM; d++) o += A.s1[ndk * NF + d] * n[d];
                    oacc += tb[4 * k + 3] * o;
                }
            }
            double gv[DIM][DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                double sp_ = 0.0;
#pragma unroll
                for (int i = 0; i < DIM; i++) sp_ += JI[d][i] * (td ? L0[i] : Lg[i][P]);
                gpn += sp_ * n[d];
#pragma unroll
                for (int q = 0; q < DIM; q++) {
                    double sv_ = 0.0;
#pragma unroll
                    for (int i = 0; i < DIM; i++) sv_ += JI[d][i] * Lg[i][q];
                    gv[q][d] = sv_;
                }
            }
#pragma unroll
            for (int d1 = 0; d1 < DIM; d1++) {
                double df = 0.0;
#pragma unroll
                for (int d2 = 0; d2 < DIM; d2++) df += gv[d1][d2] * n[d2];
                if (!p.laplace) {
#pragma unroll
                    for (int d2 = 0; d2 < DIM; d2++) df += gv[d2][d1] * n[d2];
                }
                fr[C::O_F + d1] = df * (-1.0) * nurho + pr * n[d1];     // the convective part is added below
            }
        }
    }
    // COR diffusion length: element-wide statistics of the SCVF normals (diffusion_length.h:139-172)
    if (STAB != STAB_NONE) {
        double cmn = 0.0, cav = 0.0, cmd = 0.0;
        if (p.diff_len == DIFF_COR) {
            cmn = 1.79769313486231570e308; cmd = 1.79769313486231570e308;
            for (int i = 0; i < NIP; i++) {
                double nn_[DIM], xx_[DIM], dsi;
                tile_scvf_frame<E>(S, el, i, cen, nn_, xx_, dsi);
                const double q = dotv<DIM>(nn_, nn_);
                if (q < cmn) cmn = q;
                cav += q;
                if (dsi < cmd) cmd = dsi;
            }
            cav /= NIP;
        }
        dlinv = diff_len_sq_inv<DIM>(p.diff_len, dotv<DIM>(n, n), volf, volt, ds, cmn, cav, cmd);
    }
    const double prod = sn * p.rho;
    // ---- the stabilisation's upwind (= the convective one on this path); its shapes live in the record ----
    double* up = fr + C::O_UP;
    double uplen = 1.0;
    if (p.stokes) {
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = 0.0;
    } else {
        const bool fast = S.efast[el] != 0;
        ok &= tile_upwind<E>(S, el, ip, p.upw_stab, fast, sref, n, xip, tb, std, up, uplen);
    }
    // diagonal of the ip system and numerators sb_k = qa N_k + qb up_k (stabilization.cpp:166-236)
    double inv = 0.0, qa = 0.0, qb = 0.0;
    if (STAB != STAB_NONE) {
        qa = p.visc * dlinv;
        if (!p.stokes) qb = sqrt(dotv<DIM>(std, std)) / uplen;
        double diag = qa;
        if (td) diag += 1.0 / p.dt;
        if (!p.stokes) diag += qb;
        inv = 1.0 / diag;
    }
    const double ci = inv * p.rho;
    double U[DIM] = {0.0, 0.0, 0.0};                              // upwind_vel (upwind_interface.h:334-358)
    double acc = 0.0;                                            // time-dependent closure sum  sum_k sb_k (s0_k . n)
    if (!p.stokes || (STAB != STAB_NONE && want_def && td)) {
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            const double upk = up[k];
#pragma unroll
            for (int d = 0; d < DIM; d++) U[d] += upk * S.nd[off[k] + 3 + d];
            if (STAB != STAB_NONE && want_def && td) {
                const int64_t ndk = S.lnode[off[k] / C::NDS];
                double sk = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) sk += A.s0[ndk * NF + d] * n[d];
                acc += (qa * tb[4 * k + 3] + qb * upk) * sk;
            }
        }
    }
    // defect, part 2: continuity flux (stab_vel . n) rho
    if (want_def) {
        double cont;
        if (STAB == STAB_NONE) cont = sn * p.rho;
        else {
            if (!td) {
                // stationary FIELDS: sum_k (qa N_k + qb up_k) (u_k . n) = (qa StdVel + qb U_up) . n
#pragma unroll
                for (int d = 0; d < DIM; d++) acc += (qa * std[d] + qb * U[d]) * n[d];
            }
            acc -= gpn * p.inv_rho;
            if (p.has_source) {
#pragma unroll
                for (int d = 0; d < DIM; d++) acc += p.src[d] * n[d];
            }
            if (td) acc += oacc / p.dt;
            cont = acc * ci;
        }
        fr[C::O_F + P] = cont;
    }
    // ---- Peclet blend (:871-892), convective flux ----
    double w = 1.0;
    if (!p.stokes) {
        if (p.peclet) {
            double dd = 0;
#pragma unroll
            for (int d = 0; d < DIM; d++) { const double t = NSB_TX(to, d) - NSB_TX(from, d); dd += t * t; }
            const double Pe = sn / dotv<DIM>(n, n) * sqrt(dd) / p.visc;
            const double Pe2 = Pe * Pe;
            w = Pe2 / (5.0 + Pe2);
#pragma unroll
            for (int d = 0; d < DIM; d++) U[d] = w * U[d] + (1.0 - w) * std[d];
        }
        if (want_def) {
#pragma unroll
            for (int d = 0; d < DIM; d++) fr[C::O_F + d] += U[d] * prod;
        }
    }
    // ---- Jacobian coefficients (compressed): cK_k = alpha N_k + beta up_k, dK_k = cw up_k + cpe N_k, pK_k = dN_k . mv ----
    if (want_jac) {
        const double alpha = (STAB == STAB_NONE) ? p.rho : qa * ci, beta = (STAB == STAB_NONE) ? 0.0 : qb * ci;
        const double cw = p.stokes ? 0.0 : prod * w, cpe = (p.stokes || !p.peclet) ? 0.0 : prod * (1.0 - w);
        const double mi = -1.0 * inv;
        *reinterpret_cast<double2*>(fr + C::O_N) = make_double2(n[0], n[1]);
        *reinterpret_cast<double2*>(fr + C::O_N + 2) = make_double2(n[2], alpha);
        *reinterpret_cast<double2*>(fr + C::O_BE) = make_double2(beta, cw);
        *reinterpret_cast<double2*>(fr + C::O_CPE) = make_double2(cpe, mvn[0] * mi);
        *reinterpret_cast<double2*>(fr + C::O_MV + 1) = make_double2(mvn[1] * mi, mvn[2] * mi);
    }
    return ok;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int E, int STAB, bool TD>
__global__ void __launch_bounds__(TileCfg<E>::NT, 2) fv1_tile_kernel(const TileArgs A)
{
    using C = TileCfg<E>;
    constexpr int NSH = C::NSH, NF = C::NF, NINC = C::NINC, JP = C::JP, NV = C::NV, RS = C::RS, NT = C::NT, NWARP = C::NWARP, DIM = 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TileLayout<E> L;
    const TileSmem<E> S(smem_raw, L);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    tile_stage_tables<E>(S, tid, NT);
    const KParams& p = A.p;
    const int what = p.what;
    const bool want_jac = what & (W_JAC_A | W_JAC_M), want_def = what & (W_DEF_A | W_DEF_M | W_RHS);
    const bool flux_needed = what & (W_JAC_A | W_DEF_A), jac_a = what & W_JAC_A, def_a = what & W_DEF_A;
    const int jj = lane / NSH, k = lane - jj * NSH;
    const double sa = p.scale_a;
    const double s_visc = p.visc * p.rho * p.scale_a, s_pres = p.scale_a;
    const double beta = A.beta;
    double* stg = S.stage + warp * (NV * 32);
    uint8_t* idx = S.idx + warp * (JP * 32);

    for (int ti = blockIdx.x; ti < A.n_tile; ti += gridDim.x) {
        PatchHdr H;
        {
            const int4* hp = reinterpret_cast<const int4*>(A.hdr + ti);
            const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2);
            H.node0 = h0.x; H.n_node = h0.y; H.elem0 = h0.z; H.n_elem = h0.w; H.work0 = h1.x; H.n_work = h1.y; H.adj0 = h1.z; H.n_adj = h1.w;
            H.lnode0 = h2.x; H.n_lnode = h2.y; H.pad0 = 0; H.pad1 = 0;
        }
        // ---- load: local nodes (coordinates + unknowns), element -> local node table, work list, node / adjacency tables ----
        for (int i = tid; i < H.n_lnode; i += NT) {
            const int64_t g = __ldg(A.lnodes + H.lnode0 + i);
            S.lnode[i] = (int32_t)g;
            const double2 a = __ldg(reinterpret_cast<const double2*>(A.u + g * NF)), b = __ldg(reinterpret_cast<const double2*>(A.u + g * NF) + 1);
            double* q = S.nd + i * C::NDS;
            q[0] = __ldg(A.coords + g * 3); q[1] = __ldg(A.coords + g * 3 + 1); q[2] = __ldg(A.coords + g * 3 + 2);
            q[3] = a.x; q[4] = a.y; q[5] = b.x; q[6] = b.y;
        }
        {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(A.ecorner + (int64_t)H.elem0 * NSH);
            uint32_t* dst = reinterpret_cast<uint32_t*>(S.ecor);
            for (int i = tid; i < (H.n_elem * NSH) >> 2; i += NT) dst[i] = __ldg(src + i);
        }
        for (int i = tid; i < H.n_work; i += NT) S.work[i] = __ldg(A.work + H.work0 + i);
        if (tid == 0) { S.ctr[0] = 0; S.ctr[1] = 0; }
        for (int i = tid; i < H.n_elem; i += NT) {
            const int32_t e = __ldg(A.elems + H.elem0 + i);
            S.elid[i] = e;
            S.efast[i] = A.elem_fast ? A.elem_fast[e] : (uint8_t)0;
        }
        {
            const int4* src = reinterpret_cast<const int4*>(A.adj + H.adj0);
            int4* dst = reinterpret_cast<int4*>(S.adj);
            for (int i = tid; i < H.n_adj; i += NT) dst[i] = __ldg(src + i);
            const int4* srcn = reinterpret_cast<const int4*>(A.nodes + H.node0);
            int4* dstn = reinterpret_cast<int4*>(S.nodes);
            for (int i = tid; i < H.n_node; i += NT) dstn[i] = __ldg(srcn + i);
        }
        __syncthreads();
        // the J0 rows of the tile nodes are read in the rows phase: pull them into L2 now
        if (jac_a) {
            for (int i = tid; i < H.n_node * 21; i += NT) {
                const int nl = i / 21, li = i - nl * 21;
                const PatchNode& Nd = S.nodes[nl];
                if (li * 16 < (int)Nd.cnt * (DIM * NF)) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(A.j0 + Nd.b0 * (DIM * NF) + li * 16));
            }
        }
        // ... and the tables of the next tile
        {
            const int tn = ti + (int)gridDim.x;
            if (tn < A.n_tile && tid < 32) {
                const int4* hp = reinterpret_cast<const int4*>(A.hdr + tn);
                const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2);
                // lnodes, ecorner, work, elems, adj, nodes: one 128-byte line per lane and table
                auto pf = [&](const void* base, size_t bytes) { for (size_t o = (size_t)lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(reinterpret_cast<const char*>(base) + o)); };
                pf(A.lnodes + h2.x, (size_t)h2.y * 4); pf(A.ecorner + (int64_t)h0.z * NSH, (size_t)h0.w * NSH); pf(A.work + h1.x, (size_t)h1.y * 4);
                pf(A.elems + h0.z, (size_t)h0.w * 4); pf(A.adj + h1.z, (size_t)h1.w * 16); pf(A.nodes + h0.x, (size_t)h0.y * 16);
            }
        }
        // ---- flux: the warps take 32 work items at a time ----
        if (flux_needed) {
            bool ok = true;
            for (;;) {
                int w0 = 0;
                if (lane == 0) w0 = atomicAdd(&S.ctr[0], 32);
                w0 = __shfl_sync(0xffffffffu, w0, 0);
                if (w0 >= H.n_work) break;
                const int w = w0 + lane;
                if (w < H.n_work) {
                    const uint32_t wi = S.work[w];
                    const int el = wi & 255, ip = (wi >> 8) & 15, slot = wi >> 12;
                    ok &= tile_scvf<E, STAB, TD>(A, S, el, ip, S.rec + slot * RS);
                }
            }
            if (!ok) atomicExch(A.errflag, 1);
        }
        __syncthreads();                                         // records complete; region A is dead, the staging rows may use it
        // ---- rows: one warp per tile node ----
        for (;;) {
            int nl = 0;
            if (lane == 0) nl = atomicAdd(&S.ctr[1], 1);
            nl = __shfl_sync(0xffffffffu, nl, 0);
            if (nl >= H.n_node) break;
            const PatchNode Nd = S.nodes[nl];
            const int nadj = Nd.adj_cnt, cnt = Nd.cnt;
            const bool own = lane < cnt;                         // lane = column slot b of the block row
            const double2* j0g = reinterpret_cast<const double2*>(A.j0 + Nd.b0 * (DIM * NF));
            double2 jv[DIM][2];
            if (jac_a && own) {
#pragma unroll
                for (int rf = 0; rf < DIM; rf++) { jv[rf][0] = __ldcs(j0g + (rf * cnt + lane) * 2); jv[rf][1] = __ldcs(j0g + (rf * cnt + lane) * 2 + 1); }
            }
            double aD = 0.0, aC[DIM] = {0.0, 0.0, 0.0}, aP = 0.0, fs = 0.0;
            const int self = nadj > 0 ? (int)S.adj[Nd.adj_off].self : -1;
            if (flux_needed) {
                for (int j0i = 0; j0i < nadj; j0i += JP) {
                    const int j = j0i + jj;
                    const bool on = jj < JP && j < nadj;
                    double D = 0.0, PP = 0.0, Cn[DIM] = {0.0, 0.0, 0.0};
                    uint32_t em = 0;
                    if (on) {
                        const PatchAdj& a = S.adj[Nd.adj_off + j];
                        const uint32_t lab = a.la;
                        em = reinterpret_cast<const uint32_t*>(a.emap)[k >> 2];
                        const int la = lab & 7;
#pragma unroll
                        for (int t = 0; t < NINC; t++) {
                            const double* rc = S.rec + (int)a.slot[t] * RS;
                            const bool neg = (lab >> (4 + t)) & 1;
                            if (def_a && k < NF) { const double f = rc[C::O_F + k]; fs += neg ? -f : f; }
                            if (jac_a) {
                                const int ipt = S.inctab[la * NINC + t] & 255;
                                const double* tq = S.tab4 + ipt * C::TSTR + 4 * k;
                                const double2 t0 = *reinterpret_cast<const double2*>(tq), t1 = *reinterpret_cast<const double2*>(tq + 2);
                                const double2 r2 = *reinterpret_cast<const double2*>(rc + 4), r3 = *reinterpret_cast<const double2*>(rc + 6);
                                const double2 r4 = *reinterpret_cast<const double2*>(rc + 8), r5 = *reinterpret_cast<const double2*>(rc + 10);
                                const double2 r6 = *reinterpret_cast<const double2*>(rc + 12);
                                const double upk = rc[C::O_UP + k];
                                const double sg = neg ? -sa : sa;
                                const double cK = r3.y * t1.y + r4.x * upk;           // alpha N_k + beta up_k
                                const double dK = upk * r4.y + r5.x * t1.y;           // cw up_k + cpe N_k
                                double pK = t0.x * r5.y;                               // dN_k . mv
                                pK += t0.y * r6.x; pK += t1.x * r6.y;
                                D += sg * dK; PP += sg * pK;
                                const double wv = sg * cK;
                                Cn[0] += wv * r2.x; Cn[1] += wv * r2.y; Cn[2] += wv * r3.x;
                            }
                        }
                    }
                    if (jac_a) {
                        __syncwarp();                            // the owners have consumed the previous round
#pragma unroll
                        for (int i = lane; i < JP * 8; i += 32) reinterpret_cast<uint32_t*>(idx)[i] = 0xffffffffu;
                        __syncwarp();
                        if (on) idx[jj * 32 + ((em >> (8 * (k & 3))) & 255u)] = (uint8_t)k;
                        stg[lane] = D; stg[32 + lane] = Cn[0]; stg[64 + lane] = Cn[1]; stg[96 + lane] = Cn[2]; stg[128 + lane] = PP;
                        __syncwarp();
#pragma unroll
                        for (int j2 = 0; j2 < JP; j2++) {
                            const int kk = idx[j2 * 32 + lane];
                            if (kk != 255) {
                                const int src = j2 * NSH + kk;
                                aD += stg[src]; aC[0] += stg[32 + src]; aC[1] += stg[64 + src]; aC[2] += stg[96 + src]; aP += stg[128 + src];
                            }
                        }
                    }
                }
            }
            if ((what & W_JAC_M) && lane == self) aD += p.scale_m * A.nodevol[Nd.node] * p.rho;   // lumped mass (add_jac_M_elem :781-808)
            if (want_jac && own) {
                double2* o2 = reinterpret_cast<double2*>(A.val + Nd.b0 * (NF * NF));
#pragma unroll
                for (int rf = 0; rf < DIM; rf++) {
                    double2 v0 = make_double2(0.0, 0.0), v1 = make_double2(0.0, 0.0);
                    if (jac_a) { v0.x = jv[rf][0].x * s_visc; v0.y = jv[rf][0].y * s_visc; v1.x = jv[rf][1].x * s_visc; v1.y = jv[rf][1].y * s_pres; }
                    if (rf == 0) v0.x += aD; else if (rf == 1) v0.y += aD; else v1.x += aD;
                    double2* o = o2 + (rf * cnt + lane) * 2;
                    if (beta == 0.0) { __stcs(o, v0); __stcs(o + 1, v1); }
                    else { double2 x0 = o[0], x1 = o[1]; x0.x = beta * x0.x + v0.x; x0.y = beta * x0.y + v0.y; x1.x = beta * x1.x + v1.x; x1.y = beta * x1.y + v1.y; o[0] = x0; o[1] = x1; }
                }
                double2* o = o2 + (DIM * cnt + lane) * 2;
                const double2 v0 = make_double2(aC[0], aC[1]), v1 = make_double2(aC[2], aP);
                if (beta == 0.0) { __stcs(o, v0); __stcs(o + 1, v1); }
                else { double2 x0 = o[0], x1 = o[1]; x0.x = beta * x0.x + v0.x; x0.y = beta * x0.y + v0.y; x1.x = beta * x1.x + v1.x; x1.y = beta * x1.y + v1.y; o[0] = x0; o[1] = x1; }
            }
            if (want_def) {
                // defect entry (node, component q = lane < NF): fixed-order sum over the JP lane groups
                double dsum = 0.0;
#pragma unroll
                for (int j2 = 0; j2 < JP; j2++) dsum += __shfl_sync(0xffffffffu, fs, j2 * NSH + (lane < NF ? lane : 0));
                if (lane < NF) {
                    const int64_t a = Nd.node;
                    double d = def_a ? dsum : 0.0;
                    const bool need_vol = ((what & W_RHS) && p.has_source) || (what & W_DEF_M);
                    const double vol = (need_vol && lane < DIM) ? A.nodevol[a] : 0.0;
                    if ((what & W_RHS) && p.has_source && lane < DIM) d -= p.src[lane] * vol * p.rho;
                    d *= p.scale_a;
                    if ((what & W_DEF_M) && lane < DIM) d += p.scale_m * A.u[a * NF + lane] * vol * p.rho;
                    double* q = A.def + a * NF + lane;
                    *q = (beta == 0.0) ? d : beta * (*q) + d;
                }
            }
        }
        __syncthreads();                                         // records and staging rows consumed
    }
}
#undef NSB_TX
#endif  // __CUDACC__

}  // namespace nsb
