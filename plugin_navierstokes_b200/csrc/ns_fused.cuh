// ns_fused.cuh -- FUSED owner-computes FV1 assembly (diagonal stabilisation branch, fixed-point Jacobian): one persistent
// CTA per SM assembles the CSR rows of one PATCH of grid nodes at a time (ns_patch.h). The SCVF records never leave the
// SM: per patch
//   load  : the patch tables and the corner data (coordinates, unknowns, SCV volumes) of the patch's elements are
//           gathered into shared-memory columns,
//   flux  : one lane per SCVF that touches a patch node evaluates geometry -> StdVel -> upwind (ray search) ->
//           diffusion length -> FIELDS / no-stabilisation closure -> defect fluxes -> Jacobian coefficients and writes the
//           lean record [F | n | cK | dK | pK] into its shared-memory slot (what fv1_flux_kernel<LEAN> wrote to HBM),
//   rows  : lane = (patch node, corner k): for every adjacent element the lane sums its NINC incident records and adds
//           5 values into the node's per-slot accumulators (the corners of one element are distinct nodes: conflict-free),
//           then the node's NF rows are written ONCE, coalesced: out = {nu rho, 1} scale_a J0 + state part (+ lumped mass).
// Compared with the two-kernel split path (ns_split.cuh) the 3 KB/element record round trip through HBM and the TMA
// staging / ticket / header machinery of the rows kernel are gone; the price is that SCVFs on patch boundaries are
// evaluated by both patches (see PatchPlan::n_scvf_evals).
//
// Everything that computes is written as NSB_HD "lane functions" of (thread index, shared-memory view): the CUDA kernel
// calls them between barriers, and tests/cpp/emu_fused.cpp runs the very same functions thread by thread on the CPU so
// that the arithmetic and the patch tables are parity-tested against the oracle without a GPU.
//
// Arithmetic restated from fv1/navier_stokes_fv1.cpp:250-778, fv1/stabilization.cpp:122-241,805-850,
// upwind.cpp:52-80,133-172,381-430,505-575, fv1/diffusion_length.h:47-198 (same formulas as ns_owner.cuh).
#pragma once
#include "ns_base.h"
#include "ns_patch.h"

namespace nsb {

template <int E> struct FusedCfg {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP, NF = DIM + 1, NINC = ET<E>::NINC, NSIDE = ET<E>::NSIDE;
    // threads per CTA. 2-D: 256 threads and half-size patches so that TWO CTAs are resident per SM (about 82 KB of shared memory
    // each): the phases of a patch are separated by block-wide barriers (flux -> accumulate -> output, 47 % of the stall samples
    // with one 512-thread CTA per SM), a second CTA fills the SM while the first one waits
    static constexpr int NT = ET<E>::DIM == 2 ? 256 : 512;
    static constexpr int CTAS = ET<E>::DIM == 2 ? 2 : 1;           // resident CTAs per SM the kernel is bounded for
    static constexpr int RS = LeanRec<E>::SZ;
    static constexpr int RSTR = RS + 1;                             // odd record stride in shared memory: the lanes of a half-warp (adjacent slots, same field) hit distinct banks
    static constexpr int GEO = 16;                                  // doubles per static SCVF geometry record: n[3] xip[3] JI[9] 1/L_d^2 (2-D: same slots, unused ones 0)
    static constexpr int NV = DIM + 2;                              // accumulated values per slot: D, C[DIM], PP
    static constexpr int DSTR = NSH * DIM + 1, NSTR = NSH + 1;      // odd strides of the per-ip tables
    static constexpr int NCOL = NSH * DIM + NSH * NF + NSH;         // doubles per element: corner coordinates, unknowns, SCV volumes
    static constexpr int CSTR = NCOL | 1;                           // element-major rows, odd stride: the lanes of a warp (same element, different
                                                                    // corners / different elements, same corner) hit different banks
    static constexpr int NPW = 32 / NSH;                            // rows phase, accumulation: lane = (one of NPW patch nodes, corner k)
    static constexpr int NWARP = NT / 32;
    static constexpr int JIT = 2;                                   // rows phase, output: rounds of 32 words per matrix row held in registers (J0 prefetch)
    // capacities of one patch (tile = node-box the grid is binned into; see ns_patch.h)
    static constexpr int MAXW = ET<E>::DIM == 2 ? 256 : 512;
    static constexpr int MAXE = E == E_HEX ? 80 : (E == E_TET ? 160 : (E == E_QUAD ? 64 : 112));
    static constexpr int MAXN = E == E_HEX ? 32 : (E == E_TET ? 24 : (E == E_QUAD ? 48 : 32));
    static constexpr int MAXA = E == E_HEX ? 288 : (E == E_TET ? 480 : 224);
    static PatchCaps caps()
    {
        PatchCaps c;
        c.max_work = MAXW; c.max_elem = MAXE; c.max_node = MAXN; c.max_adj = MAXA;
        if (E == E_HEX) { c.tile[0] = 4; c.tile[1] = 4; c.tile[2] = 2; }
        else if (E == E_TET) { c.tile[0] = 2; c.tile[1] = 2; c.tile[2] = 2; }
        else if (E == E_QUAD) { c.tile[0] = 8; c.tile[1] = 6; c.tile[2] = 1; }      // 48 nodes, 63 quadrilaterals, 252 SCVFs
        else { c.tile[0] = 6; c.tile[1] = 5; c.tile[2] = 1; }                        // 30 nodes, 84 triangles, 252 SCVFs
        return c;
    }
};

__host__ __device__ constexpr int fused_cnt_pad(int max_cnt) { return (max_cnt + 1) & ~1; }

// byte offsets of the shared-memory regions
template <int E> struct FusedLayout {
    using C = FusedCfg<E>;
    size_t o_rec, o_cols, o_acc, o_dnt, o_nt, o_lip, o_cor, o_side, o_iptab, o_inc, o_work, o_adj, o_nodes, o_efast, o_elid, o_misc, total;
    int cntp;
    __host__ __device__ explicit FusedLayout(int max_cnt)
    {
        cntp = fused_cnt_pad(max_cnt);
        size_t o = 0;
        auto take = [&](size_t bytes) { const size_t at = o; o = (o + bytes + 15) & ~(size_t)15; return at; };
        o_rec = take(sizeof(double) * C::MAXW * C::RSTR);
        o_cols = take(sizeof(double) * C::CSTR * C::MAXE);
        o_acc = take(sizeof(double) * (size_t)C::MAXN * C::NV * cntp);     // one accumulator block per patch node
        o_dnt = take(sizeof(double) * C::NIP * C::DSTR);
        o_nt = take(sizeof(double) * C::NIP * C::NSTR);
        o_lip = take(sizeof(double) * C::NIP * 3);
        o_cor = take(sizeof(double) * 24);
        o_side = take(sizeof(int) * 24);
        o_iptab = take(sizeof(int) * C::NIP * 12);
        o_inc = take(sizeof(int) * C::NSH * C::NINC);
        o_work = take(sizeof(uint32_t) * C::MAXW);
        o_adj = take(2 * sizeof(PatchAdj) * C::MAXA);               // ping-pong: the next patch's tables arrive while the rows are written
        o_nodes = take(2 * sizeof(PatchNode) * C::MAXN);
        o_efast = take(C::MAXE);
        o_elid = take(sizeof(int32_t) * C::MAXE);
        o_misc = take(64);
        total = o;
    }
};

// view of the CTA's shared memory (or of the emulator's buffer)
template <int E> struct FusedSmem {
    double *rec, *xs, *us, *vs, *acc, *dnt, *Nt, *lip, *cortab;
    int *sidetab, *iptab, *inctab;
    uint32_t* work; PatchAdj* adj0; PatchNode* nodes0; uint8_t* efast; int32_t* elem_id; int* misc;
    int cntp;
    NSB_HD FusedSmem(unsigned char* base, const FusedLayout<E>& L)
    {
        using C = FusedCfg<E>;
        rec = reinterpret_cast<double*>(base + L.o_rec);
        xs = reinterpret_cast<double*>(base + L.o_cols);
        us = xs + C::NSH * C::DIM;
        vs = us + C::NSH * C::NF;
        acc = reinterpret_cast<double*>(base + L.o_acc);            // [patch node][NV][cntp]
        dnt = reinterpret_cast<double*>(base + L.o_dnt);
        Nt = reinterpret_cast<double*>(base + L.o_nt);
        lip = reinterpret_cast<double*>(base + L.o_lip);
        cortab = reinterpret_cast<double*>(base + L.o_cor);
        sidetab = reinterpret_cast<int*>(base + L.o_side);
        iptab = reinterpret_cast<int*>(base + L.o_iptab);
        inctab = reinterpret_cast<int*>(base + L.o_inc);
        work = reinterpret_cast<uint32_t*>(base + L.o_work);
        adj0 = reinterpret_cast<PatchAdj*>(base + L.o_adj);           // two buffers each (ping-pong by patch parity): see FusedTab
        nodes0 = reinterpret_cast<PatchNode*>(base + L.o_nodes);
        elem_id = reinterpret_cast<int32_t*>(base + L.o_elid);
        efast = base + L.o_efast;
        misc = reinterpret_cast<int*>(base + L.o_misc);
        cntp = L.cntp;
    }
};

// the adjacency / node tables of the current patch (buffer `par` of the ping-pong pair)
template <int E> struct FusedTab {
    const PatchAdj* adj; const PatchNode* nodes;
    NSB_HD FusedTab(const FusedSmem<E>& S, int par) : adj(S.adj0 + par * FusedCfg<E>::MAXA), nodes(S.nodes0 + par * FusedCfg<E>::MAXN) {}
};

// global-memory arguments
struct FusedArgs {
    KParams p;
    int32_t n_patch;
    const PatchHdr* hdr; const PatchNode* nodes; const int32_t* elems; const int32_t* pconn; const uint32_t* work; const PatchAdj* adj;
    const double* coords; const double* scvvol; const double* nodevol;
    const double* u; const double* s0; const double* s1; const double* j0;
    double beta; double* val; double* def;
    int* errflag;
    const double* geo;             // static SCVF geometry records in work-item order ([scvf_evals][16]) or null: geometry on the fly
    const uint8_t* elem_fast;      // hex: 1 = element is star-shaped w.r.t. its ips -> predicted-side ray search allowed; null = never
    int max_adj;                   // longest adjacency list of a node
};

// reference tables -> shared memory (block-wide, once per kernel). `tid` strides over `nthreads`.
template <int E> NSB_DEV void fused_stage_tables(const FusedSmem<E>& S, int tid, int nthreads)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NSH = C::NSH, NIP = C::NIP, NINC = C::NINC;
    for (int i = tid; i < 24; i += nthreads) {
        S.cortab[i] = tab::CORNER[E][i / 3][i % 3];
        const int v = tab::SIDE[E][i / 4][i % 4];
        S.sidetab[i] = v < 0 ? 0 : v;
    }
    for (int i = tid; i < NIP * NSH * DIM; i += nthreads) S.dnt[(i / (NSH * DIM)) * C::DSTR + i % (NSH * DIM)] = tab::C_DNIP[E][i / (NSH * DIM)][(i / DIM) % NSH][i % DIM];
    for (int i = tid; i < NIP * NSH; i += nthreads) S.Nt[(i / NSH) * C::NSTR + i % NSH] = tab::NIPSH[E][i / NSH][i % NSH];
    for (int i = tid; i < NIP * 3; i += nthreads) S.lip[i] = tab::LIP[E][i / 3][i % 3];
    for (int i = tid; i < NIP * 12; i += nthreads) {
        const int ip = i / 12, j = i - ip * 12;
        int v = 0;
        if (j < 2) v = tab::EDGE[E][ip][j];
        else if (DIM == 3 && j < 10) v = (j < 6) ? tab::SIDE[E][tab::SCVF_FA[E][ip]][j - 2] : tab::SIDE[E][tab::SCVF_FB[E][ip]][j - 6];   // slots 10, 11 are padding
        S.iptab[i] = v < 0 ? 0 : v;
    }
    for (int i = tid; i < NSH * NINC; i += nthreads)
        S.inctab[i] = tab::INC[E][i / NINC][i % NINC] | (tab::INC_SIGN[E][i / NINC][i % NINC] < 0 ? 256 : 0);
}

// element column `el`, entry i
#define NSB_FCOL(base, i) (base)[(i) + el * C::CSTR]

// FV1Geometry of one SCVF from the element's shared column (see ip_geometry in ns_fv1.cuh)
template <int E>
NSB_HD void fused_ip_geometry(const FusedSmem<E>& S, int el, int ip, const double* cen, double* n, double* xip, double& ds,
                              bool wantJ, double (*JI)[ET<E>::DIM])
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NSH = C::NSH;
    const int f = S.iptab[ip * 12], t = S.iptab[ip * 12 + 1];
    double c0[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) c0[d] = 0.5 * (NSB_FCOL(S.xs, f * DIM + d) + NSB_FCOL(S.xs, t * DIM + d));
    if constexpr (DIM == 2) {
        n[0] = cen[1] - c0[1]; n[1] = -(cen[0] - c0[0]);
        xip[0] = 0.5 * (c0[0] + cen[0]); xip[1] = 0.5 * (c0[1] + cen[1]);
        ds = 0.0;
    } else {
        constexpr int NFC = (E == E_TET) ? 3 : 4;
        double c1[3] = {0, 0, 0}, c3[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < NFC; q++) {
            const int ka = S.iptab[ip * 12 + 2 + q], kb = S.iptab[ip * 12 + 6 + q];
#pragma unroll
            for (int d = 0; d < 3; d++) { c1[d] += NSB_FCOL(S.xs, ka * 3 + d); c3[d] += NSB_FCOL(S.xs, kb * 3 + d); }
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
        const double* dn = S.dnt + ip * C::DSTR;
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            double dk[DIM];
#pragma unroll
            for (int i = 0; i < DIM; i++) dk[i] = dn[k * DIM + i];
#pragma unroll
            for (int j = 0; j < DIM; j++) {
                const double xkj = NSB_FCOL(S.xs, k * DIM + j);
#pragma unroll
                for (int i = 0; i < DIM; i++) JT[i][j] += dk[i] * xkj;
            }
        }
        inv_mat<DIM>(JT, JI);
    }
}

// One side of the ray search (ElementSideRayIntersection): segment (2-D) or the triangle(s) of side `s`
// in reference order; the tests are those of side_ray_cut (ns_fv1.cuh). Returns the index of the hit triangle or -1 and
// leaves the un-divided Cramer numerators in tn / n1 / n2 / bdet.
template <int E>
NSB_HD int fused_ray_side(const FusedSmem<E>& S, int el, int s, const double* from, const double* dir, double dn2,
                          double& tn, double& n1, double& n2, double& bdet)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM;
    constexpr double SM = 1e-12;                                     // NSB_RAY_SMALL
    if constexpr (DIM == 2) {
        const int p0 = S.sidetab[s * 4], p1 = S.sidetab[s * 4 + 1];
        const double x0 = NSB_FCOL(S.xs, p0 * 2), y0 = NSB_FCOL(S.xs, p0 * 2 + 1);
        const double ex = NSB_FCOL(S.xs, p1 * 2) - x0, ey = NSB_FCOL(S.xs, p1 * 2 + 1) - y0;
        const double det = dir[0] * (-ey) + dir[1] * ex;
        const double rx = x0 - from[0], ry = y0 - from[1];
        const double t_n = rx * (-ey) + ry * ex, b_n = dir[0] * ry - dir[1] * rx;
        const double sg = det > 0.0 ? 1.0 : -1.0, ad = fabs(det);
        const bool hit = det * det > (SM * SM) * dn2 * (ex * ex + ey * ey) &&
                         b_n * sg >= -SM * ad && b_n * sg <= (1.0 + SM) * ad && t_n * sg <= 0.0;
        if (hit) { tn = t_n; n1 = b_n; n2 = 0.0; bdet = det; return 0; }
        return -1;
    } else {
        constexpr int TPS = (E == E_HEX) ? 2 : 1;
        const int p0 = S.sidetab[s * 4];
        double ed[TPS + 1][3], r[3], q[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double x0 = NSB_FCOL(S.xs, p0 * 3 + d);
            r[d] = from[d] - x0;
#pragma unroll
            for (int j = 0; j <= TPS; j++) ed[j][d] = NSB_FCOL(S.xs, S.sidetab[s * 4 + 1 + j] * 3 + d) - x0;
        }
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
}

// Ray / element-boundary intersection. Hex with ray_fast: the cut side is PREDICTED from the ray direction in reference
// coordinates (s = J^-1 dir; going upstream the first plane xi_i in {0, 1} reached) and confirmed with the exact tests of
// the reference routine; only if the confirmation fails are the sides searched in reference order. For an element that is
// star-shaped w.r.t. the ip (checked once per mesh, fused_ray_safety) exactly one boundary triangle is hit, so the predicted
// side is the one the ordered search returns; cuts on an edge shared by two sides give the same point from either side.
template <int E>
NSB_HD bool fused_ray_cut(const FusedSmem<E>& S, int el, int ip, const double* from, const double* dir, bool fast,
                          const double* sref, int& side_out, double* gcut, double* lcut)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NSIDE = C::NSIDE;
    constexpr int TPS = (E == E_HEX) ? 2 : 1;
    double tn = 0.0, n1 = 0.0, n2 = 0.0, bdet = 1.0;
    const double dn2 = dotv<DIM>(dir, dir);
    int side = -1, tri = -1;
    if constexpr (E == E_HEX) {
        if (fast) {
            // upstream along xi(t) = xi_ip + t s, t < 0: plane xi_i = 0 is reached at t = -xi_i / s_i (s_i > 0), xi_i = 1 at (1 - xi_i) / s_i (s_i < 0)
            float best = -3.0e38f; int bs = -1;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const float si = (float)sref[i];               // s = J^-1 dir (reference-space direction of the ray)
                const float xi = (float)S.lip[ip * 3 + i];
                if (si != 0.0f) {
                    const float t = si > 0.0f ? -xi / si : (1.0f - xi) / si;
                    // reference sides: zeta=0 -> 0, eta=0 -> 1, xi=1 -> 2, eta=1 -> 3, xi=0 -> 4, zeta=1 -> 5
                    const int sd = i == 0 ? (si > 0.0f ? 4 : 2) : (i == 1 ? (si > 0.0f ? 1 : 3) : (si > 0.0f ? 0 : 5));
                    if (t > best) { best = t; bs = sd; }
                }
            }
            if (bs >= 0) {
                tri = fused_ray_side<E>(S, el, bs, from, dir, dn2, tn, n1, n2, bdet);
                if (tri >= 0) side = bs;
            }
        }
    }
    if (side < 0) {
        for (int s = 0; s < NSIDE; s++) {
            tri = fused_ray_side<E>(S, el, s, from, dir, dn2, tn, n1, n2, bdet);
            if (tri >= 0) { side = s; break; }
        }
        if (side < 0) return false;
    }
    const double ibd = 1.0 / bdet;
    if constexpr (DIM == 2) {
        const double t = tn * ibd, bc = n1 * ibd;
        const int p0 = S.sidetab[side * 4], p1 = S.sidetab[side * 4 + 1];
#pragma unroll
        for (int d = 0; d < 2; d++) {
            gcut[d] = from[d] + t * dir[d];
            lcut[d] = (1 - bc) * S.cortab[p0 * 3 + d] + bc * S.cortab[p1 * 3 + d];
        }
    } else {
        const double t = tn * ibd, b1 = n1 * ibd, b2 = n2 * ibd;
        const int kk = TPS == 2 ? tri : 0;
        const int p0 = S.sidetab[side * 4], p1 = S.sidetab[side * 4 + 1 + kk], p2 = S.sidetab[side * 4 + 2 + kk];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            gcut[d] = from[d] + t * dir[d];
            lcut[d] = (1 - b1 - b2) * S.cortab[p0 * 3 + d] + b1 * S.cortab[p1 * 3 + d] + b2 * S.cortab[p2 * 3 + d];
        }
    }
    side_out = side;
    return true;
}

// upwind shapes of one ip (No / Full / Skewed / LPS); see upwind_uniform (ns_owner.cuh), upwind.cpp:52-80,133-172,381-430,505-575
template <int E>
NSB_HD bool fused_upwind(const FusedSmem<E>& S, int el, int ip, int type, bool fast, const double* sref,
                         const double* n, const double* xip, const double* N, int from, int to, const double* vel,
                         double* up, double& len)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NSH = C::NSH;
    if (type == UPW_NO) {
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = N[k];
        len = 1.0;
        return true;
    }
    if (type == UPW_FULL) {
        const double flux = dotv<DIM>(n, vel);
        const int co = flux > 0.0 ? from : to;
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = (k == co) ? 1.0 : 0.0;
        double s = 0;
#pragma unroll
        for (int d = 0; d < DIM; d++) { const double t = xip[d] - NSB_FCOL(S.xs, co * DIM + d); s += t * t; }
        len = sqrt(s);
        return true;
    }
#pragma unroll
    for (int k = 0; k < NSH; k++) up[k] = 0.0;
    if (sqrt(dotv<DIM>(vel, vel)) < 1e-14) { len = 1.0; return true; }      // upwind.cpp:407-413, 531-537
    int side = 0; double gc[DIM], lc[DIM];
    if (!fused_ray_cut<E>(S, el, ip, xip, vel, fast, sref, side, gc, lc)) { len = 1.0; return false; }
    constexpr int NSC = (DIM == 2) ? 2 : (E == E_TET ? 3 : 4);
    if (type == UPW_SKEWED) {                                    // GetNodeNextToCut, upwind.cpp:337-379
        double mn = 1.79769313486231570e308; int bestc = 0;
#pragma unroll
        for (int i = 0; i < NSC; i++) {
            const int co = S.sidetab[side * 4 + i];
            double dd = 0;
#pragma unroll
            for (int d = 0; d < DIM; d++) { const double t = gc[d] - NSB_FCOL(S.xs, co * DIM + d); dd += t * t; }
            if (dd < mn) { mn = dd; bestc = co; }
        }
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = (k == bestc) ? 1.0 : 0.0;
        double s = 0;
#pragma unroll
        for (int d = 0; d < DIM; d++) { const double t = xip[d] - NSB_FCOL(S.xs, bestc * DIM + d); s += t * t; }
        len = sqrt(s);
    } else {                                                     // LPS, upwind.cpp:562-573
        double Nc[NSH];
        lagrange<E>(lc, Nc);
        int mask = 0;
#pragma unroll
        for (int i = 0; i < NSC; i++) mask |= 1 << S.sidetab[side * 4 + i];
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = ((mask >> k) & 1) ? Nc[k] : 0.0;
        len = sqrt(dist2<DIM>(xip, gc));
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// flux phase: one SCVF (local element el, ip) -> lean record `fr` (shared memory). Same arithmetic as fv1_flux_kernel<LEAN>
// (ns_owner.cuh), staged so that few values are live across the ray search: with ~225 KB of the SM's 228 KB used as shared
// memory there is no L1 left for register spills, so everything that has to survive the upwind search is parked in the record
// itself (the upwind shapes live in the dK slots until they are scaled in place, the partial defect fluxes in the F slots,
// J^-T n in the pK slots).  nd = global node ids of the element's corners (time-dependent closure only).
// ------------------------------------------------------------------------------------------------
template <int E, int STAB, bool TD, bool GEOT>
NSB_HD bool fused_scvf(const FusedArgs& A, const FusedSmem<E>& S, int el, int ip, const int32_t* __restrict__ nd, const double* __restrict__ geo,
                       double* __restrict__ fr)
{
    using C = FusedCfg<E>;
    using LR = LeanRec<E>;
    constexpr int DIM = C::DIM, NSH = C::NSH, NIP = C::NIP, NF = C::NF, P = DIM;
    constexpr int O_GPN = LR::HEAD - 1;                          // scratch: the padding slot of the record head
    static_assert(LR::HEAD - 1 >= NF + DIM && DIM <= LR::NSHP, "record scratch layout");
    const KParams& p = A.p;
    const bool td = TD && p.time_dep;
    const double nurho = p.visc * p.rho;
    const bool want_def = p.what & W_DEF_A, want_jac = p.what & W_JAC_A;
    bool ok = true;
    const int from = S.iptab[ip * 12], to = S.iptab[ip * 12 + 1];
    const double* N = S.Nt + ip * C::NSTR;
    double n[DIM], xip[DIM], std[DIM], sref[DIM], dlinv = 0.0, sn, oacc = 0.0;
    {
        // ---- stage 1: geometry, StdVel, and everything that needs J^-T (then J^-T is dead) ----
        double JI[DIM][DIM];
        if constexpr (GEOT) {
            // static SCVF geometry record (fused_geom_record, built once per mesh): n | xip | J^-T | 1/L_d^2
#pragma unroll
            for (int d = 0; d < DIM; d++) { n[d] = geo[d]; xip[d] = geo[3 + d]; }
#pragma unroll
            for (int d = 0; d < DIM; d++)
#pragma unroll
                for (int i = 0; i < DIM; i++) JI[d][i] = geo[6 + d * DIM + i];
            dlinv = geo[15];
        } else {
            double cen[DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < NSH; k++) s += NSB_FCOL(S.xs, k * DIM + d);
                cen[d] = s * (1.0 / NSH);
            }
            // COR diffusion length: element-wide statistics of the SCVF normals (diffusion_length.h:139-172)
            double cmn = 0.0, cav = 0.0, cmd = 0.0;
            if (STAB != STAB_NONE && p.diff_len == DIFF_COR) {
                cmn = 1.79769313486231570e308; cmd = 1.79769313486231570e308;
                for (int i = 0; i < NIP; i++) {
                    double nn_[DIM], xx_[DIM], dsi;
                    fused_ip_geometry<E>(S, el, i, cen, nn_, xx_, dsi, false, nullptr);
                    const double q = dotv<DIM>(nn_, nn_);
                    if (q < cmn) cmn = q;
                    cav += q;
                    if (DIM == 3 && dsi < cmd) cmd = dsi;
                }
                cav /= NIP;
            }
            double ds = 0.0;
            fused_ip_geometry<E>(S, el, ip, cen, n, xip, ds, true, JI);
            if (STAB != STAB_NONE) dlinv = diff_len_sq_inv<DIM>(p.diff_len, dotv<DIM>(n, n), NSB_FCOL(S.vs, from), NSB_FCOL(S.vs, to), ds, cmn, cav, cmd);
        }
        // StdVel from the `u` argument (:282-293)
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < NSH; k++) s += NSB_FCOL(S.us, k * NF + d) * N[k];
            std[d] = s;
        }
        sn = dotv<DIM>(std, n);
        // reference-space direction of the upwind ray, J^-1 StdVel (predicted-side search), and J^-T n (pressure column)
#pragma unroll
        for (int i = 0; i < DIM; i++) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) { a += JI[d][i] * std[d]; b += JI[d][i] * n[d]; }
            sref[i] = a;
            fr[LR::O_PK + i] = b;                                // scratch until the pK coefficients are formed
        }
        // defect (:686-776), part 1: diffusive + pressure flux and grad p . n. The local gradient tensor is summed first and
        // mapped to global gradients once.
        if (want_def) {
            double Lg[DIM][NF], L0[DIM];
#pragma unroll
            for (int i = 0; i < DIM; i++) {
                L0[i] = 0.0;
#pragma unroll
                for (int q = 0; q < NF; q++) Lg[i][q] = 0.0;
            }
            double pr = 0.0;
#pragma unroll
            for (int k = 0; k < NSH; k++) {
                double dl[DIM], uk[NF];
#pragma unroll
                for (int i = 0; i < DIM; i++) dl[i] = S.dnt[ip * C::DSTR + k * DIM + i];
#pragma unroll
                for (int q = 0; q < NF; q++) uk[q] = NSB_FCOL(S.us, k * NF + q);
#pragma unroll
                for (int i = 0; i < DIM; i++)
#pragma unroll
                    for (int q = 0; q < NF; q++) Lg[i][q] += dl[i] * uk[q];
                pr += N[k] * uk[P];
                if (STAB != STAB_NONE && td) {                   // the closure uses solution(0) (:296, :646) and the old solution
                    const double p0k = A.s0[(int64_t)nd[k] * NF + P];
                    double o = 0.0;
#pragma unroll
                    for (int i = 0; i < DIM; i++) L0[i] += dl[i] * p0k;
#pragma unroll
                    for (int d = 0; d < DIM; d++) o += A.s1[(int64_t)nd[k] * NF + d] * n[d];
                    oacc += N[k] * o;
                }
            }
            double gv[DIM][DIM], gpn = 0.0;
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
                fr[LR::O_F + d1] = df * (-1.0) * nurho + pr * n[d1];     // the convective part is added in stage 3
            }
            fr[O_GPN] = gpn;
        }
    }
    const double prod = sn * p.rho;
    // ---- stage 2: the stabilisation's upwind; its shapes live in the dK slots of the record ----
    double* up = fr + LR::O_DK;
    double uplen = 1.0;
#pragma unroll
    for (int k = 0; k < NSH; k++) up[k] = 0.0;
    const bool fast = S.efast[el] != 0;
    if (!p.stokes) ok &= fused_upwind<E>(S, el, ip, p.upw_stab, fast, sref, n, xip, N, from, to, std, up, uplen);
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
    // everything that uses the STABILISATION's upwind shapes happens here: the convective upwind below may overwrite `up`
    double U[DIM];                                               // upwind_vel of the stabilisation's upwind (upwind_interface.h:334-358)
#pragma unroll
    for (int d = 0; d < DIM; d++) U[d] = 0.0;
    {
        const double ci = inv * p.rho;
        double acc = 0.0;                                        // time-dependent closure sum  sum_k sb_k (s0_k . n)
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            const double upk = p.stokes ? 0.0 : up[k];
            const double sb = qa * N[k] + qb * upk;
            if (want_jac) fr[LR::O_CK + k] = (STAB == STAB_NONE) ? N[k] * p.rho : sb * ci;       // continuity-row coefficients (:561-584)
#pragma unroll
            for (int d = 0; d < DIM; d++) U[d] += upk * NSB_FCOL(S.us, k * NF + d);
            if (STAB != STAB_NONE && want_def && td) {
                double sk = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) sk += A.s0[(int64_t)nd[k] * NF + d] * n[d];
                acc += sb * sk;
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
                acc -= fr[O_GPN] * p.inv_rho;
                if (p.has_source) {
#pragma unroll
                    for (int d = 0; d < DIM; d++) acc += p.src[d] * n[d];
                }
                if (td) acc += oacc / p.dt;
                cont = acc * ci;
            }
            fr[LR::O_F + P] = cont;
        }
    }
    // ---- stage 3: convective upwind, transported velocity, Peclet blend ----
    double w = 1.0;
    if (!p.stokes) {
        if (p.upw_conv != p.upw_stab) {
            double l2;
            ok &= fused_upwind<E>(S, el, ip, p.upw_conv, fast, sref, n, xip, N, from, to, std, up, l2);
#pragma unroll
            for (int d = 0; d < DIM; d++) U[d] = 0.0;
#pragma unroll
            for (int k = 0; k < NSH; k++)
#pragma unroll
                for (int d = 0; d < DIM; d++) U[d] += up[k] * NSB_FCOL(S.us, k * NF + d);
        }
        if (p.peclet) {                                           // peclet_blend :871-892
            double dd = 0;
#pragma unroll
            for (int d = 0; d < DIM; d++) { const double t = NSB_FCOL(S.xs, to * DIM + d) - NSB_FCOL(S.xs, from * DIM + d); dd += t * t; }
            const double Pe = sn / dotv<DIM>(n, n) * sqrt(dd) / p.visc;
            const double Pe2 = Pe * Pe;
            w = Pe2 / (5.0 + Pe2);
#pragma unroll
            for (int d = 0; d < DIM; d++) U[d] = w * U[d] + (1.0 - w) * std[d];
        }
        if (want_def) {
#pragma unroll
            for (int d = 0; d < DIM; d++) fr[LR::O_F + d] += U[d] * prod;
        }
    }
    // ---- stage 4: Jacobian coefficients ----
    if (want_jac) {
        const double cw = prod * w, cpe = prod * (1.0 - w);
#pragma unroll
        for (int k = 0; k < NSH; k++) {                          // convective diagonal (:430-468): the upwind shapes are scaled in place
            double D = 0.0;
            if (!p.stokes) { D = up[k] * cw; if (p.peclet) D += cpe * N[k]; }
            up[k] = D;
        }
        // pressure column of the continuity row (:586-592): -G_k.n / diag, with G_k.n = dnt_k . (J^-T n)
        double mv[DIM];
#pragma unroll
        for (int i = 0; i < DIM; i++) mv[i] = fr[LR::O_PK + i] * (-1.0 * inv);
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < DIM; i++) s += S.dnt[ip * C::DSTR + k * DIM + i] * mv[i];
            fr[LR::O_PK + k] = s;
        }
#pragma unroll
        for (int d = 0; d < DIM; d++) fr[LR::O_N + d] = n[d];
    }
    return ok;
}

// ------------------------------------------------------------------------------------------------
// lane functions of one patch (tid in [0, NT))
// ------------------------------------------------------------------------------------------------
// asynchronous global -> shared copies (cp.async, SASS LDGSTS): issued for the NEXT patch while the rows of the current one
// are assembled, completed by fused_copy_wait before the barrier that hands the buffers to the next iteration.
NSB_HD void fused_copy8(void* dst, const void* src)
{
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#else
    *reinterpret_cast<uint64_t*>(dst) = *reinterpret_cast<const uint64_t*>(src);
#endif
}
NSB_HD void fused_copy16(void* dst, const void* src)
{
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#else
    reinterpret_cast<uint64_t*>(dst)[0] = reinterpret_cast<const uint64_t*>(src)[0];
    reinterpret_cast<uint64_t*>(dst)[1] = reinterpret_cast<const uint64_t*>(src)[1];
#endif
}
NSB_HD void fused_copy4(void* dst, const void* src)
{
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#else
    *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(src);
#endif
}
NSB_HD void fused_copy_wait()
{
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

// load of one patch: tables -> work / adjbuf[par] / nodebuf[par], corner data of the patch's elements -> element rows.
// Only the connectivity lookups are synchronous loads; everything else is an asynchronous copy.
template <int E> NSB_HD void fused_load(const FusedArgs& A, const FusedSmem<E>& S, const PatchHdr& H, int par, int tid)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NSH = C::NSH, NF = C::NF;
    for (int i = tid; i < H.n_work; i += C::NT) fused_copy4(S.work + i, A.work + H.work0 + i);
    for (int i = tid; i < H.n_adj; i += C::NT) fused_copy16(S.adj0 + par * C::MAXA + i, A.adj + H.adj0 + i);
    for (int i = tid; i < H.n_node; i += C::NT) fused_copy16(S.nodes0 + par * C::MAXN + i, A.nodes + H.node0 + i);
    for (int i = tid; i < H.n_elem * NSH; i += C::NT) {
        const int el = i / NSH, k = i - el * NSH;
        const int64_t e = A.elems[H.elem0 + el];
        const int64_t ndk = A.pconn[(int64_t)(H.elem0 + el) * NSH + k];
#pragma unroll
        for (int f = 0; f < NF; f++) fused_copy8(&NSB_FCOL(S.us, k * NF + f), A.u + ndk * NF + f);
#pragma unroll
        for (int d = 0; d < DIM; d++) fused_copy8(&NSB_FCOL(S.xs, k * DIM + d), A.coords + ndk * DIM + d);
        fused_copy8(&NSB_FCOL(S.vs, k), A.scvvol + e * NSH + k);
        if (k == 0) { S.efast[el] = A.elem_fast ? A.elem_fast[e] : (uint8_t)0; S.elem_id[el] = (int32_t)e; }
    }
}

// flux: returns false if a ray search failed (the reference throws, upwind.cpp:354)
template <int E, int STAB, bool TD, bool GEOT> NSB_HD bool fused_flux(const FusedArgs& A, const FusedSmem<E>& S, const PatchHdr& H, int tid)
{
    using C = FusedCfg<E>;
    bool ok = true;
    for (int w = tid; w < H.n_work; w += C::NT) {
        const uint32_t wi = S.work[w];
        const int el = wi & 255, ip = (wi >> 8) & 15, slot = wi >> 12;
        ok &= fused_scvf<E, STAB, TD, GEOT>(A, S, el, ip, A.pconn + (int64_t)(H.elem0 + el) * C::NSH,
                                            GEOT ? A.geo + (int64_t)(H.work0 + w) * C::GEO : nullptr, S.rec + slot * C::RSTR);
    }
    return ok;
}

// rows phase, accumulation. lane = (patch node, corner k): for every adjacent element of the node (in the order of the global
// adjacency list) the lane sums the NINC incident SCVF records of its corner and adds the convective diagonal D, the
// velocity columns C[DIM] and the pressure column PP of the continuity row into the node's per-slot accumulators `accn`
// (and the defect fluxes, k < NF, into `fs`). The corners of one element are distinct nodes: a step is conflict-free; every
// lane of a warp runs the same code (no divergence between value types); steps are separated by __syncwarp on the device.
// step 1: clear the node's accumulators (its NSH lanes)
template <int E> NSB_HD void fused_rows_zero(const FusedSmem<E>& S, double* accn, int k)
{
    using C = FusedCfg<E>;
    for (int i = k; i < C::NV * S.cntp; i += C::NSH) accn[i] = 0.0;
}

// step 2: adjacency entry j of node nl. The 16-byte entry is read once: record slots, local corner + the orientation bits of
// the NINC incident SCVFs (bit 4 + t set: the node is the `to` end), CSR slots of the element's corners.
template <int E> NSB_HD void fused_rows_accum_step(const FusedArgs& A, const FusedSmem<E>& S, const FusedTab<E>& T, double* accn, int nl, int k, int j, double& fs)
{
    using C = FusedCfg<E>;
    using LR = LeanRec<E>;
    constexpr int DIM = C::DIM, NF = C::NF, NINC = C::NINC;
    const KParams& p = A.p;
    const bool jac_a = p.what & W_JAC_A, def_a = p.what & W_DEF_A;
    const PatchNode& Nd = T.nodes[nl];
    if (j >= Nd.adj_cnt) return;
    struct U4 { uint32_t x, y, z, w; };
    const U4 a = *reinterpret_cast<const U4*>(T.adj + Nd.adj_off + j);
    const uint32_t lab = a.y >> 16;                              // la | orientation bits << 4 | self << 8
    const uint32_t em = k < 4 ? a.z : a.w;                       // emap[0..3], emap[4..7]
    const int slot = (em >> (8 * (k & 3))) & 255;
    const int rs[3] = {(int)(a.x & 0xffff), (int)(a.x >> 16), (int)(a.y & 0xffff)};
    double D = 0.0, PP = 0.0, F = 0.0, Cn[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) Cn[d] = 0.0;
#pragma unroll
    for (int t = 0; t < NINC; t++) {
        const double* rc = S.rec + rs[t] * C::RSTR;
        const double sg = ((lab >> (4 + t)) & 1) ? -1.0 : 1.0;
        if (def_a) F += sg * rc[LR::O_F + (k < NF ? k : 0)];
        if (jac_a) {
            D += sg * rc[LR::O_DK + k];
            PP += sg * rc[LR::O_PK + k];
            const double w = sg * rc[LR::O_CK + k];
#pragma unroll
            for (int d = 0; d < DIM; d++) Cn[d] += w * rc[LR::O_N + d];
        }
    }
    fs += F;
    if (jac_a) {
        const double sa = p.scale_a;
        accn[slot] += sa * D;
#pragma unroll
        for (int d = 0; d < DIM; d++) accn[(1 + d) * S.cntp + slot] += sa * Cn[d];
        accn[(1 + DIM) * S.cntp + slot] += sa * PP;
    }
}

// step 3 (one lane of the node): lumped mass on the diagonal (add_jac_M_elem :781-808)
template <int E> NSB_HD void fused_rows_mass(const FusedArgs& A, const FusedTab<E>& T, double* accn, int nl)
{
    const KParams& p = A.p;
    const PatchNode& Nd = T.nodes[nl];
    if (!(p.what & W_JAC_M) || Nd.adj_cnt == 0) return;
    const int self = T.adj[Nd.adj_off].self;
    accn[self] += p.scale_m * A.nodevol[Nd.node] * p.rho;
}

// step 4: defect entry (node, component k < NF)   (add_def_A_elem / add_def_M_elem / add_rhs_elem)
template <int E> NSB_HD void fused_rows_defect(const FusedArgs& A, const FusedTab<E>& T, int nl, int k, double fs)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NF = C::NF;
    const KParams& p = A.p;
    const int64_t a = T.nodes[nl].node;
    double d = (p.what & W_DEF_A) ? fs : 0.0;
    const bool need_vol = ((p.what & W_RHS) && p.has_source) || (p.what & W_DEF_M);
    const double vol = (need_vol && k < DIM) ? A.nodevol[a] : 0.0;
    if ((p.what & W_RHS) && p.has_source && k < DIM) d -= p.src[k] * vol * p.rho;
    d *= p.scale_a;
    if ((p.what & W_DEF_M) && k < DIM) d += p.scale_m * A.u[a * NF + k] * vol * p.rho;
    double* q = A.def + a * NF + k;
    *q = (A.beta == 0.0) ? d : A.beta * (*q) + d;
}

// rows phase, output: the NF rows of node nl, written once by a whole warp (`lane`, `nlanes` = 32 on the device). Matrix row rf
// (compile-time) is written in rounds of nlanes consecutive words (3-D: 16-byte chunks, 2-D: doubles), so that consecutive
// lanes write consecutive addresses and no index arithmetic depends on rf at run time. The J0 words of the first JIT rounds of
// every momentum row are passed in registers by the device (fused_j0_prefetch, issued before the accumulation phase so that
// their latency is hidden); jpre == nullptr reads them here.
template <int E> struct FusedJ0 { double v[FusedCfg<E>::DIM][FusedCfg<E>::JIT][FusedCfg<E>::NF == 4 ? 2 : 1]; };

template <int E> NSB_HD void fused_j0_prefetch(const FusedArgs& A, const FusedTab<E>& T, int nl, int lane, int nlanes, FusedJ0<E>& J)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NF = C::NF, W = NF == 4 ? 2 : 1;
    const PatchNode& Nd = T.nodes[nl];
    const int nwr = Nd.cnt * NF / W;                             // words per row
    const double* j0g = A.j0 + Nd.b0 * (DIM * NF);
#pragma unroll
    for (int rf = 0; rf < DIM; rf++)
#pragma unroll
        for (int it = 0; it < C::JIT; it++) {
            const int i = lane + it * nlanes;
            if (i < nwr) {
#ifdef __CUDA_ARCH__
                if constexpr (W == 2) { const double2 t = __ldcs(reinterpret_cast<const double2*>(j0g) + rf * nwr + i); J.v[rf][it][0] = t.x; J.v[rf][it][1] = t.y; }
                else J.v[rf][it][0] = __ldcs(j0g + rf * nwr + i);
#else
                for (int q = 0; q < W; q++) J.v[rf][it][q] = j0g[(rf * nwr + i) * W + q];
#endif
            }
        }
}

template <int E> NSB_HD void fused_rows_out(const FusedArgs& A, const FusedSmem<E>& S, const FusedTab<E>& T, const double* acc, int nl, int lane, int nlanes, const FusedJ0<E>* jpre)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NF = C::NF;
    const KParams& p = A.p;
    const bool jac_a = p.what & W_JAC_A;
    const PatchNode& Nd = T.nodes[nl];
    const int cnt = Nd.cnt, cntp = S.cntp;
    const double s_visc = p.visc * p.rho * p.scale_a, s_pres = p.scale_a;
    double* out = A.val + Nd.b0 * (NF * NF);
    const double* j0g = A.j0 + Nd.b0 * (DIM * NF);
    const double beta = A.beta;
    if constexpr (NF == 4) {
        const int n2 = 2 * cnt;                                  // 16-byte chunks per row
        // chunk i of row rf: slot = i / 2, column pair cp = i & 1
        auto momentum = [&](int rf, int i, bool have, double jx, double jy) {
            const int slot = i >> 1, cp = i & 1;
            double vx = 0.0, vy = 0.0;
            if (jac_a) {
                if (!have) {
#ifdef __CUDA_ARCH__
                    const double2 t = __ldcs(reinterpret_cast<const double2*>(j0g) + rf * n2 + i); jx = t.x; jy = t.y;
#else
                    jx = j0g[2 * (rf * n2 + i)]; jy = j0g[2 * (rf * n2 + i) + 1];
#endif
                }
                vx = jx * s_visc; vy = jy * (cp ? s_pres : s_visc);
            }
            const double D = acc[slot];
            if (cp == (rf >> 1)) { if (rf & 1) vy += D; else vx += D; }
#ifdef __CUDA_ARCH__
            double2* o2 = reinterpret_cast<double2*>(out) + rf * n2 + i;
            if (beta == 0.0) __stcs(o2, make_double2(vx, vy));
            else { double2 o = *o2; o.x = beta * o.x + vx; o.y = beta * o.y + vy; *o2 = o; }
#else
            double* o = out + 2 * (rf * n2 + i);
            if (beta == 0.0) { o[0] = vx; o[1] = vy; } else { o[0] = beta * o[0] + vx; o[1] = beta * o[1] + vy; }
#endif
        };
#pragma unroll
        for (int rf = 0; rf < DIM; rf++) {
#pragma unroll
            for (int it = 0; it < C::JIT; it++) {
                const int i = lane + it * nlanes;
                if (i < n2) { if (jpre) momentum(rf, i, true, jpre->v[rf][it][0], jpre->v[rf][it][1]); else momentum(rf, i, false, 0.0, 0.0); }
            }
            for (int i = lane + C::JIT * nlanes; i < n2; i += nlanes) momentum(rf, i, false, 0.0, 0.0);
        }
        for (int i = lane; i < n2; i += nlanes) {                // continuity row: (C0, C1) | (C2, PP) per slot
            const int slot = i >> 1, cp = i & 1;
            const double vx = acc[(1 + 2 * cp) * cntp + slot], vy = acc[(2 + 2 * cp) * cntp + slot];
#ifdef __CUDA_ARCH__
            double2* o2 = reinterpret_cast<double2*>(out) + DIM * n2 + i;
            if (beta == 0.0) __stcs(o2, make_double2(vx, vy));
            else { double2 o = *o2; o.x = beta * o.x + vx; o.y = beta * o.y + vy; *o2 = o; }
#else
            double* o = out + 2 * (DIM * n2 + i);
            if (beta == 0.0) { o[0] = vx; o[1] = vy; } else { o[0] = beta * o[0] + vx; o[1] = beta * o[1] + vy; }
#endif
        }
    } else {
        const int rowlen = cnt * NF;
        auto momentum = [&](int rf, int i, bool have, double jv) {
            const int slot = i / NF, cf = i - slot * NF;
            double v = 0.0;
            if (jac_a) v = (have ? jv : j0g[rf * rowlen + i]) * (cf < DIM ? s_visc : s_pres);
            if (cf == rf) v += acc[slot];
            double* o = out + rf * rowlen + i;
            *o = (beta == 0.0) ? v : beta * (*o) + v;
        };
#pragma unroll
        for (int rf = 0; rf < DIM; rf++) {
#pragma unroll
            for (int it = 0; it < C::JIT; it++) {
                const int i = lane + it * nlanes;
                if (i < rowlen) { if (jpre) momentum(rf, i, true, jpre->v[rf][it][0]); else momentum(rf, i, false, 0.0); }
            }
            for (int i = lane + C::JIT * nlanes; i < rowlen; i += nlanes) momentum(rf, i, false, 0.0);
        }
        for (int i = lane; i < rowlen; i += nlanes) {
            const int slot = i / NF, cf = i - slot * NF;
            const double v = acc[(1 + cf) * cntp + slot];
            double* o = out + DIM * rowlen + i;
            *o = (beta == 0.0) ? v : beta * (*o) + v;
        }
    }
}

// static SCVF geometry record of (element corners x [NSH][DIM], SCV volumes, ip): n | xip | J^-T | 1 / L_d^2 for the given
// diffusion-length type (fv1/diffusion_length.h:47-198). Same formulas as fused_ip_geometry, reference tables read directly.
template <int E> NSB_DEV void fused_geom_record(const double* x, const double* vol, int ip, int diff_len, double* r)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NSH = C::NSH, NIP = C::NIP;
    double cen[DIM];
    for (int d = 0; d < DIM; d++) { double s = 0; for (int k = 0; k < NSH; k++) s += x[k * DIM + d]; cen[d] = s * (1.0 / NSH); }
    auto scvf = [&](int q, double* n, double* xip, double& ds) {
        const int f = tab::EDGE[E][q][0], t = tab::EDGE[E][q][1];
        double c0[DIM];
        for (int d = 0; d < DIM; d++) c0[d] = 0.5 * (x[f * DIM + d] + x[t * DIM + d]);
        if constexpr (DIM == 2) {
            n[0] = cen[1] - c0[1]; n[1] = -(cen[0] - c0[0]);
            xip[0] = 0.5 * (c0[0] + cen[0]); xip[1] = 0.5 * (c0[1] + cen[1]);
            ds = 0.0;
        } else {
            constexpr int NFC = (E == E_TET) ? 3 : 4;
            const int fa = tab::SCVF_FA[E][q], fb = tab::SCVF_FB[E][q];
            double c1[3] = {0, 0, 0}, c3[3] = {0, 0, 0};
            for (int i = 0; i < NFC; i++) {
                const int ka = tab::SIDE[E][fa][i], kb = tab::SIDE[E][fb][i];
                for (int d = 0; d < 3; d++) { c1[d] += x[ka * 3 + d]; c3[d] += x[kb * 3 + d]; }
            }
            double a[3], b[3];
            for (int d = 0; d < 3; d++) {
                c1[d] *= (1.0 / NFC); c3[d] *= (1.0 / NFC);
                a[d] = cen[d] - c0[d]; b[d] = c3[d] - c1[d];
                xip[d] = 0.25 * (c0[d] + c1[d] + cen[d] + c3[d]);
            }
            cross3(n, a, b);
            for (int d = 0; d < 3; d++) n[d] *= 0.5;
            ds = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
        }
    };
    double cmn = 0.0, cav = 0.0, cmd = 0.0;
    if (diff_len == DIFF_COR) {
        cmn = 1.79769313486231570e308; cmd = 1.79769313486231570e308;
        for (int q = 0; q < NIP; q++) {
            double nn_[DIM], xx_[DIM], dsi;
            scvf(q, nn_, xx_, dsi);
            const double v = dotv<DIM>(nn_, nn_);
            if (v < cmn) cmn = v;
            cav += v;
            if (DIM == 3 && dsi < cmd) cmd = dsi;
        }
        cav /= NIP;
    }
    double n[DIM], xip[DIM], ds;
    scvf(ip, n, xip, ds);
    double JT[DIM][DIM], JI[DIM][DIM];
    for (int i = 0; i < DIM; i++) for (int j = 0; j < DIM; j++) JT[i][j] = 0.0;
    for (int k = 0; k < NSH; k++) for (int j = 0; j < DIM; j++) for (int i = 0; i < DIM; i++) JT[i][j] += tab::C_DNIP[E][ip][k][i] * x[k * DIM + j];
    inv_mat<DIM>(JT, JI);
    for (int q = 0; q < C::GEO; q++) r[q] = 0.0;
    for (int d = 0; d < DIM; d++) { r[d] = n[d]; r[3 + d] = xip[d]; }
    for (int d = 0; d < DIM; d++) for (int i = 0; i < DIM; i++) r[6 + d * DIM + i] = JI[d][i];
    r[15] = diff_len_sq_inv<DIM>(diff_len, dotv<DIM>(n, n), vol[tab::EDGE[E][ip][0]], vol[tab::EDGE[E][ip][1]], ds, cmn, cav, cmd);
}

// true when every boundary triangle of the element (sides in reference order, quadrilaterals as (p0,p1,p2), (p0,p2,p3)) is
// seen from its inner side by every SCVF ip: the triangulated boundary is then star-shaped w.r.t. each ip, a ray from the
// ip cuts exactly one triangle, and the predicted-side ray search of fused_ray_cut returns what the ordered search returns.
template <int E> NSB_DEV bool fused_star_shaped(const double* x)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP, NSIDE = ET<E>::NSIDE;
    if constexpr (DIM == 2) return true;
    else {
        constexpr int TPS = (E == E_HEX) ? 2 : 1;
        bool ok = true;
        for (int ip = 0; ip < NIP; ip++) {
            double xi[3], N[NSH], xip[3] = {0, 0, 0};
            for (int d = 0; d < 3; d++) xi[d] = tab::LIP[E][ip][d];
            lagrange<E>(xi, N);
            for (int k = 0; k < NSH; k++) for (int d = 0; d < 3; d++) xip[d] += N[k] * x[k * 3 + d];
            int npos = 0, nneg = 0;
            for (int s = 0; s < NSIDE; s++) for (int kk = 0; kk < TPS; kk++) {
                const int p0 = tab::SIDE[E][s][0], p1 = tab::SIDE[E][s][1 + kk], p2 = tab::SIDE[E][s][2 + kk];
                double e1[3], e2[3], r[3], nrm[3];
                for (int d = 0; d < 3; d++) { e1[d] = x[p1 * 3 + d] - x[p0 * 3 + d]; e2[d] = x[p2 * 3 + d] - x[p0 * 3 + d]; r[d] = xip[d] - x[p0 * 3 + d]; }
                cross3(nrm, e1, e2);
                const double v = dotv<3>(r, nrm);
                if (v * v > 1e-16 * dotv<3>(r, r) * dotv<3>(nrm, nrm)) { if (v > 0) npos++; else nneg++; }
            }
            if (!((npos == NSIDE * TPS && nneg == 0) || (nneg == NSIDE * TPS && npos == 0))) ok = false;
        }
        return ok;
    }
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// the kernel: persistent CTAs, patches handed out by an atomic ticket
// ------------------------------------------------------------------------------------------------
template <int E, int STAB, bool TD, bool GEOT>
__global__ void __launch_bounds__(FusedCfg<E>::NT, FusedCfg<E>::CTAS) fv1_fused_kernel(const FusedArgs A, int max_cnt)
{
    using C = FusedCfg<E>;
    constexpr int NSH = C::NSH, NF = C::NF, NPW = C::NPW, NWARP = C::NWARP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FusedLayout<E> L(max_cnt);
    const FusedSmem<E> S(smem_raw, L);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    fused_stage_tables<E>(S, tid, C::NT);
    const int what = A.p.what;
    const bool want_jac = what & (W_JAC_A | W_JAC_M), want_def = what & (W_DEF_A | W_DEF_M | W_RHS);
    const bool flux_needed = what & (W_JAC_A | W_DEF_A), jac_a = what & W_JAC_A;
    // rows phase, accumulation: lane = (node jj of the NPW the warp handles at a time, corner k)
    const int jj = lane / NSH, k = lane - jj * NSH;
    const bool lane_on = jj < NPW;
    const int j0_lines = (max_cnt * C::DIM * NF * (int)sizeof(double) + 127) >> 7;
    auto load_hdr = [&](int pi) {
        PatchHdr H;
        const int4* hp = reinterpret_cast<const int4*>(A.hdr + pi);
        const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        H.node0 = h0.x; H.n_node = h0.y; H.elem0 = h0.z; H.n_elem = h0.w; H.work0 = h1.x; H.n_work = h1.y; H.adj0 = h1.z; H.n_adj = h1.w;
        return H;
    };
    // static schedule: CTA b assembles the patches b, b + gridDim.x, ... ; the tables and corner data of the next patch are
    // copied asynchronously (cp.async) into the buffers the rows phase does not use while the rows of the current one are written
    int pi = blockIdx.x, par = 0;
    if (pi >= A.n_patch) return;
    PatchHdr H = load_hdr(pi);
    fused_load<E>(A, S, H, par, tid);
    for (;;) {
        const FusedTab<E> T(S, par);
        fused_copy_wait();
        __syncthreads();                                         // tables + element rows of this patch are in shared memory
        const int pn = pi + (int)gridDim.x;
        PatchHdr Hn = H;
        if (pn < A.n_patch) Hn = load_hdr(pn);                   // consumed after the flux phase
        if (jac_a) {                                             // the J0 rows of the patch nodes are read at the end of the patch: pull them into L2 now
            for (int i = tid; i < H.n_node * j0_lines; i += C::NT) {
                const int nl = i / j0_lines, li = i - nl * j0_lines;
                const PatchNode& Nd = T.nodes[nl];
                if (li * 16 < Nd.cnt * (C::DIM * NF)) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(A.j0 + Nd.b0 * (C::DIM * NF) + li * 16));
            }
        }
        if (pn < A.n_patch) {                                    // ... and the tables of the next patch (loaded right after the flux phase)
            const int nb = (Hn.n_elem * NSH * 4 + 127) >> 7, nw = (Hn.n_work * 4 + 127) >> 7, na = (Hn.n_adj * 16 + 127) >> 7;
            const char* q = nullptr;
            if (tid < nb) q = reinterpret_cast<const char*>(A.pconn + (int64_t)Hn.elem0 * NSH) + tid * 128;
            else if (tid < nb + nw) q = reinterpret_cast<const char*>(A.work + Hn.work0) + (tid - nb) * 128;
            else if (tid < nb + nw + na) q = reinterpret_cast<const char*>(A.adj + Hn.adj0) + (tid - nb - nw) * 128;
            else if (tid < nb + nw + na + 4) q = reinterpret_cast<const char*>(A.elems + Hn.elem0) + (tid - nb - nw - na) * 128;
            if (q) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(q));
        }
        if (flux_needed && !fused_flux<E, STAB, TD, GEOT>(A, S, H, tid)) atomicExch(A.errflag, 1);
        __syncthreads();                                         // records complete; element rows, work list and flags are dead
        if (pn < A.n_patch) {
            fused_load<E>(A, S, Hn, par ^ 1, tid);
            if (GEOT && flux_needed && tid < Hn.n_work)          // the next patch's geometry records: one 128-byte line per SCVF -> L2
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(A.geo + (int64_t)(Hn.work0 + tid) * C::GEO));
        }
        // accumulation: the first warps take NPW nodes each
        for (int nl0 = warp * NPW; nl0 < H.n_node; nl0 += NWARP * NPW) {
            const int nl = nl0 + jj;
            const bool on = lane_on && nl < H.n_node;
            double* const accn = S.acc + (size_t)(on ? nl : 0) * (C::NV * S.cntp);
            double fs = 0.0;
            if (on && want_jac) fused_rows_zero<E>(S, accn, k);
            __syncwarp();
            if (flux_needed) {
                const int mycnt = on ? (int)T.nodes[nl].adj_cnt : 0;
                const int mx = __reduce_max_sync(0xffffffffu, mycnt);
                for (int j = 0; j < mx; j++) {
                    if (on) fused_rows_accum_step<E>(A, S, T, accn, nl, k, j, fs);
                    __syncwarp();
                }
            }
            if (on && k == 0) fused_rows_mass<E>(A, T, accn, nl);
            if (on && want_def && k < NF) fused_rows_defect<E>(A, T, nl, k, fs);
        }
        __syncthreads();                                         // accumulators complete
        if (want_jac) {
            // output: warp w writes the rows of the nodes w, w + NWARP, ...; the J0 words of a node (L2-resident: prefetched during
            // the flux phase) are loaded in one batch before the node's rows are formed
            for (int nl = warp; nl < H.n_node; nl += NWARP) {
                FusedJ0<E> jr;
                if (jac_a) fused_j0_prefetch<E>(A, T, nl, lane, 32, jr);
                fused_rows_out<E>(A, S, T, S.acc + (size_t)nl * (C::NV * S.cntp), nl, lane, 32, jac_a ? &jr : nullptr);
            }
        }
        if (pn >= A.n_patch) break;
        pi = pn; H = Hn; par ^= 1;
    }
}

// once per mesh (and per diffusion-length type): static SCVF geometry records in work-item order. One CTA per patch.
template <int E>
__global__ void __launch_bounds__(FusedCfg<E>::NT) fused_geom_kernel(const FusedArgs A, int diff_len, double* __restrict__ geo)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NSH = C::NSH;
    const PatchHdr H = A.hdr[blockIdx.x];
    for (int w = threadIdx.x; w < H.n_work; w += blockDim.x) {
        const uint32_t wi = A.work[H.work0 + w];
        const int el = wi & 255, ip = (wi >> 8) & 15;
        const int64_t e = A.elems[H.elem0 + el];
        double x[NSH * DIM], vol[NSH];
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            const int64_t nd = A.pconn[(int64_t)(H.elem0 + el) * NSH + k];
#pragma unroll
            for (int d = 0; d < DIM; d++) x[k * DIM + d] = A.coords[nd * DIM + d];
            vol[k] = A.scvvol[e * NSH + k];
        }
        double r[C::GEO];
        fused_geom_record<E>(x, vol, ip, diff_len, r);
        double2* o = reinterpret_cast<double2*>(geo + (int64_t)(H.work0 + w) * C::GEO);
#pragma unroll
        for (int q = 0; q < C::GEO / 2; q++) o[q] = make_double2(r[2 * q], r[2 * q + 1]);
    }
}

// once per mesh (hex): elem_fast[e] = 1 when element e is star-shaped w.r.t. every one of its ips
template <int E>
__global__ void fused_ray_safety_kernel(int64_t n_elem, const int32_t* __restrict__ conn, const double* __restrict__ coords, uint8_t* __restrict__ elem_fast)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elem) return;
    double x[NSH * DIM];
#pragma unroll
    for (int k = 0; k < NSH; k++) {
        const int64_t nd = conn[e * NSH + k];
#pragma unroll
        for (int d = 0; d < DIM; d++) x[k * DIM + d] = coords[nd * DIM + d];
    }
    elem_fast[e] = fused_star_shaped<E>(x) ? 1 : 0;
}
#endif  // __CUDACC__

}  // namespace nsb
