// ns_kernels.cuh -- FV1 assembly kernels for sm_100a (diagonal stabilisation branch).
//
//  fv1_elem_kernel   : one element (hex) or 2-3 elements (tet/quad/tri) per warp. Per-ip phase: lane = ip.
//                      Column phase: lane = (corner k, function cf) owns one COLUMN of the local Jacobian in
//                      registers. Scatter: coloured read-modify-write, red.global.add.f64, or local output.
//  (the owner-computes "gather" kernel lives in ns_gather.cuh)
#pragma once
#include <utility>
#include "ns_fv1.cuh"

namespace nsb {

// Device view of the uploaded grid + precomputed tables
struct MeshDev {
    int64_t n_elem, n_node;
    const int32_t* conn;        // [n_elem][NSH]
    const double*  coords;      // [n_node][DIM]
    const double*  scvvol;      // [n_elem][NSH]  SCV volumes (precomputed FV1Geometry table)
    const int64_t* brow;        // [n_node+1] prefix sum of block-row lengths
    const uint8_t* emap;        // [n_elem][NSH][NSH] slot of node conn[e][k] in the block row of conn[e][a]
    const int64_t* adj_ptr;     // [n_node+1] node -> adjacent (element, local corner)
    const int32_t* adj;         // elem*NSH + local corner
    int32_t max_cnt;            // longest block row
    const int32_t* node_order;  // Z-curve traversal order of the nodes (rows kernel), may be null
    int32_t l2_hints;           // rows kernel of the split path: L2 eviction-priority hints on the bulk copies
    int32_t ticket_group;       // rows kernel of the split path: nodes per atomic ticket
    const uint8_t* elem_fast;   // hex: 1 = element is star-shaped w.r.t. its ips (predicted-side ray search allowed), may be null
    // per-ip data imports (fv1/navier_stokes_fv1.cpp:184-197), null = the constants of KParams; element kernels only
    const double* ip_visc;      // [n_elem][NIP]       m_imKinViscosity at the SCVF ips
    const double* ip_rho_scvf;  // [n_elem][NIP]       m_imDensitySCVF
    const double* ip_rho_scv;   // [n_elem][NSH]       m_imDensitySCV (mass / rhs parts)
    const double* ip_src_scvf;  // [n_elem][NIP][DIM]  m_imSourceSCVF (closure of the stabilisation)
    const double* ip_src_scv;   // [n_elem][NSH][DIM]  m_imSourceSCV (add_rhs_elem)
    // phased owner-computes assembly (nsb_set_priority_nodes): the rows kernels hand out the tickets [node_begin, n_node) of the
    // node order; skip_flux = the SCVF records of the previous phase are still valid (host-side launch logic only)
    int64_t node_begin;
    int32_t skip_flux;
};

enum { SC_COLORED = 1, SC_ATOMIC = 2, SC_LOCAL = 3 };

// per-(sub-)element workspace in shared memory
template <int E, bool PAC> struct ElemWS {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP, NF = DIM + 1;
    double x[NSH * DIM];
    double u[NSH * NF];
    double s0[NSH * NF];
    double s1[NSH * NF];
    double vol[NSH];
    double nn[NIP], ds[NIP];            // COR diffusion length scratch
    int64_t rowbase[NSH];               // first value index of row (node a, fct 0)
    int32_t cnt[NSH];                   // block-row length of node a
    int32_t node[NSH];
    IpRec<E, PAC> rec[NIP];
};

template <int E> NSB_DEV void cor_stats(const double* nn, const double* ds, double& mnN, double& avN, double& mnD)
{
    constexpr int NIP = ET<E>::NIP;
    mnN = 1.79769313486231570e308; mnD = 1.79769313486231570e308; avN = 0.0;
    for (int i = 0; i < NIP; i++) { if (nn[i] < mnN) mnN = nn[i]; avN += nn[i]; if (ET<E>::DIM == 3 && ds[i] < mnD) mnD = ds[i]; }
    avN /= NIP;
}

// ------------------------------------------------------------------------------------------------
template <int E, int SC, bool PAC>
__global__ void __launch_bounds__(128) fv1_elem_kernel(KParams p, MeshDev m, const int32_t* __restrict__ elem_list,
                                                       int64_t n_list, const double* __restrict__ u,
                                                       const double* __restrict__ s0, const double* __restrict__ s1,
                                                       double* __restrict__ val, double* __restrict__ def,
                                                       double* __restrict__ Jloc, double* __restrict__ dloc,
                                                       int* __restrict__ errflag)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NIP = ET<E>::NIP, NF = DIM + 1, L = NSH * NF;
    constexpr int EPW = 32 / L;                         // elements per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ElemWS<E, PAC>* wsall = reinterpret_cast<ElemWS<E, PAC>*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / L, col = lane - sub * L;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t li = gw * EPW + sub;
    const bool active = sub < EPW && li < n_list;
    ElemWS<E, PAC>& ws = wsall[warp * EPW + (sub < EPW ? sub : 0)];
    const int64_t e = active ? (elem_list ? (int64_t)elem_list[li] : li) : 0;
    const int k = col / NF, cf = col - k * NF;

    // ---- stage element data ----
    if (active) {
        if (col < NSH) {
            const int nd = m.conn[e * NSH + col];
            ws.node[col] = nd;
            ws.vol[col] = m.scvvol[e * NSH + col];
            const int64_t b0 = m.brow[nd], b1 = m.brow[nd + 1];
            ws.rowbase[col] = b0 * (NF * NF); ws.cnt[col] = (int32_t)(b1 - b0);
        }
    }
    __syncwarp();
    if (active) {
        for (int i = col; i < NSH * NF; i += L) {
            const int kk = i / NF, ff = i - kk * NF;
            const int64_t gi = (int64_t)ws.node[kk] * NF + ff;
            ws.u[i] = u[gi];
            if (p.time_dep) { ws.s0[i] = s0[gi]; ws.s1[i] = s1[gi]; }
        }
        for (int i = col; i < NSH * DIM; i += L) {
            const int kk = i / DIM, dd = i - kk * DIM;
            ws.x[i] = m.coords[(int64_t)ws.node[kk] * DIM + dd];
        }
    }
    __syncwarp();
    // ---- per-ip phase: lane = ip ----
    double cmn = 0, cav = 0, cmd = 0;
    if (p.diff_len == DIFF_COR && p.stab != STAB_NONE) {
        if (active && col < NIP) {
            IpGeo<E> g; ip_geometry<E>(ws.x, col, g);
            ws.nn[col] = dotv<DIM>(g.n, g.n); ws.ds[col] = g.ds;
        }
        __syncwarp();
        if (active) cor_stats<E>(ws.nn, ws.ds, cmn, cav, cmd);
    }
    if (active && col < NIP) {
        const double* ps0 = p.time_dep ? ws.s0 : ws.u;
        bool ok;
        if (m.ip_visc || m.ip_rho_scvf || m.ip_src_scvf) {
            // per-ip data imports: every use of viscosity / density / source in the ip evaluation is local to the ip
            // (m_imKinViscosity[ip], m_imDensitySCVF[ip], (*pSource)[ip]; fv1/navier_stokes_fv1.cpp:336-772, stabilization.cpp:151-229)
            KParams pl = p;
            const int64_t gi = e * NIP + col;
            if (m.ip_visc) pl.visc = m.ip_visc[gi];
            if (m.ip_rho_scvf) { pl.rho = m.ip_rho_scvf[gi]; pl.inv_rho = 1.0 / pl.rho; }
            if (m.ip_src_scvf) { pl.has_source = 1; for (int d = 0; d < DIM; d++) pl.src[d] = m.ip_src_scvf[gi * DIM + d]; }
            ok = ip_eval<E, PAC>(pl, ws.x, ws.u, ps0, ws.s1, ws.vol, col, cmn, cav, cmd, ws.rec[col]);
        } else ok = ip_eval<E, PAC>(p, ws.x, ws.u, ps0, ws.s1, ws.vol, col, cmn, cav, cmd, ws.rec[col]);
        if (!ok) atomicExch(errflag, 1);
    }
    __syncwarp();
    if (!active) return;

    // ---- column phase: lane = (k, cf) ----
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
        if ((p.what & W_JAC_M) && cf < DIM) {           // add_jac_M_elem :781-808
            const double mv = p.scale_m * ws.vol[k] * (m.ip_rho_scv ? m.ip_rho_scv[e * NSH + k] : p.rho);
#pragma unroll
            for (int a = 0; a < NSH; a++)
#pragma unroll
                for (int rf = 0; rf < DIM; rf++) if (a == k && rf == cf) acc[a * NF + rf] += mv;
        }
        if (SC == SC_LOCAL) {
            // LocalMatrix layout: row index rf*NSH + a, column index cf*NSH + k
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
    // ---- defect: lane = row (a = k, rf = cf) ----
    if (p.what & (W_DEF_A | W_DEF_M | W_RHS)) {
        double d = 0.0;
        if (p.what & W_DEF_A) {
#pragma unroll
            for (int t = 0; t < ET<E>::NINC; t++) {
                const int ip = t_inc<E>(k, t);
                d += (double)t_inc_sign<E>(k, t) * ws.rec[ip].F[cf];
            }
        }
        const double rho_v = m.ip_rho_scv ? m.ip_rho_scv[e * NSH + k] : p.rho;
        if ((p.what & W_RHS) && (p.has_source || m.ip_src_scv) && cf < DIM)
            d -= (m.ip_src_scv ? m.ip_src_scv[(e * NSH + k) * DIM + cf] : p.src[cf]) * ws.vol[k] * rho_v;   // add_rhs_elem :841-869
        d *= p.scale_a;
        if ((p.what & W_DEF_M) && cf < DIM) d += p.scale_m * ws.u[k * NF + cf] * ws.vol[k] * rho_v;   // :811-838
        if (SC == SC_LOCAL) dloc[e * (int64_t)L + cf * NSH + k] = d;
        else {
            double* q = def + (int64_t)ws.node[k] * NF + cf;
            atomicAdd(q, d);
        }
    }
}

// SCV-volume table (precomputed FV1Geometry data, uploaded mesh only)
template <int E>
__global__ void scv_volume_kernel(int64_t n_elem, const int32_t* __restrict__ conn, const double* __restrict__ coords,
                                  double* __restrict__ scvvol)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elem * NSH) return;
    const int64_t e = i / NSH; const int co = (int)(i - e * NSH);
    double x[NSH * DIM];
#pragma unroll
    for (int k = 0; k < NSH; k++) {
        const int64_t nd = conn[e * NSH + k];
#pragma unroll
        for (int d = 0; d < DIM; d++) x[k * DIM + d] = coords[nd * DIM + d];
    }
    scvvol[i] = scv_volume<E>(x, co);
}

}  // namespace nsb
