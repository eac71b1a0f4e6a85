// emu_fused.cpp -- CPU EMULATION of the fused patch kernel (plugin_navierstokes_b200/csrc/ns_fused.cuh).
//
// TEST INFRASTRUCTURE ONLY: never linked into libnsb200.so, never imported by the package. It runs the very same NSB_HD
// lane functions the CUDA kernel runs (flux phase, row accumulation, row output) and the very same host-side patch builder
// (ns_patch.h, ns_graph.h), thread by thread on the CPU, so that the arithmetic and the patch tables can be compared with
// the oracle in the `-m "not gpu"` suite. Built by tests/test_fused_emu.py with g++ (no CUDA).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../plugin_navierstokes_b200/csrc/ns_fv1.cuh"
#include "../../plugin_navierstokes_b200/csrc/ns_fused.cuh"
#include "../../plugin_navierstokes_b200/csrc/ns_graph.h"

using namespace nsb;

namespace {

template <int E>
int run(int64_t n_elem, int64_t n_node, const int32_t* conn, const double* coords, const KParams& kp, const double* u,
        const double* s0, const double* s1, double beta, double* values, double* defect, int ray_fast, int use_geo, int64_t* stats, std::string& err)
{
    using C = FusedCfg<E>;
    constexpr int DIM = C::DIM, NSH = C::NSH, NF = C::NF, NINC = C::NINC, NIP = C::NIP;
    EntityGraph g;
    err = build_entity_graph(n_elem, n_node, NSH, conn, g);
    if (!err.empty()) return -1;
    std::vector<uint8_t> emap;
    build_emap(n_elem, NSH, conn, g, emap);
    // topology tables of the patch builder == reference tables of the kernels
    for (int ip = 0; ip < NIP; ip++) for (int j = 0; j < 2; j++)
        if (patch_detail::topo(E).edge[ip][j] != tab::EDGE[E][ip][j]) { err = "ns_patch.h edge table differs from ref_tables.cuh"; return -1; }
    for (int la = 0; la < NSH; la++) { int c = 0; for (int ip = 0; ip < NIP; ip++) if (tab::EDGE[E][ip][0] == la || tab::EDGE[E][ip][1] == la) {
        if (tab::INC[E][la][c] != ip) { err = "INC table is not in ascending ip order"; return -1; }
        if ((tab::INC_SIGN[E][la][c] < 0) != (tab::EDGE[E][ip][1] == la)) { err = "INC_SIGN differs from the edge orientation"; return -1; }
        c++; } }
    // precomputed tables (scv_volume_kernel, node_volume_kernel, fv1_j0_kernel restated for the host)
    std::vector<double> scvvol((size_t)n_elem * NSH), nodevol(n_node, 0.0);
    for (int64_t e = 0; e < n_elem; e++) {
        double x[NSH * DIM];
        for (int k = 0; k < NSH; k++) for (int d = 0; d < DIM; d++) x[k * DIM + d] = coords[(int64_t)conn[e * NSH + k] * DIM + d];
        for (int k = 0; k < NSH; k++) scvvol[e * NSH + k] = scv_volume<E>(x, k);
    }
    for (int64_t a = 0; a < n_node; a++) for (int64_t q = g.adj_ptr[a]; q < g.adj_ptr[a + 1]; q++) nodevol[a] += scvvol[g.adj[q]];
    std::vector<double> j0((size_t)g.brow[n_node] * DIM * NF, 0.0);
    for (int64_t a = 0; a < n_node; a++) {
        const int64_t b0 = g.brow[a];
        const int rowlen = (int)(g.brow[a + 1] - b0) * NF;
        double* out = j0.data() + b0 * (DIM * NF);
        for (int64_t q = g.adj_ptr[a]; q < g.adj_ptr[a + 1]; q++) {
            const int32_t ad = g.adj[q];
            const int e = ad / NSH, la = ad - e * NSH;
            double x[NSH * DIM];
            for (int k = 0; k < NSH; k++) for (int d = 0; d < DIM; d++) x[k * DIM + d] = coords[(int64_t)conn[(int64_t)e * NSH + k] * DIM + d];
            for (int kk = 0; kk < NSH; kk++) {
                double S[DIM][NF];
                for (int rf = 0; rf < DIM; rf++) for (int cf = 0; cf < NF; cf++) S[rf][cf] = 0.0;
                for (int t = 0; t < NINC; t++) {
                    const int ip = tab::INC[E][la][t];
                    const double sg = (double)tab::INC_SIGN[E][la][t];
                    IpGeo<E> gg;
                    ip_geometry<E>(x, ip, gg);
                    const double gn = dotv<DIM>(gg.G[kk], gg.n);
                    for (int rf = 0; rf < DIM; rf++) {
                        if (!kp.laplace) for (int cf = 0; cf < DIM; cf++) S[rf][cf] -= gg.G[kk][rf] * (sg * gg.n[cf]);
                        S[rf][rf] -= sg * gn;
                        S[rf][DIM] += gg.N[kk] * (sg * gg.n[rf]);
                    }
                }
                const int slot = emap[(int64_t)ad * NSH + kk];
                for (int rf = 0; rf < DIM; rf++) for (int cf = 0; cf < NF; cf++) out[rf * rowlen + slot * NF + cf] += S[rf][cf];
            }
        }
    }
    // star-shapedness of every element (fused_ray_safety_kernel)
    const int fast = ray_fast;
    int64_t n_bad = 0;
    std::vector<uint8_t> elem_fast(n_elem, 0);
    if (E == E_HEX) {
        for (int64_t e = 0; e < n_elem; e++) {
            double x[NSH * DIM];
            for (int k = 0; k < NSH; k++) for (int d = 0; d < DIM; d++) x[k * DIM + d] = coords[(int64_t)conn[e * NSH + k] * DIM + d];
            elem_fast[e] = fused_star_shaped<E>(x) ? 1 : 0;
            if (!elem_fast[e]) n_bad++;
        }
    }
    PatchPlan plan;
    if (!build_patch_plan(E, n_elem, n_node, conn, coords, g.adj_ptr.data(), g.adj.data(), g.brow.data(), emap.data(), C::caps(), plan, err)) return -1;
    FusedArgs A;
    std::memset(&A, 0, sizeof A);
    A.p = kp;
    A.n_patch = (int32_t)plan.hdr.size();
    A.hdr = plan.hdr.data(); A.nodes = plan.nodes.data(); A.elems = plan.elems.data(); A.pconn = plan.pconn.data();
    A.work = plan.work.data(); A.adj = plan.adj.data();
    A.coords = coords; A.scvvol = scvvol.data(); A.nodevol = nodevol.data();
    A.u = u; A.s0 = s0; A.s1 = s1; A.j0 = j0.data();
    A.beta = beta; A.val = values; A.def = defect;
    A.errflag = nullptr; A.elem_fast = fast ? elem_fast.data() : nullptr; A.max_adj = plan.max_adj_per_node;
    // static SCVF geometry records in work-item order (fused_geom_kernel)
    std::vector<double> geo;
    if (use_geo) {
        geo.resize((size_t)plan.work.size() * C::GEO);
        for (const PatchHdr& H : plan.hdr)
            for (int w = 0; w < H.n_work; w++) {
                const uint32_t wi = plan.work[H.work0 + w];
                const int el = wi & 255, ip = (wi >> 8) & 15;
                const int64_t ge = plan.elems[H.elem0 + el];
                double x[NSH * DIM], vol[NSH];
                for (int k = 0; k < NSH; k++) {
                    const int64_t nd = plan.pconn[(int64_t)(H.elem0 + el) * NSH + k];
                    for (int d = 0; d < DIM; d++) x[k * DIM + d] = coords[nd * DIM + d];
                    vol[k] = scvvol[ge * NSH + k];
                }
                fused_geom_record<E>(x, vol, ip, kp.diff_len, geo.data() + (size_t)(H.work0 + w) * C::GEO);
            }
        A.geo = geo.data();
    }
    const FusedLayout<E> L(g.max_cnt);
    std::vector<unsigned char> smem(L.total + 64);
    unsigned char* base = smem.data() + ((16 - ((uintptr_t)smem.data() & 15)) & 15);
    const FusedSmem<E> S(base, L);
    for (int tid = 0; tid < C::NT; tid++) fused_stage_tables<E>(S, tid, C::NT);
    const int what = kp.what;
    const bool want_jac = what & (W_JAC_A | W_JAC_M), want_def = what & (W_DEF_A | W_DEF_M | W_RHS), flux_needed = what & (W_JAC_A | W_DEF_A);
    bool ok = true;
    int64_t max_nodes = 0, max_work = 0, max_el = 0;
    std::vector<uint8_t> node_seen(n_node, 0);
    for (int32_t pi = 0; pi < A.n_patch; pi++) {
        const PatchHdr H = plan.hdr[pi];
        if (H.n_work > C::MAXW || H.n_elem > C::MAXE || H.n_node > C::MAXN || H.n_adj > C::MAXA) { err = "patch exceeds the kernel capacities"; return -1; }
        max_nodes = std::max<int64_t>(max_nodes, H.n_node); max_work = std::max<int64_t>(max_work, H.n_work); max_el = std::max<int64_t>(max_el, H.n_elem);
        const int par = pi & 1;
        for (int tid = 0; tid < C::NT; tid++) fused_load<E>(A, S, H, par, tid);
        const FusedTab<E> T(S, par);
        if (flux_needed) {
#define EMU_FLUX(ST, TDV) { for (int tid = 0; tid < C::NT; tid++) ok &= use_geo ? fused_flux<E, ST, TDV, true>(A, S, H, tid) : fused_flux<E, ST, TDV, false>(A, S, H, tid); }
            if (kp.stab == STAB_FIELDS) { if (kp.time_dep) EMU_FLUX(STAB_FIELDS, true) else EMU_FLUX(STAB_FIELDS, false) }
            else { if (kp.time_dep) EMU_FLUX(STAB_NONE, true) else EMU_FLUX(STAB_NONE, false) }
#undef EMU_FLUX
        }
        for (int nl = 0; nl < H.n_node; nl++) {
            if (node_seen[T.nodes[nl].node]++) { err = "a node belongs to two patches"; return -1; }
            double* accn = S.acc + (size_t)nl * (C::NV * S.cntp);
            if (want_jac) for (int k = 0; k < NSH; k++) fused_rows_zero<E>(S, accn, k);
            double fs[NSH];
            for (int k = 0; k < NSH; k++) fs[k] = 0.0;
            if (flux_needed)
                for (int j = 0; j < T.nodes[nl].adj_cnt; j++)
                    for (int k = 0; k < NSH; k++) fused_rows_accum_step<E>(A, S, T, accn, nl, k, j, fs[k]);
            fused_rows_mass<E>(A, T, accn, nl);
            if (want_def) for (int k = 0; k < NF; k++) fused_rows_defect<E>(A, T, nl, k, fs[k]);
            if (want_jac) {
                // odd nodes exercise the register-prefetched J0 path of the device, even nodes the direct reads
                FusedJ0<E> jr[32];
                for (int lane = 0; lane < 32; lane++) fused_j0_prefetch<E>(A, T, nl, lane, 32, jr[lane]);
                for (int lane = 0; lane < 32; lane++) fused_rows_out<E>(A, S, T, accn, nl, lane, 32, (nl & 1) ? &jr[lane] : nullptr);
            }
        }
    }
    for (int64_t a = 0; a < n_node; a++) if (!node_seen[a]) { err = "a node belongs to no patch"; return -1; }
    if (stats) {
        stats[0] = A.n_patch; stats[1] = plan.n_scvf_evals; stats[2] = n_elem * NIP; stats[3] = max_nodes; stats[4] = max_work;
        stats[5] = max_el; stats[6] = fast; stats[7] = n_bad;
    }
    if (!ok) { err = "GetNodeNextToCut: Cannot find cut side."; return -4; }
    return 0;
}

}  // namespace

extern "C" int emu_fused_assemble(int elem, int64_t n_elem, int64_t n_node, const int32_t* conn, const double* coords, const KParams* kp,
                                  const double* u, const double* s0, const double* s1, double beta, double* values, double* defect,
                                  int ray_fast, int use_geo, int64_t* stats, char* errbuf, int errlen)
{
    std::string err;
    int rc = -1;
    switch (elem) {
        case 0: rc = run<0>(n_elem, n_node, conn, coords, *kp, u, s0, s1, beta, values, defect, ray_fast, use_geo, stats, err); break;
        case 1: rc = run<1>(n_elem, n_node, conn, coords, *kp, u, s0, s1, beta, values, defect, ray_fast, use_geo, stats, err); break;
        case 2: rc = run<2>(n_elem, n_node, conn, coords, *kp, u, s0, s1, beta, values, defect, ray_fast, use_geo, stats, err); break;
        case 3: rc = run<3>(n_elem, n_node, conn, coords, *kp, u, s0, s1, beta, values, defect, ray_fast, use_geo, stats, err); break;
        default: err = "bad element type";
    }
    if (errbuf && errlen > 0) { std::strncpy(errbuf, err.c_str(), errlen - 1); errbuf[errlen - 1] = 0; }
    return rc;
}

extern "C" int emu_kparams_size(void) { return (int)sizeof(KParams); }
