// fv1_inst.cu -- instantiates the FV1 element / gather / SCV-volume kernels for one element type (-DNSB_ELEM=e)
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "ns_owner.cuh"
#include "ns_split.cuh"
#include "ns_launch.h"
#ifndef NSB_ELEM
#error "compile with -DNSB_ELEM=0..3"
#endif
namespace nsb {
constexpr int E = NSB_ELEM;

template <int SC, bool PAC> static cudaError_t elem_sc(NSB_ELEM_ARGS)
{
    constexpr int L = ET<E>::NSH * (ET<E>::DIM + 1), EPW = 32 / L, WPB = 4;
    const size_t smem = sizeof(ElemWS<E, PAC>) * EPW * WPB;
    auto kern = fv1_elem_kernel<E, SC, PAC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t ngrp = (n_list + EPW - 1) / EPW, nblk = (ngrp + WPB - 1) / WPB;
    kern<<<(unsigned)nblk, WPB * 32, smem, st>>>(k, m, list, n_list, u, s0, s1, val, def, jl, dl, d_err);
    return cudaGetLastError();
}
#define NSB_FWD sc, k, m, list, n_list, u, s0, s1, val, def, jl, dl, d_err, st
cudaError_t NSB_CAT(launch_elem_, NSB_ELEM)(NSB_ELEM_ARGS)
{
    if (n_list <= 0) return cudaSuccess;
    if (k.pac) {
        if (sc == SC_COLORED) return elem_sc<SC_COLORED, true>(NSB_FWD);
        if (sc == SC_ATOMIC) return elem_sc<SC_ATOMIC, true>(NSB_FWD);
        return elem_sc<SC_LOCAL, true>(NSB_FWD);
    }
    if (sc == SC_COLORED) return elem_sc<SC_COLORED, false>(NSB_FWD);
    if (sc == SC_ATOMIC) return elem_sc<SC_ATOMIC, false>(NSB_FWD);
    return elem_sc<SC_LOCAL, false>(NSB_FWD);
}

// first ticket of a rows launch (0, or the first node behind the priority nodes of a phased assembly)
static __global__ void set_ticket_kernel(unsigned long long* counter, unsigned long long v) { *counter = v; }
static cudaError_t set_ticket(unsigned long long* counter, int64_t v, cudaStream_t st)
{
    if (v == 0) return cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
    set_ticket_kernel<<<1, 1, 0, st>>>(counter, (unsigned long long)v);
    return cudaGetLastError();
}

// owner-computes path: (A) flux kernel, thread per element  ->  (B) rows kernel, warp per node
template <int STAB, bool EXACT, int CHP, int MINB>
static cudaError_t rows_t(const KParams& k, const MeshDev& m, const double* rec, const double* u, double beta, double* val,
                          double* def, cudaStream_t st, int sm_count, unsigned long long* work_counter, int WPB)
{
    constexpr int NF = ET<E>::DIM + 1;
    using WS = RowWS<E, STAB == STAB_FLOW, EXACT, CHP>;
    const size_t tab_bytes = rows_tab_bytes<E>(m.max_cnt);
    const size_t per_warp = (sizeof(WS) + sizeof(double) * NF * NF * m.max_cnt + 15) & ~(size_t)15;
    const size_t smem = tab_bytes + per_warp * WPB;
    auto kb = fv1_rows_kernel<E, STAB, EXACT, CHP, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kb, WPB * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    const int64_t nblk = std::min<int64_t>((m.n_node - m.node_begin + WPB - 1) / WPB, (int64_t)sm_count * occ);
    if (nblk <= 0) return cudaSuccess;
    e = set_ticket(work_counter, m.node_begin, st);
    if (e != cudaSuccess) return e;
    kb<<<(unsigned)nblk, WPB * 32, smem, st>>>(k, m, rec, u, beta, val, def, work_counter);
    return cudaGetLastError();
}

template <int STAB, bool EXACT> static cudaError_t gather_t(NSB_GATHER_ARGS)
{
    constexpr int NF = ET<E>::DIM + 1, NIP = ET<E>::NIP, NSH = ET<E>::NSH, DIM = ET<E>::DIM, BS = 128;
    cudaError_t e;
    if ((k.what & (W_JAC_A | W_DEF_A)) && !m.skip_flux) {
        const size_t smem_a = sizeof(double) * (NSB_CSTR(E) * BS + NIP * NSH * DIM + NIP * NSH + 24) + sizeof(int) * (NIP * 12 + 24)
                              + 16 + sizeof(double) * BS * flux_stage_stride(FluxRec<E, STAB == STAB_FLOW, EXACT>::SZ);   // staged flux records (one slot per lane)
        static const int minb = [] { const char* ev = getenv("NSB_FLUX_MINB"); return ev ? atoi(ev) : (DIM == 2 ? 4 : 2); }();   // 2-D: 4 blocks/SM (config 1: 6.85 -> 6.71 ms)
        auto ka = minb == 2 ? fv1_flux_kernel<E, STAB, EXACT, BS, 2> : minb == 4 ? fv1_flux_kernel<E, STAB, EXACT, BS, 4> : fv1_flux_kernel<E, STAB, EXACT, BS, 3>;
        e = cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
        if (e != cudaSuccess) return e;
        ka<<<(unsigned)((m.n_elem + BS - 1) / BS), BS, smem_a, st>>>(k, m, u, s0, s1, rec, d_err);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    // experiment knobs (hex only): NSB_ROWS_CH = adjacent elements staged per round, NSB_ROWS_MINB = blocks/SM the
    // register allocation is bounded for
    // measured best on B200 for hex (profiles/r1_history.md): 4 elements per round, 96 registers, 2 warps per block
    static const int WPB = [] { const char* ev = getenv("NSB_ROWS_WPB"); const int v = ev ? atoi(ev) : 2; return (v >= 1 && v <= 3) ? v : 2; }();   // 2 warps per block (config 1: 7.14 -> 6.85 ms)
    static const int CHV = [] { const char* ev = getenv("NSB_ROWS_CH"); return ev ? atoi(ev) : 4; }();
    static const int MINBV = [] { const char* ev = getenv("NSB_ROWS_MINB"); return ev ? atoi(ev) : 6; }();
    if constexpr (E == 3) {
        if (CHV == 4 && MINBV == 8) return rows_t<STAB, EXACT, 4, 8>(k, m, rec, u, beta, val, def, st, sm_count, work_counter, WPB);
        if (CHV == 4 && MINBV == 6) return rows_t<STAB, EXACT, 4, 6>(k, m, rec, u, beta, val, def, st, sm_count, work_counter, WPB);
        if (CHV == 4) return rows_t<STAB, EXACT, 4, 5>(k, m, rec, u, beta, val, def, st, sm_count, work_counter, WPB);
        if (MINBV == 6) return rows_t<STAB, EXACT, 8, 6>(k, m, rec, u, beta, val, def, st, sm_count, work_counter, WPB);
        return rows_t<STAB, EXACT, 8, 5>(k, m, rec, u, beta, val, def, st, sm_count, work_counter, WPB);
    }
    return rows_t<STAB, EXACT, 0, 5>(k, m, rec, u, beta, val, def, st, sm_count, work_counter, WPB);
}
#define NSB_GFWD k, m, rec, u, s0, s1, beta, val, def, d_err, st, sm_count, work_counter
cudaError_t NSB_CAT(launch_gather_, NSB_ELEM)(NSB_GATHER_ARGS)
{
    const bool exact = !k.stokes && k.exact_jac != 0.0;
    if (exact) switch (k.stab) {
        case STAB_FIELDS: return gather_t<STAB_FIELDS, true>(NSB_GFWD);
        case STAB_FLOW: return gather_t<STAB_FLOW, true>(NSB_GFWD);
        default: return gather_t<STAB_NONE, true>(NSB_GFWD);
    }
    switch (k.stab) {
        case STAB_FIELDS: return gather_t<STAB_FIELDS, false>(NSB_GFWD);
        case STAB_FLOW: return gather_t<STAB_FLOW, false>(NSB_GFWD);
        default: return gather_t<STAB_NONE, false>(NSB_GFWD);
    }
}
// ---- split path (ns_split.cuh): lean flux records + static table J0 ----
template <int STAB, int CHP, int MINB, bool J0D = false>
static cudaError_t split_t(NSB_GATHER_ARGS, const double* j0)
{
    constexpr int NF = ET<E>::DIM + 1, NIP = ET<E>::NIP, NSH = ET<E>::NSH, DIM = ET<E>::DIM, BS = 128;
    cudaError_t e;
    if ((k.what & (W_JAC_A | W_DEF_A)) && !m.skip_flux) {
        // NSB_FLUX_LPE = lanes per element (hex: 1 or 4), NSB_FLUX_MINB = blocks/SM the registers are bounded for
        static const int LPEV = [] { const char* ev = getenv("NSB_FLUX_LPE"); return ev ? atoi(ev) : 4; }();
        static const int FMB = [] { const char* ev = getenv("NSB_FLUX_MINB"); return ev ? atoi(ev) : 0; }();
        auto go = [&](auto ka, int lpe) -> cudaError_t {
            const int epb = BS / lpe;
            size_t smem_a = sizeof(double) * (NSB_CSTR(E) * epb + NIP * (NSH * DIM + 1) + NIP * (NSH + 1) + 24) + sizeof(int) * (NIP * 12 + 24);
            smem_a += 16 + sizeof(double) * BS * flux_stage_stride(SplitRec<E>::SZ);   // staged records (one slot per lane)
            cudaError_t e2 = cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
            if (e2 != cudaSuccess) return e2;
            ka<<<(unsigned)((m.n_elem + epb - 1) / epb), BS, smem_a, st>>>(k, m, u, s0, s1, rec, d_err);
            return cudaGetLastError();
        };
#define NSB_FLUX_GO(MB, LP) (k.time_dep ? go(fv1_flux_kernel<E, STAB, false, BS, MB, true, LP, true>, LP) : go(fv1_flux_kernel<E, STAB, false, BS, MB, true, LP, false>, LP))
        if constexpr (E == 3) {
            if (LPEV == 4) e = FMB == 3 ? NSB_FLUX_GO(3, 4) : FMB == 5 ? NSB_FLUX_GO(5, 4) : NSB_FLUX_GO(4, 4);
            else e = FMB == 3 ? NSB_FLUX_GO(3, 1) : NSB_FLUX_GO(2, 1);
        } else e = NSB_FLUX_GO(2, 1);
#undef NSB_FLUX_GO
        if (e != cudaSuccess) return e;
    }
    static const int WPB = [] { const char* ev = getenv("NSB_SPLIT_WPB"); const int v = ev ? atoi(ev) : 2; return (v >= 1 && v <= 2) ? v : 2; }();
    if constexpr (SplitRec<E>::COMP) {
        // owner-lane rows kernel: one lane per column slot of the block row (NSB_SPLIT_OWNER=0: accumulator-copy kernel)
        static const bool owner = !(getenv("NSB_SPLIT_OWNER") && atoi(getenv("NSB_SPLIT_OWNER")) == 0);
        if (owner && m.max_cnt <= 32) {
            static const int omb = [] { const char* ev = getenv("NSB_OWNER_MINB"); return ev ? atoi(ev) : 8; }();
            const size_t smem_o = split_tab_bytes<E>() + ((sizeof(OwnWS<E>) + 15) & ~(size_t)15) * 2;
            const bool fast_o = k.what == (W_JAC_A | W_DEF_A) && beta == 0.0;
            auto ko = omb == 8 ? (fast_o ? fv1_rows_owner_kernel<E, 8, true> : fv1_rows_owner_kernel<E, 8, false>)
                    : omb == 12 ? (fast_o ? fv1_rows_owner_kernel<E, 12, true> : fv1_rows_owner_kernel<E, 12, false>)
                                : (fast_o ? fv1_rows_owner_kernel<E, 10, true> : fv1_rows_owner_kernel<E, 10, false>);
            e = cudaFuncSetAttribute(ko, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_o);
            if (e != cudaSuccess) return e;
            int occ_o = 1;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_o, ko, 64, smem_o);
            if (e != cudaSuccess) return e;
            if (occ_o < 1) return cudaErrorLaunchOutOfResources;
            const int64_t nblk_o = std::min<int64_t>((m.n_node - m.node_begin + 1) / 2, (int64_t)sm_count * occ_o);
            if (nblk_o <= 0) return cudaSuccess;
            e = set_ticket(work_counter, m.node_begin, st);
            if (e != cudaSuccess) return e;
            ko<<<(unsigned)nblk_o, 64, smem_o, st>>>(k, m, rec, j0, u, beta, val, def, work_counter);
            return cudaGetLastError();
        }
    }
    constexpr size_t tab_bytes = split_tab_bytes<E>();
    const size_t smem = tab_bytes + split_warp_bytes<E, CHP, J0D>(m.max_cnt) * WPB;
    const bool fast = k.what == (W_JAC_A | W_DEF_A) && beta == 0.0;
    auto kb = fast ? fv1_rows_split_kernel<E, CHP, MINB, J0D, true> : fv1_rows_split_kernel<E, CHP, MINB, J0D, false>;
    e = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kb, WPB * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    const int64_t nblk = std::min<int64_t>((m.n_node - m.node_begin + WPB - 1) / WPB, (int64_t)sm_count * occ);
    if (nblk <= 0) return cudaSuccess;
    e = set_ticket(work_counter, m.node_begin, st);
    if (e != cudaSuccess) return e;
    kb<<<(unsigned)nblk, WPB * 32, smem, st>>>(k, m, rec, j0, u, beta, val, def, work_counter);
    return cudaGetLastError();
}
cudaError_t NSB_CAT(launch_split_, NSB_ELEM)(NSB_GATHER_ARGS, const double* j0)
{
    static const int CHV = [] { const char* ev = getenv("NSB_SPLIT_CH"); return ev ? atoi(ev) : 4; }();
    static const int MINBV = [] { const char* ev = getenv("NSB_SPLIT_MINB"); return ev ? atoi(ev) : 12; }();
#define NSB_SPLIT_GO(CH, MB) (k.stab == STAB_FIELDS ? split_t<STAB_FIELDS, CH, MB>(NSB_GFWD, j0) : split_t<STAB_NONE, CH, MB>(NSB_GFWD, j0))
    static const int J0DV = [] { const char* ev = getenv("NSB_SPLIT_J0D"); return ev ? atoi(ev) : 0; }();
    if constexpr (E == 3) {
        if (J0DV) return MINBV == 16 ? (k.stab == STAB_FIELDS ? split_t<STAB_FIELDS, 4, 16, true>(NSB_GFWD, j0) : split_t<STAB_NONE, 4, 16, true>(NSB_GFWD, j0))
                                     : (k.stab == STAB_FIELDS ? split_t<STAB_FIELDS, 4, 12, true>(NSB_GFWD, j0) : split_t<STAB_NONE, 4, 12, true>(NSB_GFWD, j0));
        if (CHV == 4) return MINBV == 16 ? NSB_SPLIT_GO(4, 16) : NSB_SPLIT_GO(4, 12);
        return MINBV == 16 ? NSB_SPLIT_GO(0, 16) : NSB_SPLIT_GO(0, 12);
    }
    return NSB_SPLIT_GO(0, 12);
#undef NSB_SPLIT_GO
}
cudaError_t NSB_CAT(launch_j0_, NSB_ELEM)(const MeshDev& m, int laplace, double* j0, cudaStream_t st, int sm_count)
{
    const int64_t nblk = std::min<int64_t>((m.n_node + 3) / 4, (int64_t)sm_count * 16);
    fv1_j0_kernel<E><<<(unsigned)nblk, 128, 0, st>>>(m, laplace, j0);
    return cudaGetLastError();
}
int NSB_CAT(lean_record_doubles_, NSB_ELEM)() { return SplitRec<E>::SZ; }

// doubles per combined SCVF record [geometry | flux] for the given stabilisation / Jacobian flavour
int NSB_CAT(scvf_record_doubles_, NSB_ELEM)(bool flow, bool exact)
{
    const int g = GeoRec<E>::SZ;
    if (flow) return g + (exact ? FluxRec<E, true, true>::SZ : FluxRec<E, true, false>::SZ);
    return g + (exact ? FluxRec<E, false, true>::SZ : FluxRec<E, false, false>::SZ);
}

cudaError_t NSB_CAT(launch_geom_, NSB_ELEM)(int64_t n_elem, const int32_t* conn, const double* coords, double* rec, int stride, cudaStream_t st)
{
    const int64_t n = n_elem * ET<E>::NIP;
    geom_kernel<E><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n_elem, conn, coords, rec, stride);
    return cudaGetLastError();
}

cudaError_t NSB_CAT(launch_scvvol_, NSB_ELEM)(int64_t n_elem, const int32_t* conn, const double* coords, double* scvvol, cudaStream_t st)
{
    const int64_t n = n_elem * ET<E>::NSH;
    scv_volume_kernel<E><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n_elem, conn, coords, scvvol);
    return cudaGetLastError();
}
}
