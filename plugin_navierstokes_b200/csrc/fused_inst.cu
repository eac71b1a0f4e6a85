// fused_inst.cu -- instantiates the fused patch kernel (ns_fused.cuh) for one element type (-DNSB_ELEM=e)
#include <algorithm>
#include <cstdlib>
#include "ns_fused.cuh"
#include "ns_launch.h"
#ifndef NSB_ELEM
#error "compile with -DNSB_ELEM=0..3"
#endif
namespace nsb {
constexpr int E = NSB_ELEM;

PatchCaps NSB_CAT(fused_caps_, NSB_ELEM)() { return FusedCfg<E>::caps(); }
size_t NSB_CAT(fused_smem_bytes_, NSB_ELEM)(int max_cnt) { return FusedLayout<E>(max_cnt).total; }

template <int STAB, bool TD, bool GEOT>
static cudaError_t fused_t(const FusedArgs& A, int max_cnt, cudaStream_t st, int sm_count, unsigned long long* work_counter)
{
    const size_t smem = FusedLayout<E>(max_cnt).total;
    auto kern = fv1_fused_kernel<E, STAB, TD, GEOT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    (void)work_counter;
    const int nblk = (int)std::min<int64_t>(A.n_patch, (int64_t)sm_count * FusedCfg<E>::CTAS);
    if (nblk <= 0) return cudaSuccess;
    kern<<<nblk, FusedCfg<E>::NT, smem, st>>>(A, max_cnt);
    return cudaGetLastError();
}

cudaError_t NSB_CAT(launch_fused_, NSB_ELEM)(const FusedArgs& A, int max_cnt, cudaStream_t st, int sm_count, unsigned long long* work_counter)
{
#define NSB_FGO(ST, TDV) (A.geo ? fused_t<ST, TDV, true>(A, max_cnt, st, sm_count, work_counter) : fused_t<ST, TDV, false>(A, max_cnt, st, sm_count, work_counter))
    if (A.p.stab == STAB_FIELDS) return A.p.time_dep ? NSB_FGO(STAB_FIELDS, true) : NSB_FGO(STAB_FIELDS, false);
    return A.p.time_dep ? NSB_FGO(STAB_NONE, true) : NSB_FGO(STAB_NONE, false);
#undef NSB_FGO
}

cudaError_t NSB_CAT(launch_fused_geom_, NSB_ELEM)(const FusedArgs& A, int diff_len, double* geo, cudaStream_t st)
{
    if (A.n_patch <= 0) return cudaSuccess;
    fused_geom_kernel<E><<<A.n_patch, FusedCfg<E>::NT, 0, st>>>(A, diff_len, geo);
    return cudaGetLastError();
}

cudaError_t NSB_CAT(launch_ray_safety_, NSB_ELEM)(int64_t n_elem, const int32_t* conn, const double* coords, uint8_t* elem_fast, cudaStream_t st)
{
    fused_ray_safety_kernel<E><<<(unsigned)((n_elem + 127) / 128), 128, 0, st>>>(n_elem, conn, coords, elem_fast);
    return cudaGetLastError();
}
}
