// tile_inst.cu -- instantiates the fused tile kernel (ns_tile.cuh) for one 3-D element type (-DNSB_ELEM=2|3)
#include <algorithm>
#include <cstdlib>
#include "ns_tile.cuh"
#include "ns_launch.h"
#ifndef NSB_ELEM
#error "compile with -DNSB_ELEM=2..3"
#endif
namespace nsb {
constexpr int E = NSB_ELEM;

PatchCaps NSB_CAT(tile_caps_, NSB_ELEM)() { return TileCfg<E>::caps(); }
size_t NSB_CAT(tile_smem_bytes_, NSB_ELEM)() { return TileLayout<E>().total; }
int NSB_CAT(tile_max_cnt_, NSB_ELEM)() { return TileCfg<E>::MAXCNT; }

template <int STAB, bool TD>
static cudaError_t tile_t(const TileArgs& A, cudaStream_t st, int sm_count)
{
    const size_t smem = TileLayout<E>().total;
    auto kern = fv1_tile_kernel<E, STAB, TD>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static const int cta_per_sm = [] { const char* ev = getenv("NSB_TILE_CTAS"); const int v = ev ? atoi(ev) : 2; return v >= 1 && v <= 2 ? v : 2; }();
    const int nblk = (int)std::min<int64_t>(A.n_tile, (int64_t)sm_count * cta_per_sm);
    if (nblk <= 0) return cudaSuccess;
    kern<<<nblk, TileCfg<E>::NT, smem, st>>>(A);
    return cudaGetLastError();
}

cudaError_t NSB_CAT(launch_tile_, NSB_ELEM)(const TileArgs& A, cudaStream_t st, int sm_count)
{
    if (A.p.stab == STAB_FIELDS) return A.p.time_dep ? tile_t<STAB_FIELDS, true>(A, st, sm_count) : tile_t<STAB_FIELDS, false>(A, st, sm_count);
    return A.p.time_dep ? tile_t<STAB_NONE, true>(A, st, sm_count) : tile_t<STAB_NONE, false>(A, st, sm_count);
}
}
