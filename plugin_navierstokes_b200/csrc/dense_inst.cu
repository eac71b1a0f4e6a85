// dense_inst.cu -- instantiates the dense-ip-system element kernel for one element type (-DNSB_ELEM=e)
#include <cstdlib>
#include "ns_dense.cuh"
#include "ns_launch.h"
#ifndef NSB_ELEM
#error "compile with -DNSB_ELEM=0..3"
#endif
namespace nsb {
constexpr int E = NSB_ELEM;
template <int SC, bool PAC> static cudaError_t dense_sc(NSB_ELEM_ARGS)
{
    // one warp = one element with a ~30 KB (hex) workspace in shared memory: small blocks pack more warps per SM
    // (1 warp/block: 7 blocks/SM; 4 warps/block: 1 block/SM)
    static const int WPB = [] { const char* ev = getenv("NSB_DENSE_WPB"); const int v = ev ? atoi(ev) : 1; return (v >= 1 && v <= 4) ? v : 1; }();
    const size_t smem = sizeof(DenseWS<E, PAC>) * WPB;
    auto kern = fv1_dense_kernel<E, SC, PAC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t nblk = (n_list + WPB - 1) / WPB;
    kern<<<(unsigned)nblk, WPB * 32, smem, st>>>(k, m, list, n_list, u, s0, s1, val, def, jl, dl, d_err);
    return cudaGetLastError();
}
#define NSB_FWD sc, k, m, list, n_list, u, s0, s1, val, def, jl, dl, d_err, st
cudaError_t NSB_CAT(launch_dense_, NSB_ELEM)(NSB_ELEM_ARGS)
{
    if (n_list <= 0) return cudaSuccess;
    if (k.pac) {
        if (sc == SC_COLORED) return dense_sc<SC_COLORED, true>(NSB_FWD);
        if (sc == SC_ATOMIC) return dense_sc<SC_ATOMIC, true>(NSB_FWD);
        return dense_sc<SC_LOCAL, true>(NSB_FWD);
    }
    if (sc == SC_COLORED) return dense_sc<SC_COLORED, false>(NSB_FWD);
    if (sc == SC_ATOMIC) return dense_sc<SC_ATOMIC, false>(NSB_FWD);
    return dense_sc<SC_LOCAL, false>(NSB_FWD);
}
}
