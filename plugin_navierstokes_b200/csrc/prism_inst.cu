// prism_inst.cu -- FV1 on prisms (fv1/navier_stokes_fv1.cpp:1542 registers the disc for Prism): the element kernels
// (coloured / atomic / local scatter; ns_kernels.cuh) and the SCV-volume table. The owner-computes paths (ns_owner.cuh,
// ns_split.cuh, ns_fused.cuh) are tuned per element type and are not instantiated for prisms: NSB_SCATTER_GATHER is served
// by the coloured element kernel (nsb_query(NSB_Q_LAST_SCATTER) reports it). The dense-ip-system kernel for PositiveUpwind
// comes from dense_inst.cu compiled with -DNSB_ELEM=4.
#include "ns_kernels.cuh"
#include "ns_launch.h"
#include "ns_launch_prism.h"
namespace nsb {
constexpr int E = E_PRISM;

template <int SC, bool PAC> static cudaError_t elem_sc(NSB_ELEM_ARGS)
{
    constexpr int WPB = 4;                               // one element per warp: lane = SCVF (9) / lane = Jacobian column (24)
    const size_t smem = sizeof(ElemWS<E, PAC>) * WPB;
    auto kern = fv1_elem_kernel<E, SC, PAC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t nblk = (n_list + WPB - 1) / WPB;
    kern<<<(unsigned)nblk, WPB * 32, smem, st>>>(k, m, list, n_list, u, s0, s1, val, def, jl, dl, d_err);
    return cudaGetLastError();
}
#define NSB_FWD sc, k, m, list, n_list, u, s0, s1, val, def, jl, dl, d_err, st
cudaError_t launch_elem_4(NSB_ELEM_ARGS)
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

cudaError_t launch_scvvol_4(int64_t n_elem, const int32_t* conn, const double* coords, double* scvvol, cudaStream_t st)
{
    const int64_t n = n_elem * ET<E>::NSH;
    scv_volume_kernel<E><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n_elem, conn, coords, scvvol);
    return cudaGetLastError();
}
}
