// fvcrq_inst.cu -- instantiates the FVCR element kernel of the non-affine CR geometries (-DNSB_ELEM=1 quadrilateral, 3 hexahedron)
#include "ns_fvcr_q.cuh"
#include "ns_launch.h"
#include "ns_launch_prism.h"
#ifndef NSB_ELEM
#error "compile with -DNSB_ELEM=1 or 3"
#endif
namespace nsb {
constexpr int E = NSB_ELEM;
template <int SC> static cudaError_t fvcrq_sc(const KParams& k, const FvcrDev& m, const int32_t* list, int64_t n_list, const double* u,
                                              double* val, double* def, int* d_err, cudaStream_t st)
{
    if (n_list <= 0) return cudaSuccess;
    constexpr int L = CRT<E>::NS * CRT<E>::DIM + 1, EPW = 32 / L, WPB = 4;
    const size_t smem = sizeof(CRQWS<E>) * EPW * WPB;
    auto kern = fvcrq_elem_kernel<E, SC, 4>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t ngrp = (n_list + EPW - 1) / EPW, nblk = (ngrp + WPB - 1) / WPB;
    kern<<<(unsigned)nblk, WPB * 32, smem, st>>>(k, m, list, n_list, u, val, def, d_err);
    return cudaGetLastError();
}
cudaError_t NSB_CAT(launch_fvcrq_, NSB_ELEM)(int sc, const KParams& k, const FvcrDev& m, const int32_t* list, int64_t n_list,
                                             const double* u, double* val, double* def, int* d_err, cudaStream_t st)
{
    if (sc == SC_ATOMIC) return fvcrq_sc<SC_ATOMIC>(k, m, list, n_list, u, val, def, d_err, st);
    return fvcrq_sc<SC_COLORED>(k, m, list, n_list, u, val, def, d_err, st);
}
}
