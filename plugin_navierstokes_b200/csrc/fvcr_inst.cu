// fvcr_inst.cu -- instantiates the FVCR element kernel for one simplex type (-DNSB_ELEM=0 tri, 2 tet)
#include <cstdlib>
#include "ns_fvcr.cuh"
#include "ns_launch.h"
#ifndef NSB_ELEM
#error "compile with -DNSB_ELEM=0 or 2"
#endif
namespace nsb {
constexpr int E = NSB_ELEM;
template <int SC> static cudaError_t fvcr_sc(const KParams& k, const FvcrDev& m, const int32_t* list, int64_t n_list, const double* u,
                                             double* val, double* def, int* d_err, cudaStream_t st)
{
    if (n_list <= 0) return cudaSuccess;
    constexpr int L = CRT<E>::NS * CRT<E>::DIM + 1, EPW = 32 / L, WPB = 4;
    const size_t smem = sizeof(CRWS<E>) * EPW * WPB;
    // NSB_FVCR_MINB = blocks/SM the register allocation is bounded for. The kernel waits on its gathers (long-scoreboard stalls
    // 5.7 per issue at 16 warps/SM): 80 registers with 0.4 KB of spill and 24 warps/SM is the measured optimum on B200
    // (config 4, 6.3 M tets, coloured: MINB 4 17.07 ms, 5 15.07, 6 13.98, 8 15.54; atomic 13.38 / 11.87 / 11.44 / 12.33)
    static const int minb = [] { const char* ev = getenv("NSB_FVCR_MINB"); return ev ? atoi(ev) : 6; }();
    auto kern = minb == 5 ? fvcr_elem_kernel<E, SC, 5> : minb == 6 ? fvcr_elem_kernel<E, SC, 6> : minb == 8 ? fvcr_elem_kernel<E, SC, 8> : fvcr_elem_kernel<E, SC, 4>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t ngrp = (n_list + EPW - 1) / EPW, nblk = (ngrp + WPB - 1) / WPB;
    kern<<<(unsigned)nblk, WPB * 32, smem, st>>>(k, m, list, n_list, u, val, def, d_err);
    return cudaGetLastError();
}
cudaError_t NSB_CAT(launch_fvcr_, NSB_ELEM)(int sc, const KParams& k, const FvcrDev& m, const int32_t* list, int64_t n_list,
                                            const double* u, double* val, double* def, int* d_err, cudaStream_t st)
{
    if (sc == SC_ATOMIC) return fvcr_sc<SC_ATOMIC>(k, m, list, n_list, u, val, def, d_err, st);
    return fvcr_sc<SC_COLORED>(k, m, list, n_list, u, val, def, d_err, st);
}
}
