// ns_dense.cuh -- FV1 element kernel for upwinds with ip-shapes (PositiveUpwind): dense nIp x nIp ip system.
#pragma once
#include "ns_kernels.cuh"
namespace nsb {
template <int E> struct DenseWS { double pad[8]; };
template <int E, int SC>
__global__ void fv1_dense_kernel(KParams p, MeshDev m, const int32_t* elem_list, int64_t n_list, const double* u,
                                 const double* s0, const double* s1, double* val, double* def, double* Jloc,
                                 double* dloc, int* errflag)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(errflag, 3);
}
}
