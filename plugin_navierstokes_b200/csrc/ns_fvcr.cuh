// ns_fvcr.cuh -- FVCR (Crouzeix-Raviart) element assembly kernels.
#pragma once
#include "ns_fv1.cuh"
namespace nsb {
struct FvcrDev {
    int64_t n_elem = 0, n_node = 0, n_side = 0, nnz = 0, n_dof = 0, prow0 = 0;
    const int32_t *conn = nullptr, *esides = nullptr, *color_order = nullptr;
    const double* coords = nullptr;
    int64_t *srow = nullptr, *sadj_ptr = nullptr;
    int32_t *scnt = nullptr, *psort = nullptr;
    uint8_t *emap = nullptr, *pslot = nullptr;
    int n_colors = 0; const int64_t* color_ptr = nullptr;
};
inline void fvcr_free(FvcrDev& f)
{
    cudaFree(f.srow); cudaFree(f.sadj_ptr); cudaFree(f.scnt); cudaFree(f.psort); cudaFree(f.emap); cudaFree(f.pslot);
    f = FvcrDev{};
}
inline int fvcr_assemble(const FvcrDev& f, const KParams& k, int elem, int mode, const double* u, double beta, double* val,
                         double* def, cudaStream_t st, int sm_count, int* d_err, int64_t* launches)
{
    return -5;
}
}
