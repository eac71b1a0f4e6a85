// ns_launch_prism.h -- host-callable launchers of the prism translation units (prism_inst.cu, dense_inst.cu -DNSB_ELEM=4)
#pragma once
#include "ns_launch.h"
namespace nsb {
cudaError_t launch_elem_4(NSB_ELEM_ARGS);
cudaError_t launch_dense_4(NSB_ELEM_ARGS);
cudaError_t launch_scvvol_4(int64_t n_elem, const int32_t* conn, const double* coords, double* scvvol, cudaStream_t st);
}
namespace nsb {
// FVCR on quadrilaterals / hexahedra (fvcrq_inst.cu -DNSB_ELEM=1 / 3)
struct FvcrDev;
cudaError_t launch_fvcrq_1(int sc, const KParams& k, const FvcrDev& m, const int32_t* list, int64_t n_list, const double* u,
                           double* val, double* def, int* d_err, cudaStream_t st);
cudaError_t launch_fvcrq_3(int sc, const KParams& k, const FvcrDev& m, const int32_t* list, int64_t n_list, const double* u,
                           double* val, double* def, int* d_err, cudaStream_t st);
}
