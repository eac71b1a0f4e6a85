// ns_launch.h -- host-callable launchers; each element type is instantiated in its own translation unit
// (fv1_inst.cu / dense_inst.cu compiled with -DNSB_ELEM=0..3) so the library builds in parallel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ns_fv1.cuh"

namespace nsb {
struct MeshDev;

struct LaunchInfo { int64_t launches = 0; };

#define NSB_ELEM_ARGS int sc, const KParams& k, const MeshDev& m, const int32_t* list, int64_t n_list, const double* u, \
    const double* s0, const double* s1, double* val, double* def, double* jl, double* dl, int* d_err, cudaStream_t st
#define NSB_GATHER_ARGS const KParams& k, const MeshDev& m, double* rec, const double* u, const double* s0, const double* s1, double beta, \
    double* val, double* def, int* d_err, cudaStream_t st, int sm_count, unsigned long long* work_counter
#define NSB_DECL(E)                                                                                   \
    cudaError_t launch_elem_##E(NSB_ELEM_ARGS);                                                       \
    cudaError_t launch_dense_##E(NSB_ELEM_ARGS);                                                      \
    cudaError_t launch_gather_##E(NSB_GATHER_ARGS);                                                   \
    cudaError_t launch_scvvol_##E(int64_t n_elem, const int32_t* conn, const double* coords, double* scvvol, cudaStream_t st); \
    cudaError_t launch_geom_##E(int64_t n_elem, const int32_t* conn, const double* coords, double* rec, int stride, cudaStream_t st); \
    int scvf_record_doubles_##E(bool flow, bool exact);                                               \
    cudaError_t launch_split_##E(NSB_GATHER_ARGS, const double* j0);                                  \
    cudaError_t launch_j0_##E(const MeshDev& m, int laplace, double* j0, cudaStream_t st, int sm_count); \
    int lean_record_doubles_##E();
NSB_DECL(0) NSB_DECL(1) NSB_DECL(2) NSB_DECL(3)
#undef NSB_DECL
struct FusedArgs;
struct PatchCaps;
#define NSB_DECLF(E)                                                                                  \
    PatchCaps fused_caps_##E();                                                                       \
    size_t fused_smem_bytes_##E(int max_cnt);                                                         \
    cudaError_t launch_fused_##E(const FusedArgs& A, int max_cnt, cudaStream_t st, int sm_count, unsigned long long* work_counter); \
    cudaError_t launch_fused_geom_##E(const FusedArgs& A, int diff_len, double* geo, cudaStream_t st);                                \
    cudaError_t launch_ray_safety_##E(int64_t n_elem, const int32_t* conn, const double* coords, uint8_t* elem_fast, cudaStream_t st);
NSB_DECLF(0) NSB_DECLF(1) NSB_DECLF(2) NSB_DECLF(3)
#undef NSB_DECLF
struct TileArgs;
#define NSB_DECLT(E)                                                                                  \
    PatchCaps tile_caps_##E();                                                                        \
    size_t tile_smem_bytes_##E();                                                                     \
    int tile_max_cnt_##E();                                                                           \
    cudaError_t launch_tile_##E(const TileArgs& A, cudaStream_t st, int sm_count);
NSB_DECLT(2) NSB_DECLT(3)
#undef NSB_DECLT
struct FvcrDev;
cudaError_t launch_fvcr_0(int sc, const KParams& k, const FvcrDev& m, const int32_t* list, int64_t n_list, const double* u,
                          double* val, double* def, int* d_err, cudaStream_t st);
cudaError_t launch_fvcr_2(int sc, const KParams& k, const FvcrDev& m, const int32_t* list, int64_t n_list, const double* u,
                          double* val, double* def, int* d_err, cudaStream_t st);
#define NSB_CAT2(a, b) a##b
#define NSB_CAT(a, b) NSB_CAT2(a, b)
}
