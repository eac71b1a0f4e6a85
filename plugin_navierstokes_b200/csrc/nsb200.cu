// nsb200.cu -- host side of libnsb200.so: context, grid preprocessing (adjacency, block-CSR pattern,
// element->CSR scatter map, colouring), kernel dispatch, and the extern "C" entry points of include/nsb200.h.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <functional>
#include <string>
#include <thread>
#include <unordered_map>
#include <chrono>
#include <vector>

#include "../../include/nsb200.h"
#include "ns_kernels.cuh"
#include "ns_launch.h"
#include "ns_launch_prism.h"
#include "ns_fvcr.cuh"
#include "ns_graph.h"
#include "ns_fused.cuh"
#include "ns_tile.cuh"
#include "ns_bnd.cuh"
#include "ns_turb.cuh"
#include "ns_crc.cuh"

using namespace nsb;

namespace nsb {
__global__ void scale_kernel(int64_t n, double beta, double* __restrict__ a)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] *= beta;
}
// once per mesh: SCV volume of every node (sum over the adjacent elements in adjacency order)
__global__ void node_volume_kernel(int64_t n_node, const int64_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj,
                                   const double* __restrict__ scvvol, double* __restrict__ nodevol)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_node) return;
    double s = 0.0;
    for (int64_t q = adj_ptr[a]; q < adj_ptr[a + 1]; q++) s += scvvol[adj[q]];
    nodevol[a] = s;
}
// ---- GPU-resident consumer of the assembled Jacobian: y = alpha * J x + beta * y ----
// FV1 block CSR (ns_graph.h): warp per node = NF consecutive rows, lane = column slot (stride 32): the 4 (3) x 4 (3) block is
// read once with 16-byte loads where NF == 4; fixed-order butterfly reduction -> bitwise deterministic.
template <int NF>
__global__ void __launch_bounds__(128) bcsr_spmv_kernel(int64_t row0, int64_t n_node, const int64_t* __restrict__ brow, const int32_t* __restrict__ bcol,
                                                        const double* __restrict__ val, const double* __restrict__ x, double alpha, double beta,
                                                        double* __restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t a = row0 + warp0; a < n_node; a += nwarp) {      // block rows [row0, n_node)
        const int64_t b0 = brow[a];
        const int cnt = (int)(brow[a + 1] - b0);
        const double* v = val + b0 * (NF * NF);
        double s[NF];
#pragma unroll
        for (int r = 0; r < NF; r++) s[r] = 0.0;
        for (int b = lane; b < cnt; b += 32) {
            const int64_t c = bcol[b0 + b];
            double xc[NF];
            if (NF == 4) { const double2 p = __ldg(reinterpret_cast<const double2*>(x + c * 4)), q = __ldg(reinterpret_cast<const double2*>(x + c * 4) + 1); xc[0] = p.x; xc[1] = p.y; xc[2] = q.x; xc[NF - 1] = q.y; }
            else { for (int f = 0; f < NF; f++) xc[f] = x[c * NF + f]; }
#pragma unroll
            for (int r = 0; r < NF; r++) {
                const double* row = v + ((int64_t)r * cnt + b) * NF;
                if (NF == 4) { const double2 p = __ldcs(reinterpret_cast<const double2*>(row)), q = __ldcs(reinterpret_cast<const double2*>(row) + 1); s[r] += p.x * xc[0] + p.y * xc[1] + q.x * xc[2] + q.y * xc[NF - 1]; }
                else { for (int f = 0; f < NF; f++) s[r] += row[f] * xc[f]; }
            }
        }
#pragma unroll
        for (int r = 0; r < NF; r++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
        if (lane < NF) {
            double sv = s[0];
#pragma unroll
            for (int r = 1; r < NF; r++) if (lane == r) sv = s[r];
            double* q = y + a * NF + lane;
            *q = (beta == 0.0) ? alpha * sv : alpha * sv + beta * (*q);
        }
    }
}
// scalar CSR (FVCR): warp per row
__global__ void __launch_bounds__(128) csr_spmv_kernel(int64_t row0, int64_t n_row, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                                       const double* __restrict__ val, const double* __restrict__ x, double alpha, double beta,
                                                       double* __restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = row0 + warp0; r < n_row; r += nwarp) {
        double s = 0.0;
        for (int64_t q = rowptr[r] + lane; q < rowptr[r + 1]; q += 32) s += __ldcs(val + q) * x[colind[q]];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[r] = (beta == 0.0) ? alpha * s : alpha * s + beta * y[r];
    }
}
// ---- Dirichlet post-pass (ugcore DirichletBoundary::adjust_jacobian / adjust_defect / adjust_solution as used by
// NavierStokesWall, bnd/wall_impl.h:44-70, and NavierStokesInflowFV1, fv1/bnd/inflow_fv1_impl.h:42-82) ----
// scalar row r of the (block) CSR matrix := unit row. rowinfo: FV1 -> block layout, FVCR -> scalar rowptr.
template <int NF>
__global__ void dirichlet_rows_bcsr_kernel(int64_t n, const int64_t* __restrict__ dofs, const int64_t* __restrict__ brow, const int32_t* __restrict__ bcol,
                                           double* __restrict__ val)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarp) {
        const int64_t dof = dofs[i], a = dof / NF; const int rf = (int)(dof - a * NF);
        const int64_t b0 = brow[a]; const int cnt = (int)(brow[a + 1] - b0);
        double* row = val + b0 * (NF * NF) + (int64_t)rf * cnt * NF;
        for (int j = lane; j < cnt * NF; j += 32) { const int b = j / NF, cf = j - b * NF; row[j] = (bcol[b0 + b] == a && cf == rf) ? 1.0 : 0.0; }
    }
}
__global__ void dirichlet_rows_csr_kernel(int64_t n, const int64_t* __restrict__ dofs, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                          double* __restrict__ val)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarp) {
        const int64_t r = dofs[i];
        for (int64_t q = rowptr[r] + lane; q < rowptr[r + 1]; q += 32) val[q] = (colind[q] == r) ? 1.0 : 0.0;
    }
}
__global__ void dirichlet_set_kernel(int64_t n, const int64_t* __restrict__ dofs, const double* __restrict__ g, double* __restrict__ v)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[dofs[i]] = g ? g[i] : 0.0;
}
__global__ void pack_kernel(int64_t n, const int64_t* __restrict__ idx, const double* __restrict__ src, double* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = src[idx[i]];
}
__global__ void unpack_add_kernel(int64_t n, const int64_t* __restrict__ idx, const double* __restrict__ in, double* __restrict__ dst)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[idx[i]] += in[i];
}

}

static std::string g_create_error;

#define NSB_ASYNC_CHUNKS 4      // row chunks of the NSB_HOST_ASYNC product (the result leaves chunk by chunk)
struct nsb_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;
    nsb_params prm;
    bool mesh_ready = false;
    int elem = -1, disc = NSB_DISC_FV1;
    int64_t n_elem = 0, n_node = 0, n_side = 0, n_dof = 0, nnz = 0;
    int n_colors = 0, max_cnt = 0;
    // host copies needed after upload
    std::vector<int64_t> h_brow;          // block-row prefix (FV1) / scalar rowptr (FVCR)
    std::vector<int32_t> h_bcol;          // block columns (FV1) / scalar colind (FVCR)
    std::vector<int64_t> h_color_ptr;
    // device
    int32_t *d_conn = nullptr, *d_adj = nullptr, *d_color_order = nullptr, *d_esides = nullptr, *d_node_order = nullptr;
    double *d_coords = nullptr, *d_scvvol = nullptr, *d_rec = nullptr;
    size_t rec_bytes = 0; int rec_stride = 0;     // combined SCVF record table [geometry | flux] of the owner-computes path
    bool rec_lean = false;                        // ... or lean flux records of the split path (no geometry half)
    double* d_j0 = nullptr; int j0_laplace = -1;  // static Jacobian part of the split path (ns_split.cuh), built once per mesh
    int64_t *d_brow = nullptr, *d_adj_ptr = nullptr;
    uint8_t *d_emap = nullptr;
    FvcrDev fvcr{};
    int *d_err = nullptr;
    unsigned long long* d_counter = nullptr;
    // staging for NSB_HOST calls
    double *d_u = nullptr, *d_s0 = nullptr, *d_s1 = nullptr, *d_val = nullptr, *d_def = nullptr;
    double *d_jloc = nullptr, *d_dloc = nullptr;
    int64_t launches = 0;
    int last_scatter = -1;                        // scatter mode that served the last assembly (NSB_Q_LAST_SCATTER)
    int64_t n_prio = 0;                           // nsb_set_priority_nodes: d_node_order = [priority nodes | the rest], assembled in two phases on request
    int64_t dev_bytes = 0;                        // device memory held by the context (grid tables, caches, staging)
    double setup_seconds = 0.0;                   // host preprocessing + table upload of the last nsb_upload_mesh*
    int sm_count = 148;
    // fused patch kernel (ns_fused.cuh): per-patch tables, node volumes, per-element ray-search flags
    bool fused_ok = false;
    std::string fused_note;
    PatchHdr* d_phdr = nullptr; PatchNode* d_pnodes = nullptr; int32_t* d_pelems = nullptr; int32_t* d_pconn = nullptr;
    uint32_t* d_pwork = nullptr; PatchAdj* d_padj = nullptr; double* d_nodevol = nullptr; uint8_t* d_elem_fast = nullptr;
    double* d_geo = nullptr; int geo_diff_len = -1;   // static SCVF geometry records of the fused kernel (per diffusion-length type)
    int32_t n_patch = 0; int max_adj = 0;
    int64_t scvf_evals = 0, patch_table_bytes = 0;
    double* d_ip[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // per-ip data imports (nsb_set_ip_data), indexed by NSB_IP_*
    // GPU-resident Jacobian hand-off + Dirichlet post-pass
    int32_t* d_bcol = nullptr; int64_t* d_rowptr = nullptr;   // block columns (FV1) / scalar pattern (FVCR), uploaded on first use
    double* d_jres = nullptr;                                  // resident CSR values (nsb_assemble_resident)
    double *d_xin = nullptr, *d_yout = nullptr;                // staging of nsb_apply_jacobian(NSB_HOST)
    // NSB_HOST_ASYNC: copy streams + events that order the staging buffers between the context stream and the copies
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_main = nullptr, ev_x_in = nullptr, ev_spmv = nullptr, ev_def_out = nullptr, ev_y_out = nullptr, ev_chunk[NSB_ASYNC_CHUNKS] = {};
    int64_t* d_dir = nullptr; int64_t n_dir = 0; double* d_dirval = nullptr;
    // boundary faces of the boundary discs (ns_bnd.cuh), per kind: BFs sorted by grid node
    struct BndSet { int64_t n_bnode = 0, n_bf = 0; int32_t* d_bnode = nullptr; int64_t* d_bptr = nullptr; nsb::BndFace* d_bf = nullptr; double* d_data = nullptr; };
    BndSet bnd[3];
    int32_t* d_sadj = nullptr; double* d_sgrad = nullptr; uint8_t* d_zgrad = nullptr;   // DiscConstraintFVCR (ns_crc.cuh): side -> element adjacency, side gradients, zero-gradient sides
    int32_t* d_bidx = nullptr; double* d_dbf = nullptr; uint8_t* d_zflag = nullptr; double* d_nut = nullptr; double* d_diag = nullptr;   // turbulent viscosity / diagnostics scratch
    // fused tile kernel (ns_tile.cuh, 3-D element types): the patch tables above built with the tile capacities + local-node tables
    bool tile_ok = false;
    int32_t* d_plnodes = nullptr; uint8_t* d_pecorner = nullptr;
};

static int set_err(nsb_ctx* c, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
#define CUDA_TRY(c, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
    return set_err(c, NSB_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); } while (0)

static const int kNSH[5] = {3, 4, 4, 8, 6}, kDIM[5] = {2, 2, 3, 3, 3}, kNSIDE[5] = {3, 4, 4, 6, 5};

// ------------------------------------------------------------------------------------------------
extern "C" const char* nsb_version(void) { return "nsb200 0.1 (sm_100a)"; }

extern "C" const char* nsb_last_error(const nsb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" void nsb_params_default(nsb_params* p)
{
    memset(p, 0, sizeof *p);
    p->disc = NSB_DISC_FV1;
    p->conv_upwind = NSB_UPWIND_UNSET;      // no default upwind / stabilisation
    p->stab = NSB_STAB_UNSET;
    p->stab_upwind = NSB_UPWIND_UNSET;
    p->diff_length = NSB_DIFF_RAW;          // stabilization.h:324
    p->defect_upwind = 1;                   // fvcr/navier_stokes_fvcr.cpp:82
    p->density = 1.0; p->density_set = 1;   // fv1/navier_stokes_fv1.cpp:82
    p->exact_jacobian = 0.0;                // navier_stokes_base.cpp:57
}

extern "C" int nsb_create(int device, nsb_ctx** out)
{
    if (!out) return set_err(nullptr, NSB_ERR_INVALID, "nsb_create: out == NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(nullptr, NSB_ERR_CUDA, "nsb_create: no CUDA device available (%s); this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return set_err(nullptr, NSB_ERR_INVALID, "nsb_create: device %d out of range", device);
    nsb_ctx* c = new nsb_ctx();
    c->device = device;
    nsb_params_default(&c->prm);
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&c->d_err, sizeof(int)) != cudaSuccess || cudaMalloc(&c->d_counter, sizeof(unsigned long long)) != cudaSuccess) {
        set_err(nullptr, NSB_ERR_CUDA, "nsb_create: cannot initialise device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        delete c; return NSB_ERR_CUDA;
    }
    cudaMemset(c->d_err, 0, sizeof(int));
    c->stream = c->own_stream;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = c;
    return NSB_OK;
}

static void fvcr_free(FvcrDev& f)
{
    cudaFree(f.srow); cudaFree(f.sadj_ptr); cudaFree(f.scnt); cudaFree(f.psort); cudaFree(f.emap); cudaFree(f.pslot);
    f = FvcrDev{};
}

static void free_mesh(nsb_ctx* c)
{
    cudaFree(c->d_conn); cudaFree(c->d_adj); cudaFree(c->d_color_order); cudaFree(c->d_esides); cudaFree(c->d_node_order); cudaFree(c->d_coords);
    cudaFree(c->d_scvvol); cudaFree(c->d_rec); cudaFree(c->d_brow); cudaFree(c->d_adj_ptr); cudaFree(c->d_emap);
    cudaFree(c->d_u); cudaFree(c->d_s0); cudaFree(c->d_s1); cudaFree(c->d_val); cudaFree(c->d_def);
    cudaFree(c->d_jloc); cudaFree(c->d_dloc); cudaFree(c->d_j0);
    cudaFree(c->d_phdr); cudaFree(c->d_pnodes); cudaFree(c->d_pelems); cudaFree(c->d_pconn); cudaFree(c->d_pwork); cudaFree(c->d_padj);
    cudaFree(c->d_nodevol); cudaFree(c->d_elem_fast); cudaFree(c->d_geo);
    for (int i = 0; i < 5; i++) { cudaFree(c->d_ip[i]); c->d_ip[i] = nullptr; }
    cudaFree(c->d_bcol); cudaFree(c->d_rowptr); cudaFree(c->d_jres); cudaFree(c->d_xin); cudaFree(c->d_yout); cudaFree(c->d_dir); cudaFree(c->d_dirval);
    c->d_bcol = nullptr; c->d_rowptr = nullptr; c->d_jres = nullptr; c->d_xin = c->d_yout = nullptr; c->d_dir = nullptr; c->n_dir = 0; c->d_dirval = nullptr;
    cudaFree(c->d_sadj); cudaFree(c->d_sgrad); cudaFree(c->d_zgrad); c->d_sadj = nullptr; c->d_sgrad = nullptr; c->d_zgrad = nullptr;
    cudaFree(c->d_bidx); cudaFree(c->d_dbf); cudaFree(c->d_zflag); cudaFree(c->d_nut); cudaFree(c->d_diag); c->d_bidx = nullptr; c->d_dbf = nullptr; c->d_zflag = nullptr; c->d_nut = nullptr; c->d_diag = nullptr;
    for (auto& b : c->bnd) { cudaFree(b.d_bnode); cudaFree(b.d_bptr); cudaFree(b.d_bf); cudaFree(b.d_data); b = nsb_ctx::BndSet(); }
    cudaFree(c->d_plnodes); cudaFree(c->d_pecorner); c->d_plnodes = nullptr; c->d_pecorner = nullptr; c->tile_ok = false;
    c->d_geo = nullptr; c->geo_diff_len = -1;
    c->d_phdr = nullptr; c->d_pnodes = nullptr; c->d_pelems = c->d_pconn = nullptr; c->d_pwork = nullptr; c->d_padj = nullptr;
    c->d_nodevol = nullptr; c->d_elem_fast = nullptr; c->fused_ok = false; c->n_patch = 0; c->scvf_evals = 0; c->patch_table_bytes = 0;
    c->dev_bytes = 0;
    c->d_j0 = nullptr; c->j0_laplace = -1; c->rec_lean = false; c->n_prio = 0;

    fvcr_free(c->fvcr);
    c->d_conn = c->d_adj = c->d_color_order = c->d_esides = c->d_node_order = nullptr; c->d_coords = c->d_scvvol = c->d_rec = nullptr; c->rec_bytes = 0; c->rec_stride = 0;
    c->d_brow = c->d_adj_ptr = nullptr; c->d_emap = nullptr;
    c->d_u = c->d_s0 = c->d_s1 = c->d_val = c->d_def = c->d_jloc = c->d_dloc = nullptr;
    c->mesh_ready = false;
}

extern "C" void nsb_destroy(nsb_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_mesh(c);
    cudaFree(c->d_err); cudaFree(c->d_counter);
    if (c->s_h2d) { cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_d2h); cudaEventDestroy(c->ev_main); cudaEventDestroy(c->ev_x_in);
                    cudaEventDestroy(c->ev_spmv); cudaEventDestroy(c->ev_def_out); cudaEventDestroy(c->ev_y_out);
                    for (int i = 0; i < NSB_ASYNC_CHUNKS; i++) cudaEventDestroy(c->ev_chunk[i]); }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int nsb_set_stream(nsb_ctx* c, void* s)
{
    if (!c) return NSB_ERR_INVALID;
    if ((cudaStream_t)s == c->stream) return NSB_OK;
    // work queued on the old stream (table builds, cached J0) must be visible to kernels on the new one
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaEvent_t ev;
    CUDA_TRY(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventRecord(ev, c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent((cudaStream_t)s, ev, 0));
    cudaEventDestroy(ev);
    c->stream = (cudaStream_t)s;          // NULL is the legacy default stream
    return NSB_OK;
}

extern "C" int nsb_set_params(nsb_ctx* c, const nsb_params* p)
{
    if (!c || !p) return NSB_ERR_INVALID;
    c->prm = *p;
    return NSB_OK;
}

extern "C" int64_t nsb_num_dofs(const nsb_ctx* c) { return c ? c->n_dof : 0; }
extern "C" int64_t nsb_nnz(const nsb_ctx* c) { return c ? c->nnz : 0; }
extern "C" int nsb_num_colors(const nsb_ctx* c) { return c ? c->n_colors : 0; }
extern "C" int64_t nsb_launch_count(const nsb_ctx* c) { return c ? c->launches : 0; }
extern "C" int nsb_query(const nsb_ctx* c, int what, double* out)
{
    if (!c || !out) return NSB_ERR_INVALID;
    switch (what) {
        case NSB_Q_DEVICE_BYTES: *out = (double)c->dev_bytes; break;
        case NSB_Q_SETUP_SECONDS: *out = c->setup_seconds; break;
        case NSB_Q_FUSED: *out = c->fused_ok ? 1.0 : (c->tile_ok ? 2.0 : 0.0); break;
        case NSB_Q_PATCHES: *out = (double)c->n_patch; break;
        case NSB_Q_SCVF_EVALS: *out = (double)c->scvf_evals; break;
        case NSB_Q_PATCH_TABLE_BYTES: *out = (double)c->patch_table_bytes; break;
        case NSB_Q_LAST_SCATTER: *out = (double)c->last_scatter; break;
        default: return NSB_ERR_INVALID;
    }
    return NSB_OK;
}
extern "C" int nsb_synchronize(nsb_ctx* c)
{
    if (!c) return NSB_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->s_h2d) { CUDA_TRY(c, cudaStreamSynchronize(c->s_h2d)); CUDA_TRY(c, cudaStreamSynchronize(c->s_d2h)); }   // NSB_HOST_ASYNC copies
    return NSB_OK;
}

// greedy colouring: no two elements of a colour share an entity; order = colour-major, element-minor.
// Colours 0..63 live in one 64-bit mask per entity; an entity whose elements need more (valence > 64 fans) gets extra
// mask words on demand, so the colour count is unbounded and no two conflicting elements ever share a launch.
static int color_elements(int64_t n_elem, int64_t n_ent, int per, const int32_t* conn, std::vector<int32_t>& order, std::vector<int64_t>& cptr)
{
    std::vector<uint64_t> mask(n_ent, 0);
    std::unordered_map<int64_t, std::vector<uint64_t>> ext;       // entity -> masks of the colours 64.., grown on demand
    std::vector<int32_t> col(n_elem);
    int ncol = 0;
    for (int64_t e = 0; e < n_elem; e++) {
        uint64_t used = 0;
        for (int k = 0; k < per; k++) used |= mask[conn[e * per + k]];
        int cc = 0;
        if (used != ~(uint64_t)0) { while ((used >> cc) & 1) cc++; }
        else {
            // all of 0..63 taken at the element's entities: search the extension words
            for (int w = 0;; w++) {
                uint64_t u = 0;
                for (int k = 0; k < per; k++) { auto it = ext.find(conn[e * per + k]); if (it != ext.end() && (size_t)w < it->second.size()) u |= it->second[w]; }
                if (u != ~(uint64_t)0) { int b = 0; while ((u >> b) & 1) b++; cc = 64 * (w + 1) + b; break; }
            }
        }
        col[e] = cc; ncol = std::max(ncol, cc + 1);
        for (int k = 0; k < per; k++) {
            const int64_t en = conn[e * per + k];
            if (cc < 64) mask[en] |= (uint64_t)1 << cc;
            else { auto& v = ext[en]; const size_t w = (size_t)(cc / 64 - 1); if (v.size() <= w) v.resize(w + 1, 0); v[w] |= (uint64_t)1 << (cc % 64); }
        }
    }
    cptr.assign(ncol + 1, 0);
    for (int64_t e = 0; e < n_elem; e++) cptr[col[e] + 1]++;
    for (int i = 0; i < ncol; i++) cptr[i + 1] += cptr[i];
    order.resize(n_elem);
    std::vector<int64_t> pos(cptr.begin(), cptr.end() - 1);
    for (int64_t e = 0; e < n_elem; e++) order[pos[col[e]]++] = (int32_t)e;
    return ncol;
}

// Z-curve (Morton) order of the grid nodes: the rows kernel walks the nodes in this order so that the SCVF records
// shared by neighbouring nodes (in every coordinate direction) are re-read from L2 instead of HBM.
static void morton_order(int64_t n, int dim, const double* coords, std::vector<int32_t>& order)
{
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = 0; i < n; i++) for (int d = 0; d < dim; d++) { lo[d] = std::min(lo[d], coords[i * dim + d]); hi[d] = std::max(hi[d], coords[i * dim + d]); }
    const int bits = dim == 3 ? 20 : 30;
    double sc[3];
    for (int d = 0; d < dim; d++) sc[d] = hi[d] > lo[d] ? ((double)((1u << bits) - 1)) / (hi[d] - lo[d]) : 0.0;
    // one common scale keeps the curve isotropic on anisotropic boxes
    double smin = 1e300; for (int d = 0; d < dim; d++) if (sc[d] > 0) smin = std::min(smin, sc[d]);
    std::vector<std::pair<uint64_t, int32_t>> key(n);
    parallel_for(n, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; i++) {
            uint64_t code = 0;
            uint32_t q[3] = {0, 0, 0};
            for (int d = 0; d < dim; d++) q[d] = (uint32_t)((coords[i * dim + d] - lo[d]) * smin);
            for (int bit = 0; bit < bits; bit++) for (int d = 0; d < dim; d++) code |= (uint64_t)((q[d] >> bit) & 1u) << (bit * dim + d);
            key[i] = {code, (int32_t)i};
        }
    });
    std::sort(key.begin(), key.end());
    order.resize(n);
    for (int64_t i = 0; i < n; i++) order[i] = key[i].second;
}

// Band order (NSB_NODE_BAND=W, 3-D): nodes sorted by (band of W cells in y, z, y, x). Walking z inside a y-band keeps the records of
// one element layer of the band (W x 0.4 MB at 184^3) in L2 until the next node plane reads them a second time, where the natural
// (z, y, x) order re-reads a whole plane of records (71 MB at 184^3) from HBM.
static void band_order(int64_t n, int dim, const double* coords, int band, std::vector<int32_t>& order)
{
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, vol = 1.0;
    for (int64_t i = 0; i < n; i++) for (int d = 0; d < dim; d++) { lo[d] = std::min(lo[d], coords[i * dim + d]); hi[d] = std::max(hi[d], coords[i * dim + d]); }
    for (int d = 0; d < dim; d++) vol *= std::max(hi[d] - lo[d], 1e-300);
    const double h = std::pow(vol / (double)n, 1.0 / dim), inv = 1.0 / h;
    std::vector<std::pair<uint64_t, int32_t>> key(n);
    parallel_for(n, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; i++) {
            uint64_t q[3] = {0, 0, 0};
            for (int d = 0; d < dim; d++) q[d] = std::min<uint64_t>((uint64_t)((coords[i * dim + d] - lo[d]) * inv + 0.5), (1u << 17) - 1);
            key[i] = {((q[1] / (uint64_t)band) << 51) | (q[2] << 34) | (q[1] << 17) | q[0], (int32_t)i};
        }
    });
    std::sort(key.begin(), key.end());
    order.resize(n);
    for (int64_t i = 0; i < n; i++) order[i] = key[i].second;
}

template <class T> static cudaError_t dev_malloc(nsb_ctx* c, T** dptr, size_t bytes)
{
    cudaError_t e = cudaMalloc((void**)dptr, std::max<size_t>(bytes, 1));
    if (e == cudaSuccess) c->dev_bytes += (int64_t)std::max<size_t>(bytes, 1);
    return e;
}
template <class T> static cudaError_t upload(nsb_ctx* c, T** dptr, const T* h, size_t n)
{
    cudaError_t e = dev_malloc(c, dptr, n * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice);
}

static cudaError_t launch_scvvol(nsb_ctx* c)
{
    c->launches++;
    switch (c->elem) {
        case 0: return launch_scvvol_0(c->n_elem, c->d_conn, c->d_coords, c->d_scvvol, c->stream);
        case 1: return launch_scvvol_1(c->n_elem, c->d_conn, c->d_coords, c->d_scvvol, c->stream);
        case 2: return launch_scvvol_2(c->n_elem, c->d_conn, c->d_coords, c->d_scvvol, c->stream);
        case NSB_PRISM: return launch_scvvol_4(c->n_elem, c->d_conn, c->d_coords, c->d_scvvol, c->stream);
        default: return launch_scvvol_3(c->n_elem, c->d_conn, c->d_coords, c->d_scvvol, c->stream);
    }
}

// per-patch tables of the fused kernel (ns_patch.h), node volumes, per-element ray-search flags. A grid the patch builder
// cannot handle leaves fused_ok = false (the two-kernel split path is used) and the reason in fused_note.
static int setup_fused(nsb_ctx* c, const int32_t* conn, const double* coords, const EntityGraph& g, const std::vector<uint8_t>& emap)
{
    c->fused_ok = false; c->tile_ok = false; c->fused_note.clear();
    // 2-D element types: fused patch kernel (ns_fused.cuh; quad 1024^2 1.15 vs 1.44 ms, tri 1.72 vs 2.01 ms against the split path).
    // 3-D element types: fused tile kernel (ns_tile.cuh, two CTAs per SM, compressed records); NSB_TILE=0 selects the two-kernel
    // split path, NSB_FUSED=1 the patch kernel of ns_fused.cuh. NSB_FUSED=0 disables every fused kernel.
    const bool three_d = (c->elem == NSB_HEX || c->elem == NSB_TET);
    const char* evf = getenv("NSB_FUSED"); const int modef = evf ? atoi(evf) : -1;
    const char* evt = getenv("NSB_TILE"); const int modet = evt ? atoi(evt) : -1;
    if (modef == 0) { c->fused_note = "disabled by NSB_FUSED=0"; return NSB_OK; }
    const bool want_tile = three_d;
    if (want_tile && modet != 1) { c->fused_note = "3-D element types: two-kernel split path (NSB_TILE=1 selects the fused tile kernel)"; return NSB_OK; }
    size_t smem = 0; PatchCaps caps;
    if (want_tile) {
        const int mc = c->elem == NSB_HEX ? tile_max_cnt_3() : tile_max_cnt_2();
        if (g.max_cnt > mc) { c->fused_note = "a block row has more column slots than the tile kernel has lanes: split path"; return NSB_OK; }
        if (c->elem == NSB_HEX) { smem = tile_smem_bytes_3(); caps = tile_caps_3(); } else { smem = tile_smem_bytes_2(); caps = tile_caps_2(); }
    } else {
        switch (c->elem) { case 0: smem = fused_smem_bytes_0(g.max_cnt); caps = fused_caps_0(); break; case 1: smem = fused_smem_bytes_1(g.max_cnt); caps = fused_caps_1(); break;
                           case 2: smem = fused_smem_bytes_2(g.max_cnt); caps = fused_caps_2(); break; default: smem = fused_smem_bytes_3(g.max_cnt); caps = fused_caps_3(); }
    }
    if (smem > 227 * 1024) { c->fused_note = "block rows too long for the shared-memory accumulators"; return NSB_OK; }
    PatchPlan plan; std::string perr;
    if (!build_patch_plan(c->elem, c->n_elem, c->n_node, conn, coords, g.adj_ptr.data(), g.adj.data(), g.brow.data(), emap.data(), caps, plan, perr)) {
        c->fused_note = perr; return NSB_OK;
    }
    CUDA_TRY(c, upload(c, &c->d_phdr, plan.hdr.data(), plan.hdr.size()));
    CUDA_TRY(c, upload(c, &c->d_pnodes, plan.nodes.data(), plan.nodes.size()));
    CUDA_TRY(c, upload(c, &c->d_pelems, plan.elems.data(), plan.elems.size()));
    if (!want_tile) CUDA_TRY(c, upload(c, &c->d_pconn, plan.pconn.data(), plan.pconn.size()));
    CUDA_TRY(c, upload(c, &c->d_pwork, plan.work.data(), plan.work.size()));
    CUDA_TRY(c, upload(c, &c->d_padj, plan.adj.data(), plan.adj.size()));
    if (want_tile) {
        CUDA_TRY(c, upload(c, &c->d_plnodes, plan.lnodes.data(), plan.lnodes.size()));
        CUDA_TRY(c, upload(c, &c->d_pecorner, plan.ecorner.data(), plan.ecorner.size()));
    }
    c->n_patch = (int32_t)plan.hdr.size(); c->max_adj = plan.max_adj_per_node; c->scvf_evals = plan.n_scvf_evals;
    c->patch_table_bytes = (int64_t)(plan.hdr.size() * sizeof(PatchHdr) + plan.nodes.size() * sizeof(PatchNode) + (plan.elems.size() + (want_tile ? plan.lnodes.size() : plan.pconn.size())) * 4 +
                                     plan.work.size() * 4 + plan.adj.size() * sizeof(PatchAdj) + plan.ecorner.size());
    CUDA_TRY(c, dev_malloc(c, &c->d_nodevol, (size_t)c->n_node * sizeof(double)));
    node_volume_kernel<<<(unsigned)((c->n_node + 255) / 256), 256, 0, c->stream>>>(c->n_node, c->d_adj_ptr, c->d_adj, c->d_scvvol, c->d_nodevol);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    if (want_tile) c->tile_ok = true; else c->fused_ok = true;
    return NSB_OK;
}

extern "C" int nsb_upload_mesh(nsb_ctx* c, int elem, int64_t n_elem, int64_t n_node, const int32_t* conn, const double* coords)
{
    const auto t_start = std::chrono::steady_clock::now();
    if (!c) return NSB_ERR_INVALID;
    if (elem < 0 || elem > NSB_PRISM || n_elem <= 0 || n_node <= 0 || !conn || !coords) return set_err(c, NSB_ERR_INVALID, "nsb_upload_mesh: bad arguments");
    CUDA_TRY(c, cudaSetDevice(c->device));
    free_mesh(c);
    const int nsh = kNSH[elem], dim = kDIM[elem], nf = dim + 1;
    c->elem = elem; c->disc = NSB_DISC_FV1; c->n_elem = n_elem; c->n_node = n_node; c->n_side = 0;
    EntityGraph g;
    { const std::string ge = build_entity_graph(n_elem, n_node, nsh, conn, g);
      if (!ge.empty()) return set_err(c, ge.find("too large") != std::string::npos ? NSB_ERR_UNSUPPORTED : NSB_ERR_INVALID, "%s", ge.c_str()); }
    if (g.max_cnt > 255) return set_err(c, NSB_ERR_UNSUPPORTED, "a node has %d neighbours (> 255)", g.max_cnt);
    if ((double)n_node * nf >= 2147483647.0) return set_err(c, NSB_ERR_UNSUPPORTED, "grid has %lld dofs: column indices are 32-bit", (long long)(n_node * nf));
    std::vector<uint8_t> emap; build_emap(n_elem, nsh, conn, g, emap);
    std::vector<int32_t> order;
    c->n_colors = color_elements(n_elem, n_node, nsh, conn, order, c->h_color_ptr);
    c->max_cnt = g.max_cnt;
    c->n_dof = n_node * nf; c->nnz = g.brow[n_node] * nf * nf;
    CUDA_TRY(c, upload(c, &c->d_conn, conn, (size_t)n_elem * nsh));
    CUDA_TRY(c, upload(c, &c->d_coords, coords, (size_t)n_node * dim));
    CUDA_TRY(c, upload(c, &c->d_brow, g.brow.data(), g.brow.size()));
    CUDA_TRY(c, upload(c, &c->d_adj_ptr, g.adj_ptr.data(), g.adj_ptr.size()));
    CUDA_TRY(c, upload(c, &c->d_adj, g.adj.data(), g.adj.size()));
    CUDA_TRY(c, upload(c, &c->d_emap, emap.data(), emap.size()));
    CUDA_TRY(c, upload(c, &c->d_color_order, order.data(), order.size()));
    { std::vector<int32_t> zo;
      const char* bev = getenv("NSB_NODE_BAND"); const int band = bev ? atoi(bev) : 0;
      if (band > 0 && dim == 3) band_order(n_node, dim, coords, band, zo); else morton_order(n_node, dim, coords, zo);
      CUDA_TRY(c, upload(c, &c->d_node_order, zo.data(), zo.size())); }
    CUDA_TRY(c, dev_malloc(c, &c->d_scvvol, (size_t)n_elem * nsh * sizeof(double)));
    CUDA_TRY(c, launch_scvvol(c));
    if (c->elem == NSB_HEX && !(getenv("NSB_RAYFAST") && atoi(getenv("NSB_RAYFAST")) == 0)) {
        // per-element flag: star-shaped w.r.t. every ip -> the upwind ray search may start with the predicted side
        CUDA_TRY(c, dev_malloc(c, &c->d_elem_fast, (size_t)c->n_elem));
        CUDA_TRY(c, launch_ray_safety_3(c->n_elem, c->d_conn, c->d_coords, c->d_elem_fast, c->stream));
        c->launches++;
    }
    if (c->elem != NSB_PRISM) { const int rcf = setup_fused(c, conn, coords, g, emap); if (rcf) return rcf; }   // prisms: element kernels only
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->h_brow.swap(g.brow); c->h_bcol.swap(g.bcol);
    c->mesh_ready = true;
    c->setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    return NSB_OK;
}

extern "C" int nsb_get_csr(const nsb_ctx* c, int64_t* rowptr, int32_t* colind)
{
    if (!c || !c->mesh_ready || !rowptr || !colind) return NSB_ERR_INVALID;
    if (c->disc == NSB_DISC_FVCR) {
        std::copy(c->h_brow.begin(), c->h_brow.end(), rowptr);
        std::copy(c->h_bcol.begin(), c->h_bcol.end(), colind);
        return NSB_OK;
    }
    const int nf = kDIM[c->elem] + 1;
    const int64_t n = c->n_node;
    parallel_for(n, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) {
            const int64_t b0 = c->h_brow[i], cnt = c->h_brow[i + 1] - b0;
            for (int rf = 0; rf < nf; rf++) {
                int64_t pos = b0 * nf * nf + rf * cnt * nf;
                rowptr[i * nf + rf] = pos;
                for (int64_t q = 0; q < cnt; q++) for (int cf = 0; cf < nf; cf++) colind[pos++] = c->h_bcol[b0 + q] * nf + cf;
            }
        }
    });
    rowptr[n * nf] = c->nnz;
    return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// prep_elem_loop validation + parameter resolution
// ------------------------------------------------------------------------------------------------
static int resolve_params(nsb_ctx* c, KParams& k, int what, const nsb_time_series* ts, double sa, double sm)
{
    const nsb_params& p = c->prm;
    memset(&k, 0, sizeof k);
    if (c->disc == NSB_DISC_FV1) {
        // fv1/navier_stokes_fv1.cpp:136-181
        if (p.stab == NSB_STAB_UNSET) return set_err(c, NSB_ERR_SETUP, "Stabilization has not been set.");
        if (p.stab < 0 || p.stab > 2) return set_err(c, NSB_ERR_INVALID, "unknown stabilization id %d", p.stab);
        int upw_stab = p.stab_upwind;
        if (p.pac_upwind) {                      // fv1/navier_stokes_fv1.h:217-225
            if (p.conv_upwind == NSB_UPWIND_UNSET) return set_err(c, NSB_ERR_SETUP, "Upwind must be specified previously.");
            upw_stab = p.conv_upwind;
        } else if (upw_stab == NSB_UPWIND_UNSET) upw_stab = p.conv_upwind;   // string overloads auto-wire (:190-215)
        if (!p.stokes) {
            if (!p.pac_upwind && p.conv_upwind == NSB_UPWIND_UNSET)
                return set_err(c, NSB_ERR_SETUP, "Upwinding for convective Term in Momentum eq. not set.");
            if (upw_stab == NSB_UPWIND_UNSET) return set_err(c, NSB_ERR_SETUP, "No upwind object set in the stabilization.");
        }
        k.upw_stab = upw_stab; k.upw_conv = p.conv_upwind; k.stab = p.stab; k.diff_len = p.diff_length;
        k.pac = p.pac_upwind ? 1 : 0;
    } else {
        // fvcr/navier_stokes_fvcr.cpp:145-181
        if (!p.stokes && p.conv_upwind == NSB_UPWIND_UNSET)
            return set_err(c, NSB_ERR_SETUP, "Upwinding for convective Term in Momentum eq. not set.");
        if (!p.stokes && p.conv_upwind == NSB_UPWIND_POSITIVE)
            return set_err(c, NSB_ERR_SETUP, "No update function registered for Geometry (upwind has no Crouzeix-Raviart overload)");
        k.upw_conv = p.conv_upwind; k.upw_stab = p.conv_upwind;
        k.defect_upwind = p.defect_upwind ? 1 : 0;
        k.grad_div = p.grad_div;
    }
    if (!p.kin_visc_set) return set_err(c, NSB_ERR_SETUP, "NavierStokes::prep_elem_loop: Kinematic Viscosity has not been set, but is required.");
    if (!p.density_set) return set_err(c, NSB_ERR_SETUP, "NavierStokes::prep_elem_loop: Density has not been set, but is required.");
    k.stokes = p.stokes ? 1 : 0; k.laplace = p.laplace ? 1 : 0; k.peclet = p.peclet_blend ? 1 : 0;
    k.has_source = p.has_source ? 1 : 0;
    k.what = what;
    k.exact_jac = p.exact_jacobian; k.visc = p.kin_visc; k.rho = p.density; k.inv_rho = 1.0 / p.density;
    k.scale_a = sa; k.scale_m = sm;
    for (int d = 0; d < 3; d++) k.src[d] = p.source[d];
    k.time_dep = (ts && ts->sol0) ? 1 : 0;
    if (k.time_dep) {
        if (!ts->sol1) return set_err(c, NSB_ERR_SETUP, "NavierStokes::add_jac_A_elem:  Stabilization needs exactly two time points.");
        k.dt = ts->dt;
    }
    return NSB_OK;
}

extern "C" int nsb_prep_elem_loop(nsb_ctx* c)
{
    if (!c) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_prep_elem_loop: no grid uploaded");
    KParams k;
    return resolve_params(c, k, 0, nullptr, 1.0, 1.0);
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
static MeshDev mesh_view(const nsb_ctx* c)
{
    MeshDev m;
    m.n_elem = c->n_elem; m.n_node = c->n_node; m.conn = c->d_conn; m.coords = c->d_coords; m.scvvol = c->d_scvvol;
    m.brow = c->d_brow; m.emap = c->d_emap; m.adj_ptr = c->d_adj_ptr; m.adj = c->d_adj; m.max_cnt = c->max_cnt;
    m.node_order = (c->n_prio > 0 || getenv("NSB_ZORDER") || (getenv("NSB_NODE_BAND") && atoi(getenv("NSB_NODE_BAND")) > 0)) ? c->d_node_order : nullptr;   // priority order, or the opt-in Z-curve (measured neutral on B200, profiles/)
    m.node_begin = 0; m.skip_flux = 0;
    { const char* ev = getenv("NSB_L2HINT"); m.l2_hints = ev ? atoi(ev) : 0; }
    m.elem_fast = c->d_elem_fast;
    m.ip_visc = c->d_ip[NSB_IP_KIN_VISC_SCVF]; m.ip_rho_scvf = c->d_ip[NSB_IP_DENSITY_SCVF]; m.ip_rho_scv = c->d_ip[NSB_IP_DENSITY_SCV];
    m.ip_src_scvf = c->d_ip[NSB_IP_SOURCE_SCVF]; m.ip_src_scv = c->d_ip[NSB_IP_SOURCE_SCV];
    { const char* ev = getenv("NSB_TICKET_GROUP"); const int v = ev ? atoi(ev) : 4; m.ticket_group = v >= 1 ? v : 4; }
    return m;
}

static bool needs_dense(const KParams& k)
{
    return !k.stokes && (k.upw_stab == UPW_POSITIVE || (!k.pac && k.upw_conv == UPW_POSITIVE));
}

static int launch_elem(nsb_ctx* c, int sc, const KParams& k, const int32_t* list, int64_t n_list, const double* u,
                       const double* s0, const double* s1, double* val, double* def, double* jl, double* dl)
{
    if (n_list <= 0) return NSB_OK;
    const MeshDev m = mesh_view(c);
    cudaError_t e;
#define NSB_GO(fn) fn(sc, k, m, list, n_list, u, s0, s1, val, def, jl, dl, c->d_err, c->stream)
    if (needs_dense(k)) switch (c->elem) { case 0: e = NSB_GO(launch_dense_0); break; case 1: e = NSB_GO(launch_dense_1); break;
                                           case 2: e = NSB_GO(launch_dense_2); break; case NSB_PRISM: e = NSB_GO(launch_dense_4); break; default: e = NSB_GO(launch_dense_3); }
    else switch (c->elem) { case 0: e = NSB_GO(launch_elem_0); break; case 1: e = NSB_GO(launch_elem_1); break;
                            case 2: e = NSB_GO(launch_elem_2); break; case NSB_PRISM: e = NSB_GO(launch_elem_4); break; default: e = NSB_GO(launch_elem_3); }
#undef NSB_GO
    c->launches++;
    CUDA_TRY(c, e);
    return NSB_OK;
}

static int launch_gather(nsb_ctx* c, const KParams& k, const double* u, const double* s0, const double* s1, double beta,
                         double* val, double* def, int phase = 0)
{
    const MeshDev m = mesh_view(c);
    // phased assembly (nsb_set_priority_nodes): phase 1 = flux kernel + the rows of the priority nodes, phase 2 = the remaining rows
    // from the records of phase 1. Paths without a separate rows kernel do everything in phase 1.
    MeshDev mp = m;
    if (phase == 1 && c->n_prio > 0) mp.n_node = c->n_prio;
    if (phase == 2) { mp.node_begin = c->n_prio; mp.skip_flux = 1; }
    cudaError_t e;
    static const int kNIP[5] = {3, 4, 6, 12, 9};
    const bool flow = k.stab == STAB_FLOW, exact = !k.stokes && k.exact_jac != 0.0;
    // split path (ns_split.cuh): static Jacobian part J0 cached per mesh + lean flux records. FLOW couples the
    // velocity components in the continuity row and exact Newton adds full blocks: those keep the general rows kernel.
    static const bool no_split = getenv("NSB_NOSPLIT") != nullptr;
    // (the split rows kernel keeps JP accumulator copies + the J0 rows per warp in shared memory: bounded row length only)
    // (3-D: the compressed record of the split path carries ONE set of upwind shapes -> same upwind for stabilisation and convection)
    const bool one_upwind = k.stokes || k.upw_conv == k.upw_stab || kDIM[c->elem] == 2;
    const bool lean = !flow && !exact && !no_split && c->max_cnt <= 64 && one_upwind;
    static const bool no_fused = getenv("NSB_NOFUSED") != nullptr;
    const bool use_fused = lean && c->fused_ok && !no_fused;     // fused patch kernel: the SCVF records stay in shared memory
    // fused tile kernel (3-D): one upwind object for stabilisation and convection (the compressed record carries one set of upwind shapes)
    const bool use_tile = lean && c->tile_ok && !no_fused;
    if (!use_fused && !use_tile) {   // per-(element, ip) record table: [static SCVF geometry | flux record] or the lean record of the split path.
        // The stride depends on the stabilisation (FLOW) and Jacobian flavour (exact Newton): (re)built on change.
        int stride = 0;
        if (lean) switch (c->elem) { case 0: stride = lean_record_doubles_0(); break; case 1: stride = lean_record_doubles_1(); break;
                                     case 2: stride = lean_record_doubles_2(); break; default: stride = lean_record_doubles_3(); }
        else switch (c->elem) { case 0: stride = scvf_record_doubles_0(flow, exact); break; case 1: stride = scvf_record_doubles_1(flow, exact); break;
                                case 2: stride = scvf_record_doubles_2(flow, exact); break; default: stride = scvf_record_doubles_3(flow, exact); }
        if (stride != c->rec_stride || lean != c->rec_lean) {
            const size_t need = (size_t)c->n_elem * kNIP[c->elem] * stride * sizeof(double);
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            if (need > c->rec_bytes) {
                cudaFree(c->d_rec); c->d_rec = nullptr; c->rec_bytes = 0; c->rec_stride = 0;
                CUDA_TRY(c, dev_malloc(c, &c->d_rec, need));
                c->rec_bytes = need;
            }
            if (!lean) {
                switch (c->elem) { case 0: e = launch_geom_0(c->n_elem, c->d_conn, c->d_coords, c->d_rec, stride, c->stream); break;
                                   case 1: e = launch_geom_1(c->n_elem, c->d_conn, c->d_coords, c->d_rec, stride, c->stream); break;
                                   case 2: e = launch_geom_2(c->n_elem, c->d_conn, c->d_coords, c->d_rec, stride, c->stream); break;
                                   default: e = launch_geom_3(c->n_elem, c->d_conn, c->d_coords, c->d_rec, stride, c->stream); }
                c->launches++;
                CUDA_TRY(c, e);
            }
            c->rec_stride = stride; c->rec_lean = lean;
        }
    }
    if (lean) {
        const int dim = kDIM[c->elem], nf = dim + 1;
        if ((k.what & W_JAC_A) && c->j0_laplace != k.laplace) {
            const size_t nj0 = (size_t)c->h_brow[c->n_node] * dim * nf;
            if (!c->d_j0) CUDA_TRY(c, dev_malloc(c, &c->d_j0, nj0 * sizeof(double)));
            CUDA_TRY(c, cudaMemsetAsync(c->d_j0, 0, nj0 * sizeof(double), c->stream));
            switch (c->elem) { case 0: e = launch_j0_0(m, k.laplace, c->d_j0, c->stream, c->sm_count); break;
                               case 1: e = launch_j0_1(m, k.laplace, c->d_j0, c->stream, c->sm_count); break;
                               case 2: e = launch_j0_2(m, k.laplace, c->d_j0, c->stream, c->sm_count); break;
                               default: e = launch_j0_3(m, k.laplace, c->d_j0, c->stream, c->sm_count); }
            c->launches++;
            CUDA_TRY(c, e);
            c->j0_laplace = k.laplace;
        }
        if ((use_tile || use_fused) && phase == 2) return NSB_OK;
        if (use_tile) {
            TileArgs A;
            memset(&A, 0, sizeof A);
            A.p = k; A.n_tile = c->n_patch;
            A.hdr = c->d_phdr; A.nodes = c->d_pnodes; A.elems = c->d_pelems; A.lnodes = c->d_plnodes; A.ecorner = c->d_pecorner; A.work = c->d_pwork; A.adj = c->d_padj;
            A.coords = c->d_coords; A.scvvol = c->d_scvvol; A.nodevol = c->d_nodevol;
            A.u = u; A.s0 = s0; A.s1 = s1; A.j0 = c->d_j0; A.beta = beta; A.val = val; A.def = def;
            A.errflag = c->d_err; A.elem_fast = c->d_elem_fast;
            e = c->elem == NSB_HEX ? launch_tile_3(A, c->stream, c->sm_count) : launch_tile_2(A, c->stream, c->sm_count);
            c->launches++;
            CUDA_TRY(c, e);
            return NSB_OK;
        }
        if (use_fused) {
            FusedArgs A;
            memset(&A, 0, sizeof A);
            A.p = k; A.n_patch = c->n_patch;
            A.hdr = c->d_phdr; A.nodes = c->d_pnodes; A.elems = c->d_pelems; A.pconn = c->d_pconn; A.work = c->d_pwork; A.adj = c->d_padj;
            A.coords = c->d_coords; A.scvvol = c->d_scvvol; A.nodevol = c->d_nodevol;
            A.u = u; A.s0 = s0; A.s1 = s1; A.j0 = c->d_j0; A.beta = beta; A.val = val; A.def = def;
            A.errflag = c->d_err; A.elem_fast = c->d_elem_fast; A.max_adj = c->max_adj;
            // static SCVF geometry records (normal, ip, J^-T, 1/L_d^2) in work-item order: 128 B per SCVF evaluation, built once
            // per mesh and diffusion-length type; NSB_GEOTAB=0 recomputes the geometry in every pass instead
            static const bool geotab = !(getenv("NSB_GEOTAB") && atoi(getenv("NSB_GEOTAB")) == 0);
            if (geotab && (k.what & (W_JAC_A | W_DEF_A))) {
                if (!c->d_geo) CUDA_TRY(c, dev_malloc(c, &c->d_geo, (size_t)c->scvf_evals * 16 * sizeof(double)));
                if (c->geo_diff_len != k.diff_len) {
                    switch (c->elem) { case 0: e = launch_fused_geom_0(A, k.diff_len, c->d_geo, c->stream); break;
                                       case 1: e = launch_fused_geom_1(A, k.diff_len, c->d_geo, c->stream); break;
                                       case 2: e = launch_fused_geom_2(A, k.diff_len, c->d_geo, c->stream); break;
                                       default: e = launch_fused_geom_3(A, k.diff_len, c->d_geo, c->stream); }
                    c->launches++;
                    CUDA_TRY(c, e);
                    c->geo_diff_len = k.diff_len;
                }
                A.geo = c->d_geo;
            }
            switch (c->elem) { case 0: e = launch_fused_0(A, c->max_cnt, c->stream, c->sm_count, c->d_counter); break;
                               case 1: e = launch_fused_1(A, c->max_cnt, c->stream, c->sm_count, c->d_counter); break;
                               case 2: e = launch_fused_2(A, c->max_cnt, c->stream, c->sm_count, c->d_counter); break;
                               default: e = launch_fused_3(A, c->max_cnt, c->stream, c->sm_count, c->d_counter); }
            c->launches++;
            CUDA_TRY(c, e);
            return NSB_OK;
        }
#define NSB_GO(fn) fn(k, mp, c->d_rec, u, s0, s1, beta, val, def, c->d_err, c->stream, c->sm_count, c->d_counter, c->d_j0)
        switch (c->elem) { case 0: e = NSB_GO(launch_split_0); break; case 1: e = NSB_GO(launch_split_1); break;
                           case 2: e = NSB_GO(launch_split_2); break; default: e = NSB_GO(launch_split_3); }
#undef NSB_GO
        const bool flux_needed = (k.what & (W_JAC_A | W_DEF_A)) && phase != 2;
        c->launches += flux_needed ? 2 : 1;
        CUDA_TRY(c, e);
        return NSB_OK;
    }
#define NSB_GO(fn) fn(k, mp, c->d_rec, u, s0, s1, beta, val, def, c->d_err, c->stream, c->sm_count, c->d_counter)
    switch (c->elem) { case 0: e = NSB_GO(launch_gather_0); break; case 1: e = NSB_GO(launch_gather_1); break;
                       case 2: e = NSB_GO(launch_gather_2); break; default: e = NSB_GO(launch_gather_3); }
#undef NSB_GO
    c->launches += ((k.what & (W_JAC_A | W_DEF_A)) && phase != 2) ? 2 : 1;
    CUDA_TRY(c, e);
    return NSB_OK;
}

static int assemble_fv1(nsb_ctx* c, const KParams& k, int mode, const double* u, const double* s0, const double* s1,
                        double beta, double* val, double* def, int phase = 0)
{
    const bool jac = k.what & (W_JAC_A | W_JAC_M), dfc = k.what & (W_DEF_A | W_DEF_M | W_RHS);
    const bool ip_data = c->d_ip[0] || c->d_ip[1] || c->d_ip[2] || c->d_ip[3] || c->d_ip[4];
    if (ip_data && needs_dense(k)) return set_err(c, NSB_ERR_UNSUPPORTED, "per-ip data imports with PositiveUpwind (dense ip systems) are not provided on the device path");
    if (mode == NSB_SCATTER_GATHER && (needs_dense(k) || k.pac || ip_data || c->elem == NSB_PRISM)) mode = NSB_SCATTER_COLORED;   // dense ip systems / PAC / per-ip data / prisms: element kernels
    c->last_scatter = mode;
    if (mode == NSB_SCATTER_GATHER) return launch_gather(c, k, u, s0, s1, beta, val, def, phase);
    if (phase == 2) return NSB_OK;                               // element kernels: everything happened in phase 1
    // element kernels accumulate into beta*old
    if (jac) {
        if (beta == 0.0) CUDA_TRY(c, cudaMemsetAsync(val, 0, sizeof(double) * c->nnz, c->stream));
        else if (beta != 1.0) { scale_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->nnz, beta, val); c->launches++; }
    }
    if (dfc) {
        if (beta == 0.0) CUDA_TRY(c, cudaMemsetAsync(def, 0, sizeof(double) * c->n_dof, c->stream));
        else if (beta != 1.0) { scale_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->n_dof, beta, def); c->launches++; }
    }
    if (mode == NSB_SCATTER_ATOMIC) return launch_elem(c, SC_ATOMIC, k, nullptr, c->n_elem, u, s0, s1, val, def, nullptr, nullptr);
    if (mode == NSB_SCATTER_COLORED) {
        for (int col = 0; col < c->n_colors; col++) {
            const int64_t lo = c->h_color_ptr[col], hi = c->h_color_ptr[col + 1];
            int rc = launch_elem(c, SC_COLORED, k, c->d_color_order + lo, hi - lo, u, s0, s1, val, def, nullptr, nullptr);
            if (rc) return rc;
        }
        return NSB_OK;
    }
    return set_err(c, NSB_ERR_INVALID, "unknown scatter mode %d", mode);
}

// FVCR: element kernel with coloured or atomic scatter (both deterministic for beta == 0, see below); the owner-computes path is FV1-only
static int assemble_fvcr(nsb_ctx* c, const KParams& k, int mode, const double* u, double beta, double* val, double* def, int phase = 0)
{
    if (phase == 2) return NSB_OK;
    const bool jac = k.what & (W_JAC_A | W_JAC_M), dfc = k.what & (W_DEF_A | W_DEF_M | W_RHS);
    // A CR velocity dof lives on an element side, which has at most TWO elements; two different sides share at most one element
    // and the pressure dof belongs to one. With beta == 0 every entry is therefore 0 + a (+ b), and a + b == b + a exactly: the
    // single-launch reduction in element order returns the bits of the coloured sweeps (tests/test_gpu_parity_fvcr.py), without
    // streaming the value array once per colour. beta != 0 adds a third term -> coloured sweeps.
    if (mode == NSB_SCATTER_GATHER) mode = (beta == 0.0) ? NSB_SCATTER_ATOMIC : NSB_SCATTER_COLORED;
    c->last_scatter = mode;
    if (jac) {
        if (beta == 0.0) CUDA_TRY(c, cudaMemsetAsync(val, 0, sizeof(double) * c->nnz, c->stream));
        else if (beta != 1.0) { scale_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->nnz, beta, val); c->launches++; }
    }
    if (dfc) {
        if (beta == 0.0) CUDA_TRY(c, cudaMemsetAsync(def, 0, sizeof(double) * c->n_dof, c->stream));
        else if (beta != 1.0) { scale_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->n_dof, beta, def); c->launches++; }
    }
    auto go = [&](int sc, const int32_t* list, int64_t n) -> cudaError_t {
        c->launches++;
        switch (c->elem) {
            case NSB_TRI: return launch_fvcr_0(sc, k, c->fvcr, list, n, u, val, def, c->d_err, c->stream);
            case NSB_QUAD: return launch_fvcrq_1(sc, k, c->fvcr, list, n, u, val, def, c->d_err, c->stream);   // non-affine CR geometry (ns_fvcr_q.cuh)
            case NSB_HEX: return launch_fvcrq_3(sc, k, c->fvcr, list, n, u, val, def, c->d_err, c->stream);
            default: return launch_fvcr_2(sc, k, c->fvcr, list, n, u, val, def, c->d_err, c->stream);
        }
    };
    if (mode == NSB_SCATTER_ATOMIC) { CUDA_TRY(c, go(SC_ATOMIC, nullptr, c->n_elem)); return NSB_OK; }
    for (int col = 0; col < c->n_colors; col++) {
        const int64_t lo = c->h_color_ptr[col], hi = c->h_color_ptr[col + 1];
        if (hi > lo) CUDA_TRY(c, go(SC_COLORED, c->d_color_order + lo, hi - lo));
    }
    return NSB_OK;
}

static int check_device_error(nsb_ctx* c)
{
    int flag = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&flag, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (flag) {
        cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream);
        return set_err(c, NSB_ERR_GEOMETRY, flag == 2 ? "Could not compute inverse." : "GetNodeNextToCut: Cannot find cut side.");
    }
    return NSB_OK;
}

static int ensure(nsb_ctx* c, double** p, size_t n)
{
    if (*p) return NSB_OK;
    CUDA_TRY(c, dev_malloc(c, p, n * sizeof(double)));
    return NSB_OK;
}

// The staging buffers (d_def, d_yout) may still be read by the D2H stream of an earlier NSB_HOST_ASYNC call: order the context
// stream behind those copies before a host-pointer call reuses them.
static int order_after_async_copies(nsb_ctx* c)
{
    if (!c->s_d2h) return NSB_OK;
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_def_out, 0));
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_y_out, 0));
    return NSB_OK;
}

extern "C" int nsb_assemble(nsb_ctx* c, int what, int mode, const double* u, const nsb_time_series* ts, double sa, double sm,
                            double beta, double* values, double* defect, int location)
{
    if (!c) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_assemble: no grid uploaded");
    if (!u) return set_err(c, NSB_ERR_INVALID, "nsb_assemble: u == NULL");
    const int phase = (what & NSB_PHASE_PRIORITY) ? 1 : (what & NSB_PHASE_REST) ? 2 : 0;
    what &= ~(NSB_PHASE_PRIORITY | NSB_PHASE_REST);
    if (location != NSB_HOST && location != NSB_DEVICE) return set_err(c, NSB_ERR_INVALID, "nsb_assemble: location must be NSB_HOST or NSB_DEVICE");
    if (phase && location != NSB_DEVICE) return set_err(c, NSB_ERR_INVALID, "nsb_assemble: the phased assembly works on device pointers");
    const bool jac = what & (NSB_JAC_A | NSB_JAC_M), dfc = what & (NSB_DEF_A | NSB_DEF_M | NSB_RHS);
    if ((jac && !values) || (dfc && !defect)) return set_err(c, NSB_ERR_INVALID, "nsb_assemble: output pointer missing for requested part");
    if (!jac && !dfc) return NSB_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    KParams k;
    int rc = resolve_params(c, k, what, ts, sa, sm);
    if (rc) return rc;
    const double *du = u, *ds0 = ts ? ts->sol0 : nullptr, *ds1 = ts ? ts->sol1 : nullptr;
    double *dv = values, *dd = defect;
    if (location == NSB_HOST) {
        const size_t nb = sizeof(double) * c->n_dof;
        if ((rc = order_after_async_copies(c))) return rc;
        if ((rc = ensure(c, &c->d_u, c->n_dof))) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_u, u, nb, cudaMemcpyHostToDevice, c->stream)); du = c->d_u;
        if (k.time_dep || (c->disc == NSB_DISC_FVCR && ds0)) {
            if ((rc = ensure(c, &c->d_s0, c->n_dof)) || (rc = ensure(c, &c->d_s1, c->n_dof))) return rc;
            CUDA_TRY(c, cudaMemcpyAsync(c->d_s0, ts->sol0, nb, cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(c, cudaMemcpyAsync(c->d_s1, ts->sol1, nb, cudaMemcpyHostToDevice, c->stream));
            ds0 = c->d_s0; ds1 = c->d_s1;
        }
        if (jac) { if ((rc = ensure(c, &c->d_val, c->nnz))) return rc; dv = c->d_val;
                   if (beta != 0.0) CUDA_TRY(c, cudaMemcpyAsync(dv, values, sizeof(double) * c->nnz, cudaMemcpyHostToDevice, c->stream)); }
        if (dfc) { if ((rc = ensure(c, &c->d_def, c->n_dof))) return rc; dd = c->d_def;
                   if (beta != 0.0) CUDA_TRY(c, cudaMemcpyAsync(dd, defect, nb, cudaMemcpyHostToDevice, c->stream)); }
    }
    if (c->disc == NSB_DISC_FVCR) rc = assemble_fvcr(c, k, mode, du, beta, dv, dd, phase);
    else rc = assemble_fv1(c, k, mode, du, ds0, ds1, beta, dv, dd, phase);
    if (rc) return rc;
    if (location == NSB_HOST) {
        if (jac) CUDA_TRY(c, cudaMemcpyAsync(values, dv, sizeof(double) * c->nnz, cudaMemcpyDeviceToHost, c->stream));
        if (dfc) CUDA_TRY(c, cudaMemcpyAsync(defect, dd, sizeof(double) * c->n_dof, cudaMemcpyDeviceToHost, c->stream));
        return check_device_error(c);          // synchronises
    }
    return NSB_OK;
}

extern "C" int nsb_check_errors(nsb_ctx* c)      /* device-pointer mode: poll the element-level error flag */
{
    if (!c) return NSB_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    return check_device_error(c);
}

extern "C" int nsb_local_contributions(nsb_ctx* c, int what, const double* u, const nsb_time_series* ts, double* Jloc,
                                       double* dloc, int location)
{
    if (!c) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_local_contributions: no grid uploaded");
    if (c->disc != NSB_DISC_FV1) return set_err(c, NSB_ERR_UNSUPPORTED, "local contributions: FV1 only");
    if (!u || !Jloc || !dloc) return set_err(c, NSB_ERR_INVALID, "nsb_local_contributions: NULL pointer");
    CUDA_TRY(c, cudaSetDevice(c->device));
    KParams k;
    int rc = resolve_params(c, k, what, ts, 1.0, 1.0);
    if (rc) return rc;
    const int L = kNSH[c->elem] * (kDIM[c->elem] + 1);
    const size_t nJ = (size_t)c->n_elem * L * L, nd = (size_t)c->n_elem * L;
    const double *du = u, *ds0 = ts ? ts->sol0 : nullptr, *ds1 = ts ? ts->sol1 : nullptr;
    double *dj = Jloc, *ddl = dloc;
    if (location == NSB_HOST) {
        const size_t nb = sizeof(double) * c->n_dof;
        if ((rc = ensure(c, &c->d_u, c->n_dof))) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_u, u, nb, cudaMemcpyHostToDevice, c->stream)); du = c->d_u;
        if (k.time_dep) {
            if ((rc = ensure(c, &c->d_s0, c->n_dof)) || (rc = ensure(c, &c->d_s1, c->n_dof))) return rc;
            CUDA_TRY(c, cudaMemcpyAsync(c->d_s0, ts->sol0, nb, cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(c, cudaMemcpyAsync(c->d_s1, ts->sol1, nb, cudaMemcpyHostToDevice, c->stream));
            ds0 = c->d_s0; ds1 = c->d_s1;
        }
        if ((rc = ensure(c, &c->d_jloc, nJ)) || (rc = ensure(c, &c->d_dloc, nd))) return rc;
        dj = c->d_jloc; ddl = c->d_dloc;
    }
    CUDA_TRY(c, cudaMemsetAsync(dj, 0, sizeof(double) * nJ, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(ddl, 0, sizeof(double) * nd, c->stream));
    rc = launch_elem(c, SC_LOCAL, k, nullptr, c->n_elem, du, ds0, ds1, nullptr, nullptr, dj, ddl);
    if (rc) return rc;
    if (location == NSB_HOST) {
        CUDA_TRY(c, cudaMemcpyAsync(Jloc, dj, sizeof(double) * nJ, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(dloc, ddl, sizeof(double) * nd, cudaMemcpyDeviceToHost, c->stream));
        return check_device_error(c);
    }
    return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// GPU-resident Jacobian and Dirichlet post-pass (8f-1)
// ------------------------------------------------------------------------------------------------
static int ensure_pattern(nsb_ctx* c)
{
    if (c->d_bcol) return NSB_OK;
    CUDA_TRY(c, upload(c, &c->d_bcol, c->h_bcol.data(), c->h_bcol.size()));
    if (c->disc == NSB_DISC_FVCR) CUDA_TRY(c, upload(c, &c->d_rowptr, c->h_brow.data(), c->h_brow.size()));
    return NSB_OK;
}

// NSB_HOST_ASYNC: an H2D and a D2H copy stream beside the context stream. An event that was never recorded counts as complete.
static int ensure_async(nsb_ctx* c)
{
    if (c->s_h2d) return NSB_OK;
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&c->ev_main, &c->ev_x_in, &c->ev_spmv, &c->ev_def_out, &c->ev_y_out}) CUDA_TRY(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (int i = 0; i < NSB_ASYNC_CHUNKS; i++) CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming));
    return NSB_OK;
}

extern "C" int nsb_assemble_resident(nsb_ctx* c, int what, int mode, const double* u, const nsb_time_series* ts, double sa, double sm,
                                     double beta, double* defect, int location)
{
    if (!c) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_assemble_resident: no grid uploaded");
    if (!u) return set_err(c, NSB_ERR_INVALID, "nsb_assemble_resident: u == NULL");
    const bool jac = what & (NSB_JAC_A | NSB_JAC_M), dfc = what & (NSB_DEF_A | NSB_DEF_M | NSB_RHS);
    if (dfc && !defect) return set_err(c, NSB_ERR_INVALID, "nsb_assemble_resident: defect pointer missing");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc;
    if (jac && !c->d_jres) {
        if ((rc = ensure(c, &c->d_jres, c->nnz))) return rc;
        CUDA_TRY(c, cudaMemsetAsync(c->d_jres, 0, sizeof(double) * c->nnz, c->stream));
    }
    if (location == NSB_DEVICE) return nsb_assemble(c, what, mode, u, ts, sa, sm, beta, c->d_jres, defect, NSB_DEVICE);
    const bool async = location == NSB_HOST_ASYNC;
    if (async && (rc = ensure_async(c))) return rc;
    if ((rc = order_after_async_copies(c))) return rc;
    // host vectors, device-resident matrix: u (and the time series) in, defect out
    KParams k;
    if ((rc = resolve_params(c, k, what, ts, sa, sm))) return rc;
    const size_t nb = sizeof(double) * c->n_dof;
    if ((rc = ensure(c, &c->d_u, c->n_dof))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_u, u, nb, cudaMemcpyHostToDevice, c->stream));
    nsb_time_series dts; const nsb_time_series* pts = nullptr;
    if (ts && ts->sol0) {
        if ((rc = ensure(c, &c->d_s0, c->n_dof)) || (rc = ensure(c, &c->d_s1, c->n_dof))) return rc;
        if (!ts->sol1) return set_err(c, NSB_ERR_SETUP, "NavierStokes::add_jac_A_elem:  Stabilization needs exactly two time points.");
        CUDA_TRY(c, cudaMemcpyAsync(c->d_s0, ts->sol0, nb, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->d_s1, ts->sol1, nb, cudaMemcpyHostToDevice, c->stream));
        dts.sol0 = c->d_s0; dts.sol1 = c->d_s1; dts.dt = ts->dt; pts = &dts;
    }
    if (dfc) { if ((rc = ensure(c, &c->d_def, c->n_dof))) return rc;
               if (beta != 0.0) CUDA_TRY(c, cudaMemcpyAsync(c->d_def, defect, nb, cudaMemcpyHostToDevice, c->stream)); }
    if ((rc = nsb_assemble(c, what, mode, c->d_u, pts, sa, sm, beta, c->d_jres, c->d_def, NSB_DEVICE))) return rc;
    if (async) {                                                 // defect leaves on the D2H stream behind whatever the caller queues next
        if (dfc) {
            CUDA_TRY(c, cudaEventRecord(c->ev_main, c->stream));
            CUDA_TRY(c, cudaStreamWaitEvent(c->s_d2h, c->ev_main, 0));
            CUDA_TRY(c, cudaMemcpyAsync(defect, c->d_def, nb, cudaMemcpyDeviceToHost, c->s_d2h));
            CUDA_TRY(c, cudaEventRecord(c->ev_def_out, c->s_d2h));
        }
        return NSB_OK;                                           // nsb_synchronize + nsb_check_errors complete the call
    }
    if (dfc) CUDA_TRY(c, cudaMemcpyAsync(defect, c->d_def, nb, cudaMemcpyDeviceToHost, c->stream));
    return check_device_error(c);
}

extern "C" int nsb_resident_jacobian(nsb_ctx* c, double** dev_values)
{
    if (!c || !dev_values) return NSB_ERR_INVALID;
    if (!c->d_jres) return set_err(c, NSB_ERR_INVALID, "nsb_resident_jacobian: nothing assembled yet (nsb_assemble_resident)");
    *dev_values = c->d_jres;
    return NSB_OK;
}

// rows [r0, r1) of y = alpha J x + beta y (block rows for FV1, scalar rows for FVCR); r1 < 0 = all rows
static int spmv_launch(nsb_ctx* c, const double* val, double alpha, const double* x, double beta, double* y, int64_t r0 = 0, int64_t r1 = -1)
{
    int rc = ensure_pattern(c);
    if (rc) return rc;
    const int64_t nrows = c->disc == NSB_DISC_FVCR ? c->n_dof : c->n_node;
    if (r1 < 0) r1 = nrows;
    if (r1 <= r0) return NSB_OK;
    const unsigned nblk = (unsigned)std::min<int64_t>((r1 - r0 + 3) / 4, (int64_t)c->sm_count * 16);
    if (c->disc == NSB_DISC_FVCR) csr_spmv_kernel<<<nblk, 128, 0, c->stream>>>(r0, r1, c->d_rowptr, c->d_bcol, val, x, alpha, beta, y);
    else if (kDIM[c->elem] == 3) bcsr_spmv_kernel<4><<<nblk, 128, 0, c->stream>>>(r0, r1, c->d_brow, c->d_bcol, val, x, alpha, beta, y);
    else bcsr_spmv_kernel<3><<<nblk, 128, 0, c->stream>>>(r0, r1, c->d_brow, c->d_bcol, val, x, alpha, beta, y);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return NSB_OK;
}

extern "C" int nsb_apply_jacobian(nsb_ctx* c, const double* values, double alpha, const double* x, double beta, double* y, int location)
{
    if (!c || !x || !y) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_apply_jacobian: no grid uploaded");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const double* val = values ? values : c->d_jres;        // NULL = the resident Jacobian
    if (!val) return set_err(c, NSB_ERR_INVALID, "nsb_apply_jacobian: no resident Jacobian (nsb_assemble_resident) and values == NULL");
    if (location == NSB_DEVICE) return spmv_launch(c, val, alpha, x, beta, y);
    int rc;
    const size_t nb = sizeof(double) * c->n_dof;
    if ((rc = ensure(c, &c->d_xin, c->n_dof)) || (rc = ensure(c, &c->d_yout, c->n_dof))) return rc;
    if (location == NSB_HOST_ASYNC) {
        // x goes up on the H2D stream while the context stream is still busy (e.g. with the assembly queued before this call);
        // J x leaves on the D2H stream. The staging buffers are handed over through events; nsb_synchronize completes the call.
        if ((rc = ensure_async(c))) return rc;
        CUDA_TRY(c, cudaStreamWaitEvent(c->s_h2d, c->ev_spmv, 0));               // the previous product has read x
        CUDA_TRY(c, cudaMemcpyAsync(c->d_xin, x, nb, cudaMemcpyHostToDevice, c->s_h2d));
        if (beta != 0.0) { CUDA_TRY(c, cudaStreamWaitEvent(c->s_h2d, c->ev_y_out, 0));
                           CUDA_TRY(c, cudaMemcpyAsync(c->d_yout, y, nb, cudaMemcpyHostToDevice, c->s_h2d)); }
        CUDA_TRY(c, cudaEventRecord(c->ev_x_in, c->s_h2d));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_x_in, 0));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_y_out, 0));             // the previous result has left the staging buffer
        // the product runs in row chunks; the rows of a chunk leave on the D2H stream while the next chunk is computed
        const int64_t nrows = c->disc == NSB_DISC_FVCR ? c->n_dof : c->n_node;
        const int64_t per_row = c->disc == NSB_DISC_FVCR ? 1 : kDIM[c->elem] + 1;
        const int nch = nrows >= 64 * NSB_ASYNC_CHUNKS ? NSB_ASYNC_CHUNKS : 1;
        for (int ch = 0; ch < nch; ch++) {
            const int64_t r0 = nrows * ch / nch, r1 = nrows * (ch + 1) / nch;
            if ((rc = spmv_launch(c, val, alpha, c->d_xin, beta, c->d_yout, r0, r1))) return rc;
            cudaEvent_t ev = ch + 1 < nch ? c->ev_chunk[ch] : c->ev_spmv;
            CUDA_TRY(c, cudaEventRecord(ev, c->stream));
            CUDA_TRY(c, cudaStreamWaitEvent(c->s_d2h, ev, 0));
            CUDA_TRY(c, cudaMemcpyAsync(y + r0 * per_row, c->d_yout + r0 * per_row, sizeof(double) * (size_t)((r1 - r0) * per_row), cudaMemcpyDeviceToHost, c->s_d2h));
        }
        CUDA_TRY(c, cudaEventRecord(c->ev_y_out, c->s_d2h));
        return NSB_OK;
    }
    if ((rc = order_after_async_copies(c))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_xin, x, nb, cudaMemcpyHostToDevice, c->stream));
    if (beta != 0.0) CUDA_TRY(c, cudaMemcpyAsync(c->d_yout, y, nb, cudaMemcpyHostToDevice, c->stream));
    if ((rc = spmv_launch(c, val, alpha, c->d_xin, beta, c->d_yout))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(y, c->d_yout, nb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return NSB_OK;
}

// per-ip data imports (fv1/navier_stokes_fv1.cpp:184-197: m_imKinViscosity / m_imDensitySCVF at the SCVF ips, m_imDensitySCV /
// m_imSourceSCV at the SCV ips, m_imSourceSCVF at the SCVF ips). data == NULL returns to the constant of nsb_params.
extern "C" int nsb_set_ip_data(nsb_ctx* c, int kind, const double* data, int location)
{
    if (!c || kind < 0 || kind > 4) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_set_ip_data: no grid uploaded");
    if (c->disc != NSB_DISC_FV1) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_set_ip_data: FV1 only");
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    static const int kNIP[5] = {3, 4, 6, 12, 9};
    const int dim = kDIM[c->elem], nsh = kNSH[c->elem], nip = kNIP[c->elem];
    const size_t per = kind == NSB_IP_KIN_VISC_SCVF || kind == NSB_IP_DENSITY_SCVF ? (size_t)nip : kind == NSB_IP_DENSITY_SCV ? (size_t)nsh
                     : kind == NSB_IP_SOURCE_SCVF ? (size_t)nip * dim : (size_t)nsh * dim;
    if (!data) { cudaFree(c->d_ip[kind]); c->d_ip[kind] = nullptr; return NSB_OK; }
    if (!c->d_ip[kind]) CUDA_TRY(c, dev_malloc(c, &c->d_ip[kind], (size_t)c->n_elem * per * sizeof(double)));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_ip[kind], data, (size_t)c->n_elem * per * sizeof(double),
                                location == NSB_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return NSB_OK;
}

extern "C" int nsb_set_dirichlet(nsb_ctx* c, int64_t n, const int64_t* dofs)
{
    if (!c || n < 0 || (n > 0 && !dofs)) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_set_dirichlet: no grid uploaded");
    CUDA_TRY(c, cudaSetDevice(c->device));
    for (int64_t i = 0; i < n; i++) if (dofs[i] < 0 || dofs[i] >= c->n_dof) return set_err(c, NSB_ERR_INVALID, "nsb_set_dirichlet: dof %lld out of range", (long long)dofs[i]);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_dir); c->d_dir = nullptr; c->n_dir = 0;
    cudaFree(c->d_dirval); c->d_dirval = nullptr;
    if (n == 0) return NSB_OK;
    CUDA_TRY(c, upload(c, &c->d_dir, dofs, (size_t)n));
    c->n_dir = n;
    return NSB_OK;
}

extern "C" int nsb_adjust_jacobian(nsb_ctx* c, double* values)
{
    if (!c) return NSB_ERR_INVALID;
    if (c->n_dir == 0) return NSB_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    double* val = values ? values : c->d_jres;
    if (!val) return set_err(c, NSB_ERR_INVALID, "nsb_adjust_jacobian: no resident Jacobian and values == NULL");
    int rc = ensure_pattern(c);
    if (rc) return rc;
    const unsigned nblk = (unsigned)std::min<int64_t>((c->n_dir + 3) / 4, (int64_t)c->sm_count * 16);
    if (c->disc == NSB_DISC_FVCR) dirichlet_rows_csr_kernel<<<nblk, 128, 0, c->stream>>>(c->n_dir, c->d_dir, c->d_rowptr, c->d_bcol, val);
    else if (kDIM[c->elem] == 3) dirichlet_rows_bcsr_kernel<4><<<nblk, 128, 0, c->stream>>>(c->n_dir, c->d_dir, c->d_brow, c->d_bcol, val);
    else dirichlet_rows_bcsr_kernel<3><<<nblk, 128, 0, c->stream>>>(c->n_dir, c->d_dir, c->d_brow, c->d_bcol, val);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return NSB_OK;
}

// vec[dof_i] := g[i] (g == NULL: 0). adjust_defect: g = NULL; adjust_solution: g = Dirichlet values (host pointer, n_dir entries)
extern "C" int nsb_adjust_vector(nsb_ctx* c, double* vec, const double* g, int location)
{
    if (!c || !vec) return NSB_ERR_INVALID;
    if (c->n_dir == 0) return NSB_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const double* dg = nullptr;
    if (g) {
        int rc = ensure(c, &c->d_dirval, (size_t)c->n_dir);
        if (rc) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_dirval, g, sizeof(double) * c->n_dir, cudaMemcpyHostToDevice, c->stream));
        dg = c->d_dirval;
    }
    double* dv = vec;
    const size_t nb = sizeof(double) * c->n_dof;
    if (location == NSB_HOST) {
        int rc = ensure(c, &c->d_yout, c->n_dof);
        if (rc) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_yout, vec, nb, cudaMemcpyHostToDevice, c->stream));
        dv = c->d_yout;
    }
    dirichlet_set_kernel<<<(unsigned)std::min<int64_t>((c->n_dir + 255) / 256, (int64_t)c->sm_count * 8), 256, 0, c->stream>>>(c->n_dir, c->d_dir, dg, dv);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    if (location == NSB_HOST) {
        CUDA_TRY(c, cudaMemcpyAsync(vec, dv, nb, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// boundary element discs on the FV1 boundary faces (ns_bnd.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_set_boundary_faces(nsb_ctx* c, int kind, int64_t n_side, const int32_t* elem, const int32_t* side, const double* data)
{
    if (!c || kind < 0 || kind > 2) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_set_boundary_faces: no grid uploaded");
    if (c->disc != NSB_DISC_FV1) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_set_boundary_faces: FV1 only");
    if (c->elem == NSB_PRISM && kind == NSB_BND_TURB_ZERO) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_set_boundary_faces: the turbulence-zero boundary belongs to the Smagorinsky provider, which is not provided for prisms");
    if (n_side < 0 || (n_side > 0 && (!elem || !side))) return NSB_ERR_INVALID;
    if (kind == NSB_BND_INFLOW && n_side > 0 && !data) return set_err(c, NSB_ERR_INVALID, "nsb_set_boundary_faces: the inflow condition needs its vector data at the boundary-face ips");
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    nsb_ctx::BndSet& b = c->bnd[kind];
    cudaFree(b.d_bnode); cudaFree(b.d_bptr); cudaFree(b.d_bf); cudaFree(b.d_data); b = nsb_ctx::BndSet();
    if (kind == NSB_BND_TURB_ZERO) { cudaFree(c->d_bidx); cudaFree(c->d_dbf); c->d_bidx = nullptr; c->d_dbf = nullptr; }
    if (n_side == 0) return NSB_OK;
    const int nsh = kNSH[c->elem], dim = kDIM[c->elem], nside = kNSIDE[c->elem];
    static const int8_t sides[5][6][4] = {{{0, 1, -1, -1}, {1, 2, -1, -1}, {2, 0, -1, -1}}, {{0, 1, -1, -1}, {1, 2, -1, -1}, {2, 3, -1, -1}, {3, 0, -1, -1}},
                                          {{0, 2, 1, -1}, {1, 2, 3, -1}, {0, 3, 2, -1}, {0, 1, 3, -1}},
                                          {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}},
                                          {{0, 2, 1, -1}, {0, 1, 4, 3}, {1, 2, 5, 4}, {2, 0, 3, 5}, {3, 4, 5, -1}}};   // prism: triangles and quadrilaterals
    std::vector<int32_t> conn((size_t)c->n_elem * nsh);
    CUDA_TRY(c, cudaMemcpy(conn.data(), c->d_conn, conn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    struct Key { int32_t node; int64_t b; int j; };
    std::vector<Key> keys; keys.reserve((size_t)n_side * 4);
    for (int64_t q = 0; q < n_side; q++) {
        if (elem[q] < 0 || elem[q] >= c->n_elem || side[q] < 0 || side[q] >= nside) return set_err(c, NSB_ERR_INVALID, "nsb_set_boundary_faces: bad (element, side) pair %lld", (long long)q);
        for (int j = 0; j < 4 && sides[c->elem][side[q]][j] >= 0; j++) keys.push_back({conn[(size_t)elem[q] * nsh + sides[c->elem][side[q]][j]], q, j});
    }
    std::sort(keys.begin(), keys.end(), [](const Key& x, const Key& y) { return x.node != y.node ? x.node < y.node : (x.b != y.b ? x.b < y.b : x.j < y.j); });
    std::vector<int32_t> bnode; std::vector<int64_t> bptr; std::vector<nsb::BndFace> bf(keys.size());
    for (size_t i = 0; i < keys.size(); i++) {
        if (i == 0 || keys[i].node != keys[i - 1].node) { bnode.push_back(keys[i].node); bptr.push_back((int64_t)i); }
        bf[i].elem = elem[keys[i].b]; bf[i].side = (int16_t)side[keys[i].b]; bf[i].j = (int16_t)keys[i].j;
        bf[i].data = kind == NSB_BND_INFLOW ? (int32_t)(keys[i].b * 4 + keys[i].j) : -1;
    }
    bptr.push_back((int64_t)keys.size());
    b.n_bnode = (int64_t)bnode.size(); b.n_bf = (int64_t)bf.size();
    CUDA_TRY(c, upload(c, &b.d_bnode, bnode.data(), bnode.size()));
    CUDA_TRY(c, upload(c, &b.d_bptr, bptr.data(), bptr.size()));
    CUDA_TRY(c, upload(c, &b.d_bf, bf.data(), bf.size()));
    if (kind == NSB_BND_INFLOW) CUDA_TRY(c, upload(c, &b.d_data, data, (size_t)n_side * 4 * dim));
    if (kind == NSB_BND_TURB_ZERO) {
        std::vector<int32_t> bidx((size_t)c->n_node, -1);
        for (size_t i = 0; i < bnode.size(); i++) bidx[bnode[i]] = (int32_t)i;
        cudaFree(c->d_bidx); cudaFree(c->d_dbf); c->d_bidx = nullptr; c->d_dbf = nullptr;
        CUDA_TRY(c, upload(c, &c->d_bidx, bidx.data(), bidx.size()));
        CUDA_TRY(c, dev_malloc(c, &c->d_dbf, sizeof(double) * bnode.size() * dim * dim));
    }
    return NSB_OK;
}

extern "C" int nsb_assemble_boundary(nsb_ctx* c, int what, const double* u, double sa, double* values, double* defect, int location)
{
    if (!c) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_assemble_boundary: no grid uploaded");
    if (c->disc != NSB_DISC_FV1) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_assemble_boundary: FV1 only");
    const bool jac = what & NSB_JAC_A, dfc = what & (NSB_DEF_A | NSB_RHS);
    if (dfc && !defect) return set_err(c, NSB_ERR_INVALID, "nsb_assemble_boundary: defect pointer missing");
    if (!u) return set_err(c, NSB_ERR_INVALID, "nsb_assemble_boundary: u == NULL");
    CUDA_TRY(c, cudaSetDevice(c->device));
    double* dv = jac ? (values ? values : c->d_jres) : nullptr;          // values: device pointer, NULL = the resident Jacobian
    if (jac && !dv) return set_err(c, NSB_ERR_INVALID, "nsb_assemble_boundary: no resident Jacobian (nsb_assemble_resident) and values == NULL");
    KParams k;
    int rc = resolve_params(c, k, what & (NSB_JAC_A | NSB_DEF_A | NSB_RHS), nullptr, sa, 1.0);
    if (rc) return rc;
    const double* du = u; double* dd = defect;
    const size_t nb = sizeof(double) * c->n_dof;
    if (location == NSB_HOST) {
        if ((rc = order_after_async_copies(c))) return rc;
        if ((rc = ensure(c, &c->d_u, c->n_dof))) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_u, u, nb, cudaMemcpyHostToDevice, c->stream)); du = c->d_u;
        if (dfc) { if ((rc = ensure(c, &c->d_def, c->n_dof))) return rc; dd = c->d_def;
                   CUDA_TRY(c, cudaMemcpyAsync(dd, defect, nb, cudaMemcpyHostToDevice, c->stream)); }
    }
    const MeshDev m = mesh_view(c);
    for (int kind = 0; kind < 2; kind++) {
        const nsb_ctx::BndSet& b = c->bnd[kind];
        if (b.n_bnode == 0) continue;
        if (kind == NSB_BND_INFLOW && !dfc) continue;
        const unsigned nblk = (unsigned)((b.n_bnode + 63) / 64);
#define NSB_BND(EE) fv1_boundary_kernel<EE><<<nblk, 64, 0, c->stream>>>(k, m, kind, b.n_bnode, b.d_bnode, b.d_bptr, b.d_bf, b.d_data, du, dv, dfc ? dd : nullptr, c->d_err)
        switch (c->elem) { case NSB_TRI: NSB_BND(E_TRI); break; case NSB_QUAD: NSB_BND(E_QUAD); break; case NSB_TET: NSB_BND(E_TET); break; case NSB_PRISM: NSB_BND(E_PRISM); break; default: NSB_BND(E_HEX); }
#undef NSB_BND
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
    }
    if (location == NSB_HOST) {
        if (dfc) CUDA_TRY(c, cudaMemcpyAsync(defect, dd, nb, cudaMemcpyDeviceToHost, c->stream));
        return check_device_error(c);
    }
    return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Smagorinsky viscosity as device-side provider of the per-ip viscosity import; diagnostics (ns_turb.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_turbulent_viscosity(nsb_ctx* c, int model, double cmodel, const double* u, int64_t n_zero, const int64_t* zero_nodes,
                                       double* nu_t, int location)
{
    if (!c) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_turbulent_viscosity: no grid uploaded");
    if (c->disc != NSB_DISC_FV1 || c->elem == NSB_PRISM) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_turbulent_viscosity: FV1 on tri / quad / tet / hex only");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (model == NSB_TURB_OFF) return nsb_set_ip_data(c, NSB_IP_KIN_VISC_SCVF, nullptr, NSB_HOST);
    if (model != NSB_TURB_SMAGORINSKY) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_turbulent_viscosity: only the Smagorinsky model is available on the device");
    if (!u) return set_err(c, NSB_ERR_INVALID, "nsb_turbulent_viscosity: u == NULL");
    if (!c->prm.kin_visc_set) return set_err(c, NSB_ERR_SETUP, "NavierStokes::prep_elem_loop: Kinematic Viscosity has not been set, but is required.");
    int rc;
    const double* du = u;
    if (location == NSB_HOST) {
        if ((rc = ensure(c, &c->d_u, c->n_dof))) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_u, u, sizeof(double) * c->n_dof, cudaMemcpyHostToDevice, c->stream)); du = c->d_u;
    }
    if ((rc = ensure(c, &c->d_nut, (size_t)c->n_node))) return rc;
    cudaFree(c->d_zflag); c->d_zflag = nullptr;
    if (n_zero > 0) {
        if (!zero_nodes) return NSB_ERR_INVALID;
        std::vector<uint8_t> z((size_t)c->n_node, 0);
        for (int64_t i = 0; i < n_zero; i++) {
            if (zero_nodes[i] < 0 || zero_nodes[i] >= c->n_node) return set_err(c, NSB_ERR_INVALID, "nsb_turbulent_viscosity: bad node index");
            z[zero_nodes[i]] = 1;
        }
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, upload(c, &c->d_zflag, z.data(), z.size()));
    }
    const int nip = c->elem == NSB_HEX ? 12 : (c->elem == NSB_TET ? 6 : kNSH[c->elem]);
    double*& ipv = c->d_ip[NSB_IP_KIN_VISC_SCVF];
    if (!ipv) CUDA_TRY(c, dev_malloc(c, &ipv, sizeof(double) * (size_t)c->n_elem * nip));
    const MeshDev m = mesh_view(c);
    const nsb_ctx::BndSet& b = c->bnd[NSB_BND_TURB_ZERO];
    const unsigned nb_node = (unsigned)((c->n_node + 127) / 128), nb_ip = (unsigned)((c->n_elem * nip + 255) / 256);
#define NSB_TURB(EE) do { \
        if (b.n_bnode > 0) fv1_smagorinsky_bf_kernel<EE><<<(unsigned)((b.n_bnode + 63) / 64), 64, 0, c->stream>>>(m, du, b.n_bnode, b.d_bnode, b.d_bptr, b.d_bf, c->d_dbf); \
        fv1_smagorinsky_kernel<EE><<<nb_node, 128, 0, c->stream>>>(m, du, cmodel, b.n_bnode > 0 ? c->d_bidx : nullptr, c->d_dbf, c->d_zflag, c->d_nut); \
        fv1_ip_visc_kernel<EE><<<nb_ip, 256, 0, c->stream>>>(m, c->d_nut, c->prm.kin_visc, ipv); } while (0)
    switch (c->elem) { case NSB_TRI: NSB_TURB(E_TRI); break; case NSB_QUAD: NSB_TURB(E_QUAD); break; case NSB_TET: NSB_TURB(E_TET); break; default: NSB_TURB(E_HEX); }
#undef NSB_TURB
    c->launches += b.n_bnode > 0 ? 3 : 2;
    CUDA_TRY(c, cudaGetLastError());
    if (nu_t) {
        if (location == NSB_HOST) { CUDA_TRY(c, cudaMemcpyAsync(nu_t, c->d_nut, sizeof(double) * c->n_node, cudaMemcpyDeviceToHost, c->stream)); CUDA_TRY(c, cudaStreamSynchronize(c->stream)); }
        else CUDA_TRY(c, cudaMemcpyAsync(nu_t, c->d_nut, sizeof(double) * c->n_node, cudaMemcpyDeviceToDevice, c->stream));
    }
    return NSB_OK;
}

extern "C" int nsb_diagnostic(nsb_ctx* c, int kind, const double* u, double dt, double* out, int location)
{
    if (!c || !u || !out) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_diagnostic: no grid uploaded");
    if (c->elem == NSB_PRISM) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_diagnostic: not provided for prisms");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc;
    const double* du = u;
    if (location == NSB_HOST) {
        if ((rc = ensure(c, &c->d_u, c->n_dof))) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_u, u, sizeof(double) * c->n_dof, cudaMemcpyHostToDevice, c->stream)); du = c->d_u;
    }
    if (kind == NSB_DIAG_VORTICITY) {
        if (c->disc != NSB_DISC_FV1) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_diagnostic: the vorticity of a Crouzeix-Raviart field (vorticityFVCR) is not available on the device");
        double* dv = out;
        if (location == NSB_HOST) { if ((rc = ensure(c, &c->d_nut, (size_t)c->n_node))) return rc; dv = c->d_nut; }
        const MeshDev m = mesh_view(c);
        const unsigned nb = (unsigned)((c->n_node + 127) / 128);
        switch (c->elem) {
            case NSB_TRI: fv1_vorticity_kernel<E_TRI><<<nb, 128, 0, c->stream>>>(m, du, dv, c->d_err); break;
            case NSB_QUAD: fv1_vorticity_kernel<E_QUAD><<<nb, 128, 0, c->stream>>>(m, du, dv, c->d_err); break;
            case NSB_TET: fv1_vorticity_kernel<E_TET><<<nb, 128, 0, c->stream>>>(m, du, dv, c->d_err); break;
            default: fv1_vorticity_kernel<E_HEX><<<nb, 128, 0, c->stream>>>(m, du, dv, c->d_err);
        }
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
        if (location == NSB_HOST) { CUDA_TRY(c, cudaMemcpyAsync(out, dv, sizeof(double) * c->n_node, cudaMemcpyDeviceToHost, c->stream)); return check_device_error(c); }
        return NSB_OK;
    }
    if (kind != NSB_DIAG_KINETIC_ENERGY && kind != NSB_DIAG_CFL) return NSB_ERR_INVALID;
    if (c->disc != NSB_DISC_FVCR) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_diagnostic: kineticEnergy / cflNumber work on a Crouzeix-Raviart velocity (navier_stokes_tools.h:731-965)");
    if (c->elem != NSB_TRI && c->elem != NSB_TET) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_diagnostic: kineticEnergy / cflNumber are provided on simplices only");
    const int64_t nblk = (c->n_elem + 255) / 256;
    if (!c->d_diag) CUDA_TRY(c, dev_malloc(c, &c->d_diag, sizeof(double) * (size_t)(nblk * 3 + 2)));
    double* res = c->d_diag + nblk * 3;
    if (c->elem == NSB_TRI) fvcr_diag_kernel<E_TRI><<<(unsigned)nblk, 256, 0, c->stream>>>(c->fvcr, du, dt, c->d_diag);
    else fvcr_diag_kernel<E_TET><<<(unsigned)nblk, 256, 0, c->stream>>>(c->fvcr, du, dt, c->d_diag);
    diag_final_kernel<<<1, 256, 0, c->stream>>>(nblk, c->d_diag, res);
    c->launches += 2;
    CUDA_TRY(c, cudaGetLastError());
    const double* src = res + (kind == NSB_DIAG_CFL ? 1 : 0);
    CUDA_TRY(c, cudaMemcpyAsync(out, src, sizeof(double), location == NSB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    if (location == NSB_HOST) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// DiscConstraintFVCR, default configuration (ns_crc.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_fvcr_constraint_defect(nsb_ctx* c, const double* u, double s_a, int lin_upwind, int lin_pressure, int64_t n_zero,
                                          const int64_t* zero_grad_sides, double* defect, int location)
{
    if (!c || !u || !defect) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_fvcr_constraint_defect: no grid uploaded");
    if (c->disc != NSB_DISC_FVCR) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_fvcr_constraint_defect: FVCR only");
    if (c->elem != NSB_TRI && c->elem != NSB_TET) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_fvcr_constraint_defect: simplices only");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc;
    const int dim = kDIM[c->elem];
    const double* du = u; double* dd = defect;
    const size_t nb = sizeof(double) * c->n_dof;
    if (location == NSB_HOST) {
        if ((rc = order_after_async_copies(c))) return rc;
        if ((rc = ensure(c, &c->d_u, c->n_dof)) || (rc = ensure(c, &c->d_def, c->n_dof))) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_u, u, nb, cudaMemcpyHostToDevice, c->stream)); du = c->d_u;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_def, defect, nb, cudaMemcpyHostToDevice, c->stream)); dd = c->d_def;
    }
    if ((rc = ensure(c, &c->d_sgrad, (size_t)c->n_side * dim * dim))) return rc;
    cudaFree(c->d_zgrad); c->d_zgrad = nullptr;
    if (n_zero > 0) {
        if (!zero_grad_sides) return NSB_ERR_INVALID;
        std::vector<uint8_t> z((size_t)c->n_side, 0);
        for (int64_t i = 0; i < n_zero; i++) {
            if (zero_grad_sides[i] < 0 || zero_grad_sides[i] >= c->n_side) return set_err(c, NSB_ERR_INVALID, "nsb_fvcr_constraint_defect: bad side index");
            z[zero_grad_sides[i]] = 1;
        }
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, upload(c, &c->d_zgrad, z.data(), z.size()));
    }
    const unsigned nblk = (unsigned)((c->n_side + 127) / 128);
    if (c->elem == NSB_TRI) {
        if (lin_upwind) fvcr_side_grad_kernel<E_TRI><<<nblk, 128, 0, c->stream>>>(c->fvcr, c->d_sadj, du, c->d_sgrad, c->d_err);
        fvcr_constraint_defect_kernel<E_TRI><<<nblk, 128, 0, c->stream>>>(c->fvcr, c->d_sadj, du, c->d_sgrad, c->d_zgrad, s_a, lin_upwind, lin_pressure, dd);
    } else {
        if (lin_upwind) fvcr_side_grad_kernel<E_TET><<<nblk, 128, 0, c->stream>>>(c->fvcr, c->d_sadj, du, c->d_sgrad, c->d_err);
        fvcr_constraint_defect_kernel<E_TET><<<nblk, 128, 0, c->stream>>>(c->fvcr, c->d_sadj, du, c->d_sgrad, c->d_zgrad, s_a, lin_upwind, lin_pressure, dd);
    }
    c->launches += lin_upwind ? 2 : 1;
    CUDA_TRY(c, cudaGetLastError());
    if (location == NSB_HOST) {
        CUDA_TRY(c, cudaMemcpyAsync(defect, dd, nb, cudaMemcpyDeviceToHost, c->stream));
        return check_device_error(c);
    }
    return NSB_OK;
}

// Nodes whose rows are assembled FIRST (e.g. the interface nodes of a partition): the owner-computes rows kernels walk the node
// order [priority nodes | the rest]; nsb_assemble(what | NSB_PHASE_PRIORITY) stops behind the priority rows,
// nsb_assemble(what | NSB_PHASE_REST) finishes the pass.
extern "C" int nsb_set_priority_nodes(nsb_ctx* c, int64_t n, const int64_t* nodes)
{
    if (!c || n < 0 || (n > 0 && !nodes)) return NSB_ERR_INVALID;
    if (!c->mesh_ready) return set_err(c, NSB_ERR_INVALID, "nsb_set_priority_nodes: no grid uploaded");
    if (c->disc != NSB_DISC_FV1) return set_err(c, NSB_ERR_UNSUPPORTED, "nsb_set_priority_nodes: FV1 only");
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    std::vector<uint8_t> flag((size_t)c->n_node, 0);
    std::vector<int32_t> order; order.reserve((size_t)c->n_node);
    for (int64_t i = 0; i < n; i++) {
        if (nodes[i] < 0 || nodes[i] >= c->n_node) return set_err(c, NSB_ERR_INVALID, "nsb_set_priority_nodes: node %lld out of range", (long long)nodes[i]);
        if (!flag[nodes[i]]) { flag[nodes[i]] = 1; order.push_back((int32_t)nodes[i]); }
    }
    c->n_prio = (int64_t)order.size();
    for (int64_t a = 0; a < c->n_node; a++) if (!flag[a]) order.push_back((int32_t)a);
    if (!c->d_node_order) CUDA_TRY(c, dev_malloc(c, &c->d_node_order, sizeof(int32_t) * (size_t)c->n_node));
    CUDA_TRY(c, cudaMemcpy(c->d_node_order, order.data(), sizeof(int32_t) * order.size(), cudaMemcpyHostToDevice));
    return NSB_OK;
}

extern "C" int nsb_pack(nsb_ctx* c, int64_t n, const int64_t* idx, const double* src, double* out)
{
    if (!c) return NSB_ERR_INVALID;
    if (n <= 0) return NSB_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    pack_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, c->sm_count * 8), 256, 0, c->stream>>>(n, idx, src, out);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return NSB_OK;
}

extern "C" int nsb_unpack_add(nsb_ctx* c, int64_t n, const int64_t* idx, const double* in, double* dst)
{
    if (!c) return NSB_ERR_INVALID;
    if (n <= 0) return NSB_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    unpack_add_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, c->sm_count * 8), 256, 0, c->stream>>>(n, idx, in, dst);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// FVCR upload (side-based velocity dofs + element pressure); kernels in ns_fvcr.cuh
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_upload_mesh_fvcr(nsb_ctx* c, int elem, int64_t n_elem, int64_t n_node, int64_t n_side, const int32_t* conn,
                                    const int32_t* esides, const double* coords)
{
    if (!c) return NSB_ERR_INVALID;
    if (elem != NSB_TRI && elem != NSB_TET && elem != NSB_QUAD && elem != NSB_HEX) return set_err(c, NSB_ERR_UNSUPPORTED, "FVCR: tri / quad / tet / hex only; hanging-node, prism and pyramid CR geometries are out of scope");
    if (n_elem <= 0 || n_node <= 0 || n_side <= 0 || !conn || !esides || !coords) return set_err(c, NSB_ERR_INVALID, "nsb_upload_mesh_fvcr: bad arguments");
    CUDA_TRY(c, cudaSetDevice(c->device));
    free_mesh(c);
    const int nco = kNSH[elem], dim = kDIM[elem], ns = kNSIDE[elem];
    c->elem = elem; c->disc = NSB_DISC_FVCR; c->n_elem = n_elem; c->n_node = n_node; c->n_side = n_side;
    for (int64_t i = 0; i < n_elem * nco; i++) if (conn[i] < 0 || conn[i] >= n_node) return set_err(c, NSB_ERR_INVALID, "connectivity entry out of range");
    EntityGraph g;
    { const std::string ge = build_entity_graph(n_elem, n_side, ns, esides, g);
      if (!ge.empty()) return set_err(c, ge.find("too large") != std::string::npos ? NSB_ERR_UNSUPPORTED : NSB_ERR_INVALID, "%s", ge.c_str()); }
    // scalar CSR: velocity rows (side s, d): neighbour sides x dim, then pressures of adjacent elements;
    // pressure row e: its sides x dim (sorted), then itself.
    const int64_t pbase = n_side * dim;
    c->n_dof = pbase + n_elem;
    if ((double)c->n_dof >= 2147483647.0) return set_err(c, NSB_ERR_UNSUPPORTED, "grid has %lld dofs: column indices are 32-bit", (long long)c->n_dof);
    std::vector<int64_t>& rp = c->h_brow; std::vector<int32_t>& ci = c->h_bcol;
    rp.assign(c->n_dof + 1, 0);
    for (int64_t s = 0; s < n_side; s++) {
        const int64_t len = (g.brow[s + 1] - g.brow[s]) * dim + (g.adj_ptr[s + 1] - g.adj_ptr[s]);
        for (int d = 0; d < dim; d++) rp[s * dim + d + 1] = len;
    }
    for (int64_t e = 0; e < n_elem; e++) rp[pbase + e + 1] = ns * dim + 1;
    for (int64_t i = 0; i < c->n_dof; i++) rp[i + 1] += rp[i];
    c->nnz = rp[c->n_dof];
    if ((double)c->nnz >= 2147483647.0 * 4) return set_err(c, NSB_ERR_UNSUPPORTED, "FVCR grid too large");
    ci.resize(c->nnz);
    parallel_for(n_side, [&](int64_t lo, int64_t hi) {
        for (int64_t s = lo; s < hi; s++) for (int d = 0; d < dim; d++) {
            int64_t q = rp[s * dim + d];
            for (int64_t kk = g.brow[s]; kk < g.brow[s + 1]; kk++) for (int d2 = 0; d2 < dim; d2++) ci[q++] = g.bcol[kk] * dim + d2;
            for (int64_t kk = g.adj_ptr[s]; kk < g.adj_ptr[s + 1]; kk++) ci[q++] = (int32_t)(pbase + g.adj[kk] / ns);
        }
    });
    std::vector<int32_t> psort((size_t)n_elem * ns);        // rank of local side k among the element's sorted sides
    parallel_for(n_elem, [&](int64_t lo, int64_t hi) {
        for (int64_t e = lo; e < hi; e++) {
            int32_t tmp[6]; for (int kk = 0; kk < ns; kk++) tmp[kk] = esides[e * ns + kk];
            std::sort(tmp, tmp + ns);
            int64_t q = rp[pbase + e];
            for (int kk = 0; kk < ns; kk++) for (int d2 = 0; d2 < dim; d2++) ci[q++] = tmp[kk] * dim + d2;
            ci[q++] = (int32_t)(pbase + e);
            for (int kk = 0; kk < ns; kk++) psort[e * ns + kk] = (int32_t)(std::lower_bound(tmp, tmp + ns, esides[e * ns + kk]) - tmp);
        }
    });
    // scatter maps: slot of side k in the row of side a; slot of element e among the elements of side a
    std::vector<uint8_t> emap; build_emap(n_elem, ns, esides, g, emap);
    std::vector<uint8_t> pslot((size_t)n_elem * ns);
    parallel_for(n_elem, [&](int64_t lo, int64_t hi) {
        for (int64_t e = lo; e < hi; e++) for (int a = 0; a < ns; a++) {
            const int32_t s = esides[e * ns + a];
            int r = 0;
            for (int64_t kk = g.adj_ptr[s]; kk < g.adj_ptr[s + 1]; kk++) if (g.adj[kk] / ns == e) r = (int)(kk - g.adj_ptr[s]);
            pslot[e * ns + a] = (uint8_t)r;
        }
    });
    std::vector<int32_t> order;
    c->n_colors = color_elements(n_elem, n_side, ns, esides, order, c->h_color_ptr);
    c->max_cnt = g.max_cnt;
    // rowptr of velocity rows only needs (first value index, row length) per side
    std::vector<int64_t> srow(n_side + 1);
    for (int64_t s = 0; s <= n_side; s++) srow[s] = s < n_side ? rp[s * dim] : rp[pbase];
    std::vector<int32_t> scnt(n_side);
    for (int64_t s = 0; s < n_side; s++) scnt[s] = (int32_t)(g.brow[s + 1] - g.brow[s]);
    CUDA_TRY(c, upload(c, &c->d_conn, conn, (size_t)n_elem * nco));
    CUDA_TRY(c, upload(c, &c->d_coords, coords, (size_t)n_node * dim));
    CUDA_TRY(c, upload(c, &c->d_esides, esides, (size_t)n_elem * ns));
    CUDA_TRY(c, upload(c, &c->d_color_order, order.data(), order.size()));
    FvcrDev& f = c->fvcr;
    f.n_elem = n_elem; f.n_node = n_node; f.n_side = n_side; f.nnz = c->nnz; f.n_dof = c->n_dof;
    f.conn = c->d_conn; f.coords = c->d_coords; f.esides = c->d_esides; f.color_order = c->d_color_order;
    f.n_colors = c->n_colors; f.color_ptr = nullptr;
    CUDA_TRY(c, upload(c, &f.srow, srow.data(), srow.size()));
    CUDA_TRY(c, upload(c, &f.scnt, scnt.data(), scnt.size()));
    CUDA_TRY(c, upload(c, &f.emap, emap.data(), emap.size()));
    CUDA_TRY(c, upload(c, &f.pslot, pslot.data(), pslot.size()));
    CUDA_TRY(c, upload(c, &f.psort, psort.data(), psort.size()));
    CUDA_TRY(c, upload(c, &f.sadj_ptr, g.adj_ptr.data(), g.adj_ptr.size()));
    CUDA_TRY(c, upload(c, &c->d_sadj, g.adj.data(), g.adj.size()));
    f.prow0 = rp[pbase];
    c->mesh_ready = true;
    return NSB_OK;
}
