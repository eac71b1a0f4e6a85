/* nsb200.h -- C ABI of libnsb200.so: B200-native (sm_100a) FV1 / FVCR defect + Jacobian assembly of the
 * incompressible Navier-Stokes system, drop-in for the element-assembly path of UG4's NavierStokes plugin.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the UG4 plugin tree).
 * The library collapses ugcore's element loop (gather LocalVector -> prep_elem -> add_*_elem ->
 * AddLocalMatrixToGlobal/AddLocalVector) into device kernels; see DESIGN.md and INTEGRATION.md.
 *
 * All functions return 0 on success, a negative nsb_status otherwise; nsb_last_error(ctx) gives the text
 * (the reference throws UG_THROW at the same conditions).  No CPU fallback exists: every compute entry
 * point fails with NSB_ERR_CUDA when no device is usable.
 */
#ifndef NSB200_H
#define NSB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsb_ctx nsb_ctx;

typedef enum {
    NSB_OK = 0,
    NSB_ERR_INVALID = -1,      /* bad argument / call order                                             */
    NSB_ERR_SETUP = -2,        /* prep_elem_loop validation failed (stabilisation / upwind / ... unset)  */
    NSB_ERR_CUDA = -3,         /* CUDA runtime error or no device                                        */
    NSB_ERR_GEOMETRY = -4,     /* element-level failure on device (ray search found no cut side, ...)    */
    NSB_ERR_UNSUPPORTED = -5
} nsb_status;

enum { NSB_TRI = 0, NSB_QUAD = 1, NSB_TET = 2, NSB_HEX = 3,
       NSB_PRISM = 4 /* FV1 only (fv1/navier_stokes_fv1.cpp:1542); served by the element kernels */ };
enum { NSB_DISC_FV1 = 0, NSB_DISC_FVCR = 1 };
/* upwind_interface.cpp:43-62 (CreateNavierStokesUpwind: "no","full","skewed","lps","pos") */
enum { NSB_UPWIND_UNSET = 0, NSB_UPWIND_NO = 1, NSB_UPWIND_FULL = 2, NSB_UPWIND_SKEWED = 3,
       NSB_UPWIND_LPS = 4, NSB_UPWIND_POSITIVE = 5 };
/* fv1/stabilization.cpp:46-57 (CreateNavierStokesStabilization: "fields","flow") + WithoutStabilization */
enum { NSB_STAB_UNSET = -1, NSB_STAB_FIELDS = 0, NSB_STAB_FLOW = 1, NSB_STAB_NONE = 2 };
/* fv1/stabilization.cpp:86-100 (set_diffusion_length: "raw","fivepoint","cor") */
enum { NSB_DIFF_RAW = 0, NSB_DIFF_FIVEPOINT = 1, NSB_DIFF_COR = 2 };

/* which element contributions to assemble (bitmask): the IElemDisc slots registered at
 * fv1/navier_stokes_fv1.cpp:1553-1572 / fvcr/navier_stokes_fvcr.cpp:825-843 */
enum { NSB_JAC_A = 1, NSB_DEF_A = 2, NSB_JAC_M = 4, NSB_DEF_M = 8, NSB_RHS = 16 };

/* how element contributions reach the global CSR matrix / defect vector */
enum {
    NSB_SCATTER_GATHER = 0,    /* owner-computes: one warp per matrix row block, written once, deterministic */
    NSB_SCATTER_COLORED = 1,   /* element kernel, one launch per colour of a greedy element colouring, deterministic */
    NSB_SCATTER_ATOMIC = 2     /* element kernel, red.global.add.f64 */
};

enum { NSB_HOST = 0, NSB_DEVICE = 1,     /* where u / values / defect pointers live */
       NSB_HOST_ASYNC = 2 };             /* nsb_assemble_resident / nsb_apply_jacobian only: (pinned) host buffers, the copies are queued on the
                                            context's own H2D / D2H streams and the call returns at once, so that x goes up while the assembly
                                            runs and the defect comes back while the product runs. nsb_synchronize completes the calls (the
                                            buffers must stay alive and untouched until then), nsb_check_errors reports element-level failures */

/* State of NavierStokesFV1 / NavierStokesFVCR and their bases that the element routines read.
 * Defaults (nsb_params_default) mirror navier_stokes_base.cpp:53-67,
 * incompressible_navier_stokes_base.cpp:53-66, fv1/navier_stokes_fv1.cpp:62-87,
 * fvcr/navier_stokes_fvcr.cpp:63-88, fv1/stabilization.h:321-325. */
typedef struct {
    int32_t disc;            /* NSB_DISC_*                                                            */
    int32_t conv_upwind;     /* set_upwind(name)            fv1/navier_stokes_fv1.h:213-215            */
    int32_t stab;            /* set_stabilization(name)     fv1/navier_stokes_fv1.h:190-199            */
    int32_t stab_upwind;     /* stab->set_upwind(); 0 = "not set" (auto-wired from conv_upwind like the
                                reference's set_upwind/set_stabilization string overloads)            */
    int32_t diff_length;     /* set_diffusion_length(name)                                            */
    int32_t stokes;          /* set_stokes      incompressible_navier_stokes_plugin.cpp:244-266        */
    int32_t laplace;         /* set_laplace                                                            */
    int32_t peclet_blend;    /* set_peclet_blend                                                       */
    int32_t pac_upwind;      /* set_pac_upwind  fv1/navier_stokes_fv1.h:217-225                        */
    int32_t defect_upwind;   /* FVCR set_defect_upwind      fvcr/navier_stokes_fvcr.cpp:82             */
    int32_t has_source;      /* set_source given                                                       */
    int32_t kin_visc_set, density_set;   /* prep_elem_loop throws when unset (:169-181)               */
    int32_t reserved;
    double  exact_jacobian;  /* set_exact_jacobian(bool|number) -> m_bFullNewtonFactor                 */
    double  grad_div;        /* FVCR set_grad_div                                                      */
    double  kin_visc;        /* set_kinematic_viscosity(number)                                        */
    double  density;         /* set_density(number), default 1                                         */
    double  source[3];       /* set_source(vector)                                                     */
} nsb_params;

/* Local time series handed to the element routines by ugcore's instationary assembling
 * (IElemDisc::local_time_solutions(); read at fv1/navier_stokes_fv1.cpp:268-280, 617-629). */
typedef struct {
    const double *sol0;      /* solution(0): current time point, NULL when stationary                 */
    const double *sol1;      /* solution(1): previous time point                                       */
    double dt;               /* time(0) - time(1)                                                      */
} nsb_time_series;

int  nsb_create(int device, nsb_ctx **out);
void nsb_destroy(nsb_ctx *ctx);
const char *nsb_last_error(const nsb_ctx *ctx);       /* also valid with ctx == NULL (creation errors)  */
/* run on an existing cudaStream_t (e.g. torch's current stream; NULL = the legacy default stream).
 * A fresh context runs on its own non-blocking stream. */
int  nsb_set_stream(nsb_ctx *ctx, void *cuda_stream);

void nsb_params_default(nsb_params *p);
int  nsb_set_params(nsb_ctx *ctx, const nsb_params *p);

/* Upload grid connectivity once (replaces the per-element FillCornerCoordinates + dd->indices() of
 * ugcore's element loop) and build on the host: node->element adjacency, the block-CSR pattern (full
 * element coupling incl. explicit zeros, dof = node*(dim+1)+fct), the element->CSR scatter map and a greedy
 * element colouring; then precompute the SCV-volume table on device.  conn: [n_elem][nsh] int32,
 * coords: [n_node][dim] double (host pointers, borrowed for the call). */
int  nsb_upload_mesh(nsb_ctx *ctx, int elem_type, int64_t n_elem, int64_t n_node,
                     const int32_t *conn, const double *coords);
/* FVCR variant: additionally elem_sides [n_elem][nside] (global side ids; side k of an element is its
 * reference side k). dofs: side*dim+d, then n_side*dim + elem for the pressure. Element types: NSB_TRI, NSB_TET,
 * NSB_QUAD, NSB_HEX (fvcr/navier_stokes_fvcr.cpp:790-813; prism / pyramid CR geometries: NSB_ERR_UNSUPPORTED). */
int  nsb_upload_mesh_fvcr(nsb_ctx *ctx, int elem_type, int64_t n_elem, int64_t n_node, int64_t n_side,
                          const int32_t *conn, const int32_t *elem_sides, const double *coords);

int64_t nsb_num_dofs(const nsb_ctx *ctx);
int64_t nsb_nnz(const nsb_ctx *ctx);
int     nsb_num_colors(const nsb_ctx *ctx);
/* scalar CSR pattern (rows sorted), so the host SparseMatrix can adopt the identical sparsity */
int  nsb_get_csr(const nsb_ctx *ctx, int64_t *rowptr /*[ndof+1]*/, int32_t *colind /*[nnz]*/);

/* prep_elem_loop (fv1/navier_stokes_fv1.cpp:136-199, fvcr/navier_stokes_fvcr.cpp:145-196): validates
 * the set-up, returns NSB_ERR_SETUP with the reference's message when it would throw. */
int  nsb_prep_elem_loop(nsb_ctx *ctx);

/* The element loop.  values[nnz] / defect[ndof] := beta*old + scale_a*(A-part) + scale_m*(M-part),
 * A-part = add_jac_A_elem / add_def_A_elem (- add_rhs_elem when NSB_RHS), M-part = add_jac_M_elem /
 * add_def_M_elem, for all elements, scattered into the global CSR matrix / vector.
 * u: evaluation point ([ndof]); ts: local time series or NULL (stationary).
 * values / defect may be NULL when the corresponding bits are absent from `what`.
 * location: NSB_HOST (buffers are copied in/out inside the call) or NSB_DEVICE (device pointers,
 * asynchronous on the context's stream). */
int  nsb_assemble(nsb_ctx *ctx, int what, int scatter_mode, const double *u, const nsb_time_series *ts,
                  double scale_a, double scale_m, double beta, double *values, double *defect, int location);

/* GPU-resident hand-off of the Jacobian: like nsb_assemble, but the CSR values stay in a context-owned device
 * buffer (beta applies to it) and only the vectors cross the host link. Replaces the copy of the assembled SparseMatrix
 * (ugcore AssembleJacobian -> J, navier_stokes_base.h:200; instationary combination fv1/navier_stokes_fv1.cpp:268-280
 * through scale_a = theta dt, scale_m = 1) to a host solver by a device-resident operator:
 *   nsb_resident_jacobian : device pointer of the values (pattern = nsb_get_csr) for a GPU solver,
 *   nsb_apply_jacobian    : y = alpha J x + beta y (matrix-vector product / residual d - J dx of a Krylov or defect-correction
 *                           loop); values == NULL selects the resident Jacobian, x / y per `location`. */
int  nsb_assemble_resident(nsb_ctx *ctx, int what, int scatter_mode, const double *u, const nsb_time_series *ts,
                           double scale_a, double scale_m, double beta, double *defect, int location);
int  nsb_resident_jacobian(nsb_ctx *ctx, double **dev_values);
int  nsb_apply_jacobian(nsb_ctx *ctx, const double *values, double alpha, const double *x, double beta, double *y, int location);

/* Dirichlet post-pass: what ugcore's DirichletBoundary does for NavierStokesWall (velocity = 0,
 * bnd/wall_impl.h:44-70) and NavierStokesInflowFV1 (velocity = user data, fv1/bnd/inflow_fv1_impl.h:42-82) after the
 * element loop: adjust_jacobian (row := unit row), adjust_defect (entry := 0), adjust_solution (entry := value).
 * dofs: scalar dof indices (host pointer, copied). values == NULL: the resident Jacobian; otherwise a device pointer.
 * nsb_adjust_vector: vec[dofs[i]] := g ? g[i] : 0  (g: host pointer with one value per Dirichlet dof). */
int  nsb_set_dirichlet(nsb_ctx *ctx, int64_t n, const int64_t *dofs);
int  nsb_adjust_jacobian(nsb_ctx *ctx, double *values);
int  nsb_adjust_vector(nsb_ctx *ctx, double *vec, const double *g, int location);

/* Boundary element discs on the FV1 boundary faces. A boundary side is given as (element, local side index in
 * the reference-element numbering); it contributes one boundary face per side corner (FV1Geometry BF).
 *   NSB_BND_OUTFLOW : NavierStokesNoNormalStressOutflowFV1::add_jac_A_elem / add_def_A_elem
 *                     (fv1/bnd/no_normal_stress_outflow_fv1.cpp:343-427): tangential diffusive flux, outflow-only convective
 *                     flux, continuity flux; constant viscosity / density of nsb_params.
 *   NSB_BND_INFLOW  : the NeumannBoundaryFV1 part of NavierStokesInflowFV1 (fv1/bnd/inflow_fv1_impl.h:42-82, :69): the given
 *                     velocity enters the continuity equation, defect(p, node) += scale_a data . n. data: host pointer,
 *                     [n_side][4][dim] = the vector datum at the ip of boundary face j of side q at (q*4 + j)*dim (the velocity
 *                     Dirichlet rows of the same condition go through nsb_set_dirichlet).
 * nsb_set_boundary_faces replaces the sides of that kind (n_side = 0 removes them). nsb_assemble_boundary ADDS the registered
 * contributions (what: NSB_JAC_A | NSB_DEF_A, scaled by scale_a) to values / defect, to be called after nsb_assemble* and
 * before the Dirichlet post-pass. values: device pointer, NULL = the resident Jacobian; u / defect per `location`.
 * Owner-computes (one thread per boundary node, faces in a fixed order): bitwise deterministic. */
enum { NSB_BND_OUTFLOW = 0, NSB_BND_INFLOW = 1, NSB_BND_TURB_ZERO = 2 /* setTurbulenceZeroBoundaries, see nsb_turbulent_viscosity */ };
int  nsb_set_boundary_faces(nsb_ctx *ctx, int kind, int64_t n_side, const int32_t *elem, const int32_t *side, const double *data);
int  nsb_assemble_boundary(nsb_ctx *ctx, int what, const double *u, double scale_a, double *values, double *defect, int location);

/* Turbulent viscosity as a device-side provider of the per-ip kinematic viscosity:
 * FV1SmagorinskyTurbViscData (fv1/turbulent_viscosity_fv1.h:200-383; assembleDeformationTensor
 * fv1/turbulent_viscosity_fv1_impl.h:504-616, FNorm :755-762, update :819-852, evaluate turbulent_viscosity_fv1.h:321-379):
 * nu_t(node) = c vol^(2/dim) sqrt(2 D:D) from the deformation tensor of `u` over the node's control volume, then
 * nu(ip) = sum_sh N_sh(ip) nu_t(sh) + the kinematic viscosity of nsb_params, written into the NSB_IP_KIN_VISC_SCVF import table
 * on the device (no host round trip) -- call it before nsb_assemble*, as the reference calls update() before an assembly.
 * setTurbulenceZeroBoundaries: the boundary sides registered as NSB_BND_TURB_ZERO add their BF closure terms; zero_nodes
 * (host pointer, the vertices of those subsets) keep nu_t = 0. nu_t: optional output [n_node] per `location`.
 * NSB_TURB_OFF removes the table again. The dynamic model (FV1DynamicTurbViscData) is not available on the device. */
enum { NSB_TURB_OFF = -1, NSB_TURB_SMAGORINSKY = 0 };
int  nsb_turbulent_viscosity(nsb_ctx *ctx, int model, double c, const double *u, int64_t n_zero, const int64_t *zero_nodes,
                             double *nu_t, int location);

/* Diagnostics of navier_stokes_tools.h on the device:
 *   NSB_DIAG_VORTICITY      vorticityFV1 (:386-525), FV1 grids: out [n_node] = d_x v - d_y u, SCV-volume weighted per vertex;
 *   NSB_DIAG_KINETIC_ENERGY kineticEnergy (:850-965), FVCR grids: out [1] = sum_e vol_e |u(barycentre)|^2 / sum_e vol_e;
 *   NSB_DIAG_CFL            cflNumber (:731-848), FVCR grids: out [1] = max dt |(x_i - x_j) . u(barycentre)| / |x_i - x_j|^2.
 * Fixed-order reductions (bitwise deterministic). u / out per `location`. */
enum { NSB_DIAG_VORTICITY = 0, NSB_DIAG_KINETIC_ENERGY = 1, NSB_DIAG_CFL = 2 };
int  nsb_diagnostic(nsb_ctx *ctx, int kind, const double *u, double dt, double *out, int location);

/* DiscConstraintFVCR (fvcr/disc_constraint_fvcr.h:164-1198) in its default configuration (:254-300: linear-upwind
 * and linear-pressure correction of the DEFECT, not adaptive, no limiter): adjust_defect (:1149-1171) -> add_defect (:770-1147).
 * ADDS the correction of the state u to `defect` (call after nsb_assemble*, per time point with its stiffness scale s_a as the
 * reference does). lin_upwind / lin_pressure = bLinUpConvDefect / bLinPressureDefect. zero_grad_sides (host pointer): the sides of
 * the zero-gradient subsets (set_zero_grad_bnd :288-292) -- elements touching one are skipped. Owner-computes over the side ->
 * element adjacency: bitwise deterministic. FVCR grids (simplices). Not available: bAdaptive (hanging nodes), the Jacobian
 * variants (bLinUpConvJacobian / bLinPressureJacobian) and the limiter. */
int  nsb_fvcr_constraint_defect(nsb_ctx *ctx, const double *u, double s_a, int lin_upwind, int lin_pressure, int64_t n_zero,
                                const int64_t *zero_grad_sides, double *defect, int location);

/* Phased assembly for the overlap of the interface exchange with the interior assembly ("hide behind interior
 * assembly"; replaces the blocking slave -> master summation of fvcr/pcr_ilut.h:182-194). nsb_set_priority_nodes names the grid
 * nodes (host pointer; n = 0 clears) whose CSR rows / defect entries are produced first -- the interface nodes of a partition.
 *   nsb_assemble(what | NSB_PHASE_PRIORITY, ...)  flux kernel + the rows of the priority nodes
 *   ... the caller packs / sends the interface rows on another stream ...
 *   nsb_assemble(what | NSB_PHASE_REST, ...)      the remaining rows, from the SCVF records of the first call (same u, same
 *                                                 parameters, same pointers; device pointers only)
 * Without a phase flag the pass is unchanged (one launch over the same node order). Paths without a separate rows kernel
 * (element kernels, fused 2-D kernel, FVCR) do the whole pass in the priority phase; the second call is then a no-op. */
enum { NSB_PHASE_PRIORITY = 256, NSB_PHASE_REST = 512 };
int  nsb_set_priority_nodes(nsb_ctx *ctx, int64_t n, const int64_t *nodes);

/* Per-ip data imports: the reference evaluates UserData for viscosity / density / source at the integration points
 * (m_imKinViscosity, m_imDensitySCVF at the SCVF ips; m_imDensitySCV, m_imSourceSCV at the SCV ips; m_imSourceSCVF at the SCVF
 * ips -- fv1/navier_stokes_fv1.cpp:184-197, read at :336,351,390,393,708,805,835,866 and fv1/stabilization.cpp:151,198,229).
 * The caller evaluates its data at those points (any C++ / Lua UserData, e.g. a turbulent viscosity) and hands the arrays
 * over; data == NULL returns to the constant of nsb_params. Layouts: [n_elem][nip], [n_elem][nsh], sources x dim. FV1 only; the
 * element kernels honour them (NSB_SCATTER_GATHER is served by the coloured element kernel while any array is set). */
enum { NSB_IP_KIN_VISC_SCVF = 0, NSB_IP_DENSITY_SCVF = 1, NSB_IP_DENSITY_SCV = 2, NSB_IP_SOURCE_SCVF = 3, NSB_IP_SOURCE_SCV = 4 };
int  nsb_set_ip_data(nsb_ctx *ctx, int kind, const double *data, int location);

/* NSB_DEVICE calls are asynchronous: element-level failures (the reference's UG_THROW inside upwind /
 * stabilisation code, upwind.cpp:354, stabilization.cpp:292,641) are latched on the device; this call
 * synchronises and reports them (NSB_ERR_GEOMETRY). NSB_HOST calls do it implicitly. */
int  nsb_check_errors(nsb_ctx *ctx);

/* compat mode of the IElemDisc slots: local matrices/vectors of all elements, laid out
 * like ugcore's LocalMatrix/LocalVector (index fct*nsh+sh): Jloc [n_elem][L][L], dloc [n_elem][L]. */
int  nsb_local_contributions(nsb_ctx *ctx, int what, const double *u, const nsb_time_series *ts,
                             double *Jloc, double *dloc, int location);

/* interface exchange helpers for the multi-GPU path (replace pcl additive->unique sums on this path):
 * out[i] = src[idx[i]]  /  dst[idx[i]] += in[i]   (device pointers, context stream) */
int  nsb_pack(nsb_ctx *ctx, int64_t n, const int64_t *idx, const double *src, double *out);
int  nsb_unpack_add(nsb_ctx *ctx, int64_t n, const int64_t *idx, const double *in, double *dst);

/* instrumentation */
enum {
    NSB_Q_DEVICE_BYTES = 0,     /* device memory held by the context (grid tables, caches, staging buffers)        */
    NSB_Q_SETUP_SECONDS = 1,    /* host preprocessing + upload time of the last nsb_upload_mesh*                    */
    NSB_Q_FUSED = 2,            /* 1 when the fused patch kernel serves NSB_SCATTER_GATHER on this grid             */
    NSB_Q_PATCHES = 3,          /* number of node patches                                                           */
    NSB_Q_SCVF_EVALS = 4,       /* SCVF evaluations per pass of the fused kernel (>= n_elem * nip: patch overlap)   */
    NSB_Q_PATCH_TABLE_BYTES = 5,/* bytes of the per-patch tables read by every pass                                 */
    NSB_Q_LAST_SCATTER = 6      /* NSB_SCATTER_* mode that actually served the last nsb_assemble*: GATHER requests are served by the
                                   coloured element kernels for FVCR, PositiveUpwind (dense ip systems), PAC and per-ip data      */
};
int  nsb_query(const nsb_ctx *ctx, int what, double *out);
int64_t nsb_launch_count(const nsb_ctx *ctx);          /* kernels launched by this context so far       */
int  nsb_synchronize(nsb_ctx *ctx);
const char *nsb_version(void);

#ifdef __cplusplus
}
#endif
#endif
