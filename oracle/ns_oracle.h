/* ns_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE)
 *
 * Plain-C restatement of the per-element FV1 / FVCR defect + Jacobian assembly of
 * UG4's NavierStokes plugin (reference: /root/reference).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
 * path, and it cannot be compiled here (every hot-path source includes ugcore headers,
 * e.g. upwind.cpp:39, that are not vendored).  The oracle is therefore pinned only by
 * (a) line-by-line restatement of the reference arithmetic cited at each function and
 * (b) the self-consistency invariants in tests/test_oracle_*.py.  ugcore conventions
 * (reference-element numbering, SCV/SCVF construction, ray/side intersection, dense
 * inverse) are restated from SURVEY.md App. B and are OUR SPEC.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libnsb200.so) never links or calls it.
 */
#ifndef NS_ORACLE_H
#define NS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORA_TRI = 0, ORA_QUAD = 1, ORA_TET = 2, ORA_HEX = 3, ORA_PRISM = 4 };
enum { ORA_UPWIND_NONE = 0, ORA_UPWIND_NO = 1, ORA_UPWIND_FULL = 2, ORA_UPWIND_SKEWED = 3,
       ORA_UPWIND_LPS = 4, ORA_UPWIND_POSITIVE = 5 };
enum { ORA_STAB_FIELDS = 0, ORA_STAB_FLOW = 1, ORA_STAB_NONE = 2 };
enum { ORA_DIFF_RAW = 0, ORA_DIFF_FIVEPOINT = 1, ORA_DIFF_COR = 2 };
enum { ORA_DISC_FV1 = 0, ORA_DISC_FVCR = 1 };
/* what-mask of ora_assemble */
enum { ORA_JAC_A = 1, ORA_DEF_A = 2, ORA_JAC_M = 4, ORA_DEF_M = 8, ORA_RHS = 16 };

/* Mirrors the state held by NavierStokesFV1 / NavierStokesFVCR and their bases
 * (navier_stokes_base.cpp:53-67, incompressible_navier_stokes_base.cpp:53-66,
 *  fv1/navier_stokes_fv1.cpp:62-87, fvcr/navier_stokes_fvcr.cpp:63-88). */
typedef struct {
    int32_t disc;          /* ORA_DISC_*                                             */
    int32_t elem;          /* ORA_TRI..ORA_PRISM                                      */
    int32_t conv_upwind;   /* upwind of the convective term (m_spConvUpwind)         */
    int32_t stab;          /* FV1 stabilisation (m_spStab)                           */
    int32_t stab_upwind;   /* upwind attached to the stabilisation (stab->upwind())  */
    int32_t diff_len;      /* ORA_DIFF_*                                             */
    int32_t stokes, laplace, peclet_blend;
    int32_t pac;           /* set_pac_upwind(true): conv-stab = stab, conv_upwind nulled */
    int32_t defect_upwind; /* FVCR only (m_bDefectUpwind)                            */
    int32_t time_dependent;/* local time series (sol0, sol1, dt) present             */
    int32_t has_source;
    int32_t pad0;
    double  exact_jac;     /* m_bFullNewtonFactor                                    */
    double  grad_div;      /* FVCR m_gradDivFactor                                   */
    double  kin_visc, density;
    double  source[3];
    double  dt;
    /* per-ip data imports (fv1/navier_stokes_fv1.cpp:184-197; FV1 only; NULL = the constants above). Global arrays over the
     * elements, the element routines read the slice of element `elem_index` (set per element by ora_assemble). */
    const double *ip_visc;       /* [n_elem][nip]       m_imKinViscosity at the SCVF ips   */
    const double *ip_rho_scvf;   /* [n_elem][nip]       m_imDensitySCVF                    */
    const double *ip_rho_scv;    /* [n_elem][nsh]       m_imDensitySCV (mass / rhs parts)  */
    const double *ip_src_scvf;   /* [n_elem][nip][dim]  m_imSourceSCVF (closure)           */
    const double *ip_src_scv;    /* [n_elem][nsh][dim]  m_imSourceSCV  (add_rhs_elem)      */
    int64_t elem_index;
} ora_params;

/* FV1 geometry of one element, for the geometry tests. Arrays sized for hex. */
typedef struct {
    int32_t dim, nsh, nip, pad;
    int32_t from[12], to[12];
    double  normal[12][3], xip[12][3], lip[12][3];
    double  shape[12][8], ggrad[12][8][3];
    double  c0c2sq[12];
    double  vol[8];
} ora_fv1_geom;

/* CR geometry of one element (simplices + quad/hex); SCV per side, SCVF per (dim-2)-object */
typedef struct {
    int32_t dim, nsh, nip, nco;
    int32_t from[12], to[12];
    double  normal[12][3], xip[12][3], lip[12][3];
    double  shape[12][6], ggrad[12][6][3];
    double  scv_normal[6][3], scv_xip[6][3], vol[6];
} ora_cr_geom;

int  ora_elem_nsh(int elem);   /* corners */
int  ora_elem_nip(int elem);   /* FV1 SCVFs (= edges) */
int  ora_elem_dim(int elem);
int  ora_elem_nside(int elem);

int  ora_fv1_geometry(int elem, const double *coords /*[nsh][dim]*/, ora_fv1_geom *out);
int  ora_cr_geometry(int elem, const double *coords, ora_cr_geom *out);

/* upwind shapes of one element: ip velocities given. returns 0 ok, <0 error */
int  ora_fv1_upwind(int elem, int upwind, const double *coords, const double *ipvel /*[nip][dim]*/,
                    double *up_sh /*[nip][nsh]*/, double *up_ip /*[nip][nip]*/, double *conv_len /*[nip]*/);

/* ray / element-side intersection (App. B-4). returns 1 if found */
int  ora_side_ray_intersection(int elem, const double *coords, const double *from, const double *dir,
                               int positive, int *side, double *gcut, double *lcut);

/* Local element contributions. u/uold are LocalVector-ordered [fct][sh] (sol0/sol1 = time
 * series, may be NULL when !time_dependent). Jloc is [L][L] with local index fct*nsh+sh,
 * dloc [L]. All outputs are ADDED to (like the reference's += semantics). */
int  ora_fv1_elem(const ora_params *p, const double *coords, const double *u,
                  const double *sol0, const double *sol1, int what, double *Jloc, double *dloc);
int  ora_fvcr_elem(const ora_params *p, const double *coords, const double *u,
                   int what, double *Jloc, double *dloc);

/* Stabilisation output for tests */
int  ora_fv1_stab(const ora_params *p, const double *coords, const double *u, const double *sol0,
                  const double *sol1, double *stab_vel /*[nip][dim]*/,
                  double *shape_vel /*[nip][dim][dim][nsh]*/, double *shape_p /*[nip][dim][nsh]*/);

/* ---- global level: CSR pattern (App. B-7/B-8) and the serial/coloured element loop ---- */
/* FV1: dof = node*(dim+1)+fct. Pass rowptr==NULL to query nnz. */
int64_t ora_fv1_csr(int elem, int64_t n_elem, int64_t n_node, const int32_t *conn,
                    int64_t *rowptr, int32_t *colind);
/* FVCR: dof = side*dim+d, pressure dof = n_side*dim + e. elem_sides [n_elem][nside]. */
int64_t ora_fvcr_csr(int elem, int64_t n_elem, int64_t n_side, const int32_t *elem_sides,
                     int64_t *rowptr, int32_t *colind);

/* Assemble. values/defect are ADDED to (caller zeroes). scale_a/scale_m multiply the
 * stiffness / mass parts (time-stepping scales). nthreads>1 uses a greedy element colouring.
 * FV1: u is [n_node][dim+1]; FVCR: u is the dof vector described above.
 * returns 0, or <0 with message in ora_last_error(). */
int  ora_assemble(const ora_params *p, int64_t n_elem, int64_t n_node_or_side,
                  const int32_t *conn, const double *coords, const int32_t *elem_sides,
                  const double *u, const double *sol0, const double *sol1,
                  const int64_t *rowptr, const int32_t *colind,
                  int what, double scale_a, double scale_m,
                  double *values, double *defect, int nthreads);

/* ---- boundary faces (FV1Geometry BF, our spec) and the boundary discs on them ---- */
int  ora_side_corners(int elem, int side);
int  ora_side_corner(int elem, int side, int j);
int  ora_fv1_bf_geometry(int elem, const double *coords, int side, int j, int *node_id, double *normal /*[dim]*/,
                         double *xip /*[dim]*/, double *shape /*[nsh]*/, double *ggrad /*[nsh][dim]*/);
/* kind 0: NavierStokesNoNormalStressOutflowFV1 (fv1/bnd/no_normal_stress_outflow_fv1.cpp:192-427);
 * kind 1: NeumannBoundaryFV1 part of NavierStokesInflowFV1 (fv1/bnd/inflow_fv1_impl.h:42-82), data [n_side][4][dim] */
int  ora_fv1_boundary(const ora_params *p, int kind, int64_t n_side, const int32_t *belem, const int32_t *bside,
                      const double *data, const int32_t *conn, const double *coords, const double *u,
                      const int64_t *rowptr, const int32_t *colind, int what, double scale_a,
                      double *values, double *defect);

/* ---- Smagorinsky turbulent viscosity as per-ip viscosity (fv1/turbulent_viscosity_fv1.h:200-383,
 * fv1/turbulent_viscosity_fv1_impl.h:504-616,755-762,819-852) and diagnostics (navier_stokes_tools.h:386-525,731-965) ---- */
int  ora_fv1_smagorinsky(int elem, int64_t n_elem, int64_t n_node, const int32_t *conn, const double *coords, const double *u,
                         double c, double kin_visc, int64_t n_bside, const int32_t *belem, const int32_t *bside,
                         const uint8_t *zero_node, double *nu_t, double *ip_visc);
int  ora_fv1_vorticity(int elem, int64_t n_elem, int64_t n_node, const int32_t *conn, const double *coords, const double *u, double *vort);
int  ora_fvcr_diagnostics(int elem, int64_t n_elem, const int32_t *conn, const double *coords, const int32_t *elem_sides,
                          const double *u, double dt, double *out /*[2]: kinetic energy, max CFL*/);

/* ---- DiscConstraintFVCR, default configuration (fvcr/disc_constraint_fvcr.h:254-300,770-1171): linear-upwind and
 * linear-pressure correction of the FVCR defect, ADDED to defect ---- */
int  ora_fvcr_constraint_defect(int elem, int64_t n_elem, int64_t n_side, const int32_t *conn, const double *coords,
                                const int32_t *elem_sides, const double *u, double s_a, int lin_upwind, int lin_pressure,
                                const uint8_t *zero_grad_side, double *defect);

const char *ora_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
