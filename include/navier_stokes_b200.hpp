// navier_stokes_b200.hpp -- C++ shim above the C ABI (nsb200.h) that mirrors the element-disc surface of the UG4
// NavierStokes plugin for the assembly path: class names, constructors, setters, defaults and throw conditions of
//   NavierStokesBase                 navier_stokes_base.h:141-208, register_navier_stokes.cpp:105-126
//   IncompressibleNavierStokesBase   incompressible/incompressible_navier_stokes_base.h:144-284
//   NavierStokesFV1                  incompressible/fv1/navier_stokes_fv1.h:185-488, fv1/register_fv1.cpp:166-184
//   NavierStokesFVCR                 incompressible/fvcr/navier_stokes_fvcr.h, fvcr/register_fvcr.cpp:289-303
// and the eight IElemDisc slots registered at fv1/navier_stokes_fv1.cpp:1553-1572.
//
// Two modes:
//   compat : prep_elem_loop() runs the GPU batch for all elements and keeps the per-element LocalMatrix /
//            LocalVector blocks on the host; prep_elem() selects the element; add_*_elem() only `+=` the block.
//            Works inside ugcore's unchanged element loop and constraint handling.
//   fast   : assemble_jacobian / assemble_defect hand u to nsb_assemble, which scatters on the device.
//
// ugcore is not available in this repository's build environment: without -DNSB_WITH_UG4 the header defines
// minimal stand-ins for LocalVector / LocalMatrix / ReferenceObjectID with ugcore's access syntax; with it, the
// ugcore types are used and register_navier_stokes_b200() adds the classes under the
// reference's registry names.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <algorithm>
#include <cctype>

#include "nsb200.h"

namespace nsb200 {

struct UGError : std::runtime_error { using std::runtime_error::runtime_error; };   // stands for UG_THROW
#define NSB_UG_THROW(msg) throw ::nsb200::UGError(msg)

#ifdef NSB_WITH_UG4
// ugcore's own types (the mock of tests/cpp/mock_ug when NSB_UG4_MOCK is defined; the real headers are included by the
// translation unit before this one otherwise)
using ug::number; using ug::LocalVector; using ug::LocalMatrix; using ug::ReferenceObjectID;
#else
typedef double number;
enum ReferenceObjectID { ROID_TRIANGLE = 2, ROID_QUADRILATERAL = 3, ROID_TETRAHEDRON = 4, ROID_HEXAHEDRON = 5, ROID_PRISM = 6 };
// u(fct, dof) / J(rfct, rdof, cfct, cdof): ugcore lib_disc/common/local_algebra.h access syntax
struct LocalVector {
    std::vector<number> v; int nfct = 0, ndof = 0;
    LocalVector() {}
    LocalVector(int nf, int nd) : v((size_t)nf * nd, 0.0), nfct(nf), ndof(nd) {}
    number& operator()(int f, int d) { return v[(size_t)f * ndof + d]; }
    number operator()(int f, int d) const { return v[(size_t)f * ndof + d]; }
};
struct LocalMatrix {
    std::vector<number> v; int nfct = 0, ndof = 0;
    LocalMatrix() {}
    LocalMatrix(int nf, int nd) : v((size_t)nf * nd * nf * nd, 0.0), nfct(nf), ndof(nd) {}
    number& operator()(int rf, int rd, int cf, int cd) { return v[((size_t)rf * ndof + rd) * (nfct * ndof) + (size_t)cf * ndof + cd]; }
    number operator()(int rf, int rd, int cf, int cd) const { return v[((size_t)rf * ndof + rd) * (nfct * ndof) + (size_t)cf * ndof + cd]; }
};
#endif

inline std::string trim_lower(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    std::string n = s.substr(a, b - a);
    std::transform(n.begin(), n.end(), n.begin(), ::tolower);
    return n;
}

// CreateNavierStokesUpwind, upwind_interface.cpp:43-62
inline int upwind_id(const std::string& name)
{
    const std::string n = trim_lower(name);
    if (n == "no") return NSB_UPWIND_NO;
    if (n == "full") return NSB_UPWIND_FULL;
    if (n == "skewed") return NSB_UPWIND_SKEWED;
    if (n == "linearprofileskewed" || n == "lps") return NSB_UPWIND_LPS;
    if (n == "positive" || n == "pos") return NSB_UPWIND_POSITIVE;
    if (n == "regular" || n == "reg") NSB_UG_THROW("NavierStokes: RegularUpwind is not provided on the device path");
    NSB_UG_THROW("NavierStokes: upwind type '" + name + "' not found. Options are: no, full, skewed, linearprofileskewed (lps), positive (pos), regular (reg)");
}
// CreateNavierStokesStabilization, fv1/stabilization.cpp:46-57 ; set_diffusion_length :86-100
inline int stab_id(const std::string& name)
{
    const std::string n = trim_lower(name);
    if (n == "fields") return NSB_STAB_FIELDS;
    if (n == "flow") return NSB_STAB_FLOW;
    NSB_UG_THROW("NavierStokes: stabilization type '" + name + "' not a valid name of a Schneider-Raw stabilization. Options are: fields, flow");
}
inline int diff_length_id(const std::string& name)
{
    const std::string n = trim_lower(name);
    if (n == "raw") return NSB_DIFF_RAW;
    if (n == "fivepoint") return NSB_DIFF_FIVEPOINT;
    if (n == "cor") return NSB_DIFF_COR;
    NSB_UG_THROW("Diffusion Length calculation method not found. Use one of [Raw, Fivepoint, Cor].");
}

class NavierStokesDeviceDisc {
  public:
    NavierStokesDeviceDisc(const char* fcts, const char* subsets, int disc, int dim, int device = 0)
        : m_dim(dim), m_device(device), m_subsets(subsets ? subsets : "")
    {
        nsb_params_default(&m_prm);
        m_prm.disc = disc;
        int n = 0; bool tok = false;
        for (const char* c = fcts; c && *c; ++c) { if (*c == ',') tok = false; else if (!std::isspace((unsigned char)*c) && !tok) { tok = true; n++; } }
        if (n != dim + 1)          // fv1/navier_stokes_fv1.cpp:66, fvcr/navier_stokes_fvcr.cpp:67-68
            NSB_UG_THROW("Wrong number of functions: The ElemDisc 'NavierStokes' needs exactly " + std::to_string(dim + 1) + " symbolic function.");
    }
    virtual ~NavierStokesDeviceDisc() { if (m_ctx) nsb_destroy(m_ctx); }
    NavierStokesDeviceDisc(const NavierStokesDeviceDisc&) = delete;
    NavierStokesDeviceDisc& operator=(const NavierStokesDeviceDisc&) = delete;

    // ---- NavierStokesBase / IncompressibleNavierStokesBase setters ----
    void set_kinematic_viscosity(number v) { m_prm.kin_visc = v; m_prm.kin_visc_set = 1; }
    void set_density(number v) { m_prm.density = v; m_prm.density_set = 1; }
    void set_source(const std::vector<number>& f) { m_prm.has_source = 1; for (size_t d = 0; d < f.size() && d < 3; d++) m_prm.source[d] = f[d]; }
    void set_exact_jacobian(bool b) { m_prm.exact_jacobian = b ? 1.0 : 0.0; }
    void set_exact_jacobian(number f) { m_prm.exact_jacobian = f; }
    void set_peclet_blend(bool b) { m_prm.peclet_blend = b; }
    void set_grad_div(number f) { m_prm.grad_div = f; }
    void set_laplace(bool b) { m_prm.laplace = b; }
    void set_stokes(bool b) { m_prm.stokes = b; }
    bool requests_local_time_series() { return true; }            // navier_stokes_base.h:200
    virtual std::string disc_type() const = 0;

    // ---- grid hand-over (replaces FillCornerCoordinates + dd->indices() of the ugcore loop) ----
    void set_grid(int elem_type, int64_t n_elem, int64_t n_node, const int32_t* conn, const number* coords,
                  int64_t n_side = 0, const int32_t* elem_sides = nullptr)
    {
        ctx();
        m_elem = elem_type; m_n_elem = n_elem;
        static const int nsh[5] = {3, 4, 4, 8, 6}, nside[5] = {3, 4, 4, 6, 5};
        if (m_prm.disc == NSB_DISC_FV1) { check(nsb_upload_mesh(m_ctx, elem_type, n_elem, n_node, conn, coords)); m_nsh = nsh[elem_type]; m_L = m_nsh * (m_dim + 1); }
        else { check(nsb_upload_mesh_fvcr(m_ctx, elem_type, n_elem, n_node, n_side, conn, elem_sides, coords)); m_nsh = nside[elem_type]; m_L = m_nsh * m_dim + 1; }
    }
    int64_t num_dofs() const { return nsb_num_dofs(m_ctx); }
    int64_t nnz() const { return nsb_nnz(m_ctx); }
    void get_csr(int64_t* rowptr, int32_t* colind) { check(nsb_get_csr(m_ctx, rowptr, colind)); }

    // ---- fast mode: whole-loop hooks (the role of DomainDiscretization::assemble_jacobian / _defect) ----
    void set_time_series(const number* sol0, const number* sol1, number dt) { m_ts.sol0 = sol0; m_ts.sol1 = sol1; m_ts.dt = dt; }
    void assemble_jacobian(number* values, const number* u, number s_a = 1.0, int scatter = NSB_SCATTER_GATHER)
    { push_params(); check(nsb_assemble(m_ctx, NSB_JAC_A, scatter, u, m_ts.sol0 ? &m_ts : nullptr, s_a, 1.0, 0.0, values, nullptr, NSB_HOST)); }
    void assemble_defect(number* defect, const number* u, number s_a = 1.0, int scatter = NSB_SCATTER_GATHER)
    { push_params(); check(nsb_assemble(m_ctx, NSB_DEF_A | NSB_RHS, scatter, u, m_ts.sol0 ? &m_ts : nullptr, s_a, 1.0, 0.0, nullptr, defect, NSB_HOST)); }

    // ---- compat mode: the IElemDisc slots ----
    /// the solution the element loop is about to assemble at (ugcore hands it per element; the batch needs it up front)
    void set_solution(const number* u) { m_u = u; }
    void prep_elem_loop(ReferenceObjectID /*roid*/, int /*si*/)
    {
        push_params();
        check(nsb_prep_elem_loop(m_ctx));                         // throws like fv1/navier_stokes_fv1.cpp:142-181
        if (m_prm.disc != NSB_DISC_FV1) NSB_UG_THROW("compat mode: FV1 only; use the fast mode for FVCR");
        if (!m_u) NSB_UG_THROW("compat mode: set_solution(u) must precede prep_elem_loop");
        const size_t nJ = (size_t)m_n_elem * m_L * m_L, nd = (size_t)m_n_elem * m_L;
        m_JA.assign(nJ, 0.0); m_dA.assign(nd, 0.0); m_JM.assign(nJ, 0.0); m_dM.assign(nd, 0.0); m_rhs.assign(nd, 0.0);
        std::vector<number> scratch(nJ);
        const nsb_time_series* ts = m_ts.sol0 ? &m_ts : nullptr;
        check(nsb_local_contributions(m_ctx, NSB_JAC_A | NSB_DEF_A, m_u, ts, m_JA.data(), m_dA.data(), NSB_HOST));
        check(nsb_local_contributions(m_ctx, NSB_JAC_M | NSB_DEF_M, m_u, ts, m_JM.data(), m_dM.data(), NSB_HOST));
        check(nsb_local_contributions(m_ctx, NSB_RHS, m_u, ts, scratch.data(), m_rhs.data(), NSB_HOST));
        for (auto& x : m_rhs) x = -x;                             // nsb returns -rhs in the defect slot
    }
    void prep_elem(const LocalVector& /*u*/, int64_t elem_index, ReferenceObjectID /*roid*/, const void* /*vCornerCoords*/)
    {
        if (elem_index < 0 || elem_index >= m_n_elem) NSB_UG_THROW("NavierStokes::prep_elem: element not part of the uploaded grid");
        m_cur = elem_index;
    }
    void add_jac_A_elem(LocalMatrix& J, const LocalVector&) { add_mat(J, m_JA); }
    void add_jac_M_elem(LocalMatrix& J, const LocalVector&) { add_mat(J, m_JM); }
    void add_def_A_elem(LocalVector& d, const LocalVector&) { add_vec(d, m_dA); }
    void add_def_M_elem(LocalVector& d, const LocalVector&) { add_vec(d, m_dM); }
    void add_rhs_elem(LocalVector& d) { add_vec(d, m_rhs); }
    void fsh_elem_loop() {}                                       // fv1/navier_stokes_fv1.cpp:201-205
    /// FVCR adapter: validation only (the device path has no per-element Crouzeix-Raviart blocks)
    void prep_elem_loop_fast_only() { push_params(); check(nsb_prep_elem_loop(m_ctx)); }

    const char* last_error() const { return nsb_last_error(m_ctx); }

  protected:
    nsb_ctx* ctx()
    {
        if (!m_ctx && nsb_create(m_device, &m_ctx) != NSB_OK) NSB_UG_THROW(nsb_last_error(nullptr));
        return m_ctx;
    }
    void check(int rc) { if (rc != NSB_OK) NSB_UG_THROW(std::string(nsb_last_error(m_ctx))); }
    virtual void resolve(nsb_params&) {}
    void push_params() { nsb_params p = m_prm; resolve(p); check(nsb_set_params(ctx(), &p)); }
    void add_mat(LocalMatrix& J, const std::vector<number>& src)
    {
        const number* b = src.data() + (size_t)m_cur * m_L * m_L;     // index fct*nsh+sh in both directions
        const int nf = m_dim + 1;
        for (int rf = 0; rf < nf; rf++) for (int rs = 0; rs < m_nsh; rs++)
            for (int cf = 0; cf < nf; cf++) for (int cs = 0; cs < m_nsh; cs++)
                J(rf, rs, cf, cs) += b[(size_t)(rf * m_nsh + rs) * m_L + cf * m_nsh + cs];
    }
    void add_vec(LocalVector& d, const std::vector<number>& src)
    {
        const number* b = src.data() + (size_t)m_cur * m_L;
        for (int f = 0; f < m_dim + 1; f++) for (int s = 0; s < m_nsh; s++) d(f, s) += b[f * m_nsh + s];
    }

    nsb_params m_prm;
    nsb_time_series m_ts{nullptr, nullptr, 0.0};
    nsb_ctx* m_ctx = nullptr;
    int m_dim, m_device, m_elem = -1, m_nsh = 0, m_L = 0;
    int64_t m_n_elem = 0, m_cur = 0;
    const number* m_u = nullptr;
    std::string m_subsets;
    std::vector<number> m_JA, m_dA, m_JM, m_dM, m_rhs;
};

/// incompressible/fv1/navier_stokes_fv1.h
template <int dim> class NavierStokesFV1 : public NavierStokesDeviceDisc {
  public:
    NavierStokesFV1(const char* fcts, const char* subsets, int device = 0) : NavierStokesDeviceDisc(fcts, subsets, NSB_DISC_FV1, dim, device) {}
    std::string disc_type() const override { return "fv1"; }
    // :190-199
    // a fresh stabilisation object: its upwind is the convective one if that is valid, EMPTY otherwise (never a stale one)
    void set_stabilization(const std::string& name) { m_prm.stab = stab_id(name); m_prm.diff_length = NSB_DIFF_RAW; m_stab_upwind = m_conv; }
    void set_stabilization(const std::string& name, const std::string& diff) { set_stabilization(name); m_prm.diff_length = diff_length_id(diff); }
    void set_no_stabilization() { m_prm.stab = NSB_STAB_NONE; m_stab_upwind = m_conv; }     // NavierStokesFV1WithoutStabilization
    void set_stabilization_upwind(const std::string& name) { m_stab_upwind = upwind_id(name); }   // stab->set_upwind(...)
    // :213-215
    void set_upwind(const std::string& name) { m_pac = false; m_conv = upwind_id(name); if (m_prm.stab != NSB_STAB_UNSET && !m_stab_upwind) m_stab_upwind = m_conv; }
    // :217-225
    void set_pac_upwind(bool b)
    {
        if (!b) return;
        if (!m_conv) NSB_UG_THROW("Upwind must be specified previously.\n");
        if (m_prm.stab == NSB_STAB_UNSET) NSB_UG_THROW("Stabilization must be specified previously.\n");
        m_stab_upwind = m_conv; m_pac = true;
    }
  protected:
    void resolve(nsb_params& p) override { p.conv_upwind = m_conv; p.stab_upwind = m_stab_upwind; p.pac_upwind = m_pac; }
    int m_conv = NSB_UPWIND_UNSET, m_stab_upwind = NSB_UPWIND_UNSET; bool m_pac = false;
};

/// incompressible/fvcr/navier_stokes_fvcr.h
template <int dim> class NavierStokesFVCR : public NavierStokesDeviceDisc {
  public:
    NavierStokesFVCR(const char* fcts, const char* subsets, int device = 0) : NavierStokesDeviceDisc(fcts, subsets, NSB_DISC_FVCR, dim, device) {}
    std::string disc_type() const override { return "fvcr"; }
    bool use_hanging() const { return true; }                       // fvcr/navier_stokes_fvcr.cpp:111-116
    void set_upwind(const std::string& name) { m_prm.conv_upwind = upwind_id(name); }
    void set_defect_upwind(bool b) { m_prm.defect_upwind = b; }
};

}  // namespace nsb200
