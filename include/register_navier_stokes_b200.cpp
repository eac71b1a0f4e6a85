// register_navier_stokes_b200.cpp -- the reference-side binding: IElemDisc adapters of the device assembly path and their
// registration under the UG4 NavierStokes plugin's registry names, so that an existing Lua script
// (`NavierStokes(fcts, subsets, "fv1")`, lua/lua-include.lua:36-47, and every setter it calls) resolves to the device classes.
//
// Compile inside the UG4 plugin tree with -DNSB_WITH_UG4 (ugcore headers), or here with -DNSB_WITH_UG4 -DNSB_UG4_MOCK against
// the mock of tests/cpp/mock_ug (tests/test_binding.py: compiles, runs the registration, compares names / groups / overloads
// with the reference's registration files and drives the slots through IElemDisc's dispatch on a GPU box).
//
// Registration mirrored (same names, bases, groups, constructors, overloads):
//   register_navier_stokes.cpp:105-126 (NavierStokesBase), :156-228 (upwind classes),
//   incompressible/incompressible_navier_stokes_plugin.cpp:244-266 (IncompressibleNavierStokesBase; the data exports velocity /
//   pressure / ... are not part of the assembly path and not provided),
//   incompressible/fv1/register_fv1.cpp:166-184 (NavierStokesFV1), :229-280 (stabilisation classes),
//   incompressible/fvcr/register_fvcr.cpp:289-303 (NavierStokesFVCR), register_navier_stokes.cpp:249-264 (InitUGPlugin_NavierStokes).
#ifdef NSB_UG4_MOCK
#include "ug_mock.h"
#else
#include "bridge/bridge.h"
#include "bridge/util.h"
#include "bridge/util_domain_dependent.h"
#include "lib_disc/spatial_disc/elem_disc/elem_disc_interface.h"
#include "lib_disc/spatial_disc/user_data/user_data.h"
#endif
#include <map>
#include <string>
#include <vector>
#include "navier_stokes_b200.hpp"

using namespace std;

namespace ug {
namespace NavierStokes {

#ifdef NSB_UG4_MOCK
// ------------------------------------------------------------------------------------------------
// (A) classes that EXIST in the UG4 plugin tree (upwind.h, fv1/stabilization.h, navier_stokes_base.h,
//     incompressible_navier_stokes_base.h). In the plugin build the originals are used; the mock build needs stand-ins
//     with the script-visible surface (they carry parameters only: the arithmetic lives in the CUDA kernels).
// ------------------------------------------------------------------------------------------------
template <int dim> class INavierStokesUpwind { public: virtual ~INavierStokesUpwind() {} virtual int nsb_id() const = 0; };
#define NSB_UPWIND_CLASS(Name, Id) template <int dim> class Name : public INavierStokesUpwind<dim> { public: int nsb_id() const { return Id; } };
NSB_UPWIND_CLASS(NavierStokesNoUpwind, NSB_UPWIND_NO)
NSB_UPWIND_CLASS(NavierStokesFullUpwind, NSB_UPWIND_FULL)
NSB_UPWIND_CLASS(NavierStokesSkewedUpwind, NSB_UPWIND_SKEWED)
NSB_UPWIND_CLASS(NavierStokesLinearProfileSkewedUpwind, NSB_UPWIND_LPS)
NSB_UPWIND_CLASS(NavierStokesPositiveUpwind, NSB_UPWIND_POSITIVE)
NSB_UPWIND_CLASS(NavierStokesRegularUpwind, 6)
#undef NSB_UPWIND_CLASS

template <int dim> class INavierStokesFV1Stabilization {
  public:
    virtual ~INavierStokesFV1Stabilization() {}
    void set_upwind(SmartPtr<INavierStokesUpwind<dim> > spUpwind) { m_spUpwind = spUpwind; }      // fv1/stabilization.h
    SmartPtr<INavierStokesUpwind<dim> > upwind() const { return m_spUpwind; }
    virtual int nsb_id() const = 0;
    virtual int nsb_diff_length() const { return NSB_DIFF_RAW; }
  protected:
    SmartPtr<INavierStokesUpwind<dim> > m_spUpwind;
};
template <int dim> class INavierStokesSRFV1Stabilization : public INavierStokesFV1Stabilization<dim> {
  public:
    void set_diffusion_length(std::string diffLength) { m_diff = nsb200::diff_length_id(diffLength); }   // fv1/stabilization.cpp:86-100
    int nsb_diff_length() const { return m_diff; }
  private:
    int m_diff = NSB_DIFF_RAW;
};
template <int dim> class NavierStokesFIELDSStabilization : public INavierStokesSRFV1Stabilization<dim> { public: int nsb_id() const { return NSB_STAB_FIELDS; } };
template <int dim> class NavierStokesFLOWStabilization : public INavierStokesSRFV1Stabilization<dim> { public: int nsb_id() const { return NSB_STAB_FLOW; } };
template <int dim> class NavierStokesFV1WithoutStabilization : public INavierStokesFV1Stabilization<dim> { public: int nsb_id() const { return NSB_STAB_NONE; } };

// navier_stokes_base.h:141-208: the import setters forward to the device object of the derived class
template <typename TDomain> class NavierStokesBase : public IElemDisc<TDomain> {
  public:
    static const int dim = TDomain::dim;
    NavierStokesBase(const char* functions, const char* subsets) : IElemDisc<TDomain>(functions, subsets) {}
    NavierStokesBase(const std::vector<std::string>& vFct, const std::vector<std::string>& vSubset) : IElemDisc<TDomain>(vFct, vSubset) {}
    virtual nsb200::NavierStokesDeviceDisc& dev() = 0;
    void set_kinematic_viscosity(SmartPtr<CplUserData<number, dim> > user)
    {
        if (!user.valid() || !user->constant()) UG_THROW("NavierStokes (device path): only constant kinematic viscosity data is supported");
        dev().set_kinematic_viscosity(user->const_value());
    }
    void set_kinematic_viscosity(number val) { dev().set_kinematic_viscosity(val); }
#ifdef UG_FOR_LUA
    void set_kinematic_viscosity(const char*) { UG_THROW("NavierStokes (device path): Lua callbacks cannot run on the device"); }
#endif
    void set_source(SmartPtr<CplUserData<MathVector<dim>, dim> > user)
    {
        if (!user.valid() || !user->constant()) UG_THROW("NavierStokes (device path): only constant source data is supported");
        const MathVector<dim> v = user->const_value();
        std::vector<number> f(dim); for (int d = 0; d < dim; d++) f[d] = v[d];
        dev().set_source(f);
    }
    void set_source(const std::vector<number>& vSource) { dev().set_source(vSource); }
#ifdef UG_FOR_LUA
    void set_source(const char*) { UG_THROW("NavierStokes (device path): Lua callbacks cannot run on the device"); }
#endif
    virtual std::string disc_type() const = 0;
    void set_exact_jacobian(bool bExactJacobian) { dev().set_exact_jacobian(bExactJacobian); }
    void set_exact_jacobian(number fullNewtonFactor) { dev().set_exact_jacobian(fullNewtonFactor); }
    virtual bool requests_local_time_series() { return true; }                  // navier_stokes_base.h:200
};

// incompressible/incompressible_navier_stokes_base.h:144-284
template <typename TDomain> class IncompressibleNavierStokesBase : public NavierStokesBase<TDomain> {
  public:
    static const int dim = TDomain::dim;
    IncompressibleNavierStokesBase(const char* functions, const char* subsets) : NavierStokesBase<TDomain>(functions, subsets) {}
    IncompressibleNavierStokesBase(const std::vector<std::string>& vFct, const std::vector<std::string>& vSubset) : NavierStokesBase<TDomain>(vFct, vSubset) {}
    void set_density(SmartPtr<CplUserData<number, dim> > user)
    {
        if (!user.valid() || !user->constant()) UG_THROW("NavierStokes (device path): only constant density data is supported");
        this->dev().set_density(user->const_value());
    }
    void set_density(number val) { this->dev().set_density(val); }
#ifdef UG_FOR_LUA
    void set_density(const char*) { UG_THROW("NavierStokes (device path): Lua callbacks cannot run on the device"); }
#endif
    void set_peclet_blend(bool pecletBlend) { this->dev().set_peclet_blend(pecletBlend); }
    void set_grad_div(number factor) { this->dev().set_grad_div(factor); }
    void set_laplace(bool bLaplace) { this->dev().set_laplace(bLaplace); }
    void set_stokes(bool bStokes) { this->dev().set_stokes(bStokes); }
};
#endif  // NSB_UG4_MOCK

static std::string join(const std::vector<std::string>& v)
{
    std::string s;
    for (size_t i = 0; i < v.size(); i++) { if (i) s += ","; s += v[i]; }
    return s;
}

// ------------------------------------------------------------------------------------------------
// (B) the adapters: IElemDisc whose slots forward to the device shim (compat mode), plus the whole-loop hooks (fast mode)
// ------------------------------------------------------------------------------------------------
/// drop-in for NavierStokesFV1<TDomain> (incompressible/fv1/navier_stokes_fv1.h:185-488)
template <typename TDomain>
class NavierStokesFV1 : public IncompressibleNavierStokesBase<TDomain> {
  public:
    static const int dim = TDomain::dim;
    typedef IncompressibleNavierStokesBase<TDomain> base_type;
    NavierStokesFV1(const char* functions, const char* subsets) : base_type(functions, subsets), m_dev(functions, subsets) { register_all_funcs(); }
    NavierStokesFV1(const std::vector<std::string>& vFct, const std::vector<std::string>& vSubset)
        : base_type(vFct, vSubset), m_dev(join(vFct).c_str(), join(vSubset).c_str()) { register_all_funcs(); }
    nsb200::NavierStokesDeviceDisc& dev() { return m_dev; }
    nsb200::NavierStokesFV1<dim>& device() { return m_dev; }
    virtual std::string disc_type() const { return "fv1"; }

    // fv1/navier_stokes_fv1.h:185-225 -------------------------------------------------------------
    void set_stabilization(SmartPtr<INavierStokesFV1Stabilization<dim> > spStab)
    {
        m_spStab = spStab;
        if (spStab->nsb_id() == NSB_STAB_NONE) m_dev.set_no_stabilization();
        else m_dev.set_stabilization(spStab->nsb_id() == NSB_STAB_FIELDS ? "fields" : "flow",
                                     spStab->nsb_diff_length() == NSB_DIFF_RAW ? "raw" : (spStab->nsb_diff_length() == NSB_DIFF_FIVEPOINT ? "fivepoint" : "cor"));
        if (spStab->upwind().valid()) m_dev.set_stabilization_upwind(name_of(spStab->upwind()->nsb_id()));
    }
    void set_stabilization(const std::string& name) { m_dev.set_stabilization(name); m_spStab = SmartPtr<INavierStokesFV1Stabilization<dim> >(); }
    void set_stabilization(const std::string& name, const std::string& diffLength) { m_dev.set_stabilization(name, diffLength); m_spStab = SmartPtr<INavierStokesFV1Stabilization<dim> >(); }
    void set_upwind(SmartPtr<INavierStokesFV1Stabilization<dim> > spStab)         // PAC: the stabilisation itself supplies the convective velocity
    {
        if (m_spStab.get() != spStab.get()) UG_THROW("NavierStokes (device path): a convective stabilisation different from the continuity stabilisation is not supported");
        if (spStab->upwind().invalid()) UG_THROW("Upwind must be specified previously.\n");
        m_dev.set_upwind(name_of(spStab->upwind()->nsb_id()));
        m_dev.set_pac_upwind(true);
    }
    void set_upwind(SmartPtr<INavierStokesUpwind<dim> > spUpwind) { m_dev.set_upwind(name_of(spUpwind->nsb_id())); }
    void set_upwind(const std::string& name) { m_dev.set_upwind(name); }
    void set_pac_upwind(bool bPac) { m_dev.set_pac_upwind(bPac); }

    // grid hand-over: element order = batch index of the slots. (In the plugin build this is filled from the DoFDistribution
    // and the position accessor at the start of an assembling; see INTEGRATION.md.)
    void set_grid(int elem_type, const std::vector<GridObject*>& elems, int64_t n_node, const int32_t* conn, const number* coords)
    {
        m_dev.set_grid(elem_type, (int64_t)elems.size(), n_node, conn, coords);
        m_elemIndex.clear();
        for (size_t i = 0; i < elems.size(); i++) m_elemIndex[elems[i]] = (int64_t)i;
    }
    void set_solution(const number* u) { m_dev.set_solution(u); }

    // the slots, signatures of fv1/navier_stokes_fv1.h:254-488 -------------------------------------
    template <typename TElem, typename TFVGeom> void prep_elem_loop(const ReferenceObjectID roid, const int si) { m_dev.prep_elem_loop(roid, si); }
    template <typename TElem, typename TFVGeom> void prep_elem(const LocalVector& u, GridObject* elem, const ReferenceObjectID roid, const MathVector<dim> vCornerCoords[])
    {
        typename std::map<GridObject*, int64_t>::const_iterator it = m_elemIndex.find(elem);
        if (it == m_elemIndex.end()) UG_THROW("NavierStokes::prep_elem: element not part of the uploaded grid");
        m_dev.prep_elem(u, it->second, roid, vCornerCoords);
    }
    template <typename TElem, typename TFVGeom> void fsh_elem_loop() { m_dev.fsh_elem_loop(); }
    template <typename TElem, typename TFVGeom> void add_jac_A_elem(LocalMatrix& J, const LocalVector& u, GridObject* elem, const MathVector<dim> vCornerCoords[]) { m_dev.add_jac_A_elem(J, u); }
    template <typename TElem, typename TFVGeom> void add_def_A_elem(LocalVector& d, const LocalVector& u, GridObject* elem, const MathVector<dim> vCornerCoords[]) { m_dev.add_def_A_elem(d, u); }
    template <typename TElem, typename TFVGeom> void add_jac_M_elem(LocalMatrix& J, const LocalVector& u, GridObject* elem, const MathVector<dim> vCornerCoords[]) { m_dev.add_jac_M_elem(J, u); }
    template <typename TElem, typename TFVGeom> void add_def_M_elem(LocalVector& d, const LocalVector& u, GridObject* elem, const MathVector<dim> vCornerCoords[]) { m_dev.add_def_M_elem(d, u); }
    template <typename TElem, typename TFVGeom> void add_rhs_elem(LocalVector& d, GridObject* elem, const MathVector<dim> vCornerCoords[]) { m_dev.add_rhs_elem(d); }

  private:
    static const char* name_of(int id)
    {
        switch (id) { case NSB_UPWIND_NO: return "no"; case NSB_UPWIND_FULL: return "full"; case NSB_UPWIND_SKEWED: return "skewed";
                      case NSB_UPWIND_LPS: return "lps"; case NSB_UPWIND_POSITIVE: return "pos"; default: return "reg"; }
    }
    // the registration of fv1/navier_stokes_fv1.cpp:1530-1572 (the element types the device path provides)
    void register_all_funcs()
    {
        if (dim == 2) { register_func<Triangle, void>(); register_func<Quadrilateral, void>(); }
        else { register_func<Tetrahedron, void>(); register_func<Hexahedron, void>(); register_func<Prism, void>(); }
    }
    template <typename TElem, typename TFVGeom> void register_func()
    {
        ReferenceObjectID id = (ReferenceObjectID)geometry_traits<TElem>::REFERENCE_OBJECT_ID;
        typedef NavierStokesFV1 T;
        this->clear_add_fct_once();
        this->set_prep_elem_loop_fct(id, &T::template prep_elem_loop<TElem, TFVGeom>);
        this->set_prep_elem_fct(id, &T::template prep_elem<TElem, TFVGeom>);
        this->set_fsh_elem_loop_fct(id, &T::template fsh_elem_loop<TElem, TFVGeom>);
        this->set_add_jac_A_elem_fct(id, &T::template add_jac_A_elem<TElem, TFVGeom>);
        this->set_add_jac_M_elem_fct(id, &T::template add_jac_M_elem<TElem, TFVGeom>);
        this->set_add_def_A_elem_fct(id, &T::template add_def_A_elem<TElem, TFVGeom>);
        this->set_add_def_M_elem_fct(id, &T::template add_def_M_elem<TElem, TFVGeom>);
        this->set_add_rhs_elem_fct(id, &T::template add_rhs_elem<TElem, TFVGeom>);
    }
    void clear_add_fct_once() {}
    nsb200::NavierStokesFV1<dim> m_dev;
    SmartPtr<INavierStokesFV1Stabilization<dim> > m_spStab;
    std::map<GridObject*, int64_t> m_elemIndex;
};

/// drop-in for NavierStokesFVCR<TDomain> (incompressible/fvcr/navier_stokes_fvcr.h). The device path assembles whole grids
/// (assemble_jacobian / assemble_defect of the shim = the IAssemble-level hooks); the per-element slots are registered for the
/// simplices and throw, because libnsb200 returns no per-element Crouzeix-Raviart blocks (nsb_local_contributions is FV1 only).
template <typename TDomain>
class NavierStokesFVCR : public IncompressibleNavierStokesBase<TDomain> {
  public:
    static const int dim = TDomain::dim;
    typedef IncompressibleNavierStokesBase<TDomain> base_type;
    NavierStokesFVCR(const char* functions, const char* subsets) : base_type(functions, subsets), m_dev(functions, subsets) { register_all_funcs(); }
    NavierStokesFVCR(const std::vector<std::string>& vFct, const std::vector<std::string>& vSubset)
        : base_type(vFct, vSubset), m_dev(join(vFct).c_str(), join(vSubset).c_str()) { register_all_funcs(); }
    nsb200::NavierStokesDeviceDisc& dev() { return m_dev; }
    nsb200::NavierStokesFVCR<dim>& device() { return m_dev; }
    virtual std::string disc_type() const { return "fvcr"; }
    virtual bool use_hanging() const { return true; }                           // fvcr/navier_stokes_fvcr.cpp:111-116
    void set_upwind(SmartPtr<INavierStokesUpwind<dim> > spUpwind)
    {
        static const char* nm[7] = {"", "no", "full", "skewed", "lps", "pos", "reg"};
        m_dev.set_upwind(nm[spUpwind->nsb_id()]);
    }
    void set_upwind(const std::string& name) { m_dev.set_upwind(name); }
    void set_defect_upwind(bool defectUpwind) { m_dev.set_defect_upwind(defectUpwind); }

    template <typename TElem, typename TFVGeom> void prep_elem_loop(const ReferenceObjectID roid, const int si) { m_dev.prep_elem_loop_fast_only(); }
    template <typename TElem, typename TFVGeom> void prep_elem(const LocalVector& u, GridObject* elem, const ReferenceObjectID roid, const MathVector<dim> vCornerCoords[]) { no_compat(); }
    template <typename TElem, typename TFVGeom> void fsh_elem_loop() {}
    template <typename TElem, typename TFVGeom> void add_jac_A_elem(LocalMatrix& J, const LocalVector& u, GridObject* elem, const MathVector<dim> vCornerCoords[]) { no_compat(); }
    template <typename TElem, typename TFVGeom> void add_def_A_elem(LocalVector& d, const LocalVector& u, GridObject* elem, const MathVector<dim> vCornerCoords[]) { no_compat(); }
    template <typename TElem, typename TFVGeom> void add_jac_M_elem(LocalMatrix& J, const LocalVector& u, GridObject* elem, const MathVector<dim> vCornerCoords[]) { no_compat(); }
    template <typename TElem, typename TFVGeom> void add_def_M_elem(LocalVector& d, const LocalVector& u, GridObject* elem, const MathVector<dim> vCornerCoords[]) { no_compat(); }
    template <typename TElem, typename TFVGeom> void add_rhs_elem(LocalVector& d, GridObject* elem, const MathVector<dim> vCornerCoords[]) { no_compat(); }

  private:
    static void no_compat() { UG_THROW("NavierStokesFVCR (device path): per-element slots are not available, assemble whole grids with assemble_jacobian / assemble_defect"); }
    void register_all_funcs()
    {
        if (dim == 2) { register_func<Triangle, void>(); register_func<Quadrilateral, void>(); }
        else { register_func<Tetrahedron, void>(); register_func<Hexahedron, void>(); }
    }
    template <typename TElem, typename TFVGeom> void register_func()
    {
        ReferenceObjectID id = (ReferenceObjectID)geometry_traits<TElem>::REFERENCE_OBJECT_ID;
        typedef NavierStokesFVCR T;
        this->set_prep_elem_loop_fct(id, &T::template prep_elem_loop<TElem, TFVGeom>);
        this->set_prep_elem_fct(id, &T::template prep_elem<TElem, TFVGeom>);
        this->set_fsh_elem_loop_fct(id, &T::template fsh_elem_loop<TElem, TFVGeom>);
        this->set_add_jac_A_elem_fct(id, &T::template add_jac_A_elem<TElem, TFVGeom>);
        this->set_add_jac_M_elem_fct(id, &T::template add_jac_M_elem<TElem, TFVGeom>);
        this->set_add_def_A_elem_fct(id, &T::template add_def_A_elem<TElem, TFVGeom>);
        this->set_add_def_M_elem_fct(id, &T::template add_def_M_elem<TElem, TFVGeom>);
        this->set_add_rhs_elem_fct(id, &T::template add_rhs_elem<TElem, TFVGeom>);
    }
    nsb200::NavierStokesFVCR<dim> m_dev;
};

// ------------------------------------------------------------------------------------------------
// (C) registration
// ------------------------------------------------------------------------------------------------
using namespace ug::bridge;

struct Functionality {

template <typename TDomain>
static void Domain(Registry& reg, string grp)
{
	static const int dim = TDomain::dim;
	string suffix = GetDomainSuffix<TDomain>();
	string tag = GetDomainTag<TDomain>();

#ifdef NSB_UG4_MOCK
//	Navier-Stokes Base (register_navier_stokes.cpp:105-126)
	{
		typedef NavierStokesBase<TDomain> T;
		typedef IElemDisc<TDomain> TBase;
		string name = string("NavierStokesBase").append(suffix);
		reg.add_class_<T, TBase >(name, grp)
			.add_method("set_kinematic_viscosity", static_cast<void (T::*)(SmartPtr<CplUserData<number, dim> >)>(&T::set_kinematic_viscosity), "", "KinematicViscosity")
			.add_method("set_kinematic_viscosity", static_cast<void (T::*)(number)>(&T::set_kinematic_viscosity), "", "KinematicViscosity")
#ifdef UG_FOR_LUA
			.add_method("set_kinematic_viscosity", static_cast<void (T::*)(const char*)>(&T::set_kinematic_viscosity), "", "KinematicViscosity")
#endif
			.add_method("set_source", static_cast<void (T::*)(SmartPtr<CplUserData<MathVector<dim>, dim> >)>(&T::set_source), "", "Source")
			.add_method("set_source", static_cast<void (T::*)(const std::vector<number>&)>(&T::set_source), "", "Source")
#ifdef UG_FOR_LUA
			.add_method("set_source", static_cast<void (T::*)(const char*)>(&T::set_source), "", "Source")
#endif
			.add_method("disc_type", &T::disc_type)
			.add_method("set_exact_jacobian", static_cast<void (T::*)(bool)>(&T::set_exact_jacobian), "", "ExactJacobian")
			.add_method("set_exact_jacobian", static_cast<void (T::*)(number)>(&T::set_exact_jacobian), "", "ExactJacobianFactor");
		reg.add_class_to_group(name, "NavierStokesBase", tag);
	}

//	Incompressible Navier-Stokes Base (incompressible_navier_stokes_plugin.cpp:244-266, setters of the assembly path)
	{
		typedef IncompressibleNavierStokesBase<TDomain> T;
		typedef NavierStokesBase<TDomain> TBase;
		string name = string("IncompressibleNavierStokesBase").append(suffix);
		reg.add_class_<T, TBase>(name, grp)
			.add_method("set_density", static_cast<void (T::*)(SmartPtr<CplUserData<number, dim> >)>(&T::set_density), "", "Density")
			.add_method("set_density", static_cast<void (T::*)(number)>(&T::set_density), "", "Density")
#ifdef UG_FOR_LUA
			.add_method("set_density", static_cast<void (T::*)(const char*)>(&T::set_density), "", "Density")
#endif
			.add_method("set_peclet_blend", &T::set_peclet_blend)
			.add_method("set_grad_div", static_cast<void (T::*)(number)>(&T::set_grad_div), "", "GradDivFactor")
			.add_method("set_laplace", &T::set_laplace)
			.add_method("set_stokes", &T::set_stokes);
		reg.add_class_to_group(name, "IncompressibleNavierStokesBase", tag);
	}
#endif

	//	Navier-Stokes FV1 (fv1/register_fv1.cpp:166-184)
	{
		typedef NavierStokesFV1<TDomain> T;
		typedef IncompressibleNavierStokesBase<TDomain> TBase;
		string name = string("NavierStokesFV1").append(suffix);
		reg.add_class_<T, TBase >(name, grp)
			.template add_constructor<void (*)(const char*,const char*)>("Functions#Subset(s)")
			.template add_constructor<void (*)(const std::vector<std::string>&, const std::vector<std::string>&)>("Functions#Subset(s)")
			.add_method("set_stabilization",  static_cast<void (T::*)(SmartPtr<INavierStokesFV1Stabilization<dim> >)>(&T::set_stabilization))
			.add_method("set_stabilization",  static_cast<void (T::*)(const std::string&)>(&T::set_stabilization))
			.add_method("set_stabilization",  static_cast<void (T::*)(const std::string&, const std::string&)>(&T::set_stabilization))
			.add_method("set_upwind",  static_cast<void (T::*)(SmartPtr<INavierStokesFV1Stabilization<dim> >)>(&T::set_upwind))
			.add_method("set_upwind",  static_cast<void (T::*)(SmartPtr<INavierStokesUpwind<dim> >)>(&T::set_upwind))
			.add_method("set_upwind",  static_cast<void (T::*)(const std::string&)>(&T::set_upwind))
			.add_method("set_pac_upwind", &T::set_pac_upwind, "", "Set pac upwind")
			.set_construct_as_smart_pointer(true);
		reg.add_class_to_group(name, "NavierStokesFV1", tag);
	}

	//	Navier-Stokes FVCR (fvcr/register_fvcr.cpp:289-303)
	{
		typedef NavierStokesFVCR<TDomain> T;
		typedef IncompressibleNavierStokesBase<TDomain> TBase;
		string name = string("NavierStokesFVCR").append(suffix);
		reg.add_class_<T, TBase >(name, grp)
			.template add_constructor<void (*)(const char*,const char*)>("Functions#Subset(s)")
			.template add_constructor<void (*)(const std::vector<std::string>&, const std::vector<std::string>&)>("Functions#Subset(s)")
			.add_method("set_upwind",  static_cast<void (T::*)(SmartPtr<INavierStokesUpwind<dim> >)>(&T::set_upwind))
			.add_method("set_upwind",  static_cast<void (T::*)(const std::string&)>(&T::set_upwind))
			.add_method("set_defect_upwind", &T::set_defect_upwind)
			.set_construct_as_smart_pointer(true);
		reg.add_class_to_group(name, "NavierStokesFVCR", tag);
	}
}

template <int dim>
static void Dimension(Registry& reg, string grp)
{
	string suffix = GetDimensionSuffix<dim>();
	string tag = GetDimensionTag<dim>();
#ifdef NSB_UG4_MOCK
//	upwind classes (register_navier_stokes.cpp:156-228)
	{
		typedef INavierStokesUpwind<dim> T;
		string name = string("INavierStokesUpwind").append(suffix);
		reg.add_class_<T>(name, grp);
		reg.add_class_to_group(name, "INavierStokesUpwind", tag);
	}
#define NSB_REG_UPWIND(Cls) { typedef Cls<dim> T; typedef INavierStokesUpwind<dim> TBase; string name = string(#Cls).append(suffix); \
		reg.add_class_<T, TBase>(name, grp).add_constructor().set_construct_as_smart_pointer(true); reg.add_class_to_group(name, #Cls, tag); }
	NSB_REG_UPWIND(NavierStokesNoUpwind)
	NSB_REG_UPWIND(NavierStokesFullUpwind)
	NSB_REG_UPWIND(NavierStokesSkewedUpwind)
	NSB_REG_UPWIND(NavierStokesLinearProfileSkewedUpwind)
	NSB_REG_UPWIND(NavierStokesPositiveUpwind)
	NSB_REG_UPWIND(NavierStokesRegularUpwind)
#undef NSB_REG_UPWIND
//	stabilisation classes (fv1/register_fv1.cpp:229-280)
	{
		typedef INavierStokesFV1Stabilization<dim> T;
		string name = string("INavierStokesFV1Stabilization").append(suffix);
		reg.add_class_<T>(name, grp)
			.add_method("set_upwind", &T::set_upwind);
		reg.add_class_to_group(name, "INavierStokesFV1Stabilization", tag);
	}
	{
		typedef INavierStokesSRFV1Stabilization<dim> T;
		typedef INavierStokesFV1Stabilization<dim> TBase;
		string name = string("INavierStokesSRFV1Stabilization").append(suffix);
		reg.add_class_<T, TBase>(name, grp)
			.add_method("set_diffusion_length", &T::set_diffusion_length);
		reg.add_class_to_group(name, "INavierStokesSRFV1Stabilization", tag);
	}
#define NSB_REG_STAB(Cls, Base) { typedef Cls<dim> T; typedef Base<dim> TBase; string name = string(#Cls).append(suffix); \
		reg.add_class_<T, TBase>(name, grp).add_constructor().set_construct_as_smart_pointer(true); reg.add_class_to_group(name, #Cls, tag); }
	NSB_REG_STAB(NavierStokesFIELDSStabilization, INavierStokesSRFV1Stabilization)
	NSB_REG_STAB(NavierStokesFLOWStabilization, INavierStokesSRFV1Stabilization)
	NSB_REG_STAB(NavierStokesFV1WithoutStabilization, INavierStokesFV1Stabilization)
#undef NSB_REG_STAB
#endif
}

};  // end Functionality
}  // namespace NavierStokes

/// called when the plugin is loaded (register_navier_stokes.cpp:249-264)
extern "C" void
InitUGPlugin_NavierStokes(ug::bridge::Registry* reg, string grp)
{
	grp.append("SpatialDisc/NavierStokes/");
	typedef NavierStokes::Functionality Functionality;

	try{
		ug::bridge::RegisterDimension2d3dDependent<Functionality>(*reg,grp);
		ug::bridge::RegisterDomain2d3dDependent<Functionality>(*reg,grp);
	}
	UG_REGISTRY_CATCH_THROW(grp);
}

}  // namespace ug
