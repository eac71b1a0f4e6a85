// ug_mock.h -- a MOCK of the handful of ugcore interfaces the NavierStokes plugin's registration and element-disc code
// touches (test infrastructure: lets include/register_navier_stokes_b200.cpp compile with -DNSB_WITH_UG4 and run without
// ugcore). Shapes follow ugcore (lib_disc/spatial_disc/elem_disc/elem_disc_interface.h, bridge/bridge.h, registry/registry.h,
// common/util/smart_pointer.h, lib_disc/common/local_algebra.h); nothing here is copied, only the call syntax is kept.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <typeinfo>
#include <vector>

namespace ug {
typedef double number;
struct UGError : std::runtime_error { using std::runtime_error::runtime_error; };
#define UG_THROW(msg) do { std::stringstream ss__; ss__ << msg; throw ::ug::UGError(ss__.str()); } while (0)
#define UG_REGISTRY_CATCH_THROW(grp) catch (const ::ug::UGError& e__) { throw; }

template <class T> class SmartPtr {
    std::shared_ptr<T> p;
    template <class U> friend class SmartPtr;
  public:
    SmartPtr() {}
    explicit SmartPtr(T* q) : p(q) {}
    SmartPtr(std::shared_ptr<T> q) : p(std::move(q)) {}
    template <class U, class = typename std::enable_if<std::is_convertible<U*, T*>::value>::type> SmartPtr(const SmartPtr<U>& o) : p(o.p) {}
    T* operator->() const { return p.get(); }
    T& operator*() const { return *p; }
    T* get() const { return p.get(); }
    bool valid() const { return (bool)p; }
    bool invalid() const { return !p; }
    template <class U> SmartPtr<U> cast_dynamic() const { return SmartPtr<U>(std::dynamic_pointer_cast<U>(p)); }
};
template <class T> using ConstSmartPtr = SmartPtr<const T>;
template <class T, class... A> SmartPtr<T> make_sp(A&&... a) { return SmartPtr<T>(std::make_shared<T>(std::forward<A>(a)...)); }

template <int N> struct MathVector { number v[N]; number& operator[](int i) { return v[i]; } const number& operator[](int i) const { return v[i]; } };

enum ReferenceObjectID { ROID_UNKNOWN = -1, ROID_VERTEX, ROID_EDGE, ROID_TRIANGLE, ROID_QUADRILATERAL, ROID_TETRAHEDRON, ROID_HEXAHEDRON,
                         ROID_PRISM, ROID_PYRAMID, ROID_OCTAHEDRON, NUM_REFERENCE_OBJECTS };
struct GridObject { virtual ~GridObject() {} };
struct Triangle : GridObject {}; struct Quadrilateral : GridObject {}; struct Tetrahedron : GridObject {}; struct Hexahedron : GridObject {}; struct Prism : GridObject {};
template <class TElem> struct geometry_traits;
template <> struct geometry_traits<Triangle> { enum { REFERENCE_OBJECT_ID = ROID_TRIANGLE }; };
template <> struct geometry_traits<Quadrilateral> { enum { REFERENCE_OBJECT_ID = ROID_QUADRILATERAL }; };
template <> struct geometry_traits<Tetrahedron> { enum { REFERENCE_OBJECT_ID = ROID_TETRAHEDRON }; };
template <> struct geometry_traits<Hexahedron> { enum { REFERENCE_OBJECT_ID = ROID_HEXAHEDRON }; };
template <> struct geometry_traits<Prism> { enum { REFERENCE_OBJECT_ID = ROID_PRISM }; };

// lib_disc/common/local_algebra.h access syntax: u(fct, dof), J(rfct, rdof, cfct, cdof)
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

struct Domain2d { static const int dim = 2; };
struct Domain3d { static const int dim = 3; };

// user data: only constant data can be evaluated by this mock
template <typename TData, int dim> struct CplUserData { virtual ~CplUserData() {} virtual bool constant() const { return false; } virtual TData const_value() const { return TData(); } };
template <int dim> struct ConstUserNumber : CplUserData<number, dim> { number c; explicit ConstUserNumber(number v) : c(v) {} bool constant() const override { return true; } number const_value() const override { return c; } };
template <int dim> struct ConstUserVector : CplUserData<MathVector<dim>, dim> { MathVector<dim> c; bool constant() const override { return true; } MathVector<dim> const_value() const override { return c; } };

// IElemDisc: the slot tables of elem_disc_interface.h (member-function pointers per reference element) and the do_* dispatch
template <typename TDomain> class IElemDisc {
  public:
    static const int dim = TDomain::dim;
    typedef IElemDisc<TDomain> T;
    typedef void (T::*PrepareElemLoopFct)(ReferenceObjectID roid, int si);
    typedef void (T::*PrepareElemFct)(const LocalVector& u, GridObject* elem, const ReferenceObjectID roid, const MathVector<dim> vCornerCoords[]);
    typedef void (T::*FinishElemLoopFct)();
    typedef void (T::*ElemJAFct)(LocalMatrix& J, const LocalVector& u, GridObject* elem, const MathVector<dim> vCornerCoords[]);
    typedef void (T::*ElemdAFct)(LocalVector& d, const LocalVector& u, GridObject* elem, const MathVector<dim> vCornerCoords[]);
    typedef void (T::*ElemRHSFct)(LocalVector& rhs, GridObject* elem, const MathVector<dim> vCornerCoords[]);

    IElemDisc(const char* functions, const char* subsets) { split(functions, m_vFct); split(subsets, m_vSubset); clear(); }
    IElemDisc(const std::vector<std::string>& vFct, const std::vector<std::string>& vSubset) : m_vFct(vFct), m_vSubset(vSubset) { clear(); }
    virtual ~IElemDisc() {}
    const std::vector<std::string>& symb_fcts() const { return m_vFct; }
    const std::vector<std::string>& symb_subsets() const { return m_vSubset; }
    size_t num_fct() const { return m_vFct.size(); }
    virtual bool requests_local_time_series() { return false; }
    virtual bool use_hanging() const { return false; }
    void clear_add_fct() { clear(); }

    template <class F> void set_prep_elem_loop_fct(ReferenceObjectID id, F f) { m_prepLoop[id] = static_cast<PrepareElemLoopFct>(f); }
    template <class F> void set_prep_elem_fct(ReferenceObjectID id, F f) { m_prep[id] = static_cast<PrepareElemFct>(f); }
    template <class F> void set_fsh_elem_loop_fct(ReferenceObjectID id, F f) { m_fsh[id] = static_cast<FinishElemLoopFct>(f); }
    template <class F> void set_add_jac_A_elem_fct(ReferenceObjectID id, F f) { m_jacA[id] = static_cast<ElemJAFct>(f); }
    template <class F> void set_add_jac_M_elem_fct(ReferenceObjectID id, F f) { m_jacM[id] = static_cast<ElemJAFct>(f); }
    template <class F> void set_add_def_A_elem_fct(ReferenceObjectID id, F f) { m_defA[id] = static_cast<ElemdAFct>(f); }
    template <class F> void set_add_def_M_elem_fct(ReferenceObjectID id, F f) { m_defM[id] = static_cast<ElemdAFct>(f); }
    template <class F> void set_add_rhs_elem_fct(ReferenceObjectID id, F f) { m_rhs[id] = static_cast<ElemRHSFct>(f); }

    bool has_slots(ReferenceObjectID id) const { return m_prepLoop[id] && m_prep[id] && m_fsh[id] && m_jacA[id] && m_jacM[id] && m_defA[id] && m_defM[id] && m_rhs[id]; }
    void do_prep_elem_loop(ReferenceObjectID id, int si) { need(m_prepLoop[id]); (this->*m_prepLoop[id])(id, si); m_roid = id; }
    void do_prep_elem(const LocalVector& u, GridObject* e, const MathVector<dim> c[]) { (this->*m_prep[m_roid])(u, e, m_roid, c); }
    void do_fsh_elem_loop() { (this->*m_fsh[m_roid])(); }
    void do_add_jac_A_elem(LocalMatrix& J, const LocalVector& u, GridObject* e, const MathVector<dim> c[]) { (this->*m_jacA[m_roid])(J, u, e, c); }
    void do_add_jac_M_elem(LocalMatrix& J, const LocalVector& u, GridObject* e, const MathVector<dim> c[]) { (this->*m_jacM[m_roid])(J, u, e, c); }
    void do_add_def_A_elem(LocalVector& d, const LocalVector& u, GridObject* e, const MathVector<dim> c[]) { (this->*m_defA[m_roid])(d, u, e, c); }
    void do_add_def_M_elem(LocalVector& d, const LocalVector& u, GridObject* e, const MathVector<dim> c[]) { (this->*m_defM[m_roid])(d, u, e, c); }
    void do_add_rhs_elem(LocalVector& r, GridObject* e, const MathVector<dim> c[]) { (this->*m_rhs[m_roid])(r, e, c); }

  private:
    template <class P> static void need(P p) { if (!p) UG_THROW("IElemDisc: no function registered for this reference element"); }
    static void split(const char* s, std::vector<std::string>& out)
    {
        std::string cur;
        for (const char* c = s; c && *c; ++c) { if (*c == ',') { if (!cur.empty()) out.push_back(cur); cur.clear(); } else if (*c != ' ' && *c != '\t') cur.push_back(*c); }
        if (!cur.empty()) out.push_back(cur);
    }
    void clear()
    {
        for (int i = 0; i < NUM_REFERENCE_OBJECTS; i++) { m_prepLoop[i] = nullptr; m_prep[i] = nullptr; m_fsh[i] = nullptr; m_jacA[i] = m_jacM[i] = nullptr; m_defA[i] = m_defM[i] = nullptr; m_rhs[i] = nullptr; }
    }
    std::vector<std::string> m_vFct, m_vSubset;
    PrepareElemLoopFct m_prepLoop[NUM_REFERENCE_OBJECTS]; PrepareElemFct m_prep[NUM_REFERENCE_OBJECTS]; FinishElemLoopFct m_fsh[NUM_REFERENCE_OBJECTS];
    ElemJAFct m_jacA[NUM_REFERENCE_OBJECTS], m_jacM[NUM_REFERENCE_OBJECTS]; ElemdAFct m_defA[NUM_REFERENCE_OBJECTS], m_defM[NUM_REFERENCE_OBJECTS];
    ElemRHSFct m_rhs[NUM_REFERENCE_OBJECTS];
    ReferenceObjectID m_roid = ROID_UNKNOWN;
};

// ---- bridge::Registry: records what is registered (class, bases, group, constructors, methods, class groups) ----
namespace bridge {
struct MockMethod { std::string name, signature; };
struct MockClass {
    std::string name, group; std::vector<std::string> bases; std::vector<std::string> constructors; std::vector<MockMethod> methods;
    bool smart_ptr = false;
};
template <class T> class ExportedClass {
    MockClass* c;
  public:
    explicit ExportedClass(MockClass* m) : c(m) {}
    ExportedClass& add_constructor(const std::string& = "") { c->constructors.push_back("void (*)()"); return *this; }
    template <class Sig> ExportedClass& add_constructor(const std::string& = "", const std::string& = "", const std::string& = "") { c->constructors.push_back(typeid(Sig).name()); return *this; }
    template <class M> ExportedClass& add_method(const std::string& name, M, const std::string& = "", const std::string& = "", const std::string& = "")
    { c->methods.push_back({name, typeid(M).name()}); return *this; }
    ExportedClass& set_construct_as_smart_pointer(bool b) { c->smart_ptr = b; return *this; }
};
class Registry {
  public:
    template <class T> ExportedClass<T> add_class_(const std::string& name, const std::string& grp) { return ExportedClass<T>(mk(name, grp, {})); }
    template <class T, class B> ExportedClass<T> add_class_(const std::string& name, const std::string& grp) { return ExportedClass<T>(mk(name, grp, {typeid(B).name()})); }
    template <class T, class B, class B2> ExportedClass<T> add_class_(const std::string& name, const std::string& grp) { return ExportedClass<T>(mk(name, grp, {typeid(B).name(), typeid(B2).name()})); }
    void add_class_to_group(const std::string& name, const std::string& group, const std::string& tag) { groups.push_back({name, group, tag}); }
    const MockClass* get_class(const std::string& name) const { auto it = index.find(name); return it == index.end() ? nullptr : classes[it->second].get(); }
    struct GroupEntry { std::string name, group, tag; };
    std::vector<std::unique_ptr<MockClass>> classes; std::vector<GroupEntry> groups; std::map<std::string, size_t> index;
  private:
    MockClass* mk(const std::string& name, const std::string& grp, std::vector<std::string> bases)
    {
        if (index.count(name)) UG_THROW("Registry: class '" << name << "' registered twice");
        classes.emplace_back(new MockClass()); MockClass* c = classes.back().get();
        c->name = name; c->group = grp; c->bases = std::move(bases); index[name] = classes.size() - 1;
        return c;
    }
};
template <class TDomain> std::string GetDomainSuffix() { return TDomain::dim == 2 ? "2d" : "3d"; }
template <class TDomain> std::string GetDomainTag() { return TDomain::dim == 2 ? "dim=2d;" : "dim=3d;"; }
template <int dim> std::string GetDimensionSuffix() { return dim == 2 ? "2d" : "3d"; }
template <int dim> std::string GetDimensionTag() { return dim == 2 ? "dim=2d;" : "dim=3d;"; }
template <class F> void RegisterDimension2d3dDependent(Registry& reg, std::string grp) { F::template Dimension<2>(reg, grp); F::template Dimension<3>(reg, grp); }
template <class F> void RegisterDomain2d3dDependent(Registry& reg, std::string grp) { F::template Domain<Domain2d>(reg, grp); F::template Domain<Domain3d>(reg, grp); }
}  // namespace bridge
}  // namespace ug
