// Drives include/register_navier_stokes_b200.cpp (compiled with -DNSB_WITH_UG4 -DNSB_UG4_MOCK against tests/cpp/mock_ug):
//   (no argument)  runs InitUGPlugin_NavierStokes on the mock registry, prints every registered class as one line
//                  "class <name> | group <grp> | bases <n> | ctors <n> | smart <0/1> | methods m1 m1 m2 ..." and "group <name> <group> <tag>",
//                  then exercises constructors, SmartPtr / string overloads and throw conditions without a GPU;
//   gpu <outfile>  assembles a small quadrilateral grid through the IElemDisc slot DISPATCH (do_prep_elem_loop, do_prep_elem,
//                  do_add_jac_A_elem, ...: what ugcore's element loop calls) and writes grid, state, Jacobian and defect to
//                  <outfile> for the comparison with the CPU oracle (tests/test_binding.py).
#include <cmath>
#include <cstdio>
#include <map>
#include "../../include/register_navier_stokes_b200.cpp"

using namespace ug;
using namespace ug::NavierStokes;
static int fails = 0;
#define EXPECT(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); fails++; } } while (0)
template <class F> static bool throws(F f, const char* needle)
{
    try { f(); } catch (const std::exception& e) { return std::string(e.what()).find(needle) != std::string::npos; }
    return false;
}

int main(int argc, char** argv)
{
    bridge::Registry reg;
    InitUGPlugin_NavierStokes(&reg, "/ug4/");
    if (argc == 1) {
        for (auto& c : reg.classes) {
            printf("class %s | group %s | bases %zu | ctors %zu | smart %d | methods", c->name.c_str(), c->group.c_str(), c->bases.size(), c->constructors.size(), (int)c->smart_ptr);
            for (auto& m : c->methods) printf(" %s", m.name.c_str());
            printf("\n");
        }
        for (auto& g : reg.groups) printf("group %s %s %s\n", g.name.c_str(), g.group.c_str(), g.tag.c_str());
        EXPECT(throws([&] { InitUGPlugin_NavierStokes(&reg, "/ug4/"); }, "registered twice"));
        // constructors (both), names, throw conditions of the setters, slot tables
        EXPECT(throws([] { NavierStokesFV1<Domain2d> d("u,v,w,p", "Inner"); }, "Wrong number of functions"));
        NavierStokesFV1<Domain2d> d("u, v, p", "Inner");
        NavierStokesFV1<Domain3d> d3(std::vector<std::string>{"u", "v", "w", "p"}, std::vector<std::string>{"Inner"});
        EXPECT(d.disc_type() == "fv1" && d3.disc_type() == "fv1" && d.requests_local_time_series() && d.symb_fcts().size() == 3);
        EXPECT(d.has_slots(ROID_TRIANGLE) && d.has_slots(ROID_QUADRILATERAL) && !d.has_slots(ROID_HEXAHEDRON));
        EXPECT(d3.has_slots(ROID_TETRAHEDRON) && d3.has_slots(ROID_HEXAHEDRON) && d3.has_slots(ROID_PRISM) && !d3.has_slots(ROID_PYRAMID) && !d3.has_slots(ROID_TRIANGLE));
        EXPECT(throws([&] { d.set_pac_upwind(true); }, "Upwind must be specified previously"));
        d.set_upwind(make_sp<NavierStokesLinearProfileSkewedUpwind<2> >());
        EXPECT(throws([&] { d.set_pac_upwind(true); }, "Stabilization must be specified previously"));
        SmartPtr<NavierStokesFIELDSStabilization<2> > st = make_sp<NavierStokesFIELDSStabilization<2> >();
        st->set_diffusion_length("COR");
        EXPECT(throws([&] { st->set_diffusion_length("foo"); }, "Diffusion Length"));
        st->set_upwind(make_sp<NavierStokesFullUpwind<2> >());
        d.set_stabilization(st);
        d.set_upwind(SmartPtr<INavierStokesFV1Stabilization<2> >(st));           // PAC through the stabilisation object
        EXPECT(throws([&] { d.set_upwind(SmartPtr<INavierStokesFV1Stabilization<2> >(make_sp<NavierStokesFLOWStabilization<2> >())); }, "different from the continuity"));
        EXPECT(throws([&] { d.set_upwind(std::string("central")); }, "not found"));
        EXPECT(throws([&] { d.set_stabilization(std::string("supg")); }, "not a valid name"));
        d.set_kinematic_viscosity(make_sp<ConstUserNumber<2> >(0.01));
        d.set_kinematic_viscosity(0.02);
        EXPECT(throws([&] { d.set_kinematic_viscosity(SmartPtr<CplUserData<number, 2> >(new CplUserData<number, 2>())); }, "only constant"));
        d.set_density(1.5); d.set_peclet_blend(true); d.set_laplace(false); d.set_stokes(false); d.set_grad_div(0.0);
        d.set_exact_jacobian(true); d.set_exact_jacobian(0.5); d.set_source(std::vector<number>{0.1, 0.2});
        NavierStokesFVCR<Domain3d> c("u,v,w,p", "Inner");
        NavierStokesFVCR<Domain2d> c2(std::vector<std::string>{"u", "v", "p"}, std::vector<std::string>{"Inner"});
        EXPECT(c.disc_type() == "fvcr" && c.use_hanging() && c.has_slots(ROID_TETRAHEDRON) && c.has_slots(ROID_HEXAHEDRON) && !c.has_slots(ROID_PRISM) && c2.has_slots(ROID_TRIANGLE) && c2.has_slots(ROID_QUADRILATERAL));
        c.set_upwind(make_sp<NavierStokesFullUpwind<3> >()); c.set_upwind(std::string("no")); c.set_defect_upwind(false);
        EXPECT(throws([&] { IElemDisc<Domain3d>& b = c; b.do_prep_elem_loop(ROID_PRISM, 0); }, "no function registered"));
        printf(fails ? "FAILED\n" : "OK\n");
        return fails ? 1 : 0;
    }
    // ---- GPU: ugcore-like element loop through the slot dispatch ----
    const int n = 5, nn = n + 1, nf = 3;
    std::vector<int32_t> conn; std::vector<double> xy, u;
    for (int j = 0; j < nn; j++) for (int i = 0; i < nn; i++) {
        xy.push_back(i / (double)n + 0.03 * std::sin(3.0 * i + j)); xy.push_back(j / (double)n + 0.02 * std::cos(2.0 * j + i));
        u.push_back(std::sin(1.0 + i)); u.push_back(std::cos(0.5 * j) - 0.3); u.push_back(0.1 * i * j);
    }
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) { int a = i + nn * j; conn.insert(conn.end(), {a, a + 1, a + 1 + nn, a + nn}); }
    std::vector<Quadrilateral> quads(n * n); std::vector<GridObject*> elems;
    for (auto& q : quads) elems.push_back(&q);
    NavierStokesFV1<Domain2d> d("u,v,p", "Inner");
    d.set_kinematic_viscosity(0.01);
    d.set_upwind(make_sp<NavierStokesLinearProfileSkewedUpwind<2> >());
    SmartPtr<NavierStokesFLOWStabilization<2> > st = make_sp<NavierStokesFLOWStabilization<2> >();
    st->set_diffusion_length("cor");
    d.set_stabilization(st);
    d.set_upwind(SmartPtr<INavierStokesUpwind<2> >(make_sp<NavierStokesLinearProfileSkewedUpwind<2> >()));
    d.set_exact_jacobian(true);
    d.set_source(std::vector<number>{0.2, -0.1});
    d.set_grid(NSB_QUAD, elems, nn * nn, conn.data(), xy.data());
    d.set_solution(u.data());
    IElemDisc<Domain2d>& disc = d;                                     // what ugcore's loop sees
    const int64_t ndof = d.device().num_dofs(), nnz = d.device().nnz();
    std::vector<int64_t> rowptr(ndof + 1); std::vector<int32_t> colind(nnz);
    d.device().get_csr(rowptr.data(), colind.data());
    std::map<std::pair<int64_t, int64_t>, double> G; std::vector<double> gd(ndof, 0.0);
    disc.do_prep_elem_loop(ROID_QUADRILATERAL, 0);
    for (int e = 0; e < n * n; e++) {
        LocalVector lu(nf, 4), ld(nf, 4), lr(nf, 4); LocalMatrix lJ(nf, 4);
        MathVector<2> cc[4];
        for (int s = 0; s < 4; s++) { cc[s][0] = xy[conn[e * 4 + s] * 2]; cc[s][1] = xy[conn[e * 4 + s] * 2 + 1]; }
        for (int f = 0; f < nf; f++) for (int s = 0; s < 4; s++) lu(f, s) = u[conn[e * 4 + s] * nf + f];
        disc.do_prep_elem(lu, elems[e], cc);
        disc.do_add_jac_A_elem(lJ, lu, elems[e], cc);
        disc.do_add_def_A_elem(ld, lu, elems[e], cc);
        disc.do_add_rhs_elem(lr, elems[e], cc);
        for (int rf = 0; rf < nf; rf++) for (int rs = 0; rs < 4; rs++) {
            const int64_t gr = conn[e * 4 + rs] * nf + rf;
            gd[gr] += ld(rf, rs) - lr(rf, rs);
            for (int cf = 0; cf < nf; cf++) for (int cs = 0; cs < 4; cs++) G[{gr, conn[e * 4 + cs] * nf + cf}] += lJ(rf, rs, cf, cs);
        }
    }
    disc.do_fsh_elem_loop();
    FILE* f = fopen(argc > 2 ? argv[2] : "binding_out.txt", "w");
    fprintf(f, "%d %d %lld %lld\n", n * n, nn * nn, (long long)ndof, (long long)nnz);
    for (int v : conn) fprintf(f, "%d\n", v);
    for (double v : xy) fprintf(f, "%.17g\n", v);
    for (double v : u) fprintf(f, "%.17g\n", v);
    for (int64_t r = 0; r < ndof; r++) for (int64_t q = rowptr[r]; q < rowptr[r + 1]; q++) fprintf(f, "%.17g\n", G[{r, (int64_t)colind[q]}]);
    for (double v : gd) fprintf(f, "%.17g\n", v);
    fclose(f);
    printf("OK\n");
    return 0;
}
