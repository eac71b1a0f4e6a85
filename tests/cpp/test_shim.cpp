// Exercises include/navier_stokes_b200.hpp: names / defaults / throw conditions (CPU), and with "gpu" as argv[1] the
// compat mode (IElemDisc slots inside a ugcore-like element loop) against the fast mode on a small quad grid.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include "navier_stokes_b200.hpp"

using namespace nsb200;
static int fails = 0;
#define EXPECT(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); fails++; } } while (0)
template <class F> static bool throws(F f, const char* needle)
{
    try { f(); } catch (const UGError& e) { return std::string(e.what()).find(needle) != std::string::npos; }
    return false;
}

int main(int argc, char** argv)
{
    EXPECT(upwind_id(" LPS ") == NSB_UPWIND_LPS && upwind_id("LinearProfileSkewed") == NSB_UPWIND_LPS);
    EXPECT(upwind_id("pos") == NSB_UPWIND_POSITIVE && upwind_id("Full") == NSB_UPWIND_FULL && upwind_id("no") == NSB_UPWIND_NO);
    EXPECT(throws([] { upwind_id("central"); }, "not found"));
    EXPECT(stab_id("FIELDS") == NSB_STAB_FIELDS && stab_id("flow") == NSB_STAB_FLOW);
    EXPECT(throws([] { stab_id("supg"); }, "not a valid name"));
    EXPECT(throws([] { diff_length_id("foo"); }, "Diffusion Length"));
    EXPECT(throws([] { NavierStokesFV1<2> d("u,v,w,p", "Inner"); }, "Wrong number of functions"));
    {
        NavierStokesFV1<2> d("u, v, p", "Inner");
        EXPECT(d.disc_type() == "fv1" && d.requests_local_time_series());
        EXPECT(throws([&] { d.set_pac_upwind(true); }, "Upwind must be specified previously"));
        d.set_upwind("full");
        EXPECT(throws([&] { d.set_pac_upwind(true); }, "Stabilization must be specified previously"));
        NavierStokesFVCR<3> c("u,v,w,p", "Inner");
        EXPECT(c.disc_type() == "fvcr" && c.use_hanging());
    }
    if (argc > 1 && std::string(argv[1]) == "gpu") {
        const int n = 4, nn = n + 1, nf = 3;
        std::vector<int32_t> conn; std::vector<double> xy, u;
        for (int j = 0; j < nn; j++) for (int i = 0; i < nn; i++) {
            xy.push_back(i / (double)n + 0.03 * std::sin(3.0 * i + j)); xy.push_back(j / (double)n + 0.02 * std::cos(2.0 * j + i));
            u.push_back(std::sin(1.0 + i)); u.push_back(std::cos(0.5 * j) - 0.3); u.push_back(0.1 * i * j);
        }
        for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) { int a = i + nn * j; conn.insert(conn.end(), {a, a + 1, a + 1 + nn, a + nn}); }
        NavierStokesFV1<2> d("u,v,p", "Inner");
        d.set_kinematic_viscosity(0.01);
        EXPECT(throws([&] { d.set_grid(NSB_QUAD, n * n, nn * nn, conn.data(), xy.data()); d.set_solution(u.data()); d.prep_elem_loop(ROID_QUADRILATERAL, 0); },
                      "Stabilization has not been set"));
        d.set_upwind("lps");
        d.set_stabilization("flow", "cor");
        d.set_exact_jacobian(true);
        d.set_source({0.2, -0.1});
        // fast mode
        std::vector<int64_t> rowptr(d.num_dofs() + 1); std::vector<int32_t> colind(d.nnz());
        d.get_csr(rowptr.data(), colind.data());
        std::vector<double> vals(d.nnz()), dfc(d.num_dofs());
        d.assemble_jacobian(vals.data(), u.data());
        d.assemble_defect(dfc.data(), u.data());
        // compat mode inside a ugcore-like element loop
        d.set_solution(u.data());
        d.prep_elem_loop(ROID_QUADRILATERAL, 0);
        std::map<std::pair<int64_t, int64_t>, double> G; std::vector<double> gd(d.num_dofs(), 0.0);
        for (int e = 0; e < n * n; e++) {
            LocalVector lu(nf, 4), ld(nf, 4); LocalMatrix lJ(nf, 4);
            for (int f = 0; f < nf; f++) for (int s = 0; s < 4; s++) lu(f, s) = u[conn[e * 4 + s] * nf + f];
            d.prep_elem(lu, e, ROID_QUADRILATERAL, nullptr);
            d.add_jac_A_elem(lJ, lu); d.add_def_A_elem(ld, lu);
            LocalVector lr(nf, 4); d.add_rhs_elem(lr);
            for (int rf = 0; rf < nf; rf++) for (int rs = 0; rs < 4; rs++) {
                const int64_t gr = conn[e * 4 + rs] * nf + rf;
                gd[gr] += ld(rf, rs) - lr(rf, rs);
                for (int cf = 0; cf < nf; cf++) for (int cs = 0; cs < 4; cs++) G[{gr, conn[e * 4 + cs] * nf + cf}] += lJ(rf, rs, cf, cs);
            }
        }
        d.fsh_elem_loop();
        double scale = 0, err = 0, derr = 0, dscale = 0;
        for (int64_t r = 0; r < d.num_dofs(); r++) {
            for (int64_t q = rowptr[r]; q < rowptr[r + 1]; q++) { scale = std::fmax(scale, std::fabs(vals[q])); err = std::fmax(err, std::fabs(vals[q] - G[{r, colind[q]}])); }
            dscale = std::fmax(dscale, std::fabs(dfc[r])); derr = std::fmax(derr, std::fabs(dfc[r] - gd[r]));
        }
        printf("compat vs fast: jac %.2e / %.2e   def %.2e / %.2e\n", err, scale, derr, dscale);
        EXPECT(err < 1e-12 * scale && derr < 1e-12 * dscale && scale > 0);
    }
    printf(fails ? "FAILED\n" : "OK\n");
    return fails ? 1 : 0;
}
