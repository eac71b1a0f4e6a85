// ns_bnd.cuh -- boundary element discs on the FV1 boundary faces:
//   kind 0  NavierStokesNoNormalStressOutflowFV1::add_jac_A_elem / add_def_A_elem
//           (fv1/bnd/no_normal_stress_outflow_fv1.cpp:343-427; diffusive_flux_Jac :192-236, diffusive_flux_defect :239-279,
//            convective_flux_Jac :282-313, convective_flux_defect :316-338), constant viscosity / density;
//   kind 1  the NeumannBoundaryFV1 part of NavierStokesInflowFV1 (fv1/bnd/inflow_fv1_impl.h:42-82): vector data on the
//           pressure function, rhs(p, co) -= data . n  (ugcore neumann_boundary_fv1.cpp, absent: our spec), i.e.
//           defect(p, co) += scale_a data . n.
// Boundary faces of FV1Geometry (ugcore fv1_geom.cpp, absent: our spec, restated in oracle/ns_oracle.c bf_update): one BF per
// corner of a boundary side; 2-D segment [corner, edge midpoint]; 3-D quadrilateral [corner, midpoint of the edge to the next
// side corner, side centre, midpoint of the edge to the previous side corner]; ip = mean of the BF corners, normal turned away
// from the element barycentre, shapes / global gradients of all element shape functions at the BF ip.
//
// Owner-computes, like the volume path: the host sorts the BFs by their grid node; ONE THREAD per boundary node walks its BFs in
// that fixed order and adds their rows to the node's block row (slots through the element -> CSR scatter map), so the result is
// bitwise deterministic and needs no atomics. O(surface) work.
#pragma once
#include "ns_base.h"
#include "ns_kernels.cuh"

namespace nsb {

struct BndFace { int32_t elem; int16_t side, j; int32_t data; };   // data: index of the BF's vector datum (kind 1), else -1

template <int E>
NSB_DEV void bf_corner_set(const double (*x)[ET<E>::DIM], int side, int j, double (*c)[ET<E>::DIM], int& nc)
{
    constexpr int DIM = ET<E>::DIM;
    const int ns = t_side_n<E>(side), co = t_side<E>(side, j);
#pragma unroll
    for (int d = 0; d < DIM; d++) c[0][d] = x[co][d];
    if constexpr (DIM == 2) {
        const int other = t_side<E>(side, 1 - j);
#pragma unroll
        for (int d = 0; d < DIM; d++) c[1][d] = 0.5 * (x[co][d] + x[other][d]);
        nc = 2;
    } else {
        const int nx = t_side<E>(side, (j + 1) % ns), pv = t_side<E>(side, (j + ns - 1) % ns);
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            c[1][d] = 0.5 * (x[co][d] + x[nx][d]); c[3][d] = 0.5 * (x[co][d] + x[pv][d]);
            double s = 0.0;
            for (int k = 0; k < ns; k++) s += x[t_side<E>(side, k)][d];
            c[2][d] = s / ns;
        }
        nc = 4;
    }
}

// outward, area-scaled normal and local ip of boundary face j of a side (x: global element corners)
template <int E>
NSB_DEV void bf_normal_lip(const double (*x)[ET<E>::DIM], int side, int j, double* n, double* lip)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    double xr[NSH][DIM], c[4][DIM], lc[4][DIM];
#pragma unroll
    for (int k = 0; k < NSH; k++)
#pragma unroll
        for (int d = 0; d < DIM; d++) xr[k][d] = t_corner<E>(k, d);
    int nc;
    bf_corner_set<E>(x, side, j, c, nc);
    bf_corner_set<E>(xr, side, j, lc, nc);
#pragma unroll
    for (int d = 0; d < DIM; d++) { double t = 0.0; for (int k = 0; k < nc; k++) t += lc[k][d]; lip[d] = t / nc; }
    if constexpr (DIM == 2) { n[0] = c[1][1] - c[0][1]; n[1] = -(c[1][0] - c[0][0]); }
    else {
        double av[3], bv[3];
#pragma unroll
        for (int d = 0; d < 3; d++) { av[d] = c[2][d] - c[0][d]; bv[d] = c[3][d] - c[1][d]; }
        cross3(n, av, bv);
#pragma unroll
        for (int d = 0; d < 3; d++) n[d] *= 0.5;
    }
    // outward: away from the element barycentre
    const int ns = t_side_n<E>(side);
    double o = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        double bary = 0.0, sc = 0.0;
        for (int k = 0; k < NSH; k++) bary += x[k][d];
        for (int k = 0; k < ns; k++) sc += x[t_side<E>(side, k)][d];
        o += n[d] * (sc / ns - bary / NSH);
    }
    if (o < 0) {
#pragma unroll
        for (int d = 0; d < DIM; d++) n[d] = -n[d];
    }
}

template <int E>
__global__ void __launch_bounds__(64) fv1_boundary_kernel(KParams p, MeshDev m, int kind, int64_t n_bnode, const int32_t* __restrict__ bnode,
                                                          const int64_t* __restrict__ bptr, const BndFace* __restrict__ bf,
                                                          const double* __restrict__ data, const double* __restrict__ u,
                                                          double* __restrict__ val, double* __restrict__ def, int* __restrict__ errflag)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1, P = DIM;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bnode) return;
    const int64_t a = bnode[i];
    const int64_t b0 = m.brow[a]; const int cnt = (int)(m.brow[a + 1] - b0);
    double* rows = val ? val + b0 * (NF * NF) : nullptr;                 // scalar row rf: rows + rf * cnt * NF, entry (slot, cf)
    const bool want_jac = (p.what & W_JAC_A) && val, want_def = (p.what & (W_DEF_A | W_RHS)) && def;
    const double nurho = p.visc * p.rho, sa = p.scale_a;
    double dacc[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) dacc[f] = 0.0;
    for (int64_t q = bptr[i]; q < bptr[i + 1]; q++) {
        const BndFace f = bf[q];
        const int32_t* nd = m.conn + (int64_t)f.elem * NSH;
        double x[NSH][DIM], ul[NSH][NF];
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            const int64_t g = nd[k];
#pragma unroll
            for (int d = 0; d < DIM; d++) x[k][d] = m.coords[g * DIM + d];
#pragma unroll
            for (int c = 0; c < NF; c++) ul[k][c] = u ? u[g * NF + c] : 0.0;
        }
        double n[DIM], lip[DIM];
        bf_normal_lip<E>(x, f.side, f.j, n, lip);
        if (kind == 1) {
            if (want_def) {
                double s = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) s += data[(int64_t)f.data * DIM + d] * n[d];
                dacc[P] += sa * s;
            }
            continue;
        }
        double N[NSH], dN[NSH][DIM], JT[DIM][DIM], JI[DIM][DIM], G[NSH][DIM];
        lagrange<E>(lip, N);
        lagrange_grad<E>(lip, dN);
#pragma unroll
        for (int r = 0; r < DIM; r++)
#pragma unroll
            for (int s = 0; s < DIM; s++) { double t = 0.0; for (int k = 0; k < NSH; k++) t += dN[k][r] * x[k][s]; JT[r][s] = t; }
        const double det = inv_mat<DIM>(JT, JI);
        if (!(fabs(det) > 0.0)) { atomicExch(errflag, 1); continue; }
#pragma unroll
        for (int k = 0; k < NSH; k++)
#pragma unroll
            for (int s = 0; s < DIM; s++) { double t = 0.0; for (int r = 0; r < DIM; r++) t += JI[s][r] * dN[k][r]; G[k][s] = t; }
        double sv[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) { double t = 0.0; for (int k = 0; k < NSH; k++) t += ul[k][d] * N[k]; sv[d] = t; }
        const double svn = dotv<DIM>(sv, n);
        double flux = svn * p.rho;                                     // no inflow through the outflow boundary (:299, :329)
        if (flux < 0) flux = 0.0;
        const int co = t_side<E>(f.side, f.j);
        if (want_jac) {
            const uint8_t* em = m.emap + ((int64_t)f.elem * NSH + co) * NSH;
            for (int sh = 0; sh < NSH; sh++) {
                const int slot = em[sh];
                const double gn = dotv<DIM>(G[sh], n);
                double T[DIM][DIM], nst[DIM];
#pragma unroll
                for (int d1 = 0; d1 < DIM; d1++)
#pragma unroll
                    for (int d2 = 0; d2 < DIM; d2++) { T[d1][d2] = d1 == d2 ? gn : 0.0; if (!p.laplace) T[d1][d2] += G[sh][d1] * n[d2]; }
#pragma unroll
                for (int d2 = 0; d2 < DIM; d2++) { double t = 0.0; for (int d1 = 0; d1 < DIM; d1++) t += T[d1][d2] * n[d1]; nst[d2] = t; }
#pragma unroll
                for (int d1 = 0; d1 < DIM; d1++)
#pragma unroll
                    for (int d2 = 0; d2 < DIM; d2++) {
                        double v = (T[d1][d2] - n[d1] * nst[d2]) * (-nurho);
                        if (d1 == d2 && !p.stokes) v += flux * N[sh];
                        rows[(int64_t)d1 * cnt * NF + slot * NF + d2] += sa * v;
                    }
#pragma unroll
                for (int d2 = 0; d2 < DIM; d2++) rows[(int64_t)P * cnt * NF + slot * NF + d2] += sa * (N[sh] * n[d2] * p.rho);
            }
        }
        if (want_def && (p.what & W_DEF_A)) {
            double gv[DIM][DIM], df[DIM];
#pragma unroll
            for (int d1 = 0; d1 < DIM; d1++)
#pragma unroll
                for (int d2 = 0; d2 < DIM; d2++) { double t = 0.0; for (int sh = 0; sh < NSH; sh++) t += G[sh][d2] * ul[sh][d1]; gv[d1][d2] = t; }
#pragma unroll
            for (int d1 = 0; d1 < DIM; d1++) {
                double t = 0.0;
                for (int d2 = 0; d2 < DIM; d2++) t += gv[d1][d2] * n[d2];
                if (!p.laplace) for (int d2 = 0; d2 < DIM; d2++) t += gv[d2][d1] * n[d2];
                df[d1] = t;
            }
            const double dn = dotv<DIM>(df, n);
#pragma unroll
            for (int d1 = 0; d1 < DIM; d1++) {
                double v = (df[d1] - dn * n[d1]) * (-nurho);          // VecScaleAppend(diffFlux, -dot, normal) :270 (normal not normalised)
                if (!p.stokes) v += flux * sv[d1];
                dacc[d1] += sa * v;
            }
            dacc[P] += sa * (svn * p.rho);
        }
    }
    if (want_def) {
#pragma unroll
        for (int f = 0; f < NF; f++) def[a * NF + f] += dacc[f];
    }
}

}  // namespace nsb
