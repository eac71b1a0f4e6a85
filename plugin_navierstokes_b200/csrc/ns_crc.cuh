// ns_crc.cuh -- SURVEY 8f-3: DiscConstraintFVCR in its default configuration (fvcr/disc_constraint_fvcr.h:254-300:
// bLinUpConvDefect = true, bLinPressureDefect = true, not adaptive, no limiter), the post-assembly correction of the FVCR DEFECT
// (adjust_defect :1149-1171 -> add_defect :770-1147):
//   side gradients (:780-870)     acGrad(side) = sum_elem vol_scv [sum_sh u_sh,d0 grad_sh,d1] / sum_elem vol_scv
//   per element and SCVF (:1022-1144), elements with a side in a zero-gradient subset skipped (:321-327):
//     flux = s_a StdVel . n, base = flux > 0 ? from : to
//     linear upwind   : upwindVel_d1 = acGrad(base)_d1 . (x_ip - x_scv(base));   d(d1, from) += upwindVel flux,  d(d1, to) -= ...
//     linear pressure : pGrad = 1/|elem| sum_sides n_side (boundary ? p_e : (p_e + p_nb) / 2),
//                       pressure = s_a pGrad . (x_ip - barycentre);              d(d1, from) += pressure n_d1,    d(d1, to) -= ...
// The reference loops over elements and scatters into the side dofs; here both passes are OWNER-COMPUTES over the
// side -> (element, local side) adjacency: one thread per side gathers the contributions of its (at most two) elements in a fixed
// order -- no atomics, no colouring, bitwise deterministic. Simplices (the element types of the FVCR device path).
#pragma once
#include "ns_base.h"
#include "ns_fvcr.cuh"

namespace nsb {

// element-level CR geometry shared by both passes
template <int E> struct CRElemGeo {
    static constexpr int DIM = CRT<E>::DIM, NCO = CRT<E>::NCO, NS = CRT<E>::NS;
    double x[NCO][DIM], bary[DIM], G[NS][DIM], scvn[NS][DIM], scvx[NS][DIM], vol;   // vol = SCV volume = |elem| / NS
};

template <int E> NSB_DEV bool cr_elem_geo(const FvcrDev& f, int64_t e, CRElemGeo<E>& g)
{
    constexpr int DIM = CRT<E>::DIM, NCO = CRT<E>::NCO, NS = CRT<E>::NS;
    for (int k = 0; k < NCO; k++) {
        const int64_t nd = f.conn[e * NCO + k];
#pragma unroll
        for (int d = 0; d < DIM; d++) g.x[k][d] = f.coords[nd * DIM + d];
    }
    double JT[DIM][DIM], JI[DIM][DIM], dl[NCO][DIM], xi0[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) xi0[d] = 0.0;
    lagrange_grad<E>(xi0, dl);
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) { double s = 0; for (int k = 0; k < NCO; k++) s += dl[k][i] * g.x[k][j]; JT[i][j] = s; }
    const double det = inv_mat<DIM>(JT, JI);
    if (!(fabs(det) > 0.0)) return false;
    g.vol = fabs(det) / (DIM == 2 ? 2.0 : 6.0) / NS;
#pragma unroll
    for (int d = 0; d < DIM; d++) { double s = 0; for (int k = 0; k < NCO; k++) s += g.x[k][d]; g.bary[d] = s / NCO; }
    for (int s = 0; s < NS; s++) {
        const int o = tab::CR_OPP[E][s];
#pragma unroll
        for (int j = 0; j < DIM; j++) { double t = 0; for (int i = 0; i < DIM; i++) t += JI[j][i] * dl[o][i]; g.G[s][j] = -1.0 * DIM * t; }
        double xb[DIM], nn[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) { double t = 0.0; for (int q = 0; q < DIM; q++) t += g.x[tab::SIDE[E][s][q]][d]; xb[d] = t / DIM; }
        if constexpr (DIM == 2) {
            const double* a = g.x[tab::SIDE[E][s][0]]; const double* b = g.x[tab::SIDE[E][s][1]];
            nn[0] = b[1] - a[1]; nn[1] = -(b[0] - a[0]);
        } else {
            double e1[3], e2[3], c[3];
            for (int d = 0; d < 3; d++) { e1[d] = g.x[tab::SIDE[E][s][1]][d] - g.x[tab::SIDE[E][s][0]][d]; e2[d] = g.x[tab::SIDE[E][s][2]][d] - g.x[tab::SIDE[E][s][0]][d]; }
            cross3(c, e1, e2);
            for (int d = 0; d < 3; d++) nn[d] = 0.5 * c[d];
        }
        double outw = 0;
#pragma unroll
        for (int d = 0; d < DIM; d++) outw += nn[d] * (xb[d] - g.bary[d]);
        const double sg = outw < 0 ? -1.0 : 1.0;
#pragma unroll
        for (int d = 0; d < DIM; d++) { g.scvn[s][d] = sg * nn[d]; g.scvx[s][d] = xb[d]; }
    }
    return true;
}

// pass 1: acGrad [n_side][DIM][DIM]. sadj: (element * NS + local side) entries of the side -> element adjacency
template <int E>
__global__ void __launch_bounds__(128) fvcr_side_grad_kernel(FvcrDev f, const int32_t* __restrict__ sadj, const double* __restrict__ u,
                                                             double* __restrict__ grad, int* __restrict__ errflag)
{
    constexpr int DIM = CRT<E>::DIM, NS = CRT<E>::NS;
    const int64_t sd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sd >= f.n_side) return;
    double acc[DIM][DIM], vol = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) acc[i][j] = 0.0;
    for (int64_t q = f.sadj_ptr[sd]; q < f.sadj_ptr[sd + 1]; q++) {
        const int64_t e = sadj[q] / NS;
        CRElemGeo<E> g;
        if (!cr_elem_geo<E>(f, e, g)) { atomicExch(errflag, 2); continue; }
        for (int sh = 0; sh < NS; sh++) {
            const int64_t s2 = f.esides[e * NS + sh];
#pragma unroll
            for (int d0 = 0; d0 < DIM; d0++) {
                const double uv = u[s2 * DIM + d0];
#pragma unroll
                for (int d1 = 0; d1 < DIM; d1++) acc[d0][d1] += uv * g.G[sh][d1] * g.vol;
            }
        }
        vol += g.vol;
    }
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) grad[sd * (DIM * DIM) + i * DIM + j] = vol > 0.0 ? acc[i][j] / vol : 0.0;
}

// pass 2: defect(side, d1) += the corrections of the SCVFs of the adjacent elements that start or end at this side
template <int E>
__global__ void __launch_bounds__(128) fvcr_constraint_defect_kernel(FvcrDev f, const int32_t* __restrict__ sadj, const double* __restrict__ u,
                                                                     const double* __restrict__ grad, const uint8_t* __restrict__ zflag,
                                                                     double s_a, int lin_up, int lin_p, double* __restrict__ def)
{
    constexpr int DIM = CRT<E>::DIM, NCO = CRT<E>::NCO, NS = CRT<E>::NS, NIP = CRT<E>::NIP;
    const int64_t sd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sd >= f.n_side) return;
    const int64_t pbase = f.n_side * DIM;
    double dacc[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) dacc[d] = 0.0;
    for (int64_t q = f.sadj_ptr[sd]; q < f.sadj_ptr[sd + 1]; q++) {
        const int64_t e = sadj[q] / NS; const int a = (int)(sadj[q] - e * NS);
        int64_t sides[NS];
        bool skip = false;
        for (int s = 0; s < NS; s++) { sides[s] = f.esides[e * NS + s]; if (zflag && zflag[sides[s]]) skip = true; }
        if (skip) continue;                                      // zeroGradBndElem (:321-327, :1022-1024)
        CRElemGeo<E> g;
        if (!cr_elem_geo<E>(f, e, g)) continue;
        double ul[NS][DIM], pg[DIM];
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int d = 0; d < DIM; d++) ul[s][d] = u[sides[s] * DIM + d];
#pragma unroll
        for (int d = 0; d < DIM; d++) pg[d] = 0.0;
        if (lin_p) {
            const double pe = u[pbase + e];
            for (int s = 0; s < NS; s++) {
                const int64_t q0 = f.sadj_ptr[sides[s]], q1 = f.sadj_ptr[sides[s] + 1];
                double pv = pe;                                   // boundary side (:1080-1082)
                if (q1 - q0 > 1) { const int64_t e0 = sadj[q0] / NS, e1 = sadj[q0 + 1] / NS; pv = 0.5 * (pe + u[pbase + (e0 == e ? e1 : e0)]); }
#pragma unroll
                for (int d = 0; d < DIM; d++) pg[d] += g.scvn[s][d] * pv;
            }
#pragma unroll
            for (int d = 0; d < DIM; d++) pg[d] /= (g.vol * NS);
        }
        double lb[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) { double s = 0; for (int k = 0; k < NCO; k++) s += tab::CORNER[E][k][d]; lb[d] = s / NCO; }
        for (int ip = 0; ip < NIP; ip++) {
            const int from = tab::CR_FROM[E][ip], to = tab::CR_TO[E][ip];
            if (from != a && to != a) continue;
            double n[DIM], xip[DIM], lip[DIM], N[NS];
            if constexpr (DIM == 2) {
                const double* c0 = g.x[ip];
                n[0] = g.bary[1] - c0[1]; n[1] = -(g.bary[0] - c0[0]);
                for (int d = 0; d < 2; d++) { xip[d] = 0.5 * (c0[d] + g.bary[d]); lip[d] = 0.5 * (tab::CORNER[E][ip][d] + lb[d]); }
            } else {
                const int c0 = tab::EDGE[E][ip][0], c1 = tab::EDGE[E][ip][1];
                double e1[3], e2[3], c[3];
                for (int d = 0; d < 3; d++) { e1[d] = g.x[c1][d] - g.x[c0][d]; e2[d] = g.bary[d] - g.x[c0][d]; }
                cross3(c, e1, e2);
                for (int d = 0; d < 3; d++) {
                    n[d] = 0.5 * c[d];
                    xip[d] = (g.x[c0][d] + g.x[c1][d] + g.bary[d]) / 3.0;
                    lip[d] = (tab::CORNER[E][c0][d] + tab::CORNER[E][c1][d] + lb[d]) / 3.0;
                }
            }
            double ft = 0;
#pragma unroll
            for (int d = 0; d < DIM; d++) ft += n[d] * (g.scvx[to][d] - g.scvx[from][d]);
            if (ft < 0) {
#pragma unroll
                for (int d = 0; d < DIM; d++) n[d] = -n[d];
            }
            cr_shapes<E>(lip, N);
            double sv[DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) { double s = 0; for (int k = 0; k < NS; k++) s += ul[k][d] * N[k]; sv[d] = s; }
            const double flux = s_a * dotv<DIM>(sv, n);
            const int base = flux > 0 ? from : to;
            const double sgn = from == a ? 1.0 : -1.0;
            double pressure = 0.0;
            if (lin_p) { double t = 0.0; for (int j = 0; j < DIM; j++) t += pg[j] * (xip[j] - g.bary[j]); pressure = s_a * t; }
            const double* gb = grad + sides[base] * (DIM * DIM);
#pragma unroll
            for (int d1 = 0; d1 < DIM; d1++) {
                double v = 0.0;
                if (lin_up) { double uv = 0.0; for (int d2 = 0; d2 < DIM; d2++) uv += gb[d1 * DIM + d2] * (xip[d2] - g.scvx[base][d2]); v += uv * flux; }
                if (lin_p) v += pressure * n[d1];
                dacc[d1] += sgn * v;
            }
        }
    }
#pragma unroll
    for (int d = 0; d < DIM; d++) def[sd * DIM + d] += dacc[d];
}

}  // namespace nsb
