// ns_fv1.cuh -- device-side building blocks of the FV1 (vertex-centred, Schneider-Raw stabilised)
// incompressible Navier-Stokes element assembly for sm_100a.
//
// What is computed follows the UG4 NavierStokes plugin (file:line cited per function, relative to the
// reference tree); how it is computed is organised for the GPU: every quantity that belongs to one
// sub-control-volume face (SCVF, one integration point "ip") is produced by ONE lane in registers
// (`ip_eval`), parked in shared memory as an `IpRec`, and consumed by lanes that each own one COLUMN
// (corner k, function cf) of the local Jacobian (`jac_col`), so a flux derivative is evaluated once and
// added to row `from` / subtracted from row `to` without cross-lane traffic.
#pragma once
#include <utility>
#include "ns_base.h"

namespace nsb {

// ------------------------------------------------------------------------------------------------
// FV1Geometry of one SCVF (ugcore FV1Geometry::update, called from prep_elem,
// fv1/navier_stokes_fv1.cpp:208-248; conventions SURVEY.md App. B-2)
// ------------------------------------------------------------------------------------------------
template <int E> struct IpGeo {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    double n[DIM], xip[DIM], N[NSH], G[NSH][DIM], ds;
    int from, to;
};

// x: element corner coordinates [NSH][DIM] (shared memory)
template <int E> NSB_DEV void ip_geometry(const double* __restrict__ x, int ip, IpGeo<E>& g)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    const int f = t_edge<E>(ip, 0), t = t_edge<E>(ip, 1);
    g.from = f; g.to = t;
    double cen[DIM], c0[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < NSH; k++) s += x[k * DIM + d];
        cen[d] = s * (1.0 / NSH);
        c0[d] = 0.5 * (x[f * DIM + d] + x[t * DIM + d]);
    }
    if constexpr (DIM == 2) {
        // SCVF corners [edge midpoint, barycentre]; n = (dy, -dx) of c1 - c0
        g.n[0] = cen[1] - c0[1]; g.n[1] = -(cen[0] - c0[0]);
        g.xip[0] = 0.5 * (c0[0] + cen[0]); g.xip[1] = 0.5 * (c0[1] + cen[1]);
        g.ds = 0.0;
    } else {
        // SCVF corners [edge midpoint, centre of face A, barycentre, centre of face B]
        const int fa = t_fa<E>(ip), fb = t_fb<E>(ip);
        constexpr int NFC = (E == E_TET) ? 3 : 4;
        double c1[3] = {0, 0, 0}, c3[3] = {0, 0, 0};
        double w1 = 1.0 / NFC, w3 = 1.0 / NFC;
        if constexpr (E == E_PRISM) {
            // mixed sides: triangle (3 corners) or quadrilateral (4)
            const int na = t_side_n<E>(fa), nb = t_side_n<E>(fb);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (q < na) { const int ka = t_side<E>(fa, q);
#pragma unroll
                    for (int d = 0; d < 3; d++) c1[d] += x[ka * 3 + d]; }
                if (q < nb) { const int kb = t_side<E>(fb, q);
#pragma unroll
                    for (int d = 0; d < 3; d++) c3[d] += x[kb * 3 + d]; }
            }
            w1 = 1.0 / na; w3 = 1.0 / nb;
        } else {
#pragma unroll
        for (int q = 0; q < NFC; q++) {
            const int ka = tab::SIDE[E][fa][q], kb = tab::SIDE[E][fb][q];
#pragma unroll
            for (int d = 0; d < 3; d++) { c1[d] += x[ka * 3 + d]; c3[d] += x[kb * 3 + d]; }
        }
        }
        double a[3], b[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            if constexpr (E == E_PRISM) { c1[d] *= w1; c3[d] *= w3; } else { c1[d] *= (1.0 / NFC); c3[d] *= (1.0 / NFC); }
            a[d] = cen[d] - c0[d]; b[d] = c3[d] - c1[d];
            g.xip[d] = 0.25 * (c0[d] + c1[d] + cen[d] + c3[d]);
        }
        cross3(g.n, a, b);
#pragma unroll
        for (int d = 0; d < 3; d++) g.n[d] *= 0.5;
        g.ds = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];      // |corner(0) - corner(2)|^2
    }
    // shapes and global gradients at the local ip
    double xi[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) xi[d] = t_lip<E>(ip, d);
    lagrange<E>(xi, g.N);
    double dN[NSH][DIM];
    lagrange_grad<E>(xi, dN);
    double JT[DIM][DIM], JI[DIM][DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < NSH; k++) s += dN[k][i] * x[k * DIM + j];
            JT[i][j] = s;
        }
    inv_mat<DIM>(JT, JI);
#pragma unroll
    for (int k = 0; k < NSH; k++)
#pragma unroll
        for (int j = 0; j < DIM; j++) {
            double s = 0;
#pragma unroll
            for (int i = 0; i < DIM; i++) s += JI[j][i] * dN[k][i];
            g.G[k][j] = s;
        }
}

// SCV volume of corner `co` (ugcore FV1Geometry SCV::volume; App. B-2). Simplices: |T|/(dim+1).
template <int E> NSB_DEV double scv_volume(const double* __restrict__ x, int co)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    if constexpr (E == E_TRI) {
        const double a = (x[2] - x[0]) * (x[5] - x[1]) - (x[4] - x[0]) * (x[3] - x[1]);
        return fabs(a) * (0.5 / 3.0);
    } else if constexpr (E == E_TET) {
        double a[3], b[3], c[3], t[3];
#pragma unroll
        for (int d = 0; d < 3; d++) { a[d] = x[3 + d] - x[d]; b[d] = x[6 + d] - x[d]; c[d] = x[9 + d] - x[d]; }
        cross3(t, a, b);
        return fabs(dotv<3>(t, c)) * (1.0 / 24.0);
    } else if constexpr (E == E_QUAD) {
        const int nx = (co + 1) & 3, pv = (co + 3) & 3;
        double bc[2], m1[2], m2[2];
#pragma unroll
        for (int d = 0; d < 2; d++) {
            bc[d] = 0.25 * (x[d] + x[2 + d] + x[4 + d] + x[6 + d]);
            m1[d] = 0.5 * (x[co * 2 + d] + x[nx * 2 + d]);
            m2[d] = 0.5 * (x[pv * 2 + d] + x[co * 2 + d]);
        }
        const double ax = bc[0] - x[co * 2], ay = bc[1] - x[co * 2 + 1], bx = m2[0] - m1[0], by = m2[1] - m1[1];
        return 0.5 * fabs(ax * by - ay * bx);
    } else if constexpr (E == E_PRISM) {
        // hexahedron (corner, edge midpoint, triangle centre, edge midpoint | axis-edge midpoint, quadrilateral centre,
        // barycentre, quadrilateral centre): volume of the trilinear hexahedron through these eight points
        const int base = co < 3 ? 0 : 3, t = co - base, ca = base + (t + 1) % 3, cb = base + (t + 2) % 3, up = co < 3 ? co + 3 : co - 3;
        const int oa = ca < 3 ? ca + 3 : ca - 3, ob = cb < 3 ? cb + 3 : cb - 3;
        double p[8][3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double xc = x[co * 3 + d], xa = x[ca * 3 + d], xb = x[cb * 3 + d], xu = x[up * 3 + d], xoa = x[oa * 3 + d], xob = x[ob * 3 + d];
            p[0][d] = xc;
            p[1][d] = (xc + xa) / 2;
            p[2][d] = (xc + xa + xb) / 3;
            p[3][d] = (xc + xb) / 2;
            p[4][d] = (xc + xu) / 2;
            p[5][d] = (xc + xa + xoa + xu) / 4;
            p[6][d] = (x[d] + x[3 + d] + x[6 + d] + x[9 + d] + x[12 + d] + x[15 + d]) / 6;
            p[7][d] = (xc + xb + xob + xu) / 4;
        }
        double a[3], b[3], c[3], tt[3], v = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) { a[d] = (p[6][d] - p[1][d]) + (p[7][d] - p[0][d]); b[d] = p[6][d] - p[3][d]; c[d] = p[2][d] - p[0][d]; }
        cross3(tt, b, c); v += dotv<3>(a, tt);
#pragma unroll
        for (int d = 0; d < 3; d++) { a[d] = p[7][d] - p[0][d]; b[d] = (p[6][d] - p[3][d]) + (p[5][d] - p[0][d]); c[d] = p[6][d] - p[4][d]; }
        cross3(tt, b, c); v += dotv<3>(a, tt);
#pragma unroll
        for (int d = 0; d < 3; d++) { a[d] = p[6][d] - p[1][d]; b[d] = p[5][d] - p[0][d]; c[d] = (p[6][d] - p[4][d]) + (p[2][d] - p[0][d]); }
        cross3(tt, b, c); v += dotv<3>(a, tt);
        return fabs(v) * (1.0 / 12.0);
    } else {
        // trilinear image of the reference octant adjacent to corner `co`; exact volume by the
        // long-diagonal formula
        const int cx = (co & 1) ^ ((co >> 1) & 1), cy = (co >> 1) & 1, cz = (co >> 2) & 1;
        const double lo[3] = {cx ? 0.5 : 0.0, cy ? 0.5 : 0.0, cz ? 0.5 : 0.0};
        double p[8][3];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int qx = (q & 1) ^ ((q >> 1) & 1), qy = (q >> 1) & 1, qz = (q >> 2) & 1;
            const double xi[3] = {lo[0] + 0.5 * qx, lo[1] + 0.5 * qy, lo[2] + 0.5 * qz};
            double N[8];
            lagrange<E_HEX>(xi, N);
#pragma unroll
            for (int d = 0; d < 3; d++) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) s += N[k] * x[k * 3 + d];
                p[q][d] = s;
            }
        }
        double a[3], b[3], c[3], t[3], v = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) { a[d] = (p[6][d] - p[1][d]) + (p[7][d] - p[0][d]); b[d] = p[6][d] - p[3][d]; c[d] = p[2][d] - p[0][d]; }
        cross3(t, b, c); v += dotv<3>(a, t);
#pragma unroll
        for (int d = 0; d < 3; d++) { a[d] = p[7][d] - p[0][d]; b[d] = (p[6][d] - p[3][d]) + (p[5][d] - p[0][d]); c[d] = p[6][d] - p[4][d]; }
        cross3(t, b, c); v += dotv<3>(a, t);
#pragma unroll
        for (int d = 0; d < 3; d++) { a[d] = p[6][d] - p[1][d]; b[d] = p[5][d] - p[0][d]; c[d] = (p[6][d] - p[4][d]) + (p[2][d] - p[0][d]); }
        cross3(t, b, c); v += dotv<3>(a, t);
        return fabs(v) * (1.0 / 12.0);
    }
    (void)DIM; (void)NSH;
}

// ------------------------------------------------------------------------------------------------
// ElementSideRayIntersection (ugcore geometry_util.h): sides in reference order,
// quadrilateral sides as triangles (p0,p1,p2),(p0,p2,p3); first hit with t<=0 (upwind search) wins.
// Call sites: upwind.cpp:351,547.
// ------------------------------------------------------------------------------------------------
#define NSB_RAY_SMALL 1e-12

template <int N, class F, int... I> NSB_DEV void static_for_impl(F&& f, std::integer_sequence<int, I...>)
{ (f(std::integral_constant<int, I>{}), ...); }
template <int N, class F> NSB_DEV void static_for(F&& f) { static_for_impl<N>(f, std::make_integer_sequence<int, N>{}); }

// compile-time reference-element tables (fold into immediates once loops are unrolled)
template <int E> __host__ __device__ constexpr int edge_corner(int ip, int j)
{
    if (E == E_TRI)  { constexpr int T[3][2]  = {{0,1},{1,2},{2,0}}; return T[ip][j]; }
    if (E == E_QUAD) { constexpr int T[4][2]  = {{0,1},{1,2},{2,3},{3,0}}; return T[ip][j]; }
    if (E == E_TET)  { constexpr int T[6][2]  = {{0,1},{1,2},{2,0},{0,3},{1,3},{2,3}}; return T[ip][j]; }
    if (E == E_PRISM) { constexpr int T[9][2] = {{0,1},{1,2},{2,0},{0,3},{1,4},{2,5},{3,4},{4,5},{5,3}}; return T[ip][j]; }
    constexpr int T[12][2] = {{0,1},{1,2},{2,3},{3,0},{0,4},{1,5},{2,6},{3,7},{4,5},{5,6},{6,7},{7,4}};
    return T[ip][j];
}
template <int E> __host__ __device__ constexpr int side_corner(int s, int i)
{
    if (E == E_TRI)  { constexpr int T[3][2] = {{0,1},{1,2},{2,0}}; return T[s][i]; }
    if (E == E_QUAD) { constexpr int T[4][2] = {{0,1},{1,2},{2,3},{3,0}}; return T[s][i]; }
    if (E == E_TET)  { constexpr int T[4][3] = {{0,2,1},{1,2,3},{0,3,2},{0,1,3}}; return T[s][i]; }
    if (E == E_PRISM) { constexpr int T[5][4] = {{0,2,1,0},{0,1,4,3},{1,2,5,4},{2,0,3,5},{3,4,5,3}}; return T[s][i]; }
    constexpr int T[6][4] = {{0,3,2,1},{0,1,5,4},{1,2,6,5},{2,3,7,6},{3,0,4,7},{4,5,6,7}};
    return T[s][i];
}
// corners of a side: only the prism mixes triangles and quadrilaterals
template <int E> __host__ __device__ constexpr int side_ncorner(int s)
{
    if (E == E_PRISM) return (s == 0 || s == 4) ? 3 : 4;
    return ET<E>::DIM == 2 ? 2 : (E == E_TET ? 3 : 4);
}

// All candidate segments / triangles are tested in reference order with compile-time corner indices; the
// inside / upstream tests are done on the un-divided Cramer numerators (sign-corrected by det), so the
// only divisions are the three of the winning triangle.
template <int E> NSB_DEV bool side_ray_cut(const double* __restrict__ x, const double* from, const double* dir,
                                           int& side_out, double* gcut, double* lcut)
{
    constexpr int DIM = ET<E>::DIM, NSIDE = ET<E>::NSIDE;
    constexpr double S = NSB_RAY_SMALL;
    bool found = false;
    int best = 0;
    double tn = 0.0, n1 = 0.0, n2 = 0.0, bdet = 1.0;
    if constexpr (DIM == 2) {
        const double dn2 = dir[0] * dir[0] + dir[1] * dir[1];
        static_for<NSIDE>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            constexpr int p0 = side_corner<E>(s, 0), p1 = side_corner<E>(s, 1);
            if (!found) {
                const double ex = x[p1 * 2] - x[p0 * 2], ey = x[p1 * 2 + 1] - x[p0 * 2 + 1];
                const double det = dir[0] * (-ey) + dir[1] * ex;
                if (det * det > (S * S) * dn2 * (ex * ex + ey * ey)) {
                    const double rx = x[p0 * 2] - from[0], ry = x[p0 * 2 + 1] - from[1];
                    const double t_n = rx * (-ey) + ry * ex, b_n = dir[0] * ry - dir[1] * rx;
                    const double sg = det > 0.0 ? 1.0 : -1.0, ad = fabs(det);
                    if (b_n * sg >= -S * ad && b_n * sg <= (1.0 + S) * ad && t_n * sg <= 0.0) {
                        found = true; best = s; tn = t_n; n1 = b_n; bdet = det;
                    }
                }
            }
        });
        if (!found) return false;
        const double t = tn / bdet, bc = n1 / bdet;
        const int p0 = t_side<E>(best, 0), p1 = t_side<E>(best, 1);
#pragma unroll
        for (int d = 0; d < 2; d++) {
            gcut[d] = from[d] + t * dir[d];
            lcut[d] = (1 - bc) * t_corner<E>(p0, d) + bc * t_corner<E>(p1, d);
        }
        side_out = best;
        return true;
    } else {
        const double dn2 = dotv<3>(dir, dir);
        constexpr int TPS = (E == E_HEX || E == E_PRISM) ? 2 : 1;            // triangles per side
        static_for<NSIDE * TPS>([&](auto ic) {
            constexpr int i = decltype(ic)::value, s = i / TPS, kk = i % TPS;
            constexpr int p0 = side_corner<E>(s, 0), p1 = side_corner<E>(s, 1 + kk), p2 = side_corner<E>(s, (2 + kk) & 3);
            if (kk + 3 <= side_ncorner<E>(s) && !found) {
                double e1[3], e2[3], r[3], nrm[3], q[3];
#pragma unroll
                for (int d = 0; d < 3; d++) { e1[d] = x[p1 * 3 + d] - x[p0 * 3 + d]; e2[d] = x[p2 * 3 + d] - x[p0 * 3 + d]; r[d] = from[d] - x[p0 * 3 + d]; }
                cross3(nrm, e1, e2);
                const double det = -dotv<3>(dir, nrm);
                if (det * det > (S * S) * dn2 * dotv<3>(nrm, nrm)) {
                    const double t_n = dotv<3>(r, nrm);
                    cross3(q, r, dir);
                    const double b1n = dotv<3>(e2, q), b2n = -dotv<3>(e1, q);
                    const double sg = det > 0.0 ? 1.0 : -1.0, ad = fabs(det);
                    if (b1n * sg >= -S * ad && b2n * sg >= -S * ad && (b1n + b2n) * sg <= (1.0 + S) * ad && t_n * sg <= 0.0) {
                        found = true; best = i; tn = t_n; n1 = b1n; n2 = b2n; bdet = det;
                    }
                }
            }
        });
        if (!found) return false;
        const double t = tn / bdet, b1 = n1 / bdet, b2 = n2 / bdet;
        const int s = best / TPS, kk = best - s * TPS;
        const int p0 = t_side<E>(s, 0), p1 = t_side<E>(s, 1 + kk), p2 = t_side<E>(s, 2 + kk);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            gcut[d] = from[d] + t * dir[d];
            lcut[d] = (1 - b1 - b2) * t_corner<E>(p0, d) + b1 * t_corner<E>(p1, d) + b2 * t_corner<E>(p2, d);
        }
        side_out = s;
        return true;
    }
}

// ------------------------------------------------------------------------------------------------
// Upwind shapes of ONE ip for the upwinds without ip-shapes (upwind.cpp:52-80 No, :133-172 Full,
// :337-430 Skewed, :505-575 LinearProfileSkewed). `type` folds away when it is a compile-time constant.
// returns false if the ray search found no cut side (reference throws, upwind.cpp:354).
// ------------------------------------------------------------------------------------------------
template <int E> NSB_DEV bool upwind_ip(int type, const double* __restrict__ x, const IpGeo<E>& g,
                                        const double* vel, double* up, double& len)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    if (type == UPW_NO) {
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = g.N[k];
        len = 1.0;
        return true;
    }
#pragma unroll
    for (int k = 0; k < NSH; k++) up[k] = 0.0;
    if (type == UPW_FULL) {
        const double flux = dotv<DIM>(g.n, vel);
        const int co = flux > 0.0 ? g.from : g.to;
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = (k == co) ? 1.0 : 0.0;
        len = sqrt(dist2<DIM>(g.xip, x + co * DIM));
        return true;
    }
    // skewed / linear profile skewed
    if (sqrt(dotv<DIM>(vel, vel)) < 1e-14) { len = 1.0; return true; }
    int side = 0; double gc[DIM], lc[DIM];
    if (!side_ray_cut<E>(x, g.xip, vel, side, gc, lc)) { len = 1.0; return false; }
    constexpr int NSC = (DIM == 2) ? 2 : (E == E_TET ? 3 : 4);
    if (type == UPW_SKEWED) {
        double mn = 1.79769313486231570e308; int best = 0;
#pragma unroll
        for (int i = 0; i < NSC; i++) {
            if (E == E_PRISM && i >= t_side_n<E>(side)) continue;
            const int co = t_side<E>(side, i);
            const double dd = dist2<DIM>(gc, x + co * DIM);
            if (dd < mn) { mn = dd; best = co; }
        }
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = (k == best) ? 1.0 : 0.0;
        len = sqrt(dist2<DIM>(g.xip, x + best * DIM));
    } else {
        double Nc[NSH];
        lagrange<E>(lc, Nc);
        int mask = 0;
#pragma unroll
        for (int i = 0; i < NSC; i++) if (E != E_PRISM || i < t_side_n<E>(side)) mask |= 1 << t_side<E>(side, i);
#pragma unroll
        for (int k = 0; k < NSH; k++) up[k] = ((mask >> k) & 1) ? Nc[k] : 0.0;
        len = sqrt(dist2<DIM>(g.xip, gc));
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Record of one ip, parked in shared memory between the per-ip phase and the column phase.
// It holds the flux derivative B(rf, cf; k) of add_jac_A_elem (fv1/navier_stokes_fv1.cpp:317-594) in a
// column-ready factored form, so the column phase is (almost) pure accumulation:
//     B(d1,d2;k) = A[k][d1]*n[d2] + delta(d1,d2)*D[k]  (+ Q[k][d1][d2] when the stabilisation is the upwind, PAC)
//     B(d1,P ;k) = N[k]*n[d1]                          (+ Pm[k][d1] for PAC)
//     B(P ,d2;k) = C[k][d2]            B(P,P;k) = CP[k]
// ------------------------------------------------------------------------------------------------
template <int E, bool PAC> struct IpRecPacPart {};
template <int E> struct IpRecPacPart<E, true> {
    double Q[ET<E>::NSH][ET<E>::DIM][ET<E>::DIM];
    double Pm[ET<E>::NSH][ET<E>::DIM];
};
template <int E, bool PAC> struct IpRec : IpRecPacPart<E, PAC> {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1;
    double n[DIM];
    double F[NF];           // defect fluxes (momentum d, continuity), add_def_A_elem :686-776
    double A[NSH][DIM];
    double D[NSH];
    double N[NSH];
    double C[NSH][DIM];
    double CP[NSH];
};

// Stabilisation shape accessors used while building the record. Diagonal branch: closed forms
// (stabilization.cpp:213-236 FIELDS, :536-582 FLOW, :826-849 none) from per-ip data in registers.
template <int E> struct StabDiag {
    static constexpr int DIM = ET<E>::DIM;
    const double* N; const double (*G)[DIM]; const double* sb; const double* std;
    double invdiag, inv_rho; int stab;
    NSB_DEV double sv(int d, int d2, int k) const
    {
        if (stab == STAB_NONE) return d == d2 ? N[k] : 0.0;
        if (stab == STAB_FIELDS) return d == d2 ? sb[k] * invdiag : 0.0;
        if (d == d2) {
            double s = sb[k];
#pragma unroll
            for (int q = 0; q < DIM; q++) if (q != d) s -= std[q] * G[k][q];
            return s * invdiag;
        }
        return std[d] * G[k][d2] * invdiag;
    }
    NSB_DEV double sp(int d, int k) const
    {
        if (stab == STAB_NONE) return 0.0;
        return -1.0 * G[k][d] * inv_rho * invdiag;
    }
};
// Dense branch: shapes live in shared memory arrays sv[ip][d][d2][k], sp[ip][d][k]
template <int E> struct StabDense {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    const double* svp; const double* spp;       // already offset to this ip
    NSB_DEV double sv(int d, int d2, int k) const { return svp[(d * DIM + d2) * NSH + k]; }
    NSB_DEV double sp(int d, int k) const { return spp[d * NSH + k]; }
};

// Builds the factored flux derivative of one ip (see IpRec). Terms and quirks of
// add_jac_A_elem, fv1/navier_stokes_fv1.cpp:336-592:
//   diffusion :336-356, pressure :363-368, convection by upwind :430-457 / by stabilisation (PAC) :400-427,
//   Peclet part :460-468, exact-Newton extras :475-550 (factor NOT applied to the upwind and Peclet
//   parts, :528-529,:542-545; un-connected PAC term :494-496), continuity :561-592.
// up/cvx: convective upwind shapes and sum_ip2 N_k(ip2)*up_ip(ip,ip2); U: transported velocity (blended).
template <int E, bool PAC, class SV>
NSB_DEV void ip_coeffs(const KParams& p, const double* n, const double* N, const double (*G)[ET<E>::DIM],
                       const double* up, const double* cvx, const double* U, double w, double prod,
                       const SV& S, bool connected, IpRec<E, PAC>& r)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    const double nurho = p.visc * p.rho;
#pragma unroll
    for (int d = 0; d < DIM; d++) r.n[d] = n[d];
#pragma unroll
    for (int k = 0; k < NSH; k++) {
        double D = -1.0 * nurho * dotv<DIM>(G[k], n);
        double A[DIM];
#pragma unroll
        for (int d1 = 0; d1 < DIM; d1++) A[d1] = p.laplace ? 0.0 : -1.0 * nurho * G[k][d1];
        if (!p.stokes) {
            if (!PAC) D += (up[k] + cvx[k]) * (prod * w);
            if (p.peclet) D += prod * (1.0 - w) * N[k];
            if (p.exact_jac != 0.0) {
                double e = 0.0;
                if (!PAC) e += w * up[k] * p.rho;
                if (p.peclet) e += (1.0 - w) * N[k] * p.rho;
#pragma unroll
                for (int d1 = 0; d1 < DIM; d1++) A[d1] += e * U[d1];
            }
        }
        r.D[k] = D; r.N[k] = N[k];
#pragma unroll
        for (int d1 = 0; d1 < DIM; d1++) r.A[k][d1] = A[d1];
        // continuity row
        double cp = 0.0;
#pragma unroll
        for (int q = 0; q < DIM; q++) cp += S.sp(q, k) * n[q] * p.rho;
        r.CP[k] = cp;
#pragma unroll
        for (int d2 = 0; d2 < DIM; d2++) {
            double cv = 0.0;
            if (connected) {
#pragma unroll
                for (int q = 0; q < DIM; q++) cv += S.sv(q, d2, k) * n[q] * p.rho;
            } else cv = S.sv(d2, d2, k) * n[d2] * p.rho;
            r.C[k][d2] = cv;
        }
        if constexpr (PAC) {
            double pp = 0.0;
            if (!p.stokes && p.exact_jac != 0.0) {
#pragma unroll
                for (int q = 0; q < DIM; q++) pp += S.sp(q, k) * n[q];
                pp *= p.exact_jac * p.rho;
            }
#pragma unroll
            for (int d1 = 0; d1 < DIM; d1++) {
                r.Pm[k][d1] = p.stokes ? 0.0 : prod * w * S.sp(d1, k) + pp * U[d1];
#pragma unroll
                for (int d2 = 0; d2 < DIM; d2++) {
                    double q = 0.0;
                    if (!p.stokes) {
                        if (connected || d1 == d2) q = prod * w * S.sv(d1, d2, k);
                        if (p.exact_jac != 0.0) {
                            double pv = 0.0;
                            if (connected) {
#pragma unroll
                                for (int z = 0; z < DIM; z++) pv += w * S.sv(z, d2, k) * n[z];
                            } else pv = S.sv(d1, d1, k) * n[d1];
                            pv *= p.exact_jac * p.rho;
                            q += pv * U[d1];
                        }
                    }
                    r.Q[k][d1][d2] = q;
                }
            }
        }
    }
}

// One COLUMN (corner k, function cf) of the flux derivative of one ip: val[rf] is added to row
// (rf, from) and subtracted from row (rf, to).
template <int E, bool PAC>
NSB_DEV void jac_col(const IpRec<E, PAC>& r, int k, int cf, double* val)
{
    constexpr int DIM = ET<E>::DIM, P = DIM;
    if (cf < DIM) {
        const double ncf = r.n[cf];
#pragma unroll
        for (int d1 = 0; d1 < DIM; d1++) {
            double v = r.A[k][d1] * ncf;
            if (d1 == cf) v += r.D[k];
            if constexpr (PAC) v += r.Q[k][d1][cf];
            val[d1] = v;
        }
        val[P] = r.C[k][cf];
    } else {
        const double Nk = r.N[k];
#pragma unroll
        for (int d1 = 0; d1 < DIM; d1++) {
            double v = Nk * r.n[d1];
            if constexpr (PAC) v += r.Pm[k][d1];
            val[d1] = v;
        }
        val[P] = r.CP[k];
    }
}

// peclet_blend, fv1/navier_stokes_fv1.cpp:871-892
template <int E> NSB_DEV double peclet_blend(double* U, const IpGeo<E>& g, const double* __restrict__ x,
                                             const double* std, double visc)
{
    constexpr int DIM = ET<E>::DIM;
    const double Pe = dotv<DIM>(std, g.n) / dotv<DIM>(g.n, g.n) * sqrt(dist2<DIM>(x + g.to * DIM, x + g.from * DIM)) / visc;
    const double Pe2 = Pe * Pe, w = Pe2 / (5.0 + Pe2);
#pragma unroll
    for (int d = 0; d < DIM; d++) U[d] = w * U[d] + (1.0 - w) * std[d];
    return w;
}

// ------------------------------------------------------------------------------------------------
// Everything of one ip for the DIAGONAL stabilisation branch (upwinds without ip shapes):
// geometry -> StdVel -> upwind(s) -> diffusion length -> FIELDS/FLOW/none closure -> defect fluxes.
// Restates the common prologue (fv1/navier_stokes_fv1.cpp:261-314), stabilization.cpp:142-241 /
// :456-587 / :805-850 and add_def_A_elem (:667-777) for one SCVF.
//   x  [NSH][DIM] corner coordinates, u [NSH][NF] the `u` argument, s0/s1 the local time series
//   (pSol/pOldSol; s0==u and s1==nullptr when stationary), vol [NSH] SCV volumes.
// returns false when a ray search failed.
// ------------------------------------------------------------------------------------------------
template <int E, bool PAC>
NSB_DEV bool ip_eval(const KParams& p, const double* __restrict__ x, const double* __restrict__ u,
                     const double* __restrict__ s0, const double* __restrict__ s1,
                     const double* __restrict__ vol, int ip,
                     double cor_minN, double cor_avgN, double cor_minD, IpRec<E, PAC>& r)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1, P = DIM;
    IpGeo<E> g;
    ip_geometry<E>(x, ip, g);
    bool ok = true;
    double std[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < NSH; k++) s += u[k * NF + d] * g.N[k];
        std[d] = s;
    }
    // ---- stabilisation's upwind (+ downwind for FLOW) ----
    double up[NSH], dn[NSH], uplen = 1.0, dnlen = 1.0;
    if (!p.stokes) {
        ok &= upwind_ip<E>(p.upw_stab, x, g, std, up, uplen);
        if (p.stab == STAB_FLOW) {
            double neg[DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) neg[d] = -1.0 * std[d];
            ok &= upwind_ip<E>(p.upw_stab, x, g, neg, dn, dnlen);
        }
    }
    // ---- Schneider-Raw closure, diagonal branch ----
    double stabvel[DIM], invdiag = 0.0, sb[NSH];
    if (p.stab == STAB_NONE) {
#pragma unroll
        for (int d = 0; d < DIM; d++) stabvel[d] = std[d];
#pragma unroll
        for (int k = 0; k < NSH; k++) sb[k] = 0.0;
    } else {
        const double nn = dotv<DIM>(g.n, g.n);
        const double a = p.visc * diff_len_sq_inv<DIM>(p.diff_len, nn, vol[g.from], vol[g.to], g.ds, cor_minN, cor_avgN, cor_minD);
        double b = 0.0, c = 0.0;
        if (!p.stokes) {
            const double nrm = sqrt(dotv<DIM>(std, std));
            b = nrm / uplen;
            if (p.stab == STAB_FLOW) c = nrm / (dnlen + uplen);
        }
        double diag = a;
        if (p.time_dep) diag += 1.0 / p.dt;
        if (!p.stokes) diag += b;
        invdiag = 1.0 / diag;
#pragma unroll
        for (int k = 0; k < NSH; k++) {
            double sv = a * g.N[k];
            if (!p.stokes) {
                sv += b * up[k];
                if (p.stab == STAB_FLOW) sv += c * (dn[k] - up[k]);
            }
            sb[k] = sv;
        }
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            double rhs = p.has_source ? p.src[d] : 0.0;
            if (p.time_dep) {
                double o = 0.0;
#pragma unroll
                for (int k = 0; k < NSH; k++) o += g.N[k] * s1[k * NF + d];
                rhs += o / p.dt;
            }
#pragma unroll
            for (int k = 0; k < NSH; k++) {
                double sv = sb[k];
                if (p.stab == STAB_FLOW) {
#pragma unroll
                    for (int q = 0; q < DIM; q++) if (q != d) sv -= std[q] * g.G[k][q];
                }
                rhs += sv * s0[k * NF + d];
                if (p.stab == STAB_FLOW) {
#pragma unroll
                    for (int q = 0; q < DIM; q++) if (q != d) rhs += std[d] * g.G[k][q] * s0[k * NF + q];
                }
                rhs += (-1.0 * g.G[k][d] * p.inv_rho) * s0[k * NF + P];
            }
            stabvel[d] = rhs * invdiag;
        }
    }
    // ---- convective upwind ----
    double cup[NSH];
    double U[DIM], w = 1.0;
#pragma unroll
    for (int d = 0; d < DIM; d++) U[d] = 0.0;
#pragma unroll
    for (int k = 0; k < NSH; k++) cup[k] = 0.0;
    if (!p.stokes) {
        if constexpr (PAC) {
#pragma unroll
            for (int d = 0; d < DIM; d++) U[d] = stabvel[d];
        } else {
            if (p.upw_conv == p.upw_stab) {
#pragma unroll
                for (int k = 0; k < NSH; k++) cup[k] = up[k];
            } else {
                double l2;
                ok &= upwind_ip<E>(p.upw_conv, x, g, std, cup, l2);
            }
#pragma unroll
            for (int k = 0; k < NSH; k++)
#pragma unroll
                for (int d = 0; d < DIM; d++) U[d] += cup[k] * u[k * NF + d];     // upwind_vel, upwind_interface.h:334-358
        }
        if (p.peclet) w = peclet_blend<E>(U, g, x, std, p.visc);
    }
    const double prod = dotv<DIM>(std, g.n) * p.rho;
    // ---- defect fluxes :686-776 ----
    if (p.what & W_DEF_A) {
        double gv[DIM][DIM];
#pragma unroll
        for (int d1 = 0; d1 < DIM; d1++)
#pragma unroll
            for (int d2 = 0; d2 < DIM; d2++) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < NSH; k++) s += g.G[k][d2] * u[k * NF + d1];
                gv[d1][d2] = s;
            }
        double pr = 0;
#pragma unroll
        for (int k = 0; k < NSH; k++) pr += g.N[k] * u[k * NF + P];
#pragma unroll
        for (int d1 = 0; d1 < DIM; d1++) {
            double df = 0;
#pragma unroll
            for (int d2 = 0; d2 < DIM; d2++) df += gv[d1][d2] * g.n[d2];
            if (!p.laplace) {
#pragma unroll
                for (int d2 = 0; d2 < DIM; d2++) df += gv[d2][d1] * g.n[d2];
            }
            double f = df * (-1.0) * p.visc * p.rho;
            if (!p.stokes) f += U[d1] * prod;
            f += pr * g.n[d1];
            r.F[d1] = f;
        }
        r.F[P] = dotv<DIM>(stabvel, g.n) * p.rho;
    }
    // ---- park the record: factored flux derivative ----
    if (p.what & W_JAC_A) {
        double zero[NSH];
#pragma unroll
        for (int k = 0; k < NSH; k++) zero[k] = 0.0;
        StabDiag<E> S{g.N, g.G, sb, std, invdiag, p.inv_rho, p.stab};
        ip_coeffs<E, PAC>(p, g.n, g.N, g.G, cup, zero, U, w, prod, S, p.stab == STAB_FLOW, r);
    }
    return ok;
}

}  // namespace nsb
