// ns_base.h -- types and small math helpers shared by every kernel of libnsb200 and by the host-side
// emulation harness of the fused patch kernel (tests/cpp/emu_fused.cpp compiles the NSB_HD functions with g++).
#pragma once
#include <math.h>
#include <stdint.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define NSB_HD __host__ __device__ __forceinline__
#define NSB_DEV __device__ __forceinline__
#else
#define NSB_HD inline
#define NSB_DEV inline
#ifndef __device__
#define __device__
#define __host__
#define __constant__ static
#define __forceinline__ inline
#endif
#endif
#include "ref_tables.cuh"

namespace nsb {

enum { E_TRI = 0, E_QUAD = 1, E_TET = 2, E_HEX = 3, E_PRISM = 4 };
enum { UPW_NONE = 0, UPW_NO = 1, UPW_FULL = 2, UPW_SKEWED = 3, UPW_LPS = 4, UPW_POSITIVE = 5 };
enum { STAB_FIELDS = 0, STAB_FLOW = 1, STAB_NONE = 2 };
enum { DIFF_RAW = 0, DIFF_FIVEPOINT = 1, DIFF_COR = 2 };
enum { W_JAC_A = 1, W_DEF_A = 2, W_JAC_M = 4, W_DEF_M = 8, W_RHS = 16 };

template <int E> struct ET;
template <> struct ET<E_TRI>  { static constexpr int DIM = 2, NSH = 3, NIP = 3,  NSIDE = 3, NINC = 2; };
template <> struct ET<E_QUAD> { static constexpr int DIM = 2, NSH = 4, NIP = 4,  NSIDE = 4, NINC = 2; };
template <> struct ET<E_TET>  { static constexpr int DIM = 3, NSH = 4, NIP = 6,  NSIDE = 4, NINC = 3; };
template <> struct ET<E_HEX>  { static constexpr int DIM = 3, NSH = 8, NIP = 12, NSIDE = 6, NINC = 3; };
template <> struct ET<E_PRISM> { static constexpr int DIM = 3, NSH = 6, NIP = 9, NSIDE = 5, NINC = 3; };   // element kernels only (fv1 / dense)

// reference-element tables by element type: the four original types index the [4][..] tables, the prism its own arrays
template <int E> NSB_DEV int t_edge(int ip, int j) { if constexpr (E == E_PRISM) return tab::P_EDGE[ip][j]; else return tab::EDGE[E][ip][j]; }
template <int E> NSB_DEV int t_side(int s, int i) { if constexpr (E == E_PRISM) return tab::P_SIDE[s][i]; else return tab::SIDE[E][s][i]; }
template <int E> NSB_DEV int t_side_n(int s) { if constexpr (E == E_PRISM) return tab::P_SIDE_N[s]; else return tab::SIDE_N[E][s]; }
template <int E> NSB_DEV int t_fa(int ip) { if constexpr (E == E_PRISM) return tab::P_SCVF_FA[ip]; else return tab::SCVF_FA[E][ip]; }
template <int E> NSB_DEV int t_fb(int ip) { if constexpr (E == E_PRISM) return tab::P_SCVF_FB[ip]; else return tab::SCVF_FB[E][ip]; }
template <int E> NSB_DEV double t_lip(int ip, int d) { if constexpr (E == E_PRISM) return tab::P_LIP[ip][d]; else return tab::LIP[E][ip][d]; }
template <int E> NSB_DEV double t_corner(int k, int d) { if constexpr (E == E_PRISM) return tab::P_CORNER[k][d]; else return tab::CORNER[E][k][d]; }
template <int E> NSB_DEV int t_inc(int k, int t) { if constexpr (E == E_PRISM) return tab::P_INC[k][t]; else return tab::INC[E][k][t]; }
template <int E> NSB_DEV int t_inc_sign(int k, int t) { if constexpr (E == E_PRISM) return tab::P_INC_SIGN[k][t]; else return tab::INC_SIGN[E][k][t]; }

// Runtime-uniform state of the disc (NavierStokesFV1 members; fv1/navier_stokes_fv1.h:604-614).
struct KParams {
    int upw_stab, upw_conv, stab, diff_len;
    int stokes, laplace, peclet, pac, time_dep, has_source;
    int what, defect_upwind;
    double exact_jac, visc, rho, inv_rho, dt, scale_a, scale_m, grad_div;
    double src[3];
};


// ------------------------------------------------------------------------------------------------
// P1 / Q1 Lagrange shapes (ugcore LagrangeP1) at a local point
// ------------------------------------------------------------------------------------------------
template <int E> NSB_HD void lagrange(const double* xi, double* N)
{
    if constexpr (E == E_TRI) { N[0] = 1.0 - xi[0] - xi[1]; N[1] = xi[0]; N[2] = xi[1]; }
    else if constexpr (E == E_TET) { N[0] = 1.0 - xi[0] - xi[1] - xi[2]; N[1] = xi[0]; N[2] = xi[1]; N[3] = xi[2]; }
    else if constexpr (E == E_QUAD) {
        const double x = xi[0], y = xi[1];
        N[0] = (1 - x) * (1 - y); N[1] = x * (1 - y); N[2] = x * y; N[3] = (1 - x) * y;
    } else if constexpr (E == E_PRISM) {
        // P1 on the triangle times P1 along the axis (ugcore LagrangeP1<ReferencePrism>)
        const double l0 = 1.0 - xi[0] - xi[1], l1 = xi[0], l2 = xi[1], z = xi[2];
        N[0] = l0 * (1 - z); N[1] = l1 * (1 - z); N[2] = l2 * (1 - z); N[3] = l0 * z; N[4] = l1 * z; N[5] = l2 * z;
    } else {
        const double x = xi[0], y = xi[1], z = xi[2];
        const double a0 = (1 - x) * (1 - y), a1 = x * (1 - y), a2 = x * y, a3 = (1 - x) * y;
        N[0] = a0 * (1 - z); N[1] = a1 * (1 - z); N[2] = a2 * (1 - z); N[3] = a3 * (1 - z);
        N[4] = a0 * z;       N[5] = a1 * z;       N[6] = a2 * z;       N[7] = a3 * z;
    }
}

template <int E> NSB_HD void lagrange_grad(const double* xi, double (*dN)[ET<E>::DIM])
{
    if constexpr (E == E_TRI) { dN[0][0] = -1; dN[0][1] = -1; dN[1][0] = 1; dN[1][1] = 0; dN[2][0] = 0; dN[2][1] = 1; }
    else if constexpr (E == E_TET) {
        dN[0][0] = -1; dN[0][1] = -1; dN[0][2] = -1; dN[1][0] = 1; dN[1][1] = 0; dN[1][2] = 0;
        dN[2][0] = 0; dN[2][1] = 1; dN[2][2] = 0; dN[3][0] = 0; dN[3][1] = 0; dN[3][2] = 1;
    } else if constexpr (E == E_QUAD) {
        const double x = xi[0], y = xi[1];
        dN[0][0] = -(1 - y); dN[0][1] = -(1 - x); dN[1][0] = (1 - y); dN[1][1] = -x;
        dN[2][0] = y;        dN[2][1] = x;        dN[3][0] = -y;      dN[3][1] = (1 - x);
    } else if constexpr (E == E_PRISM) {
        const double l[3] = {1.0 - xi[0] - xi[1], xi[0], xi[1]}, z = xi[2];
        const double dl[3][2] = {{-1, -1}, {1, 0}, {0, 1}};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int t = k % 3;
            const double fz = k < 3 ? 1 - z : z, sz = k < 3 ? -1.0 : 1.0;
            dN[k][0] = dl[t][0] * fz; dN[k][1] = dl[t][1] * fz; dN[k][2] = l[t] * sz;
        }
    } else {
        const double x = xi[0], y = xi[1], z = xi[2];
        const double fx[2] = {1 - x, x}, fy[2] = {1 - y, y}, fz[2] = {1 - z, z};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int bx = (k & 1) ^ ((k >> 1) & 1), by = (k >> 1) & 1, bz = (k >> 2) & 1;
            dN[k][0] = (bx ? 1.0 : -1.0) * fy[by] * fz[bz];
            dN[k][1] = fx[bx] * (by ? 1.0 : -1.0) * fz[bz];
            dN[k][2] = fx[bx] * fy[by] * (bz ? 1.0 : -1.0);
        }
    }
}

template <int DIM> NSB_HD double dotv(const double* a, const double* b)
{
    double s = a[0] * b[0];
#pragma unroll
    for (int d = 1; d < DIM; d++) s += a[d] * b[d];
    return s;
}
template <int DIM> NSB_HD double dist2(const double* a, const double* b)
{
    double s = 0;
#pragma unroll
    for (int d = 0; d < DIM; d++) { const double t = a[d] - b[d]; s += t * t; }
    return s;
}
NSB_HD void cross3(double* o, const double* a, const double* b)
{ o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }

// inverse of a DIM x DIM matrix; returns the determinant
template <int DIM> NSB_HD double inv_mat(const double (*a)[DIM], double (*inv)[DIM])
{
    if constexpr (DIM == 2) {
        const double det = a[0][0] * a[1][1] - a[0][1] * a[1][0], r = 1.0 / det;
        inv[0][0] = a[1][1] * r; inv[0][1] = -a[0][1] * r; inv[1][0] = -a[1][0] * r; inv[1][1] = a[0][0] * r;
        return det;
    } else {
        const double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
        const double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
        const double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
        const double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02, r = 1.0 / det;
        inv[0][0] = c00 * r; inv[1][0] = c01 * r; inv[2][0] = c02 * r;
        inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * r;
        inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * r;
        inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * r;
        inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * r;
        inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * r;
        inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * r;
        return det;
    }
}

// Diffusion length (fv1/diffusion_length.h:47-198). cor_* are the element-wide min/avg needed by COR.
template <int DIM> NSB_HD double diff_len_sq_inv(int type, double nn, double volf, double volt, double ds,
                                                  double cor_minN, double cor_avgN, double cor_minD)
{
    double A = 0.5 * (volf + volt); A *= A;
    if (type == DIFF_RAW)       return DIM == 2 ? 1.0 / (0.5 * A / nn + 3.0 * nn / 8.0) : 1.0 / (0.5 * A / nn + 3.0 * ds / 8.0);
    if (type == DIFF_FIVEPOINT) return DIM == 2 ? 2.0 * nn / A + 8.0 / nn : 2.0 * nn / A + 8.0 * ds / nn;
    return DIM == 2 ? 2.0 * cor_minN / A + 8.0 / (3.0 * cor_avgN) : 2.0 * cor_minN / A + 8.0 * cor_minD / (3.0 * cor_avgN);
}


// lean SCVF record of the split / fused paths: [F | n | cK | dK | pK = -G_k.n / diag] (see ns_split.cuh)
template <int E> struct LeanRec {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1;
    static constexpr int NSHP = (NSH + 1) & ~1;
    static constexpr int O_F = 0, O_N = NF, HEAD = (NF + DIM + 1) & ~1;
    static constexpr int O_CK = HEAD, O_DK = O_CK + NSHP, O_PK = O_DK + NSHP, RAW = O_PK + NSHP;
    static constexpr int SZ = (RAW + 3) & ~3;              // hex 32 doubles = 256 B; tet / quad / tri 20 doubles = 160 B
};

// compressed SCVF record of the split path for the 3-D element types (and of the fused tile kernel, ns_tile.cuh):
//   [ F[4] | n[3] alpha | beta cw | cpe mv0 | mv1 mv2 | up[NSH] ]     (hex 22 doubles = 176 B, tet 18 doubles = 144 B)
// from which the consumer forms, with the constant tables N_k(ip), dN_k(ip):
//   cK_k = alpha N_k + beta up_k (continuity row, :561-584), dK_k = cw up_k + cpe N_k (convective diagonal, :430-468),
//   pK_k = dN_k . mv (-G_k.n / diag, :586-592). One set of upwind shapes: stabilisation and convection use the same upwind.
template <int E> struct CompRec {
    static constexpr int NSH = ET<E>::NSH;
    static constexpr int O_F = 0, O_N = 4, O_AL = 7, O_BE = 8, O_CW = 9, O_CPE = 10, O_MV = 11, O_UP = 14, SZ = 14 + NSH, HEAD = 4;
    static constexpr int TSTR = NSH * 4 + 2;                // doubles per ip of the consumer's table [ip][k] -> (dN0, dN1, dN2, N), 16-byte rows, ip rows in distinct banks
};

}  // namespace nsb
