// ns_turb.cuh -- SURVEY 8f-4: the Smagorinsky turbulent viscosity as a DEVICE-SIDE provider of the per-ip kinematic viscosity
// import, and the diagnostics of navier_stokes_tools.h.
//
//   FV1SmagorinskyTurbViscData (fv1/turbulent_viscosity_fv1.h:200-383):
//     assembleDeformationTensor (fv1/turbulent_viscosity_fv1_impl.h:504-616)  D_a = 1/vol_a [ sum_scvf +-(1/2)(u_ip n^T + n u_ip^T)
//                                   + sum_{BF of the turbulence-zero subsets} (1/2)(u_a n^T + n u_a^T) ]
//     update (:819-852), FNorm (:755-762)                                     nu_t(a) = c vol_a^(2/dim) sqrt(2 sum D_ij^2), 0 on the
//                                                                             turbulence-zero subsets
//     evaluate (turbulent_viscosity_fv1.h:321-379)                            nu(ip) = sum_sh N_sh(ip) nu_t(sh) + kinematic viscosity
//   The reference scatters element by element into vertex attachments; here ONE THREAD per grid node gathers its incident SCVFs
//   over the node -> (element, corner) adjacency of the owner-computes path (fixed order, no atomics), a second kernel interpolates
//   to the SCVF ips and writes the import table that the element kernels read (MeshDev::ip_visc).
//   vorticityFV1 (navier_stokes_tools.h:386-525): the same gather with the shape gradients at the SCV ips (= corners).
//   kineticEnergy (:850-965) / cflNumber (:731-848) of a Crouzeix-Raviart field: per-element values, fixed-order tree reduction.
#pragma once
#include "ns_base.h"
#include "ns_kernels.cuh"
#include "ns_bnd.cuh"
#include "ns_fvcr.cuh"

namespace nsb {

// area-scaled SCVF normal from the global element corners (App. B-2: 2-D (dy, -dx) of barycentre - edge midpoint;
// 3-D 0.5 (c2 - c0) x (c3 - c1) with c = edge midpoint, centre of face A, barycentre, centre of face B)
template <int E>
NSB_DEV void scvf_normal_of(const double (*x)[ET<E>::DIM], int ip, double* n)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH;
    const int ea = tab::EDGE[E][ip][0], eb = tab::EDGE[E][ip][1];
    double c0[DIM], c2[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        c0[d] = (x[ea][d] + x[eb][d]) / 2;
        double s = 0.0;
        for (int k = 0; k < NSH; k++) s += x[k][d];
        c2[d] = s / NSH;
    }
    if constexpr (DIM == 2) { n[0] = c2[1] - c0[1]; n[1] = -(c2[0] - c0[0]); }
    else {
        const int fa = tab::SCVF_FA[E][ip], fb = tab::SCVF_FB[E][ip];
        const int na = tab::SIDE_N[E][fa], nb = tab::SIDE_N[E][fb];
        double av[3], bv[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double sa = 0.0, sb = 0.0;
            for (int k = 0; k < na; k++) sa += x[tab::SIDE[E][fa][k]][d];
            for (int k = 0; k < nb; k++) sb += x[tab::SIDE[E][fb][k]][d];
            av[d] = c2[d] - c0[d]; bv[d] = sb / nb - sa / na;
        }
        cross3(n, av, bv);
#pragma unroll
        for (int d = 0; d < 3; d++) n[d] *= 0.5;
    }
}

// nodal Smagorinsky viscosity. bidx [n_node]: index of the node among the nodes of the turbulence-zero boundary sides (-1: not
// on them) into dbf = their BF closure sums (fv1_smagorinsky_bf_kernel); zflag [n_node]: 1 = vertex of a turbulence-zero subset
// (nu_t = 0). All three may be null.
template <int E>
__global__ void __launch_bounds__(128) fv1_smagorinsky_kernel(MeshDev m, const double* __restrict__ u, double cmodel,
                                                              const int32_t* __restrict__ bidx, const double* __restrict__ dbf,
                                                              const uint8_t* __restrict__ zflag, double* __restrict__ nu_t)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1, NINC = ET<E>::NINC;
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m.n_node) return;
    double D[DIM][DIM], vol = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) D[i][j] = 0.0;
    for (int64_t q = m.adj_ptr[a]; q < m.adj_ptr[a + 1]; q++) {
        const int32_t ad = m.adj[q];
        const int64_t e = ad / NSH; const int la = ad - (int)e * NSH;
        double x[NSH][DIM], uv[NSH][DIM];
        for (int k = 0; k < NSH; k++) {
            const int64_t g = m.conn[e * NSH + k];
#pragma unroll
            for (int d = 0; d < DIM; d++) { x[k][d] = m.coords[g * DIM + d]; uv[k][d] = u[g * NF + d]; }
        }
        vol += m.scvvol[e * NSH + la];
        for (int t = 0; t < NINC; t++) {
            const int ip = tab::INC[E][la][t];
            const double sg = tab::INC_SIGN[E][la][t] < 0 ? -0.5 : 0.5;
            double n[DIM], v[DIM];
            scvf_normal_of<E>(x, ip, n);
#pragma unroll
            for (int d = 0; d < DIM; d++) { double s = 0.0; for (int k = 0; k < NSH; k++) s += tab::NIPSH[E][ip][k] * uv[k][d]; v[d] = s; }
#pragma unroll
            for (int i = 0; i < DIM; i++)
#pragma unroll
                for (int j = 0; j < DIM; j++) D[i][j] += sg * (v[i] * n[j] + v[j] * n[i]);
        }
    }
    if ((zflag && zflag[a]) || !(vol > 0.0)) { nu_t[a] = 0.0; return; }   // update(): vertices of the turbulence-zero subsets are skipped (:833)
    const int bi = bidx ? bidx[a] : -1;
    if (bi >= 0) {                                              // BF closure of the turbulence-zero sides (:591-602)
#pragma unroll
        for (int i = 0; i < DIM; i++)
#pragma unroll
            for (int j = 0; j < DIM; j++) D[i][j] += dbf[(int64_t)bi * (DIM * DIM) + i * DIM + j];
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) { const double t = D[i][j] / vol; s += t * t; }
    const double delta = pow(vol, 1.0 / DIM);
    nu_t[a] = cmodel * delta * delta * sqrt(2.0 * s);
}

// BF closure sums (1/2)(u_a n^T + n u_a^T) of the nodes on the turbulence-zero sides, one thread per such node
template <int E>
__global__ void __launch_bounds__(64) fv1_smagorinsky_bf_kernel(MeshDev m, const double* __restrict__ u, int64_t n_bnode,
                                                                const int32_t* __restrict__ bnode, const int64_t* __restrict__ bptr,
                                                                const BndFace* __restrict__ bf, double* __restrict__ dbf)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bnode) return;
    const int64_t a = bnode[i];
    double D[DIM][DIM];
#pragma unroll
    for (int r = 0; r < DIM; r++)
#pragma unroll
        for (int s = 0; s < DIM; s++) D[r][s] = 0.0;
    for (int64_t q = bptr[i]; q < bptr[i + 1]; q++) {
        const BndFace f = bf[q];
        double x[NSH][DIM], n[DIM], lip[DIM];
        for (int k = 0; k < NSH; k++) {
            const int64_t g = m.conn[(int64_t)f.elem * NSH + k];
#pragma unroll
            for (int d = 0; d < DIM; d++) x[k][d] = m.coords[g * DIM + d];
        }
        bf_normal_lip<E>(x, f.side, f.j, n, lip);
#pragma unroll
        for (int r = 0; r < DIM; r++)
#pragma unroll
            for (int s = 0; s < DIM; s++) D[r][s] += 0.5 * (u[a * NF + r] * n[s] + u[a * NF + s] * n[r]);
    }
#pragma unroll
    for (int r = 0; r < DIM; r++)
#pragma unroll
        for (int s = 0; s < DIM; s++) dbf[i * (DIM * DIM) + r * DIM + s] = D[r][s];
}

// nu(ip) = sum_sh N_sh(ip) nu_t(sh) + kinematic viscosity  -> the per-ip import table [n_elem][NIP]
template <int E>
__global__ void __launch_bounds__(256) fv1_ip_visc_kernel(MeshDev m, const double* __restrict__ nu_t, double visc, double* __restrict__ ipv)
{
    constexpr int NSH = ET<E>::NSH, NIP = ET<E>::NIP;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.n_elem * NIP) return;
    const int64_t e = i / NIP; const int ip = (int)(i - e * NIP);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NSH; k++) s += tab::NIPSH[E][ip][k] * nu_t[m.conn[e * NSH + k]];
    ipv[i] = s + visc;
}

// vorticityFV1: one thread per node
template <int E>
__global__ void __launch_bounds__(128) fv1_vorticity_kernel(MeshDev m, const double* __restrict__ u, double* __restrict__ vort, int* __restrict__ errflag)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1;
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m.n_node) return;
    double w = 0.0, vol = 0.0;
    for (int64_t q = m.adj_ptr[a]; q < m.adj_ptr[a + 1]; q++) {
        const int32_t ad = m.adj[q];
        const int64_t e = ad / NSH; const int la = ad - (int)e * NSH;
        double x[NSH][DIM], uv[NSH][2], xi[DIM], dN[NSH][DIM], JT[DIM][DIM], JI[DIM][DIM];
        for (int k = 0; k < NSH; k++) {
            const int64_t g = m.conn[e * NSH + k];
#pragma unroll
            for (int d = 0; d < DIM; d++) x[k][d] = m.coords[g * DIM + d];
            uv[k][0] = u[g * NF]; uv[k][1] = u[g * NF + 1];
        }
#pragma unroll
        for (int d = 0; d < DIM; d++) xi[d] = tab::CORNER[E][la][d];
        lagrange_grad<E>(xi, dN);
#pragma unroll
        for (int r = 0; r < DIM; r++)
#pragma unroll
            for (int s = 0; s < DIM; s++) { double t = 0.0; for (int k = 0; k < NSH; k++) t += dN[k][r] * x[k][s]; JT[r][s] = t; }
        const double det = inv_mat<DIM>(JT, JI);
        if (!(fabs(det) > 0.0)) { atomicExch(errflag, 2); continue; }
        double lw = 0.0;
        for (int k = 0; k < NSH; k++) {
            double g0 = 0.0, g1 = 0.0;
#pragma unroll
            for (int r = 0; r < DIM; r++) { g0 += JI[0][r] * dN[k][r]; g1 += JI[1][r] * dN[k][r]; }
            lw += uv[k][1] * g0 - uv[k][0] * g1;
        }
        const double v = m.scvvol[e * NSH + la];
        w += lw * v; vol += v;
    }
    vort[a] = vol > 0.0 ? w / vol : 0.0;
}

// kinetic energy / CFL number of a Crouzeix-Raviart field: per-block partial sums (fixed tree order -> deterministic)
template <int E>
__global__ void __launch_bounds__(256) fvcr_diag_kernel(FvcrDev f, const double* __restrict__ u, double dt, double* __restrict__ part /*[nblk][3]*/)
{
    constexpr int DIM = CRT<E>::DIM, NCO = CRT<E>::NCO, NS = CRT<E>::NS;
    __shared__ double sh[3][256];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double en = 0.0, vo = 0.0, cfl = 0.0;
    if (e < f.n_elem) {
        double x[NCO][DIM], val[DIM], xs[NS][DIM];
        for (int k = 0; k < NCO; k++) {
            const int64_t g = f.conn[e * NCO + k];
#pragma unroll
            for (int d = 0; d < DIM; d++) x[k][d] = f.coords[g * DIM + d];
        }
#pragma unroll
        for (int d = 0; d < DIM; d++) val[d] = 0.0;
        const double Nb = 1.0 - DIM * (1.0 / (DIM + 1));                  // CR shape of every side at the barycentre of a simplex
        for (int s = 0; s < NS; s++) {
            const int64_t sd = f.esides[e * NS + s];
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                val[d] += Nb * u[sd * DIM + d];
                double t = 0.0;
                for (int k = 0; k < DIM; k++) t += x[tab::SIDE[E][s][k]][d];
                xs[s][d] = t / DIM;
            }
        }
        double ve;
        if constexpr (DIM == 2) ve = 0.5 * fabs((x[1][0] - x[0][0]) * (x[2][1] - x[0][1]) - (x[2][0] - x[0][0]) * (x[1][1] - x[0][1]));
        else {
            double av[3], bv[3], cv[3], t[3];
#pragma unroll
            for (int d = 0; d < 3; d++) { av[d] = x[1][d] - x[0][d]; bv[d] = x[2][d] - x[0][d]; cv[d] = x[3][d] - x[0][d]; }
            cross3(t, av, bv);
            ve = fabs(t[0] * cv[0] + t[1] * cv[1] + t[2] * cv[2]) / 6.0;
        }
        vo = ve;
#pragma unroll
        for (int d = 0; d < DIM; d++) en += ve * val[d] * val[d];
        for (int i = 0; i < NS; i++)
            for (int j = i + 1; j < NS; j++) {
                double q = 0.0, dd = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) { const double sb = xs[i][d] - xs[j][d]; q += sb * val[d]; dd += sb * sb; }
                const double l = dt * 1.0 / dd * fabs(q);
                if (l > cfl) cfl = l;
            }
    }
    sh[0][threadIdx.x] = en; sh[1][threadIdx.x] = vo; sh[2][threadIdx.x] = cfl;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
            sh[2][threadIdx.x] = fmax(sh[2][threadIdx.x], sh[2][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[blockIdx.x * 3] = sh[0][0]; part[blockIdx.x * 3 + 1] = sh[1][0]; part[blockIdx.x * 3 + 2] = sh[2][0]; }
}

// final stage: one block, the partials are summed in index order by strided serial loops + the same fixed tree
__global__ void __launch_bounds__(256) diag_final_kernel(int64_t nblk, const double* __restrict__ part, double* __restrict__ out /*[2]*/)
{
    __shared__ double sh[3][256];
    double en = 0.0, vo = 0.0, cfl = 0.0;
    for (int64_t i = threadIdx.x; i < nblk; i += 256) { en += part[i * 3]; vo += part[i * 3 + 1]; cfl = fmax(cfl, part[i * 3 + 2]); }
    sh[0][threadIdx.x] = en; sh[1][threadIdx.x] = vo; sh[2][threadIdx.x] = cfl;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
            sh[2][threadIdx.x] = fmax(sh[2][threadIdx.x], sh[2][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sh[0][0] / sh[1][0]; out[1] = sh[2][0]; }
}

}  // namespace nsb
