// ns_split.cuh -- owner-computes FV1 assembly with the Jacobian split into a STATIC and a STATE part
// (FIELDS / no stabilisation, fixed-point Jacobian; FLOW and exact-Newton keep the general rows kernel of ns_owner.cuh).
//
// For constant viscosity / density the local Jacobian of add_jac_A_elem (fv1/navier_stokes_fv1.cpp:317-594) is
//     J = nu rho * S  +  P  +  state part,
//   S : diffusion  -(G_k,d1 n_d2 [unless laplace] + delta_d1d2 G_k.n)      (:336-356)   geometry only
//   P : pressure gradient  N_k n_d1 in the pressure column                  (:363-368)   geometry only
//   state part : convective diagonal dK_k (:430-468), continuity row cK_k n_d2 (:561-584) and -G_k.n / diag (:586-592).
// S and P are assembled ONCE per mesh into the table J0 (momentum rows only, same slot order as the CSR rows).
// Every pass then moves, per SCVF, a lean 256-byte record (hex) instead of the 480-byte geometry + flux record, and
// accumulates 5 instead of 16 values per (node, element, corner):
//
//   fv1_j0_kernel          once per mesh: warp per node, deterministic (adjacency order), J0[node][rf < DIM][slot][cf]
//   fv1_flux_kernel<LEAN>  (ns_owner.cuh) writes [F | n | cK | dK | pK = -G_k.n / diag] per SCVF
//   fv1_rows_split_kernel  warp per node: lane = (element, corner) loads its coefficients of the NINC incident lean records
//                          directly and sums them in registers; per-slot accumulators in shared memory; the node's J0
//                          rows arrive by one TMA bulk copy; rows written once:
//                          out = {nu rho, 1} * scale_a * J0 + state part (+ lumped mass).
#pragma once
#include "ns_owner.cuh"

namespace nsb {


// ---- static part, once per mesh ---------------------------------------------------------------------
// J0 block row of node a: [rf < DIM][slot < cnt][cf < NF] at offset DIM*NF*brow[a]:
//   cf < DIM : sum over incident SCVFs of  -sign * (G_k[rf] n[cf] (unless laplace) + delta_rf,cf G_k.n)   (to be scaled by nu rho)
//   cf = DIM : sum of  sign * N_k n[rf]
template <int E>
__global__ void __launch_bounds__(128) fv1_j0_kernel(MeshDev m, int laplace, double* __restrict__ j0)
{
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1, NINC = ET<E>::NINC;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t a = warp0; a < m.n_node; a += nwarp) {
        const int64_t q0 = m.adj_ptr[a], q1 = m.adj_ptr[a + 1], b0 = m.brow[a];
        const int rowlen = (int)(m.brow[a + 1] - b0) * NF;
        double* out = j0 + b0 * (DIM * NF);
        for (int64_t q = q0; q < q1; q++) {                     // adjacency order (ascending element index): deterministic
            const int32_t ad = m.adj[q];
            const int e = ad / NSH, la = ad - e * NSH;
            double x[NSH * DIM];
#pragma unroll
            for (int k = 0; k < NSH; k++) {
                const int64_t nd = m.conn[(int64_t)e * NSH + k];
#pragma unroll
                for (int d = 0; d < DIM; d++) x[k * DIM + d] = m.coords[nd * DIM + d];
            }
            double S[DIM][NF];
#pragma unroll
            for (int rf = 0; rf < DIM; rf++)
#pragma unroll
                for (int cf = 0; cf < NF; cf++) S[rf][cf] = 0.0;
            for (int t = 0; t < NINC; t++) {
                const int ip = tab::INC[E][la][t];
                const double sg = (double)tab::INC_SIGN[E][la][t];
                IpGeo<E> g;
                ip_geometry<E>(x, ip, g);
                double Gk[DIM], Nk = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) Gk[d] = 0.0;
#pragma unroll
                for (int kk = 0; kk < NSH; kk++)
                    if (kk == lane) {
                        Nk = g.N[kk];
#pragma unroll
                        for (int d = 0; d < DIM; d++) Gk[d] = g.G[kk][d];
                    }
                const double gn = dotv<DIM>(Gk, g.n);
#pragma unroll
                for (int rf = 0; rf < DIM; rf++) {
                    if (!laplace) {
#pragma unroll
                        for (int cf = 0; cf < DIM; cf++) S[rf][cf] -= Gk[rf] * (sg * g.n[cf]);
                    }
                    S[rf][rf] -= sg * gn;
                    S[rf][DIM] += Nk * (sg * g.n[rf]);
                }
            }
            if (lane < NSH) {
                const int slot = m.emap[(int64_t)ad * NSH + lane];
#pragma unroll
                for (int rf = 0; rf < DIM; rf++)
#pragma unroll
                    for (int cf = 0; cf < NF; cf++) out[rf * rowlen + slot * NF + cf] += S[rf][cf];
            }
            __syncwarp();                                        // orders the read-modify-writes of successive elements
        }
    }
}

// ---- rows kernel of the split path --------------------------------------------------------------------
// warp per node (atomic tickets). Per round of CH adjacent elements: one TMA bulk copy per incident lean record, all in
// flight at once on a warp-private mbarrier (plus, in the first round, the node's J0 rows). lane = (jj, k) = corner k of
// the jj-th of JP = 32 / NSH elements handled in parallel sums its NINC SCVFs in registers and adds the 5 values into
// the per-slot accumulators of ITS copy jj (the same neighbour may be a corner of several adjacent elements; the
// corners of one element are distinct nodes -> no conflicts inside a copy). The copies are merged in fixed order ->
// bitwise deterministic. Rows are written once: out = {nu rho, 1} * scale_a * J0 + state part (+ lumped mass).
template <int E, int CHP = 0> struct SplitCfg {
    static constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1, NINC = ET<E>::NINC, NIP = ET<E>::NIP;
    static constexpr int JP = 32 / NSH;                            // adjacent elements handled in parallel
    static constexpr int CH = CHP ? CHP : ((DIM == 3) ? 8 : 16);   // adjacent elements staged per round (CH * NINC <= 32)
    static constexpr int NREC = CH * NINC;
    static constexpr int NV = DIM + 2;                             // accumulated values per slot: D, C[DIM], PP
    static_assert(NREC <= 32, "one lane issues one record copy");
};
// record format of the split path: compressed for the 3-D element types (CompRec, ns_base.h), lean otherwise
template <int E> struct SplitRec {
    static constexpr bool COMP = ET<E>::DIM == 3;
    static constexpr int SZ = COMP ? CompRec<E>::SZ : LeanRec<E>::SZ, HEAD = COMP ? CompRec<E>::HEAD : LeanRec<E>::HEAD;
};
template <int E> __host__ __device__ constexpr size_t split_tab_bytes()
{
    return ((sizeof(int32_t) * ET<E>::NSH * ET<E>::NINC + 15) & ~(size_t)15) + (SplitRec<E>::COMP ? sizeof(double) * ET<E>::NIP * CompRec<E>::TSTR : 0);
}
template <int E, int CHP = 0> struct SplitWS {
    using C = SplitCfg<E, CHP>;
    static constexpr int RS = SplitRec<E>::SZ, SS = rows_smem_stride<E>(RS);
    alignas(16) double rec[C::NREC][SS];
    unsigned long long bar;
    int32_t ipx[C::NREC];
    uint8_t slot[C::CH][8];
};
__host__ __device__ constexpr int split_cnt_pad(int max_cnt) { return (max_cnt + 1) & ~1; }
// J0D: the node's J0 rows are read straight from global memory in the output stage (prefetched into L2 at the top of
// the node) instead of being staged in shared memory by a bulk copy: smaller footprint per warp -> more warps per SM.
template <int E, int CHP, bool J0D = false> __host__ __device__ constexpr size_t split_warp_bytes(int max_cnt)
{
    using C = SplitCfg<E, CHP>;
    return (sizeof(SplitWS<E, CHP>) + sizeof(double) * ((J0D ? 0 : (size_t)C::DIM * C::NF * max_cnt) + (size_t)C::JP * C::NV * split_cnt_pad(max_cnt)) + 15) & ~(size_t)15;
}

// FAST: the standard pass (what = JAC_A | DEF_A, beta = 0) with the flags folded at compile time.
template <int E, int CHP = 0, int MINB = 12, bool J0D = false, bool FAST = false>
__global__ void __launch_bounds__(64, MINB) fv1_rows_split_kernel(KParams p, MeshDev m, const double* __restrict__ rec,
                                                                const double* __restrict__ j0, const double* __restrict__ u,
                                                                double beta, double* __restrict__ val, double* __restrict__ def,
                                                                unsigned long long* __restrict__ work_counter)
{
    using C = SplitCfg<E, CHP>;
    using LR = LeanRec<E>;
    using WS = SplitWS<E, CHP>;
    constexpr int DIM = C::DIM, NSH = C::NSH, NF = C::NF, NINC = C::NINC, CH = C::CH, NIP = C::NIP, NV = C::NV, JP = C::JP, RS = WS::RS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // block layout: [inc table NSH*NINC ints, padded to 16 B][per warp: WS | j0 rows DIM*NF*max_cnt | acc JP*NV*cntp]
    int32_t* inctab = reinterpret_cast<int32_t*>(smem_raw);
    constexpr size_t tab_bytes = split_tab_bytes<E>();
    constexpr bool COMP = SplitRec<E>::COMP;
    using CR = CompRec<E>;
    const double* tab4 = reinterpret_cast<const double*>(smem_raw + ((sizeof(int32_t) * NSH * NINC + 15) & ~(size_t)15));   // [NIP][TSTR]: (dN0, dN1, dN2, N) per corner
    const int cntp = split_cnt_pad(m.max_cnt);
    const int accn = NV * cntp;                                  // doubles per accumulator copy
    const size_t per_warp = split_warp_bytes<E, CHP, J0D>(m.max_cnt);
    WS& ws = *reinterpret_cast<WS*>(smem_raw + tab_bytes + warp * per_warp);
    double* j0s = reinterpret_cast<double*>(smem_raw + tab_bytes + warp * per_warp + sizeof(WS));
    double* acc = j0s + (J0D ? 0 : (size_t)DIM * NF * m.max_cnt);
    for (int i = threadIdx.x; i < NSH * NINC; i += blockDim.x)
        inctab[i] = tab::INC[E][i / NINC][i % NINC] | (tab::INC_SIGN[E][i / NINC][i % NINC] < 0 ? 256 : 0);
    if constexpr (COMP) {
        double* t4 = const_cast<double*>(tab4);
        for (int i = threadIdx.x; i < NIP * NSH * 4; i += blockDim.x) {
            const int ip = i / (NSH * 4), kk = (i >> 2) % NSH, c = i & 3;
            t4[ip * CR::TSTR + kk * 4 + c] = c < 3 ? tab::C_DNIP[E][ip][kk][c < DIM ? c : 0] : tab::NIPSH[E][ip][kk];
        }
    }
    if (lane == 0) mbar_init(&ws.bar, 1);
    __syncthreads();
    unsigned phase = 0;
    // L2 priorities: a lean record has a second reader (the other node of its edge) -> evict_last; the J0 rows are a stream
    const unsigned long long pol_rec = l2_policy_evict_last(), pol_j0 = l2_policy_evict_first();
    const bool l2hint = m.l2_hints != 0;
    const int what = FAST ? (W_JAC_A | W_DEF_A) : p.what;
    if (FAST) beta = 0.0;
    const bool want_jac = what & (W_JAC_A | W_JAC_M), want_def = what & (W_DEF_A | W_DEF_M | W_RHS);
    const bool jac_a = what & W_JAC_A, def_a = what & W_DEF_A;
    // a defect-only pass needs the fluxes only (head of the record)
    const unsigned cp_bytes = (unsigned)sizeof(double) * (jac_a ? RS : SplitRec<E>::HEAD);
    const int jj = lane / NSH, k = lane - jj * NSH;
    const bool lane_on = jj < JP;
    double* accj = acc + (lane_on ? jj : 0) * accn;
    const double s_visc = p.visc * p.rho * p.scale_a, s_pres = p.scale_a;
    (void)NIP;

    // The header of a node is a chain of dependent global loads (ticket -> adjacency range / block row -> adjacency
    // entries). It is fetched one node ahead: the ticket is taken at the top of the previous node, the ranges are
    // loaded while that node's records are in flight, the first adjacency entries while its rows are written.
    // Tickets are taken TG nodes at a time (one atomic per TG nodes; neighbouring warps still work on a tight window).
    const int TG = m.ticket_group;
    auto take = [&]() -> unsigned long long { unsigned long long t = 0; if (lane == 0) t = atomicAdd(work_counter, (unsigned long long)TG); return t; };
    int64_t tk_base = (int64_t)__shfl_sync(0xffffffffu, take(), 0);
    int tk_off = 0;
    int64_t nx_ai = tk_base;
    int64_t nx_a = 0, nx_q0 = 0, nx_q1 = 0, nx_b0 = 0, nx_b1 = 0;
    int32_t nx_ad = 0;
    if (nx_ai < m.n_node) {
        nx_a = m.node_order ? (int64_t)m.node_order[nx_ai] : nx_ai;
        nx_q0 = m.adj_ptr[nx_a]; nx_q1 = m.adj_ptr[nx_a + 1]; nx_b0 = m.brow[nx_a]; nx_b1 = m.brow[nx_a + 1];
        nx_ad = (nx_q0 + lane < nx_q1 && lane < CH) ? m.adj[nx_q0 + lane] : 0;
    }
    for (;;) {
        if (nx_ai >= m.n_node) break;
        const int64_t a = nx_a, q0 = nx_q0, q1 = nx_q1, b0 = nx_b0;
        const int cnt = (int)(nx_b1 - nx_b0);
        const int32_t ad_first = nx_ad;
        const bool need_tk = (++tk_off == TG);                   // next node: a new ticket batch, taken now and consumed below
        const unsigned long long tk = need_tk ? take() : 0ULL;
        const int rowlen = cnt * NF;
        const unsigned j0_bytes = (jac_a && !J0D) ? (unsigned)(sizeof(double) * DIM * NF) * (unsigned)cnt : 0u;
        const double* j0g = j0 + b0 * (DIM * NF);
        if (J0D && jac_a) {
            const int nline = (cnt * DIM * NF * (int)sizeof(double) + 127) >> 7;
            for (int i = lane; i < nline; i += 32) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(j0g + i * 16));
        }
        __syncwarp();                                            // the previous node's output stage has read j0s / acc
        if (want_jac) {
            double2* z = reinterpret_cast<double2*>(acc);
            const int nz = (JP * accn) >> 1;
            for (int i = lane; i < nz; i += 32) z[i] = make_double2(0.0, 0.0);
        }
        if (q0 >= q1) {                                          // unreferenced node: no records will be waited for
            if (need_tk) { tk_base = (int64_t)__shfl_sync(0xffffffffu, tk, 0); tk_off = 0; }
                nx_ai = tk_base + tk_off;
            if (nx_ai < m.n_node) {
                nx_a = m.node_order ? (int64_t)m.node_order[nx_ai] : nx_ai;
                nx_q0 = m.adj_ptr[nx_a]; nx_q1 = m.adj_ptr[nx_a + 1]; nx_b0 = m.brow[nx_a]; nx_b1 = m.brow[nx_a + 1];
            }
        }
        double fs = 0.0, vsum = 0.0;                             // lane (jj, k < NF): signed flux component k; SCV volumes
        int self_slot = 0;
        for (int64_t qb = q0; qb < q1; qb += CH) {
            const int nj = (int)((q1 - qb) < CH ? (q1 - qb) : CH);
            const int nrec = nj * NINC;
            __syncwarp();
            const bool first = (qb == q0);
            const int32_t ad = first ? ad_first : ((lane < nj) ? m.adj[qb + lane] : 0);
            const int e_l = ad / NSH, la_l = ad - e_l * NSH;
            const int rj = lane / NINC, rt = lane - rj * NINC;
            const int e_r = __shfl_sync(0xffffffffu, e_l, rj < CH ? rj : 0);
            const int la_r = __shfl_sync(0xffffffffu, la_l, rj < CH ? rj : 0);
            const int ipx_r = inctab[la_r * NINC + rt];
            const int64_t gi_r = (int64_t)e_r * NIP + (ipx_r & 255);
            if (lane < nrec) ws.ipx[lane] = ipx_r;
            // asynchronous staging: one TMA bulk copy per incident SCVF record (and the node's J0 rows), all in flight at once
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            if (lane == 0) mbar_arrive_expect_tx(&ws.bar, cp_bytes * (unsigned)nrec + (first ? j0_bytes : 0u));
            __syncwarp();
            if (l2hint) {
                if (lane < nrec) bulk_g2s_hint(&ws.rec[lane][0], rec + gi_r * RS, cp_bytes, &ws.bar, pol_rec);
                if (first && j0_bytes && lane == 31) bulk_g2s_hint(j0s, j0g, j0_bytes, &ws.bar, pol_j0);
            } else {
                if (lane < nrec) bulk_g2s(&ws.rec[lane][0], rec + gi_r * RS, cp_bytes, &ws.bar);
                if (first && j0_bytes && lane == 31) bulk_g2s(j0s, j0g, j0_bytes, &ws.bar);
            }
            // scatter slots + the node's SCV volume in the adjacent elements: loaded now, used after the records have landed
            uint2 emv = make_uint2(0u, 0u);
            double vv = 0.0;
            if (lane < nj) {
                const uint8_t* em = m.emap + (int64_t)ad * NSH;
                if (NSH == 8) emv = __ldg(reinterpret_cast<const uint2*>(em));
                else if (NSH == 4) emv.x = __ldg(reinterpret_cast<const uint32_t*>(em));
                else { for (int q = 0; q < NSH; q++) emv.x |= (uint32_t)em[q] << (8 * q); }
                vv = m.scvvol[ad];
            }
            const int sslot = __shfl_sync(0xffffffffu, la_l, 0);
            if (first) {                                         // next node: ticket arrived -> load its ranges while the records fly
                if (need_tk) { tk_base = (int64_t)__shfl_sync(0xffffffffu, tk, 0); tk_off = 0; }
                nx_ai = tk_base + tk_off;
                if (nx_ai < m.n_node) {
                    nx_a = m.node_order ? (int64_t)m.node_order[nx_ai] : nx_ai;
                    nx_q0 = m.adj_ptr[nx_a]; nx_q1 = m.adj_ptr[nx_a + 1]; nx_b0 = m.brow[nx_a]; nx_b1 = m.brow[nx_a + 1];
                }
            }
            __syncwarp();
            mbar_wait(&ws.bar, phase);
            phase ^= 1u;
            if (lane < nj) {
                if (NSH == 8) *reinterpret_cast<uint2*>(ws.slot[lane]) = emv;
                else *reinterpret_cast<uint32_t*>(ws.slot[lane]) = emv.x;
            }
            vsum += vv;
            __syncwarp();
            if (first) self_slot = ws.slot[0][sslot];
            for (int jb = 0; jb < nj; jb += JP) {
                const int j = jb + jj;
                if (lane_on && j < nj) {
                    double D = 0.0, PP = 0.0, Cn[DIM];
#pragma unroll
                    for (int d = 0; d < DIM; d++) Cn[d] = 0.0;
#pragma unroll
                    for (int t = 0; t < NINC; t++) {
                        const int r = j * NINC + t;
                        const double* rc = ws.rec[r];
                        const int ipx = ws.ipx[r];
                        const bool neg = ipx & 256;
                        if (def_a && k < NF) { const double f = rc[LR::O_F + k]; fs += neg ? -f : f; }
                        if constexpr (COMP) {
                            if (jac_a) {
                                const double* tq = tab4 + (ipx & 255) * CR::TSTR + 4 * k;
                                const double2 t0 = *reinterpret_cast<const double2*>(tq), t1 = *reinterpret_cast<const double2*>(tq + 2);
                                const double2 r2 = *reinterpret_cast<const double2*>(rc + 4), r3 = *reinterpret_cast<const double2*>(rc + 6);
                                const double2 r4 = *reinterpret_cast<const double2*>(rc + 8), r5 = *reinterpret_cast<const double2*>(rc + 10);
                                const double2 r6 = *reinterpret_cast<const double2*>(rc + 12);
                                const double upk = rc[CR::O_UP + k];
                                const double sg = neg ? -p.scale_a : p.scale_a;
                                const double cK = r3.y * t1.y + r4.x * upk;           // alpha N_k + beta up_k
                                const double dK = upk * r4.y + r5.x * t1.y;           // cw up_k + cpe N_k
                                double pK = t0.x * r5.y;                               // dN_k . mv
                                pK += t0.y * r6.x; pK += t1.x * r6.y;
                                D += sg * dK; PP += sg * pK;
                                const double wv = sg * cK;
                                Cn[0] += wv * r2.x; Cn[1] += wv * r2.y; Cn[DIM - 1] += wv * r3.x;
                            }
                        } else if (jac_a) {
                            const double sg = neg ? -p.scale_a : p.scale_a;
                            D += sg * rc[LR::O_DK + k];
                            PP += sg * rc[LR::O_PK + k];
                            const double w = sg * rc[LR::O_CK + k];
#pragma unroll
                            for (int d = 0; d < DIM; d++) Cn[d] += w * rc[LR::O_N + d];
                        }
                    }
                    if (jac_a) {
                        const int slot = ws.slot[j][k];
                        accj[slot] += D;
#pragma unroll
                        for (int d = 0; d < DIM; d++) accj[(1 + d) * cntp + slot] += Cn[d];
                        accj[(1 + DIM) * cntp + slot] += PP;
                    }
                }
            }
        }
        __syncwarp();
        // deterministic reductions: SCV volume of the node (butterfly), defect fluxes (lane q < NF sums its component over jj)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        const double volsum = vsum;
        double dsum = 0.0;
        if (def_a) {
#pragma unroll
            for (int j2 = 0; j2 < JP; j2++) dsum += __shfl_sync(0xffffffffu, fs, j2 * NSH + (lane < NF ? lane : 0));
        }
        nx_ad = (nx_ai < m.n_node && nx_q0 + lane < nx_q1 && lane < CH) ? m.adj[nx_q0 + lane] : 0;   // used at the top of the next node
        if (want_jac) {
            // merge the JP accumulator copies (fixed order) into copy 0, add the lumped mass (add_jac_M_elem :781-808)
            for (int i = lane; i < accn; i += 32) {
                double sacc = acc[i];
#pragma unroll
                for (int c = 1; c < JP; c++) sacc += acc[c * accn + i];
                if ((what & W_JAC_M) && i == self_slot) sacc += p.scale_m * volsum * p.rho;
                acc[i] = sacc;
            }
            __syncwarp();
            double* out = val + b0 * (NF * NF);
            if constexpr (NF == 4) {
                // the NF rows of the node are contiguous (in the CSR values and in the staged J0 rows): one loop over all
                // 16-byte chunks; chunk i = row rf, slot, column pair cp
                const int n2 = 2 * cnt;
                double2* o2 = reinterpret_cast<double2*>(out);
                const double2* j2 = reinterpret_cast<const double2*>(J0D ? j0g : j0s);
                for (int i = lane; i < NF * n2; i += 32) {
                    const int rf = (i >= n2) + (i >= 2 * n2) + (i >= 3 * n2);
                    const int ii = i - rf * n2, slot = ii >> 1, cp = ii & 1;
                    double2 v;
                    if (rf < DIM) {
                        v = make_double2(0.0, 0.0);
                        if (jac_a) { const double2 jv = J0D ? __ldcs(j2 + i) : j2[i]; v.x = jv.x * s_visc; v.y = jv.y * (cp ? s_pres : s_visc); }
                        const double D = acc[slot];
                        if (rf == 2 * cp) v.x += D;
                        if (rf == 2 * cp + 1) v.y += D;
                    } else {
                        v = make_double2(acc[(1 + 2 * cp) * cntp + slot], acc[(2 + 2 * cp) * cntp + slot]);
                    }
                    if (beta == 0.0) __stcs(o2 + i, v);
                    else { double2 o = o2[i]; o.x = beta * o.x + v.x; o.y = beta * o.y + v.y; o2[i] = o; }
                }
            } else {
                for (int rf = 0; rf < DIM; rf++) {
                    double* orow = out + rf * rowlen;
                    for (int i = lane; i < rowlen; i += 32) {
                        const int slot = i / NF, cf = i - slot * NF;
                        double v = jac_a ? (J0D ? __ldcs(j0g + rf * rowlen + i) : j0s[rf * rowlen + i]) * (cf < DIM ? s_visc : s_pres) : 0.0;
                        if (cf == rf) v += acc[slot];
                        if (beta == 0.0) __stcs(orow + i, v);
                        else orow[i] = beta * orow[i] + v;
                    }
                }
                double* orow = out + DIM * rowlen;
                for (int i = lane; i < rowlen; i += 32) {
                    const int slot = i / NF, cf = i - slot * NF;
                    const double v = acc[(1 + cf) * cntp + slot];
                    if (beta == 0.0) __stcs(orow + i, v);
                    else orow[i] = beta * orow[i] + v;
                }
            }
        }
        if (want_def && lane < NF) {
            double d = def_a ? dsum : 0.0;
            if ((what & W_RHS) && p.has_source && lane < DIM) d -= p.src[lane] * volsum * p.rho;
            d *= p.scale_a;
            if ((what & W_DEF_M) && lane < DIM) d += p.scale_m * u[a * NF + lane] * volsum * p.rho;
            double* q = def + a * NF + lane;
            *q = (beta == 0.0) ? d : beta * (*q) + d;
        }
    }
}

// ---- rows kernel of the split path, owner-lane form (3-D element types, compressed records, block rows up to 32 slots) ----
// warp per node (atomic tickets, TG nodes per ticket, their headers loaded as one batch). Per round of JP adjacent elements:
// one TMA bulk copy per incident compressed record on a warp-private mbarrier; lane = (element jj, corner k) forms its 5
// values (D, C[3], PP) from the NINC records and parks them in a per-warp staging row; then lane = COLUMN SLOT b of the block
// row picks up the values whose scatter slot is b (slot -> corner table written by the compute lanes) and accumulates them in
// REGISTERS in fixed order: no shared-memory accumulators, no copies to merge, no zero fill. The lane then owns the 4x4 block
// (node, b): out = {nu rho, 1} scale_a J0 + state part, J0 read straight from global memory into registers at the top of the
// node. 3.9 KB of shared memory per warp instead of 11 KB.
template <int E> struct OwnWS {
    static constexpr int NSH = ET<E>::NSH, NINC = ET<E>::NINC, JP = 32 / NSH, NREC = JP * NINC, NV = ET<E>::DIM + 2;
    static constexpr int RS = CompRec<E>::SZ, SS = rows_smem_stride<E>(RS);
    alignas(16) double rec[2][NREC][SS];           // two rounds in flight
    alignas(16) double stage[NV * 32];
    unsigned long long bar[2];
    int32_t ipx[2][NREC];
    alignas(4) uint8_t idx[JP * 32];
    alignas(8) uint8_t slot[2][JP][8];
};

template <int E, int MINB = 12, bool FAST = false>
__global__ void __launch_bounds__(64, MINB) fv1_rows_owner_kernel(KParams p, MeshDev m, const double* __restrict__ rec,
                                                                const double* __restrict__ j0, const double* __restrict__ u,
                                                                double beta, double* __restrict__ val, double* __restrict__ def,
                                                                unsigned long long* __restrict__ work_counter)
{
    using WS = OwnWS<E>;
    using CR = CompRec<E>;
    constexpr int DIM = ET<E>::DIM, NSH = ET<E>::NSH, NF = DIM + 1, NINC = ET<E>::NINC, NIP = ET<E>::NIP, JP = WS::JP, RS = WS::RS;
    static_assert(DIM == 3, "compressed records: 3-D element types");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t* inctab = reinterpret_cast<int32_t*>(smem_raw);
    double* tab4 = reinterpret_cast<double*>(smem_raw + ((sizeof(int32_t) * NSH * NINC + 15) & ~(size_t)15));
    constexpr size_t tab_bytes = split_tab_bytes<E>();
    WS& ws = *reinterpret_cast<WS*>(smem_raw + tab_bytes + warp * ((sizeof(WS) + 15) & ~(size_t)15));
    for (int i = threadIdx.x; i < NSH * NINC; i += blockDim.x)
        inctab[i] = tab::INC[E][i / NINC][i % NINC] | (tab::INC_SIGN[E][i / NINC][i % NINC] < 0 ? 256 : 0);
    for (int i = threadIdx.x; i < NIP * NSH * 4; i += blockDim.x) {
        const int ip = i / (NSH * 4), kk = (i >> 2) % NSH, c = i & 3;
        tab4[ip * CR::TSTR + kk * 4 + c] = c < 3 ? tab::C_DNIP[E][ip][kk][c < DIM ? c : 0] : tab::NIPSH[E][ip][kk];
    }
    if (lane == 0) { mbar_init(&ws.bar[0], 1); mbar_init(&ws.bar[1], 1); }
    __syncthreads();
    unsigned phase0 = 0, phase1 = 0;
    const int what = FAST ? (W_JAC_A | W_DEF_A) : p.what;
    if (FAST) beta = 0.0;
    const bool want_jac = what & (W_JAC_A | W_JAC_M), want_def = what & (W_DEF_A | W_DEF_M | W_RHS);
    const bool jac_a = what & W_JAC_A, def_a = what & W_DEF_A, flux_needed = jac_a || def_a;
    const bool need_vol = what & (W_JAC_M | W_DEF_M | W_RHS);
    const unsigned cp_bytes = (unsigned)sizeof(double) * (jac_a ? RS : CR::HEAD);
    const int jj = lane / NSH, k = lane - jj * NSH;
    const double sa = p.scale_a, s_visc = p.visc * p.rho * p.scale_a, s_pres = p.scale_a;
    int TG = m.ticket_group;                                     // nodes per ticket: 1, 2, 4 or 8
    TG = TG >= 8 ? 8 : (TG >= 4 ? 4 : (TG >= 2 ? 2 : 1));
    const int EPN = 32 / TG;                                     // batched adjacency entries per node
    double* stg = ws.stage;
    uint8_t* idx = ws.idx;

    for (;;) {
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(work_counter, (unsigned long long)TG);
        const int64_t base = (int64_t)__shfl_sync(0xffffffffu, tk, 0);
        if (base >= m.n_node) break;
        // headers of the TG nodes of this ticket, one per lane
        int64_t h_a = 0, h_q0 = 0, h_b0 = 0;
        int h_nadj = 0, h_cnt = 0;
        if (lane < TG && base + lane < m.n_node) {
            h_a = m.node_order ? (int64_t)m.node_order[base + lane] : base + lane;
            h_q0 = m.adj_ptr[h_a]; h_nadj = (int)(m.adj_ptr[h_a + 1] - h_q0); h_b0 = m.brow[h_a]; h_cnt = (int)(m.brow[h_a + 1] - h_b0);
        }
        // adjacency entries, slot maps and SCV volumes of ALL nodes of the ticket in one batch (lane = (node, entry)) when every
        // node has at most EPN adjacent elements; otherwise they are loaded round by round
        const bool batched = __all_sync(0xffffffffu, h_nadj <= EPN);
        int32_t b_ad = 0; uint2 b_em = make_uint2(0u, 0u); double b_vol = 0.0;
        if (batched) {
            const int tn = lane / EPN, jx = lane - tn * EPN;
            const int64_t tq0 = __shfl_sync(0xffffffffu, h_q0, tn);
            const int tna = __shfl_sync(0xffffffffu, h_nadj, tn);
            if (jx < tna) {
                b_ad = m.adj[tq0 + jx];
                const uint8_t* em = m.emap + (int64_t)b_ad * NSH;
                if (NSH == 8) b_em = __ldg(reinterpret_cast<const uint2*>(em));
                else b_em.x = __ldg(reinterpret_cast<const uint32_t*>(em));
                if (need_vol) b_vol = m.scvvol[b_ad];
            }
        }
        for (int ti = 0; ti < TG && base + ti < m.n_node; ti++) {
            const int64_t a = __shfl_sync(0xffffffffu, h_a, ti), q0 = __shfl_sync(0xffffffffu, h_q0, ti), b0 = __shfl_sync(0xffffffffu, h_b0, ti);
            const int nadj = __shfl_sync(0xffffffffu, h_nadj, ti), cnt = __shfl_sync(0xffffffffu, h_cnt, ti);
            const bool own = lane < cnt;                         // lane = column slot b of the block row
            const double2* j0g = reinterpret_cast<const double2*>(j0 + b0 * (DIM * NF));
            double2 jv[DIM][2];
            if (jac_a && own) {
#pragma unroll
                for (int rf = 0; rf < DIM; rf++) { jv[rf][0] = __ldcs(j0g + (rf * cnt + lane) * 2); jv[rf][1] = __ldcs(j0g + (rf * cnt + lane) * 2 + 1); }
            }
            double aD = 0.0, aC[DIM] = {0.0, 0.0, 0.0}, aP = 0.0, fs = 0.0, vsum = 0.0;
            int self_slot = -1;
            const int nround = (nadj + JP - 1) / JP;
            // round r: lane < nj holds adjacency entry r * JP + lane; the slot maps / record copies go to buffer r & 1
            auto issue = [&](int r) {
                const int buf = r & 1;
                const int nj = (nadj - r * JP) < JP ? (nadj - r * JP) : JP;
                const int nrec = nj * NINC;
                int32_t ad = 0; uint2 emv = make_uint2(0u, 0u); double vv = 0.0;
                if (batched) {
                    const int srcl = ti * EPN + r * JP + (lane < JP ? lane : 0);
                    ad = __shfl_sync(0xffffffffu, b_ad, srcl); emv.x = __shfl_sync(0xffffffffu, b_em.x, srcl);
                    if (NSH == 8) emv.y = __shfl_sync(0xffffffffu, b_em.y, srcl);
                    if (need_vol) vv = __shfl_sync(0xffffffffu, b_vol, srcl);
                } else if (lane < nj) {
                    ad = m.adj[q0 + r * JP + lane];
                    const uint8_t* em = m.emap + (int64_t)ad * NSH;
                    if (NSH == 8) emv = __ldg(reinterpret_cast<const uint2*>(em));
                    else emv.x = __ldg(reinterpret_cast<const uint32_t*>(em));
                    if (need_vol) vv = m.scvvol[ad];
                }
                if (lane < nj) vsum += vv;
                const int e_l = ad / NSH, la_l = ad - e_l * NSH;
                if (r == 0) {                                    // slot of the node itself in its block row
                    const int sla = __shfl_sync(0xffffffffu, la_l, 0);
                    const uint32_t w0 = __shfl_sync(0xffffffffu, emv.x, 0), w1 = __shfl_sync(0xffffffffu, emv.y, 0);
                    self_slot = (int)(((sla < 4 ? w0 : w1) >> (8 * (sla & 3))) & 255u);
                }
                const int rj = lane / NINC, rt = lane - rj * NINC;
                const int e_r = __shfl_sync(0xffffffffu, e_l, rj < JP ? rj : 0);
                const int la_r = __shfl_sync(0xffffffffu, la_l, rj < JP ? rj : 0);
                const int ipx_r = inctab[la_r * NINC + rt];
                const int64_t gi_r = (int64_t)e_r * NIP + (ipx_r & 255);
                if (lane < nrec) ws.ipx[buf][lane] = ipx_r;
                if (lane < nj) {
                    if (NSH == 8) *reinterpret_cast<uint2*>(ws.slot[buf][lane]) = emv;
                    else *reinterpret_cast<uint32_t*>(ws.slot[buf][lane]) = emv.x;
                }
                if (flux_needed) {
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    if (lane == 0) mbar_arrive_expect_tx(&ws.bar[buf], cp_bytes * (unsigned)nrec);
                    __syncwarp();
                    if (lane < nrec) bulk_g2s(&ws.rec[buf][lane][0], rec + gi_r * RS, cp_bytes, &ws.bar[buf]);
                } else __syncwarp();
            };
            __syncwarp();                                        // the previous node has consumed every buffer
            if (nround > 0) issue(0);
            if (nround > 1) issue(1);
            for (int r = 0; r < nround; r++) {
                const int buf = r & 1;
                const int nj = (nadj - r * JP) < JP ? (nadj - r * JP) : JP;
                if (flux_needed) {
                    if (buf == 0) { mbar_wait(&ws.bar[0], phase0); phase0 ^= 1u; } else { mbar_wait(&ws.bar[1], phase1); phase1 ^= 1u; }
                }
                double D = 0.0, PP = 0.0, Cn[DIM] = {0.0, 0.0, 0.0};
                const bool on = jj < nj;
                if (on && flux_needed) {
#pragma unroll
                    for (int t = 0; t < NINC; t++) {
                        const int rr = jj * NINC + t;
                        const double* rc = ws.rec[buf][rr];
                        const int ipx = ws.ipx[buf][rr];
                        const bool neg = ipx & 256;
                        if (def_a && k < NF) { const double f = rc[CR::O_F + k]; fs += neg ? -f : f; }
                        if (jac_a) {
                            const double* tq = tab4 + (ipx & 255) * CR::TSTR + 4 * k;
                            const double2 t0 = *reinterpret_cast<const double2*>(tq), t1 = *reinterpret_cast<const double2*>(tq + 2);
                            const double2 r2 = *reinterpret_cast<const double2*>(rc + 4), r3 = *reinterpret_cast<const double2*>(rc + 6);
                            const double2 r4 = *reinterpret_cast<const double2*>(rc + 8), r5 = *reinterpret_cast<const double2*>(rc + 10);
                            const double2 r6 = *reinterpret_cast<const double2*>(rc + 12);
                            const double upk = rc[CR::O_UP + k];
                            const double sg = neg ? -sa : sa;
                            const double cK = r3.y * t1.y + r4.x * upk;           // alpha N_k + beta up_k
                            const double dK = upk * r4.y + r5.x * t1.y;           // cw up_k + cpe N_k
                            double pK = t0.x * r5.y;                               // dN_k . mv
                            pK += t0.y * r6.x; pK += t1.x * r6.y;
                            D += sg * dK; PP += sg * pK;
                            const double wv = sg * cK;
                            Cn[0] += wv * r2.x; Cn[1] += wv * r2.y; Cn[2] += wv * r3.x;
                        }
                    }
                }
                if (jac_a) {
#pragma unroll
                    for (int i = lane; i < JP * 8; i += 32) reinterpret_cast<uint32_t*>(idx)[i] = 0xffffffffu;
                    __syncwarp();
                    if (on) idx[jj * 32 + ws.slot[buf][jj][k]] = (uint8_t)k;
                    stg[lane] = D; stg[32 + lane] = Cn[0]; stg[64 + lane] = Cn[1]; stg[96 + lane] = Cn[2]; stg[128 + lane] = PP;
                    __syncwarp();
#pragma unroll
                    for (int j2 = 0; j2 < JP; j2++) {
                        const int kk = idx[j2 * 32 + lane];
                        if (kk != 255) {
                            const int src = j2 * NSH + kk;
                            aD += stg[src]; aC[0] += stg[32 + src]; aC[1] += stg[64 + src]; aC[2] += stg[96 + src]; aP += stg[128 + src];
                        }
                    }
                }
                __syncwarp();                                    // rec / ipx / slot of this buffer and the staging row are free again
                if (r + 2 < nround) issue(r + 2);
            }
            double volsum = 0.0;
            if (need_vol) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
                volsum = vsum;
            }
            if ((what & W_JAC_M) && lane == self_slot) aD += p.scale_m * volsum * p.rho;   // lumped mass (add_jac_M_elem :781-808)
            if (want_jac && own) {
                double2* o2 = reinterpret_cast<double2*>(val + b0 * (NF * NF));
#pragma unroll
                for (int rf = 0; rf < DIM; rf++) {
                    double2 v0 = make_double2(0.0, 0.0), v1 = make_double2(0.0, 0.0);
                    if (jac_a) { v0.x = jv[rf][0].x * s_visc; v0.y = jv[rf][0].y * s_visc; v1.x = jv[rf][1].x * s_visc; v1.y = jv[rf][1].y * s_pres; }
                    if (rf == 0) v0.x += aD; else if (rf == 1) v0.y += aD; else v1.x += aD;
                    double2* o = o2 + (rf * cnt + lane) * 2;
                    if (beta == 0.0) { __stcs(o, v0); __stcs(o + 1, v1); }
                    else { double2 x0 = o[0], x1 = o[1]; x0.x = beta * x0.x + v0.x; x0.y = beta * x0.y + v0.y; x1.x = beta * x1.x + v1.x; x1.y = beta * x1.y + v1.y; o[0] = x0; o[1] = x1; }
                }
                double2* o = o2 + (DIM * cnt + lane) * 2;
                const double2 v0 = make_double2(aC[0], aC[1]), v1 = make_double2(aC[2], aP);
                if (beta == 0.0) { __stcs(o, v0); __stcs(o + 1, v1); }
                else { double2 x0 = o[0], x1 = o[1]; x0.x = beta * x0.x + v0.x; x0.y = beta * x0.y + v0.y; x1.x = beta * x1.x + v1.x; x1.y = beta * x1.y + v1.y; o[0] = x0; o[1] = x1; }
            }
            if (want_def) {
                double dsum = 0.0;
                if (def_a) {
#pragma unroll
                    for (int j2 = 0; j2 < JP; j2++) dsum += __shfl_sync(0xffffffffu, fs, j2 * NSH + (lane < NF ? lane : 0));
                }
                if (lane < NF) {
                    double d = def_a ? dsum : 0.0;
                    if ((what & W_RHS) && p.has_source && lane < DIM) d -= p.src[lane] * volsum * p.rho;
                    d *= p.scale_a;
                    if ((what & W_DEF_M) && lane < DIM) d += p.scale_m * u[a * NF + lane] * volsum * p.rho;
                    double* q = def + a * NF + lane;
                    *q = (beta == 0.0) ? d : beta * (*q) + d;
                }
            }
        }
    }
}

}  // namespace nsb
