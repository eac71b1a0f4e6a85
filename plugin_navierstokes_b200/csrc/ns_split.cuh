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
template <int E, int CHP = 0> struct SplitWS {
    using C = SplitCfg<E, CHP>;
    static constexpr int RS = LeanRec<E>::SZ, SS = rows_smem_stride<E>(RS);
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
    constexpr size_t tab_bytes = (sizeof(int32_t) * NSH * NINC + 15) & ~(size_t)15;
    const int cntp = split_cnt_pad(m.max_cnt);
    const int accn = NV * cntp;                                  // doubles per accumulator copy
    const size_t per_warp = split_warp_bytes<E, CHP, J0D>(m.max_cnt);
    WS& ws = *reinterpret_cast<WS*>(smem_raw + tab_bytes + warp * per_warp);
    double* j0s = reinterpret_cast<double*>(smem_raw + tab_bytes + warp * per_warp + sizeof(WS));
    double* acc = j0s + (J0D ? 0 : (size_t)DIM * NF * m.max_cnt);
    for (int i = threadIdx.x; i < NSH * NINC; i += blockDim.x)
        inctab[i] = tab::INC[E][i / NINC][i % NINC] | (tab::INC_SIGN[E][i / NINC][i % NINC] < 0 ? 256 : 0);
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
    const unsigned cp_bytes = (unsigned)sizeof(double) * (jac_a ? RS : LR::HEAD);
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
                        const bool neg = ws.ipx[r] & 256;
                        if (def_a && k < NF) { const double f = rc[LR::O_F + k]; fs += neg ? -f : f; }
                        if (jac_a) {
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

}  // namespace nsb
