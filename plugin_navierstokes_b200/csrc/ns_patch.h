// ns_patch.h -- host side of the fused patch kernel (ns_fused.cuh): partition of the grid nodes into PATCHES and the
// per-patch tables the kernel consumes. Pure C++ (no CUDA), also compiled by the CPU emulation harness in tests/cpp.
//
// A patch is a spatially compact set of grid nodes whose CSR rows one CTA assembles. The CTA evaluates every SCVF
// (element e, integration point ip) that touches a patch node ONCE into a shared-memory record slot and then sums,
// for every patch node, the records of its incident SCVFs into the node's rows (owner-computes, no atomics).
// SCVFs that join nodes of two patches are evaluated by both (redundancy 1 + 1/p per direction for a p-node edge).
//
// Tables per patch (all indices local to the patch, so they fit 8/16 bits):
//   nodes[]  : global node id, first value index of its block row, row length, range of its adjacency entries
//   elems[]  : global element ids of the elements touching a patch node (ascending)
//   pconn[]  : global node ids of their corners (saves one dependent load in the kernel)
//   work[]   : one entry per SCVF to evaluate: local element | ip << 8 | record slot << 12, sorted by ip (a warp of the flux
//              phase then reads the reference tables of one or two ips)
//   lnodes[] / ecorner[] (tile kernel only): the patch's local nodes and the element -> local node table
//   adj[]    : per (patch node, adjacent element), in the order of the global adjacency list (ascending element id =
//              the summation order of every owner-computes kernel in this library): record slots of the NINC incident
//              SCVFs, local corner, CSR slots of the element's corners in the node's block row
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace nsb {

struct PatchHdr { int32_t node0, n_node, elem0, n_elem, work0, n_work, adj0, n_adj, lnode0, n_lnode, pad0, pad1; };   // 48 B
struct PatchNode { int64_t b0; int32_t node; uint16_t adj_off; uint8_t adj_cnt, cnt; };          // 16 B
struct PatchAdj { uint16_t slot[3]; uint8_t la, self; uint8_t emap[8]; };                        // 16 B; la: bits 0-2 local corner, bit 4+t: SCVF t enters the node (sign -1)
static_assert(sizeof(PatchHdr) == 48 && sizeof(PatchNode) == 16 && sizeof(PatchAdj) == 16, "table layouts are read as 16-byte words on the device");

struct PatchCaps {
    int max_work, max_elem, max_node, max_adj;
    int max_lnode = 0;                 // > 0: build the local-node tables (lnodes / ecorner) of the tile kernel, at most this many per patch (<= 256)
    int tile[3];                       // tile edge lengths in grid cells (estimated spacing) the nodes are binned into
};

struct PatchPlan {
    std::vector<PatchHdr> hdr;
    std::vector<PatchNode> nodes;
    std::vector<int32_t> elems, pconn;
    std::vector<int32_t> lnodes;       // tile kernel (ns_tile.cuh): global ids of the patch's local nodes (every corner of a patch element), ascending
    std::vector<uint8_t> ecorner;      // ... and the local node id of corner k of patch element el, [elems][nsh]
    std::vector<uint32_t> work;
    std::vector<PatchAdj> adj;
    int64_t n_scvf_evals = 0;          // = work.size(); / (n_elem * nip) = redundancy of the flux phase
    int max_adj_per_node = 0;
};

namespace patch_detail {
struct ElemTopo { int nsh, dim, nip, ninc; int edge[12][2]; };
inline const ElemTopo& topo(int elem)
{
    // SURVEY.md App. B-1 (edge = SCVF numbering); identical to tab::EDGE / tab::INC of ref_tables.cuh (checked by the emulator test)
    static const ElemTopo T[4] = {
        {3, 2, 3, 2, {{0, 1}, {1, 2}, {2, 0}}},
        {4, 2, 4, 2, {{0, 1}, {1, 2}, {2, 3}, {3, 0}}},
        {4, 3, 6, 3, {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}}},
        {8, 3, 12, 3, {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {0, 4}, {1, 5}, {2, 6}, {3, 7}, {4, 5}, {5, 6}, {6, 7}, {7, 4}}}};
    return T[elem];
}
template <class F> inline void par_for(int64_t n, F fn)
{
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (n < 64) nt = 1;
    std::vector<std::thread> th;
    const int64_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const int64_t lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([=]() { fn((int)t, lo, hi); });
    }
    for (auto& x : th) x.join();
}
}  // namespace patch_detail

// Builds the plan. adj_ptr / adj: node -> (element * nsh + local corner), ascending; brow: block-row prefix; emap: slot of
// corner k in the block row of corner a, [n_elem][nsh][nsh]. Returns false (with a message) when a single node exceeds the caps.
inline bool build_patch_plan(int elem, int64_t n_elem, int64_t n_node, const int32_t* conn, const double* coords,
                             const int64_t* adj_ptr, const int32_t* adj, const int64_t* brow, const uint8_t* emap,
                             const PatchCaps& caps, PatchPlan& out, std::string& err)
{
    using namespace patch_detail;
    const ElemTopo& T = topo(elem);
    const int nsh = T.nsh, dim = T.dim, nip = T.nip, ninc = T.ninc;
    // ips incident to corner la, ascending ip (= tab::INC)
    int inc[8][3];
    for (int la = 0; la < nsh; la++) { int c = 0; for (int ip = 0; ip < nip; ip++) if (T.edge[ip][0] == la || T.edge[ip][1] == la) inc[la][c++] = ip; }
    // ---- grid spacing per axis: mean |dx_d| over the element edges whose dominant direction is d ----
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, hs[3] = {0, 0, 0};
    int64_t hc[3] = {0, 0, 0};
    for (int64_t i = 0; i < n_node; i++) for (int d = 0; d < dim; d++) { lo[d] = std::min(lo[d], coords[i * dim + d]); hi[d] = std::max(hi[d], coords[i * dim + d]); }
    const int64_t estep = std::max<int64_t>(1, n_elem / 200000);
    for (int64_t e = 0; e < n_elem; e += estep) for (int ip = 0; ip < nip; ip++) {
        const int64_t a = conn[e * nsh + T.edge[ip][0]], b = conn[e * nsh + T.edge[ip][1]];
        double dd[3] = {0, 0, 0}; int dm = 0;
        for (int d = 0; d < dim; d++) { dd[d] = std::fabs(coords[a * dim + d] - coords[b * dim + d]); if (dd[d] > dd[dm]) dm = d; }
        hs[dm] += dd[dm]; hc[dm]++;
    }
    double h[3];
    for (int d = 0; d < dim; d++) { h[d] = hc[d] ? hs[d] / hc[d] : (hi[d] - lo[d]); if (!(h[d] > 0)) h[d] = 1.0; }
    // ---- bin the nodes into tiles of caps.tile cells; order: tile (x fastest), then cell (z, y, x) ----
    int64_t ntile[3] = {1, 1, 1};
    for (int d = 0; d < dim; d++) ntile[d] = (int64_t)std::floor((hi[d] - lo[d]) / h[d] + 0.5) / caps.tile[d] + 1;
    struct Key { uint64_t tile; uint32_t cell; int32_t node; };
    std::vector<Key> keys(n_node);
    par_for(n_node, [&](int, int64_t a, int64_t b) {
        for (int64_t i = a; i < b; i++) {
            int64_t c[3] = {0, 0, 0}, t[3] = {0, 0, 0};
            for (int d = 0; d < dim; d++) {
                c[d] = (int64_t)std::floor((coords[i * dim + d] - lo[d]) / h[d] + 0.5);
                t[d] = std::min(std::max<int64_t>(c[d] / caps.tile[d], 0), ntile[d] - 1);
                c[d] -= t[d] * caps.tile[d];
                c[d] = std::min<int64_t>(std::max<int64_t>(c[d], 0), 1023);
            }
            keys[i].tile = (uint64_t)((t[2] * ntile[1] + t[1]) * ntile[0] + t[0]);
            keys[i].cell = (uint32_t)((c[2] << 20) | (c[1] << 10) | c[0]);
            keys[i].node = (int32_t)i;
        }
    });
    std::sort(keys.begin(), keys.end(), [](const Key& x, const Key& y) { return x.tile != y.tile ? x.tile < y.tile : (x.cell != y.cell ? x.cell < y.cell : x.node < y.node); });
    std::vector<int64_t> gstart;                                   // tile groups
    for (int64_t i = 0; i < n_node; i++) if (i == 0 || keys[i].tile != keys[i - 1].tile) gstart.push_back(i);
    gstart.push_back(n_node);
    const int64_t ngroup = (int64_t)gstart.size() - 1;

    // ---- per thread: local plans over contiguous ranges of groups ----
    struct Local { PatchPlan p; std::string err; };
    const unsigned maxthreads = 16;
    std::vector<Local> locals(maxthreads);
    std::vector<int> used(maxthreads, 0);
    par_for(ngroup, [&](int tix, int64_t g0, int64_t g1) {
        used[tix] = 1;
        Local& L = locals[tix];
        std::vector<int32_t> mark_node(0), elocal(0);
        // stamps in hash-free form: small sorted vectors (patches are tiny)
        std::vector<int32_t> nodes, elems, stackbuf;
        std::function<void(std::vector<int32_t>&)> emit = [&](std::vector<int32_t>& nd) {
            if (!L.err.empty() || nd.empty()) return;
            // node set (sorted copy for membership tests; the patch keeps the spatial order of `nd`)
            std::vector<int32_t> sorted_nodes(nd);
            std::sort(sorted_nodes.begin(), sorted_nodes.end());
            auto in_patch = [&](int32_t n) { return std::binary_search(sorted_nodes.begin(), sorted_nodes.end(), n); };
            elems.clear();
            int64_t nadj = 0; int maxadj = 0;
            for (int32_t n : nd) {
                const int64_t q0 = adj_ptr[n], q1 = adj_ptr[n + 1];
                nadj += q1 - q0; maxadj = std::max<int>(maxadj, (int)(q1 - q0));
                for (int64_t q = q0; q < q1; q++) elems.push_back(adj[q] / nsh);
            }
            std::sort(elems.begin(), elems.end());
            elems.erase(std::unique(elems.begin(), elems.end()), elems.end());
            // SCVFs with at least one end in the patch
            int64_t nwork = 0;
            for (int32_t e : elems) for (int ip = 0; ip < nip; ip++)
                if (in_patch(conn[(int64_t)e * nsh + T.edge[ip][0]]) || in_patch(conn[(int64_t)e * nsh + T.edge[ip][1]])) nwork++;
            std::vector<int32_t> ln;                           // local nodes (tile kernel)
            if (caps.max_lnode > 0) {
                for (int32_t e : elems) for (int k = 0; k < nsh; k++) ln.push_back(conn[(int64_t)e * nsh + k]);
                std::sort(ln.begin(), ln.end());
                ln.erase(std::unique(ln.begin(), ln.end()), ln.end());
            }
            const bool fits = nwork <= caps.max_work && (int64_t)elems.size() <= caps.max_elem && (int64_t)nd.size() <= caps.max_node &&
                              nadj <= caps.max_adj && maxadj <= 255 && (int64_t)ln.size() <= std::min(caps.max_lnode > 0 ? caps.max_lnode : 256, 256);
            if (!fits) {
                if (nd.size() == 1) { L.err = "a single node exceeds the patch capacities (valence too high for the fused kernel)"; return; }
                // split at the median of the axis with the largest extent
                double blo[3] = {1e300, 1e300, 1e300}, bhi[3] = {-1e300, -1e300, -1e300};
                for (int32_t n : nd) for (int d = 0; d < dim; d++) { blo[d] = std::min(blo[d], coords[(int64_t)n * dim + d]); bhi[d] = std::max(bhi[d], coords[(int64_t)n * dim + d]); }
                int ax = 0; for (int d = 1; d < dim; d++) if ((bhi[d] - blo[d]) / h[d] > (bhi[ax] - blo[ax]) / h[ax]) ax = d;
                std::vector<int32_t> a(nd), b;
                std::stable_sort(a.begin(), a.end(), [&](int32_t x, int32_t y) { return coords[(int64_t)x * dim + ax] < coords[(int64_t)y * dim + ax]; });
                // cut between two distinct coordinate layers when possible (keeps structured tiles box-shaped)
                size_t mid = a.size() / 2;
                { size_t up = mid; const double eps = 1e-9 * h[ax];
                  while (up < a.size() && std::fabs(coords[(int64_t)a[up] * dim + ax] - coords[(int64_t)a[up - 1] * dim + ax]) <= eps) up++;
                  size_t dn = mid;
                  while (dn > 0 && dn < a.size() && std::fabs(coords[(int64_t)a[dn] * dim + ax] - coords[(int64_t)a[dn - 1] * dim + ax]) <= eps) dn--;
                  if (up < a.size() && (up - mid <= mid - dn || dn == 0)) mid = up; else if (dn > 0) mid = dn; }
                b.assign(a.begin() + mid, a.end()); a.resize(mid);
                emit(a); emit(b);
                return;
            }
            PatchPlan& P = L.p;
            PatchHdr H;
            H.node0 = (int32_t)P.nodes.size(); H.n_node = (int32_t)nd.size();
            H.elem0 = (int32_t)P.elems.size(); H.n_elem = (int32_t)elems.size();
            H.work0 = (int32_t)P.work.size(); H.adj0 = (int32_t)P.adj.size();
            H.lnode0 = (int32_t)P.lnodes.size(); H.n_lnode = (int32_t)ln.size(); H.pad0 = H.pad1 = 0;
            P.lnodes.insert(P.lnodes.end(), ln.begin(), ln.end());
            // work items, slot = running index; slot of (local element, ip) for the adjacency table
            std::vector<uint16_t> slot_of(elems.size() * nip, 0xffff);
            uint32_t nw = 0;
            for (size_t el = 0; el < elems.size(); el++) {
                const int64_t e = elems[el];
                P.elems.push_back((int32_t)e);
                for (int k = 0; k < nsh; k++) {
                    P.pconn.push_back(conn[e * nsh + k]);
                    if (caps.max_lnode > 0) P.ecorner.push_back((uint8_t)(std::lower_bound(ln.begin(), ln.end(), conn[e * nsh + k]) - ln.begin()));
                }
            }
            for (int ip = 0; ip < nip; ip++)
                for (size_t el = 0; el < elems.size(); el++) {
                    const int64_t e = elems[el];
                    if (in_patch(conn[e * nsh + T.edge[ip][0]]) || in_patch(conn[e * nsh + T.edge[ip][1]])) {
                        slot_of[el * nip + ip] = (uint16_t)nw;
                        P.work.push_back((uint32_t)el | ((uint32_t)ip << 8) | (nw << 12));
                        nw++;
                    }
                }
            H.n_work = (int32_t)nw;
            uint32_t aoff = 0;
            for (int32_t n : nd) {
                PatchNode N;
                N.node = n; N.b0 = brow[n]; N.cnt = (uint8_t)(brow[n + 1] - brow[n]);
                N.adj_off = (uint16_t)aoff; N.adj_cnt = (uint8_t)(adj_ptr[n + 1] - adj_ptr[n]);
                P.nodes.push_back(N);
                for (int64_t q = adj_ptr[n]; q < adj_ptr[n + 1]; q++) {
                    const int32_t ad = adj[q];
                    const int32_t e = ad / nsh; const int la = ad - e * nsh;
                    const size_t el = (size_t)(std::lower_bound(elems.begin(), elems.end(), e) - elems.begin());
                    PatchAdj A;
                    std::memset(&A, 0, sizeof A);
                    for (int t = 0; t < ninc; t++) A.slot[t] = slot_of[el * nip + inc[la][t]];
                    A.la = (uint8_t)la;
                    for (int t = 0; t < ninc; t++) if (T.edge[inc[la][t]][1] == la) A.la |= (uint8_t)(16 << t);   // the node is the `to` end of SCVF t
                    for (int k = 0; k < nsh; k++) A.emap[k] = emap[((int64_t)e * nsh + la) * nsh + k];
                    A.self = A.emap[la];
                    P.adj.push_back(A);
                    aoff++;
                }
                P.max_adj_per_node = std::max<int>(P.max_adj_per_node, N.adj_cnt);
            }
            H.n_adj = (int32_t)aoff;
            P.hdr.push_back(H);
        };
        for (int64_t g = g0; g < g1; g++) {
            std::vector<int32_t> nd;
            nd.reserve(gstart[g + 1] - gstart[g]);
            for (int64_t i = gstart[g]; i < gstart[g + 1]; i++) nd.push_back(keys[i].node);
            emit(nd);
        }
    });
    // ---- concatenate in thread order (= tile order) ----
    out = PatchPlan();
    for (unsigned t = 0; t < maxthreads; t++) {
        if (!used[t]) continue;
        Local& L = locals[t];
        if (!L.err.empty()) { err = L.err; return false; }
        const int32_t n0 = (int32_t)out.nodes.size(), e0 = (int32_t)out.elems.size(), w0 = (int32_t)out.work.size(), a0 = (int32_t)out.adj.size(), l0 = (int32_t)out.lnodes.size();
        if ((double)out.adj.size() + L.p.adj.size() >= 2147483647.0 || (double)out.work.size() + L.p.work.size() >= 2147483647.0) { err = "grid too large for 32-bit patch tables"; return false; }
        for (PatchHdr H : L.p.hdr) { H.node0 += n0; H.elem0 += e0; H.work0 += w0; H.adj0 += a0; H.lnode0 += l0; out.hdr.push_back(H); }
        out.nodes.insert(out.nodes.end(), L.p.nodes.begin(), L.p.nodes.end());
        out.elems.insert(out.elems.end(), L.p.elems.begin(), L.p.elems.end());
        out.pconn.insert(out.pconn.end(), L.p.pconn.begin(), L.p.pconn.end());
        out.lnodes.insert(out.lnodes.end(), L.p.lnodes.begin(), L.p.lnodes.end());
        out.ecorner.insert(out.ecorner.end(), L.p.ecorner.begin(), L.p.ecorner.end());
        out.work.insert(out.work.end(), L.p.work.begin(), L.p.work.end());
        out.adj.insert(out.adj.end(), L.p.adj.begin(), L.p.adj.end());
        out.max_adj_per_node = std::max(out.max_adj_per_node, L.p.max_adj_per_node);
        L.p = PatchPlan();
    }
    out.n_scvf_evals = (int64_t)out.work.size();
    (void)n_elem;
    return true;
}

}  // namespace nsb
