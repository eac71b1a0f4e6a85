// ns_graph.h -- host-side grid preprocessing shared by libnsb200 (nsb200.cu) and the CPU emulation harness of the fused
// kernel (tests/cpp/emu_fused.cpp): entity -> element adjacency, block-CSR pattern (full element coupling incl. explicit
// zeros), element -> CSR scatter map. "entities" are nodes (FV1) or sides (FVCR velocity dofs). Pure C++.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

namespace nsb {

template <class F> static void parallel_for(int64_t n, F fn)
{
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (n < 4096) nt = 1;
    std::vector<std::thread> th;
    const int64_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const int64_t lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([=]() { fn(lo, hi); });
    }
    for (auto& x : th) x.join();
}

struct EntityGraph {
    std::vector<int64_t> adj_ptr;   // entity -> incident (element, local index)
    std::vector<int32_t> adj;       // elem*per + local
    std::vector<int64_t> brow;      // entity -> neighbouring entities (sorted, incl. itself)
    std::vector<int32_t> bcol;
    int max_cnt = 0;
};

// returns an empty string on success, the error text otherwise
static std::string build_entity_graph(int64_t n_elem, int64_t n_ent, int per, const int32_t* conn, EntityGraph& g)
{
    if ((double)n_elem * per >= 2147483647.0) return "grid too large for 32-bit adjacency ids";
    g.adj_ptr.assign(n_ent + 1, 0);
    for (int64_t e = 0; e < n_elem; e++) for (int k = 0; k < per; k++) {
        const int32_t nd = conn[e * per + k];
        if (nd < 0 || nd >= n_ent) return "connectivity entry out of range (element " + std::to_string((long long)e) + ")";
        g.adj_ptr[nd + 1]++;
    }
    for (int64_t i = 0; i < n_ent; i++) g.adj_ptr[i + 1] += g.adj_ptr[i];
    g.adj.resize(g.adj_ptr[n_ent]);
    { std::vector<int64_t> pos(g.adj_ptr.begin(), g.adj_ptr.end() - 1);
      for (int64_t e = 0; e < n_elem; e++) for (int k = 0; k < per; k++) g.adj[pos[conn[e * per + k]]++] = (int32_t)(e * per + k); }
    // neighbour lists: count, prefix, fill
    std::vector<int32_t> cnt(n_ent);
    auto gather = [&](int64_t i, std::vector<int32_t>& tmp) {
        tmp.clear();
        for (int64_t q = g.adj_ptr[i]; q < g.adj_ptr[i + 1]; q++) { const int64_t e = g.adj[q] / per; for (int k = 0; k < per; k++) tmp.push_back(conn[e * per + k]); }
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    };
    parallel_for(n_ent, [&](int64_t lo, int64_t hi) { std::vector<int32_t> tmp; for (int64_t i = lo; i < hi; i++) { gather(i, tmp); cnt[i] = (int32_t)tmp.size(); } });
    g.brow.assign(n_ent + 1, 0);
    g.max_cnt = 0;
    for (int64_t i = 0; i < n_ent; i++) { g.brow[i + 1] = g.brow[i] + cnt[i]; g.max_cnt = std::max(g.max_cnt, (int)cnt[i]); }
    g.bcol.resize(g.brow[n_ent]);
    parallel_for(n_ent, [&](int64_t lo, int64_t hi) { std::vector<int32_t> tmp; for (int64_t i = lo; i < hi; i++) { gather(i, tmp); std::copy(tmp.begin(), tmp.end(), g.bcol.begin() + g.brow[i]); } });
    return std::string();
}

// slot of entity conn[e][k] in the neighbour list of entity conn[e][a]
static void build_emap(int64_t n_elem, int per, const int32_t* conn, const EntityGraph& g, std::vector<uint8_t>& emap)
{
    emap.resize((size_t)n_elem * per * per);
    parallel_for(n_elem, [&](int64_t lo, int64_t hi) {
        for (int64_t e = lo; e < hi; e++) for (int a = 0; a < per; a++) {
            const int32_t na = conn[e * per + a];
            const int32_t* b = g.bcol.data() + g.brow[na]; const int32_t* en = g.bcol.data() + g.brow[na + 1];
            for (int k = 0; k < per; k++) emap[(e * per + a) * per + k] = (uint8_t)(std::lower_bound(b, en, conn[e * per + k]) - b);
        }
    });
}

}  // namespace nsb
