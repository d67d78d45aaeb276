// Host-side construction of the tile / gather plan (init-time; see fem_layout.cuh for the layout).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <string>
#include <vector>

namespace sb {

struct HostPlan {
    int n_nodes = 0, n_elems = 0, npe = 0, tile_e = 0, n_tiles = 0, maxval = 0;
    std::vector<uint32_t> order;          // [n_tiles*tile_e] element slot -> original element (0xFFFFFFFF = padding)
    std::vector<uint32_t> tile_node_off;  // [n_tiles+1]
    std::vector<uint32_t> tile_nodes;
    std::vector<uint32_t> tile_shslot;    // aligned with tile_nodes: for a shared node, its index in sh_nodes (chunk * chunk_size + rank); 0xFFFFFFFF for interior entries
    std::vector<uint32_t> tile_nint;
    std::vector<uint32_t> tile_nb;        // [n_tiles] elements of the tile that feed at least one shared node, when they come FIRST in the tile (else 0xFFFFFFFF)
    std::vector<uint16_t> tile_val;
    std::vector<uint16_t> tile_jds;       // [n_tiles][maxval+1]
    std::vector<uint16_t> lnode;          // [n_tiles*tile_e*npe]
    std::vector<uint32_t> slot;           // [n_tiles*tile_e*npe]
    int n_shared = 0, n_chunks = 0, n_forced_shared = 0;
    std::vector<uint32_t> sh_nodes;       // [n_chunks*chunk]
    std::vector<uint16_t> sh_val;
    std::vector<uint32_t> sh_base;        // [n_chunks] first staging entry of the chunk's ELL block
    size_t stage_n = 0;
    int max_touched = 0, max_slots = 0, max_int = 1, max_shtouch = 1;
    size_t n_interior = 0, n_staged_corners = 0, n_demoted = 0;
};

inline uint64_t morton_spread(uint64_t v) {  // 21 bits -> every third bit
    v &= 0x1fffffULL;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}

// elems: n_elems x npe node indices (original topology order); pos: 3*n_nodes doubles (rest positions).
// Returns "" or an error text.
// force_shared (optional): n_nodes flags, nodes that must take the staging path whatever their incident elements.
// reorder_from_tile (optional): tiles from this index on list their elements that feed shared nodes first (tile_nb).
// smem_limit / sv_bytes / slot_bytes (optional): shared-memory budget of a tile CTA, bytes of one staged nodal vector and of
// one slot.  When a tile's interior nodes need more slots than fit, its highest-valence interior nodes are demoted to
// shared nodes (their contributions go through the HBM/L2 staging buffer instead), so any tile size can be made to fit.
inline std::string build_plan(HostPlan& P, int n_nodes, int n_elems, int npe, const uint32_t* elems, const double* pos,
                              int tile_e, int chunk, uint32_t stage_flag, size_t smem_limit = 0, size_t sv_bytes = 0, size_t slot_bytes = 0,
                              const unsigned char* force_shared = nullptr, int reorder_from_tile = 0x7fffffff) {
    P = HostPlan();
    P.n_nodes = n_nodes; P.n_elems = n_elems; P.npe = npe; P.tile_e = tile_e;
    P.n_tiles = std::max(1, (n_elems + tile_e - 1) / tile_e);
    for (size_t i = 0; i < size_t(n_elems) * npe; ++i)
        if (elems[i] >= uint32_t(n_nodes)) return "element refers to a node index out of range";

    // ---- spatial order of the elements: Morton code of the integer cell that holds the rest centroid.  The cell size is
    // the edge of the average hexahedron-equivalent element, so that on a regular grid the 6 tetrahedra of one cube share
    // a cell and power-of-8 tiles are exact cubes of cells (minimal tile surface => fewest shared nodes).
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n_nodes; ++i) for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], pos[3 * size_t(i) + c]); hi[c] = std::max(hi[c], pos[3 * size_t(i) + c]); }
    double ext = 0, vol = 1; int flat = 0;
    for (int c = 0; c < 3; ++c) { const double e = hi[c] - lo[c]; ext = std::max(ext, e); if (e > 0) vol *= e; else flat++; }
    if (!(ext > 0)) ext = 1;
    const double cells = std::max(1.0, npe == 4 ? n_elems / 6.0 : double(n_elems));
    double h = flat == 0 ? std::cbrt(vol / cells) : (flat == 1 ? std::sqrt(vol / cells) : (flat == 2 ? vol / cells : 1.0));
    if (!(h > 0) || ext / h > 2000000.0) h = ext / 2000000.0;
    std::vector<uint64_t> key(n_elems);
    for (int e = 0; e < n_elems; ++e) {
        uint64_t k = 0;
        for (int c = 0; c < 3; ++c) {
            double s = 0; for (int a = 0; a < npe; ++a) s += pos[3 * size_t(elems[size_t(e) * npe + a]) + c];
            double u = std::floor((s / npe - lo[c]) / h); u = std::min(std::max(u, 0.0), 2097151.0);
            k |= morton_spread(uint64_t(u)) << c;
        }
        key[e] = k;
    }
    std::vector<uint32_t> sorted(n_elems);
    std::iota(sorted.begin(), sorted.end(), 0u);
    std::stable_sort(sorted.begin(), sorted.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    const size_t n_slots = size_t(P.n_tiles) * tile_e;
    P.order.assign(n_slots, 0xFFFFFFFFu);
    std::vector<uint32_t> tile_of(n_elems), slot_of_elem(n_elems);
    for (int i = 0; i < n_elems; ++i) { P.order[i] = sorted[i]; tile_of[sorted[i]] = uint32_t(i / tile_e); slot_of_elem[sorted[i]] = uint32_t(i); }

    // ---- node -> incident (element, corner) in ascending original element order
    std::vector<uint32_t> inc_off(size_t(n_nodes) + 1, 0);
    for (size_t i = 0; i < size_t(n_elems) * npe; ++i) inc_off[elems[i] + 1]++;
    for (int i = 0; i < n_nodes; ++i) inc_off[i + 1] += inc_off[i];
    std::vector<uint32_t> inc(size_t(n_elems) * npe), fill(inc_off.begin(), inc_off.end() - 1);
    for (int e = 0; e < n_elems; ++e) for (int c = 0; c < npe; ++c) inc[fill[elems[size_t(e) * npe + c]]++] = uint32_t(e) * npe + c;
    int maxval = 1;
    for (int i = 0; i < n_nodes; ++i) maxval = std::max<int>(maxval, inc_off[i + 1] - inc_off[i]);
    if (maxval > 1023) return "node valence above 1023 is not supported";
    P.maxval = maxval;

    // ---- interior / shared classification
    std::vector<int32_t> interior_tile(n_nodes, -1);
    for (int i = 0; i < n_nodes; ++i) {
        const uint32_t b = inc_off[i], e = inc_off[i + 1];
        if (b == e) continue;
        if (force_shared && force_shared[i]) continue;   // (e.g. partition-interface nodes of a multi-GPU run)
        const uint32_t t0 = tile_of[inc[b] / npe];
        bool same = true;
        for (uint32_t k = b + 1; k < e && same; ++k) same = tile_of[inc[k] / npe] == t0;
        if (same) interior_tile[i] = int32_t(t0);
    }
    std::vector<std::vector<uint32_t>> tile_int(P.n_tiles);
    for (int i = 0; i < n_nodes; ++i) if (interior_tile[i] >= 0) tile_int[interior_tile[i]].push_back(uint32_t(i));
    if (smem_limit) {
        // nodes touched per tile (independent of the classification) -> slot budget -> demotion
        std::vector<uint32_t> mark(n_nodes, 0xFFFFFFFFu);
        size_t max_touched = 1;
        for (int t = 0; t < P.n_tiles; ++t) {
            size_t cnt = 0;
            const size_t s0 = size_t(t) * tile_e, s1 = std::min(n_slots, s0 + tile_e);
            for (size_t s = s0; s < s1; ++s) {
                const uint32_t e = P.order[s];
                if (e == 0xFFFFFFFFu) continue;
                for (int c = 0; c < npe; ++c) { const uint32_t n = elems[size_t(e) * npe + c]; if (mark[n] != uint32_t(t)) { mark[n] = uint32_t(t); ++cnt; } }
            }
            max_touched = std::max(max_touched, cnt);
        }
        const size_t in_bytes = (sv_bytes * max_touched + 15) & ~size_t(15);
        if (in_bytes + slot_bytes * size_t(maxval) > smem_limit) return "tile touches too many nodes for shared memory; use a smaller tile";
        const size_t budget = std::min<size_t>(65535, (smem_limit - in_bytes) / slot_bytes);
        for (int t = 0; t < P.n_tiles; ++t) {
            std::vector<uint32_t>& in = tile_int[t];
            auto val = [&](uint32_t n) { return size_t(inc_off[n + 1] - inc_off[n]); };
            std::stable_sort(in.begin(), in.end(), [&](uint32_t a, uint32_t b) { return val(a) > val(b); });
            size_t total = 0;
            for (uint32_t n : in) total += val(n);
            size_t drop = 0;
            while (total > budget && drop < in.size()) { total -= val(in[drop]); interior_tile[in[drop]] = -1; ++drop; }
            in.erase(in.begin(), in.begin() + drop);
            P.n_demoted += drop;
        }
    }

    // ---- inside a tile, the elements that touch a shared node come first (both groups keep their spatial order): once they
    // are done the tile has nothing more to stage, which lets the persistent CG kernel announce "my staged contributions are
    // complete" before the tile's remaining elements and hide the grid barrier behind them.  The order of the elements inside a
    // tile has no influence on any result (every sum runs in ORIGINAL element order).  Only the tiles a CTA processes LAST
    // (t >= reorder_from_tile) are reordered: the mixed order streams a little faster (measured: +0.45 us per 3328-tet tile).
    P.tile_nb.assign(P.n_tiles, 0xFFFFFFFFu);
    for (int t = std::max(0, reorder_from_tile); t < P.n_tiles; ++t) {
        const size_t s0 = size_t(t) * tile_e, s1 = std::min(n_slots, s0 + tile_e);
        auto feeds_shared = [&](uint32_t e) {
            if (e == 0xFFFFFFFFu) return false;
            for (int c = 0; c < npe; ++c) if (interior_tile[elems[size_t(e) * npe + c]] < 0) return true;
            return false;
        };
        auto mid = std::stable_partition(P.order.begin() + s0, P.order.begin() + s1, feeds_shared);
        P.tile_nb[t] = uint32_t(mid - (P.order.begin() + s0));
        // (padding slots, if any, stay behind the real elements)
        std::stable_partition(mid, P.order.begin() + s1, [](uint32_t e) { return e != 0xFFFFFFFFu; });
    }

    std::vector<uint32_t> corner_slot(size_t(n_elems) * npe, 0);
    std::vector<int32_t> local_idx(n_nodes, -1);
    P.tile_node_off.assign(P.n_tiles + 1, 0);
    P.tile_nint.assign(P.n_tiles, 0);
    P.tile_jds.assign(size_t(P.n_tiles) * (maxval + 1), 0);
    P.lnode.assign(n_slots * npe, 0xFFFFu);
    for (int t = 0; t < P.n_tiles; ++t) {
        std::vector<uint32_t>& in = tile_int[t];
        auto val = [&](uint32_t n) { return inc_off[n + 1] - inc_off[n]; };
        std::stable_sort(in.begin(), in.end(), [&](uint32_t a, uint32_t b) { return val(a) > val(b); });
        // jagged-diagonal offsets
        std::vector<uint32_t> jds(maxval + 1, 0);
        for (int j = 0; j < maxval; ++j) {
            uint32_t cnt = 0;  // nodes with valence > j: a prefix of the ranked list
            while (cnt < in.size() && val(in[cnt]) > uint32_t(j)) ++cnt;
            jds[j + 1] = jds[j] + cnt;
        }
        if (jds[maxval] > 65535u) return "tile needs more than 65535 shared-memory slots; use a smaller tile";
        for (int j = 0; j <= maxval; ++j) P.tile_jds[size_t(t) * (maxval + 1) + j] = uint16_t(jds[j]);
        P.max_slots = std::max<int>(P.max_slots, jds[maxval]);
        const uint32_t base = uint32_t(P.tile_nodes.size());
        P.tile_node_off[t] = base;
        P.tile_nint[t] = uint32_t(in.size());
        for (size_t k = 0; k < in.size(); ++k) {
            const uint32_t n = in[k];
            local_idx[n] = int32_t(k);
            P.tile_nodes.push_back(n);
            P.tile_val.push_back(uint16_t(val(n)));

            for (uint32_t j = 0; j < val(n); ++j) corner_slot[inc[inc_off[n] + j]] = jds[j] + uint32_t(k);
        }
        P.n_interior += in.size();
        P.max_int = std::max<int>(P.max_int, int(in.size()));
        // shared nodes touched by this tile, ascending id
        std::vector<uint32_t> sh;
        const size_t s0 = size_t(t) * tile_e, s1 = std::min(n_slots, s0 + tile_e);
        for (size_t s = s0; s < s1; ++s) {
            const uint32_t e = P.order[s];
            if (e == 0xFFFFFFFFu) continue;
            for (int c = 0; c < npe; ++c) { const uint32_t n = elems[size_t(e) * npe + c]; if (local_idx[n] < 0) { sh.push_back(n); local_idx[n] = 0x40000000; } }
        }
        std::sort(sh.begin(), sh.end());
        for (size_t k = 0; k < sh.size(); ++k) { local_idx[sh[k]] = int32_t(in.size() + k); P.tile_nodes.push_back(sh[k]); P.tile_val.push_back(0); }
        const size_t touched = in.size() + sh.size();
        if (touched >= 65535) return "tile touches more than 65534 nodes; use a smaller tile";
        P.max_touched = std::max<int>(P.max_touched, int(touched));
        P.max_shtouch = std::max<int>(P.max_shtouch, int(sh.size()));
        for (size_t s = s0; s < s1; ++s) {
            const uint32_t e = P.order[s];
            if (e == 0xFFFFFFFFu) continue;
            for (int c = 0; c < npe; ++c) P.lnode[s * npe + c] = uint16_t(local_idx[elems[size_t(e) * npe + c]]);
        }
        for (uint32_t n : in) local_idx[n] = -1;
        for (uint32_t n : sh) local_idx[n] = -1;
    }
    P.tile_node_off[P.n_tiles] = uint32_t(P.tile_nodes.size());
    if (P.max_touched < 1) P.max_touched = 1;
    if (P.max_slots < 1) P.max_slots = 1;

    // ---- shared nodes, ascending id, in chunks of `chunk` nodes.  Contribution j (element order) of the node of rank k in
    // chunk c is staged at sh_base[c] + j*chunk + k: an ELL block per chunk, padded to the chunk's largest valence, so that
    // the staging address needs no index table and consecutive threads (k) read consecutive 16-byte entries.
    std::vector<uint32_t> shared;
    // (nodes flagged in force_shared come first: in a multi-GPU run these are the partition-interface nodes, whose partial sums must
    // leave for the other GPUs as early as possible in the shared-node phase)
    if (force_shared) for (int i = 0; i < n_nodes; ++i) if (interior_tile[i] < 0 && force_shared[i] && inc_off[i + 1] > inc_off[i]) shared.push_back(uint32_t(i));
    P.n_forced_shared = int(shared.size());
    for (int i = 0; i < n_nodes; ++i) if (interior_tile[i] < 0 && !(force_shared && force_shared[i] && inc_off[i + 1] > inc_off[i])) shared.push_back(uint32_t(i));
    P.n_shared = int(shared.size());
    P.n_chunks = std::max(1, (P.n_shared + chunk - 1) / chunk);
    P.sh_nodes.assign(size_t(P.n_chunks) * chunk, 0xFFFFFFFFu);
    P.sh_val.assign(size_t(P.n_chunks) * chunk, 0);
    P.sh_base.assign(P.n_chunks, 0);
    size_t stage = 0;
    for (int c = 0; c < P.n_chunks; ++c) {
        const size_t b = size_t(c) * chunk, e = std::min(shared.size(), b + chunk);
        auto val = [&](uint32_t n) { return inc_off[n + 1] - inc_off[n]; };
        uint32_t mv = 0;
        for (size_t i = b; i < e; ++i) mv = std::max(mv, val(shared[i]));
        if (stage + size_t(mv) * chunk >= 0x7fffffffu) return "staging buffer exceeds 2^31 entries";
        P.sh_base[c] = uint32_t(stage);
        for (size_t i = b; i < e; ++i) {
            const uint32_t n = shared[i];
            const size_t k = i - b;
            P.sh_nodes[i] = n;
            P.sh_val[i] = uint16_t(val(n));
            for (uint32_t j = 0; j < val(n); ++j) corner_slot[inc[inc_off[n] + j]] = stage_flag | uint32_t(stage + size_t(j) * chunk + k);
            P.n_staged_corners += val(n);
        }
        stage += size_t(mv) * chunk;
    }
    P.stage_n = std::max<size_t>(stage, 1);
    {
        std::vector<uint32_t> slot_of(n_nodes, 0xFFFFFFFFu);
        for (size_t i = 0; i < shared.size(); ++i) slot_of[shared[i]] = uint32_t(i);   // (chunks are consecutive runs of `chunk` shared nodes: index == chunk * chunk_size + rank)
        P.tile_shslot.assign(P.tile_nodes.size(), 0xFFFFFFFFu);
        for (int t = 0; t < P.n_tiles; ++t)
            for (uint32_t i = P.tile_node_off[t] + P.tile_nint[t]; i < P.tile_node_off[t + 1]; ++i) P.tile_shslot[i] = slot_of[P.tile_nodes[i]];
    }

    // ---- per element-slot destination table
    P.slot.assign(n_slots * npe, 0);
    for (size_t s = 0; s < n_slots; ++s) {
        const uint32_t e = P.order[s];
        if (e == 0xFFFFFFFFu) continue;
        for (int c = 0; c < npe; ++c) P.slot[s * npe + c] = corner_slot[size_t(e) * npe + c];
    }
    (void)slot_of_elem;
    return "";
}

}  // namespace sb
