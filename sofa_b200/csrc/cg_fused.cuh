// The whole CGLinearSolver::solve loop (CGLinearSolver.inl:130-272) in ONE persistent cooperative kernel, second generation:
// ONE grid-wide reduction per iteration instead of two, and the shared-node sums hidden behind the element stream.
//
//   reference iteration k                           here
//   ---------------------                           ----
//   p = r + beta p            (cgstep_beta)         local: every CTA updates the p of ALL nodes its tiles touch (redundantly for shared nodes:
//   q = A p                                         same operands, same order => same bits in every copy), no exchange
//   den = p.q ; alpha = rho/den                     ONE reduction of four dot products taken while q is produced:
//   x += alpha p ; r -= alpha q (cgstep_alpha)          r.r (the TRUE rho_k), p.q, r.q, q.q
//   rho' = r.r ; beta = rho'/rho                    rho_{k+1} = rho_k - 2 alpha (r.q) + alpha^2 (q.q)   [exact identity for r' = r - alpha q]
//
// alpha_k uses the true rho_k = r_k.r_k measured in iteration k's own reduction; only beta_k and the tolerance test of iteration k use the
// one-step prediction of rho_{k+1}, whose error does not accumulate (it is replaced by the measured value one iteration later).  x, r, p
// follow the reference's vOp formulas with those scalars.  Measured against the oracle's CG (tests/test_cg_fused_recurrence.py, CPU):
// identical iteration counts, solutions equal to ~1e-7 (Vec3f) / 1e-15 (Vec3d) relative.
//
// Dependencies of an iteration:
//   element pass (tiles) --> [S1: staged contributions complete] --> shared-node sums --> [S2: the four dot products] --> local update
// S1 is waited for by the threads that sum shared nodes only: dedicated warps start on them while the element warps are still inside the
// interior elements of their last tile (the plan lists the elements that feed shared nodes first), and the element warps join when done.
// After S2 everything is CTA-local, so the next element pass starts at once.
//
// Ownership: a tile's interior nodes belong to its CTA (x, r in tile-ordered private arrays; p, q in shared memory when the CTA's tiles
// fit -- "cached" -- or in tile-ordered HBM scratch -- "streamed" -- so any number of tiles per CTA works).  A shared node is finished by
// whichever warp takes its 32-node unit; its x, r, p, q live in slot-indexed arrays.  Its owner applies the x/r/p update of iteration k
// lazily at the start of iteration k+1's sum; the tiles that touch it rebuild r_{k+1}, p_{k+1} from the published r_k, q_k themselves.
#pragma once
#include "cg_persist.cuh"

namespace sb {

constexpr int kUnit = 32;                 // shared nodes per gather unit (one warp)
constexpr int kFusedDots = 4;             // r.r, p.q, r.q, q.q
constexpr int kMaxFusedTiles = 64;        // tiles per CTA in streamed mode (bounds the loops only)

struct FusedLayout {
    int cached;             // 1: p / q / node tables of the CTA's tiles stay in shared memory for the whole solve; 0: streamed from HBM scratch
    int tiles_per_cta, units_per_cta;
    int max_touched, max_slots, max_int, max_shtouch, maxval;
    unsigned off_slot, off_tidx, off_nrec, off_q, off_jds, off_upart, off_extra, total;
};
template <class R> inline FusedLayout fused_layout(bool cached, int tiles_per_cta, int units_per_cta, int max_touched, int max_slots, int max_int, int max_shtouch, int maxval,
                                                   size_t extra_bytes = 0) {
    FusedLayout L;
    L.cached = cached ? 1 : 0; L.tiles_per_cta = tiles_per_cta; L.units_per_cta = units_per_cta;
    L.max_touched = max_touched; L.max_slots = max_slots; L.max_int = max_int; L.max_shtouch = max_shtouch; L.maxval = maxval;
    const int tc = cached ? tiles_per_cta : 1;
    auto up = [](size_t o) { return (o + 15) & ~size_t(15); };
    size_t o = up(sizeof(typename SVec<R>::T) * size_t(max_touched) * tc);
    L.off_slot = unsigned(o); o = up(o + sizeof(R) * 3 * size_t(max_slots));
    L.off_tidx = unsigned(o); if (cached) o = up(o + sizeof(uint32_t) * size_t(max_shtouch) * tc);
    L.off_nrec = unsigned(o); if (cached) o = up(o + sizeof(NodeRec<R>) * size_t(max_int) * tc);
    L.off_q = unsigned(o); if (cached) o = up(o + sizeof(R) * 3 * size_t(max_int) * tc);
    L.off_jds = unsigned(o); o = up(o + sizeof(uint16_t) * size_t(maxval + 1) * tc);
    L.off_upart = unsigned(o); o = up(o + sizeof(double) * kFusedDots * size_t(std::max(units_per_cta, 1)));
    L.off_extra = unsigned(o); o = up(o + extra_bytes);
    L.total = unsigned(o);
    return L;
}

template <class R> struct FusedCG {
    typedef typename SVec<R>::T SV;
    NodeEpilogue<R> ep;     // epilogue of q = A p: mass / projection / plane terms
    R* x; R* r;             // flat vectors of the caller: x is read at the start and written at the end, r holds r_0 (= b unless warm start)
    const R* b;
    SV* xt; SV* rt;         // x, r of the interior nodes in tile order (private to the owner CTA)
    SV* gP; R* gQ; NodeRec<R>* gNrec;   // streamed mode: p of the touched nodes, q and static records of the interior nodes, tile order
    SV* xS; SV* rS; SV* pS; SV* qS;     // shared nodes by slot (index in sh_nodes)
    GRec<R>* shrec;         // [n_chunks * kGatherChunk] static record of each shared node, rebuilt at every solve (mass / fixed may change)
    size_t n3;
    CGDev* cg;
    unsigned long long* sync;   // zero at launch: [3][grid][kFusedDots] doubles, then the S1 and S2 arrival counters (one 128-byte line each)
    FusedLayout lay;
    PeerDev<R> peer;
};
inline size_t fused_sync_words(int grid) { return size_t(3) * grid * kFusedDots + 64; }

// ---- CTA-scope flag in shared memory (release / acquire) ---------------------------------------------------------------------------
__device__ __forceinline__ void st_release_cta_shared(unsigned* p, unsigned v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(unsigned(__cvta_generic_to_shared(p))), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_cta_shared(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(unsigned(__cvta_generic_to_shared(p))) : "memory");
    return v;
}
// barrier over the first N threads of the CTA (the element warps) -- id 1; the whole CTA uses __syncthreads (id 0)
template <int N> __device__ __forceinline__ void bar_first() { asm volatile("bar.sync 1, %0;" :: "n"(N) : "memory"); }

// per-CTA scalars of the solve, in shared memory (kept out of the element loop's registers)
template <class R> struct FusedScal {
    double rho, normb, tol, thr;
    int it; unsigned tsc, max_iter;
    R alpha, malpha, beta;          // of the LAST completed iteration (pending update of the shared nodes)
    int a_one, ma_one;
    int first;                      // no update has been made yet (iteration 1: p = r)
    unsigned iter;                  // iterations started in this launch (sequence number of S1 / S2)
    unsigned vsync;                 // value syncs done (S2 and the prologue's)
    int failed;
};

template <class R> struct FusedAcc { double rr, pq, rq, qq; };
template <class R> __device__ __forceinline__ void acc_node(FusedAcc<R>& a, R r0, R r1, R r2, R p0, R p1, R p2, R q0, R q1, R q2) {
    a.rr += double(r0) * double(r0) + double(r1) * double(r1) + double(r2) * double(r2);
    a.pq += double(p0) * double(q0) + double(p1) * double(q1) + double(p2) * double(q2);
    a.rq += double(r0) * double(q0) + double(r1) * double(q1) + double(r2) * double(q2);
    a.qq += double(q0) * double(q0) + double(q1) * double(q1) + double(q2) * double(q2);
}
// fixed-order warp sum of the four accumulators (result in lane 0)
template <class R> __device__ __forceinline__ void acc_warp_sum(FusedAcc<R>& a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.rr += __shfl_down_sync(0xffffffffu, a.rr, o); a.pq += __shfl_down_sync(0xffffffffu, a.pq, o);
        a.rq += __shfl_down_sync(0xffffffffu, a.rq, o); a.qq += __shfl_down_sync(0xffffffffu, a.qq, o);
    }
}
__device__ __forceinline__ void stcg_sv(float4* p, float4 v) { __stcg(p, v); }
__device__ __forceinline__ void stcg_sv(SVec<double>::T* p, SVec<double>::T v) { double* d = reinterpret_cast<double*>(p); __stcg(d, v.x); __stcg(d + 1, v.y); __stcg(d + 2, v.z); }

// pointers to the per-tile state of tile slot c of this CTA (shared memory when cached, HBM scratch when streamed)
template <class R> struct TileState {
    typedef typename SVec<R>::T SV;
    SV* P; R* Q; const NodeRec<R>* nrec; const uint32_t* tidx; const uint16_t* jds;
    uint32_t node_off; int n_touched, n_int;
};
template <class R> __device__ __forceinline__ TileState<R> tile_state(const TileDev<R>& t, const FusedCG<R>& a, unsigned char* smem_raw, int c, int tile) {
    typedef typename SVec<R>::T SV;
    const FusedLayout& L = a.lay;
    TileState<R> s;
    s.node_off = t.tile_node_off[tile];
    s.n_touched = int(t.tile_node_off[tile + 1] - s.node_off); s.n_int = int(t.tile_nint[tile]);
    if (L.cached) {
        s.P = reinterpret_cast<SV*>(smem_raw) + c * L.max_touched;
        s.Q = reinterpret_cast<R*>(smem_raw + L.off_q) + 3 * size_t(c * L.max_int);
        s.nrec = reinterpret_cast<const NodeRec<R>*>(smem_raw + L.off_nrec) + c * L.max_int;
        s.tidx = reinterpret_cast<const uint32_t*>(smem_raw + L.off_tidx) + c * L.max_shtouch;
        s.jds = reinterpret_cast<const uint16_t*>(smem_raw + L.off_jds) + c * (L.maxval + 1);
    } else {
        s.P = a.gP + s.node_off; s.Q = a.gQ + 3 * size_t(s.node_off); s.nrec = a.gNrec + s.node_off;
        s.tidx = t.tile_shslot + s.node_off + s.n_int;
        s.jds = reinterpret_cast<const uint16_t*>(smem_raw + L.off_jds);      // (copied there per tile)
    }
    return s;
}

// ---- grid-wide sync that also sums kFusedDots doubles per CTA ------------------------------------------------------------------------
// arrival: `v` (lane 0 of warp 0 holds the CTA's four sums) is posted, everything the CTA wrote before is released.  Must be called by
// warp 0 after a CTA-wide barrier.
__device__ __forceinline__ void fused_arrive_values(unsigned long long* sync, unsigned vs, const double v[kFusedDots]) {
    const unsigned G = gridDim.x;
    double* cur = reinterpret_cast<double*>(sync) + (size_t(vs % 3) * G + blockIdx.x) * kFusedDots;
    unsigned* counter = reinterpret_cast<unsigned*>(sync + size_t(3) * G * kFusedDots + 32);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kFusedDots; ++i) cur[i] = v[i];
        __threadfence();
        atomicAdd(counter, 1u);
    }
}
// wait (warp 0) + fixed-order sum over the CTAs; the totals are left in out[] (shared memory) for the CTA-wide barrier that follows
__device__ __forceinline__ bool fused_wait_values(unsigned long long* sync, unsigned vs, double* out /* smem [kFusedDots] */) {
    const unsigned G = gridDim.x;
    const double* cur = reinterpret_cast<const double*>(sync) + size_t(vs % 3) * G * kFusedDots;
    const unsigned* counter = reinterpret_cast<const unsigned*>(sync + size_t(3) * G * kFusedDots + 32);
    bool ok = true;
    if (threadIdx.x == 0) {
        const unsigned target = (vs + 1) * G;
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_u32(counter) < target) { if (globaltimer_ns() - t0 > kSyncTimeoutNs) { ok = false; break; } }
    }
    ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
    __syncwarp();
    double v[kFusedDots];
#pragma unroll
    for (int i = 0; i < kFusedDots; ++i) v[i] = 0.0;
    for (unsigned c = threadIdx.x; c < G; c += 32) {
#pragma unroll
        for (int i = 0; i < kFusedDots; ++i) v[i] += __ldcg(cur + size_t(c) * kFusedDots + i);
    }
#pragma unroll
    for (int i = 0; i < kFusedDots; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], o);
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kFusedDots; ++i) out[i] = v[i];
    }
    return ok;
}
// S1: "this CTA's staged contributions are complete" -- arrival only (the waiters are the gather threads)
__device__ __forceinline__ void fused_arrive_s1(unsigned long long* sync) {
    unsigned* counter = reinterpret_cast<unsigned*>(sync + size_t(3) * gridDim.x * kFusedDots);
    __threadfence();
    atomicAdd(counter, 1u);
}
__device__ __forceinline__ bool fused_poll_s1(const unsigned long long* sync, unsigned iter) {
    const unsigned* counter = reinterpret_cast<const unsigned*>(sync + size_t(3) * gridDim.x * kFusedDots);
    const unsigned target = (iter + 1) * gridDim.x;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_u32(counter) < target) { if (globaltimer_ns() - t0 > kSyncTimeoutNs) return false; }
    return true;
}

// ---- once per solve: static tables ------------------------------------------------------------------------------------------------
template <class R, int NT> __device__ __forceinline__ void fused_load_tables(const TileDev<R>& t, const FusedCG<R>& a, unsigned char* smem_raw) {
    const FusedLayout& L = a.lay;
    const NodeEpilogue<R>& ep = a.ep;
    const int G = int(gridDim.x);
    for (int c = 0; c < L.tiles_per_cta; ++c) {
        const int tile = blockIdx.x + c * G;
        if (tile >= t.n_tiles) break;
        const uint32_t node_off = t.tile_node_off[tile];
        const int n_touched = int(t.tile_node_off[tile + 1] - node_off), n_int = int(t.tile_nint[tile]);
        NodeRec<R>* nrec = L.cached ? reinterpret_cast<NodeRec<R>*>(smem_raw + L.off_nrec) + c * L.max_int : a.gNrec + node_off;
        for (int k = threadIdx.x; k < n_touched; k += NT) {
            if (k < n_int) {
                const uint32_t g = t.tile_nodes[node_off + k];
                nrec[k] = NodeRec<R>{g, unsigned(t.tile_val[node_off + k]) | ((ep.fixed && ep.fixed[g]) ? 0x10000u : 0u), ep.mass ? ep.mass[g] : R(0)};
            } else if (L.cached) reinterpret_cast<uint32_t*>(smem_raw + L.off_tidx)[c * L.max_shtouch + (k - n_int)] = t.tile_shslot[node_off + k];
        }
        if (L.cached)
            for (int j = threadIdx.x; j <= t.maxval; j += NT) reinterpret_cast<uint16_t*>(smem_raw + L.off_jds)[c * (L.maxval + 1) + j] = t.tile_jds[size_t(tile) * (t.maxval + 1) + j];
    }
    // the shared nodes of this CTA's units
    for (int lu = 0; lu < L.units_per_cta; ++lu) {
        const int unit = lu * G + int(blockIdx.x);
        if (size_t(unit) * kUnit >= size_t(t.n_chunks) * kGatherChunk) break;
        for (int k = threadIdx.x; k < kUnit; k += NT) {
            const size_t slot = size_t(unit) * kUnit + k;
            const uint32_t g = t.sh_nodes[slot];
            GRec<R> rec{0xFFFFFFFFu, 0u, 0u, R(0)};
            if (g != 0xFFFFFFFFu)
                rec = GRec<R>{g, unsigned(t.sh_val[slot]) | ((ep.fixed && ep.fixed[g]) ? 0x10000u : 0u), t.sh_base[slot / kGatherChunk] + uint32_t(slot % kGatherChunk), ep.mass ? ep.mass[g] : R(0)};
            a.shrec[slot] = rec;
        }
    }
}

// ---- start of the solve: |b|, rho_0 = r.r (CGLinearSolver.inl:130-180) and the initial state p = r of every copy --------------------
template <class R, int NT> __device__ __forceinline__ bool fused_init(const TileDev<R>& t, const FusedCG<R>& a, FusedScal<R>* sc, unsigned char* smem_raw, double* s_wpart, double* s_tot) {
    typedef typename SVec<R>::T SV;
    const FusedLayout& L = a.lay;
    const PeerDev<R>& P = a.peer;
    CGDev* cg = a.cg;
    const int G = int(gridDim.x);
    const size_t n = a.n3 / 3;
    double sb = 0.0, sr = 0.0;
    for (size_t g = size_t(blockIdx.x) * NT + threadIdx.x; g < n; g += size_t(G) * NT) {
        if (P.enabled && !P.owned[g]) continue;
        const R b0 = a.b[3 * g], b1 = a.b[3 * g + 1], b2 = a.b[3 * g + 2], r0 = a.r[3 * g], r1 = a.r[3 * g + 1], r2 = a.r[3 * g + 2];
        sb += double(b0) * double(b0) + double(b1) * double(b1) + double(b2) * double(b2);
        sr += double(r0) * double(r0) + double(r1) * double(r1) + double(r2) * double(r2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sb += __shfl_down_sync(0xffffffffu, sb, o); sr += __shfl_down_sync(0xffffffffu, sr, o); }
    if ((threadIdx.x & 31) == 0) { s_wpart[(threadIdx.x >> 5) * kFusedDots] = sb; s_wpart[(threadIdx.x >> 5) * kFusedDots + 1] = sr; }
    // p = r, private copies of x and r
    for (int c = 0; c < L.tiles_per_cta; ++c) {
        const int tile = blockIdx.x + c * G;
        if (tile >= t.n_tiles) break;
        const TileState<R> s = tile_state<R>(t, a, smem_raw, c, tile);
        for (int k = threadIdx.x; k < s.n_touched; k += NT) {
            const size_t g = t.tile_nodes[s.node_off + k];
            const SV rv = SVec<R>::make(a.r[3 * g], a.r[3 * g + 1], a.r[3 * g + 2]);
            s.P[k] = rv;
            if (k < s.n_int) { a.rt[s.node_off + k] = rv; a.xt[s.node_off + k] = SVec<R>::make(a.x[3 * g], a.x[3 * g + 1], a.x[3 * g + 2]); }
        }
    }
    for (int lu = 0; lu < L.units_per_cta; ++lu) {
        const int unit = lu * G + int(blockIdx.x);
        if (size_t(unit) * kUnit >= size_t(t.n_chunks) * kGatherChunk) break;
        for (int k = threadIdx.x; k < kUnit; k += NT) {
            const size_t slot = size_t(unit) * kUnit + k;
            const size_t g = t.sh_nodes[slot];
            if (g == 0xFFFFFFFFu) continue;
            const SV rv = SVec<R>::make(a.r[3 * g], a.r[3 * g + 1], a.r[3 * g + 2]);
            stcg_sv(a.rS + slot, rv); stcg_sv(a.pS + slot, rv);
            stcg_sv(a.xS + slot, SVec<R>::make(a.x[3 * g], a.x[3 * g + 1], a.x[3 * g + 2]));
        }
    }
    __syncthreads();
    bool ok = true;
    if (threadIdx.x < 32) {
        double v[kFusedDots] = {0.0, 0.0, 0.0, 0.0};
        for (int w = threadIdx.x; w < NT / 32; w += 32) { v[0] += s_wpart[w * kFusedDots]; v[1] += s_wpart[w * kFusedDots + 1]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v[0] += __shfl_down_sync(0xffffffffu, v[0], o); v[1] += __shfl_down_sync(0xffffffffu, v[1], o); }
        fused_arrive_values(a.sync, 0u, v);
        ok = fused_wait_values(a.sync, 0u, s_tot);
    }
    __syncthreads();
    const double nb2 = s_tot[0], rho0 = s_tot[1];
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    if (threadIdx.x == 0) {
        sc->vsync = 1; sc->iter = 0; sc->first = 1; sc->failed = ok ? 0 : 1;
        sc->normb = sqrt(nb2); sc->rho = rho0; sc->it = 1;
        sc->tol = cg->tolerance; sc->thr = cg->threshold; sc->tsc = cg->time_step_count; sc->max_iter = cg->max_iter;
        sc->alpha = R(0); sc->malpha = R(0); sc->beta = R(0); sc->a_one = 0; sc->ma_one = 0;
    }
    __syncthreads();
    if (sc->failed) { if (lead) { cg->done = 1; cg->end_cond = 99; } return false; }
    if (lead) cg->normb = sc->normb;
    if (sc->normb == 0.0) { if (lead) { cg->done = 1; cg->nb_iter = 0; cg->end_cond = 4; } return false; }
    if (lead) cg_after_rho(cg, rho0);                      // it: 0 -> 1, first entry of the error graph, tolerance test
    if (1u > sc->max_iter) return false;
    const double err = sqrt(rho0) / sc->normb;
    if (err <= sc->tol && !(sc->tsc == 0)) return false;
    return true;
}

// ---- interior nodes of one tile: ordered sum of their slots, epilogue, q kept for the update, the four dot products -------------------
template <class R, int ET> __device__ __forceinline__ void fused_interior(const FusedCG<R>& a, const TileState<R>& s, const R* s_slot, FusedAcc<R>& acc) {
    typedef typename SVec<R>::T SV;
    const NodeEpilogue<R>& ep = a.ep;
    const int max_slots = a.lay.max_slots;
    const bool plus = ep.sign > 0;
    const bool counted_all = !a.peer.enabled;
    for (int k = threadIdx.x; k < s.n_int; k += ET) {
        const SV rv = sv_ldcg(a.rt + s.node_off + k);
        const NodeRec<R> rec = s.nrec[k];
        const int val = int(rec.val_fixed & 0xFFFFu);
        const SV pv = s.P[k];
        R ax = R(0), ay = R(0), az = R(0);
        node_mass_m(ep, ep.pre_kind, rec.mass, R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
        int jj = 0;
        for (; jj + 4 <= val; jj += 4) {
            R cx[4], cy[4], cz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int sl = s.jds[jj + u] + k; cx[u] = s_slot[sl]; cy[u] = s_slot[max_slots + sl]; cz[u] = s_slot[2 * max_slots + sl]; }
            if (plus) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { ax += cx[u]; ay += cy[u]; az += cz[u]; }
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) { ax -= cx[u]; ay -= cy[u]; az -= cz[u]; }
            }
        }
        for (; jj < val; ++jj) {
            const int sl = s.jds[jj] + k;
            if (plus) { ax += s_slot[sl]; ay += s_slot[max_slots + sl]; az += s_slot[2 * max_slots + sl]; }
            else { ax -= s_slot[sl]; ay -= s_slot[max_slots + sl]; az -= s_slot[2 * max_slots + sl]; }
        }
        node_finish_m(ep, rec.g, rec.mass, (rec.val_fixed & 0x10000u) != 0, R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
        s.Q[3 * k] = ax; s.Q[3 * k + 1] = ay; s.Q[3 * k + 2] = az;
        if (counted_all || a.peer.owned[rec.g]) acc_node<R>(acc, R(rv.x), R(rv.y), R(rv.z), R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
    }
}

// ---- one unit of 32 shared nodes (one warp, lane = node) ---------------------------------------------------------------------------------
// Applies the pending update of the previous iteration, sums the staged contributions in element order, publishes r and q for the tiles.
template <class R> struct GatherBatch;
template <> struct GatherBatch<float> { static constexpr int N = 8; };
template <> struct GatherBatch<double> { static constexpr int N = 4; };
template <class R> __device__ __forceinline__ void fused_unit(const TileDev<R>& t, const FusedCG<R>& a, const FusedScal<R>* sc, int unit, double* upart /* smem [kFusedDots] */) {
    typedef typename SVec<R>::T SV;
    constexpr int B = GatherBatch<R>::N;
    const NodeEpilogue<R>& ep = a.ep;
    const int lane = threadIdx.x & 31;
    const size_t slot = size_t(unit) * kUnit + lane;
    FusedAcc<R> acc{0.0, 0.0, 0.0, 0.0};
    const GRec<R> rec = a.shrec[slot];
    if (rec.g != 0xFFFFFFFFu) {
        const int val = int(rec.val_fixed & 0xFFFFu);
        const Quad<R>* stg = t.stage + rec.base;
        const uint64_t pol = l2_policy_evict_first();
        const Quad<R> zero{R(0), R(0), R(0), R(0)};
        Quad<R> b0[B], b1[B];
#pragma unroll
        for (int u = 0; u < B; ++u) b0[u] = u < val ? stage_load(stg + size_t(u) * kGatherChunk, pol) : zero;
#pragma unroll
        for (int u = 0; u < B; ++u) b1[u] = B + u < val ? stage_load(stg + size_t(B + u) * kGatherChunk, pol) : zero;
        SV pv = sv_ldcg(a.pS + slot), rv = sv_ldcg(a.rS + slot);
        R p0 = R(pv.x), p1 = R(pv.y), p2 = R(pv.z), r0 = R(rv.x), r1 = R(rv.y), r2 = R(rv.z);
        if (!sc->first) {
            // x += alpha p ; r -= alpha q ; p = p beta + r   of the previous iteration (cgstep_alpha, cgstep_beta)
            const SV xv = sv_ldcg(a.xS + slot), qo = sv_ldcg(a.qS + slot);
            R x0 = R(xv.x), x1 = R(xv.y), x2 = R(xv.z);
            const R alpha = sc->alpha, malpha = sc->malpha, beta = sc->beta;
            const bool a_one = sc->a_one != 0, ma_one = sc->ma_one != 0;
            x_one<R>(x0, p0, alpha, a_one); x_one<R>(x1, p1, alpha, a_one); x_one<R>(x2, p2, alpha, a_one);
            r_one<R>(r0, R(qo.x), malpha, ma_one); r_one<R>(r1, R(qo.y), malpha, ma_one); r_one<R>(r2, R(qo.z), malpha, ma_one);
            p0 = p_update<R>(p0, beta, r0); p1 = p_update<R>(p1, beta, r1); p2 = p_update<R>(p2, beta, r2);
            stcg_sv(a.xS + slot, SVec<R>::make(x0, x1, x2));
            stcg_sv(a.rS + slot, SVec<R>::make(r0, r1, r2));
            stcg_sv(a.pS + slot, SVec<R>::make(p0, p1, p2));
        }
        R q0 = R(0), q1 = R(0), q2 = R(0);
        node_mass_m(ep, ep.pre_kind, rec.mass, p0, p1, p2, q0, q1, q2);
        const bool plus = ep.sign > 0;
        for (int j0 = 0; j0 < val; j0 += 2 * B) {
#pragma unroll
            for (int u = 0; u < B; ++u)
                if (j0 + u < val) { if (plus) { q0 += b0[u].a; q1 += b0[u].b; q2 += b0[u].c; } else { q0 -= b0[u].a; q1 -= b0[u].b; q2 -= b0[u].c; } }
            if (j0 + B >= val) break;
#pragma unroll
            for (int u = 0; u < B; ++u) b0[u] = j0 + 2 * B + u < val ? stage_load(stg + size_t(j0 + 2 * B + u) * kGatherChunk, pol) : zero;
#pragma unroll
            for (int u = 0; u < B; ++u)
                if (j0 + B + u < val) { if (plus) { q0 += b1[u].a; q1 += b1[u].b; q2 += b1[u].c; } else { q0 -= b1[u].a; q1 -= b1[u].b; q2 -= b1[u].c; } }
            if (j0 + 2 * B >= val) break;
#pragma unroll
            for (int u = 0; u < B; ++u) b1[u] = j0 + 3 * B + u < val ? stage_load(stg + size_t(j0 + 3 * B + u) * kGatherChunk, pol) : zero;
        }
        node_finish_m(ep, rec.g, rec.mass, (rec.val_fixed & 0x10000u) != 0, p0, p1, p2, q0, q1, q2);
        stcg_sv(a.qS + slot, SVec<R>::make(q0, q1, q2));
        acc_node<R>(acc, r0, r1, r2, p0, p1, p2, q0, q1, q2);
    }
    acc_warp_sum<R>(acc);
    if (lane == 0) { upart[0] = acc.rr; upart[1] = acc.pq; upart[2] = acc.rq; upart[3] = acc.qq; }
}

// ---- after S2: x, r, p of every node the CTA's tiles touch ----------------------------------------------------------------------------------
// with_p = false: the last iteration of the solve (only x matters; r is updated too so that the vector the caller sees is consistent)
template <class R, int NT> __device__ __forceinline__ void fused_update(const TileDev<R>& t, const FusedCG<R>& a, const FusedScal<R>* sc, unsigned char* smem_raw, bool with_p) {
    typedef typename SVec<R>::T SV;
    const FusedLayout& L = a.lay;
    const int G = int(gridDim.x);
    const R alpha = sc->alpha, malpha = sc->malpha, beta = sc->beta;
    const bool a_one = sc->a_one != 0, ma_one = sc->ma_one != 0;
    for (int c = 0; c < L.tiles_per_cta; ++c) {
        const int tile = blockIdx.x + c * G;
        if (tile >= t.n_tiles) break;
        const TileState<R> s = tile_state<R>(t, a, smem_raw, c, tile);
        constexpr int U = 2;        // rounds whose L2 requests are all issued before the first use
        for (int k0 = threadIdx.x; k0 < s.n_touched; k0 += U * NT) {
            SV va[U], vb[U];        // interior: x, r (private arrays) ; shared: r, q (the owner's published values)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = k0 + u * NT;
                if (k < s.n_int) { va[u] = sv_ldcg(a.xt + s.node_off + k); vb[u] = sv_ldcg(a.rt + s.node_off + k); }
                else if (k < s.n_touched) { const size_t slot = s.tidx[k - s.n_int]; va[u] = sv_ldcg(a.rS + slot); vb[u] = sv_ldcg(a.qS + slot); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = k0 + u * NT;
                if (k >= s.n_touched) continue;
                const SV pv = s.P[k];
                R p0 = R(pv.x), p1 = R(pv.y), p2 = R(pv.z), r0, r1, r2;
                if (k < s.n_int) {
                    R x0 = R(va[u].x), x1 = R(va[u].y), x2 = R(va[u].z);
                    r0 = R(vb[u].x); r1 = R(vb[u].y); r2 = R(vb[u].z);
                    x_one<R>(x0, p0, alpha, a_one); x_one<R>(x1, p1, alpha, a_one); x_one<R>(x2, p2, alpha, a_one);
                    r_one<R>(r0, s.Q[3 * k], malpha, ma_one); r_one<R>(r1, s.Q[3 * k + 1], malpha, ma_one); r_one<R>(r2, s.Q[3 * k + 2], malpha, ma_one);
                    stcg_sv(a.xt + s.node_off + k, SVec<R>::make(x0, x1, x2));
                    stcg_sv(a.rt + s.node_off + k, SVec<R>::make(r0, r1, r2));
                } else {
                    r0 = R(va[u].x); r1 = R(va[u].y); r2 = R(va[u].z);
                    r_one<R>(r0, R(vb[u].x), malpha, ma_one); r_one<R>(r1, R(vb[u].y), malpha, ma_one); r_one<R>(r2, R(vb[u].z), malpha, ma_one);
                }
                if (with_p) s.P[k] = SVec<R>::make(p_update<R>(p0, beta, r0), p_update<R>(p1, beta, r1), p_update<R>(p2, beta, r2));
            }
        }
    }
}

// ---- end of the solve: x back in the caller's flat vector -----------------------------------------------------------------------------------
// pending: the shared nodes still owe the x update of the last iteration (the solve ended after an update)
template <class R, int NT> __device__ __forceinline__ void fused_finish(const TileDev<R>& t, const FusedCG<R>& a, const FusedScal<R>* sc, unsigned char* smem_raw, bool updated, bool pending) {
    typedef typename SVec<R>::T SV;
    const FusedLayout& L = a.lay;
    const int G = int(gridDim.x);
    if (a.peer.enabled && blockIdx.x == 0 && threadIdx.x == 0) *a.peer.epoch += 1ull;   // (every CTA read it at the start)
    if (!updated) return;                    // no update was made: x is untouched
    const R alpha = sc->alpha;
    const bool a_one = sc->a_one != 0;
    for (int c = 0; c < L.tiles_per_cta; ++c) {
        const int tile = blockIdx.x + c * G;
        if (tile >= t.n_tiles) break;
        const uint32_t node_off = t.tile_node_off[tile];
        const int n_int = int(t.tile_nint[tile]);
        for (int k = threadIdx.x; k < n_int; k += NT) {
            const SV xv = sv_ldcg(a.xt + node_off + k);
            R* d = a.x + 3 * size_t(t.tile_nodes[node_off + k]);
            d[0] = R(xv.x); d[1] = R(xv.y); d[2] = R(xv.z);
        }
    }
    for (int lu = 0; lu < L.units_per_cta; ++lu) {
        const int unit = lu * G + int(blockIdx.x);
        if (size_t(unit) * kUnit >= size_t(t.n_chunks) * kGatherChunk) break;
        for (int k = threadIdx.x; k < kUnit; k += NT) {
            const size_t slot = size_t(unit) * kUnit + k;
            const size_t g = t.sh_nodes[slot];
            if (g == 0xFFFFFFFFu) continue;
            const SV xv = sv_ldcg(a.xS + slot);
            R x0 = R(xv.x), x1 = R(xv.y), x2 = R(xv.z);
            if (pending) { const SV pv = sv_ldcg(a.pS + slot); x_one<R>(x0, R(pv.x), alpha, a_one); x_one<R>(x1, R(pv.y), alpha, a_one); x_one<R>(x2, R(pv.z), alpha, a_one); }
            R* d = a.x + 3 * g;
            d[0] = x0; d[1] = x1; d[2] = x2;
        }
    }
}

// ---- the kernel ----------------------------------------------------------------------------------------------------------------------------
// Pass: the element type's policy --
//   typedef Dev;  static const TileDev<R>& tiles(const Dev&);
//   template <int ET, class OnBoundary> static void elements(const Dev&, int tile, const SV* s_in, R* s_slot, int max_slots, unsigned char* s_extra, int arrive_at, OnBoundary f)
//       one pass over the tile's elements by the ET element threads; calls f() once (from every element thread, at the same trip count) when
//       the elements [0, arrive_at) are done (arrive_at < 0: never).
// ET element threads + GT dedicated gather threads (GT may be 0: everybody does everything).
constexpr int kTrF = kTraceTail;     // trace records of the fused kernel: same area as the first-generation kernel's
template <class R, class Pass, int ET, int GT>
__global__ void __launch_bounds__(ET + GT, 1) fused_cg_kernel(typename Pass::Dev d, FusedCG<R> a) {
    typedef typename SVec<R>::T SV;
    constexpr int NT = ET + GT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_wpart[(NT / 32) * kFusedDots];     // per element warp: running sums over the CTA's tiles
    __shared__ double s_tot[kFusedDots];
    __shared__ FusedScal<R> s_sc;
    __shared__ unsigned s_flag1, s_next_unit;
    CGDev* cg = a.cg;
    if (cg->done) return;
    const TileDev<R>& t = Pass::tiles(d);
    const FusedLayout& L = a.lay;
    const int G = int(gridDim.x);
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    SV* s_in0 = reinterpret_cast<SV*>(smem_raw);
    R* s_slot = reinterpret_cast<R*>(smem_raw + L.off_slot);
    double* s_upart = reinterpret_cast<double*>(smem_raw + L.off_upart);
    unsigned char* s_extra = smem_raw + L.off_extra;
    if (threadIdx.x == 0) { s_flag1 = 0u; s_next_unit = 0u; }
    trace_mark(a.ep.trace, kTrF, 8);
    fused_load_tables<R, NT>(t, a, smem_raw);
    __syncthreads();
    trace_mark(a.ep.trace, kTrF, 9);
    bool updated = false, pending = false;
    if (fused_init<R, NT>(t, a, &s_sc, smem_raw, s_wpart, s_tot)) {
        // units of this CTA: unit = lu * G + blockIdx.x
        const int n_units_total = t.n_chunks * (kGatherChunk / kUnit);
        const int n_my_units = (n_units_total - int(blockIdx.x) + G - 1) / G;
        const bool elem_thread = threadIdx.x < ET;
        for (;;) {
            const unsigned iter = s_sc.iter;
            // phase stamps of ONE iteration in the middle of the solve (the 10th of this launch; mark 12 = start of the 11th)
            unsigned long long* const tr = (a.ep.trace && iter == 9u) ? a.ep.trace : nullptr;
            if (a.ep.trace && iter == 10u) trace_mark(a.ep.trace, kTrF, 12);
            trace_mark(tr, kTrF, 0);
            // ---- element warps: the CTA's tiles
            if (elem_thread) {
                FusedAcc<R> acc{0.0, 0.0, 0.0, 0.0};
                bool arrived = false;
                for (int c = 0; c < L.tiles_per_cta; ++c) {
                    const int tile = blockIdx.x + c * G;
                    if (tile >= t.n_tiles) break;
                    const TileState<R> s = tile_state<R>(t, a, smem_raw, c, tile);
                    const SV* s_in = s.P;
                    if (!L.cached) {
                        for (int k = threadIdx.x; k < s.n_touched; k += ET) s_in0[k] = sv_ldcg(s.P + k);
                        uint16_t* jd = reinterpret_cast<uint16_t*>(smem_raw + L.off_jds);
                        for (int j = threadIdx.x; j <= t.maxval; j += ET) jd[j] = t.tile_jds[size_t(tile) * (t.maxval + 1) + j];
                        s_in = s_in0;
                        bar_first<ET>();
                    }
                    // the CTA's last tile lists its elements that feed shared nodes first: after them the staged contributions of this CTA are complete
                    const bool last = c + 1 == L.tiles_per_cta || tile + G >= t.n_tiles;
                    int arrive_at = -1;
                    if (last && t.tile_nb && t.tile_nb[tile] != 0xFFFFFFFFu) {
                        const int nb_up = (int(t.tile_nb[tile]) + ET - 1) / ET * ET;
                        if (nb_up + ET <= t.tile_e) arrive_at = nb_up;
                    }
                    Pass::template elements<ET>(d, tile, s_in, s_slot, L.max_slots, s_extra, arrive_at, [&]() {
                        bar_first<ET>();
                        if (threadIdx.x == ET - 32) fused_arrive_s1(a.sync);      // (the last element warp has the shortest tail of the tile)
                    });
                    if (arrive_at >= 0) arrived = true;
                    bar_first<ET>();
                    if (last && !arrived) { if (threadIdx.x == ET - 32) fused_arrive_s1(a.sync); arrived = true; }
                    if (c == 0) trace_mark(tr, kTrF, 1);
                    TileState<R> s2 = s;
                    if (!L.cached) s2.P = s_in0;        // (p of the tile's nodes is in shared memory right now)
                    fused_interior<R, ET>(a, s2, s_slot, acc);
                    bar_first<ET>();                    // the slots are free for the next tile
                    if (c == 0) trace_mark(tr, kTrF, 2);
                }
                acc_warp_sum<R>(acc);
                if ((threadIdx.x & 31) == 0) {
                    double* w = s_wpart + (threadIdx.x >> 5) * kFusedDots;
                    w[0] = acc.rr; w[1] = acc.pq; w[2] = acc.rq; w[3] = acc.qq;
                }
                trace_mark(tr, kTrF, 3);
            }
            // ---- shared nodes: wait for S1 (one poller per CTA), then take units until none is left
            if (threadIdx.x == (GT > 0 ? ET : 0)) {
                const bool ok = fused_poll_s1(a.sync, iter);
                if (!ok) s_sc.failed = 1;
                st_release_cta_shared(&s_flag1, iter + 1u);
            }
            if ((threadIdx.x & 31) == 0) { while (ld_acquire_cta_shared(&s_flag1) != iter + 1u) { } }
            __syncwarp();
            trace_mark_by(tr, kTrF, 4, GT > 0 ? ET : 0);
            for (;;) {
                int lu = 0;
                if ((threadIdx.x & 31) == 0) lu = int(atomicAdd(&s_next_unit, 1u));
                lu = __shfl_sync(0xffffffffu, lu, 0);
                if (lu >= n_my_units) break;
                fused_unit<R>(t, a, &s_sc, lu * G + int(blockIdx.x), s_upart + size_t(lu) * kFusedDots);
            }
            __syncthreads();
            pending = false;                        // (the units have applied the previous iteration's update of the shared nodes)
            trace_mark(tr, kTrF, 5);
            // ---- S2: the four dot products
            if (threadIdx.x < 32) {
                double v[kFusedDots] = {0.0, 0.0, 0.0, 0.0};
                for (int w = threadIdx.x; w < ET / 32; w += 32) {
#pragma unroll
                    for (int i = 0; i < kFusedDots; ++i) v[i] += s_wpart[w * kFusedDots + i];
                }
                for (int u = threadIdx.x; u < n_my_units; u += 32) {
#pragma unroll
                    for (int i = 0; i < kFusedDots; ++i) v[i] += s_upart[size_t(u) * kFusedDots + i];
                }
#pragma unroll
                for (int i = 0; i < kFusedDots; ++i) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], o);
                }
                fused_arrive_values(a.sync, s_sc.vsync, v);
                const bool ok = fused_wait_values(a.sync, s_sc.vsync, s_tot);
                if (threadIdx.x == 0) {
                    if (!ok) s_sc.failed = 1;
                    s_sc.vsync += 1; s_sc.iter = iter + 1u; s_next_unit = 0u;
                }
            }
            __syncthreads();
            trace_mark(tr, kTrF, 6);
            if (s_sc.failed) { if (lead) { cg->done = 1; cg->end_cond = 99; } break; }
            // ---- scalars (every thread of every CTA computes the same values from the same sums)
            const double rr = s_tot[0], den = s_tot[1], rq = s_tot[2], qq = s_tot[3];
            const double rho = rr;                  // the measured rho_k = r_k.r_k replaces last iteration's prediction
            const int it = s_sc.it;
            bool stop = false;
            if (den != 0.0) { if (fabs(den) <= s_sc.thr && !(it == 1 && s_sc.tsc == 0)) stop = true; } else stop = true;
            if (lead) {
                cg->rho = rho;
                if (cg->n_err > 0 && cg->n_err <= kMaxGraph) cg->graph_error[cg->n_err - 1] = sqrt(rho) / s_sc.normb;
                cg_after_den(cg, den);
            }
            if (stop) break;
            const double alpha_d = rho / den;
            const R alpha = R(alpha_d), malpha = R(-alpha_d);
            // r' = r + q * malpha  =>  r'.r' = r.r + 2 malpha r.q + malpha^2 q.q  (the update uses malpha as rounded to Real)
            const double ma = double(malpha);
            double rho_new = rho + 2.0 * ma * rq + ma * ma * qq;
            if (!(rho_new > 0.0)) rho_new = 0.0;
            const int it2 = it + 1;
            bool stop2 = unsigned(it2) > s_sc.max_iter;
            if (!stop2) { const double err = sqrt(rho_new) / s_sc.normb; if (err <= s_sc.tol && !(it2 == 1 && s_sc.tsc == 0)) stop2 = true; }
            if (lead) cg_after_rho(cg, rho_new);
            __syncthreads();                        // everybody has read the scalars of the previous iteration
            if (threadIdx.x == 0) {
                s_sc.alpha = alpha; s_sc.malpha = malpha; s_sc.a_one = alpha_d == 1.0 ? 1 : 0; s_sc.ma_one = -alpha_d == 1.0 ? 1 : 0;
                s_sc.beta = R(rho_new / rho); s_sc.rho = rho_new; s_sc.it = it2; s_sc.first = 0;
            }
            __syncthreads();
            // ---- local update of x, r (and p unless the solve is over)
            fused_update<R, NT>(t, a, &s_sc, smem_raw, !stop2);
            updated = true; pending = true;
            if (stop2) break;
            __syncthreads();
            trace_mark(tr, kTrF, 13);
        }
    }
    __syncthreads();
    fused_finish<R, NT>(t, a, &s_sc, smem_raw, updated, pending);
    trace_mark(a.ep.trace, kTrF, 11);
}

}  // namespace sb
