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
    L.off_upart = unsigned(o); o = up(o + sizeof(double) * kFusedDots * 2 * size_t(std::max(units_per_cta, 1)));   // [2][units]: second half = interface lanes (multi-GPU)
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
    int n_if_units;         // multi-GPU: the partition-interface nodes are the first shared nodes; units [0, n_if_units) hold them
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
    unsigned long long seq_base;    // multi-GPU: sequence numbers of this launch start here (launches so far * 65536)
    int n_err, n_den;               // entries of CGDev::graph_error / graph_den so far (the lead thread records with plain stores, no read-modify-write)
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
    __syncwarp();   // reconverge first: after a loop whose trip count differs between lanes the shuffles would take the per-shuffle WARPSYNC slow path (measured: 4.6 us for a 4 x 5 double tree)
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
template <class R, bool CACHED> __device__ __forceinline__ TileState<R> tile_state(const TileDev<R>& t, const FusedCG<R>& a, unsigned char* smem_raw, int c, int tile) {
    typedef typename SVec<R>::T SV;
    const FusedLayout& L = a.lay;
    TileState<R> s;
    s.node_off = t.tile_node_off[tile];
    s.n_touched = int(t.tile_node_off[tile + 1] - s.node_off); s.n_int = int(t.tile_nint[tile]);
    if constexpr (CACHED) {     // (compile-time: the pointers keep their shared-memory address space => LDS/STS, not generic accesses)
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
// Wait until *counter >= target; every lane of the calling warp runs the loop (uniform control flow), ~4 s time-out.
__device__ __forceinline__ bool poll_counter_warp(const unsigned* counter, unsigned target) {
    const long long t0 = poll_clock();
    bool ok = true;
    for (;;) {
        if (ld_acquire_u32(counter) >= target) break;
        if (poll_clock() - t0 > kSyncTimeoutCycles) { ok = false; break; }
    }
    return ok;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fused_arrive_values(unsigned long long* sync, unsigned vs, const double v[kFusedDots]) {
    const unsigned G = gridDim.x;
    double2* cur = reinterpret_cast<double2*>(reinterpret_cast<double*>(sync) + (size_t(vs % 3) * G + blockIdx.x) * kFusedDots);
    unsigned* counter = reinterpret_cast<unsigned*>(sync + size_t(3) * G * kFusedDots + 32);
    if (threadIdx.x == 0) {
        __stcg(cur, make_double2(v[0], v[1])); __stcg(cur + 1, make_double2(v[2], v[3]));
        red_release_gpu_add(counter, 1u);        // release: the CTA's writes (ordered before this thread by the CTA barrier) and the values above
    }
}
// wait (warp 0) + fixed-order sum over the CTAs; the totals are left in out[] (shared memory) for the CTA-wide barrier that follows.
// The values of up to 160 CTAs are requested at once (one L2 round trip), then added in a fixed order.
__device__ __forceinline__ bool fused_wait_values(unsigned long long* sync, unsigned vs, double* out /* smem [kFusedDots] */) {
    const unsigned G = gridDim.x;
    const double2* cur = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(sync) + size_t(vs % 3) * G * kFusedDots);
    const unsigned* counter = reinterpret_cast<const unsigned*>(sync + size_t(3) * G * kFusedDots + 32);
    // (all 32 lanes poll together -- one transaction per trip, the same value for every lane: a loop run by lane 0 alone leaves that lane
    // split from its warp for good on this compiler, and every later shuffle of the warp then takes the WARPSYNC slow path)
    const bool ok = poll_counter_warp(counter, (vs + 1) * G);
    double v[kFusedDots];
#pragma unroll
    for (int i = 0; i < kFusedDots; ++i) v[i] = 0.0;
    for (unsigned c0 = 0; c0 < G; c0 += 160u) {
        double2 lo[5], hi[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const unsigned c = c0 + threadIdx.x + 32u * u;
            lo[u] = make_double2(0.0, 0.0); hi[u] = lo[u];
            if (c < G) { lo[u] = __ldcg(cur + 2 * size_t(c)); hi[u] = __ldcg(cur + 2 * size_t(c) + 1); }
        }
#pragma unroll
        for (int u = 0; u < 5; ++u) { v[0] += lo[u].x; v[1] += lo[u].y; v[2] += hi[u].x; v[3] += hi[u].y; }
    }
    __syncwarp();
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
// ---- the same across the GPUs of one NVSwitch node (peer memory, see cg_persist.cuh) ---------------------------------------------------
// Local arrival as above; CTA 0 is this GPU's leader: once the local CTAs have arrived it adds their values and stores the four sums into
// slot [rank] of EVERY rank's table (lane r serves rank r: the W remote stores leave together), as 8-byte words carrying 32 bits of payload
// and the 32-bit sequence number.  Warp 0 of every CTA then reads the W slots of its own GPU (lane r polls rank r) and adds them in rank
// order: same operands, same order, same bits on every GPU.  The leader's own slot is stored last with release (it publishes this GPU's
// CTAs' writes to the local pollers, which fence after seeing it).
template <class R> __device__ __forceinline__ bool fused_sync_values_dist(const FusedCG<R>& a, unsigned vs, unsigned long long seq64, const double v[kFusedDots], double* out) {
    const PeerDev<R>& P = a.peer;
    const unsigned G = gridDim.x;
    const unsigned seq = unsigned(seq64);
    const int set = int(seq64 & 1ull), lane = int(threadIdx.x & 31);
    fused_arrive_values(a.sync, vs, v);
    bool ok = true;
    if (blockIdx.x == 0) {
        double tot[kFusedDots];
        ok = fused_wait_values(a.sync, vs, out);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < kFusedDots; ++i) tot[i] = out[i];
        __syncwarp();
        if (lane < P.world) {
            unsigned long long* dst = (lane == P.rank ? P.ar4 : P.peer_ar4[lane]) + (size_t(set) * kMaxPeers + P.rank) * 8;
#pragma unroll
            for (int i = 0; i < kFusedDots; ++i) {
                const unsigned long long b = (unsigned long long)__double_as_longlong(tot[i]);
                const unsigned long long w0 = (b & 0xFFFFFFFFull) | ((unsigned long long)seq << 32), w1 = (b >> 32) | ((unsigned long long)seq << 32);
                if (lane == P.rank) { if (i < kFusedDots - 1) { dst[2 * i] = w0; dst[2 * i + 1] = w1; } else { dst[2 * i] = w0; st_release_gpu_u64(dst + 2 * i + 1, w1); } }
                else { st_relaxed_sys_u64(dst + 2 * i, w0); st_relaxed_sys_u64(dst + 2 * i + 1, w1); }
            }
        }
    }
    __syncwarp();
    double val[kFusedDots] = {0.0, 0.0, 0.0, 0.0};
    if (lane < P.world) {
        const unsigned long long* src = P.ar4 + (size_t(set) * kMaxPeers + lane) * 8;
        const long long t0 = poll_clock();
        for (;;) {
            unsigned long long w[8];
            bool all = true;
#pragma unroll
            for (int i = 0; i < 8; ++i) { w[i] = ld_relaxed_sys_u64(src + i); all = all && unsigned(w[i] >> 32) == seq; }
            if (all) {
#pragma unroll
                for (int i = 0; i < kFusedDots; ++i) val[i] = __longlong_as_double((long long)((w[2 * i] & 0xFFFFFFFFull) | (w[2 * i + 1] << 32)));
                break;
            }
            if (poll_clock() - t0 > kSyncTimeoutCycles) { ok = false; break; }
        }
    }
    __syncwarp();
    asm volatile("fence.acq_rel.gpu;" ::: "memory");       // acquire side of the leader's release (lighter than the sequentially consistent __threadfence)
    ok = __all_sync(0xffffffffu, ok ? 1 : 0) != 0;
    double tot[kFusedDots] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < P.world; ++r) {
#pragma unroll
        for (int i = 0; i < kFusedDots; ++i) tot[i] += __shfl_sync(0xffffffffu, val[i], r);      // rank order
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kFusedDots; ++i) out[i] = tot[i];
    }
    return ok;
}
// S1: "this CTA's staged contributions are complete" -- arrival only (the waiters are the threads that sum shared nodes)
__device__ __forceinline__ void fused_arrive_s1(unsigned long long* sync) {
    unsigned* counter = reinterpret_cast<unsigned*>(sync + size_t(3) * gridDim.x * kFusedDots);
    red_release_gpu_add(counter, 1u);
}
// (called by all 32 lanes of one warp)
__device__ __forceinline__ bool fused_poll_s1(const unsigned long long* sync, unsigned iter) {
    const unsigned* counter = reinterpret_cast<const unsigned*>(sync + size_t(3) * gridDim.x * kFusedDots);
    return poll_counter_warp(counter, (iter + 1) * gridDim.x);
}

// ---- once per solve: static tables ------------------------------------------------------------------------------------------------
template <class R, int NT, bool CACHED> __device__ __forceinline__ void fused_load_tables(const TileDev<R>& t, const FusedCG<R>& a, unsigned char* smem_raw) {
    const FusedLayout& L = a.lay;
    const NodeEpilogue<R>& ep = a.ep;
    const int G = int(gridDim.x);
    for (int c = 0; c < L.tiles_per_cta; ++c) {
        const int tile = blockIdx.x + c * G;
        if (tile >= t.n_tiles) break;
        const uint32_t node_off = t.tile_node_off[tile];
        const int n_touched = int(t.tile_node_off[tile + 1] - node_off), n_int = int(t.tile_nint[tile]);
        NodeRec<R>* nrec = CACHED ? reinterpret_cast<NodeRec<R>*>(smem_raw + L.off_nrec) + c * L.max_int : a.gNrec + node_off;
        for (int k = threadIdx.x; k < n_touched; k += NT) {
            if (k < n_int) {
                const uint32_t g = t.tile_nodes[node_off + k];
                nrec[k] = NodeRec<R>{g, unsigned(t.tile_val[node_off + k]) | ((ep.fixed && ep.fixed[g]) ? 0x10000u : 0u), ep.mass ? ep.mass[g] : R(0)};
            } else if (CACHED) reinterpret_cast<uint32_t*>(smem_raw + L.off_tidx)[c * L.max_shtouch + (k - n_int)] = t.tile_shslot[node_off + k];
        }
        if (CACHED)
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
template <class R, int NT, bool CACHED> __device__ __forceinline__ bool fused_init(const TileDev<R>& t, const FusedCG<R>& a, FusedScal<R>* sc, unsigned char* smem_raw, double* s_wpart, double* s_tot) {
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
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sb += __shfl_down_sync(0xffffffffu, sb, o); sr += __shfl_down_sync(0xffffffffu, sr, o); }
    if ((threadIdx.x & 31) == 0) { s_wpart[(threadIdx.x >> 5) * kFusedDots] = sb; s_wpart[(threadIdx.x >> 5) * kFusedDots + 1] = sr; }
    // p = r, private copies of x and r
    for (int c = 0; c < L.tiles_per_cta; ++c) {
        const int tile = blockIdx.x + c * G;
        if (tile >= t.n_tiles) break;
        const TileState<R> s = tile_state<R, CACHED>(t, a, smem_raw, c, tile);
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
        __syncwarp();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v[0] += __shfl_down_sync(0xffffffffu, v[0], o); v[1] += __shfl_down_sync(0xffffffffu, v[1], o); }
        if (P.enabled) ok = fused_sync_values_dist<R>(a, 0u, (*P.epoch) * 65536ull + 1ull, v, s_tot);
        else { fused_arrive_values(a.sync, 0u, v); ok = fused_wait_values(a.sync, 0u, s_tot); }
    }
    __syncthreads();
    const double nb2 = s_tot[0], rho0 = s_tot[1];
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    if (threadIdx.x == 0) {
        sc->vsync = 1; sc->iter = 0; sc->first = 1; sc->failed = ok ? 0 : 1;
        sc->normb = sqrt(nb2); sc->rho = rho0; sc->it = 1;
        sc->tol = cg->tolerance; sc->thr = cg->threshold; sc->tsc = cg->time_step_count; sc->max_iter = cg->max_iter;
        sc->alpha = R(0); sc->malpha = R(0); sc->beta = R(0); sc->a_one = 0; sc->ma_one = 0;
        sc->n_err = 2; sc->n_den = 0;
        sc->seq_base = P.enabled ? (*P.epoch) * 65536ull : 0ull;       // (cg_begin_kernel: one entry; cg_after_rho(rho0) below: the second)
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
            for (int u = 0; u < 4; ++u) { const int sl = s.jds[jj + u] + k; cx[u] = s_slot[3 * sl]; cy[u] = s_slot[3 * sl + 1]; cz[u] = s_slot[3 * sl + 2]; }
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
            if (plus) { ax += s_slot[3 * sl]; ay += s_slot[3 * sl + 1]; az += s_slot[3 * sl + 2]; }
            else { ax -= s_slot[3 * sl]; ay -= s_slot[3 * sl + 1]; az -= s_slot[3 * sl + 2]; }
        }
        node_finish_m(ep, rec.g, rec.mass, (rec.val_fixed & 0x10000u) != 0, R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
        s.Q[3 * k] = ax; s.Q[3 * k + 1] = ay; s.Q[3 * k + 2] = az;
        if (counted_all || a.peer.owned[rec.g]) acc_node<R>(acc, R(rv.x), R(rv.y), R(rv.z), R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
    }
}

// ---- one unit of 32 shared nodes (one warp, lane = node) ---------------------------------------------------------------------------------
// Applies the pending update of the previous iteration, sums the staged contributions in element order, publishes r and q for the tiles.
template <class R> struct GatherBatch;
template <> struct GatherBatch<float> { static constexpr int N = 6; };
template <> struct GatherBatch<double> { static constexpr int N = 4; };
// what a unit needs that does not depend on S1: requested before the wait, so its L2 latency hides behind it
template <class R> struct UnitPre { GRec<R> rec; typename SVec<R>::T pv, rv, xv, qo; };
template <class R> __device__ __forceinline__ void fused_unit_preload(const FusedCG<R>& a, const FusedScal<R>* sc, int unit, UnitPre<R>& u) {
    const size_t slot = size_t(unit) * kUnit + (threadIdx.x & 31);
    u.rec = a.shrec[slot];
    u.pv = sv_ldcg(a.pS + slot); u.rv = sv_ldcg(a.rS + slot);
    if (!sc->first) { u.xv = sv_ldcg(a.xS + slot); u.qo = sv_ldcg(a.qS + slot); }
    else { u.xv = u.pv; u.qo = u.pv; }
}
template <class R> __device__ __forceinline__ void fused_unit(const TileDev<R>& t, const FusedCG<R>& a, const FusedScal<R>* sc, int unit, const UnitPre<R>& pre, unsigned long long hseq64, double* upart /* smem [kFusedDots] */) {
    typedef typename SVec<R>::T SV;
    constexpr int B = GatherBatch<R>::N;
    const NodeEpilogue<R>& ep = a.ep;
    const int lane = threadIdx.x & 31;
    const size_t slot = size_t(unit) * kUnit + lane;
    FusedAcc<R> acc{0.0, 0.0, 0.0, 0.0};
    const GRec<R> rec = pre.rec;
    if (rec.g != 0xFFFFFFFFu) {
        const int val = int(rec.val_fixed & 0xFFFFu);
        const Quad<R>* stg = t.stage + rec.base;
        const uint64_t pol = l2_policy_evict_first();
        const Quad<R> zero{R(0), R(0), R(0), R(0)};
        R p0 = R(pre.pv.x), p1 = R(pre.pv.y), p2 = R(pre.pv.z), r0 = R(pre.rv.x), r1 = R(pre.rv.y), r2 = R(pre.rv.z);
        if (!sc->first) {
            // x += alpha p ; r -= alpha q ; p = p beta + r   of the previous iteration (cgstep_alpha, cgstep_beta).  A dozen ALU operations on
            // preloaded values: done before the staged contributions are requested so that x and the old q leave the registers first.
            R x0 = R(pre.xv.x), x1 = R(pre.xv.y), x2 = R(pre.xv.z);
            const R alpha = sc->alpha, malpha = sc->malpha, beta = sc->beta;
            const bool a_one = sc->a_one != 0, ma_one = sc->ma_one != 0;
            x_one<R>(x0, p0, alpha, a_one); x_one<R>(x1, p1, alpha, a_one); x_one<R>(x2, p2, alpha, a_one);
            r_one<R>(r0, R(pre.qo.x), malpha, ma_one); r_one<R>(r1, R(pre.qo.y), malpha, ma_one); r_one<R>(r2, R(pre.qo.z), malpha, ma_one);
            p0 = p_update<R>(p0, beta, r0); p1 = p_update<R>(p1, beta, r1); p2 = p_update<R>(p2, beta, r2);
            stcg_sv(a.xS + slot, SVec<R>::make(x0, x1, x2));
            stcg_sv(a.rS + slot, SVec<R>::make(r0, r1, r2));
            stcg_sv(a.pS + slot, SVec<R>::make(p0, p1, p2));
        }
        Quad<R> b0[B], b1[B];
#pragma unroll
        for (int u = 0; u < B; ++u) b0[u] = u < val ? stage_load(stg + size_t(u) * kGatherChunk, pol) : zero;
#pragma unroll
        for (int u = 0; u < B; ++u) b1[u] = B + u < val ? stage_load(stg + size_t(B + u) * kGatherChunk, pol) : zero;
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
        const PeerDev<R>& P = a.peer;
        const int if_row = P.enabled ? P.sh_if_row[slot] : -1;
        if (if_row >= 0) {
            // multi-GPU, a node of the partition interface: (q) is this rank's PARTIAL sum.  Every other sharing rank gets it in its inbox now
            // (NVLink stores); the node is finished in the second pass (fused_unit_interface), after the warp's other units.
            const unsigned hseq = unsigned(hseq64);
            for (int e = 0; e < P.max_sh - 1; ++e) {
                const int2 to = P.if_send[if_row * (P.max_sh - 1) + e];
                if (to.x >= 0) {
                    unsigned long long* d = P.nb_inbox[to.x] + size_t(InboxWords<R>::N) * size_t(to.y);
                    inbox_put(d, 0, q0, hseq); inbox_put(d, 1, q1, hseq); inbox_put(d, 2, q2, hseq);
                }
            }
        } else acc_node<R>(acc, r0, r1, r2, p0, p1, p2, q0, q1, q2);
    }
    acc_warp_sum<R>(acc);
    if (lane == 0) { upart[0] = acc.rr; upart[1] = acc.pq; upart[2] = acc.rq; upart[3] = acc.qq; }
}
// second pass over a unit that holds interface nodes: q = the sharing ranks' partial sums in ascending rank order (own partial at its rank's
// place) -- same operands, same order, same bits on every rank; the rank that owns the node counts it in the dot products.
template <class R> __device__ __forceinline__ bool fused_unit_interface(const FusedCG<R>& a, int unit, unsigned long long hseq64, double* upart2 /* smem [kFusedDots] */) {
    typedef typename SVec<R>::T SV;
    const PeerDev<R>& P = a.peer;
    const int lane = threadIdx.x & 31;
    const size_t slot = size_t(unit) * kUnit + lane;
    FusedAcc<R> acc{0.0, 0.0, 0.0, 0.0};
    bool ok = true;
    const GRec<R> rec = a.shrec[slot];
    const int if_row = rec.g != 0xFFFFFFFFu ? P.sh_if_row[slot] : -1;
    if (if_row >= 0) {
        const SV pv = sv_ldcg(a.pS + slot), rv = sv_ldcg(a.rS + slot), qv = sv_ldcg(a.qS + slot);
        const unsigned hseq = unsigned(hseq64);
        R s0 = R(0), s1 = R(0), s2 = R(0);
        for (int j = 0; j < P.max_sh; ++j) {
            const int sj = P.src[if_row * P.max_sh + j];
            R c0 = R(0), c1 = R(0), c2 = R(0);
            if (sj == -1) { c0 = R(qv.x); c1 = R(qv.y); c2 = R(qv.z); }
            else if (sj >= 0) {
                const unsigned long long* row = P.inbox + size_t(InboxWords<R>::N) * size_t(sj);
                const long long t0 = poll_clock();
                while (!(inbox_get(row, 0, hseq, c0) && inbox_get(row, 1, hseq, c1) && inbox_get(row, 2, hseq, c2)))
                    if (poll_clock() - t0 > kSyncTimeoutCycles) { ok = false; break; }
            }
            if (j == 0) { s0 = c0; s1 = c1; s2 = c2; } else { s0 += c0; s1 += c1; s2 += c2; }
        }
        stcg_sv(a.qS + slot, SVec<R>::make(s0, s1, s2));
        if (P.owned[rec.g]) acc_node<R>(acc, R(rv.x), R(rv.y), R(rv.z), R(pv.x), R(pv.y), R(pv.z), s0, s1, s2);
    }
    acc_warp_sum<R>(acc);
    if (lane == 0) { upart2[0] = acc.rr; upart2[1] = acc.pq; upart2[2] = acc.rq; upart2[3] = acc.qq; }
    return __all_sync(0xffffffffu, ok ? 1 : 0) != 0;
}

// ---- after S2: x, r, p of every node the CTA's tiles touch ----------------------------------------------------------------------------------
// with_p = false: the last iteration of the solve (only x matters; r is updated too so that the vector the caller sees is consistent)
template <class R, int NT, bool CACHED> __device__ __forceinline__ void fused_update(const TileDev<R>& t, const FusedCG<R>& a, const FusedScal<R>* sc, unsigned char* smem_raw, bool with_p) {
    typedef typename SVec<R>::T SV;
    const FusedLayout& L = a.lay;
    const int G = int(gridDim.x);
    const R alpha = sc->alpha, malpha = sc->malpha, beta = sc->beta;
    const bool a_one = sc->a_one != 0, ma_one = sc->ma_one != 0;
    constexpr int U = 2;            // rounds per tile whose L2 requests are all issued before the first use
    // interior: x, r (private arrays) ; shared: r, q (the owner's published values)
    auto request = [&](const TileState<R>& s, int k, SV& va, SV& vb) {
        if (k < s.n_int) { va = sv_ldcg(a.xt + s.node_off + k); vb = sv_ldcg(a.rt + s.node_off + k); }
        else if (k < s.n_touched) { const size_t slot = s.tidx[k - s.n_int]; va = sv_ldcg(a.rS + slot); vb = sv_ldcg(a.qS + slot); }
    };
    auto apply = [&](const TileState<R>& s, int k, const SV& va, const SV& vb) {
        if (k >= s.n_touched) return;
        const SV pv = s.P[k];
        R p0 = R(pv.x), p1 = R(pv.y), p2 = R(pv.z), r0, r1, r2;
        if (k < s.n_int) {
            R x0 = R(va.x), x1 = R(va.y), x2 = R(va.z);
            r0 = R(vb.x); r1 = R(vb.y); r2 = R(vb.z);
            x_one<R>(x0, p0, alpha, a_one); x_one<R>(x1, p1, alpha, a_one); x_one<R>(x2, p2, alpha, a_one);
            r_one<R>(r0, s.Q[3 * k], malpha, ma_one); r_one<R>(r1, s.Q[3 * k + 1], malpha, ma_one); r_one<R>(r2, s.Q[3 * k + 2], malpha, ma_one);
            stcg_sv(a.xt + s.node_off + k, SVec<R>::make(x0, x1, x2));
            stcg_sv(a.rt + s.node_off + k, SVec<R>::make(r0, r1, r2));
        } else {
            r0 = R(va.x); r1 = R(va.y); r2 = R(va.z);
            r_one<R>(r0, R(vb.x), malpha, ma_one); r_one<R>(r1, R(vb.y), malpha, ma_one); r_one<R>(r2, R(vb.z), malpha, ma_one);
        }
        if (with_p) s.P[k] = SVec<R>::make(p_update<R>(p0, beta, r0), p_update<R>(p1, beta, r1), p_update<R>(p2, beta, r2));
    };
    // two tiles at a time: the requests of both are in flight together (one L2 round trip for a CTA with two ~900-node tiles)
    for (int c = 0; c < L.tiles_per_cta; c += 2) {
        const int tile0 = blockIdx.x + c * G, tile1 = tile0 + G;
        if (tile0 >= t.n_tiles) break;
        const bool two = c + 1 < L.tiles_per_cta && tile1 < t.n_tiles;
        const TileState<R> s0 = tile_state<R, CACHED>(t, a, smem_raw, c, tile0);
        TileState<R> s1 = s0;
        if (two) s1 = tile_state<R, CACHED>(t, a, smem_raw, c + 1, tile1); else s1.n_touched = 0;
        const int nmax = max(s0.n_touched, s1.n_touched);
        for (int k0 = threadIdx.x; k0 < nmax; k0 += U * NT) {
            SV va0[U], vb0[U], va1[U], vb1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { request(s0, k0 + u * NT, va0[u], vb0[u]); request(s1, k0 + u * NT, va1[u], vb1[u]); }
#pragma unroll
            for (int u = 0; u < U; ++u) { apply(s0, k0 + u * NT, va0[u], vb0[u]); apply(s1, k0 + u * NT, va1[u], vb1[u]); }
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
//   typedef First;  static void prefetch(const Dev&, int tile, First&)   requests what the thread needs for its first element of the tile
//   template <int ET, class OnBoundary> static void elements(const Dev&, int tile, const SV* s_in, R* s_slot, int max_slots, unsigned char* s_extra, int arrive_at, OnBoundary f, const First&)
//       one pass over the tile's elements by the ET element threads; calls f() once (from every element thread, at the same trip count) when
//       the elements [0, arrive_at) are done (arrive_at < 0: never).
// ET element threads + GT dedicated gather threads (GT may be 0: everybody does everything).  The CTA's units of shared nodes are dealt
// statically: the first `ded_units` to the dedicated warps, the others to the element warps (which take them once their tiles are done) --
// a fixed assignment keeps every partial sum, hence every bit of the result, independent of timing.
constexpr int kTrF = kTraceTail;     // trace records of the fused kernel: same area as the first-generation kernel's
template <class R, class Pass, int ET, int GT, bool CACHED>
__global__ void __launch_bounds__(ET + GT, 1) fused_cg_kernel(typename Pass::Dev d, FusedCG<R> a, int ded_share_pct) {
    typedef typename SVec<R>::T SV;
    constexpr int NT = ET + GT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_wpart[(NT / 32) * kFusedDots];     // per element warp: sums over the CTA's tiles
    __shared__ double s_tot[kFusedDots];
    __shared__ FusedScal<R> s_sc;
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
    trace_mark(a.ep.trace, kTrF, 8);
    fused_load_tables<R, NT, CACHED>(t, a, smem_raw);
    __syncthreads();
    trace_mark(a.ep.trace, kTrF, 9);
    bool updated = false, pending = false;
    if (fused_init<R, NT, CACHED>(t, a, &s_sc, smem_raw, s_wpart, s_tot)) {
        // units of this CTA: unit = lu * G + blockIdx.x, lu < n_my_units
        const int n_units_total = t.n_chunks * (kGatherChunk / kUnit);
        const int n_my_units = (n_units_total - int(blockIdx.x) + G - 1) / G;
        const int n_if_local = a.peer.enabled ? max(0, (a.n_if_units - int(blockIdx.x) + G - 1) / G) : 0;
        const int ded_units = GT > 0 ? min(n_my_units, (n_my_units * ded_share_pct + 99) / 100) : 0;
        const bool elem_thread = threadIdx.x < ET;
        const int warp = threadIdx.x >> 5;
        // this warp's units: lu = lu0, lu0 + lu_step, ... < lu_end
        // the record of the thread's first element of the CTA's first tile is requested before the phase that precedes the tile
        typename Pass::First first;
        if (elem_thread) Pass::prefetch(d, int(blockIdx.x), first);
        const int lu0 = elem_thread ? ded_units + warp : (warp - ET / 32), lu_step = elem_thread ? ET / 32 : GT / 32, lu_end = elem_thread ? n_my_units : ded_units;
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
                    const TileState<R> s = tile_state<R, CACHED>(t, a, smem_raw, c, tile);
                    const SV* s_in = s.P;
                    if (!CACHED) {
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
                    }, first);
                    if (c + 1 < L.tiles_per_cta && tile + G < t.n_tiles) Pass::prefetch(d, tile + G, first);     // the next tile's, under this tile's interior sums
                    if (arrive_at >= 0) arrived = true;
                    bar_first<ET>();
                    if (last && !arrived) { if (threadIdx.x == ET - 32) fused_arrive_s1(a.sync); arrived = true; }
                    if (c == 0) trace_mark(tr, kTrF, 1);
                    TileState<R> s2 = s;
                    if (!CACHED) s2.P = s_in0;        // (p of the tile's nodes is in shared memory right now)
                    fused_interior<R, ET>(a, s2, s_slot, acc);
                    bar_first<ET>();                    // the slots are free for the next tile
                    if (c == 0) trace_mark(tr, kTrF, 2);
                }
                acc_warp_sum<R>(acc);
                if ((threadIdx.x & 31) == 0) {
                    double* w = s_wpart + warp * kFusedDots;
                    w[0] = acc.rr; w[1] = acc.pq; w[2] = acc.rq; w[3] = acc.qq;
                }
                trace_mark(tr, kTrF, 3);
            }
            // ---- shared nodes.  What the warp's first unit needs from the previous iteration is requested before the wait for S1.
            UnitPre<R> pre;
            if (lu0 < lu_end) fused_unit_preload<R>(a, &s_sc, lu0 * G + int(blockIdx.x), pre);
            if (elem_thread) {
                if (threadIdx.x < 32) { const bool ok = fused_poll_s1(a.sync, iter); if (!ok && threadIdx.x == 0) s_sc.failed = 1; }      // (normally complete long ago when GT > 0)
                bar_first<ET>();
                trace_mark(tr, kTrF, 4);
            } else {
                if (threadIdx.x < ET + 32) { const bool ok = fused_poll_s1(a.sync, iter); if (!ok && threadIdx.x == ET) s_sc.failed = 1; }
                asm volatile("bar.sync 2, %0;" :: "n"(GT > 0 ? GT : 32) : "memory");
                trace_mark_by(tr, kTrF, 14, ET);
            }
            for (int lu = lu0; lu < lu_end; lu += lu_step) {
                if (lu != lu0) fused_unit_preload<R>(a, &s_sc, lu * G + int(blockIdx.x), pre);
                fused_unit<R>(t, a, &s_sc, lu * G + int(blockIdx.x), pre, s_sc.seq_base + iter + 1ull, s_upart + size_t(lu) * kFusedDots);
            }
            // multi-GPU: second pass over the warp's units that hold interface nodes (their partial sums left for the other GPUs in the first)
            for (int lu = lu0; lu < lu_end && lu < n_if_local; lu += lu_step) {
                const bool ok = fused_unit_interface<R>(a, lu * G + int(blockIdx.x), s_sc.seq_base + iter + 1ull, s_upart + (size_t(L.units_per_cta) + lu) * kFusedDots);
                if (!ok && (threadIdx.x & 31) == 0) s_sc.failed = 1;
            }
            if (GT > 0 && !elem_thread) trace_mark_by(tr, kTrF, 15, ET);
            __syncthreads();
            pending = false;                        // (the units have applied the previous iteration's update of the shared nodes)
            trace_mark(tr, kTrF, 5);
            // ---- S2: the four dot products
            if (threadIdx.x < 32) {
                double v[kFusedDots] = {0.0, 0.0, 0.0, 0.0};
                // (uniform trip counts: a loop whose trip count differs between lanes can leave the warp split when it reaches the shuffles,
                // which then take the per-shuffle WARPSYNC slow path)
                for (int w0 = 0; w0 < ET / 32; w0 += 32) {
                    const int w = w0 + int(threadIdx.x);
                    if (w < ET / 32) {
#pragma unroll
                        for (int i = 0; i < kFusedDots; ++i) v[i] += s_wpart[w * kFusedDots + i];
                    }
                }
                for (int u0 = 0; u0 < n_my_units; u0 += 32) {
                    const int u = u0 + int(threadIdx.x);
                    if (u < n_my_units) {
#pragma unroll
                        for (int i = 0; i < kFusedDots; ++i) v[i] += s_upart[size_t(u) * kFusedDots + i];
                    }
                }
                for (int u0 = 0; u0 < n_if_local; u0 += 32) {
                    const int u = u0 + int(threadIdx.x);
                    if (u < n_if_local) {
#pragma unroll
                        for (int i = 0; i < kFusedDots; ++i) v[i] += s_upart[(size_t(L.units_per_cta) + u) * kFusedDots + i];
                    }
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kFusedDots; ++i) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], o);
                }
                trace_mark(tr, kTrF, 10);
                bool ok;
                if (a.peer.enabled) ok = fused_sync_values_dist<R>(a, s_sc.vsync, s_sc.seq_base + s_sc.vsync + 1ull, v, s_tot);
                else {
                    fused_arrive_values(a.sync, s_sc.vsync, v);
                    if (GT == 0) trace_mark(tr, kTrF, 14);
                    ok = fused_wait_values(a.sync, s_sc.vsync, s_tot);
                }
                if (GT == 0) trace_mark(tr, kTrF, 15);
                if (threadIdx.x == 0) {
                    if (!ok) s_sc.failed = 1;
                    s_sc.vsync += 1; s_sc.iter = iter + 1u;
                }
            }
            __syncthreads();
            trace_mark(tr, kTrF, 6);
            if (s_sc.failed) { if (lead) { cg->done = 1; cg->end_cond = 99; } break; }
            // ---- scalars (every thread of every CTA computes the same values from the same sums).  The lead thread records them in
            // CGDev with plain stores -- no read-modify-write of global memory on anybody's path (cg_after_den / cg_after_rho semantics).
            const double rr = s_tot[0], den = s_tot[1], rq = s_tot[2], qq = s_tot[3];
            const double rho = rr;                  // the measured rho_k = r_k.r_k replaces last iteration's prediction
            const int it = s_sc.it;
            const double normb = s_sc.normb;
            bool stop = false;
            if (den != 0.0) { if (fabs(den) <= s_sc.thr && !(it == 1 && s_sc.tsc == 0)) stop = true; } else stop = true;
            int n_err = s_sc.n_err, n_den = s_sc.n_den;
            if (lead) {
                if (n_err >= 1 && n_err <= kMaxGraph) cg->graph_error[n_err - 1] = sqrt(rho) / normb;
                if (n_den < kMaxGraph) { cg->graph_den[n_den] = den; ++n_den; cg->n_den = n_den; }
                cg->den = den;
                if (stop) { cg->rho = rho; cg->done = 1; cg->nb_iter = it; cg->end_cond = den != 0.0 ? 2 : 3; }
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
            const double err2 = sqrt(rho_new) / normb;
            const bool tol_hit = !stop2 && err2 <= s_sc.tol && !(it2 == 1 && s_sc.tsc == 0);
            if (lead) {
                cg->alpha = alpha_d; cg->rho_1 = rho; cg->rho = rho_new;
                if (stop2) { cg->done = 1; cg->nb_iter = it2; cg->end_cond = 0; }
                else {
                    cg->it = it2;
                    if (n_err < kMaxGraph) { cg->graph_error[n_err] = err2; ++n_err; cg->n_err = n_err; }
                    if (tol_hit) { cg->done = 1; cg->nb_iter = it2; cg->end_cond = 1; }
                }
            }
            stop2 = stop2 || tol_hit;
            __syncthreads();                        // everybody has read the scalars of the previous iteration
            if (threadIdx.x == 0) {
                s_sc.alpha = alpha; s_sc.malpha = malpha; s_sc.a_one = alpha_d == 1.0 ? 1 : 0; s_sc.ma_one = -alpha_d == 1.0 ? 1 : 0;
                s_sc.beta = R(rho_new / rho); s_sc.rho = rho_new; s_sc.it = it2; s_sc.first = 0; s_sc.n_err = n_err; s_sc.n_den = n_den;
            }
            __syncthreads();
            // ---- local update of x, r (and p unless the solve is over); the next iteration's first element record travels meanwhile
            if (!stop2 && elem_thread) Pass::prefetch(d, int(blockIdx.x), first);
            fused_update<R, NT, CACHED>(t, a, &s_sc, smem_raw, !stop2);
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
