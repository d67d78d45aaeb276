// Pieces of the persistent CG kernel (the whole CGLinearSolver::solve loop in one cooperative launch) that do not depend
// on the element type: the grid barrier, the p-update fused into the nodal staging of a tile, the shared-node phase, and
// the x/r update with its reductions.  The element pass itself is supplied by tet_kernels.cuh / hex_kernels.cuh.
#pragma once
#include "fem_layout.cuh"

namespace sb {

// ---- small helpers shared with the multi-kernel CG path (vec_ops.cuh) ------------------------------------------------
template <class R> struct Vec4T;
template <> struct Vec4T<float> { typedef float4 T; static constexpr int N = 4; };
template <> struct Vec4T<double> { typedef double2 T; static constexpr int N = 2; };
__device__ __forceinline__ void v4_avf(float4& p, const float4& r, float b) { p.x *= b; p.x += r.x; p.y *= b; p.y += r.y; p.z *= b; p.z += r.z; p.w *= b; p.w += r.w; }
__device__ __forceinline__ void v4_avf(double2& p, const double2& r, double b) { p.x *= b; p.x += r.x; p.y *= b; p.y += r.y; }
// x += p*alpha ; r += q*(-alpha)  (cgstep_alpha -> two vOp_v_inc_bf), then the term of rho' = r.r
template <class R> __device__ __forceinline__ double xr_one(R& x, R& r, R p, R q, R alpha, R malpha, bool a_one, bool ma_one) {
    if (a_one) x += p; else x += p * alpha;      // vOp takes `r += b` when k == 1
    if (ma_one) r += q; else r += q * malpha;
    return double(r) * double(r);
}
template <class R> __device__ __forceinline__ double r_one(R& r, R q, R malpha, bool ma_one) {
    if (ma_one) r += q; else r += q * malpha;
    return double(r) * double(r);
}
template <class R> __device__ __forceinline__ void x_one(R& x, R p, R alpha, bool a_one) {
    if (a_one) x += p; else x += p * alpha;      // vOp takes `r += b` when k == 1
}
__device__ __forceinline__ double xr_vec(float4& x, float4& r, const float4& p, const float4& q, float alpha, float malpha, bool a_one, bool ma_one) {
    double s = xr_one<float>(x.x, r.x, p.x, q.x, alpha, malpha, a_one, ma_one);
    s += xr_one<float>(x.y, r.y, p.y, q.y, alpha, malpha, a_one, ma_one);
    s += xr_one<float>(x.z, r.z, p.z, q.z, alpha, malpha, a_one, ma_one);
    s += xr_one<float>(x.w, r.w, p.w, q.w, alpha, malpha, a_one, ma_one);
    return s;
}
__device__ __forceinline__ double xr_vec(double2& x, double2& r, const double2& p, const double2& q, double alpha, double malpha, bool a_one, bool ma_one) {
    double s = xr_one<double>(x.x, r.x, p.x, q.x, alpha, malpha, a_one, ma_one);
    s += xr_one<double>(x.y, r.y, p.y, q.y, alpha, malpha, a_one, ma_one);
    return s;
}
// every thread of the CTA gets the sum of partials[0..n), added in a fixed order
template <class R> __device__ __forceinline__ double sum_partials_all(const double* partials, int n, double* red, double* bcast) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += __ldcg(partials + i);
    __syncthreads();
    s = block_sum(s, red);
    if (threadIdx.x == 0) *bcast = s;
    __syncthreads();
    return *bcast;
}

constexpr int kPersistNotEligible = 1;
// ---- persistent CG --------------------------------------------------------------------------------------------------
// Node ownership inside the kernel: a tile's interior nodes belong to the tile's CTA, a shared node belongs to the one thread
// that sums its staged contributions.  The owner alone reads and writes x, r and q of its nodes, so for interior nodes p and q
// never leave shared memory, and the x / r values an owner needs after a grid sync can be requested before it.  Only the p
// and r of SHARED nodes travel through HBM/L2 (p double-buffered), for the tiles that touch them.
// Shared-memory plan (bytes from the start of dynamic shared memory), computed by persist_layout().
struct PersistLayout {
    int tiles_cached;       // tiles per CTA (1 or 2): their node tables and staged vectors stay in shared memory
    int max_touched, max_slots, max_int, max_shtouch, maxval;
    unsigned off_slot, off_tidx, off_nrec, off_qr, off_jds, off_extra, total;   // off_extra: element-type scratch (hexahedra: cached stiffness matrices)
};
// static per-node data of a tile's interior nodes, kept in shared memory for the whole solve
template <class R> struct NodeRec { uint32_t g; uint32_t val_fixed; R mass; };        // val | fixed<<16
// ... and of the shared node a thread owns
template <class R> struct GRec { uint32_t g; uint32_t val_fixed; uint32_t base; R mass; };
template <class R> inline PersistLayout persist_layout(int tiles_per_cta, int max_touched, int max_slots, int max_int, int max_shtouch, int maxval, size_t extra_bytes = 0) {
    PersistLayout L;
    L.tiles_cached = tiles_per_cta; L.max_touched = max_touched; L.max_slots = max_slots; L.max_int = max_int; L.max_shtouch = max_shtouch; L.maxval = maxval;
    auto up = [](size_t o) { return (o + 15) & ~size_t(15); };
    size_t o = up(sizeof(typename SVec<R>::T) * size_t(max_touched) * tiles_per_cta);
    L.off_slot = unsigned(o); o = up(o + sizeof(R) * 3 * size_t(max_slots));
    L.off_tidx = unsigned(o); o = up(o + sizeof(uint32_t) * size_t(max_shtouch) * tiles_per_cta);
    L.off_nrec = unsigned(o); o = up(o + sizeof(NodeRec<R>) * size_t(max_int) * tiles_per_cta);
    L.off_qr = unsigned(o); o = up(o + sizeof(R) * 3 * size_t(max_int) * tiles_per_cta);
    L.off_jds = unsigned(o); o = up(o + sizeof(uint16_t) * size_t(maxval + 1) * tiles_per_cta);
    L.off_extra = unsigned(o); o = up(o + extra_bytes);
    L.total = unsigned(o);
    return L;
}
// static shared memory of the kernel for a CTA of `threads` (the per-thread shared-node records + reduction scratch)
template <class R> inline size_t persist_static_smem(int threads) { return sizeof(GRec<R>) * size_t(threads) + 33 * sizeof(double); }

template <class R> struct InboxWords;
template <> struct InboxWords<float> { static constexpr int N = 3; };
template <> struct InboxWords<double> { static constexpr int N = 6; };
// ---- multi-GPU: peer memory over NVLink (CUDA IPC mappings of every rank's mailbox), no NCCL inside the loop ---------------
constexpr int kMaxPeers = 8;
// Everything that crosses NVLink is written as 8-byte words carrying 32 bits of payload and a 32-bit sequence number (the scheme
// of NCCL's LL protocol): an aligned 8-byte store is single-copy atomic, so a reader that sees the sequence number it waits for
// has the payload too -- no system-scope fence on either side, no separate "data is ready" flag.
struct ARSlot { unsigned long long w[2]; };          // a double: {lo32 | seq<<32, hi32 | seq<<32}
template <class R> struct PeerDev {
    int enabled, rank, world, n_nb, max_sh;
    int nb_rank[kMaxPeers];
    ARSlot* ar;                                // local: [2][kMaxPeers] all-reduce slots (double-buffered by sequence parity)
    ARSlot* peer_ar[kMaxPeers];                // rank r's ar array (peer memory, own included)
    unsigned long long* ar4;                   // local: [2][kMaxPeers][8 words] all-reduce slots of the fused kernel (four doubles per rank)
    unsigned long long* peer_ar4[kMaxPeers];   // rank r's ar4 array
    unsigned long long* epoch;                 // local: launches done so far (sequence numbers are never reset)
    unsigned long long* inbox;                 // local: partial q of interface nodes received from the neighbours, kInboxWords<R> words per row
    unsigned long long* nb_inbox[kMaxPeers];   // neighbour k's inbox (peer memory)
    const int32_t* sh_if_row;                  // aligned with sh_nodes: row of the node in the interface table, or -1
    const int32_t* src;                        // [n_if][max_sh] in ascending-rank order: -1 own partial, -2 nobody, else inbox row
    const int2* if_send;                       // [n_if][max_sh-1]: {neighbour index or -1, row in that neighbour's inbox}
    const unsigned char* owned;                // [n_nodes] 1 = this rank counts the node in dot products
};

template <class R> struct PersistCG {
    NodeEpilogue<R> ep;     // epilogue of q = A p: mass / projection terms, dot_kind = DOT_STORE (out is not used)
    R* x; R* r;
    const R* b;             // right-hand side: |b| and the first rho = r.r are computed by the kernel itself (r == b unless warm start)
    typename SVec<R>::T* xt; typename SVec<R>::T* rt;   // x and r of the interior nodes in tile order (private to the owner CTA: one coalesced access per node)
    R* gstate;              // [9][gridDim.x * blockDim.x] p, r, x of each thread's shared node (private, coalesced), between iterations
    typename SVec<R>::T* p0; typename SVec<R>::T* p1;   // p of the shared nodes by node id, padded (one 16-byte access in Vec3f), double-buffered:
                            // tiles read p_old while the owners write p_new
    typename SVec<R>::T* rs; // r of the shared nodes by node id, padded (the flat r is only read in the first iteration)
    size_t n3;
    CGDev* cg;
    unsigned long long* sync;   // [3 * gridDim.x + 1] grid_sync_sum slots and arrival counter, zero at launch
    PersistLayout lay;
    PeerDev<R> peer;
};

// ---- grid-wide barrier that also sums one double per CTA --------------------------------------------------------------
// Arrival counter + one polling thread per CTA (all-to-all polling of per-CTA flags was measured 4x slower: 148 x 148 spinning
// loads on nine cache lines).  The per-CTA values go through three rotating arrays so that a slot is never rewritten while a
// slower CTA may still read it.  All CTAs add the same values in the same order.
// Writes made by any thread of any CTA before the sync are visible to every thread after it.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// The barrier in two halves, so that work which nobody else waits for can run between them:
//   grid_arrive  -- everything the CTA wrote so far is released (and, with a value, the CTA's term of the sum is posted)
//   grid_wait    -- returns once every CTA has arrived; what they wrote before arriving is visible to all threads
// (`who`: the thread that releases -- its warp stalls in the fence until the CTA's stores have drained, so a caller with more
// work ahead picks a warp that has the least of it)
__device__ __forceinline__ void grid_arrive(unsigned long long* slots, unsigned s, unsigned who = 0u) {
    const unsigned G = gridDim.x;
    unsigned* counter = reinterpret_cast<unsigned*>(slots + size_t(3) * G);
    (void)s;
    __syncthreads();                       // the CTA's writes precede the release
    if (threadIdx.x == who) { __threadfence(); atomicAdd(counter, 1u); }
}
__device__ __forceinline__ void grid_arrive_value(unsigned long long* slots, unsigned s, double cta_value /* thread 0 */) {
    const unsigned G = gridDim.x;
    double* cur = reinterpret_cast<double*>(slots) + size_t(s % 3) * G;
    unsigned* counter = reinterpret_cast<unsigned*>(slots + size_t(3) * G);
    __syncthreads();
    if (threadIdx.x == 0) { cur[blockIdx.x] = cta_value; __threadfence(); atomicAdd(counter, 1u); }
}
__device__ __forceinline__ void grid_wait(unsigned long long* slots, unsigned& s) {
    const unsigned G = gridDim.x;
    const unsigned* counter = reinterpret_cast<const unsigned*>(slots + size_t(3) * G);
    if (threadIdx.x == 0) {
        const unsigned target = (s + 1) * G;
        while (ld_acquire_u32(counter) < target) { }
    }
    __syncthreads();                       // thread 0's acquire precedes every thread's later reads
    ++s;
}
__device__ __forceinline__ double grid_wait_sum(unsigned long long* slots, unsigned& s, double* bcast) {
    const unsigned G = gridDim.x;
    const double* cur = reinterpret_cast<const double*>(slots) + size_t(s % 3) * G;
    grid_wait(slots, s);
    // warp 0 adds the G values in a fixed order and broadcasts
    if (threadIdx.x < 32) {
        double v = 0.0;
        for (unsigned i = threadIdx.x; i < G; i += 32) v += __ldcg(cur + i);
        __syncwarp();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) *bcast = v;
    }
    __syncthreads();
    return *bcast;
}
// arrive + wait, no value
__device__ __forceinline__ void grid_sync(unsigned long long* slots, unsigned& s) {
    grid_arrive(slots, s);
    grid_wait(slots, s);
}
__device__ __forceinline__ double grid_sync_sum(unsigned long long* slots, unsigned& s, double cta_value /* thread 0 */, double* red, double* bcast) {
    (void)red;
    grid_arrive_value(slots, s, cta_value);
    return grid_wait_sum(slots, s, bcast);
}

// ---- grid sync that also crosses the GPUs (multi-GPU mode) ------------------------------------------------------------------
// Local arrival as above; CTA 0 is the leader: once every CTA of this GPU has arrived it adds their values, talks to the other
// GPUs through peer memory (kind 1: tells the neighbours "my halo rows are in your inbox" and waits for theirs; kind 2: all-reduce
// of one double, every rank adds the ranks' values in rank order), publishes the result and releases the local CTAs.
// Every wait gives up after ~4 s (a rank that never arrives must not hang the GPUs): the solve is then marked as failed.
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
constexpr unsigned long long kSyncTimeoutNs = 4000000000ull;
// Time-outs of the polling loops are taken on the SM's cycle counter: a read of %globaltimer inside a polling loop costs microseconds
// per trip (measured: a grid sync with the timer in its loop took 7 us after the last arrival, 2 us without).
constexpr long long kSyncTimeoutCycles = 8000000000ll;          // ~4 s at the 2 GHz boost clock
__device__ __forceinline__ long long poll_clock() { return clock64(); }
struct DistSeq { unsigned long long base; unsigned halo, ar; };
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
// payload words of the interface rows: one per float, two per double
__device__ __forceinline__ void inbox_put(unsigned long long* row, int c, float v, unsigned seq) { st_relaxed_sys_u64(row + c, (unsigned long long)__float_as_uint(v) | ((unsigned long long)seq << 32)); }
__device__ __forceinline__ void inbox_put(unsigned long long* row, int c, double v, unsigned seq) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    st_relaxed_sys_u64(row + 2 * c, (b & 0xFFFFFFFFull) | ((unsigned long long)seq << 32));
    st_relaxed_sys_u64(row + 2 * c + 1, (b >> 32) | ((unsigned long long)seq << 32));
}
__device__ __forceinline__ bool inbox_get(const unsigned long long* row, int c, unsigned seq, float& v) {
    const unsigned long long w = ld_relaxed_sys_u64(row + c);
    if (unsigned(w >> 32) != seq) return false;
    v = __uint_as_float(unsigned(w)); return true;
}
__device__ __forceinline__ bool inbox_get(const unsigned long long* row, int c, unsigned seq, double& v) {
    const unsigned long long w0 = ld_relaxed_sys_u64(row + 2 * c), w1 = ld_relaxed_sys_u64(row + 2 * c + 1);
    if (unsigned(w0 >> 32) != seq || unsigned(w1 >> 32) != seq) return false;
    v = __longlong_as_double((long long)((w0 & 0xFFFFFFFFull) | (w1 << 32))); return true;
}
// All-reduce of one double across the GPUs, fused with the grid sync: CTA 0 adds the CTAs' values once all have arrived and
// stores the sum into every rank's slot (own slot first, with release: it also publishes this GPU's CTAs' writes); every
// CTA then polls the `world` slots of its own GPU and adds them in rank order.  Waits give up after ~4 s.
template <class R> __device__ __forceinline__ double dist_sync(const PersistCG<R>& a, unsigned& s, DistSeq& xs, double cta_value, double* bcast, bool& failed) {
    const unsigned G = gridDim.x;
    unsigned long long* slots = a.sync;
    double* cur = reinterpret_cast<double*>(slots) + size_t(s % 3) * G;
    unsigned* counter = reinterpret_cast<unsigned*>(slots + size_t(3) * G);
    const PeerDev<R>& P = a.peer;
    const unsigned long long aseq64 = xs.base + xs.ar + 1;
    const unsigned aseq = unsigned(aseq64);
    const int set = int(aseq64 & 1ull);
    __syncthreads();
    if (threadIdx.x == 0) { cur[blockIdx.x] = cta_value; __threadfence(); atomicAdd(counter, 1u); }
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            const long long t0 = poll_clock();
            while (ld_acquire_u32(counter) < (s + 1) * G) { if (poll_clock() - t0 > kSyncTimeoutCycles) break; }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            double v = 0.0;
            for (unsigned i = threadIdx.x; i < G; i += 32) v += __ldcg(cur + i);
            __syncwarp();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (threadIdx.x == 0) {
                const unsigned long long b = (unsigned long long)__double_as_longlong(v);
                const unsigned long long w0 = (b & 0xFFFFFFFFull) | ((unsigned long long)aseq << 32), w1 = (b >> 32) | ((unsigned long long)aseq << 32);
                ARSlot* mine = P.ar + set * kMaxPeers + P.rank;
                mine->w[0] = w0;
                st_release_gpu_u64(&mine->w[1], w1);
                for (int r = 0; r < P.world; ++r)
                    if (r != P.rank) { ARSlot* dst = P.peer_ar[r] + set * kMaxPeers + P.rank; st_relaxed_sys_u64(&dst->w[0], w0); st_relaxed_sys_u64(&dst->w[1], w1); }
            }
        }
    }
    if (threadIdx.x == 0) {
        const long long t0 = poll_clock();
        double tot = 0.0;
        bool fail = false;
        for (int r = 0; r < P.world && !fail; ++r) {
            const ARSlot* src = P.ar + set * kMaxPeers + r;
            unsigned long long w0, w1;
            for (;;) {
                w0 = ld_relaxed_sys_u64(&src->w[0]); w1 = ld_relaxed_sys_u64(&src->w[1]);
                if (unsigned(w0 >> 32) == aseq && unsigned(w1 >> 32) == aseq) break;
                if (poll_clock() - t0 > kSyncTimeoutCycles) { fail = true; break; }
            }
            tot += __longlong_as_double((long long)((w0 & 0xFFFFFFFFull) | (w1 << 32)));      // rank order
        }
        __threadfence();
        *bcast = fail ? __longlong_as_double(0x7FF8000000000001ll) : tot;
    }
    __syncthreads();
    ++s; ++xs.ar;
    const double r = *bcast;
    if (r != r && (__double_as_longlong(r) & 0xFFFFFFFFll) == 1ll) failed = true;
    return r;
}

template <class R> struct PersistState {
    typename SVec<R>::T* pold; typename SVec<R>::T* pnew;
    double rho, normb, tol, thr;
    int it;
    unsigned tsc, max_iter;
    bool first;
    R beta;
    unsigned sync_count;
    DistSeq xs;
    bool failed;
    bool updated;           // x has been updated at least once (persist_finish must write it back)
    __device__ explicit PersistState(const PersistCG<R>& a) {
        pold = a.p0; pnew = a.p1;
        const CGDev* cg = a.cg;
        rho = cg->rho; normb = cg->normb; tol = cg->tolerance; thr = cg->threshold;
        it = cg->it; tsc = cg->time_step_count; max_iter = cg->max_iter;
        first = true; beta = R(0); sync_count = 0; failed = false; updated = false;
        xs.base = a.peer.enabled ? (*a.peer.epoch) * 65536ull : 0ull; xs.halo = 0; xs.ar = 0;
    }
};

// L2-coherent load of a padded nodal vector (written by another CTA earlier in the same kernel)
__device__ __forceinline__ float4 sv_ldcg(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ SVec<double>::T sv_ldcg(const SVec<double>::T* p) {
    SVec<double>::T v; const double* d = reinterpret_cast<const double*>(p); v.x = __ldcg(d); v.y = __ldcg(d + 1); v.z = __ldcg(d + 2); return v;
}
// p = p*beta + r  (cgstep_beta -> vOp_avf, CGLinearSolver.inl:184-197), one component
template <class R> __device__ __forceinline__ R p_update(R p, R beta, R r) { p *= beta; p += r; return p; }

// Chunks of shared nodes are dealt to the CTAs round-robin (group g of CTA b sums chunk g * gridDim.x + b), so that every CTA gets
// n_chunks / gridDim.x of them, give or take one: with consecutive chunks per CTA the last CTAs of the grid would have none.
__device__ __forceinline__ int persist_chunk_of_thread() { return int(threadIdx.x / kGatherChunk) * int(gridDim.x) + int(blockIdx.x); }
// once per solve: the node tables of the CTA's tiles and of the thread's shared node go to shared memory
template <class R> __device__ __forceinline__ void persist_load_tables(const TileDev<R>& t, const PersistCG<R>& a, unsigned char* smem_raw, GRec<R>* s_grec) {
    const PersistLayout& L = a.lay;
    uint32_t* s_tidx = reinterpret_cast<uint32_t*>(smem_raw + L.off_tidx);
    NodeRec<R>* s_nrec = reinterpret_cast<NodeRec<R>*>(smem_raw + L.off_nrec);
    uint16_t* s_jds = reinterpret_cast<uint16_t*>(smem_raw + L.off_jds);
    const NodeEpilogue<R>& ep = a.ep;
    for (int c = 0; c < L.tiles_cached; ++c) {
        const int tile = blockIdx.x + c * gridDim.x;
        if (tile >= t.n_tiles) break;
        const uint32_t node_off = t.tile_node_off[tile];
        const int n_touched = int(t.tile_node_off[tile + 1] - node_off), n_int = int(t.tile_nint[tile]);
        for (int k = threadIdx.x; k < n_touched; k += blockDim.x) {
            const uint32_t g = t.tile_nodes[node_off + k];
            if (k < n_int) s_nrec[c * L.max_int + k] = NodeRec<R>{g, unsigned(t.tile_val[node_off + k]) | ((ep.fixed && ep.fixed[g]) ? 0x10000u : 0u), ep.mass ? ep.mass[g] : R(0)};
            else s_tidx[c * L.max_shtouch + (k - n_int)] = g;
        }
        for (int j = threadIdx.x; j <= t.maxval; j += blockDim.x) s_jds[c * (L.maxval + 1) + j] = t.tile_jds[size_t(tile) * (t.maxval + 1) + j];
    }
    // the thread's shared node (one round: checked by the host)
    const int chunk = persist_chunk_of_thread(), k = threadIdx.x % kGatherChunk;
    GRec<R> rec{0xFFFFFFFFu, 0u, 0u, R(0)};
    if (chunk < t.n_chunks) {
        const uint32_t g = t.sh_nodes[size_t(chunk) * kGatherChunk + k];
        if (g != 0xFFFFFFFFu)
            rec = GRec<R>{g, unsigned(t.sh_val[size_t(chunk) * kGatherChunk + k]) | ((ep.fixed && ep.fixed[g]) ? 0x10000u : 0u), t.sh_base[chunk] + k, ep.mass ? ep.mass[g] : R(0)};
    }
    s_grec[threadIdx.x] = rec;
    __syncthreads();
}

// [A0] the new search direction of every node touched by the CTA's tiles, in shared memory.  Interior nodes: from the p and
// the r the CTA itself left in shared memory (from HBM on the first iteration); shared nodes: from their owners' p_old and r.
template <class R> __device__ __forceinline__ void persist_phase1(const TileDev<R>& t, const PersistCG<R>& a, const PersistState<R>& st, unsigned char* smem_raw) {
    typedef typename SVec<R>::T SV;
    const PersistLayout& L = a.lay;
    SV* s_in = reinterpret_cast<SV*>(smem_raw);
    const uint32_t* s_tidx = reinterpret_cast<const uint32_t*>(smem_raw + L.off_tidx);
    const NodeRec<R>* s_nrec = reinterpret_cast<const NodeRec<R>*>(smem_raw + L.off_nrec);
    const R* s_qr = reinterpret_cast<const R*>(smem_raw + L.off_qr);
    // one node: interior (k < n_int) from shared memory, shared from the owners' p_old and r (rv, po: requested by the caller)
    auto one = [&](int c, int k, int n_int, const SV& rv, const SV& po_sh) {
        R p0, p1, p2;
        if (k < n_int) {
            const SV po = s_in[c * L.max_touched + k];
            const R* rn = s_qr + 3 * size_t(c * L.max_int + k);
            p0 = p_update<R>(R(po.x), st.beta, rn[0]); p1 = p_update<R>(R(po.y), st.beta, rn[1]); p2 = p_update<R>(R(po.z), st.beta, rn[2]);
        } else {
            p0 = p_update<R>(R(po_sh.x), st.beta, R(rv.x)); p1 = p_update<R>(R(po_sh.y), st.beta, R(rv.y)); p2 = p_update<R>(R(po_sh.z), st.beta, R(rv.z));
        }
        s_in[c * L.max_touched + k] = SVec<R>::make(p0, p1, p2);
    };
    if (st.first) {
        for (int c = 0; c < L.tiles_cached; ++c) {
            const int tile = blockIdx.x + c * gridDim.x;
            if (tile >= t.n_tiles) break;
            const int n_touched = int(t.tile_node_off[tile + 1] - t.tile_node_off[tile]), n_int = int(t.tile_nint[tile]);
            for (int k = threadIdx.x; k < n_touched; k += blockDim.x) {
                const size_t g = k < n_int ? size_t(s_nrec[c * L.max_int + k].g) : size_t(s_tidx[c * L.max_shtouch + (k - n_int)]);
                s_in[c * L.max_touched + k] = SVec<R>::make(__ldcg(a.r + 3 * g), __ldcg(a.r + 3 * g + 1), __ldcg(a.r + 3 * g + 2));
            }
        }
    } else {
        // The p_old / r of the shared nodes come out of L2: all the requests of a thread (two tiles x two rounds) are issued
        // before the first is used, so the phase pays one L2 round trip instead of four.
        constexpr int kTiles = 2, kRounds = 2;
        SV rv[kTiles][kRounds], po[kTiles][kRounds];
        int n_touched[kTiles], n_int[kTiles];
#pragma unroll
        for (int c = 0; c < kTiles; ++c)
#pragma unroll
            for (int u = 0; u < kRounds; ++u) { rv[c][u] = SVec<R>::make(R(0), R(0), R(0)); po[c][u] = rv[c][u]; }
#pragma unroll
        for (int c = 0; c < kTiles; ++c) {
            const int tile = blockIdx.x + c * gridDim.x;
            const bool on = c < L.tiles_cached && tile < t.n_tiles;
            n_touched[c] = on ? int(t.tile_node_off[tile + 1] - t.tile_node_off[tile]) : 0;
            n_int[c] = on ? int(t.tile_nint[tile]) : 0;
        }
#pragma unroll
        for (int c = 0; c < kTiles; ++c)
#pragma unroll
            for (int u = 0; u < kRounds; ++u) {
                const int k = threadIdx.x + u * blockDim.x;
                if (k >= n_int[c] && k < n_touched[c]) {
                    const size_t g = s_tidx[c * L.max_shtouch + (k - n_int[c])];
                    rv[c][u] = sv_ldcg(a.rs + g); po[c][u] = sv_ldcg(st.pold + g);
                }
            }
#pragma unroll
        for (int c = 0; c < kTiles; ++c)
#pragma unroll
            for (int u = 0; u < kRounds; ++u) {
                const int k = threadIdx.x + u * blockDim.x;
                if (k < n_touched[c]) one(c, k, n_int[c], rv[c][u], po[c][u]);
            }
        for (int c = 0; c < kTiles; ++c)
            for (int k = threadIdx.x + kRounds * blockDim.x; k < n_touched[c]; k += blockDim.x) {
                SV r1 = SVec<R>::make(R(0), R(0), R(0)), p1 = r1;
                if (k >= n_int[c]) { const size_t g = s_tidx[c * L.max_shtouch + (k - n_int[c])]; r1 = sv_ldcg(a.rs + g); p1 = sv_ldcg(st.pold + g); }
                one(c, k, n_int[c], r1, p1);
            }
    }
    __syncthreads();
}

// phase 3 of a tile inside the CG loop: like tile_phase3, with the node's valence / mass / fixed flag from shared memory, and
// q left in shared memory for the x / r update
template <class R> __device__ __forceinline__ double persist_phase3(const TileDev<R>& t, int tile, int c, const PersistCG<R>& a, unsigned char* smem_raw) {
    typedef typename SVec<R>::T SV;
    const PersistLayout& L = a.lay;
    const NodeEpilogue<R>& ep = a.ep;
    const SV* s_in = reinterpret_cast<const SV*>(smem_raw) + c * L.max_touched;
    const R* s_slot = reinterpret_cast<const R*>(smem_raw + L.off_slot);
    const NodeRec<R>* s_nrec = reinterpret_cast<const NodeRec<R>*>(smem_raw + L.off_nrec) + c * L.max_int;
    R* s_qr = reinterpret_cast<R*>(smem_raw + L.off_qr) + 3 * size_t(c * L.max_int);
    const uint16_t* s_jds = reinterpret_cast<const uint16_t*>(smem_raw + L.off_jds) + c * (L.maxval + 1);
    const int max_slots = L.max_slots;
    const int n_int = int(t.tile_nint[tile]);
    const bool plus = ep.sign > 0;
    double part = 0.0;
    for (int k = threadIdx.x; k < n_int; k += blockDim.x) {
        const NodeRec<R> rec = s_nrec[k];
        const int val = int(rec.val_fixed & 0xFFFFu);
        const SV pv = s_in[k];
        R ax = R(0), ay = R(0), az = R(0);
        node_mass_m(ep, ep.pre_kind, rec.mass, R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
        int jj = 0;
        for (; jj + 4 <= val; jj += 4) {
            R cx[4], cy[4], cz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int s = s_jds[jj + u] + k; cx[u] = s_slot[3 * s]; cy[u] = s_slot[3 * s + 1]; cz[u] = s_slot[3 * s + 2]; }
            if (plus) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { ax += cx[u]; ay += cy[u]; az += cz[u]; }
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) { ax -= cx[u]; ay -= cy[u]; az -= cz[u]; }
            }
        }
        for (; jj < val; ++jj) {
            const int s = s_jds[jj] + k;
            if (plus) { ax += s_slot[3 * s]; ay += s_slot[3 * s + 1]; az += s_slot[3 * s + 2]; }
            else { ax -= s_slot[3 * s]; ay -= s_slot[3 * s + 1]; az -= s_slot[3 * s + 2]; }
        }
        part += node_finish_m(ep, rec.g, rec.mass, (rec.val_fixed & 0x10000u) != 0, R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
        s_qr[3 * k] = ax; s_qr[3 * k + 1] = ay; s_qr[3 * k + 2] = az;
    }
    return part;
}

// Everything of an iteration after the tiles: returns false when the solve is over.  `part`: this thread's share of
// p.q over the interior nodes of the CTA's tiles.
// `arrived`: the CTA has already arrived at the staging barrier (during its last tile); only the wait is left.
template <class R> __device__ __forceinline__ bool persist_rest(const TileDev<R>& t, const PersistCG<R>& a, PersistState<R>& st, double part, double* red, double* bcast,
                                                                unsigned char* smem_raw, const GRec<R>* s_grec, bool arrived = false) {
    typedef typename SVec<R>::T SV;
    CGDev* cg = a.cg;
    const NodeEpilogue<R>& ep = a.ep;
    const PersistLayout& L = a.lay;
    // ---- [B] the thread's shared node.  Its p and r live in registers; x is requested before the barrier.
    const GRec<R> grec = s_grec[threadIdx.x];
    const bool has_node = grec.g != 0xFFFFFFFFu;
    // p, r, x of the node are private to this thread; between iterations they rest in a coalesced scratch array (registers
    // held across the element pass would spill there)
    const size_t gs_n = size_t(gridDim.x) * blockDim.x, gs_i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const PeerDev<R>& P = a.peer;
    // multi-GPU: is the node on the partition interface (row of the halo tables), and does this rank count it in dot products
    const int if_row = (P.enabled && has_node) ? P.sh_if_row[size_t(persist_chunk_of_thread()) * kGatherChunk + threadIdx.x % kGatherChunk] : -1;
    const bool counted = if_row < 0 || P.owned[grec.g] != 0;
    R gp0 = R(0), gp1 = R(0), gp2 = R(0), gr0 = R(0), gr1 = R(0), gr2 = R(0);
    if (has_node) {
        const size_t g3 = 3 * size_t(grec.g);
        if (st.first) { gr0 = a.r[g3]; gr1 = a.r[g3 + 1]; gr2 = a.r[g3 + 2]; gp0 = gr0; gp1 = gr1; gp2 = gr2; }
        else {
            gr0 = a.gstate[3 * gs_n + gs_i]; gr1 = a.gstate[4 * gs_n + gs_i]; gr2 = a.gstate[5 * gs_n + gs_i];
            gp0 = p_update<R>(a.gstate[gs_i], st.beta, gr0); gp1 = p_update<R>(a.gstate[gs_n + gs_i], st.beta, gr1); gp2 = p_update<R>(a.gstate[2 * gs_n + gs_i], st.beta, gr2);
        }
        st.pnew[grec.g] = SVec<R>::make(gp0, gp1, gp2);               // for the tiles that touch the node, next iteration
    }
    trace_mark(ep.trace, kTraceTail, 1);
    if (!arrived) grid_arrive(a.sync, st.sync_count);
    grid_wait(a.sync, st.sync_count);                  // staged contributions are complete
    trace_mark(ep.trace, kTraceTail, 2);
    double part2 = part;                               // p.q of the interior nodes + (below) of the thread's shared node
    R gq0 = R(0), gq1 = R(0), gq2 = R(0);
    if (has_node) {
        const int val = int(grec.val_fixed & 0xFFFFu);
        const Quad<R>* stg = t.stage + grec.base;
        const uint64_t pol = l2_policy_evict_first();
        Quad<R> b0[kGatherBatch], b1[kGatherBatch];
        gather_load<R>(b0, stg, 0, val, pol);
        gather_load<R>(b1, stg, kGatherBatch, val, pol);
        node_mass_m(ep, ep.pre_kind, grec.mass, gp0, gp1, gp2, gq0, gq1, gq2);
        gather_sum<R>(b0, b1, stg, val, ep.sign > 0, gq0, gq1, gq2, pol);
        // p.q of the node.  Multi-GPU: (gq) of an interface node is this rank's PARTIAL sum; since p is the same on every sharing
        // rank, the ranks' p.q_partial add up to p.q, so den needs no exchanged q: halo and all-reduce share one cross-GPU sync.
        part2 += node_finish_m(ep, grec.g, grec.mass, (grec.val_fixed & 0x10000u) != 0, gp0, gp1, gp2, gq0, gq1, gq2);
        if (if_row >= 0) {
            // every other sharing rank gets the partial sum in its inbox (NVLink store)
            for (int e = 0; e < P.max_sh - 1; ++e) {
                const int2 to = P.if_send[if_row * (P.max_sh - 1) + e];
                if (to.x >= 0) {
                    unsigned long long* d = P.nb_inbox[to.x] + size_t(InboxWords<R>::N) * size_t(to.y);
                    const unsigned hseq = unsigned(st.xs.base + st.xs.halo + 1);
                    inbox_put(d, 0, gq0, hseq); inbox_put(d, 1, gq1, hseq); inbox_put(d, 2, gq2, hseq);
                }
            }
        }
    }
    const SV* s_in = reinterpret_cast<const SV*>(smem_raw);
    const NodeRec<R>* s_nrec = reinterpret_cast<const NodeRec<R>*>(smem_raw + L.off_nrec);
    R* s_qr = reinterpret_cast<R*>(smem_raw + L.off_qr);
    __syncthreads();
    part2 = block_sum(part2, red);
    trace_mark(ep.trace, kTraceTail, 3);
    const double den = P.enabled ? dist_sync<R>(a, st.sync_count, st.xs, part2, bcast, st.failed) : grid_sync_sum(a.sync, st.sync_count, part2, red, bcast);
    bool halo_timeout = false, prr_poison = false;
    if (P.enabled && if_row >= 0 && !st.failed) {
        // q of an interface node: the sharing ranks' partial sums in ascending rank order -- same operands, same order, same bits on every rank
        R s0 = R(0), s1 = R(0), s2 = R(0);
        for (int j = 0; j < P.max_sh; ++j) {
            const int sj = P.src[if_row * P.max_sh + j];
            R c0 = R(0), c1 = R(0), c2 = R(0);
            if (sj == -1) { c0 = gq0; c1 = gq1; c2 = gq2; }
            else if (sj >= 0) {
                // (the words were sent before the neighbour's contribution to den, so they have normally landed by now)
                const unsigned long long* row = P.inbox + size_t(InboxWords<R>::N) * size_t(sj);
                const unsigned hseq = unsigned(st.xs.base + st.xs.halo + 1);
                const long long t0 = poll_clock();
                while (!(inbox_get(row, 0, hseq, c0) && inbox_get(row, 1, hseq, c1) && inbox_get(row, 2, hseq, c2)))
                    if (poll_clock() - t0 > kSyncTimeoutCycles) { halo_timeout = true; break; }
            }
            if (j == 0) { s0 = c0; s1 = c1; s2 = c2; } else { s0 += c0; s1 += c1; s2 += c2; }
        }
        gq0 = s0; gq1 = s1; gq2 = s2;
    }
    if (P.enabled) { ++st.xs.halo; if (__syncthreads_or(halo_timeout ? 1 : 0)) prr_poison = true; }
    if (st.failed) { if (blockIdx.x == 0 && threadIdx.x == 0) { cg->done = 1; cg->end_cond = 99; } return false; }
    trace_mark(ep.trace, kTraceTail, 4);
    bool stop = false;
    if (den != 0.0) { if (fabs(den) <= st.thr && !(st.it == 1 && st.tsc == 0)) stop = true; } else stop = true;
    if (blockIdx.x == 0 && threadIdx.x == 0) cg_after_den(cg, den);
    if (stop) return false;
    // ---- [C] x += alpha p ; r -= alpha q, by the owners
    const double alpha_d = st.rho / den;
    const R alpha = R(alpha_d), malpha = R(-alpha_d);
    const bool a_one = (alpha_d == 1.0), ma_one = (-alpha_d == 1.0);
    double prr = prr_poison ? __longlong_as_double(0x7FF8000000000001ll) : 0.0;    // a lost halo row fails the solve on every rank through the next all-reduce
    // The r half of the update comes first: rho = r.r is what the other CTAs wait for.  The x half (nobody reads x before the
    // solve ends) runs between the arrival at the rho sync and the wait, under the barrier's latency.
    if (has_node) {
        double own = r_one<R>(gr0, gq0, malpha, ma_one);
        own += r_one<R>(gr1, gq1, malpha, ma_one);
        own += r_one<R>(gr2, gq2, malpha, ma_one);
        if (counted) prr += own;
        a.rs[grec.g] = SVec<R>::make(gr0, gr1, gr2);                           // for the tiles that touch the node
    }
    constexpr int kTiles = 2, kRounds = 2;
    int n_int[kTiles]; uint32_t node_off[kTiles];
#pragma unroll
    for (int c = 0; c < kTiles; ++c) {
        const int tile = blockIdx.x + c * gridDim.x;
        const bool on = c < L.tiles_cached && tile < t.n_tiles;
        n_int[c] = on ? int(t.tile_nint[tile]) : 0;
        node_off[c] = on ? t.tile_node_off[tile] : 0u;
    }
    auto r_node = [&](int c, int k, R r0, R r1, R r2) {
        R* q = s_qr + 3 * size_t(c * L.max_int + k);
        prr += r_one<R>(r0, q[0], malpha, ma_one);
        prr += r_one<R>(r1, q[1], malpha, ma_one);
        prr += r_one<R>(r2, q[2], malpha, ma_one);
        a.rt[node_off[c] + k] = SVec<R>::make(r0, r1, r2);
        q[0] = r0; q[1] = r1; q[2] = r2;
    };
    auto x_node = [&](int c, int k, R x0, R x1, R x2) {
        const SV pv = s_in[c * L.max_touched + k];
        x_one<R>(x0, R(pv.x), alpha, a_one); x_one<R>(x1, R(pv.y), alpha, a_one); x_one<R>(x2, R(pv.z), alpha, a_one);
        a.xt[node_off[c] + k] = SVec<R>::make(x0, x1, x2);
    };
    if (st.first) {
        for (int c = 0; c < kTiles; ++c)
            for (int k = threadIdx.x; k < n_int[c]; k += blockDim.x) {
                const size_t g3 = 3 * size_t(s_nrec[c * L.max_int + k].g);
                r_node(c, k, a.r[g3], a.r[g3 + 1], a.r[g3 + 2]);
            }
    } else {
        SV rv[kTiles][kRounds];
#pragma unroll
        for (int c = 0; c < kTiles; ++c)
#pragma unroll
            for (int u = 0; u < kRounds; ++u) {
                const int k = threadIdx.x + u * blockDim.x;
                rv[c][u] = k < n_int[c] ? a.rt[node_off[c] + k] : SVec<R>::make(R(0), R(0), R(0));
            }
#pragma unroll
        for (int c = 0; c < kTiles; ++c)
#pragma unroll
            for (int u = 0; u < kRounds; ++u) {
                const int k = threadIdx.x + u * blockDim.x;
                if (k < n_int[c]) r_node(c, k, R(rv[c][u].x), R(rv[c][u].y), R(rv[c][u].z));
            }
        for (int c = 0; c < kTiles; ++c)
            for (int k = threadIdx.x + kRounds * blockDim.x; k < n_int[c]; k += blockDim.x) { const SV v = a.rt[node_off[c] + k]; r_node(c, k, R(v.x), R(v.y), R(v.z)); }
    }
    __syncthreads();
    prr = block_sum(prr, red);
    trace_mark(ep.trace, kTraceTail, 5);
    if (!P.enabled) grid_arrive_value(a.sync, st.sync_count, prr);
    // ---- x += alpha p (owners only; private data, needs no release)
    if (has_node) {
        const size_t g3 = 3 * size_t(grec.g);
        R gx0, gx1, gx2;
        if (st.first) { gx0 = a.x[g3]; gx1 = a.x[g3 + 1]; gx2 = a.x[g3 + 2]; }
        else { gx0 = a.gstate[6 * gs_n + gs_i]; gx1 = a.gstate[7 * gs_n + gs_i]; gx2 = a.gstate[8 * gs_n + gs_i]; }
        x_one<R>(gx0, gp0, alpha, a_one); x_one<R>(gx1, gp1, alpha, a_one); x_one<R>(gx2, gp2, alpha, a_one);
        a.gstate[gs_i] = gp0; a.gstate[gs_n + gs_i] = gp1; a.gstate[2 * gs_n + gs_i] = gp2;
        a.gstate[3 * gs_n + gs_i] = gr0; a.gstate[4 * gs_n + gs_i] = gr1; a.gstate[5 * gs_n + gs_i] = gr2;
        a.gstate[6 * gs_n + gs_i] = gx0; a.gstate[7 * gs_n + gs_i] = gx1; a.gstate[8 * gs_n + gs_i] = gx2;
    }
    if (st.first) {
        for (int c = 0; c < kTiles; ++c)
            for (int k = threadIdx.x; k < n_int[c]; k += blockDim.x) {
                const size_t g3 = 3 * size_t(s_nrec[c * L.max_int + k].g);
                x_node(c, k, a.x[g3], a.x[g3 + 1], a.x[g3 + 2]);
            }
    } else {
        SV xv[kTiles][kRounds];
#pragma unroll
        for (int c = 0; c < kTiles; ++c)
#pragma unroll
            for (int u = 0; u < kRounds; ++u) {
                const int k = threadIdx.x + u * blockDim.x;
                xv[c][u] = k < n_int[c] ? a.xt[node_off[c] + k] : SVec<R>::make(R(0), R(0), R(0));
            }
#pragma unroll
        for (int c = 0; c < kTiles; ++c)
#pragma unroll
            for (int u = 0; u < kRounds; ++u) {
                const int k = threadIdx.x + u * blockDim.x;
                if (k < n_int[c]) x_node(c, k, R(xv[c][u].x), R(xv[c][u].y), R(xv[c][u].z));
            }
        for (int c = 0; c < kTiles; ++c)
            for (int k = threadIdx.x + kRounds * blockDim.x; k < n_int[c]; k += blockDim.x) { const SV v = a.xt[node_off[c] + k]; x_node(c, k, R(v.x), R(v.y), R(v.z)); }
    }
    st.updated = true;
    const double rho_new = P.enabled ? dist_sync<R>(a, st.sync_count, st.xs, prr, bcast, st.failed) : grid_wait_sum(a.sync, st.sync_count, bcast);
    if (st.failed) { if (blockIdx.x == 0 && threadIdx.x == 0) { cg->done = 1; cg->end_cond = 99; } return false; }
    trace_mark(ep.trace, kTraceTail, 6);
    const int it2 = st.it + 1;
    bool stop2 = unsigned(it2) > st.max_iter;
    if (!stop2) { const double err = sqrt(rho_new) / st.normb; if (err <= st.tol && !(it2 == 1 && st.tsc == 0)) stop2 = true; }
    if (blockIdx.x == 0 && threadIdx.x == 0) cg_after_rho(cg, rho_new);
    if (stop2) return false;
    st.beta = R(rho_new / st.rho);
    st.rho = rho_new; st.it = it2; st.first = false;
    typename SVec<R>::T* tmp = st.pold; st.pold = st.pnew; st.pnew = tmp;
    return true;
}

// Start of the solve: normb = |b| and rho = r.r (CGLinearSolver.inl:130-180), all-reduced over the GPUs in multi-GPU mode (owned
// nodes only).  Returns false when the solve is already over (b == 0, or the initial residual meets the tolerance).
template <class R> __device__ __forceinline__ bool persist_init(const PersistCG<R>& a, PersistState<R>& st, double* red, double* bcast) {
    const PeerDev<R>& P = a.peer;
    CGDev* cg = a.cg;
    const size_t n = a.n3 / 3;
    double sb = 0.0, sr = 0.0;
    for (size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x; g < n; g += size_t(gridDim.x) * blockDim.x) {
        if (P.enabled && !P.owned[g]) continue;
        const R b0 = a.b[3 * g], b1 = a.b[3 * g + 1], b2 = a.b[3 * g + 2], r0 = a.r[3 * g], r1 = a.r[3 * g + 1], r2 = a.r[3 * g + 2];
        sb += double(b0) * double(b0) + double(b1) * double(b1) + double(b2) * double(b2);
        sr += double(r0) * double(r0) + double(r1) * double(r1) + double(r2) * double(r2);
    }
    __syncthreads();
    sb = block_sum(sb, red);
    const double nb2 = P.enabled ? dist_sync<R>(a, st.sync_count, st.xs, sb, bcast, st.failed) : grid_sync_sum(a.sync, st.sync_count, sb, red, bcast);
    __syncthreads();
    sr = block_sum(sr, red);
    const double rho0 = P.enabled ? dist_sync<R>(a, st.sync_count, st.xs, sr, bcast, st.failed) : grid_sync_sum(a.sync, st.sync_count, sr, red, bcast);
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    if (st.failed) { if (lead) { cg->done = 1; cg->end_cond = 99; } return false; }
    st.normb = sqrt(nb2);
    if (lead) cg->normb = st.normb;
    if (st.normb == 0.0) { if (lead) { cg->done = 1; cg->nb_iter = 0; cg->end_cond = 4; } return false; }
    if (lead) cg_after_rho(cg, rho0);                      // it: 0 -> 1, first entry of the error graph, tolerance test
    st.rho = rho0; st.it = 1;
    if (1u > st.max_iter) return false;
    const double err = sqrt(rho0) / st.normb;
    if (err <= st.tol && !(st.tsc == 0)) return false;
    return true;
}

// after the last iteration: x back in the caller's flat vector (during the solve the owners keep it to themselves)
template <class R> __device__ __forceinline__ void persist_finish(const TileDev<R>& t, const PersistCG<R>& a, const PersistState<R>& st, unsigned char* smem_raw, const GRec<R>* s_grec) {
    typedef typename SVec<R>::T SV;
    const PersistLayout& L = a.lay;
    if (a.peer.enabled && blockIdx.x == 0 && threadIdx.x == 0) *a.peer.epoch += 1ull;   // (every CTA read it at the start)
    if (!st.updated) return;                 // no update was made: x is untouched
    const GRec<R> grec = s_grec[threadIdx.x];
    if (grec.g != 0xFFFFFFFFu) {
        const size_t gs_n = size_t(gridDim.x) * blockDim.x, gs_i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
        R* d = a.x + 3 * size_t(grec.g); d[0] = a.gstate[6 * gs_n + gs_i]; d[1] = a.gstate[7 * gs_n + gs_i]; d[2] = a.gstate[8 * gs_n + gs_i];
    }
    const NodeRec<R>* s_nrec = reinterpret_cast<const NodeRec<R>*>(smem_raw + L.off_nrec);
    for (int c = 0; c < L.tiles_cached; ++c) {
        const int tile = blockIdx.x + c * gridDim.x;
        if (tile >= t.n_tiles) break;
        const int n_int = int(t.tile_nint[tile]);
        const uint32_t node_off = t.tile_node_off[tile];
        for (int k = threadIdx.x; k < n_int; k += blockDim.x) {
            const SV xv = a.xt[node_off + k];
            R* d = a.x + 3 * size_t(s_nrec[c * L.max_int + k].g);
            d[0] = R(xv.x); d[1] = R(xv.y); d[2] = R(xv.z);
        }
    }
}

}  // namespace sb
