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
// Shared-memory plan of the kernel (bytes from the start of dynamic shared memory), computed by persist_layout().
struct PersistLayout {
    int tiles_cached;       // tiles per CTA (1 or 2): their node tables and staged vectors stay in shared memory
    int max_touched, max_slots, max_int;
    unsigned off_slot, off_tidx, off_nrec, total;
};
template <class R> inline PersistLayout persist_layout(int tiles_per_cta, int max_touched, int max_slots, int max_int) {
    PersistLayout L;
    L.tiles_cached = tiles_per_cta; L.max_touched = max_touched; L.max_slots = max_slots; L.max_int = max_int;
    size_t o = sizeof(typename SVec<R>::T) * size_t(max_touched) * tiles_per_cta;
    o = (o + 15) & ~size_t(15); L.off_slot = unsigned(o);
    o += sizeof(R) * 3 * size_t(max_slots);
    o = (o + 15) & ~size_t(15); L.off_tidx = unsigned(o);
    o += sizeof(uint32_t) * size_t(max_touched) * tiles_per_cta;
    o = (o + 15) & ~size_t(15); L.off_nrec = unsigned(o);
    o += 16 * size_t(max_int) * tiles_per_cta;
    L.total = unsigned(o);
    return L;
}
// static per-node data of the nodes a thread finishes, kept in shared memory for the whole solve
struct alignas(16) NodeRec { uint32_t g; uint32_t val_fixed; uint32_t mass_lo, mass_hi; };   // val | fixed<<16 ; mass bits (float in lo, double in lo/hi)
template <class R> __device__ __forceinline__ NodeRec make_node_rec(uint32_t g, unsigned val, bool fx, R m);
template <> __device__ __forceinline__ NodeRec make_node_rec<float>(uint32_t g, unsigned val, bool fx, float m) { return NodeRec{g, val | (fx ? 0x10000u : 0u), __float_as_uint(m), 0u}; }
template <> __device__ __forceinline__ NodeRec make_node_rec<double>(uint32_t g, unsigned val, bool fx, double m) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(m);
    return NodeRec{g, val | (fx ? 0x10000u : 0u), unsigned(b), unsigned(b >> 32)};
}
template <class R> __device__ __forceinline__ R node_rec_mass(const NodeRec& n);
template <> __device__ __forceinline__ float node_rec_mass<float>(const NodeRec& n) { return __uint_as_float(n.mass_lo); }
template <> __device__ __forceinline__ double node_rec_mass<double>(const NodeRec& n) { return __longlong_as_double((long long)((unsigned long long)n.mass_lo | ((unsigned long long)n.mass_hi << 32))); }

template <class R> struct PersistCG {
    NodeEpilogue<R> ep;     // epilogue of q = A p: out = q, mass / projection terms, dot_kind = DOT_STORE
    R* x; R* r;
    R* p0; R* p1;           // the search direction is double-buffered: tiles read p_old of shared nodes while p_new is written
    size_t n3;
    CGDev* cg;
    unsigned long long* sync;   // [3 * gridDim.x + 1] grid_sync_sum slots and arrival counter, zero at launch
    PersistLayout lay;
    int debug;              // tuning experiments (SOFAB200_DEBUG_MODE)
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
__device__ __forceinline__ double grid_sync_sum(unsigned long long* slots, unsigned& s, double cta_value /* thread 0 */, double* red, double* bcast) {
    const unsigned G = gridDim.x;
    double* cur = reinterpret_cast<double*>(slots) + size_t(s % 3) * G;
    unsigned* counter = reinterpret_cast<unsigned*>(slots + size_t(3) * G);
    __syncthreads();                       // the CTA's writes precede thread 0's release
    if (threadIdx.x == 0) {
        cur[blockIdx.x] = cta_value;
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned target = (s + 1) * G;
        while (ld_acquire_u32(counter) < target) { }
        __threadfence();
    }
    __syncthreads();
    double v = 0.0;
    for (unsigned i = threadIdx.x; i < G; i += blockDim.x) v += __ldcg(cur + i);
    v = block_sum(v, red);
    if (threadIdx.x == 0) *bcast = v;
    __syncthreads();
    ++s;
    return *bcast;
}

template <class R> struct PersistState {
    R* pold; R* pnew;
    double rho, normb, tol, thr;
    int it;
    unsigned tsc, max_iter;
    bool first;
    R beta;
    unsigned sync_count;
    __device__ explicit PersistState(const PersistCG<R>& a) {
        pold = a.p0; pnew = a.p1;
        const CGDev* cg = a.cg;
        rho = cg->rho; normb = cg->normb; tol = cg->tolerance; thr = cg->threshold;
        it = cg->it; tsc = cg->time_step_count; max_iter = cg->max_iter;
        first = true; beta = R(0); sync_count = 0;
    }
};

// p_new = r (first iteration) or p_old*beta + r (cgstep_beta -> vOp_avf, CGLinearSolver.inl:184-197) for one node
template <class R> __device__ __forceinline__ void persist_p_node(const PersistState<R>& st, const R* r, size_t g, R& p0, R& p1, R& p2) {
    const R r0 = __ldcg(r + 3 * g), r1 = __ldcg(r + 3 * g + 1), r2 = __ldcg(r + 3 * g + 2);
    if (st.first) { p0 = r0; p1 = r1; p2 = r2; return; }
    p0 = __ldcg(st.pold + 3 * g); p1 = __ldcg(st.pold + 3 * g + 1); p2 = __ldcg(st.pold + 3 * g + 2);
    p0 *= st.beta; p0 += r0; p1 *= st.beta; p1 += r1; p2 *= st.beta; p2 += r2;
}

// once per solve: the node tables of the CTA's tiles and of the thread's shared node go to shared memory
template <class R> __device__ __forceinline__ void persist_load_tables(const TileDev<R>& t, const PersistCG<R>& a, unsigned char* smem_raw, NodeRec* s_grec, uint32_t* s_gbase) {
    const PersistLayout& L = a.lay;
    uint32_t* s_tidx = reinterpret_cast<uint32_t*>(smem_raw + L.off_tidx);
    NodeRec* s_nrec = reinterpret_cast<NodeRec*>(smem_raw + L.off_nrec);
    const NodeEpilogue<R>& ep = a.ep;
    for (int c = 0; c < L.tiles_cached; ++c) {
        const int tile = blockIdx.x + c * gridDim.x;
        if (tile >= t.n_tiles) break;
        const uint32_t node_off = t.tile_node_off[tile];
        const int n_touched = int(t.tile_node_off[tile + 1] - node_off), n_int = int(t.tile_nint[tile]);
        for (int k = threadIdx.x; k < n_touched; k += blockDim.x) {
            const uint32_t g = t.tile_nodes[node_off + k];
            s_tidx[c * L.max_touched + k] = g;
            if (k < n_int) s_nrec[c * L.max_int + k] = make_node_rec<R>(g, t.tile_val[node_off + k], ep.fixed && ep.fixed[g], ep.mass ? ep.mass[g] : R(0));
        }
    }
    // the thread's shared node: chunk = blockIdx.x * groups + threadIdx.x / kGatherChunk (one round: checked by the host)
    const int chunk = blockIdx.x * (blockDim.x / kGatherChunk) + threadIdx.x / kGatherChunk, k = threadIdx.x % kGatherChunk;
    NodeRec rec{0xFFFFFFFFu, 0u, 0u, 0u};
    uint32_t base = 0;
    if (chunk < t.n_chunks) {
        const uint32_t g = t.sh_nodes[size_t(chunk) * kGatherChunk + k];
        if (g != 0xFFFFFFFFu) {
            rec = make_node_rec<R>(g, t.sh_val[size_t(chunk) * kGatherChunk + k], ep.fixed && ep.fixed[g], ep.mass ? ep.mass[g] : R(0));
            base = t.sh_base[chunk] + k;
        }
    }
    s_grec[threadIdx.x] = rec; s_gbase[threadIdx.x] = base;
    __syncthreads();
}

// phase 1 of an iteration: the new search direction of every node touched by the CTA's tiles goes to shared memory (one
// round trip for all tiles); the tile that holds a node as interior also writes it to the p_new vector.
template <class R> __device__ __forceinline__ void persist_phase1(const TileDev<R>& t, const PersistCG<R>& a, const PersistState<R>& st, unsigned char* smem_raw) {
    typedef typename SVec<R>::T SV;
    const PersistLayout& L = a.lay;
    SV* s_in = reinterpret_cast<SV*>(smem_raw);
    const uint32_t* s_tidx = reinterpret_cast<const uint32_t*>(smem_raw + L.off_tidx);
    for (int c = 0; c < L.tiles_cached; ++c) {
        const int tile = blockIdx.x + c * gridDim.x;
        if (tile >= t.n_tiles) break;
        const int n_touched = int(t.tile_node_off[tile + 1] - t.tile_node_off[tile]), n_int = int(t.tile_nint[tile]);
        for (int k = threadIdx.x; k < n_touched; k += blockDim.x) {
            const uint32_t g = s_tidx[c * L.max_touched + k];
            R p0, p1, p2;
            persist_p_node<R>(st, a.r, g, p0, p1, p2);
            s_in[c * L.max_touched + k] = SVec<R>::make(p0, p1, p2);
            if (k < n_int) { R* d = st.pnew + 3 * size_t(g); d[0] = p0; d[1] = p1; d[2] = p2; }
        }
    }
    __syncthreads();
}

// phase 3 of a tile inside the CG loop: like tile_phase3, with the node's id / valence / mass / fixed flag from shared memory
template <class R> __device__ __forceinline__ double persist_phase3(const TileDev<R>& t, int tile, int c, const PersistCG<R>& a, unsigned char* smem_raw, const uint16_t* s_jds) {
    typedef typename SVec<R>::T SV;
    const PersistLayout& L = a.lay;
    const NodeEpilogue<R>& ep = a.ep;
    const SV* s_in = reinterpret_cast<const SV*>(smem_raw) + c * L.max_touched;
    const R* s_slot = reinterpret_cast<const R*>(smem_raw + L.off_slot);
    const NodeRec* s_nrec = reinterpret_cast<const NodeRec*>(smem_raw + L.off_nrec) + c * L.max_int;
    const int max_slots = L.max_slots;
    const int n_int = int(t.tile_nint[tile]);
    const bool plus = ep.sign > 0;
    double part = 0.0;
    for (int k = threadIdx.x; k < n_int; k += blockDim.x) {
        const NodeRec rec = s_nrec[k];
        const int val = int(rec.val_fixed & 0xFFFFu);
        const R m = node_rec_mass<R>(rec);
        const SV pv = s_in[k];
        R ax = R(0), ay = R(0), az = R(0);
        node_mass_m(ep, ep.pre_kind, m, R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
        int jj = 0;
        for (; jj + 4 <= val; jj += 4) {
            R cx[4], cy[4], cz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int s = s_jds[jj + u] + k; cx[u] = s_slot[s]; cy[u] = s_slot[max_slots + s]; cz[u] = s_slot[2 * max_slots + s]; }
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
            if (plus) { ax += s_slot[s]; ay += s_slot[max_slots + s]; az += s_slot[2 * max_slots + s]; }
            else { ax -= s_slot[s]; ay -= s_slot[max_slots + s]; az -= s_slot[2 * max_slots + s]; }
        }
        part += node_post_m(ep, rec.g, m, (rec.val_fixed & 0x10000u) != 0, R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
    }
    return part;
}

// Everything of an iteration after the tiles: returns false when the solve is over.  `part`: this thread's share of
// p.q over the interior nodes of the CTA's tiles.
template <class R> __device__ __forceinline__ bool persist_rest(const TileDev<R>& t, const PersistCG<R>& a, PersistState<R>& st, double part, double* red, double* bcast,
                                                                const NodeRec* s_grec, const uint32_t* s_gbase) {
    CGDev* cg = a.cg;
    const NodeEpilogue<R>& ep = a.ep;
    // ---- [B] shared nodes.  Everything that does not depend on the other CTAs is requested before the barrier.
    const NodeRec grec = s_grec[threadIdx.x];
    const bool has_node = grec.g != 0xFFFFFFFFu;
    R gp0 = R(0), gp1 = R(0), gp2 = R(0);
    if (has_node) persist_p_node<R>(st, a.r, grec.g, gp0, gp1, gp2);
    part = block_sum(part, red);
    trace_mark(ep.trace, kTraceTail, 1);
    const double den_tiles = grid_sync_sum(a.sync, st.sync_count, part, red, bcast);     // staged contributions are complete
    trace_mark(ep.trace, kTraceTail, 2);
    double part2 = 0.0;
    if (has_node) {
        const int val = int(grec.val_fixed & 0xFFFFu);
        const Quad<R>* stg = t.stage + s_gbase[threadIdx.x];
        const uint64_t pol = l2_policy_evict_first();
        Quad<R> b0[kGatherBatch], b1[kGatherBatch];
        gather_load<R>(b0, stg, 0, val, pol);
        gather_load<R>(b1, stg, kGatherBatch, val, pol);
        const R m = node_rec_mass<R>(grec);
        R* d = st.pnew + 3 * size_t(grec.g); d[0] = gp0; d[1] = gp1; d[2] = gp2;       // nobody else writes a shared node's p
        R ax = R(0), ay = R(0), az = R(0);
        node_mass_m(ep, ep.pre_kind, m, gp0, gp1, gp2, ax, ay, az);
        if (b0[0].a == R(123456789)) trace_mark(ep.trace, kTraceTail, 15);   // (waits for the first staged entry)
        trace_mark(ep.trace, kTraceTail, 8);
        gather_sum<R>(b0, b1, stg, val, ep.sign > 0, ax, ay, az, pol);
        if (ax == R(123456789)) trace_mark(ep.trace, kTraceTail, 15);
        trace_mark(ep.trace, kTraceTail, 9);
        part2 = node_post_m(ep, grec.g, m, (grec.val_fixed & 0x10000u) != 0, gp0, gp1, gp2, ax, ay, az);
    }
    __syncthreads();
    part2 = block_sum(part2, red);
    trace_mark(ep.trace, kTraceTail, 3);
    const double den = den_tiles + grid_sync_sum(a.sync, st.sync_count, part2, red, bcast);   // q and p_new are complete
    trace_mark(ep.trace, kTraceTail, 4);
    bool stop = false;
    if (den != 0.0) { if (fabs(den) <= st.thr && !(st.it == 1 && st.tsc == 0)) stop = true; } else stop = true;
    if (blockIdx.x == 0 && threadIdx.x == 0) cg_after_den(cg, den);
    if (stop) return false;
    // ---- [C]
    const double alpha_d = st.rho / den;
    const R alpha = R(alpha_d), malpha = R(-alpha_d);
    const bool a_one = (alpha_d == 1.0), ma_one = (-alpha_d == 1.0);
    const size_t stride = size_t(gridDim.x) * blockDim.x, t0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    typedef typename Vec4T<R>::T V;
    constexpr int N = Vec4T<R>::N;
    const size_t nv = a.n3 / N;
    V* xv = reinterpret_cast<V*>(a.x); V* rv = reinterpret_cast<V*>(a.r);
    const V* pv = reinterpret_cast<const V*>(st.pnew); const V* qv = reinterpret_cast<const V*>(ep.out);
    double prr = 0.0;
    for (size_t i = t0; i < nv; i += stride) {
        V xx = xv[i], rr = __ldcg(rv + i); const V pp = __ldcg(pv + i), qq = __ldcg(qv + i);
        prr += xr_vec(xx, rr, pp, qq, alpha, malpha, a_one, ma_one);
        xv[i] = xx; rv[i] = rr;
    }
    for (size_t i = nv * N + t0; i < a.n3; i += stride) { const R pp = __ldcg(st.pnew + i), qq = __ldcg(ep.out + i); prr += xr_one<R>(a.x[i], a.r[i], pp, qq, alpha, malpha, a_one, ma_one); }
    __syncthreads();
    prr = block_sum(prr, red);
    trace_mark(ep.trace, kTraceTail, 5);
    const double rho_new = grid_sync_sum(a.sync, st.sync_count, prr, red, bcast);          // x, r are complete
    trace_mark(ep.trace, kTraceTail, 6);
    const int it2 = st.it + 1;
    bool stop2 = unsigned(it2) > st.max_iter;
    if (!stop2) { const double err = sqrt(rho_new) / st.normb; if (err <= st.tol && !(it2 == 1 && st.tsc == 0)) stop2 = true; }
    if (blockIdx.x == 0 && threadIdx.x == 0) cg_after_rho(cg, rho_new);
    if (stop2) return false;
    st.beta = R(rho_new / st.rho);
    st.rho = rho_new; st.it = it2; st.first = false;
    R* tmp = st.pold; st.pold = st.pnew; st.pnew = tmp;
    return true;
}

}  // namespace sb
