// Device data layout shared by the tetra / hexa force fields: CTA tiles of elements, a deterministic
// gather plan, and the fused per-node epilogue (mass, projection, dot) of the solver node.
//
// Layout (see DESIGN.md "Data layout in HBM"):
//   * elements are sorted along a Morton curve of their rest centroid and cut into tiles of `tile_e`;
//     one CTA processes one tile.
//   * a node whose incident elements all lie in one tile is INTERIOR to it: its corner contributions
//     are staged in shared memory and summed in-kernel.  Every other node is SHARED: its contributions
//     are staged in HBM/L2 (`stage`) and summed by one thread per node -- in the boundary kernel
//     (gather_shared_kernel), in the fused CG tail, or inside the persistent CG kernel (cg_persist.cuh).
//   * in both cases the contributions of one node are added one by one in ascending ORIGINAL element
//     index, starting from the incoming value -- the exact order of the reference's sequential
//     `f[index[k]] += ...` loop (TetrahedronFEMForceField.inl:928-929,1220-1235).  No atomics:
//     results are bit-reproducible run to run and independent of the tiling.
//   * shared-memory slots are laid out as jagged diagonals: the interior nodes of a tile are ranked by
//     descending valence and contribution j of rank k lives at jds[j] + k (conflict-free sequential sums,
//     no padding).  The HBM stage is an ELL block per chunk of kGatherChunk consecutive shared nodes:
//     contribution j of the chunk's node k at sh_base[chunk] + j*kGatherChunk + k (coalesced, no table).
#pragma once
#include "common.cuh"

namespace sb {

constexpr int kGatherChunk = 128;      // shared nodes per chunk (= per CTA of the boundary kernel, per thread group of the CG kernels)
constexpr unsigned kStageFlag = 0x80000000u;

// epilogue selection for the per-node gather
enum PreKind { PRE_NONE = 0, PRE_GRAVITY = 1, PRE_MDX = 2 };
enum DotKind { DOT_NONE = 0, DOT_STORE = 1, DOT_CG_DEN = 2 };

// Device-resident CG scalars (CGLinearSolver.inl:73-315 keeps these on the host; here the whole
// solve runs without a host round trip, so they live in HBM).
constexpr int kMaxGraph = 1026;
struct CGDev {
    int done;               // set once a break condition is met; later kernels of the solve become no-ops
    int nb_iter;            // "CG iterations" as the reference reports it
    int end_cond;           // 0 iterations, 1 tolerance, 2 threshold, 3 den==0, 4 b==0, 99 a wait inside the persistent kernel timed out
    int it;                 // current iteration (1-based)
    unsigned time_step_count;
    unsigned max_iter;
    double tolerance, threshold;
    double normb, rho, rho_1, den, alpha;
    int n_err, n_den;
    double graph_error[kMaxGraph];
    double graph_den[kMaxGraph];
};

// PlaneForceField parameters after setPlane's normalisation (PlaneForceField.inl:139-145)
template <class R> struct PlaneDev { R nx, ny, nz, d, stiff, damp, limit2; int bilateral; };

template <class R> struct NodeEpilogue {
    const R* init_src;      // acc starts from init_src[node] (may alias out), or 0 when null
    int sign;               // +1: acc += c_i ; -1: acc -= c_i  (per contribution, in order)
    int pre_kind;           // mass term applied BEFORE the element contributions (mass first in the scene)
    int post_kind;          // ... or AFTER them
    const R* mass;          // DiagonalMass vertexMass
    const R* mdx_src;       // dx of addMDx (== the kernel's input vector p / v)
    R mass_factor;          // addMDx factor (narrowed to Real as the reference does)
    int mass_factor_is_one; // reference takes `res += dx*m` when factor == 1.0
    int mass_uniform;       // UniformMass: res += dx * um_f with um_f = vertexMass (* Real(factor) if factor != 1), UniformMass.inl:414-419
    R um_f;
    R gx, gy, gz;           // gravity (narrowed)
    int has_scale;          // b.teq(h)
    R scale;
    const unsigned char* fixed;  // projectResponse mask (1 = fixed), may be null
    R* out;
    // dot(out, dot_with) accumulated in double, per-CTA partials, fixed-order final sum by the last CTA
    int dot_kind;
    const R* dot_with;
    double* partials;       // [partial_base + blockIdx.x]
    int partial_base;
    int partial_total;      // number of partials to sum when finishing (tile CTAs + boundary CTAs)
    unsigned* counter;      // last-CTA detection (boundary kernel only)
    double* dot_result;
    CGDev* cg;
    unsigned long long* trace;   // diagnostics: per-CTA phase timestamps (null = off)
    // PlaneForceField as the node's LAST force field, fused: 0 none, 1 addForce (input vector = positions), 2 addDForce
    int plane_mode;
    PlaneDev<R> plane;
    R plane_fact;                          // addDForce: Real(-stiffness * kFactorIncludingRayleighDamping)
    const R* plane_v;                      // addForce: velocities (null = zero)
    unsigned char* plane_contacts;         // m_contacts as a per-node flag: written by addForce, read by addDForce
    const R* plane_in;                     // the pass's input vector, for the callers that do not hold the node's entry already
};
constexpr int kTraceWords = 16;          // words per CTA
constexpr int kTraceTail = 4096 * kTraceWords;   // the CG tail kernel's records start here
#ifdef __CUDA_ARCH__
// The stamps are taken by ALL lanes of the marking thread's warp and stored by one: a branch taken by a single lane would split that lane
// from its warp (measured: the warp's later shuffles then take the per-shuffle WARPSYNC slow path, microseconds per reduction tree).
__device__ __forceinline__ void trace_mark_by(unsigned long long* trace, int base, int i, int who) {
    if (trace && int(threadIdx.x >> 5) == (who >> 5)) {
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        unsigned sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
        if (int(threadIdx.x) == who) {
            trace[base + blockIdx.x * kTraceWords + i] = t;
            if (i == 0) trace[base + blockIdx.x * kTraceWords + 7] = sm;
        }
    }
}
__device__ __forceinline__ void trace_mark(unsigned long long* trace, int base, int i) { trace_mark_by(trace, base, i, 0); }
#else
inline void trace_mark(unsigned long long*, int, int) {}
inline void trace_mark_by(unsigned long long*, int, int, int) {}
#endif

template <class R> struct TileDev {
    int n_nodes, n_elems, n_tiles, tile_e, maxval;
    // per tile
    const uint32_t* tile_node_off;   // [n_tiles+1] into tile_nodes
    const uint32_t* tile_nodes;      // global node ids: interior (ranked by valence desc) then shared
    const uint32_t* tile_shslot;     // aligned with tile_nodes: a shared node's index in sh_nodes (fused CG kernel: the owner's state arrays are indexed by it)
    const uint32_t* tile_nint;       // [n_tiles]
    const uint32_t* tile_nb;         // [n_tiles] leading elements of the tile that feed shared nodes (null: unknown)
    const uint16_t* tile_val;        // valence of each interior node, aligned with tile_nodes (shared entries unused)
    const uint16_t* tile_jds;        // [n_tiles][maxval+1]
    // shared nodes
    int n_shared, n_chunks;
    const uint32_t* sh_nodes;        // [n_chunks*kGatherChunk] (padded with 0xFFFFFFFF)
    const uint16_t* sh_val;
    const uint32_t* sh_base;         // [n_chunks] first entry of the chunk's ELL block: contribution j of node k at sh_base + j*kGatherChunk + k
    Quad<R>* stage;                  // one (x,y,z,-) quad per staged contribution: a single 16-byte store / load
    size_t stage_n;
};

// L2 residency control (sm_80+ createpolicy / cache_hint).  The staged corner contributions are written by the element
// pass and read once by the boundary kernel a few microseconds later: they are stored evict_last so that they are still in
// the 126 MB L2 when read (no HBM round trip), while the element records, which stream through exactly once per pass, are
// loaded evict_first so that they do not push the staged data out.
HD uint64_t l2_policy_evict_last() {
    uint64_t p = 0;
#ifdef __CUDA_ARCH__
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
#endif
    return p;
}
HD uint64_t l2_policy_evict_first() {
    uint64_t p = 0;
#ifdef __CUDA_ARCH__
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
#endif
    return p;
}
#ifdef __CUDA_ARCH__
__device__ __forceinline__ float4 ldg_hint(const float4* a, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ double2 ldg_hint(const double2* a, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ uint4 ldg_hint(const uint4* a, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ uint2 ldg_hint(const uint2* a, uint64_t pol) {
    uint2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg_hint(float4* a, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_hint(double2* a, double2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" :: "l"(a), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
#endif
// pol: an L2 cache policy on the device; ignored on the host (tests/emu executes the same functions on the CPU)
HD void stage_store(Quad<float>* p, float x, float y, float z, uint64_t pol) {
#ifdef __CUDA_ARCH__
    stg_hint(reinterpret_cast<float4*>(p), make_float4(x, y, z, 0.f), pol);
#else
    (void)pol; *p = Quad<float>{x, y, z, 0.f};
#endif
}
HD void stage_store(Quad<double>* p, double x, double y, double z, uint64_t pol) {
#ifdef __CUDA_ARCH__
    stg_hint(reinterpret_cast<double2*>(p), make_double2(x, y), pol);
    stg_hint(reinterpret_cast<double2*>(p) + 1, make_double2(z, 0.0), pol);
#else
    (void)pol; *p = Quad<double>{x, y, z, 0.0};
#endif
}
// without a cache hint (no policy descriptor to move into uniform registers per store); the fourth word is padding and is never read as a value
__device__ __forceinline__ void stage_store_plain(unsigned long long a, float x, float y, float z) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%3};" :: "l"(a), "f"(x), "f"(y), "f"(z) : "memory");
}
__device__ __forceinline__ void stage_store_plain(unsigned long long a, double x, double y, double z) {
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" :: "l"(a), "d"(x), "d"(y) : "memory");
    asm volatile("st.global.v2.f64 [%0+16], {%1,%1};" :: "l"(a), "d"(z) : "memory");
}
HD Quad<float> stage_load(const Quad<float>* p, uint64_t pol) {
#ifdef __CUDA_ARCH__
    const float4 v = ldg_hint(reinterpret_cast<const float4*>(p), pol);
    return Quad<float>{v.x, v.y, v.z, v.w};
#else
    (void)pol; return *p;
#endif
}
HD Quad<double> stage_load(const Quad<double>* p, uint64_t pol) {
#ifdef __CUDA_ARCH__
    const double2 a = ldg_hint(reinterpret_cast<const double2*>(p), pol), b = ldg_hint(reinterpret_cast<const double2*>(p) + 1, pol);
    return Quad<double>{a.x, a.y, b.x, b.y};
#else
    (void)pol; return *p;
#endif
}
// index words of the element records
HD uint2 idx_load(const uint2* p, uint64_t pol) {
#ifdef __CUDA_ARCH__
    return ldg_hint(p, pol);
#else
    (void)pol; return *p;
#endif
}
HD uint4 idx_load(const uint4* p, uint64_t pol) {
#ifdef __CUDA_ARCH__
    return ldg_hint(p, pol);
#else
    (void)pol; return *p;
#endif
}
// element records: streamed once per pass
HD Quad<float> rec_load(const Quad<float>* p, uint64_t pol) { return stage_load(p, pol); }
HD Quad<double> rec_load(const Quad<double>* p, uint64_t pol) { return stage_load(p, pol); }

// acc (op)= contribution, in the reference's order; then mass/scale/projection/dot.
template <class R> HD void node_pre(const NodeEpilogue<R>& ep, uint32_t g, R& ax, R& ay, R& az) {
    if (ep.init_src) { ax = ep.init_src[3 * size_t(g)]; ay = ep.init_src[3 * size_t(g) + 1]; az = ep.init_src[3 * size_t(g) + 2]; }
    else { ax = R(0); ay = R(0); az = R(0); }
}
// (vx,vy,vz): the node's entry of mdx_src / dot_with when the caller already holds it (shared-memory copy of the input vector)
// m: the node's vertexMass (already loaded)
template <class R> HD void node_mass_m(const NodeEpilogue<R>& ep, int kind, R m, R vx, R vy, R vz, R& ax, R& ay, R& az) {
    if (kind == PRE_GRAVITY) {  // DiagonalMass::addForce: f[i] += theGravity*masses[i]
        ax += ep.gx * m; ay += ep.gy * m; az += ep.gz * m;
    } else if (kind == PRE_MDX) {  // DiagonalMass::addMDx / UniformMass::addMDx
        if (ep.mass_uniform) { ax += vx * ep.um_f; ay += vy * ep.um_f; az += vz * ep.um_f; }
        else if (ep.mass_factor_is_one) { ax += vx * m; ay += vy * m; az += vz * m; }
        else { ax += (vx * m) * ep.mass_factor; ay += (vy * m) * ep.mass_factor; az += (vz * m) * ep.mass_factor; }
    }
}
template <class R> HD void node_mass_v(const NodeEpilogue<R>& ep, int kind, uint32_t g, R vx, R vy, R vz, R& ax, R& ay, R& az) {
    if (kind == PRE_GRAVITY || kind == PRE_MDX) node_mass_m(ep, kind, ep.mass[g], vx, vy, vz, ax, ay, az);
}
template <class R> HD void node_mass(const NodeEpilogue<R>& ep, int kind, uint32_t g, R& ax, R& ay, R& az) {
    R vx = 0, vy = 0, vz = 0;
    if (kind == PRE_MDX) { vx = ep.mdx_src[3 * size_t(g)]; vy = ep.mdx_src[3 * size_t(g) + 1]; vz = ep.mdx_src[3 * size_t(g) + 2]; }
    node_mass_v(ep, kind, g, vx, vy, vz, ax, ay, az);
}
// PlaneForceField::addForce :158-205 / addDForce :208-226 for one node, (vx, vy, vz) = the node's entry of the pass's input vector
template <class R> HD void node_plane(const NodeEpilogue<R>& ep, uint32_t g, R vx, R vy, R vz, R& ax, R& ay, R& az) {
    const PlaneDev<R>& P = ep.plane;
    if (ep.plane_mode == 2) {
        if (ep.plane_contacts[g]) {
            R s = vx * P.nx; s += vy * P.ny; s += vz * P.nz;
            const R t = ep.plane_fact * s;
            ax = ax + P.nx * t; ay = ay + P.ny * t; az = az + P.nz * t;
        }
    } else if (ep.plane_mode == 1) {
        R d = vx * P.nx; d += vy * P.ny; d += vz * P.nz;
        d = d - P.d;
        unsigned char hit = 0;
        if (P.bilateral || d < R(0)) {
            const R fi = -P.stiff * d, di = -P.damp * d;
            const R w0 = ep.plane_v ? ep.plane_v[3 * size_t(g)] : R(0), w1 = ep.plane_v ? ep.plane_v[3 * size_t(g) + 1] : R(0), w2 = ep.plane_v ? ep.plane_v[3 * size_t(g) + 2] : R(0);
            R f0 = P.nx * fi - w0 * di, f1 = P.ny * fi - w1 * di, f2 = P.nz * fi - w2 * di;
            R amp = f0 * f0; amp += f1 * f1; amp += f2 * f2;
            if (P.limit2 > R(0) && amp > P.limit2) { const R sc = sqrt(P.limit2 / amp); f0 *= sc; f1 *= sc; f2 *= sc; }
            ax += f0; ay += f1; az += f2;
            hit = 1;
        }
        ep.plane_contacts[g] = hit;
    }
}
// returns this node's contribution to the dot product
template <class R> HD double node_post_v(const NodeEpilogue<R>& ep, uint32_t g, R vx, R vy, R vz, R ax, R ay, R az) {
    node_mass_v(ep, ep.post_kind, g, vx, vy, vz, ax, ay, az);
    if (ep.plane_mode) node_plane(ep, g, vx, vy, vz, ax, ay, az);
    if (ep.has_scale) { ax *= ep.scale; ay *= ep.scale; az *= ep.scale; }
    if (ep.fixed && ep.fixed[g]) { ax = R(0); ay = R(0); az = R(0); }
    ep.out[3 * size_t(g)] = ax; ep.out[3 * size_t(g) + 1] = ay; ep.out[3 * size_t(g) + 2] = az;
    if (ep.dot_kind != DOT_NONE) return double(ax) * double(vx) + double(ay) * double(vy) + double(az) * double(vz);
    return 0.0;
}
// same with the node's mass and fixed flag already loaded
// ... and the result left in (ax, ay, az) instead of ep.out
template <class R> HD double node_finish_m(const NodeEpilogue<R>& ep, uint32_t g, R m, bool is_fixed, R vx, R vy, R vz, R& ax, R& ay, R& az) {
    node_mass_m(ep, ep.post_kind, m, vx, vy, vz, ax, ay, az);
    if (ep.plane_mode) node_plane(ep, g, vx, vy, vz, ax, ay, az);
    if (ep.has_scale) { ax *= ep.scale; ay *= ep.scale; az *= ep.scale; }
    if (is_fixed) { ax = R(0); ay = R(0); az = R(0); }
    if (ep.dot_kind != DOT_NONE) return double(ax) * double(vx) + double(ay) * double(vy) + double(az) * double(vz);
    return 0.0;
}
template <class R> HD double node_post_m(const NodeEpilogue<R>& ep, uint32_t g, R m, bool is_fixed, R vx, R vy, R vz, R ax, R ay, R az) {
    const double d = node_finish_m(ep, g, m, is_fixed, vx, vy, vz, ax, ay, az);
    ep.out[3 * size_t(g)] = ax; ep.out[3 * size_t(g) + 1] = ay; ep.out[3 * size_t(g) + 2] = az;
    return d;
}
template <class R> HD double node_post(const NodeEpilogue<R>& ep, uint32_t g, R ax, R ay, R az) {
    R vx = 0, vy = 0, vz = 0;
    const R* src = ep.dot_kind != DOT_NONE ? ep.dot_with : (ep.post_kind == PRE_MDX ? ep.mdx_src : nullptr);
    if (ep.plane_mode) src = ep.plane_in;
    if (src) { vx = src[3 * size_t(g)]; vy = src[3 * size_t(g) + 1]; vz = src[3 * size_t(g) + 2]; }
    return node_post_v(ep, g, vx, vy, vz, ax, ay, az);
}

// block-wide sum in a fixed order: warp shuffles, then warp 0 over the per-warp partials
__device__ __forceinline__ double block_sum(double v, double* warp_scratch /* >= 32 doubles of smem */) {
    __syncwarp();   // (callers come out of loops with lane-dependent trip counts: without reconvergence every shuffle takes the WARPSYNC slow path)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_scratch[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? warp_scratch[threadIdx.x] : 0.0;
    if (w == 0) {
        __syncwarp();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    return v;  // valid in thread 0
}

// CG bookkeeping executed by one thread once den = p.q is known (CGLinearSolver.inl:211-265)
__host__ __device__ inline void cg_after_den(CGDev* cg, double den) {
    if (cg->n_den < kMaxGraph) cg->graph_den[cg->n_den++] = den;
    cg->den = den;
    if (den != 0.0) {
        if (fabs(den) <= cg->threshold) {
            if (cg->it == 1 && cg->time_step_count == 0) { /* reference only warns on the very first step */ }
            else { cg->done = 1; cg->nb_iter = cg->it; cg->end_cond = 2; return; }
        }
        cg->alpha = cg->rho / den;
    } else { cg->done = 1; cg->nb_iter = cg->it; cg->end_cond = 3; }
}
// ... and once the next rho = r.r is known (CGLinearSolver.inl:130-180,268)
__host__ __device__ inline void cg_after_rho(CGDev* cg, double rho_new) {
    cg->rho_1 = cg->rho;
    cg->rho = rho_new;
    const int it = cg->it + 1;
    if (unsigned(it) > cg->max_iter) { cg->done = 1; cg->nb_iter = it; cg->end_cond = 0; return; }
    cg->it = it;
    const double err = sqrt(rho_new) / cg->normb;
    if (cg->n_err < kMaxGraph) cg->graph_error[cg->n_err++] = err;
    if (err <= cg->tolerance) {
        if (it == 1 && cg->time_step_count == 0) { /* warning only */ }
        else { cg->done = 1; cg->nb_iter = it; cg->end_cond = 1; }
    }
}

// Finish a dot product whose per-CTA partials are complete: the LAST CTA to arrive sums them in index order.
template <class R> __device__ __forceinline__ void finish_dot(const NodeEpilogue<R>& ep, double block_total, double* scratch, bool count_blocks) {
    if (ep.dot_kind == DOT_NONE) return;
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        ep.partials[ep.partial_base + blockIdx.x] = block_total;
        is_last = false;
        if (count_blocks) {
            __threadfence();
            const unsigned ticket = atomicInc(ep.counter, gridDim.x - 1);  // wraps to 0 for the next launch
            is_last = (ticket == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!count_blocks || !is_last) return;
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < ep.partial_total; i += blockDim.x) s += __ldcg(ep.partials + i);
    __syncthreads();
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) {
        if (ep.dot_result) *ep.dot_result = s;
        if (ep.dot_kind == DOT_CG_DEN) cg_after_den(ep.cg, s);
    }
}

// ---- pieces of the tile kernels shared by the tetra and hexa force fields -------------------------------------------
// shared-memory staging of nodal vectors: float -> float4 (one LDS.128), double -> 3 doubles
template <class R> struct SVec;
template <> struct SVec<float> { typedef float4 T; static __device__ __forceinline__ T make(float x, float y, float z) { return make_float4(x, y, z, 0.f); } };
template <> struct SVec<double> { struct T { double x, y, z; }; static __device__ __forceinline__ T make(double x, double y, double z) { T t; t.x = x; t.y = y; t.z = z; return t; } };

template <class R> __host__ __device__ inline size_t tile_smem_bytes(int max_touched, int max_slots) {
    size_t a = sizeof(typename SVec<R>::T) * size_t(max_touched);
    a = (a + 15) & ~size_t(15);
    return a + sizeof(R) * 3 * size_t(max_slots);
}
// phase 1: stage the input vector of the tile's touched nodes (and the tile's jagged-diagonal table) in shared memory
template <class R> __device__ __forceinline__ void tile_phase1(const TileDev<R>& t, int tile, const R* __restrict__ in, typename SVec<R>::T* s_in, uint16_t* s_jds) {
    const uint32_t node_off = t.tile_node_off[tile];
    const int n_touched = int(t.tile_node_off[tile + 1] - node_off);
    // two nodes per thread and per round: both node ids are requested before either nodal vector (dependent loads overlap)
    for (int k = threadIdx.x; k < n_touched; k += 2 * blockDim.x) {
        const int k2 = k + blockDim.x;
        const uint32_t g1 = t.tile_nodes[node_off + k];
        const uint32_t g2 = k2 < n_touched ? t.tile_nodes[node_off + k2] : g1;
        const R* p1 = in + 3 * size_t(g1); const R* p2 = in + 3 * size_t(g2);
        const R a0 = p1[0], a1 = p1[1], a2 = p1[2], b0 = p2[0], b1 = p2[1], b2 = p2[2];
        s_in[k] = SVec<R>::make(a0, a1, a2);
        if (k2 < n_touched) s_in[k2] = SVec<R>::make(b0, b1, b2);
    }
    for (int j = threadIdx.x; j <= t.maxval; j += blockDim.x) s_jds[j] = t.tile_jds[size_t(tile) * (t.maxval + 1) + j];
    __syncthreads();
}
// phase 2 tail: one corner contribution to its slot (shared memory for interior nodes, L2-resident HBM stage for shared ones).
// Shared-memory slots are (x, y, z) triples at a stride of three Reals: one address per contribution, the other two stores at immediate
// offsets, and the ordered sums (consecutive threads -> consecutive slots) stay conflict-free because 3 is odd.  The staging address is
// one 32 x 32 -> 64-bit multiply-add.  (The scatter is ~20 % of the element loop's issue slots: measured, profiles/README.md.)
template <class R> __device__ __forceinline__ void tile_scatter(const TileDev<R>& t, unsigned s, R cx, R cy, R cz, R* s_slot, int max_slots, uint64_t pol_keep) {
    (void)max_slots; (void)pol_keep;
    if (s & kStageFlag) {
        unsigned long long addr;
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(s & ~kStageFlag), "r"(unsigned(sizeof(Quad<R>))), "l"(reinterpret_cast<unsigned long long>(t.stage)));
        stage_store_plain(addr, cx, cy, cz);
    } else { R* dst = s_slot + 3u * s; dst[0] = cx; dst[1] = cy; dst[2] = cz; }
}
// phase 3: interior nodes, sequential sum in element order + fused epilogue; returns the thread's share of the dot product.
// mdx_src / dot_with are the kernel's own input vector whenever they are used (A*p: both are p), so the shared-memory
// copy staged in phase 1 serves them.
template <class R> __device__ __forceinline__ double tile_phase3(const TileDev<R>& t, int tile, const NodeEpilogue<R>& ep, const typename SVec<R>::T* s_in,
                                                                 const R* s_slot, int max_slots, const uint16_t* s_jds) {
    const uint32_t node_off = t.tile_node_off[tile];
    const int n_int = int(t.tile_nint[tile]);
    double part = 0.0;
    for (int k = threadIdx.x; k < n_int; k += blockDim.x) {
        const uint32_t g = t.tile_nodes[node_off + k];
        const int val = t.tile_val[node_off + k];
        const typename SVec<R>::T pv = s_in[k];
        R ax, ay, az;
        node_pre(ep, g, ax, ay, az);
        node_mass_v(ep, ep.pre_kind, g, R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
        // slots are read four contributions ahead of the adds; the adds themselves stay strictly in element order
        const bool plus = ep.sign > 0;
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
        part += node_post_v(ep, g, R(pv.x), R(pv.y), R(pv.z), ax, ay, az);
    }
    return part;
}

// ---- shared nodes: ordered sum of the staged contributions ---------------------------------------------------------
// One thread per node.  The node's contributions are fetched in batches of kGatherBatch independent 16-byte loads, three
// batches in flight, and added ONE BY ONE IN ELEMENT ORDER.  st: the node's first staged entry (stride kGatherChunk).
constexpr int kGatherBatch = 8;
template <class R> __device__ __forceinline__ void gather_load(Quad<R> (&b)[kGatherBatch], const Quad<R>* st, int j, int val, uint64_t pol) {
    const Quad<R> zero{R(0), R(0), R(0), R(0)};
#pragma unroll
    for (int u = 0; u < kGatherBatch; ++u) b[u] = (j + u < val) ? stage_load(st + size_t(j + u) * kGatherChunk, pol) : zero;
}
template <class R> __device__ __forceinline__ void gather_add(const Quad<R> (&b)[kGatherBatch], int j, int val, bool plus, R& ax, R& ay, R& az) {
#pragma unroll
    for (int u = 0; u < kGatherBatch; ++u) {
        if (j + u < val) {
            if (plus) { ax += b[u].a; ay += b[u].b; az += b[u].c; } else { ax -= b[u].a; ay -= b[u].b; az -= b[u].c; }
        }
    }
}
// first two batches are requested by the caller as early as possible (gather_load b0 @0, b1 @kGatherBatch)
template <class R> __device__ __forceinline__ void gather_sum(Quad<R> (&b0)[kGatherBatch], Quad<R> (&b1)[kGatherBatch], const Quad<R>* st, int val, bool plus,
                                                              R& ax, R& ay, R& az, uint64_t pol) {
    constexpr int B = kGatherBatch;
    Quad<R> b2[B];
    gather_load<R>(b2, st, 2 * B, val, pol);
    for (int j0 = 0; j0 < val; j0 += 3 * B) {
        gather_add<R>(b0, j0, val, plus, ax, ay, az);
        if (j0 + B >= val) break;
        gather_load<R>(b0, st, j0 + 3 * B, val, pol);
        gather_add<R>(b1, j0 + B, val, plus, ax, ay, az);
        if (j0 + 2 * B >= val) break;
        gather_load<R>(b1, st, j0 + 4 * B, val, pol);
        gather_add<R>(b2, j0 + 2 * B, val, plus, ax, ay, az);
        gather_load<R>(b2, st, j0 + 5 * B, val, pol);
    }
}
// node k of `chunk`; returns the node's share of the dot product
template <class R> __device__ __forceinline__ double gather_chunk(const TileDev<R>& d, const NodeEpilogue<R>& ep, int chunk, int k) {
    const uint32_t g = d.sh_nodes[size_t(chunk) * kGatherChunk + k];
    if (g == 0xFFFFFFFFu) return 0.0;
    const int val = int(d.sh_val[size_t(chunk) * kGatherChunk + k]);
    const Quad<R>* st = d.stage + d.sh_base[chunk] + k;
    const uint64_t pol = l2_policy_evict_first();
    Quad<R> b0[kGatherBatch], b1[kGatherBatch];
    gather_load<R>(b0, st, 0, val, pol);
    gather_load<R>(b1, st, kGatherBatch, val, pol);
    R ax, ay, az;
    node_pre(ep, g, ax, ay, az);
    node_mass(ep, ep.pre_kind, g, ax, ay, az);
    gather_sum<R>(b0, b1, st, val, ep.sign > 0, ax, ay, az, pol);
    return node_post(ep, g, ax, ay, az);
}
template <class R> __global__ void __launch_bounds__(kGatherChunk) gather_shared_kernel(TileDev<R> d, NodeEpilogue<R> ep) {
    __shared__ double red[32];
    if (ep.cg && ep.cg->done) return;
    const double part = gather_chunk<R>(d, ep, blockIdx.x, threadIdx.x);
    if (ep.dot_kind != DOT_NONE) {
        const double tot = block_sum(part, red);
        finish_dot(ep, tot, red, true);
    }
}

}  // namespace sb
