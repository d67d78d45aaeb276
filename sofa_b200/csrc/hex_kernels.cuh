// HexahedronFEMForceField on the device (same tile / gather structure as the tetra force field).
//   addForce  : HexahedronFEMForceField.inl:194-246  (accumulateForceSmall :740-786, Large :836-884, Polar :1027-1076)
//   addDForce : HexahedronFEMForceField.inl:248-286
// The element stiffness K_e (24x24) is a per-element Data of the reference (`stiffnessMatrices`).  Matrices that are
// bit-identical (all elements of a grid whose spacing is exactly representable) are stored once: the element record keeps
// an index into a table of unique matrices, and a CTA whose tile refers to few matrices serves them from shared memory.
#pragma once
#include "cg_fused.cuh"
#include "math3.cuh"

namespace sb {

enum HexMode { HM_DF = 0, HM_F_SMALL = 1, HM_F_LARGE = 2, HM_F_POLAR = 3 };
constexpr int kHexSmemMatrices = 4;   // unique K_e cached per CTA in shared memory
constexpr int kHexKStride = 28;       // row stride of a cached K_e in the cooperative kernel: 24 padded to 28 floats, so that the eight lanes of a hexahedron
                                      // (rows 3w..3w+2, 84w floats apart) hit eight different 16-byte bank groups
constexpr int kHexKPadded = 24 * kHexKStride;

template <class R> struct HexDev {
    TileDev<R> t;
    const uint4* lnode;            // 8 x u16 local node ids (0xFFFF first = padding element)
    const uint4* slot_a; const uint4* slot_b;   // 8 contribution destinations
    Quad<R>* r0; Quad<R>* r1; Quad<R>* r2;      // _rotations[e] (9, row-major; R, not transposed) ; r2.d unused
    const uint32_t* kidx;          // index of the element's K_e in ktab
    const R* ktab;                 // [n_unique][576] row-major
    const uint32_t* tile_kuniq;    // [n_tiles][kHexSmemMatrices+1]: count (0 = too many: read K_e from global) then indices
    const Quad<R>* x0;             // 6 planes of n_slots: _rotatedInitialElements (24 Reals)
    size_t n_slots;
    R k_factor;
};

// F = K*Depl : Mat<24,24>*Vec<24>, row by row, left to right (HexahedronFEMForceField.inl:711-715, Mat.h:577-587)
template <class R> HD void hex_matvec(R F[24], const R* __restrict__ K, const R D[24]) {
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const R* row = K + 24 * i;
        R s = row[0] * D[0];
#pragma unroll
        for (int j = 1; j < 24; ++j) s += row[j] * D[j];
        F[i] = s;
    }
}

// mean edges of the 8 nodes (HexahedronFEMForceField.inl:797-801, 920-931)
template <class R> HD void hex_mean_edges(const V3<R> n[8], V3<R>& ex, V3<R>& ey, V3<R>& ez) {
    ex = (n[1] - n[0] + n[2] - n[3] + n[5] - n[4] + n[6] - n[7]) * R(.25);
    ey = (n[3] - n[0] + n[2] - n[1] + n[7] - n[4] + n[6] - n[5]) * R(.25);
    ez = (n[4] - n[0] + n[5] - n[1] + n[7] - n[3] + n[6] - n[2]) * R(.25);
}
// computeRotationLarge :816-834 / computeRotationPolar :918-943
template <class R> HD void hex_rotation(M3<R>& r, const V3<R> n[8], bool polar) {
    V3<R> ex, ey, ez;
    hex_mean_edges(n, ex, ey, ez);
    if (!polar) {
        normalize3(ex);
        V3<R> z = cross3(ex, ey);
        normalize3(z);
        ey = cross3(z, ex);
        set_row(r, 0, ex); set_row(r, 1, ey); set_row(r, 2, z);
    } else {
        M3<R> A;
        set_row(A, 0, ex); set_row(A, 1, ey); set_row(A, 2, ez);
        polar_decomposition(A, r);
    }
}

// One element.  P: the 8 nodal vectors (positions or dx); C: the 8 corner contributions as the reference adds/subtracts them.
template <class R, int MODE> HD void hex_element(const HexDev<R>& d, size_t es, const R* __restrict__ K, const V3<R> P[8], V3<R> C[8]) {
    R D[24], F[24];
    if (MODE == HM_DF) {
        const Quad<R> q0 = d.r0[es], q1 = d.r1[es], q2 = d.r2[es];
        M3<R> rot;
        rot.m[0][0] = q0.a; rot.m[0][1] = q0.b; rot.m[0][2] = q0.c; rot.m[1][0] = q0.d; rot.m[1][1] = q1.a; rot.m[1][2] = q1.b;
        rot.m[2][0] = q1.c; rot.m[2][1] = q1.d; rot.m[2][2] = q2.a;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const V3<R> x2 = mul(rot, P[w]); D[3 * w] = x2.x; D[3 * w + 1] = x2.y; D[3 * w + 2] = x2.z; }
        hex_matvec(F, K, D);
#pragma unroll
        for (int w = 0; w < 8; ++w) C[w] = mul_t(rot, mk3<R>(F[3 * w], F[3 * w + 1], F[3 * w + 2])) * d.k_factor;   // _rotations[i].multTranspose(F_w) * kFactor
    } else {
        R X0[24];
#pragma unroll
        for (int q = 0; q < 6; ++q) { const Quad<R> v = d.x0[size_t(q) * d.n_slots + es]; X0[4 * q] = v.a; X0[4 * q + 1] = v.b; X0[4 * q + 2] = v.c; X0[4 * q + 3] = v.d; }
        if (MODE == HM_F_SMALL) {
#pragma unroll
            for (int w = 0; w < 8; ++w) { D[3 * w] = X0[3 * w] - P[w].x; D[3 * w + 1] = X0[3 * w + 1] - P[w].y; D[3 * w + 2] = X0[3 * w + 2] - P[w].z; }
            hex_matvec(F, K, D);
#pragma unroll
            for (int w = 0; w < 8; ++w) C[w] = mk3<R>(F[3 * w], F[3 * w + 1], F[3 * w + 2]);
        } else {
            M3<R> rot;
            hex_rotation(rot, P, MODE == HM_F_POLAR);
            d.r0[es] = Quad<R>{rot.m[0][0], rot.m[0][1], rot.m[0][2], rot.m[1][0]};
            d.r1[es] = Quad<R>{rot.m[1][1], rot.m[1][2], rot.m[2][0], rot.m[2][1]};
            d.r2[es] = Quad<R>{rot.m[2][2], R(0), R(0), R(0)};
#pragma unroll
            for (int w = 0; w < 8; ++w) { const V3<R> def = mul(rot, P[w]); D[3 * w] = X0[3 * w] - def.x; D[3 * w + 1] = X0[3 * w + 1] - def.y; D[3 * w + 2] = X0[3 * w + 2] - def.z; }
            hex_matvec(F, K, D);
#pragma unroll
            for (int w = 0; w < 8; ++w) C[w] = mul_t(rot, mk3<R>(F[3 * w], F[3 * w + 1], F[3 * w + 2]));
        }
    }
}

template <class R> __host__ __device__ inline size_t hex_smem_bytes(int max_touched, int max_slots) {
    size_t a = tile_smem_bytes<R>(max_touched, max_slots);
    a = (a + 15) & ~size_t(15);
    return a + sizeof(R) * kHexKPadded * kHexSmemMatrices;
}

// phase 2 of a tile: one thread per hexahedron, 8 corner contributions scattered to their slots.  The tile's (few) distinct
// element stiffness matrices are cached in shared memory first (s_k: kHexSmemMatrices x 576 Reals); ends with no barrier.
// NTHR: number of threads that share the tile (0: the whole CTA, synchronised with __syncthreads; else the first NTHR threads, named barrier 1)
template <class R, int MODE, int NTHR = 0>
__device__ __forceinline__ void hex_tile_elements(const HexDev<R>& d, int tile, const typename SVec<R>::T* s_in, R* s_slot, int max_slots, R* s_k) {
    typedef typename SVec<R>::T SV;
    const TileDev<R>& t = d.t;
    const int nthr = NTHR > 0 ? NTHR : int(blockDim.x);
    const uint32_t* ku = d.tile_kuniq + size_t(tile) * (kHexSmemMatrices + 1);
    const int n_ku = int(ku[0]);
    for (int i = threadIdx.x; i < n_ku * 576; i += nthr) s_k[i] = d.ktab[size_t(ku[1 + i / 576]) * 576 + i % 576];
    if (NTHR > 0) bar_first<(NTHR > 0 ? NTHR : 32)>(); else __syncthreads();
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    for (int le = threadIdx.x; le < t.tile_e; le += nthr) {
        const size_t es = size_t(tile) * t.tile_e + le;
        const uint4 ln = idx_load(d.lnode + es, pol_stream);
        if ((ln.x & 0xFFFFu) == 0xFFFFu) continue;
        const uint4 sa = idx_load(d.slot_a + es, pol_stream), sb2 = idx_load(d.slot_b + es, pol_stream);
        const unsigned lid[8] = {ln.x & 0xFFFFu, ln.x >> 16, ln.y & 0xFFFFu, ln.y >> 16, ln.z & 0xFFFFu, ln.z >> 16, ln.w & 0xFFFFu, ln.w >> 16};
        V3<R> P[8], C[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) { const SV pv = s_in[lid[w]]; P[w] = mk3<R>(pv.x, pv.y, pv.z); }
        const uint32_t ki = d.kidx[es];
        const R* K = d.ktab + size_t(ki) * 576;
        for (int u = 0; u < n_ku; ++u) if (ku[1 + u] == ki) K = s_k + 576 * u;
        hex_element<R, MODE>(d, es, K, P, C);
        const unsigned s8[8] = {sa.x, sa.y, sa.z, sa.w, sb2.x, sb2.y, sb2.z, sb2.w};
#pragma unroll
        for (int w = 0; w < 8; ++w) tile_scatter<R>(t, s8[w], C[w].x, C[w].y, C[w].z, s_slot, max_slots, pol_keep);
    }
}
template <class R, int NTHR> __device__ __forceinline__ void hex_coop_elements(const HexDev<R>& d, int tile, const typename SVec<R>::T* s_in, R* s_slot, int max_slots, R* s_k);
// element policy of the fused CG kernel (cg_fused.cuh); the plan does not order a tile's elements that feed shared nodes first (arrive_at < 0)
template <class R> struct HexPass {
    typedef HexDev<R> Dev;
    static __device__ __forceinline__ const TileDev<R>& tiles(const Dev& d) { return d.t; }
    struct First {};
    static __device__ __forceinline__ void prefetch(const Dev&, int, First&) {}
    template <int ET, class OnBoundary>
    static __device__ __forceinline__ void elements(const Dev& d, int tile, const typename SVec<R>::T* s_in, R* s_slot, int max_slots, unsigned char* s_extra, int, OnBoundary, const First&) {
        if (sizeof(R) == 4) hex_coop_elements<R, ET>(d, tile, s_in, s_slot, max_slots, reinterpret_cast<R*>(s_extra));      // eight lanes per hexahedron
        else hex_tile_elements<R, HM_DF, ET>(d, tile, s_in, s_slot, max_slots, reinterpret_cast<R*>(s_extra));           // Vec3d: one thread per hexahedron (the 24 + 72 doubles of the cooperative pass spill)
    }
};

template <class R, int MODE>
__global__ void __launch_bounds__(256) hex_tile_kernel(HexDev<R> d, const R* __restrict__ in, NodeEpilogue<R> ep, int max_touched, int max_slots) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ uint16_t s_jds[1024];
    typedef typename SVec<R>::T SV;
    if (ep.cg && ep.cg->done) return;
    SV* s_in = reinterpret_cast<SV*>(smem_raw);
    size_t off = (sizeof(SV) * size_t(max_touched) + 15) & ~size_t(15);
    R* s_slot = reinterpret_cast<R*>(smem_raw + off);
    off = (off + sizeof(R) * 3 * size_t(max_slots) + 15) & ~size_t(15);
    R* s_k = reinterpret_cast<R*>(smem_raw + off);   // kHexSmemMatrices x 576

    const TileDev<R>& t = d.t;
    const int tile = blockIdx.x;
    tile_phase1<R>(t, tile, in, s_in, s_jds);   // ends with __syncthreads()
    hex_tile_elements<R, MODE>(d, tile, s_in, s_slot, max_slots, s_k);
    __syncthreads();
    const double part = tile_phase3<R>(t, tile, ep, s_in, s_slot, max_slots, s_jds);
    if (ep.dot_kind != DOT_NONE) {
        const double tot = block_sum(part, red);
        finish_dot(ep, tot, red, false);
    }
}

// phase 2 of a tile with eight lanes per hexahedron (see hex_tile_df_coop_kernel); NTHR as in hex_tile_elements.  s_k: kHexSmemMatrices padded matrices.
template <class R, int NTHR>
__device__ __forceinline__ void hex_coop_elements(const HexDev<R>& d, int tile, const typename SVec<R>::T* s_in, R* s_slot, int max_slots, R* s_k) {
    typedef typename SVec<R>::T SV;
    const TileDev<R>& t = d.t;
    const int nthr = NTHR > 0 ? NTHR : int(blockDim.x);
    const uint32_t* ku = d.tile_kuniq + size_t(tile) * (kHexSmemMatrices + 1);
    const int n_ku = int(ku[0]);
    for (int i = threadIdx.x; i < n_ku * 576; i += nthr) {
        const int m = i / 576, rc = i % 576;
        s_k[m * kHexKPadded + (rc / 24) * kHexKStride + rc % 24] = d.ktab[size_t(ku[1 + m]) * 576 + rc];
    }
    if (NTHR > 0) bar_first<(NTHR > 0 ? NTHR : 32)>(); else __syncthreads();
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const int w = threadIdx.x & 7, grp = threadIdx.x >> 3, ngrp = nthr >> 3;
    const int gbase = (threadIdx.x & 31) & ~7;           // first lane of the group inside its warp
    // One element, lane w of its group.  KREG: the tile refers to ONE stiffness matrix (every tile of a regular grid) and the lane keeps its three
    // rows in registers for the whole tile -- no shared-memory traffic for K_e at all; otherwise the rows come from the padded shared-memory copy
    // (16-byte loads) or from HBM.
    // what a lane reads from HBM for one element; requested one element ahead (the loop is bound by the latency of these loads otherwise: ncu showed
    // long-scoreboard stalls of 4.5 per issue with 16 warps per SM)
    struct Rec { unsigned lid, slot; Quad<R> q0, q1, q2; uint32_t ki; };
    auto load = [&](int le) {
        Rec r;
        const size_t es = size_t(tile) * t.tile_e + le;
        r.lid = reinterpret_cast<const uint16_t*>(d.lnode + es)[w];
        r.slot = w < 4 ? reinterpret_cast<const uint32_t*>(d.slot_a + es)[w] : reinterpret_cast<const uint32_t*>(d.slot_b + es)[w - 4];
        r.q0 = rec_load(d.r0 + es, pol_stream); r.q1 = rec_load(d.r1 + es, pol_stream); r.q2 = rec_load(d.r2 + es, pol_stream);
        r.ki = d.kidx[es];
        return r;
    };
    auto element = [&](const Rec& r, auto row_value) {
        const bool valid = __shfl_sync(0xffffffffu, r.lid, gbase) != 0xFFFFu;       // (a padding element has 0xFFFF in its first corner)
        M3<R> rot;
        rot.m[0][0] = r.q0.a; rot.m[0][1] = r.q0.b; rot.m[0][2] = r.q0.c; rot.m[1][0] = r.q0.d; rot.m[1][1] = r.q1.a; rot.m[1][2] = r.q1.b;
        rot.m[2][0] = r.q1.c; rot.m[2][1] = r.q1.d; rot.m[2][2] = r.q2.a;
        SV pv = SVec<R>::make(R(0), R(0), R(0));
        if (valid) pv = s_in[r.lid];
        const V3<R> x2 = mul(rot, mk3<R>(pv.x, pv.y, pv.z));                      // const Coord x_2 = _rotations[i] * dx[elem[w]]
        R D[24];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            D[3 * j] = __shfl_sync(0xffffffffu, x2.x, gbase + j); D[3 * j + 1] = __shfl_sync(0xffffffffu, x2.y, gbase + j); D[3 * j + 2] = __shfl_sync(0xffffffffu, x2.z, gbase + j);
        }
        R F[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            R s = row_value(c, 0) * D[0];
#pragma unroll
            for (int j = 1; j < 24; ++j) s += row_value(c, j) * D[j];
            F[c] = s;
        }
        const V3<R> C = mul_t(rot, mk3<R>(F[0], F[1], F[2])) * d.k_factor;       // _rotations[i].multTranspose(F_w) * kFactor
        if (valid) tile_scatter<R>(t, r.slot, C.x, C.y, C.z, s_slot, max_slots, pol_keep);
    };
    // tile_e is a multiple of 32 and a warp holds four consecutive groups: the loop condition is uniform inside a warp (full-mask shuffles are safe)
    if (n_ku == 1) {
        R kr[3][24];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 24; ++j) kr[c][j] = s_k[(3 * w + c) * kHexKStride + j];
        int le = grp;
        Rec cur;
        if (le < t.tile_e) cur = load(le);
        while (le < t.tile_e) {
            const int nle = le + ngrp;
            Rec nxt = cur;
            if (nle < t.tile_e) nxt = load(nle);
            element(cur, [&](int c, int j) { return kr[c][j]; });
            cur = nxt; le = nle;
        }
    } else {
        for (int le = grp; le < t.tile_e; le += ngrp) {
            const Rec r = load(le);
            int u_hit = -1;
            for (int u = 0; u < n_ku; ++u) if (ku[1 + u] == r.ki) u_hit = u;
            if (u_hit >= 0) { const R* Ks = s_k + kHexKPadded * u_hit + 3 * w * kHexKStride; element(r, [&](int c, int j) { return Ks[c * kHexKStride + j]; }); }
            else { const R* Kg = d.ktab + size_t(r.ki) * 576 + 3 * w * 24; element(r, [&](int c, int j) { return Kg[c * 24 + j]; }); }
        }
    }
}

// ---- addDForce, eight lanes per hexahedron -------------------------------------------------------------------------------------------
// The one-thread-per-hexahedron pass needs 255 registers (24 + 24 element values, 8 nodal vectors, the rotation), i.e. 8 warps per SM, and is
// issue-bound at a third of the fp32 rate.  Here lane w of a group of eight owns node w: it rotates its nodal vector (D_w = R p_w), the group
// exchanges the 24 values with shuffles, the lane forms rows 3w..3w+2 of F = K_e D -- each row the same left-to-right sum of 24 products as
// HexahedronFEMForceField.inl:711-715, so every bit of the result is the reference's -- rotates back and scatters its own corner.  When the tile refers to ONE matrix (every tile of a regular grid) the lane keeps its three
// rows of K_e in registers for the whole tile: two 256-thread CTAs per SM at ~128 registers (16 warps), no shared-memory traffic for K_e.  Otherwise
// the rows are read from the padded shared-memory copy, or from HBM when the tile refers to more than kHexSmemMatrices distinct matrices.
template <class R>
__global__ void __launch_bounds__(256, sizeof(R) == 4 ? 2 : 1) hex_tile_df_coop_kernel(HexDev<R> d, const R* __restrict__ in, NodeEpilogue<R> ep, int max_touched, int max_slots) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ uint16_t s_jds[1024];
    typedef typename SVec<R>::T SV;
    if (ep.cg && ep.cg->done) return;
    SV* s_in = reinterpret_cast<SV*>(smem_raw);
    size_t off = (sizeof(SV) * size_t(max_touched) + 15) & ~size_t(15);
    R* s_slot = reinterpret_cast<R*>(smem_raw + off);
    off = (off + sizeof(R) * 3 * size_t(max_slots) + 15) & ~size_t(15);
    R* s_k = reinterpret_cast<R*>(smem_raw + off);   // kHexSmemMatrices x 24 x kHexKStride
    const TileDev<R>& t = d.t;
    const int tile = blockIdx.x;
    tile_phase1<R>(t, tile, in, s_in, s_jds);   // ends with __syncthreads()
    hex_coop_elements<R, 0>(d, tile, s_in, s_slot, max_slots, s_k);
    __syncthreads();
    const double part = tile_phase3<R>(t, tile, ep, s_in, s_slot, max_slots, s_jds);
    if (ep.dot_kind != DOT_NONE) {
        const double tot = block_sum(part, red);
        finish_dot(ep, tot, red, false);
    }
}

// The whole CG loop in one persistent cooperative kernel: same skeleton as tet_cg_persistent_kernel (cg_persist.cuh), with the
// hexahedral element pass.
template <class R>
__global__ void __launch_bounds__(256) hex_cg_persistent_kernel(HexDev<R> d, PersistCG<R> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ double bcast;
    __shared__ GRec<R> s_grec[256];
    typedef typename SVec<R>::T SV;
    CGDev* cg = a.cg;
    if (cg->done) return;
    const TileDev<R>& t = d.t;
    const PersistLayout& L = a.lay;
    SV* s_in = reinterpret_cast<SV*>(smem_raw);
    R* s_slot = reinterpret_cast<R*>(smem_raw + L.off_slot);
    R* s_k = reinterpret_cast<R*>(smem_raw + L.off_extra);
    PersistState<R> st(a);
    persist_load_tables<R>(t, a, smem_raw, s_grec);
    if (!persist_init<R>(a, st, red, &bcast)) { persist_finish<R>(t, a, st, smem_raw, s_grec); return; }
    for (;;) {
        persist_phase1<R>(t, a, st, smem_raw);
        double part = 0.0;
        for (int c = 0; c < L.tiles_cached; ++c) {
            const int tile = blockIdx.x + c * gridDim.x;
            if (tile >= t.n_tiles) break;
            hex_tile_elements<R, HM_DF>(d, tile, s_in + c * L.max_touched, s_slot, L.max_slots, s_k);
            __syncthreads();
            part += persist_phase3<R>(t, tile, c, a, smem_raw);
            __syncthreads();
        }
        if (!persist_rest<R>(t, a, st, part, red, &bcast, smem_raw, s_grec)) break;
    }
    persist_finish<R>(t, a, st, smem_raw, s_grec);
}

template <class R> __global__ void hex_export_rotations_kernel(HexDev<R> d, const uint32_t* __restrict__ orig, R* __restrict__ out) {
    const size_t es = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (es >= d.n_slots) return;
    const uint32_t e = orig[es];
    if (e == 0xFFFFFFFFu) return;
    const Quad<R> q0 = d.r0[es], q1 = d.r1[es], q2 = d.r2[es];
    R* o = out + 9 * size_t(e);
    o[0] = q0.a; o[1] = q0.b; o[2] = q0.c; o[3] = q0.d; o[4] = q1.a; o[5] = q1.b; o[6] = q1.c; o[7] = q1.d; o[8] = q2.a;
}

// getNodeRotation, HexahedronFEMForceField.inl:946-974 (getRotations :976-1023): per node, identity + the sum of _rotations[h] * _initialrotations[h]^T
// over the hexahedra around it (ascending index), divided by their number, made orthogonal by polarDecomposition.  The reference starts
// the sum from the identity, not from zero; reproduced as is.  rot0 = _initialrotations in ORIGINAL element order.
template <class R> __global__ void hex_node_rotations_kernel(HexDev<R> d, const uint32_t* __restrict__ inc_off, const uint32_t* __restrict__ inc_es,
                                                             const uint32_t* __restrict__ inc_e, const R* __restrict__ rot0, R* __restrict__ out) {
    const size_t n = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (n >= size_t(d.t.n_nodes)) return;
    M3<R> acc;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc.m[i][j] = i == j ? R(1) : R(0);
    const uint32_t b = inc_off[n], e = inc_off[n + 1];
    for (uint32_t k = b; k < e; ++k) {
        const uint32_t es = inc_es[k];
        const Quad<R> q0 = d.r0[es], q1 = d.r1[es], q2 = d.r2[es];
        M3<R> rot, r0t;
        rot.m[0][0] = q0.a; rot.m[0][1] = q0.b; rot.m[0][2] = q0.c; rot.m[1][0] = q0.d; rot.m[1][1] = q1.a; rot.m[1][2] = q1.b;
        rot.m[2][0] = q1.c; rot.m[2][1] = q1.d; rot.m[2][2] = q2.a;
        const R* p = rot0 + 9 * size_t(inc_e[k]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) r0t.m[i][j] = p[3 * j + i];
        const M3<R> pr = mul(rot, r0t);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc.m[i][j] += pr.m[i][j];
    }
    const R cnt = R(e - b);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc.m[i][j] = acc.m[i][j] / cnt;
    M3<R> q;
    polar_decomposition(acc, q);
    R* o = out + 9 * n;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o[3 * i + j] = q.m[i][j];
}


}  // namespace sb
