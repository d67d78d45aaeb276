// FastTetrahedralCorotationalForceField<B200Vec3Types>:
//   Sofa/Component/SolidMechanics/FEM/Elastic/src/sofa/component/solidmechanics/fem/elastic/FastTetrahedralCorotationalForceField.{h,inl}  -- [FTC]
// The class keeps, per tetrahedron, the six 3x3 edge blocks of the linear stiffness (linearDfDx), the rest edge vectors and the rotation of the
// last addForce ([FTC].h:83-106); addDForce ([FTC].inl:402-470) runs over the EDGES of the topology with one 3x3 matrix per edge, assembled
// from the tetrahedra at the first call after each addForce.  Device layout:
//   * addForce      one tile pass over the tetrahedra (the plan of fem_layout.cuh with 4 corners per element): rotation, six edge forces,
//                   four corner contributions summed per node in ascending tetrahedron index; the rotation (R^T, as the class stores it)
//                   is written back per tetrahedron;
//   * edge matrices the addForce pass also leaves, per tetrahedron, the six 3x3 blocks R^T (L_j R) / (L_j R)^T R it will add to its edges' matrices;
//                   one thread per edge then sums the blocks of the tetrahedra around it in ascending index ([FTC].inl:414-450) -- run by the
//                   first addDForce after an addForce;
//   * addDForce     one tile pass over the EDGES (the same plan machinery with 2 corners per element): df[e1] += M dx, df[e0] -= M^T dx in
//                   ascending edge index per node, the order of the reference's loop.
// Both passes end in the shared fused epilogue (mass term, projection, dot product), so a solver node drives this class like the others.
// Edge numbering: TetrahedronSetTopologyContainer::createEdgesInTetrahedronArray (first appearance over the tetrahedra, local edges
// {0,1},{0,2},{0,3},{1,2},{1,3},{2,3}, vertices sorted), or the caller's list when the topology already holds edges.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>

#include "cg_fused.cuh"
#include "fem_layout.cuh"
#include "math3.cuh"
#include "plan.h"
#include "tet_handle.h"

using namespace sb;

namespace sb {

enum FastMethod { FAST_POLAR = 0, FAST_QR = 1, FAST_POLAR2 = 2, FAST_LINEAR = 3 };   // RotationDecompositionMethod, [FTC].h:70-76
// per-tetrahedron record, one plane of NS Reals per entry (coalesced over the element slots)
constexpr int kFastEdgeVec = 0;      // restEdgeVector[6]            18
constexpr int kFastDfDx = 18;        // linearDfDx[6], row-major     54
constexpr int kFastRestRot = 72;     // restRotation                  9
constexpr int kFastShape = 81;       // shapeVector[1..3]             9
constexpr int kFastRec = 90;
constexpr int kFastPmat = 56;        // six 3x3 matrices per tetrahedron, padded to a multiple of four Reals (written as 16-byte quads)
// edgesInTetrahedronArray, core/topology/Topology.cpp:44: {0,1},{0,2},{0,3},{1,2},{1,3},{2,3}
static const int kFastLh[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

template <class R> struct FastDev {
    TileDev<R> t;                 // the plan of the pass being run (tetrahedra or edges)
    const ushort4* lnode; const uint4* slot;        // tetrahedra: local node and destination of each corner
    const R* rec; size_t NS;      // rec[k * NS + es]
    R* rot;                       // rot[k * NS + es], k < 9: tetraInfo.rotation (the transposed rotation)
    R* pmat;                      // pmat[es * kFastPmat + 9 j + k]: the element's contribution to the matrix of its j-th edge, R^T (L_j R) or its transpose
    const unsigned char* orient;  // bit j: edgeOrientation[j] == 1
    const ushort2* elnode; const uint2* eslot;      // edges
    const R* emat; size_t NSe;    // emat[k * NSe + es], k < 9: edgeDfDx
    R k_factor;
};

template <class R> __device__ __forceinline__ M3<R> fast_load_mat(const R* base, size_t stride, size_t es) {
    M3<R> m;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) m.m[i][j] = base[size_t(3 * i + j) * stride + es];
    return m;
}
// computeQRRotation, [FTC].inl:272-294
template <class R> HD void fast_qr_rotation(M3<R>& r, const V3<R>& d0, const V3<R>& d1) {
    V3<R> edgex = d0;
    normalize3(edgex);
    V3<R> edgey = d1;
    V3<R> edgez = cross3(edgex, edgey);
    normalize3(edgez);
    edgey = cross3(edgez, edgex);
    set_row(r, 0, edgex); set_row(r, 1, edgey); set_row(r, 2, edgez);
}

// addForce of one tetrahedron, [FTC].inl:312-393
template <class R, int METHOD> __device__ __forceinline__ void fast_tet_element(const FastDev<R>& d, size_t es, const V3<R> P[4], V3<R> C[4]) {
    const int L0[6] = {0, 0, 0, 1, 1, 2}, L1[6] = {1, 2, 3, 2, 3, 3};
    V3<R> displ[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) displ[j] = P[L1[j]] - P[L0[j]];
    M3<R> Rm;
    if (METHOD == FAST_POLAR) {
        M3<R> F;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const R sv[3] = {d.rec[size_t(kFastShape + 3 * j) * d.NS + es], d.rec[size_t(kFastShape + 3 * j + 1) * d.NS + es], d.rec[size_t(kFastShape + 3 * j + 2) * d.NS + es]};
            const R dv[3] = {displ[j].x, displ[j].y, displ[j].z};
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    if (j == 0) F.m[k][l] = dv[k] * sv[l];
                    else F.m[k][l] += dv[k] * sv[l];
                }
        }
        polar_decomposition(F, Rm);
    } else if (METHOD == FAST_QR) {
        M3<R> S;
        fast_qr_rotation(S, displ[0], displ[1]);
        const M3<R> rest = fast_load_mat(d.rec + size_t(kFastRestRot) * d.NS, d.NS, es);
        Rm = mul_atb(S, rest);                       // S.multTranspose(restRotation)
    } else if (METHOD == FAST_POLAR2) {
        M3<R> S;
        set_row(S, 0, displ[0]); set_row(S, 1, displ[1]); set_row(S, 2, displ[2]);
        polar_decomposition(S, Rm);
        const M3<R> rest = fast_load_mat(d.rec + size_t(kFastRestRot) * d.NS, d.NS, es);
        Rm = mul(transpose(Rm), rest);               // R.transposed() * restRotation
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Rm.m[i][j] = i == j ? R(1) : R(0);
    }
    const M3<R> rot = transpose(Rm);                 // tetraInfo.rotation
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) d.rot[size_t(3 * i + j) * d.NS + es] = rot.m[i][j];
    V3<R> force[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) force[n] = mk3<R>(R(0), R(0), R(0));
    const unsigned orient = d.orient[es];
    R pm[kFastPmat];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const V3<R> rest_edge = mk3<R>(d.rec[size_t(kFastEdgeVec + 3 * j) * d.NS + es], d.rec[size_t(kFastEdgeVec + 3 * j + 1) * d.NS + es], d.rec[size_t(kFastEdgeVec + 3 * j + 2) * d.NS + es]);
        const V3<R> dj = mul(rot, displ[j]) - rest_edge;
        const M3<R> Lj = fast_load_mat(d.rec + size_t(kFastDfDx + 9 * j) * d.NS, d.NS, es);
        force[L1[j]] = force[L1[j]] + mul(Lj, dj);
        force[L0[j]] = force[L0[j]] - mul_t(Lj, dj);
        // what this tetrahedron adds to the matrix of its j-th edge at the next addDForce ([FTC].inl:436-447): rotation^T (L_j rotation) when the
        // edge runs the way the topology's edge does, (L_j rotation)^T rotation otherwise -- computed here, where both operands are in registers
        const M3<R> tmp = mul(Lj, rot);
        const M3<R> add = ((orient >> j) & 1u) ? mul_atb(rot, tmp) : mul_atb(tmp, rot);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) pm[9 * j + 3 * a + b] = add.m[a][b];
    }
    pm[54] = R(0); pm[55] = R(0);
    {
        Quad<R>* dst = reinterpret_cast<Quad<R>*>(d.pmat + es * size_t(kFastPmat));
#pragma unroll
        for (int q = 0; q < kFastPmat / 4; ++q) dst[q] = Quad<R>{pm[4 * q], pm[4 * q + 1], pm[4 * q + 2], pm[4 * q + 3]};
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) C[n] = mul(Rm, force[n]);
}

template <class R, int METHOD>
__global__ void __launch_bounds__(256) fast_tet_kernel(FastDev<R> d, const R* __restrict__ in, NodeEpilogue<R> ep, int max_touched, int max_slots) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ uint16_t s_jds[1024];
    typedef typename SVec<R>::T SV;
    if (ep.cg && ep.cg->done) return;
    SV* s_in = reinterpret_cast<SV*>(smem_raw);
    const size_t off = (sizeof(SV) * size_t(max_touched) + 15) & ~size_t(15);
    R* s_slot = reinterpret_cast<R*>(smem_raw + off);
    const TileDev<R>& t = d.t;
    const int tile = blockIdx.x;
    const uint64_t pol_keep = l2_policy_evict_last();
    tile_phase1<R>(t, tile, in, s_in, s_jds);
    for (int le = threadIdx.x; le < t.tile_e; le += blockDim.x) {
        const size_t es = size_t(tile) * t.tile_e + le;
        const ushort4 ln = d.lnode[es];
        if (ln.x == 0xFFFFu) continue;
        const uint4 sl = d.slot[es];
        const SV pa = s_in[ln.x], pb = s_in[ln.y], pc = s_in[ln.z], pd = s_in[ln.w];
        const V3<R> P[4] = {mk3<R>(pa.x, pa.y, pa.z), mk3<R>(pb.x, pb.y, pb.z), mk3<R>(pc.x, pc.y, pc.z), mk3<R>(pd.x, pd.y, pd.z)};
        V3<R> C[4];
        fast_tet_element<R, METHOD>(d, es, P, C);
        const unsigned s4[4] = {sl.x, sl.y, sl.z, sl.w};
#pragma unroll
        for (int n = 0; n < 4; ++n) tile_scatter<R>(t, s4[n], C[n].x, C[n].y, C[n].z, s_slot, max_slots, pol_keep);
    }
    __syncthreads();
    const double part = tile_phase3<R>(t, tile, ep, s_in, s_slot, max_slots, s_jds);
    if (ep.dot_kind != DOT_NONE) {
        const double tot = block_sum(part, red);
        finish_dot(ep, tot, red, false);
    }
}

// the per-edge matrices, [FTC].inl:414-450: the sum, over the tetrahedra around the edge in ascending index, of the 3x3 blocks fast_tet_kernel left in
// pmat.  inc: (tetrahedron slot << 4) | local edge.
template <class R>
__global__ void fast_edge_assemble_kernel(size_t NSe, const uint32_t* __restrict__ eorder, const uint32_t* __restrict__ inc_off, const uint32_t* __restrict__ inc,
                                          const R* __restrict__ pmat, R* __restrict__ emat) {
    const size_t s = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= NSe) return;
    const uint32_t e = eorder[s];
    if (e == 0xFFFFFFFFu) return;
    R M[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) M[k] = R(0);
    for (uint32_t q = inc_off[e]; q < inc_off[e + 1]; ++q) {
        const uint32_t w = inc[q];
        const R* src = pmat + size_t(w >> 4) * kFastPmat + 9 * (w & 7u);
#pragma unroll
        for (int k = 0; k < 9; ++k) M[k] += src[k];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) emat[size_t(k) * NSe + s] = M[k];
}

// addDForce over the edges, [FTC].inl:455-466.  The epilogue subtracts (sign -1, like the other force fields' addDForce): the corner of edge[1]
// carries -(M deltax), the corner of edge[0] carries +(M^T deltax); a - (-c) and a + c are the same IEEE operation.
template <class R>
__global__ void __launch_bounds__(256) fast_edge_kernel(FastDev<R> d, const R* __restrict__ in, NodeEpilogue<R> ep, int max_touched, int max_slots) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ uint16_t s_jds[1024];
    typedef typename SVec<R>::T SV;
    if (ep.cg && ep.cg->done) return;
    SV* s_in = reinterpret_cast<SV*>(smem_raw);
    const size_t off = (sizeof(SV) * size_t(max_touched) + 15) & ~size_t(15);
    R* s_slot = reinterpret_cast<R*>(smem_raw + off);
    const TileDev<R>& t = d.t;
    const int tile = blockIdx.x;
    const uint64_t pol_keep = l2_policy_evict_last();
    tile_phase1<R>(t, tile, in, s_in, s_jds);
    for (int le = threadIdx.x; le < t.tile_e; le += blockDim.x) {
        const size_t es = size_t(tile) * t.tile_e + le;
        const ushort2 ln = d.elnode[es];
        if (ln.x == 0xFFFFu) continue;
        const uint2 sl = d.eslot[es];
        const SV p0 = s_in[ln.x], p1 = s_in[ln.y];
        const V3<R> deltax = (mk3<R>(p1.x, p1.y, p1.z) - mk3<R>(p0.x, p0.y, p0.z)) * d.k_factor;
        const M3<R> M = fast_load_mat(d.emat, d.NSe, es);
        const V3<R> c1 = mul(M, deltax), c0 = mul_t(M, deltax);
        const bool minus = ep.sign < 0;
        tile_scatter<R>(t, sl.x, minus ? c0.x : -c0.x, minus ? c0.y : -c0.y, minus ? c0.z : -c0.z, s_slot, max_slots, pol_keep);
        tile_scatter<R>(t, sl.y, minus ? -c1.x : c1.x, minus ? -c1.y : c1.y, minus ? -c1.z : c1.z, s_slot, max_slots, pol_keep);
    }
    __syncthreads();
    const double part = tile_phase3<R>(t, tile, ep, s_in, s_slot, max_slots, s_jds);
    if (ep.dot_kind != DOT_NONE) {
        const double tot = block_sum(part, red);
        finish_dot(ep, tot, red, false);
    }
}

// element policy of the fused CG kernel (cg_fused.cuh): the whole CGLinearSolver loop in one persistent launch, A*p over the edges.
// (The kernel's epilogue subtracts the contributions -- it computes q = (m M + b B + k K) p with the stiffness term as `df -= ...` -- hence
// the signs of fast_edge_kernel's `minus` branch.)
template <class R> struct EdgePass {
    typedef FastDev<R> Dev;
    static __device__ __forceinline__ const TileDev<R>& tiles(const Dev& d) { return d.t; }
    struct First {};
    static __device__ __forceinline__ void prefetch(const Dev&, int, First&) {}
    template <int ET, class OnBoundary>
    static __device__ __forceinline__ void elements(const Dev& d, int tile, const typename SVec<R>::T* s_in, R* s_slot, int max_slots, unsigned char*, int arrive_at, OnBoundary on_boundary,
                                                    const First&) {
        typedef typename SVec<R>::T SV;
        const TileDev<R>& t = d.t;
        const uint64_t pol_keep = l2_policy_evict_last();
        for (int le = threadIdx.x; le < t.tile_e; le += ET) {
            if (le - int(threadIdx.x) == arrive_at) on_boundary();
            const size_t es = size_t(tile) * t.tile_e + le;
            const ushort2 ln = d.elnode[es];
            if (ln.x == 0xFFFFu) continue;
            const uint2 sl = d.eslot[es];
            const SV p0 = s_in[ln.x], p1 = s_in[ln.y];
            const V3<R> deltax = (mk3<R>(p1.x, p1.y, p1.z) - mk3<R>(p0.x, p0.y, p0.z)) * d.k_factor;
            const M3<R> M = fast_load_mat(d.emat, d.NSe, es);
            const V3<R> c1 = mul(M, deltax), c0 = mul_t(M, deltax);
            tile_scatter<R>(t, sl.x, c0.x, c0.y, c0.z, s_slot, max_slots, pol_keep);
            tile_scatter<R>(t, sl.y, -c1.x, -c1.y, -c1.z, s_slot, max_slots, pol_keep);
        }
    }
};

// ---- host side ---------------------------------------------------------------------------------------------------------------------------------
template <class R> struct PlanBufs {
    DevBuf<uint32_t> tile_node_off, tile_nodes, tile_shslot, tile_nint, tile_nb, sh_nodes, sh_base;
    DevBuf<uint16_t> tile_val, tile_jds, sh_val;
    DevBuf<Quad<R>> stage;
    int upload(const HostPlan& P, cudaStream_t s) {
        SB_TRY(tile_node_off.upload(P.tile_node_off, s)); SB_TRY(tile_nodes.upload(P.tile_nodes, s)); SB_TRY(tile_shslot.upload(P.tile_shslot, s));
        SB_TRY(tile_nint.upload(P.tile_nint, s)); SB_TRY(tile_nb.upload(P.tile_nb, s)); SB_TRY(tile_val.upload(P.tile_val, s)); SB_TRY(tile_jds.upload(P.tile_jds, s));
        SB_TRY(sh_nodes.upload(P.sh_nodes, s)); SB_TRY(sh_val.upload(P.sh_val, s)); SB_TRY(sh_base.upload(P.sh_base, s));
        SB_TRY(stage.alloc(P.stage_n)); SB_TRY(stage.zero(s));
        return SOFAB200_OK;
    }
    TileDev<R> dev(const HostPlan& P) const {
        TileDev<R> t;
        t.n_nodes = P.n_nodes; t.n_elems = P.n_elems; t.n_tiles = P.n_tiles; t.tile_e = P.tile_e; t.maxval = P.maxval;
        t.tile_node_off = tile_node_off.p; t.tile_nodes = tile_nodes.p; t.tile_shslot = tile_shslot.p; t.tile_nint = tile_nint.p; t.tile_nb = tile_nb.p;
        t.tile_val = tile_val.p; t.tile_jds = tile_jds.p;
        t.n_shared = P.n_shared; t.n_chunks = P.n_chunks; t.sh_nodes = sh_nodes.p; t.sh_val = sh_val.p; t.sh_base = sh_base.p;
        t.stage = stage.p; t.stage_n = P.stage_n;
        return t;
    }
};

template <class R> struct FastFF : sofab200_tetfem {
    HostPlan tplan, eplan;
    PlanBufs<R> tb, eb;
    size_t n_edges = 0, tsmem = 0, esmem = 0;
    DevBuf<ushort4> lnode; DevBuf<uint4> slot;
    DevBuf<ushort2> elnode; DevBuf<uint2> eslot;
    DevBuf<R> rec, rot, emat, pmat;
    DevBuf<unsigned char> orient;
    DevBuf<uint32_t> eorder, inc_off, inc;
    bool update_matrix = true;
    // element-ordered host copies of what init computed (inspection / parity)
    std::vector<R> h_shape, h_dfdx, h_dfdx_diag, h_rest_rot, h_rest_edge, h_orient;
    std::vector<uint32_t> h_edges;
    FastDev<R> dev(bool edges) const {
        FastDev<R> d;
        d.t = edges ? eb.dev(eplan) : tb.dev(tplan);
        d.lnode = lnode.p; d.slot = slot.p; d.rec = rec.p; d.NS = size_t(tplan.n_tiles) * tplan.tile_e; d.rot = rot.p; d.pmat = pmat.p; d.orient = orient.p;
        d.elnode = elnode.p; d.eslot = eslot.p; d.emat = emat.p; d.NSe = size_t(eplan.n_tiles) * eplan.tile_e;
        d.k_factor = R(0);
        return d;
    }
};

// a tile size that cuts `n_elems` into k * sm_count equal tiles and fits the shared-memory budget
static std::string fast_plan(HostPlan& P, size_t& smem, int n_nodes, int n_elems, int npe, const uint32_t* elems, const double* pos, int sm_count, int cap, int tile_elems,
                             size_t sv_bytes, size_t slot_bytes) {
    const size_t limit = 150 * 1024;
    int k_waves = std::max<int>(1, int((size_t(n_elems) + size_t(sm_count) * cap - 1) / (size_t(sm_count) * cap)));
    auto tile_for = [&](int k) { return std::max(32, (int((size_t(n_elems) + size_t(sm_count) * k - 1) / (size_t(sm_count) * k)) + 31) / 32 * 32); };
    int tile_e = tile_elems > 0 ? std::max(32, (tile_elems + 31) / 32 * 32) : tile_for(k_waves);
    for (;;) {
        const std::string err = build_plan(P, n_nodes, n_elems, npe, elems, pos, tile_e, kGatherChunk, kStageFlag);
        smem = ((sv_bytes * size_t(P.max_touched) + 15) & ~size_t(15)) + slot_bytes * size_t(P.max_slots);
        const bool too_big = smem > limit || err.find("use a smaller tile") != std::string::npos;
        if (too_big && tile_elems <= 0 && tile_e > 32) { tile_e = tile_for(++k_waves); continue; }
        if (!err.empty()) return err;
        if (too_big) return "tile does not fit in shared memory; use a smaller tile_elems";
        if (P.maxval > 1023) return "a node with more than 1023 incident elements";
        return "";
    }
}

template <class R> static int fast_create_t(sofab200_ctx* ctx, size_t n_nodes, const void* rest, size_t n_tets, const uint32_t* tets, const sofab200_tetfem_desc* desc, sofab200_tetfem** out) {
    std::unique_ptr<FastFF<R>> ff(new FastFF<R>());
    ff->ctx = ctx; ff->real = sizeof(R) == 4 ? SOFAB200_F32 : SOFAB200_F64; ff->kind = 1; ff->n_nodes = n_nodes; ff->n_tets = n_tets;
    // Data `method` ([FTC].inl:191-203): "polar" / "qr","large" / "polar2" / "none","linear","small"
    int method;
    switch (desc->method) {
    case SOFAB200_TET_POLAR: method = FAST_POLAR; break;
    case SOFAB200_TET_LARGE: method = FAST_QR; break;
    case SOFAB200_TET_SMALL: method = FAST_LINEAR; break;
    case SOFAB200_TET_POLAR2: method = FAST_POLAR2; break;
    default: return fail(SOFAB200_ERR_INVALID, "FastTetrahedralCorotationalForceField: method must be polar, qr (large), polar2 or none (small)");
    }
    ff->method = method;
    const R* x0 = static_cast<const R*>(rest);
    // ---- edges of the topology
    std::vector<uint32_t> edges, eit(6 * n_tets);
    {
        std::map<std::pair<uint32_t, uint32_t>, uint32_t> idx;
        if (desc->n_edges > 0 && desc->edges) {
            edges.assign(desc->edges, desc->edges + 2 * desc->n_edges);
            for (size_t e = 0; e < desc->n_edges; ++e) {
                if (edges[2 * e] >= n_nodes || edges[2 * e + 1] >= n_nodes) return fail(SOFAB200_ERR_INVALID, "edge refers to a node index out of range");
                idx.emplace(std::make_pair(std::min(edges[2 * e], edges[2 * e + 1]), std::max(edges[2 * e], edges[2 * e + 1])), uint32_t(e));
            }
        }
        const bool given = !edges.empty();
        for (size_t i = 0; i < n_tets; ++i)
            for (int j = 0; j < 6; ++j) {
                const uint32_t v1 = tets[4 * i + kFastLh[j][0]], v2 = tets[4 * i + kFastLh[j][1]];
                if (v1 >= n_nodes || v2 >= n_nodes) return fail(SOFAB200_ERR_INVALID, "element refers to a node index out of range");
                const std::pair<uint32_t, uint32_t> key(std::min(v1, v2), std::max(v1, v2));
                auto it = idx.find(key);
                if (it == idx.end()) {
                    if (given) return fail(SOFAB200_ERR_INVALID, "the edge list does not hold every edge of the tetrahedra");
                    it = idx.emplace(key, uint32_t(idx.size())).first;
                    edges.push_back(key.first); edges.push_back(key.second);
                }
                eit[6 * i + j] = it->second;
            }
    }
    const size_t E = edges.size() / 2;
    ff->n_edges = E; ff->h_edges = edges;
    // ---- createTetrahedronRestInformation, [FTC].inl:38-150, in the reference's arithmetic
    ff->h_shape.assign(12 * n_tets, R(0)); ff->h_dfdx.assign(54 * n_tets, R(0)); ff->h_dfdx_diag.assign(36 * n_tets, R(0));
    ff->h_rest_rot.assign(9 * n_tets, R(0)); ff->h_rest_edge.assign(18 * n_tets, R(0)); ff->h_orient.assign(6 * n_tets, R(0));
    for (size_t i = 0; i < n_tets; ++i) {
        const R E_ = R(desc->n_young > i ? desc->young[i] : desc->young[0]), nu = R(desc->n_poisson > i ? desc->poisson[i] : desc->poisson[0]);
        R mu = E_ / (2 * (1 + nu));                                       // toLameParameters<3, Real>, impl/LameParameters.h:57-65
        R lambda = E_ * nu / ((1 + nu) * (1 - (3 - 1) * nu));
        V3<R> point[4];
        for (int j = 0; j < 4; ++j) { const size_t n = tets[4 * i + j]; point[j] = mk3<R>(x0[3 * n], x0[3 * n + 1], x0[3 * n + 2]); }
        const R vol = -(dot3(cross3(point[1] - point[0], point[2] - point[0]), point[3] - point[0]) / R(6));     // -signedVolume, geometry/Tetrahedron.h:73-83
        mu *= std::fabs(vol); lambda *= std::fabs(vol);
        V3<R> sv[4];
        for (int j = 0; j < 4; ++j) {
            V3<R> c = cross3(point[(j + 2) % 4] - point[(j + 1) % 4], point[(j + 3) % 4] - point[(j + 1) % 4]);
            if (j % 2) c = mk3<R>(-c.x, -c.y, -c.z);
            const R den = vol * 6;
            sv[j] = mk3<R>(c.x / den, c.y / den, c.z / den);
            ff->h_shape[12 * i + 3 * j] = sv[j].x; ff->h_shape[12 * i + 3 * j + 1] = sv[j].y; ff->h_shape[12 * i + 3 * j + 2] = sv[j].z;
        }
        auto comp = [](const V3<R>& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); };
        for (int j = 0; j < 4; ++j) {
            const R val = mu * dot3(sv[j], sv[j]);
            R* D = &ff->h_dfdx_diag[36 * i + 9 * j];
            for (int m = 0; m < 3; ++m)
                for (int n = m; n < 3; ++n) {
                    D[3 * m + n] = lambda * comp(sv[j], n) * comp(sv[j], m) + mu * comp(sv[j], n) * comp(sv[j], m);
                    if (m == n) D[3 * m + m] += R(val); else D[3 * n + m] = D[3 * m + n];
                }
        }
        V3<R> rest_edge[6];
        for (int j = 0; j < 6; ++j) {
            const int k = kFastLh[j][0], l = kFastLh[j][1];
            rest_edge[j] = point[l] - point[k];
            ff->h_rest_edge[18 * i + 3 * j] = rest_edge[j].x; ff->h_rest_edge[18 * i + 3 * j + 1] = rest_edge[j].y; ff->h_rest_edge[18 * i + 3 * j + 2] = rest_edge[j].z;
            const R val = mu * dot3(sv[l], sv[k]);
            R* D = &ff->h_dfdx[54 * i + 9 * j];
            for (int m = 0; m < 3; ++m)
                for (int n = 0; n < 3; ++n) {
                    D[3 * m + n] = lambda * comp(sv[k], n) * comp(sv[l], m) + mu * comp(sv[l], n) * comp(sv[k], m);
                    if (m == n) D[3 * m + m] += R(val);
                }
            ff->h_orient[6 * i + j] = tets[4 * i + k] == edges[2 * eit[6 * i + j]] ? R(1) : R(-1);      // updateTopologyInformation, [FTC].inl:249-270
        }
        M3<R> rr;
        std::memset(&rr, 0, sizeof(rr));
        if (method == FAST_QR) fast_qr_rotation(rr, rest_edge[0], rest_edge[1]);
        else if (method == FAST_POLAR2) {
            M3<R> T;
            set_row(T, 0, point[1] - point[0]); set_row(T, 1, point[2] - point[0]); set_row(T, 2, point[3] - point[0]);
            polar_decomposition(T, rr);
        }
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) ff->h_rest_rot[9 * i + 3 * a + b] = rr.m[a][b];
    }
    // ---- the two plans
    std::vector<double> pos(3 * n_nodes);
    for (size_t i = 0; i < 3 * n_nodes; ++i) pos[i] = double(x0[i]);
    typedef typename SVec<R>::T SV;
    std::string err = fast_plan(ff->tplan, ff->tsmem, int(n_nodes), int(n_tets), 4, tets, pos.data(), ctx->sm_count, sizeof(R) == 4 ? 2048 : 1024, desc->tile_elems, sizeof(SV), 3 * sizeof(R));
    if (!err.empty()) return fail(SOFAB200_ERR_INVALID, err);
    err = fast_plan(ff->eplan, ff->esmem, int(n_nodes), int(E), 2, edges.data(), pos.data(), ctx->sm_count, sizeof(R) == 4 ? 4096 : 2048, 0, sizeof(SV), 3 * sizeof(R));
    if (!err.empty()) return fail(SOFAB200_ERR_INVALID, err);
    const HostPlan& TP = ff->tplan; const HostPlan& EP = ff->eplan;
    const size_t NS = size_t(TP.n_tiles) * TP.tile_e, NSe = size_t(EP.n_tiles) * EP.tile_e;
    SB_CHECK(NS < (size_t(1) << 28), "mesh too large for the edge incidence words");
    cudaStream_t s = ctx->stream;
    {
        std::vector<ushort4> ln(NS, make_ushort4(0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF)); std::vector<uint4> sl(NS, make_uint4(0, 0, 0, 0));
        std::vector<R> rec(size_t(kFastRec) * NS, R(0));
        std::vector<uint32_t> slot_of_tet(n_tets, 0);
        for (size_t es = 0; es < NS; ++es) {
            const uint32_t e = TP.order[es];
            if (e == 0xFFFFFFFFu) continue;
            slot_of_tet[e] = uint32_t(es);
            ln[es] = make_ushort4(TP.lnode[4 * es], TP.lnode[4 * es + 1], TP.lnode[4 * es + 2], TP.lnode[4 * es + 3]);
            sl[es] = make_uint4(TP.slot[4 * es], TP.slot[4 * es + 1], TP.slot[4 * es + 2], TP.slot[4 * es + 3]);
            for (int k = 0; k < 18; ++k) rec[size_t(kFastEdgeVec + k) * NS + es] = ff->h_rest_edge[18 * size_t(e) + k];
            for (int k = 0; k < 54; ++k) rec[size_t(kFastDfDx + k) * NS + es] = ff->h_dfdx[54 * size_t(e) + k];
            for (int k = 0; k < 9; ++k) rec[size_t(kFastRestRot + k) * NS + es] = ff->h_rest_rot[9 * size_t(e) + k];
            for (int k = 0; k < 9; ++k) rec[size_t(kFastShape + k) * NS + es] = ff->h_shape[12 * size_t(e) + 3 + k];
        }
        SB_TRY(ff->lnode.upload(ln, s)); SB_TRY(ff->slot.upload(sl, s)); SB_TRY(ff->rec.upload(rec, s));
        SB_TRY(ff->rot.alloc(9 * NS)); SB_TRY(ff->rot.zero(s));
        SB_TRY(ff->pmat.alloc(size_t(kFastPmat) * NS)); SB_TRY(ff->pmat.zero(s));
        {
            std::vector<unsigned char> orient(NS, 0);
            for (size_t es = 0; es < NS; ++es) {
                const uint32_t e = TP.order[es];
                if (e == 0xFFFFFFFFu) continue;
                for (int j = 0; j < 6; ++j) if (ff->h_orient[6 * size_t(e) + j] == R(1)) orient[es] |= (unsigned char)(1u << j);
            }
            SB_TRY(ff->orient.upload(orient, s));
        }
        // edges in tile order + the tetrahedra around each edge, ascending index (the order in which [FTC].inl:425-448 accumulates)
        std::vector<ushort2> eln(NSe, make_ushort2(0xFFFF, 0xFFFF)); std::vector<uint2> esl(NSe, make_uint2(0, 0));
        for (size_t es = 0; es < NSe; ++es) {
            if (EP.order[es] == 0xFFFFFFFFu) continue;
            eln[es] = make_ushort2(EP.lnode[2 * es], EP.lnode[2 * es + 1]);
            esl[es] = make_uint2(EP.slot[2 * es], EP.slot[2 * es + 1]);
        }
        std::vector<uint32_t> inc_off(E + 1, 0), inc(6 * n_tets);
        for (size_t q = 0; q < 6 * n_tets; ++q) inc_off[eit[q] + 1]++;
        for (size_t e = 0; e < E; ++e) inc_off[e + 1] += inc_off[e];
        std::vector<uint32_t> cur(inc_off.begin(), inc_off.end() - 1);
        for (size_t i = 0; i < n_tets; ++i)
            for (int j = 0; j < 6; ++j)
                inc[cur[eit[6 * i + j]]++] = (slot_of_tet[i] << 4) | uint32_t(j);
        SB_TRY(ff->elnode.upload(eln, s)); SB_TRY(ff->eslot.upload(esl, s)); SB_TRY(ff->eorder.upload(EP.order, s));
        SB_TRY(ff->inc_off.upload(inc_off, s)); SB_TRY(ff->inc.upload(inc, s));
        SB_TRY(ff->emat.alloc(9 * NSe)); SB_TRY(ff->emat.zero(s));
        SB_TRY(ff->tb.upload(TP, s)); SB_TRY(ff->eb.upload(EP, s));
        SB_CUDA(cudaStreamSynchronize(s));
    }
    if (getenv("SOFAB200_VERBOSE"))
        fprintf(stderr, "[sofa_b200] fast tet plan: %d tiles x %d tetrahedra (smem %zu B), %zu edges in %d tiles x %d (smem %zu B)\n", TP.n_tiles, TP.tile_e, ff->tsmem, E, EP.n_tiles, EP.tile_e, ff->esmem);
    *out = ff.release();
    return SOFAB200_OK;
}

int fast_create(sofab200_ctx* ctx, int real, size_t n_nodes, const void* rest, size_t n_tets, const uint32_t* tets, const sofab200_tetfem_desc* desc, sofab200_tetfem** out) {
    if (desc->plastic_max_threshold > 0 || desc->compute_von_mises || desc->update_stiffness_matrix || desc->n_local_stiffness > 0 || desc->shared_nodes)
        return fail(SOFAB200_ERR_UNSUPPORTED, "FastTetrahedralCorotationalForceField has no plasticity, von Mises, updateStiffnessMatrix or localStiffnessFactor Data, and is not partitioned over GPUs");
    if (real == SOFAB200_F32) return fast_create_t<float>(ctx, n_nodes, rest, n_tets, tets, desc, out);
    return fast_create_t<double>(ctx, n_nodes, rest, n_tets, tets, desc, out);
}

template <class R, int METHOD> static int fast_launch_tets(FastFF<R>& ff, const FastDev<R>& d, const R* in, const NodeEpilogue<R>& ep) {
    auto kern = fast_tet_kernel<R, METHOD>;
    if (ff.tsmem > 48 * 1024) SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ff.tsmem)));
    ff.ctx->prof_start(2);
    kern<<<ff.tplan.n_tiles, 256, ff.tsmem, ff.ctx->stream>>>(d, in, ep, ff.tplan.max_touched, ff.tplan.max_slots);
    ff.ctx->prof_stop(2);
    ff.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}

template <class R> int fast_run(sofab200_tetfem* base, bool dforce, const R* in, R k_factor, NodeEpilogue<R> ep, bool skip_gather) {
    FastFF<R>& ff = *static_cast<FastFF<R>*>(base);
    SB_CHECK((!ep.mdx_src || ep.mdx_src == in) && (!ep.dot_with || ep.dot_with == in) && (!ep.plane_mode || ep.plane_in == in), "mass / dot / plane operands must be the pass's input vector");
    FastDev<R> d = ff.dev(dforce);
    d.k_factor = k_factor;
    const HostPlan& plan = dforce ? ff.eplan : ff.tplan;
    ep.partial_base = 0;
    ep.partial_total = plan.n_tiles + plan.n_chunks;
    if (!dforce) {
        switch (ff.method) {
        case FAST_POLAR: SB_TRY((fast_launch_tets<R, FAST_POLAR>(ff, d, in, ep))); break;
        case FAST_QR: SB_TRY((fast_launch_tets<R, FAST_QR>(ff, d, in, ep))); break;
        case FAST_POLAR2: SB_TRY((fast_launch_tets<R, FAST_POLAR2>(ff, d, in, ep))); break;
        default: SB_TRY((fast_launch_tets<R, FAST_LINEAR>(ff, d, in, ep))); break;
        }
        ff.update_matrix = true;          // "next time assemble the matrix", [FTC].inl:396
    } else {
        SB_TRY(fast_assemble(ff));
        auto kern = fast_edge_kernel<R>;
        if (ff.esmem > 48 * 1024) SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ff.esmem)));
        ff.ctx->prof_start(0);
        kern<<<ff.eplan.n_tiles, 256, ff.esmem, ff.ctx->stream>>>(d, in, ep, ff.eplan.max_touched, ff.eplan.max_slots);
        ff.ctx->prof_stop(0);
        ff.ctx->launches++;
        SB_CUDA(cudaGetLastError());
    }
    if (skip_gather) return SOFAB200_OK;
    ep.partial_base = plan.n_tiles;
    ff.ctx->prof_start(1);
    gather_shared_kernel<R><<<plan.n_chunks, kGatherChunk, 0, ff.ctx->stream>>>(d.t, ep);
    ff.ctx->prof_stop(1);
    ff.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}
template int fast_run<float>(sofab200_tetfem*, bool, const float*, float, NodeEpilogue<float>, bool);
template int fast_run<double>(sofab200_tetfem*, bool, const double*, double, NodeEpilogue<double>, bool);

template <class R> static int fast_assemble(FastFF<R>& ff) {
    if (!ff.update_matrix) return SOFAB200_OK;
    ff.update_matrix = false;
    const size_t NSe = size_t(ff.eplan.n_tiles) * ff.eplan.tile_e;
    fast_edge_assemble_kernel<R><<<unsigned((NSe + 127) / 128), 128, 0, ff.ctx->stream>>>(NSe, ff.eorder.p, ff.inc_off.p, ff.inc.p, ff.pmat.p, ff.emat.p);
    ff.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}

// the whole CG loop in the fused persistent kernel, A*p over the edges: cached tile state when the CTA's (at most two) tiles fit, else streamed
template <class R, bool CACHED> static int fast_fused_launch(FastFF<R>& ff, FastDev<R> d, FusedCG<R> a, const FusedLayout& L, int grid, bool dry_run, int* info) {
    constexpr int ET = sizeof(R) == 4 ? 512 : 256;
    auto kern = fused_cg_kernel<R, EdgePass<R>, ET, 0, CACHED>;
    cudaFuncAttributes fa;
    SB_CUDA(cudaFuncGetAttributes(&fa, kern));
    int dev_smem_optin = 0;
    SB_CUDA(cudaDeviceGetAttribute(&dev_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ff.ctx->device));
    if (L.total + fa.sharedSizeBytes > size_t(dev_smem_optin)) return kPersistNotEligible;
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(std::max<unsigned>(L.total, 1024))));
    int per_sm = 0;
    SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ET, L.total));
    if (per_sm < 1) return kPersistNotEligible;
    if (info) { info[0] = grid; info[1] = L.tiles_per_cta; info[2] = L.cached; info[3] = int(L.total); info[4] = ET; info[5] = 0; }
    if (dry_run) return SOFAB200_OK;
    SB_TRY(fast_assemble(ff));
    a.lay = L;
    int ded_share = 0;
    void* args[] = {&d, &a, &ded_share};
    ff.ctx->prof_start(4);
    SB_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(ET), args, L.total, ff.ctx->stream));
    ff.ctx->prof_stop(4);
    ff.ctx->launches++;
    return SOFAB200_OK;
}
template <class R> int fast_cg_fused(sofab200_tetfem* base, R k_factor, FusedCG<R> a, size_t sync_capacity, bool dry_run, int* info) {
    FastFF<R>& ff = *static_cast<FastFF<R>*>(base);
    { const char* env = getenv("SOFAB200_FAST_FUSED"); if (env && atoi(env) == 0) return kPersistNotEligible; }
    if (a.ep.sign >= 0) return kPersistNotEligible;      // (EdgePass writes the contributions for a subtracting epilogue)
    FastDev<R> d = ff.dev(true);
    d.k_factor = k_factor;
    const HostPlan& P = ff.eplan;
    const int grid = std::max(1, std::min(ff.ctx->sm_count, P.n_tiles));
    const int tiles_per_cta = (P.n_tiles + grid - 1) / grid;
    const int n_units = P.n_chunks * (kGatherChunk / kUnit);
    const int units_per_cta = (n_units + grid - 1) / grid;
    if (fused_sync_words(grid) > sync_capacity) return fail(SOFAB200_ERR_INVALID, "sync buffer too small for the fused CG kernel");
    const FusedLayout Lc = fused_layout<R>(true, tiles_per_cta, units_per_cta, P.max_touched, P.max_slots, P.max_int, P.max_shtouch, P.maxval);
    if (tiles_per_cta <= 2 && Lc.total + 2048 <= 196 * 1024) {
        const int rc = fast_fused_launch<R, true>(ff, d, a, Lc, grid, dry_run, info);
        if (rc != kPersistNotEligible) return rc;
    }
    const FusedLayout Ls = fused_layout<R>(false, tiles_per_cta, units_per_cta, P.max_touched, P.max_slots, P.max_int, P.max_shtouch, P.maxval);
    return fast_fused_launch<R, false>(ff, d, a, Ls, grid, dry_run, info);
}
template int fast_cg_fused<float>(sofab200_tetfem*, float, FusedCG<float>, size_t, bool, int*);
template int fast_cg_fused<double>(sofab200_tetfem*, double, FusedCG<double>, size_t, bool, int*);

template <class R> TileDev<R> fast_tiledev(sofab200_tetfem* base) { FastFF<R>& ff = *static_cast<FastFF<R>*>(base); return ff.eb.dev(ff.eplan); }
template TileDev<float> fast_tiledev<float>(sofab200_tetfem*);
template TileDev<double> fast_tiledev<double>(sofab200_tetfem*);

#define FAST_BOTH(base, expr)                                                                                       \
    do {                                                                                                            \
        if ((base)->real == SOFAB200_F32) { auto& ff = *static_cast<FastFF<float>*>(base); return (expr); }         \
        auto& ff = *static_cast<FastFF<double>*>(base); return (expr);                                              \
    } while (0)
int fast_partial_count(sofab200_tetfem* base) { FAST_BOTH(base, std::max(ff.tplan.n_tiles + ff.tplan.n_chunks, ff.eplan.n_tiles + ff.eplan.n_chunks)); }
size_t fast_tile_node_count(sofab200_tetfem* base) { FAST_BOTH(base, ff.eplan.tile_nodes.size()); }
size_t fast_shared_slot_count(sofab200_tetfem* base) { FAST_BOTH(base, size_t(ff.eplan.n_chunks) * kGatherChunk); }

template <class R> static int fast_get_t(FastFF<R>& ff, const std::string& what, void* out) {
    auto cp = [&](const std::vector<R>& v) { std::memcpy(out, v.data(), v.size() * sizeof(R)); return SOFAB200_OK; };
    if (what == "shapeVectors") return cp(ff.h_shape);
    if (what == "linearDfDx") return cp(ff.h_dfdx);
    if (what == "linearDfDxDiag") return cp(ff.h_dfdx_diag);
    if (what == "restRotations") return cp(ff.h_rest_rot);
    if (what == "restEdgeVectors") return cp(ff.h_rest_edge);
    if (what == "edgeOrientations") return cp(ff.h_orient);
    if (what == "edges") { std::memcpy(out, ff.h_edges.data(), ff.h_edges.size() * sizeof(uint32_t)); return SOFAB200_OK; }
    if (what == "n_edges") { *static_cast<uint64_t*>(out) = ff.n_edges; return SOFAB200_OK; }
    SB_CUDA(cudaStreamSynchronize(ff.ctx->stream));
    if (what == "rotations") {        // tetraInfo.rotation, element order
        const size_t NS = size_t(ff.tplan.n_tiles) * ff.tplan.tile_e;
        std::vector<R> tmp(9 * NS);
        SB_CUDA(cudaMemcpy(tmp.data(), ff.rot.p, tmp.size() * sizeof(R), cudaMemcpyDeviceToHost));
        R* o = static_cast<R*>(out);
        for (size_t es = 0; es < NS; ++es) { const uint32_t e = ff.tplan.order[es]; if (e != 0xFFFFFFFFu) for (int k = 0; k < 9; ++k) o[9 * size_t(e) + k] = tmp[size_t(k) * NS + es]; }
        return SOFAB200_OK;
    }
    if (what == "edgeInfo") {         // d_edgeInfo as of the last addDForce, edge order
        const size_t NSe = size_t(ff.eplan.n_tiles) * ff.eplan.tile_e;
        std::vector<R> tmp(9 * NSe);
        SB_CUDA(cudaMemcpy(tmp.data(), ff.emat.p, tmp.size() * sizeof(R), cudaMemcpyDeviceToHost));
        R* o = static_cast<R*>(out);
        for (size_t es = 0; es < NSe; ++es) { const uint32_t e = ff.eplan.order[es]; if (e != 0xFFFFFFFFu) for (int k = 0; k < 9; ++k) o[9 * size_t(e) + k] = tmp[size_t(k) * NSe + es]; }
        return SOFAB200_OK;
    }
    return fail(SOFAB200_ERR_INVALID, "unknown array name");
}
int fast_get(sofab200_tetfem* base, const char* what, void* out) { FAST_BOTH(base, fast_get_t(ff, what, out)); }
int fast_stats(const sofab200_tetfem* cbase, uint64_t out[8]) {
    sofab200_tetfem* base = const_cast<sofab200_tetfem*>(cbase);
    auto fill = [&](auto& ff) {
        const HostPlan& P = ff.eplan;      // the plan of the pass the CG loop runs
        out[0] = P.n_tiles; out[1] = P.tile_e; out[2] = P.n_interior; out[3] = P.n_shared; out[4] = P.n_staged_corners; out[5] = ff.esmem; out[6] = P.maxval; out[7] = ff.n_tets;
        return SOFAB200_OK;
    };
    FAST_BOTH(base, fill(ff));
}

}  // namespace sb
