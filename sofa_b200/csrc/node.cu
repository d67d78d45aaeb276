// MechanicalObject vector ops, DiagonalMass / FixedProjectiveConstraint kernels, and the device-resident
// solver node (EulerImplicitSolver + CGLinearSolver<GraphScattered>).
#include <cmath>
#include <cstring>
#include <memory>

#include <nccl.h>

#include "vec_ops.cuh"
#include "cg_fused.cuh"

using namespace sb;

namespace sb {
// implemented in tet_fem.cu / hex_fem.cu
template <class R> int tet_run(sofab200_tetfem* ff, bool dforce, const R* in, R k_factor, NodeEpilogue<R> ep, bool skip_gather);
template <class R> int tet_cg_persistent(sofab200_tetfem* ff, R k_factor, PersistCG<R> a, size_t part_capacity, bool dry_run);
template <class R> int tet_cg_fused(sofab200_tetfem* ff, R k_factor, FusedCG<R> a, size_t sync_capacity, bool dry_run, int* info);
size_t tet_shared_slot_count(sofab200_tetfem* ff);
template <class R> int hex_cg_fused(sofab200_hexfem* ff, R k_factor, FusedCG<R> a, size_t sync_capacity, bool dry_run, int* info);
size_t hex_shared_slot_count(sofab200_hexfem* ff);
size_t tet_tile_node_count(sofab200_tetfem* ff);
const std::vector<uint32_t>& tet_shared_node_table(sofab200_tetfem* ff);
template <class R> TileDev<R> tet_tiledev(sofab200_tetfem* ff);
template <class R> TileDev<R> hex_tiledev(sofab200_hexfem* ff);
int tet_partial_count(sofab200_tetfem* ff);
template <class R> int hex_run(sofab200_hexfem* ff, bool dforce, const R* in, R k_factor, NodeEpilogue<R> ep, bool skip_gather);
template <class R> int hex_cg_persistent(sofab200_hexfem* ff, R k_factor, PersistCG<R> a, size_t sync_capacity, bool dry_run);
size_t hex_tile_node_count(sofab200_hexfem* ff);
int hex_partial_count(sofab200_hexfem* ff);
int tet_real(sofab200_tetfem* ff); size_t tet_nodes(sofab200_tetfem* ff);
bool tet_is_fast(sofab200_tetfem* ff);
int hex_real(sofab200_hexfem* ff); size_t hex_nodes(sofab200_hexfem* ff);
}  // namespace sb

#define LAUNCH(ctx, kernel, grid, block, ...)                              \
    do {                                                                   \
        (ctx)->prof_start(3);                                              \
        kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);        \
        (ctx)->prof_stop(3);                                               \
        (ctx)->launches++;                                                 \
        SB_CUDA(cudaGetLastError());                                       \
    } while (0)

#define SB_NCCL(call)                                                                                   \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess) return ::sb::fail(SOFAB200_ERR_CUDA, std::string(#call) + ": " + ncclGetErrorString(r_)); \
    } while (0)

struct sofab200_comm {
    sofab200_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
};

// ---------------------------------------------------------------------------------------------------
// MechanicalObject::vOp dispatch (MechanicalObject.inl:2075-2203)
// ---------------------------------------------------------------------------------------------------
template <class R> static int vop_impl(sofab200_ctx* ctx, size_t n, R* r, const R* a, const R* b, double k) {
    const size_t n3 = 3 * n;
    if (n3 == 0) return SOFAB200_OK;
    const int g = vec_grid(n3, ctx->sm_count);
    const R kr = R(k);
    if (!a) {
        if (!b) LAUNCH(ctx, (vop_kernel<R, VOP_CLEAR>), g, kVecBlock, n3, r, a, b, kr);
        else if (r == b) LAUNCH(ctx, (vop_kernel<R, VOP_SCALE>), g, kVecBlock, n3, r, a, b, kr);
        else LAUNCH(ctx, (vop_kernel<R, VOP_EQ_BF>), g, kVecBlock, n3, r, a, b, kr);
    } else if (!b) {
        if (r != a) LAUNCH(ctx, (vop_kernel<R, VOP_COPY>), g, kVecBlock, n3, r, a, b, kr);
    } else if (r == a) {
        if (k == 1.0) LAUNCH(ctx, (vop_kernel<R, VOP_PEQ>), g, kVecBlock, n3, r, a, b, kr);
        else LAUNCH(ctx, (vop_kernel<R, VOP_PEQ_BF>), g, kVecBlock, n3, r, a, b, kr);
    } else if (r == b) {
        if (k == 1.0) LAUNCH(ctx, (vop_kernel<R, VOP_PEQ>), g, kVecBlock, n3, r, b, a, kr);  // r += a
        else LAUNCH(ctx, (vop_kernel<R, VOP_AVF>), g, kVecBlock, n3, r, a, b, kr);
    } else {
        if (k == 1.0) LAUNCH(ctx, (vop_kernel<R, VOP_EQ_AB>), g, kVecBlock, n3, r, a, b, kr);
        else LAUNCH(ctx, (vop_kernel<R, VOP_EQ_ABF>), g, kVecBlock, n3, r, a, b, kr);
    }
    return SOFAB200_OK;
}

template <class R> static int vdot_impl(sofab200_ctx* ctx, size_t n, const R* a, const R* b, double* result_host) {
    const size_t n3 = 3 * n;
    int g = vec_grid(n3, ctx->sm_count);
    if (g > 4096) g = 4096;
    LAUNCH(ctx, (vdot_kernel<R>), g, kVecBlock, n3, a, b, ctx->red_partials.p, ctx->red_counter.p, int(DF_STORE), ctx->red_result.p, (CGDev*)nullptr);
    SB_CUDA(cudaMemcpyAsync(result_host, ctx->red_result.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SOFAB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// solver node
// ---------------------------------------------------------------------------------------------------
struct sofab200_node {
    virtual ~sofab200_node() {}
    sofab200_ctx* ctx = nullptr;
    int real = 0;
    size_t n = 0;
    sofab200_solver_params prm;
};

namespace sb {
template <class R> struct Node : sofab200_node {
    sofab200_tetfem* tet = nullptr;
    sofab200_hexfem* hex = nullptr;
    bool has_mass = false, mass_first = true;
    sofab200_meshmass* mesh_mass = nullptr;   // MeshMatrixMass instead of a diagonal mass: its terms run as their own kernels before the element pass (per-node start values)
    bool uniform_mass = false; double um = 0.0;   // UniformMass: every entry of `mass` holds Real(um)
    bool has_plane = false; PlaneDev<R> plane; double plane_stiffness = 0, plane_rayleigh = 0;   // PlaneForceField, last force field of the node
    DevBuf<unsigned char> plane_contacts;
    DevBuf<R> mass;
    DevBuf<unsigned char> fixed;
    bool has_fixed = false;
    DevBuf<R> f, b, dx, p, q, r;
    DevBuf<R> hx, hv;           // device copies of host state for step_host
    DevBuf<CGDev> cg;
    DevBuf<double> partials;
    DevBuf<unsigned> counters;  // [0] boundary kernel, [1] vector kernels
    int n_fem_partials = 0;
    double mF = 0, bF = 0, kF = 0;
    // CUDA graph of one whole EulerImplicit step (about 110 kernels): replayed while (x, v, params) stay the same
    struct StepGraph { cudaGraphExec_t exec = nullptr; R* x = nullptr; R* v = nullptr; sofab200_solver_params prm; uint64_t launches = 0; int seen = 0; } sg, sg_rest;
    cudaStream_t side_stream = nullptr;    // step_host: the v copy runs here while addForce (which only needs x) runs on the main stream
    cudaEvent_t side_event = nullptr;
    // MechanicalObject::externalForce (accumulateForce, MechanicalObject.inl:1356-1375): f starts from it instead of zero
    DevBuf<R> ext; bool has_ext = false;
    // pipelined host coupling (sofab200_node_step_pipelined): the positions of step k travel to the host on a copy stream while step k + 1 runs
    cudaStream_t copy_stream = nullptr, up_stream = nullptr;
    DevBuf<R> ext_stage; cudaEvent_t ev_ext_used = nullptr; bool ext_used_recorded = false;
    cudaEvent_t ev_step[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr}, ev_h2d = nullptr;
    bool copy_pending[2] = {false, false};
    DevBuf<R> xout[2];
    unsigned long long pipe_k = 0;
    bool use_graph = true;
    // a captured step bakes in which mass / halo path the node takes: forget it whenever that changes
    void invalidate_graphs() {
        for (StepGraph* g : {&sg, &sg_rest}) { if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; } g->seen = 0; }
    }
    ~Node() {
        if (sg.exec) cudaGraphExecDestroy(sg.exec);
        if (sg_rest.exec) cudaGraphExecDestroy(sg_rest.exec);
        if (side_event) cudaEventDestroy(side_event);
        if (side_stream) cudaStreamDestroy(side_stream);
        for (int i = 0; i < 2; ++i) { if (ev_step[i]) cudaEventDestroy(ev_step[i]); if (ev_copy[i]) cudaEventDestroy(ev_copy[i]); }
        if (ev_h2d) cudaEventDestroy(ev_h2d);
        if (ev_ext_used) cudaEventDestroy(ev_ext_used);
        if (up_stream) cudaStreamDestroy(up_stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
    }

    // ---- multi-GPU state (sofab200_node_set_distributed) ------------------------------------------------
    struct Halo {
        sofab200_comm* comm = nullptr;
        size_t n_if = 0, n_send = 0;
        int max_sh = 1;
        std::vector<int> nb_rank; std::vector<size_t> nb_count, nb_off;
        DevBuf<uint32_t> if_idx, send_idx;
        DevBuf<int32_t> src;
        DevBuf<R> sendbuf, recvbuf;
        DevBuf<unsigned char> owned;
        DevBuf<double> scal;
        std::vector<uint32_t> if_idx_host;
        std::vector<std::vector<uint32_t>> nb_rows_host;
    } halo;
    // peer-memory mode (sofab200_node_set_peer): the persistent CG kernel talks to the other GPUs itself
    struct Peer {
        bool ready = false;
        PeerDev<R> dev;
        DevBuf<int32_t> sh_if_row;
        DevBuf<int2> if_send;
        DevBuf<int> fail_flag;
        unsigned long long* hcount = nullptr;   // in the mailbox: halo_peer_kernel calls so far
        size_t buf_words = 0;                   // 8-byte words of one inbox buffer (three buffers: CG kernel, halo even, halo odd)
    } peer;
    bool distributed() const { return halo.comm != nullptr; }
    ncclDataType_t nccl_real() const { return sizeof(R) == 4 ? ncclFloat : ncclDouble; }
    // interface rows of q: sum over the sharing ranks, ascending rank order, same bits on every rank
    int halo_sum(R* qv, const CGDev* cgp) {
        if (!distributed() || halo.n_if == 0) return SOFAB200_OK;
        if (peer.ready && !cgp) {
            // over peer memory, one small kernel (the NCCL route below costs three launches and a send/recv group)
            LAUNCH(ctx, (halo_peer_kernel<R>), 1, 1024, peer.dev, halo.n_if, (const uint32_t*)halo.if_idx.p, qv, peer.hcount, peer.buf_words, peer.fail_flag.p, cg.p);
            return SOFAB200_OK;
        }
        LAUNCH(ctx, (halo_pack_kernel<R>), vec_grid(halo.n_send, ctx->sm_count), kVecBlock, halo.n_send, (const uint32_t*)halo.send_idx.p, (const R*)qv, halo.sendbuf.p, cgp);
        SB_NCCL(ncclGroupStart());
        for (size_t k = 0; k < halo.nb_rank.size(); ++k) {
            SB_NCCL(ncclSend(halo.sendbuf.p + 3 * halo.nb_off[k], 3 * halo.nb_count[k], nccl_real(), halo.nb_rank[k], halo.comm->comm, ctx->stream));
            SB_NCCL(ncclRecv(halo.recvbuf.p + 3 * halo.nb_off[k], 3 * halo.nb_count[k], nccl_real(), halo.nb_rank[k], halo.comm->comm, ctx->stream));
        }
        SB_NCCL(ncclGroupEnd());
        ctx->launches += 1;
        LAUNCH(ctx, (halo_sum_kernel<R>), vec_grid(halo.n_if, ctx->sm_count), kVecBlock, halo.n_if, halo.max_sh, (const uint32_t*)halo.if_idx.p, (const int32_t*)halo.src.p,
               (const R*)halo.recvbuf.p, qv, cgp);
        return SOFAB200_OK;
    }
    // dot over owned nodes -> allreduce -> CG bookkeeping, all on the stream
    int dist_dot(const R* a, const R* b, int action) {
        int g = vec_grid(n, ctx->sm_count); if (g > 2048) g = 2048;
        LAUNCH(ctx, (vdot_masked_kernel<R>), g, kVecBlock, n, a, b, (const unsigned char*)halo.owned.p, partials.p, counters.p + 1, halo.scal.p, (const CGDev*)cg.p);
        SB_NCCL(ncclAllReduce(halo.scal.p, halo.scal.p, 1, ncclDouble, ncclSum, halo.comm->comm, ctx->stream));
        ctx->launches += 1;
        LAUNCH(ctx, cg_scalar_kernel, 1, 1, cg.p, (const double*)halo.scal.p, action);
        return SOFAB200_OK;
    }

    int fem_run(bool dforce, const R* in, R k_factor, const NodeEpilogue<R>& ep, bool skip_gather = false) {
        if (tet) return tet_run<R>(tet, dforce, in, k_factor, ep, skip_gather);
        return hex_run<R>(hex, dforce, in, k_factor, ep, skip_gather);
    }
    // ---- fused CG tail (cooperative kernel): gather + den | x,r update + rho | p update
    bool fused_tail = true;
    bool persistent = true;      // SOFAB200_CG_PERSISTENT=0 selects the multi-kernel loop
    DevBuf<unsigned long long> sync_slots;
    DevBuf<typename SVec<R>::T> xt, rt, ps0, ps1, rs;
    DevBuf<R> gstate;
    int tail_grid = 0;
    DevBuf<double> partials_rho;
    int cg_tail(R* x, double m, double bfac, double k) {
        NodeEpilogue<R> ep = make_mbk_ep(q.p, nullptr, p.p, m, bfac, k, false, 1.0, true, DOT_STORE, cg.p);
        TileDev<R> td = tet ? tet_tiledev<R>(tet) : hex_tiledev<R>(hex);
        if (!tail_grid) {
            int per_sm = 0;
            SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_tail_kernel<R>, kTailBlock, 0));
            tail_grid = std::max(1, std::min(per_sm, 4)) * ctx->sm_count;
            SB_TRY(partials_rho.alloc(size_t(tail_grid) + 8));
            if (partials.n < size_t(td.n_tiles + tail_grid + 8)) return fail(SOFAB200_ERR_INVALID, "partials buffer too small");
        }
        size_t n3 = 3 * n;
        R* rp = r.p; R* pp = p.p; const R* qp = q.p; CGDev* cgp = cg.p; double* pd = partials.p; int ntp = td.n_tiles; double* pr = partials_rho.p;
        void* args[] = {&td, &ep, &n3, &x, &rp, &pp, &qp, &cgp, &pd, &ntp, &pr};
        ctx->prof_start(3);
        SB_CUDA(cudaLaunchCooperativeKernel((void*)cg_tail_kernel<R>, dim3(tail_grid), dim3(kTailBlock), args, 0, ctx->stream));
        ctx->prof_stop(3);
        ctx->launches++;
        return SOFAB200_OK;
    }
    // the whole CG loop in one cooperative launch (cg_persist.cuh); pd != null: multi-GPU over peer memory
    int launch_persistent(R* x, const R* bvec, double m, double bfac, double k, const PeerDev<R>* pd) {
        const double kf = k + bfac * prm.ff_rayleigh_stiffness;
        const size_t n3 = 3 * n;
        if (!ps0.p) { SB_TRY(ps0.alloc(n)); SB_TRY(ps1.alloc(n)); SB_TRY(rs.alloc(n)); }
        const size_t n_tile_nodes = tet ? tet_tile_node_count(tet) : hex_tile_node_count(hex);
        if (xt.n < n_tile_nodes) { SB_TRY(xt.alloc(n_tile_nodes)); SB_TRY(rt.alloc(n_tile_nodes)); }
        if (!gstate.p) SB_TRY(gstate.alloc(size_t(9) * ctx->sm_count * 2048));
        if (!sync_slots.p) SB_TRY(sync_slots.alloc(3 * 2048 + 8));
        PersistCG<R> a;
        a.ep = make_mbk_ep(q.p, nullptr, p.p, m, bfac, k, false, 1.0, true, DOT_STORE, cg.p);
        a.x = x; a.r = r.p; a.b = bvec; a.xt = xt.p; a.rt = rt.p; a.gstate = gstate.p; a.p0 = ps0.p; a.p1 = ps1.p; a.rs = rs.p; a.n3 = n3; a.cg = cg.p; a.sync = sync_slots.p;
        if (pd) a.peer = *pd; else std::memset(&a.peer, 0, sizeof(a.peer));
        SB_CUDA(cudaMemsetAsync(sync_slots.p, 0, sync_slots.n * sizeof(unsigned long long), ctx->stream));
        const int rc = tet ? tet_cg_persistent<R>(tet, R(kf), a, sync_slots.n - 8, false) : hex_cg_persistent<R>(hex, R(kf), a, sync_slots.n - 8, false);
        if (rc == SOFAB200_OK) ctx->launches++;
        return rc;
    }
    // second-generation kernel (cg_fused.cuh): one reduction per iteration, any number of tiles per CTA
    bool fused = true;           // SOFAB200_CG_FUSED=0 selects the first-generation persistent kernel
    DevBuf<typename SVec<R>::T> gP, xS, rS, pS, qS;
    DevBuf<R> gQ;
    DevBuf<NodeRec<R>> gNrec;
    DevBuf<GRec<R>> shrec;
    int fused_info[6] = {0, 0, 0, 0, 0, 0};
    int launch_fused(R* x, const R* bvec, double m, double bfac, double k, const PeerDev<R>* pd = nullptr) {
        const double kf = k + bfac * prm.ff_rayleigh_stiffness;
        const size_t n_tile_nodes = tet ? tet_tile_node_count(tet) : hex_tile_node_count(hex), n_slots = tet ? tet_shared_slot_count(tet) : hex_shared_slot_count(hex);
        if (xt.n < n_tile_nodes) { SB_TRY(xt.alloc(n_tile_nodes)); SB_TRY(rt.alloc(n_tile_nodes)); }
        if (gP.n < n_tile_nodes) { SB_TRY(gP.alloc(n_tile_nodes)); SB_TRY(gQ.alloc(3 * n_tile_nodes)); SB_TRY(gNrec.alloc(n_tile_nodes)); }
        if (xS.n < n_slots) { SB_TRY(xS.alloc(n_slots)); SB_TRY(rS.alloc(n_slots)); SB_TRY(pS.alloc(n_slots)); SB_TRY(qS.alloc(n_slots)); SB_TRY(shrec.alloc(n_slots)); }
        if (!sync_slots.p) SB_TRY(sync_slots.alloc(3 * 2048 + 8));
        FusedCG<R> a;
        a.ep = make_mbk_ep(q.p, nullptr, p.p, m, bfac, k, false, 1.0, true, DOT_STORE, cg.p);
        a.x = x; a.r = r.p; a.b = bvec; a.xt = xt.p; a.rt = rt.p; a.gP = gP.p; a.gQ = gQ.p; a.gNrec = gNrec.p;
        a.xS = xS.p; a.rS = rS.p; a.pS = pS.p; a.qS = qS.p; a.shrec = shrec.p; a.n3 = 3 * n; a.cg = cg.p; a.sync = sync_slots.p;
        if (pd) a.peer = *pd; else std::memset(&a.peer, 0, sizeof(a.peer));
        a.n_if_units = pd ? int((halo.n_if + kUnit - 1) / kUnit) : 0;
        SB_CUDA(cudaMemsetAsync(sync_slots.p, 0, sync_slots.n * sizeof(unsigned long long), ctx->stream));
        return tet ? tet_cg_fused<R>(tet, R(kf), a, sync_slots.n, false, fused_info) : hex_cg_fused<R>(hex, R(kf), a, sync_slots.n, false, fused_info);
    }
    NodeEpilogue<R> base_ep() {
        NodeEpilogue<R> ep{};
        ep.mass = mass.p; ep.partials = partials.p; ep.counter = counters.p; ep.cg = nullptr; ep.trace = ctx->trace.p;
        return ep;
    }
    void set_mass_term(NodeEpilogue<R>& ep, int kind, const R* src, double factor) {
        if (!has_mass || mesh_mass) return;
        if (kind == PRE_MDX && factor == 0.0) return;  // Mass::addMBKdx skips a null factor (Mass.inl:96-99)
        if (mass_first) ep.pre_kind = kind; else ep.post_kind = kind;
        ep.mdx_src = src; ep.mass_factor = R(factor); ep.mass_factor_is_one = (factor == 1.0);
        ep.mass_uniform = uniform_mass ? 1 : 0;
        if (uniform_mass) { R m = R(um); if (factor != 1.0) m *= R(factor); ep.um_f = m; }
        ep.gx = R(prm.gravity[0]); ep.gy = R(prm.gravity[1]); ep.gz = R(prm.gravity[2]);
    }
    // mop.computeForce
    int compute_force(R* f_out, const R* x, const R* v = nullptr, bool skip_halo = false) {
        NodeEpilogue<R> ep = base_ep();
        ep.init_src = has_ext ? ext.p : nullptr; ep.sign = +1; ep.out = f_out;     // resetForce + accumulateForce: f = 0 (+ externalForce)
        set_mass_term(ep, PRE_GRAVITY, nullptr, 1.0);
        if (mesh_mass) {   // f = 0; MeshMatrixMass::addForce(f); the element pass then starts every node from that value (same order as the reference's visitor)
            if (has_ext) SB_CUDA(cudaMemcpyAsync(f_out, ext.p, 3 * n * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream));
            else LAUNCH(ctx, (vop_kernel<R, VOP_CLEAR>), vec_grid(3 * n, ctx->sm_count), kVecBlock, 3 * n, f_out, (const R*)nullptr, (const R*)nullptr, R(0));
            SB_TRY(sofab200_meshmass_add_force(mesh_mass, f_out, prm.gravity));
            ep.init_src = f_out;
        }
        if (has_plane) { ep.plane_mode = 1; ep.plane = plane; ep.plane_v = v; ep.plane_contacts = plane_contacts.p; ep.plane_in = x; }
        SB_TRY(fem_run(false, x, R(0), ep));
        return skip_halo ? SOFAB200_OK : halo_sum(f_out, nullptr);     // (skip_halo: the caller keeps every rank's PARTIAL forces, see step_direct)
    }
    // df = init + (m M + b B + k K) d, optionally scaled and projected; dot(out, dot_with) optional
    NodeEpilogue<R> make_mbk_ep(R* out, const R* init, const R* d, double m, double bfac, double k, bool scale, double s, bool project, int dot_kind, CGDev* cgp) {
        NodeEpilogue<R> ep = base_ep();
        ep.init_src = init; ep.sign = -1; ep.out = out;
        const double mf = m - bfac * prm.mass_rayleigh_mass;          // MechanicalParams.h:64
        set_mass_term(ep, PRE_MDX, d, mf);
        ep.has_scale = scale; ep.scale = R(s);
        ep.fixed = (project && has_fixed) ? fixed.p : nullptr;
        ep.dot_kind = dot_kind; ep.dot_with = d; ep.cg = cgp;
        if (mesh_mass) ep.init_src = out;   // add_mbk has put init + m M d there
        if (has_plane) {   // BaseForceField::addMBKdx for the plane: skipped when its kFactor and bFactor are both zero
            const double kfp = k + bfac * plane_rayleigh;
            if (kfp != 0.0 || bfac != 0.0) {
                ep.plane_mode = 2; ep.plane = plane; ep.plane_contacts = plane_contacts.p; ep.plane_in = d;
                ep.plane_fact = R(-double(R(plane_stiffness)) * kfp);
            }
        }
        return ep;
    }
    int add_mbk(R* out, const R* init, const R* d, double m, double bfac, double k, bool scale, double s, bool project, int dot_kind, CGDev* cgp, bool skip_gather = false) {
        NodeEpilogue<R> ep = make_mbk_ep(out, init, d, m, bfac, k, scale, s, project, dot_kind, cgp);
        const double kf = k + bfac * prm.ff_rayleigh_stiffness;       // MechanicalParams.h:62
        if (mesh_mass) {
            // Mass::addMBKdx (Mass.inl:93-105) for the MeshMatrixMass, as its own kernel: out = init (or 0) + M d * mFactor
            if (!mass_first) return fail(SOFAB200_ERR_UNSUPPORTED, "a MeshMatrixMass must precede the force field in the node");
            if (distributed()) return fail(SOFAB200_ERR_UNSUPPORTED, "MeshMatrixMass is not available in a distributed node");
            if (init) { if (init != out) SB_CUDA(cudaMemcpyAsync(out, init, 3 * n * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream)); }
            else LAUNCH(ctx, (vop_kernel<R, VOP_CLEAR>), vec_grid(3 * n, ctx->sm_count), kVecBlock, 3 * n, out, (const R*)nullptr, (const R*)nullptr, R(0));
            const double mf = m - bfac * prm.mass_rayleigh_mass;
            if (mf != 0.0) SB_TRY(sofab200_meshmass_add_mdx(mesh_mass, out, d, mf));
        }
        if (kf != 0.0 || bfac != 0.0) return fem_run(true, d, R(kf), ep, skip_gather);   // BaseForceField::addMBKdx, BaseForceField.cpp:38-47
        if (dot_kind != DOT_NONE) return fail(SOFAB200_ERR_UNSUPPORTED, "system without a stiffness term is not supported in the CG loop");
        LAUNCH(ctx, (node_only_kernel<R>), vec_grid(n, ctx->sm_count), kVecBlock, n, ep);
        return SOFAB200_OK;
    }
    int apply(R* q_out, const R* p_in, double m, double bfac, double k) {
        SB_TRY(add_mbk(q_out, nullptr, p_in, m, bfac, k, false, 1.0, true, DOT_NONE, nullptr));
        return halo_sum(q_out, nullptr);
    }
    // CGLinearSolver::solve, device resident
    int cg_solve(R* x, const R* bvec, double m, double bfac, double k) {
        const size_t n3 = 3 * n;
        const int g = vec_grid(n3, ctx->sm_count);
        if (mesh_mass) persistent = false;   // the persistent kernel keeps p of interior nodes in shared memory: no neighbour access for the edge terms
        CGBegin cb{prm.iterations, prm.tolerance, prm.threshold};
        LAUNCH(ctx, cg_begin_kernel, 1, 1, cg.p, cb, (const int*)(peer.ready ? peer.fail_flag.p : nullptr));
        if (prm.warm_start) {
            SB_TRY(apply(r.p, x, m, bfac, k));                                             // r = A x
            LAUNCH(ctx, (vop_kernel<R, VOP_AVF>), g, kVecBlock, n3, r.p, bvec, (const R*)nullptr, R(-1.0));  // r = b + r*(-1)
        } else {
            LAUNCH(ctx, (vop_kernel<R, VOP_CLEAR>), g, kVecBlock, n3, x, (const R*)nullptr, (const R*)nullptr, R(0));
            SB_CUDA(cudaMemcpyAsync(r.p, bvec, n3 * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        // streaming CG kernels: two CTAs per SM, 16-byte accesses, grid-stride
        const int gd = std::max(1, std::min<int>(2 * ctx->sm_count, int((n3 / 4 + kVecBlock - 1) / kVecBlock)));
        if (distributed()) {
            // same loop, with the interface rows of q exchanged and the three dot products all-reduced over the ranks
            const double kf_d = k + bfac * prm.ff_rayleigh_stiffness;
            if (peer.ready && persistent && tet && (kf_d != 0.0 || bfac != 0.0)) {
                // the loop in ONE persistent kernel per GPU; halo rows and dot products go through peer memory (cg_fused.cuh / cg_persist.cuh)
                const int rc = fused ? launch_fused(x, bvec, m, bfac, k, &peer.dev) : launch_persistent(x, bvec, m, bfac, k, &peer.dev);
                if (rc == SOFAB200_OK) { LAUNCH(ctx, cg_end_kernel, 1, 1, cg.p); return SOFAB200_OK; }
                if (rc != kPersistNotEligible) return rc;
                peer.ready = false;      // (cannot happen after sofab200_node_set_peer's probe; kept for safety)
            }
            SB_TRY(dist_dot(bvec, bvec, DF_CG_NORMB));
            SB_TRY(dist_dot(r.p, r.p, DF_CG_RHO));
            for (unsigned it = 1; it <= prm.iterations; ++it) {
                LAUNCH(ctx, (cg_p_update_kernel<R>), gd, kVecBlock, n3, p.p, (const R*)r.p, (const CGDev*)cg.p);
                SB_TRY(add_mbk(q.p, nullptr, p.p, m, bfac, k, false, 1.0, true, DOT_NONE, cg.p));
                SB_TRY(halo_sum(q.p, cg.p));
                SB_TRY(dist_dot(p.p, q.p, -1));
                LAUNCH(ctx, (cg_xr_update_kernel<R>), gd, kVecBlock, n3, x, r.p, (const R*)p.p, (const R*)q.p, cg.p, partials.p, counters.p + 1, 0);
                SB_TRY(dist_dot(r.p, r.p, DF_CG_RHO));
            }
            LAUNCH(ctx, cg_end_kernel, 1, 1, cg.p);
            return SOFAB200_OK;
        }
        const double kf_chk = k + bfac * prm.ff_rayleigh_stiffness;
        if (persistent && fused && (kf_chk != 0.0 || bfac != 0.0)) {
            const int rc = launch_fused(x, bvec, m, bfac, k);                   // (|b| and the first rho included)
            if (rc == SOFAB200_OK) { LAUNCH(ctx, cg_end_kernel, 1, 1, cg.p); return SOFAB200_OK; }
            if (rc != kPersistNotEligible) return rc;
            fused = false;
        }
        if (persistent && (kf_chk != 0.0 || bfac != 0.0)) {
            const int rc = launch_persistent(x, bvec, m, bfac, k, nullptr);     // (|b| and the first rho included)
            if (rc == SOFAB200_OK) { LAUNCH(ctx, cg_end_kernel, 1, 1, cg.p); return SOFAB200_OK; }
            if (rc != kPersistNotEligible) return rc;
            persistent = false;      // this mesh does not fit: multi-kernel loop from now on
        }
        LAUNCH(ctx, (vdot_kernel<R>), gd, kVecBlock, n3, bvec, bvec, partials.p, counters.p + 1, int(DF_CG_NORMB), (double*)nullptr, cg.p);
        LAUNCH(ctx, (vdot_kernel<R>), gd, kVecBlock, n3, (const R*)r.p, (const R*)r.p, partials.p, counters.p + 1, int(DF_CG_RHO), (double*)nullptr, cg.p);
        if (fused_tail && (kf_chk != 0.0 || bfac != 0.0)) {
            // two launches per iteration: the element pass, then the fused cooperative tail (which also prepares the next p)
            SB_CUDA(cudaMemcpyAsync(p.p, r.p, n3 * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream));   // p = r (first iteration)
            for (unsigned it = 1; it <= prm.iterations; ++it) {
                SB_TRY(add_mbk(q.p, nullptr, p.p, m, bfac, k, false, 1.0, true, DOT_STORE, cg.p, true));
                SB_TRY(cg_tail(x, m, bfac, k));
            }
            LAUNCH(ctx, cg_end_kernel, 1, 1, cg.p);
            return SOFAB200_OK;
        }
        for (unsigned it = 1; it <= prm.iterations; ++it) {
            LAUNCH(ctx, (cg_p_update_kernel<R>), gd, kVecBlock, n3, p.p, (const R*)r.p, (const CGDev*)cg.p);
            SB_TRY(add_mbk(q.p, nullptr, p.p, m, bfac, k, false, 1.0, true, DOT_CG_DEN, cg.p));  // q = A p ; den = p.q
            LAUNCH(ctx, (cg_xr_update_kernel<R>), gd, kVecBlock, n3, x, r.p, (const R*)p.p, (const R*)q.p, cg.p, partials.p, counters.p + 1, 1);
        }
        LAUNCH(ctx, cg_end_kernel, 1, 1, cg.p);
        return SOFAB200_OK;
    }
    // EulerImplicitSolver::solve, replayed from a captured CUDA graph once the same (x, v, params) have been seen twice
    // skip_force: the caller has already run compute_force(f, x) (step_host overlaps it with the copy of v)
    int step(R* x, R* v, bool skip_force = false) {
        if (!use_graph || ctx->profiling || ctx->trace.p) return step_direct(x, v, skip_force);
        StepGraph& sg = skip_force ? this->sg_rest : this->sg;
        const bool same = sg.x == x && sg.v == v && std::memcmp(&sg.prm, &prm, sizeof(prm)) == 0;
        if (!same) {
            if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
            sg.x = x; sg.v = v; sg.prm = prm; sg.seen = 0;
        }
        if (!sg.exec) {
            if (sg.seen++ == 0) return step_direct(x, v, skip_force);   // first time: plain launches (also configures the kernels)
            if (!ctx->capture_stream) SB_CUDA(cudaStreamCreateWithFlags(&ctx->capture_stream, cudaStreamNonBlocking));
            cudaStream_t user = ctx->stream;
            const uint64_t l0 = ctx->launches;
            ctx->stream = ctx->capture_stream;
            cudaError_t e = cudaStreamBeginCapture(ctx->capture_stream, cudaStreamCaptureModeThreadLocal);
            int rc = SOFAB200_OK;
            cudaGraph_t graph = nullptr;
            if (e == cudaSuccess) {
                rc = step_direct(x, v, skip_force);
                e = cudaStreamEndCapture(ctx->capture_stream, &graph);
            }
            ctx->stream = user;
            sg.launches = ctx->launches - l0;
            ctx->launches = l0;
            if (rc != SOFAB200_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess || !graph) return fail(SOFAB200_ERR_CUDA, std::string("stream capture failed: ") + cudaGetErrorString(e));
            e = cudaGraphInstantiate(&sg.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { sg.exec = nullptr; return fail(SOFAB200_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
        }
        SB_CUDA(cudaGraphLaunch(sg.exec, ctx->stream));
        ctx->launches += sg.launches;
        return SOFAB200_OK;
    }
    int step_direct(R* x, R* v, bool skip_force = false) {
        const double h = prm.dt, tr = prm.trapezoidal ? 0.5 : 1.0;
        const bool fo = prm.first_order != 0;
        // Distributed: the right-hand side is linear in f, so every rank starts its partial b from its PARTIAL f (mass term on the owner only)
        // and ONE exchange, of b, completes both -- the interface rows of f itself are left partial (nothing downstream reads them).
        const bool partial_f = distributed() && halo.n_if && !fo && !skip_force;
        if (!skip_force) SB_TRY(compute_force(f.p, x, v, partial_f));
        if (!fo) {
            // b = (f + (-rM M + (h tr + rK) K) v) * h, projected          EulerImplicitSolver.cpp:147-162
            const R* finit = f.p;
            if (distributed() && halo.n_if && !partial_f) {   // f was completed over the ranks: its interface rows enter the distributed sum on their owner only
                LAUNCH(ctx, (mask_rows_kernel<R>), vec_grid(n, ctx->sm_count), kVecBlock, n, (const unsigned char*)halo.owned.p, (const R*)f.p, p.p);
                finit = p.p;
            }
            SB_TRY(add_mbk(b.p, finit, v, -prm.rayleigh_mass, 0.0, h * tr + prm.rayleigh_stiffness, true, h, true, DOT_NONE, nullptr));
            SB_TRY(halo_sum(b.p, nullptr));
        } else {
            NodeEpilogue<R> ep = base_ep();
            ep.init_src = f.p; ep.out = b.p; ep.sign = -1; ep.fixed = has_fixed ? fixed.p : nullptr;
            LAUNCH(ctx, (node_only_kernel<R>), vec_grid(n, ctx->sm_count), kVecBlock, n, ep);
        }
        mF = fo ? 1 : 1 + tr * h * prm.rayleigh_mass;
        bF = fo ? 0 : -tr * h;
        kF = fo ? -h * tr : -tr * h * (tr * h + prm.rayleigh_stiffness);
        SB_TRY(cg_solve(dx.p, b.p, mF, bF, kF));
        const size_t n3 = 3 * n;
        const int g = vec_grid(n3, ctx->sm_count);
        if (fo) {
            SB_CUDA(cudaMemcpyAsync(v, dx.p, n3 * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream));   // newVel.eq(x)
            LAUNCH(ctx, (vop_kernel<R, VOP_PEQ_BF>), g, kVecBlock, n3, x, (const R*)x, (const R*)v, R(h));  // newPos.eq(pos,newVel,h)
        } else {
            LAUNCH(ctx, (integrate_kernel<R>), g, kVecBlock, n3, v, x, (const R*)dx.p, R(1), 1, R(h));
        }
        if (prm.vdamping != 0.0) LAUNCH(ctx, (vop_kernel<R, VOP_SCALE>), g, kVecBlock, n3, v, (const R*)nullptr, (const R*)v, R(std::exp(-h * prm.vdamping)));
        return SOFAB200_OK;
    }
};

template <class R> static PlaneDev<R> plane_dev(const sofab200_plane_desc* p);
template <class R> static int node_create(sofab200_ctx* ctx, size_t n, const sofab200_node_desc* d, sofab200_node** out) {
    std::unique_ptr<Node<R>> nd(new Node<R>());
    nd->ctx = ctx; nd->real = sizeof(R) == 4 ? SOFAB200_F32 : SOFAB200_F64; nd->n = n;
    nd->tet = d->tetfem; nd->hex = d->hexfem; nd->mass_first = d->mass_first != 0;
    if (const char* env = getenv("SOFAB200_GRAPH")) nd->use_graph = atoi(env) != 0;
    if (const char* env = getenv("SOFAB200_FUSED_TAIL")) nd->fused_tail = atoi(env) != 0;
    if (const char* env = getenv("SOFAB200_CG_PERSISTENT")) nd->persistent = atoi(env) != 0;
    if (const char* env = getenv("SOFAB200_CG_FUSED")) nd->fused = atoi(env) != 0;
    std::memset(&nd->prm, 0, sizeof(nd->prm));
    nd->prm.gravity[1] = -9.81; nd->prm.dt = 0.01; nd->prm.iterations = 25; nd->prm.tolerance = 1e-5; nd->prm.threshold = 1e-5;
    cudaStream_t s = ctx->stream;
    if (d->plane) {
        nd->has_plane = true; nd->plane = plane_dev<R>(d->plane); nd->plane_stiffness = d->plane->stiffness; nd->plane_rayleigh = d->plane_rayleigh_stiffness;
        SB_TRY(nd->plane_contacts.alloc(n)); SB_TRY(nd->plane_contacts.zero(s));
    }
    if (d->uniform_mass) {
        nd->has_mass = true; nd->uniform_mass = true; nd->um = d->uniform_vertex_mass;
        std::vector<R> um(n, R(d->uniform_vertex_mass));
        SB_TRY(nd->mass.upload(um, s));
        SB_CUDA(cudaStreamSynchronize(s));
    } else if (d->vertex_mass_host) {
        nd->has_mass = true;
        SB_TRY(nd->mass.alloc(n));
        SB_CUDA(cudaMemcpyAsync(nd->mass.p, d->vertex_mass_host, n * sizeof(R), cudaMemcpyHostToDevice, s));
    }
    if (d->fix_all || d->n_fixed > 0) {
        std::vector<unsigned char> mask(n, d->fix_all ? 1 : 0);
        for (size_t i = 0; i < d->n_fixed; ++i) { SB_CHECK(d->fixed_host[i] < n, "fixed index out of range"); mask[d->fixed_host[i]] = 1; }
        nd->has_fixed = true;
        SB_TRY(nd->fixed.upload(mask, s));
    }
    for (DevBuf<R>* v : {&nd->f, &nd->b, &nd->dx, &nd->p, &nd->q, &nd->r}) { SB_TRY(v->alloc(3 * n)); SB_TRY(v->zero(s)); }
    SB_TRY(nd->cg.alloc(1)); SB_TRY(nd->cg.zero(s));
    nd->n_fem_partials = d->tetfem ? tet_partial_count(d->tetfem) : hex_partial_count(d->hexfem);
    SB_TRY(nd->partials.alloc(size_t(std::max(nd->n_fem_partials, 4096)) + 16)); SB_TRY(nd->partials.zero(s));
    SB_TRY(nd->counters.alloc(4)); SB_TRY(nd->counters.zero(s));
    SB_CUDA(cudaStreamSynchronize(s));
    *out = nd.release();
    return SOFAB200_OK;
}
}  // namespace sb

#define NODE_DISPATCH(node, EXPR_F, EXPR_D) ((node)->real == SOFAB200_F32 ? (EXPR_F) : (EXPR_D))
#define NF(node) (static_cast<Node<float>*>(node))
#define ND(node) (static_cast<Node<double>*>(node))

extern "C" {

int sofab200_mo_vop(sofab200_ctx* ctx, sofab200_real real, size_t n, void* r_dev, const void* a_dev, const void* b_dev, double k) {
    SB_CHECK(ctx && r_dev, "null argument");
    if (real == SOFAB200_F32) return vop_impl<float>(ctx, n, (float*)r_dev, (const float*)a_dev, (const float*)b_dev, k);
    return vop_impl<double>(ctx, n, (double*)r_dev, (const double*)a_dev, (const double*)b_dev, k);
}
int sofab200_mo_vdot(sofab200_ctx* ctx, sofab200_real real, size_t n, const void* a_dev, const void* b_dev, double* result_host) {
    SB_CHECK(ctx && a_dev && b_dev && result_host, "null argument");
    if (n == 0) { *result_host = 0.0; return SOFAB200_OK; }
    if (real == SOFAB200_F32) return vdot_impl<float>(ctx, n, (const float*)a_dev, (const float*)b_dev, result_host);
    return vdot_impl<double>(ctx, n, (const double*)a_dev, (const double*)b_dev, result_host);
}
int sofab200_mo_vdot_dev(sofab200_ctx* ctx, sofab200_real real, size_t n, const void* a_dev, const void* b_dev, const unsigned char* mask_dev, double* result_dev) {
    SB_CHECK(ctx && a_dev && b_dev && result_dev, "null argument");
    int g = vec_grid(n, ctx->sm_count);
    if (g > 2048) g = 2048;
    if (real == SOFAB200_F32) LAUNCH(ctx, (vdot_masked_kernel<float>), g, kVecBlock, n, (const float*)a_dev, (const float*)b_dev, mask_dev, ctx->red_partials.p, ctx->red_counter.p, result_dev, (const CGDev*)nullptr);
    else LAUNCH(ctx, (vdot_masked_kernel<double>), g, kVecBlock, n, (const double*)a_dev, (const double*)b_dev, mask_dev, ctx->red_partials.p, ctx->red_counter.p, result_dev, (const CGDev*)nullptr);
    return SOFAB200_OK;
}
int sofab200_mo_vmultiop_integrate(sofab200_ctx* ctx, sofab200_real real, size_t n, void* v_dev, void* x_dev, const void* a_dev, double f_v_a, double f_x_v) {
    SB_CHECK(ctx && v_dev && x_dev && a_dev, "null argument");
    const size_t n3 = 3 * n;
    if (n3 == 0) return SOFAB200_OK;
    const int g = vec_grid(n3, ctx->sm_count);
    if (real == SOFAB200_F32) LAUNCH(ctx, (integrate_kernel<float>), g, kVecBlock, n3, (float*)v_dev, (float*)x_dev, (const float*)a_dev, float(f_v_a), int(float(f_v_a) == 1.0f), float(f_x_v));
    else LAUNCH(ctx, (integrate_kernel<double>), g, kVecBlock, n3, (double*)v_dev, (double*)x_dev, (const double*)a_dev, f_v_a, int(f_v_a == 1.0), f_x_v);
    return SOFAB200_OK;
}
int sofab200_mo_accumulate_force(sofab200_ctx* ctx, sofab200_real real, size_t n, void* f_dev, const void* ext_dev) {
    SB_CHECK(ctx && f_dev && ext_dev, "null argument");
    if (n == 0) return SOFAB200_OK;
    const int g = vec_grid(n, ctx->sm_count);
    if (real == SOFAB200_F32) LAUNCH(ctx, (accumulate_force_kernel<float>), g, kVecBlock, n, (float*)f_dev, (const float*)ext_dev);
    else LAUNCH(ctx, (accumulate_force_kernel<double>), g, kVecBlock, n, (double*)f_dev, (const double*)ext_dev);
    return SOFAB200_OK;
}
int sofab200_mass_add_mdx(sofab200_ctx* ctx, sofab200_real real, size_t n, void* res_dev, const void* dx_dev, const void* m_dev, double factor) {
    SB_CHECK(ctx && res_dev && dx_dev && m_dev, "null argument");
    if (n == 0) return SOFAB200_OK;
    const int g = vec_grid(3 * n, ctx->sm_count);
    if (real == SOFAB200_F32) LAUNCH(ctx, (mass_mdx_kernel<float>), g, kVecBlock, n, (float*)res_dev, (const float*)dx_dev, (const float*)m_dev, float(factor), int(factor == 1.0));
    else LAUNCH(ctx, (mass_mdx_kernel<double>), g, kVecBlock, n, (double*)res_dev, (const double*)dx_dev, (const double*)m_dev, factor, int(factor == 1.0));
    return SOFAB200_OK;
}
}  // extern "C"
namespace sb {
template <class R> static PlaneDev<R> plane_dev(const sofab200_plane_desc* p) {
    // setPlane, PlaneForceField.inl:139-145: n = |normal| (sqrt of norm2 accumulated x, y, z); planeNormal = normal / n; planeD = d / n
    const R a = R(p->normal[0]), b = R(p->normal[1]), c = R(p->normal[2]);
    R n2 = a * a; n2 += b * b; n2 += c * c;
    const R nn = R(std::sqrt(n2));
    PlaneDev<R> P;
    P.nx = a / nn; P.ny = b / nn; P.nz = c / nn; P.d = R(p->d) / nn;
    P.stiff = R(p->stiffness); P.damp = R(p->damping);
    R lim = R(p->max_force); lim *= lim; P.limit2 = lim;
    P.bilateral = p->bilateral;
    return P;
}
}  // namespace sb
extern "C" {
int sofab200_plane_add_force(sofab200_ctx* ctx, sofab200_real real, size_t n, void* f_dev, const void* x_dev, const void* v_dev, const sofab200_plane_desc* plane,
                             unsigned char* contacts_dev) {
    SB_CHECK(ctx && f_dev && x_dev && v_dev && plane && contacts_dev, "null argument");
    if (n == 0) return SOFAB200_OK;
    const int g = vec_grid(n, ctx->sm_count);
    if (real == SOFAB200_F32) LAUNCH(ctx, (plane_add_force_kernel<float>), g, kVecBlock, n, plane_dev<float>(plane), (float*)f_dev, (const float*)x_dev, (const float*)v_dev, contacts_dev);
    else LAUNCH(ctx, (plane_add_force_kernel<double>), g, kVecBlock, n, plane_dev<double>(plane), (double*)f_dev, (const double*)x_dev, (const double*)v_dev, contacts_dev);
    return SOFAB200_OK;
}
int sofab200_plane_add_dforce(sofab200_ctx* ctx, sofab200_real real, size_t n, void* df_dev, const void* dx_dev, const sofab200_plane_desc* plane,
                              const unsigned char* contacts_dev, double k_factor) {
    SB_CHECK(ctx && df_dev && dx_dev && plane && contacts_dev, "null argument");
    if (n == 0) return SOFAB200_OK;
    const int g = vec_grid(n, ctx->sm_count);
    if (real == SOFAB200_F32) {
        const PlaneDev<float> P = plane_dev<float>(plane);
        LAUNCH(ctx, (plane_add_dforce_kernel<float>), g, kVecBlock, n, P, float(-double(float(plane->stiffness)) * k_factor), (float*)df_dev, (const float*)dx_dev, contacts_dev);
    } else {
        const PlaneDev<double> P = plane_dev<double>(plane);
        LAUNCH(ctx, (plane_add_dforce_kernel<double>), g, kVecBlock, n, P, -plane->stiffness * k_factor, (double*)df_dev, (const double*)dx_dev, contacts_dev);
    }
    return SOFAB200_OK;
}
int sofab200_uniform_mass_add_mdx(sofab200_ctx* ctx, sofab200_real real, size_t n, void* res_dev, const void* dx_dev, double vertex_mass, double factor) {
    SB_CHECK(ctx && res_dev && dx_dev, "null argument");
    if (n == 0) return SOFAB200_OK;
    const int g = vec_grid(3 * n, ctx->sm_count);
    // res[i] += dx[i] * m, m = vertexMass (*= Real(factor) if factor != 1): the vOp_v_inc_bf kernel computes exactly r[i] += b[i]*k
    if (real == SOFAB200_F32) { float m = float(vertex_mass); if (factor != 1.0) m *= float(factor); LAUNCH(ctx, (vop_kernel<float, VOP_PEQ_BF>), g, kVecBlock, 3 * n, (float*)res_dev, (const float*)nullptr, (const float*)dx_dev, m); }
    else { double m = vertex_mass; if (factor != 1.0) m *= factor; LAUNCH(ctx, (vop_kernel<double, VOP_PEQ_BF>), g, kVecBlock, 3 * n, (double*)res_dev, (const double*)nullptr, (const double*)dx_dev, m); }
    return SOFAB200_OK;
}
int sofab200_uniform_mass_add_force(sofab200_ctx* ctx, sofab200_real real, size_t n, void* f_dev, double vertex_mass, const double gravity[3]) {
    SB_CHECK(ctx && f_dev && gravity, "null argument");
    if (n == 0) return SOFAB200_OK;
    const int g = vec_grid(3 * n, ctx->sm_count);
    if (real == SOFAB200_F32) {
        const float m = float(vertex_mass);
        LAUNCH(ctx, (add_const3_kernel<float>), g, kVecBlock, n, (float*)f_dev, float(gravity[0]) * m, float(gravity[1]) * m, float(gravity[2]) * m);
    } else LAUNCH(ctx, (add_const3_kernel<double>), g, kVecBlock, n, (double*)f_dev, gravity[0] * vertex_mass, gravity[1] * vertex_mass, gravity[2] * vertex_mass);
    return SOFAB200_OK;
}
int sofab200_mass_add_force(sofab200_ctx* ctx, sofab200_real real, size_t n, void* f_dev, const void* m_dev, const double gravity[3]) {
    SB_CHECK(ctx && f_dev && m_dev && gravity, "null argument");
    if (n == 0) return SOFAB200_OK;
    const int g = vec_grid(3 * n, ctx->sm_count);
    if (real == SOFAB200_F32) LAUNCH(ctx, (mass_gravity_kernel<float>), g, kVecBlock, n, (float*)f_dev, (const float*)m_dev, float(gravity[0]), float(gravity[1]), float(gravity[2]));
    else LAUNCH(ctx, (mass_gravity_kernel<double>), g, kVecBlock, n, (double*)f_dev, (const double*)m_dev, gravity[0], gravity[1], gravity[2]);
    return SOFAB200_OK;
}
int sofab200_mass_acc_from_f(sofab200_ctx* ctx, sofab200_real real, size_t n, void* a_dev, const void* f_dev, const void* m_dev) {
    SB_CHECK(ctx && a_dev && f_dev && m_dev, "null argument");
    if (n == 0) return SOFAB200_OK;
    const int g = vec_grid(3 * n, ctx->sm_count);
    if (real == SOFAB200_F32) LAUNCH(ctx, (mass_acc_kernel<float>), g, kVecBlock, n, (float*)a_dev, (const float*)f_dev, (const float*)m_dev);
    else LAUNCH(ctx, (mass_acc_kernel<double>), g, kVecBlock, n, (double*)a_dev, (const double*)f_dev, (const double*)m_dev);
    return SOFAB200_OK;
}
int sofab200_fixed_project_response(sofab200_ctx* ctx, sofab200_real real, size_t n, void* res_dev, size_t n_idx, const uint32_t* idx_dev, int fix_all) {
    SB_CHECK(ctx && res_dev, "null argument");
    if (fix_all) return sofab200_mo_vop(ctx, real, n, res_dev, nullptr, nullptr, 0.0);
    if (n_idx == 0) return SOFAB200_OK;
    SB_CHECK(idx_dev != nullptr, "indices are null");
    const int g = vec_grid(n_idx, ctx->sm_count);
    if (real == SOFAB200_F32) LAUNCH(ctx, (fixed_project_kernel<float>), g, kVecBlock, n_idx, idx_dev, (float*)res_dev);
    else LAUNCH(ctx, (fixed_project_kernel<double>), g, kVecBlock, n_idx, idx_dev, (double*)res_dev);
    return SOFAB200_OK;
}

int sofab200_comm_get_unique_id(void* out_bytes) {
    SB_CHECK(out_bytes != nullptr, "null argument");
    static_assert(sizeof(ncclUniqueId) <= SOFAB200_UNIQUE_ID_BYTES, "unique id does not fit");
    ncclUniqueId id;
    SB_NCCL(ncclGetUniqueId(&id));
    std::memset(out_bytes, 0, SOFAB200_UNIQUE_ID_BYTES);
    std::memcpy(out_bytes, &id, sizeof(id));
    return SOFAB200_OK;
}
int sofab200_comm_create(sofab200_ctx* ctx, int world, int rank, const void* unique_id_bytes, sofab200_comm** out) {
    SB_CHECK(ctx && unique_id_bytes && out && world >= 1 && rank >= 0 && rank < world, "bad argument");
    SB_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, unique_id_bytes, sizeof(id));
    std::unique_ptr<sofab200_comm> c(new sofab200_comm());
    c->ctx = ctx; c->world = world; c->rank = rank;
    SB_NCCL(ncclCommInitRank(&c->comm, world, id, rank));
    *out = c.release();
    return SOFAB200_OK;
}
int sofab200_comm_destroy(sofab200_comm* comm) {
    if (!comm) return SOFAB200_OK;
    if (comm->comm) ncclCommDestroy(comm->comm);
    delete comm;
    return SOFAB200_OK;
}
}  // extern "C"
namespace sb {
template <class R> static int node_set_distributed(Node<R>* nd, sofab200_comm* comm, const sofab200_halo_desc* h) {
    cudaStream_t s = nd->ctx->stream;
    // every rank applies the per-node epilogue to its copy of an interface node and the copies are then summed over the sharers: only the
    // DiagonalMass term is masked by ownership (the caller zeroes vertexMass on the non-owned copies).  PlaneForceField, UniformMass and
    // MeshMatrixMass terms would be counted once per sharing rank.
    if (nd->tet && tet_is_fast(nd->tet)) return fail(SOFAB200_ERR_UNSUPPORTED, "FastTetrahedralCorotationalForceField is not partitioned over GPUs (its per-edge matrices sum over tetrahedra of several partitions)");
    if (nd->has_plane) return fail(SOFAB200_ERR_UNSUPPORTED, "a node with a PlaneForceField cannot be distributed (its term would be counted once per sharing rank on interface nodes)");
    if (nd->uniform_mass) return fail(SOFAB200_ERR_UNSUPPORTED, "a node with a UniformMass cannot be distributed; use a DiagonalMass whose vertexMass is zero on non-owned interface nodes");
    if (nd->mesh_mass) return fail(SOFAB200_ERR_UNSUPPORTED, "MeshMatrixMass is not available in a distributed node");
    auto& H = nd->halo;
    H.n_if = h->n_interface; H.max_sh = std::max(1, h->max_sharers);
    std::vector<unsigned char> owned(h->owned, h->owned + nd->n);
    SB_TRY(H.owned.upload(owned, s));
    std::vector<uint32_t> if_idx(h->interface, h->interface + h->n_interface), send_idx;
    std::vector<int32_t> src(h->n_interface * size_t(H.max_sh), -2);
    for (size_t i = 0; i < h->n_interface; ++i) { SB_CHECK(h->my_slot[i] >= 0 && h->my_slot[i] < H.max_sh, "my_slot out of range"); src[i * H.max_sh + h->my_slot[i]] = -1; }
    H.nb_rank.clear(); H.nb_count.clear(); H.nb_off.clear();
    size_t off = 0;
    for (int k = 0; k < h->n_neighbours; ++k) {
        H.nb_rank.push_back(h->nb_rank[k]); H.nb_count.push_back(h->nb_count[k]); H.nb_off.push_back(off);
        for (size_t i = 0; i < h->nb_count[k]; ++i) {
            const uint32_t row = h->nb_rows[k][i];
            SB_CHECK(row < h->n_interface && h->nb_slot[k][i] >= 0 && h->nb_slot[k][i] < H.max_sh, "halo plan out of range");
            send_idx.push_back(if_idx[row]);
            src[size_t(row) * H.max_sh + h->nb_slot[k][i]] = int32_t(off + i);
        }
        off += h->nb_count[k];
    }
    H.n_send = off;
    H.if_idx_host = if_idx;
    H.nb_rows_host.clear();
    for (int k = 0; k < h->n_neighbours; ++k) H.nb_rows_host.emplace_back(h->nb_rows[k], h->nb_rows[k] + h->nb_count[k]);
    nd->peer.ready = false;
    SB_TRY(H.if_idx.upload(if_idx, s)); SB_TRY(H.send_idx.upload(send_idx, s)); SB_TRY(H.src.upload(src, s));
    SB_TRY(H.sendbuf.alloc(3 * std::max<size_t>(off, 1))); SB_TRY(H.recvbuf.alloc(3 * std::max<size_t>(off, 1)));
    SB_TRY(H.scal.alloc(2)); SB_TRY(H.scal.zero(s));
    SB_CUDA(cudaStreamSynchronize(s));
    H.comm = comm;
    nd->invalidate_graphs();
    return SOFAB200_OK;
}
// mailbox layout (bytes): 64 all-reduce slots [2][kMaxPeers][2 words] (first-generation kernel) | 320 epoch u64 | 328 halo-call counter u64 |
// 1024 all-reduce slots of the fused kernel [2][kMaxPeers][8 words] | 2048 three inbox buffers of inbox_rows rows each (the same
// inbox_rows on every rank, so that the offsets in a peer's mailbox are known)
constexpr size_t kMailboxFlags = 0, kMailboxAr = 64, kMailboxEpoch = 320, kMailboxHcount = 328, kMailboxAr4 = 1024, kMailboxInbox = 2048;
template <class R> static size_t inbox_buf_words(size_t rows) { return (size_t(InboxWords<R>::N) * std::max<size_t>(rows, 1) + 31) & ~size_t(31); }
template <class R> static size_t node_peer_bytes(const Node<R>* nd, size_t rows) {
    return kMailboxInbox + 3 * inbox_buf_words<R>(std::max(rows, nd->halo.n_send)) * sizeof(unsigned long long) + 256;
}
template <class R> static int node_set_peer(Node<R>* nd, const sofab200_peer_desc* d) {
    if (!d->peer_base) { nd->peer.ready = false; nd->invalidate_graphs(); return SOFAB200_OK; }
    SB_CHECK(nd->distributed(), "sofab200_node_set_distributed must come first");
    SB_CHECK(nd->tet != nullptr, "peer mode is implemented for the tetrahedral force field");
    if (nd->fused) {
        FusedCG<R> probe; std::memset(&probe, 0, sizeof(probe));
        const int rc = tet_cg_fused<R>(nd->tet, R(1), probe, size_t(3) * 2048 + 8, true, nullptr);
        if (rc == kPersistNotEligible) return fail(SOFAB200_ERR_UNSUPPORTED, "this partition does not fit the fused CG kernel");
        if (rc != SOFAB200_OK) return rc;
    } else {
        PersistCG<R> probe; std::memset(&probe, 0, sizeof(probe));
        const int rc = tet_cg_persistent<R>(nd->tet, R(1), probe, size_t(3) * 2048, true);
        if (rc == kPersistNotEligible) return fail(SOFAB200_ERR_UNSUPPORTED, "this partition does not fit the persistent CG kernel (more than two tiles per SM)");
        if (rc != SOFAB200_OK) return rc;
    }
    SB_CHECK(d->world >= 1 && d->world <= kMaxPeers && d->rank >= 0 && d->rank < d->world, "rank / world out of range (world <= 8)");
    auto& H = nd->halo;
    SB_CHECK(int(H.nb_rank.size()) <= kMaxPeers, "too many neighbours");
    cudaStream_t s = nd->ctx->stream;
    PeerDev<R>& P = nd->peer.dev;
    std::memset(&P, 0, sizeof(P));
    P.rank = d->rank; P.world = d->world; P.n_nb = int(H.nb_rank.size()); P.max_sh = H.max_sh;
    unsigned char* mine = static_cast<unsigned char*>(d->peer_base[d->rank]);
    P.ar = reinterpret_cast<ARSlot*>(mine + kMailboxAr);
    P.ar4 = reinterpret_cast<unsigned long long*>(mine + kMailboxAr4);
    P.epoch = reinterpret_cast<unsigned long long*>(mine + kMailboxEpoch);
    nd->peer.hcount = reinterpret_cast<unsigned long long*>(mine + kMailboxHcount);
    SB_CHECK(d->inbox_rows >= H.n_send, "inbox_rows must be the largest number of received rows over all ranks");
    nd->peer.buf_words = inbox_buf_words<R>(d->inbox_rows);
    SB_TRY(nd->peer.fail_flag.alloc(1)); SB_TRY(nd->peer.fail_flag.zero(s));
    P.inbox = reinterpret_cast<unsigned long long*>(mine + kMailboxInbox);
    for (int r = 0; r < d->world; ++r) { SB_CHECK(d->peer_base[r] != nullptr, "peer_base entry is null"); P.peer_ar[r] = reinterpret_cast<ARSlot*>(static_cast<unsigned char*>(d->peer_base[r]) + kMailboxAr);
        P.peer_ar4[r] = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(d->peer_base[r]) + kMailboxAr4); }
    for (int k = 0; k < P.n_nb; ++k) {
        const int r = H.nb_rank[k];
        SB_CHECK(r >= 0 && r < d->world && r != d->rank, "neighbour rank out of range");
        unsigned char* base = static_cast<unsigned char*>(d->peer_base[r]);
        P.nb_rank[k] = r;
        P.nb_inbox[k] = reinterpret_cast<unsigned long long*>(base + kMailboxInbox);
    }
    // interface row of every entry of the plan's shared-node table; every interface node must be there
    const std::vector<uint32_t>& sh = tet_shared_node_table(nd->tet);
    std::vector<int32_t> row_of(nd->n, -1);
    for (size_t i = 0; i < H.if_idx_host.size(); ++i) row_of[H.if_idx_host[i]] = int32_t(i);
    std::vector<int32_t> sh_if_row(sh.size(), -1);
    size_t found = 0;
    for (size_t i = 0; i < sh.size(); ++i) if (sh[i] != 0xFFFFFFFFu && row_of[sh[i]] >= 0) { sh_if_row[i] = row_of[sh[i]]; ++found; }
    SB_CHECK(found == H.if_idx_host.size(), "interface nodes must be flagged in sofab200_tetfem_desc::shared_nodes when the force field is created");
    // where each interface row goes: {neighbour, row in its inbox}
    const int ms1 = std::max(1, H.max_sh - 1);
    std::vector<int2> if_send(H.if_idx_host.size() * size_t(ms1), make_int2(-1, 0));
    std::vector<int> fill(H.if_idx_host.size(), 0);
    for (int k = 0; k < P.n_nb; ++k)
        for (size_t i = 0; i < H.nb_rows_host[k].size(); ++i) {
            const uint32_t row = H.nb_rows_host[k][i];
            SB_CHECK(fill[row] < ms1, "interface node shared by more ranks than max_sharers");
            if_send[size_t(row) * ms1 + fill[row]++] = make_int2(k, int(d->remote_off[k] + i));
        }
    if (if_send.empty()) if_send.push_back(make_int2(-1, 0));
    SB_TRY(nd->peer.sh_if_row.upload(sh_if_row, s)); SB_TRY(nd->peer.if_send.upload(if_send, s));
    SB_CUDA(cudaStreamSynchronize(s));
    P.sh_if_row = nd->peer.sh_if_row.p; P.if_send = nd->peer.if_send.p; P.src = H.src.p; P.owned = H.owned.p;
    P.enabled = 1;
    nd->peer.ready = true;
    nd->invalidate_graphs();
    return SOFAB200_OK;
}
}  // namespace sb
extern "C" {
size_t sofab200_node_peer_bytes(const sofab200_node* node, size_t inbox_rows) {
    if (!node) return 0;
    return node->real == SOFAB200_F32 ? node_peer_bytes<float>(static_cast<const Node<float>*>(node), inbox_rows) : node_peer_bytes<double>(static_cast<const Node<double>*>(node), inbox_rows);
}
int sofab200_node_set_peer(sofab200_node* node, const sofab200_peer_desc* peer) {
    SB_CHECK(node && peer, "null argument");   /* peer->peer_base == NULL: leave peer mode */
    return NODE_DISPATCH(node, node_set_peer<float>(NF(node), peer), node_set_peer<double>(ND(node), peer));
}
int sofab200_node_set_distributed(sofab200_node* node, sofab200_comm* comm, const sofab200_halo_desc* halo) {
    SB_CHECK(node && comm && halo && halo->owned, "null argument");
    SB_CHECK(comm->ctx == node->ctx, "communicator and node belong to different contexts");
    return NODE_DISPATCH(node, node_set_distributed<float>(NF(node), comm, halo), node_set_distributed<double>(ND(node), comm, halo));
}

int sofab200_node_create(sofab200_ctx* ctx, sofab200_real real, size_t n_nodes, const sofab200_node_desc* desc, sofab200_node** out) {
    SB_CHECK(ctx && desc && out, "null argument");
    SB_CHECK((desc->tetfem != nullptr) != (desc->hexfem != nullptr), "exactly one of tetfem / hexfem must be given");
    if (desc->tetfem) SB_CHECK(tet_real(desc->tetfem) == int(real) && tet_nodes(desc->tetfem) == n_nodes, "force field and node disagree on Real or size");
    if (desc->hexfem) SB_CHECK(hex_real(desc->hexfem) == int(real) && hex_nodes(desc->hexfem) == n_nodes, "force field and node disagree on Real or size");
    SB_CHECK(desc->n_fixed == 0 || desc->fixed_host, "fixed indices are null");
    SB_CUDA(cudaSetDevice(ctx->device));
    if (real == SOFAB200_F32) return node_create<float>(ctx, n_nodes, desc, out);
    return node_create<double>(ctx, n_nodes, desc, out);
}
int sofab200_node_destroy(sofab200_node* node) { delete node; return SOFAB200_OK; }
int sofab200_node_set_params(sofab200_node* node, const sofab200_solver_params* p) {
    SB_CHECK(node && p, "null argument");
    SB_CHECK(p->iterations + 2 < unsigned(kMaxGraph), "iterations too large");
    node->prm = *p;
    return SOFAB200_OK;
}
int sofab200_node_compute_force(sofab200_node* node, void* f_dev, const void* x_dev) {
    SB_CHECK(node && f_dev && x_dev, "null argument");
    return NODE_DISPATCH(node, NF(node)->compute_force((float*)f_dev, (const float*)x_dev), ND(node)->compute_force((double*)f_dev, (const double*)x_dev));
}
int sofab200_node_apply(sofab200_node* node, void* q_dev, const void* p_dev, double m, double b, double k) {
    SB_CHECK(node && q_dev && p_dev && q_dev != p_dev, "null or aliased argument");
    return NODE_DISPATCH(node, NF(node)->apply((float*)q_dev, (const float*)p_dev, m, b, k), ND(node)->apply((double*)q_dev, (const double*)p_dev, m, b, k));
}
int sofab200_node_add_mbkdx(sofab200_node* node, void* out_dev, const void* init_dev, const void* d_dev, double m, double b, double k, int scale, double sf, int project) {
    SB_CHECK(node && out_dev && d_dev && out_dev != d_dev, "null or aliased argument");
    return NODE_DISPATCH(node, NF(node)->add_mbk((float*)out_dev, (const float*)init_dev, (const float*)d_dev, m, b, k, scale != 0, sf, project != 0, DOT_NONE, nullptr),
                         ND(node)->add_mbk((double*)out_dev, (const double*)init_dev, (const double*)d_dev, m, b, k, scale != 0, sf, project != 0, DOT_NONE, nullptr));
}
int sofab200_node_set_mesh_mass(sofab200_node* node, sofab200_meshmass* mesh_mass) {
    SB_CHECK(node && mesh_mass, "null argument");
    int mreal = 0; size_t mn = 0; const sofab200_ctx* mctx = nullptr;
    sb::meshmass_info(mesh_mass, &mreal, &mn, &mctx);
    SB_CHECK(mreal == node->real && mn == node->n && mctx == node->ctx, "the MeshMatrixMass must have the node's real type, size and context");
    if (node->real == SOFAB200_F32) { SB_CHECK(!NF(node)->distributed(), "MeshMatrixMass is not available in a distributed node"); NF(node)->mesh_mass = mesh_mass; NF(node)->has_mass = true; NF(node)->invalidate_graphs(); }
    else { SB_CHECK(!ND(node)->distributed(), "MeshMatrixMass is not available in a distributed node"); ND(node)->mesh_mass = mesh_mass; ND(node)->has_mass = true; ND(node)->invalidate_graphs(); }
    return SOFAB200_OK;
}
int sofab200_node_set_vertex_mass(sofab200_node* node, const void* vertex_mass_host) {
    SB_CHECK(node && vertex_mass_host, "null argument");
    cudaStream_t s = node->ctx->stream;
    if (node->real == SOFAB200_F32) { auto* n = NF(node); if (!n->mass.p) { SB_TRY(n->mass.alloc(n->n)); n->invalidate_graphs(); } if (!n->has_mass) n->invalidate_graphs(); n->has_mass = true; SB_CUDA(cudaMemcpyAsync(n->mass.p, vertex_mass_host, n->n * 4, cudaMemcpyHostToDevice, s)); }
    else { auto* n = ND(node); if (!n->mass.p) { SB_TRY(n->mass.alloc(n->n)); n->invalidate_graphs(); } if (!n->has_mass) n->invalidate_graphs(); n->has_mass = true; SB_CUDA(cudaMemcpyAsync(n->mass.p, vertex_mass_host, n->n * 8, cudaMemcpyHostToDevice, s)); }
    SB_CUDA(cudaStreamSynchronize(s));
    return SOFAB200_OK;
}
int sofab200_node_cg_solve(sofab200_node* node, void* x_dev, const void* b_dev, double m, double b, double k, int* nb_iter_host) {
    SB_CHECK(node && x_dev && b_dev, "null argument");
    SB_TRY(NODE_DISPATCH(node, NF(node)->cg_solve((float*)x_dev, (const float*)b_dev, m, b, k), ND(node)->cg_solve((double*)x_dev, (const double*)b_dev, m, b, k)));
    if (nb_iter_host) return sofab200_node_last_solve(node, nb_iter_host, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
    return SOFAB200_OK;
}
int sofab200_node_step(sofab200_node* node, void* x_dev, void* v_dev) {
    SB_CHECK(node && x_dev && v_dev, "null argument");
    return NODE_DISPATCH(node, NF(node)->step((float*)x_dev, (float*)v_dev), ND(node)->step((double*)x_dev, (double*)v_dev));
}
}  // extern "C"
namespace sb {
template <class R> static int node_step_host(Node<R>* n, void* x_host, void* v_host) {
    const size_t bytes = 3 * n->n * sizeof(R);
    cudaStream_t s = n->ctx->stream;
    if (!n->hx.p) { SB_TRY(n->hx.alloc(3 * n->n)); SB_TRY(n->hv.alloc(3 * n->n)); }
    if (!n->side_stream) { SB_CUDA(cudaStreamCreateWithFlags(&n->side_stream, cudaStreamNonBlocking)); SB_CUDA(cudaEventCreateWithFlags(&n->side_event, cudaEventDisableTiming)); }
    // x first on the main stream, v on the side stream: addForce needs only x and runs while v is still on the bus
    SB_CUDA(cudaEventRecord(n->side_event, s));
    SB_CUDA(cudaStreamWaitEvent(n->side_stream, n->side_event, 0));           // (the previous step's readers of hv are done)
    SB_CUDA(cudaMemcpyAsync(n->hx.p, x_host, bytes, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(n->hv.p, v_host, bytes, cudaMemcpyHostToDevice, n->side_stream));
    SB_CUDA(cudaEventRecord(n->side_event, n->side_stream));
    if (n->has_plane) SB_CUDA(cudaStreamWaitEvent(s, n->side_event, 0));      // (the plane's damping term reads v: no overlap then)
    SB_TRY(n->compute_force(n->f.p, n->hx.p, n->has_plane ? n->hv.p : nullptr));
    SB_CUDA(cudaStreamWaitEvent(s, n->side_event, 0));
    SB_TRY(n->step(n->hx.p, n->hv.p, true));
    SB_CUDA(cudaMemcpyAsync(x_host, n->hx.p, bytes, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(v_host, n->hv.p, bytes, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    return SOFAB200_OK;
}
// accumulateForce (MechanicalObject.inl:1356-1375): f[i] += externalForce[i] on a freshly reset f, i.e. every component becomes +0 + ext
// (a negative zero does not survive; rows equal to Deriv() are skipped by the reference, which leaves the same +0)
template <class R> __global__ void __launch_bounds__(kVecBlock) ext_accumulate_kernel(size_t n3, R* __restrict__ r) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n3; i += size_t(gridDim.x) * blockDim.x) r[i] = R(0) + r[i];
}
// (distributed node: every rank holds a copy of its interface nodes and their forces are summed over the sharing ranks, so only the owner's copy
// carries the external force -- the same masking as for the DiagonalMass term)
template <class R> static int node_finish_external_force(Node<R>* n) {
    LAUNCH(n->ctx, (ext_accumulate_kernel<R>), vec_grid(3 * n->n, n->ctx->sm_count), kVecBlock, 3 * n->n, n->ext.p);
    if (n->distributed()) LAUNCH(n->ctx, (mask_rows_kernel<R>), vec_grid(n->n, n->ctx->sm_count), kVecBlock, n->n, (const unsigned char*)n->halo.owned.p, (const R*)n->ext.p, n->ext.p);
    return SOFAB200_OK;
}
template <class R> static int node_upload_external_force(Node<R>* n, const void* ext_host, bool sync) {
    cudaStream_t s = n->ctx->stream;
    const bool on = ext_host != nullptr;
    if (on != n->has_ext) { n->has_ext = on; n->invalidate_graphs(); }
    if (!on) return SOFAB200_OK;
    if (!n->ext.p) SB_TRY(n->ext.alloc(3 * n->n));
    SB_CUDA(cudaMemcpyAsync(n->ext.p, ext_host, 3 * n->n * sizeof(R), cudaMemcpyHostToDevice, s));
    SB_TRY(node_finish_external_force(n));
    if (sync) SB_CUDA(cudaStreamSynchronize(s));
    return SOFAB200_OK;
}
// One step of a device-resident state coupled to a host loop: this step's external forces come up from (pinned) host memory, the new positions
// go down on a copy stream into x_out_host while the caller already submits the next step; the call returns once the PREVIOUS step's positions
// are complete on the host (so two output buffers must alternate), sofab200_node_flush waits for the last ones.
template <class R> static int node_step_pipelined(Node<R>* n, R* x_dev, R* v_dev, const void* ext_host, void* x_out_host) {
    const size_t bytes = 3 * n->n * sizeof(R);
    cudaStream_t s = n->ctx->stream;
    if (!n->copy_stream) {
        SB_CUDA(cudaStreamCreateWithFlags(&n->copy_stream, cudaStreamNonBlocking));
        SB_CUDA(cudaStreamCreateWithFlags(&n->up_stream, cudaStreamNonBlocking));
        SB_CUDA(cudaEventCreateWithFlags(&n->ev_h2d, cudaEventDisableTiming));
        SB_CUDA(cudaEventCreateWithFlags(&n->ev_ext_used, cudaEventDisableTiming));
        SB_TRY(n->ext_stage.alloc(3 * n->n));
        for (int i = 0; i < 2; ++i) {
            SB_CUDA(cudaEventCreateWithFlags(&n->ev_step[i], cudaEventDisableTiming)); SB_CUDA(cudaEventCreateWithFlags(&n->ev_copy[i], cudaEventDisableTiming));
            SB_TRY(n->xout[i].alloc(3 * n->n));
        }
    }
    if (ext_host) {
        // this step's forces go up on their own stream, under whatever step is still running; the main stream only moves them from the staging
        // buffer into the vector the captured step reads (a device-to-device copy) once the previous step is through
        if (!n->has_ext) { n->has_ext = true; n->invalidate_graphs(); }
        if (!n->ext.p) SB_TRY(n->ext.alloc(3 * n->n));
        if (n->ext_used_recorded) SB_CUDA(cudaStreamWaitEvent(n->up_stream, n->ev_ext_used, 0));    // (the staging buffer has been consumed)
        SB_CUDA(cudaMemcpyAsync(n->ext_stage.p, ext_host, bytes, cudaMemcpyHostToDevice, n->up_stream));
        SB_CUDA(cudaEventRecord(n->ev_h2d, n->up_stream));
        SB_CUDA(cudaStreamWaitEvent(s, n->ev_h2d, 0));
        SB_CUDA(cudaMemcpyAsync(n->ext.p, n->ext_stage.p, bytes, cudaMemcpyDeviceToDevice, s));
        SB_TRY(node_finish_external_force(n));
        SB_CUDA(cudaEventRecord(n->ev_ext_used, s));
        n->ext_used_recorded = true;
    } else SB_TRY(node_upload_external_force(n, nullptr, false));
    SB_TRY(n->step(x_dev, v_dev, false));
    const int slot = int(n->pipe_k & 1ull);
    if (n->copy_pending[slot]) SB_CUDA(cudaStreamWaitEvent(s, n->ev_copy[slot], 0));       // (the copy of two steps ago has left this staging buffer)
    SB_CUDA(cudaMemcpyAsync(n->xout[slot].p, x_dev, bytes, cudaMemcpyDeviceToDevice, s));
    SB_CUDA(cudaEventRecord(n->ev_step[slot], s));
    SB_CUDA(cudaStreamWaitEvent(n->copy_stream, n->ev_step[slot], 0));
    SB_CUDA(cudaMemcpyAsync(x_out_host, n->xout[slot].p, bytes, cudaMemcpyDeviceToHost, n->copy_stream));
    SB_CUDA(cudaEventRecord(n->ev_copy[slot], n->copy_stream));
    n->copy_pending[slot] = true;
    if (n->copy_pending[slot ^ 1]) SB_CUDA(cudaEventSynchronize(n->ev_copy[slot ^ 1]));    // the previous step's positions are on the host now
    if (ext_host) SB_CUDA(cudaEventSynchronize(n->ev_h2d));                                // ... and this step's forces have left ext_host (the caller may refill it)
    ++n->pipe_k;
    return SOFAB200_OK;
}
template <class R> static int node_flush(Node<R>* n) {
    for (int i = 0; i < 2; ++i) if (n->copy_pending[i]) SB_CUDA(cudaEventSynchronize(n->ev_copy[i]));
    SB_CUDA(cudaStreamSynchronize(n->ctx->stream));
    return SOFAB200_OK;
}
// x round trip only: the velocities stay resident in HBM between steps (what the reference's own loop does with a device-typed state)
template <class R> static int node_step_host_x(Node<R>* n, void* x_host, const void* v_host_in, void* v_host_out) {
    const size_t bytes = 3 * n->n * sizeof(R);
    cudaStream_t s = n->ctx->stream;
    if (!n->hx.p) { SB_TRY(n->hx.alloc(3 * n->n)); SB_TRY(n->hv.alloc(3 * n->n)); SB_TRY(n->hv.zero(s)); }
    SB_CUDA(cudaMemcpyAsync(n->hx.p, x_host, bytes, cudaMemcpyHostToDevice, s));
    if (v_host_in) SB_CUDA(cudaMemcpyAsync(n->hv.p, v_host_in, bytes, cudaMemcpyHostToDevice, s));
    SB_TRY(n->step(n->hx.p, n->hv.p, false));
    SB_CUDA(cudaMemcpyAsync(x_host, n->hx.p, bytes, cudaMemcpyDeviceToHost, s));
    if (v_host_out) SB_CUDA(cudaMemcpyAsync(v_host_out, n->hv.p, bytes, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    return SOFAB200_OK;
}
}  // namespace sb
extern "C" {
int sofab200_node_step_host_x(sofab200_node* node, void* x_host, const void* v_host_in, void* v_host_out) {
    SB_CHECK(node && x_host, "null argument");
    return NODE_DISPATCH(node, sb::node_step_host_x<float>(NF(node), x_host, v_host_in, v_host_out), sb::node_step_host_x<double>(ND(node), x_host, v_host_in, v_host_out));
}
int sofab200_node_set_external_force(sofab200_node* node, const void* ext_host) {
    SB_CHECK(node != nullptr, "null argument");
    return NODE_DISPATCH(node, sb::node_upload_external_force<float>(NF(node), ext_host, true), sb::node_upload_external_force<double>(ND(node), ext_host, true));
}
int sofab200_node_step_pipelined(sofab200_node* node, void* x_dev, void* v_dev, const void* ext_host, void* x_out_host) {
    SB_CHECK(node && x_dev && v_dev && x_out_host, "null argument");
    return NODE_DISPATCH(node, sb::node_step_pipelined<float>(NF(node), (float*)x_dev, (float*)v_dev, ext_host, x_out_host),
                         sb::node_step_pipelined<double>(ND(node), (double*)x_dev, (double*)v_dev, ext_host, x_out_host));
}
int sofab200_node_flush(sofab200_node* node) {
    SB_CHECK(node != nullptr, "null argument");
    return NODE_DISPATCH(node, sb::node_flush<float>(NF(node)), sb::node_flush<double>(ND(node)));
}
int sofab200_node_step_host(sofab200_node* node, void* x_host, void* v_host) {
    SB_CHECK(node && x_host && v_host, "null argument");
    return NODE_DISPATCH(node, sb::node_step_host<float>(NF(node), x_host, v_host), sb::node_step_host<double>(ND(node), x_host, v_host));
}
int sofab200_node_last_solve(sofab200_node* node, int* nb_iter, int* end_cond, double* graph_error, size_t* n_error, double* graph_den, size_t* n_den, size_t cap) {
    SB_CHECK(node != nullptr, "null argument");
    CGDev* dev = node->real == SOFAB200_F32 ? NF(node)->cg.p : ND(node)->cg.p;
    static thread_local CGDev h;
    SB_CUDA(cudaMemcpyAsync(&h, dev, sizeof(CGDev), cudaMemcpyDeviceToHost, node->ctx->stream));
    SB_CUDA(cudaStreamSynchronize(node->ctx->stream));
    if (nb_iter) *nb_iter = h.nb_iter;
    if (end_cond) *end_cond = h.end_cond;
    if (n_error) *n_error = size_t(h.n_err);
    if (n_den) *n_den = size_t(h.n_den);
    if (graph_error) for (size_t i = 0; i < cap && i < size_t(h.n_err); ++i) graph_error[i] = h.graph_error[i];
    if (graph_den) for (size_t i = 0; i < cap && i < size_t(h.n_den); ++i) graph_den[i] = h.graph_den[i];
    return SOFAB200_OK;
}
int sofab200_node_get(sofab200_node* node, const char* what, void* out_host) {
    SB_CHECK(node && what && out_host, "null argument");
    const std::string w(what);
    const void* src = nullptr;
    const size_t es = node->real == SOFAB200_F32 ? 4 : 8;
    if (w == "plane_contacts") {   // PlaneForceField m_contacts as n bytes (1 = in contact at the last addForce)
        const unsigned char* c = node->real == SOFAB200_F32 ? NF(node)->plane_contacts.p : ND(node)->plane_contacts.p;
        SB_CHECK(c != nullptr, "the node has no PlaneForceField");
        SB_CUDA(cudaMemcpyAsync(out_host, c, node->n, cudaMemcpyDeviceToHost, node->ctx->stream));
        SB_CUDA(cudaStreamSynchronize(node->ctx->stream));
        return SOFAB200_OK;
    }
    if (node->real == SOFAB200_F32) { auto* n = NF(node); src = w == "f" ? n->f.p : w == "b" ? n->b.p : w == "dx" ? n->dx.p : nullptr; }
    else { auto* n = ND(node); src = w == "f" ? n->f.p : w == "b" ? n->b.p : w == "dx" ? n->dx.p : nullptr; }
    SB_CHECK(src != nullptr, "unknown vector name");
    SB_CUDA(cudaMemcpyAsync(out_host, src, 3 * node->n * es, cudaMemcpyDeviceToHost, node->ctx->stream));
    SB_CUDA(cudaStreamSynchronize(node->ctx->stream));
    return SOFAB200_OK;
}
int sofab200_node_cg_kernel_info(const sofab200_node* node, int out[8]) {
    SB_CHECK(node && out, "null argument");
    const int* fi = node->real == SOFAB200_F32 ? static_cast<const Node<float>*>(node)->fused_info : static_cast<const Node<double>*>(node)->fused_info;
    const bool fused = node->real == SOFAB200_F32 ? static_cast<const Node<float>*>(node)->fused : static_cast<const Node<double>*>(node)->fused;
    const bool persistent = node->real == SOFAB200_F32 ? static_cast<const Node<float>*>(node)->persistent : static_cast<const Node<double>*>(node)->persistent;
    for (int i = 0; i < 6; ++i) out[i] = fi[i];
    out[6] = fused ? 1 : 0; out[7] = persistent ? 1 : 0;
    return SOFAB200_OK;
}
int sofab200_node_reset(sofab200_node* node) {
    SB_CHECK(node != nullptr, "null argument");
    CGDev* dev = node->real == SOFAB200_F32 ? NF(node)->cg.p : ND(node)->cg.p;
    SB_CUDA(cudaMemsetAsync(dev, 0, sizeof(CGDev), node->ctx->stream));
    return SOFAB200_OK;
}

}  // extern "C"
