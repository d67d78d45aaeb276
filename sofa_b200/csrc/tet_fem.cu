// TetrahedronFEMForceField<B200Vec3Types>: device upload + launches + C ABI.
#include <cstring>
#include <memory>

#include "tet_host.h"

using namespace sb;

#include "tet_handle.h"

namespace sb {

template <class R> struct TetFF : sofab200_tetfem {
    HostTet<R> h;
    DevBuf<ushort4> lnode; DevBuf<uint4> slot; DevBuf<uint32_t> orig;
    DevBuf<Quad<R>> rk0, rk1, rk2, j0, j1, j2, js0, js1, js2, x0a, x0b, x0c, sv0, sv1, sv2, sv3, sv4;
    DevBuf<uint32_t> tile_node_off, tile_nodes, tile_shslot, tile_nint, tile_nb, sh_nodes, sh_base;
    DevBuf<uint16_t> tile_val, tile_jds, sh_val;
    DevBuf<Quad<R>> stage;
    bool update_j = false;   // updateStiffnessMatrix: addForce rewrites the cofactor planes
    bool split_j = false;    // TetrahedronFEMForceField, method large, updateStiffnessMatrix: only the normal-strain copies are rewritten; js0..2 keep the shear copies (TM_*_JS)
    bool sibling = false; DevBuf<R> a0_el;   // TetrahedralCorotationalFEMForceField (its getRotation differs)
    DevBuf<R> vm_shf, vm_lambda, vm_mu, vm_rest, vm_elem; int von_mises = 0;   // computeVonMisesStress (uploaded at the first call)
    DevBuf<Quad<R>> pl0, pl1; double plastic[3] = {0, 0, 0};   // _plasticStrains in tile order (plasticMaxThreshold > 0 only)
    DevBuf<R> rot_export;
    DevBuf<uint32_t> inc_off, inc_es, inc_e; DevBuf<R> r0t_el; uint32_t es_of_first = 0;   // getRotations: node -> incident elements (built at the first call)
    int threads = 256;   // CTA size of the addDForce tile kernel
    bool prefetch = true;
    TetDev<R> dev() {
        const HostPlan& plan = h.plan;
        TetDev<R> d;
        d.t.n_nodes = int(n_nodes); d.t.n_elems = int(n_tets); d.t.n_tiles = plan.n_tiles; d.t.tile_e = plan.tile_e; d.t.maxval = plan.maxval;
        d.t.tile_node_off = tile_node_off.p; d.t.tile_nodes = tile_nodes.p; d.t.tile_shslot = tile_shslot.p; d.t.tile_nint = tile_nint.p; d.t.tile_nb = tile_nb.p; d.t.tile_val = tile_val.p; d.t.tile_jds = tile_jds.p;
        d.t.n_shared = plan.n_shared; d.t.n_chunks = plan.n_chunks; d.t.sh_nodes = sh_nodes.p; d.t.sh_val = sh_val.p; d.t.sh_base = sh_base.p;
        d.t.stage = stage.p; d.t.stage_n = plan.stage_n;
        d.lnode = lnode.p; d.slot = slot.p;
        d.rk0 = rk0.p; d.rk1 = rk1.p; d.rk2 = rk2.p; d.j0 = j0.p; d.j1 = j1.p; d.j2 = j2.p;
        d.x0a = x0a.p; d.x0b = x0b.p; d.x0c = x0c.p; d.sv0 = sv0.p; d.sv1 = sv1.p; d.sv2 = sv2.p; d.sv3 = sv3.p; d.sv4 = sv4.p;
        d.k_factor = R(0);
        d.js0 = js0.p; d.js1 = js1.p; d.js2 = js2.p;
        d.j0w = update_j ? j0.p : nullptr; d.j1w = update_j ? j1.p : nullptr; d.j2w = update_j ? j2.p : nullptr;
        d.pl0 = pl0.p; d.pl1 = pl1.p; d.plastic_max = R(plastic[0]); d.plastic_yield = R(plastic[1]); d.plastic_creep = R(plastic[2]);
        return d;
    }
};

template <class R> static int tet_upload(TetFF<R>& ff) {
    HostTet<R>& H = ff.h;
    const HostPlan& P = H.plan;
    cudaStream_t s = ff.ctx->stream;
    SB_TRY(ff.lnode.upload(H.lnode, s)); SB_TRY(ff.slot.upload(H.slot, s)); SB_TRY(ff.orig.upload(P.order, s));
    SB_TRY(ff.rk0.upload(H.rk0, s)); SB_TRY(ff.rk1.upload(H.rk1, s)); SB_TRY(ff.rk2.upload(H.rk2, s));
    SB_TRY(ff.j0.upload(H.j0, s)); SB_TRY(ff.j1.upload(H.j1, s)); SB_TRY(ff.j2.upload(H.j2, s));
    if (ff.split_j) { SB_TRY(ff.js0.upload(H.j0, s)); SB_TRY(ff.js1.upload(H.j1, s)); SB_TRY(ff.js2.upload(H.j2, s)); }
    SB_TRY(ff.x0a.upload(H.x0a, s)); SB_TRY(ff.x0b.upload(H.x0b, s)); SB_TRY(ff.x0c.upload(H.x0c, s));
    if (H.method == SOFAB200_TET_SVD) {
        SB_TRY(ff.sv0.upload(H.sv[0], s)); SB_TRY(ff.sv1.upload(H.sv[1], s)); SB_TRY(ff.sv2.upload(H.sv[2], s));
        SB_TRY(ff.sv3.upload(H.sv[3], s)); SB_TRY(ff.sv4.upload(H.sv[4], s));
    }
    SB_TRY(ff.tile_node_off.upload(P.tile_node_off, s)); SB_TRY(ff.tile_nodes.upload(P.tile_nodes, s)); SB_TRY(ff.tile_shslot.upload(P.tile_shslot, s)); SB_TRY(ff.tile_nint.upload(P.tile_nint, s)); SB_TRY(ff.tile_nb.upload(P.tile_nb, s));
    SB_TRY(ff.tile_val.upload(P.tile_val, s)); SB_TRY(ff.tile_jds.upload(P.tile_jds, s));
    SB_TRY(ff.sh_nodes.upload(P.sh_nodes, s)); SB_TRY(ff.sh_val.upload(P.sh_val, s)); SB_TRY(ff.sh_base.upload(P.sh_base, s));
    SB_TRY(ff.stage.alloc(P.stage_n)); SB_TRY(ff.stage.zero(s));
    if (ff.plastic[0] > 0) { const size_t NS = size_t(P.n_tiles) * P.tile_e; SB_TRY(ff.pl0.alloc(NS)); SB_TRY(ff.pl0.zero(s)); SB_TRY(ff.pl1.alloc(NS)); SB_TRY(ff.pl1.zero(s)); }
    SB_CUDA(cudaStreamSynchronize(s));
    // the tile-ordered host planes are no longer needed once they are resident in HBM
    for (auto* v : {&H.rk0, &H.rk1, &H.rk2, &H.j0, &H.j1, &H.j2, &H.x0a, &H.x0b, &H.x0c}) { v->clear(); v->shrink_to_fit(); }
    for (auto& v : H.sv) { v.clear(); v.shrink_to_fit(); }
    H.lnode.clear(); H.lnode.shrink_to_fit(); H.slot.clear(); H.slot.shrink_to_fit();
    return SOFAB200_OK;
}

template <class R, int MODE, int MAXT, bool PF> static int tet_launch_variant(TetFF<R>& ff, const TetDev<R>& d, const R* in, const NodeEpilogue<R>& ep) {
    auto kern = tet_tile_kernel<R, MODE, MAXT, PF>;
    // (the attribute is per device and context, not per thread: set it for the current device on every launch -- a host-side table lookup)
    const size_t smem_total = ff.h.smem_bytes;
    if (smem_total > 48 * 1024) SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_total)));
    const int cls = (MODE == TM_DF_COROT || MODE == TM_DF_SMALL || MODE == TM_DF_COROT_JS) ? 0 : 2;
    ff.ctx->prof_start(cls);
    kern<<<ff.h.plan.n_tiles, ((MODE == TM_DF_COROT || (MODE == TM_F_LARGE && MAXT > 256)) && sizeof(R) == 4) ? std::min(ff.threads, MAXT) : 256, smem_total, ff.ctx->stream>>>(d, in, ep, ff.h.plan.max_touched, ff.h.plan.max_slots);
    ff.ctx->prof_stop(cls);
    ff.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}
template <class R, int MODE> static int tet_launch_mode(TetFF<R>& ff, const TetDev<R>& d, const R* in, const NodeEpilogue<R>& ep) {
    // the addForce passes run once per step: one conservative variant; the addDForce pass (26x per step) is tuned
    if (MODE == TM_F_LARGE && sizeof(R) == 4 && ff.threads > 256) return ff.prefetch ? tet_launch_variant<R, MODE, 512, true>(ff, d, in, ep) : tet_launch_variant<R, MODE, 512, false>(ff, d, in, ep);
    if (MODE != TM_DF_COROT) return tet_launch_variant<R, MODE, 256, false>(ff, d, in, ep);
    if (sizeof(R) == 8) return tet_launch_variant<R, MODE, 256, false>(ff, d, in, ep);
    if (ff.threads > 512) return tet_launch_variant<R, MODE, 1024, false>(ff, d, in, ep);
    if (ff.threads > 256) return ff.prefetch ? tet_launch_variant<R, MODE, 512, true>(ff, d, in, ep) : tet_launch_variant<R, MODE, 512, false>(ff, d, in, ep);
    return ff.prefetch ? tet_launch_variant<R, MODE, 256, true>(ff, d, in, ep) : tet_launch_variant<R, MODE, 256, false>(ff, d, in, ep);
}

// ---- persistent CG kernel (cooperative launch: every CTA must be resident, the kernel crosses grid-wide syncs) -----------
// Returns SOFAB200_OK, an error, or kPersistNotEligible (> 0) when the mesh does not fit the kernel's assumptions (at most
// two tiles and one chunk of shared nodes per thread group and CTA): the caller then runs the multi-kernel loop.
template <class R, int MODE, int MAXT, bool PF> static int tet_persist_variant(TetFF<R>& ff, TetDev<R> d, PersistCG<R> a, int threads, size_t sync_capacity, bool dry_run = false) {
    auto kern = tet_cg_persistent_kernel<R, MODE, MAXT, PF>;
    const HostPlan& P = ff.h.plan;
    const int groups = threads / kGatherChunk;
    // grid: as many CTAs as the GPU holds at once, at most one per tile (or per group of chunks if there are more of those)
    cudaFuncAttributes fa;
    SB_CUDA(cudaFuncGetAttributes(&fa, kern));
    int dev_smem_optin = 0;
    SB_CUDA(cudaDeviceGetAttribute(&dev_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ff.ctx->device));
    for (int tiles_per_cta = 1; tiles_per_cta <= 2; ++tiles_per_cta) {
        const PersistLayout L = persist_layout<R>(tiles_per_cta, P.max_touched, P.max_slots, P.max_int, P.max_shtouch, P.maxval);
        if (L.total + fa.sharedSizeBytes > size_t(dev_smem_optin)) break;
        SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(std::max<unsigned>(L.total, 1024))));
        int per_sm = 0;
        SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, L.total));
        const int max_grid = per_sm * ff.ctx->sm_count;
        if (max_grid < 1) break;
        const int need_tiles = (P.n_tiles + tiles_per_cta - 1) / tiles_per_cta, need_chunks = (P.n_chunks + groups - 1) / groups;
        const int grid = std::max(need_tiles, need_chunks);
        if (grid > max_grid) continue;
        if (size_t(3) * grid + 1 > sync_capacity) return fail(SOFAB200_ERR_INVALID, "sync buffer too small for the persistent CG kernel");
        if (dry_run) return SOFAB200_OK;
        a.lay = L;
        void* args[] = {&d, &a};
        ff.ctx->prof_start(4);
        SB_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(threads), args, L.total, ff.ctx->stream));
        ff.ctx->prof_stop(4);
        ff.ctx->launches++;
        return SOFAB200_OK;
    }
    return kPersistNotEligible;
}
// dry_run: only tell whether the mesh fits the kernel (SOFAB200_OK) or not (kPersistNotEligible)
template <class R> int tet_cg_persistent(sofab200_tetfem* base, R k_factor, PersistCG<R> a, size_t sync_capacity, bool dry_run) {
    if (base->kind == 1) return kPersistNotEligible;      // FastTetrahedralCorotationalForceField: the multi-kernel loop
    TetFF<R>& ff = *static_cast<TetFF<R>*>(base);
    TetDev<R> d = ff.dev();
    d.k_factor = k_factor;
    if (ff.split_j) return kPersistNotEligible;       // two cofactor sets per element: the multi-kernel loop serves it
    if (ff.method == SOFAB200_TET_SMALL) return tet_persist_variant<R, TM_DF_SMALL, 256, false>(ff, d, a, 256, sync_capacity, dry_run);
    if (sizeof(R) == 8) return tet_persist_variant<R, TM_DF_COROT, 256, false>(ff, d, a, 256, sync_capacity, dry_run);
    if (ff.threads > 256) return tet_persist_variant<R, TM_DF_COROT, 512, true>(ff, d, a, 512, sync_capacity, dry_run);
    return tet_persist_variant<R, TM_DF_COROT, 256, true>(ff, d, a, 256, sync_capacity, dry_run);
}
template int tet_cg_persistent<float>(sofab200_tetfem*, float, PersistCG<float>, size_t, bool);
template int tet_cg_persistent<double>(sofab200_tetfem*, double, PersistCG<double>, size_t, bool);

// ---- fused persistent CG kernel (cg_fused.cuh): any number of tiles per CTA -----------------------------------------------------------
// Returns SOFAB200_OK, an error, or kPersistNotEligible when not even the streamed layout fits.  info (optional): {grid, tiles per CTA,
// cached, dynamic shared memory, element threads, gather threads}.
template <class R, int MODE, bool PF, int ET, int GT, bool CACHED> static int tet_fused_launch(TetFF<R>& ff, TetDev<R> d, FusedCG<R> a, const FusedLayout& L, int grid, bool dry_run, int* info) {
    auto kern = fused_cg_kernel<R, TetPass<R, MODE, PF>, ET, GT, CACHED>;
    cudaFuncAttributes fa;
    SB_CUDA(cudaFuncGetAttributes(&fa, kern));
    int dev_smem_optin = 0;
    SB_CUDA(cudaDeviceGetAttribute(&dev_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ff.ctx->device));
    if (L.total + fa.sharedSizeBytes > size_t(dev_smem_optin)) return kPersistNotEligible;
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(std::max<unsigned>(L.total, 1024))));
    int per_sm = 0;
    SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ET + GT, L.total));
    if (per_sm < 1) return kPersistNotEligible;
    if (info) { info[0] = grid; info[1] = L.tiles_per_cta; info[2] = L.cached; info[3] = int(L.total); info[4] = ET; info[5] = GT; }
    if (dry_run) return SOFAB200_OK;
    a.lay = L;
    int ded_share = 50;       // per cent of the CTA's units of shared nodes that go to the dedicated warps
    if (const char* env = getenv("SOFAB200_FUSED_GATHER_SHARE")) ded_share = std::max(0, std::min(100, atoi(env)));
    void* args[] = {&d, &a, &ded_share};
    ff.ctx->prof_start(4);
    SB_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(ET + GT), args, L.total, ff.ctx->stream));
    ff.ctx->prof_stop(4);
    ff.ctx->launches++;
    return SOFAB200_OK;
}
template <class R, int MODE, bool PF, int ET, int GT> static int tet_fused_variant(TetFF<R>& ff, TetDev<R> d, FusedCG<R> a, size_t sync_capacity, bool dry_run, int* info) {
    const HostPlan& P = ff.h.plan;
    const int grid = std::max(1, std::min(ff.ctx->sm_count, P.n_tiles));
    const int tiles_per_cta = (P.n_tiles + grid - 1) / grid;
    const int n_units = P.n_chunks * (kGatherChunk / kUnit);
    const int units_per_cta = (n_units + grid - 1) / grid;
    if (fused_sync_words(grid) > sync_capacity) return fail(SOFAB200_ERR_INVALID, "sync buffer too small for the fused CG kernel");
    // cached layout when the CTA's tiles fit next to the slots inside the 196 KB carve-out (a larger one throttles the element stream), else streamed
    size_t cached_limit = 196 * 1024;
    if (const char* env = getenv("SOFAB200_FUSED_CACHED_KB")) { const int v = atoi(env); if (v >= 0 && v <= 227) cached_limit = size_t(v) * 1024; }
    const FusedLayout Lc = fused_layout<R>(true, tiles_per_cta, units_per_cta, P.max_touched, P.max_slots, P.max_int, P.max_shtouch, P.maxval);
    if (tiles_per_cta <= 2 && Lc.total + 2048 <= cached_limit) {
        const int rc = tet_fused_launch<R, MODE, PF, ET, GT, true>(ff, d, a, Lc, grid, dry_run, info);
        if (rc != kPersistNotEligible) return rc;
    }
    const FusedLayout Ls = fused_layout<R>(false, tiles_per_cta, units_per_cta, P.max_touched, P.max_slots, P.max_int, P.max_shtouch, P.maxval);
    return tet_fused_launch<R, MODE, PF, ET, GT, false>(ff, d, a, Ls, grid, dry_run, info);
}
template <class R> int tet_cg_fused(sofab200_tetfem* base, R k_factor, FusedCG<R> a, size_t sync_capacity, bool dry_run, int* info) {
    if (base->kind == 1) return fast_cg_fused<R>(base, k_factor, a, sync_capacity, dry_run, info);      // FastTetrahedralCorotationalForceField: A*p over the edges (fast_fem.cu)
    TetFF<R>& ff = *static_cast<TetFF<R>*>(base);
    TetDev<R> d = ff.dev();
    d.k_factor = k_factor;
    if (ff.split_j) return kPersistNotEligible;       // two cofactor sets per element: the multi-kernel loop serves it
    int gw = 0;   // dedicated gather warps
    if (const char* env = getenv("SOFAB200_FUSED_GATHER_WARPS")) gw = atoi(env);
    if (ff.method == SOFAB200_TET_SMALL) return tet_fused_variant<R, TM_DF_SMALL, false, 256, 0>(ff, d, a, sync_capacity, dry_run, info);
    if (sizeof(R) == 8) return tet_fused_variant<R, TM_DF_COROT, false, 256, 0>(ff, d, a, sync_capacity, dry_run, info);
    if (ff.threads > 256) {
        if (gw >= 4) return tet_fused_variant<R, TM_DF_COROT, true, 512, 128>(ff, d, a, sync_capacity, dry_run, info);
        if (gw >= 2) return tet_fused_variant<R, TM_DF_COROT, true, 512, 64>(ff, d, a, sync_capacity, dry_run, info);
        return tet_fused_variant<R, TM_DF_COROT, true, 512, 0>(ff, d, a, sync_capacity, dry_run, info);
    }
    return tet_fused_variant<R, TM_DF_COROT, true, 256, 0>(ff, d, a, sync_capacity, dry_run, info);
}
template int tet_cg_fused<float>(sofab200_tetfem*, float, FusedCG<float>, size_t, bool, int*);
template int tet_cg_fused<double>(sofab200_tetfem*, double, FusedCG<double>, size_t, bool, int*);
size_t tet_shared_slot_count(sofab200_tetfem* base) {
    if (base->kind == 1) return fast_shared_slot_count(base);
    if (base->real == SOFAB200_F32) return size_t(static_cast<TetFF<float>*>(base)->h.plan.n_chunks) * kGatherChunk;
    return size_t(static_cast<TetFF<double>*>(base)->h.plan.n_chunks) * kGatherChunk;
}

// Element pass + boundary gather with a caller-provided epilogue (also used by the solver node).
template <class R> TileDev<R> tet_tiledev(sofab200_tetfem* base) { if (base->kind == 1) return fast_tiledev<R>(base); return static_cast<TetFF<R>*>(base)->dev().t; }
template TileDev<float> tet_tiledev<float>(sofab200_tetfem*);
template TileDev<double> tet_tiledev<double>(sofab200_tetfem*);

// skip_gather: the caller sums the shared nodes itself (fused CG tail kernel)
template <class R> int tet_run(sofab200_tetfem* base, bool dforce, const R* in, R k_factor, NodeEpilogue<R> ep, bool skip_gather) {
    if (base->kind == 1) return fast_run<R>(base, dforce, in, k_factor, ep, skip_gather);
    TetFF<R>& ff = *static_cast<TetFF<R>*>(base);
    const HostPlan& plan = ff.h.plan;
    SB_CHECK((!ep.mdx_src || ep.mdx_src == in) && (!ep.dot_with || ep.dot_with == in) && (!ep.plane_mode || ep.plane_in == in), "mass / dot / plane operands must be the pass's input vector");
    TetDev<R> d = ff.dev();
    d.k_factor = k_factor;
    ep.partial_base = 0;
    ep.partial_total = plan.n_tiles + plan.n_chunks;
    if (dforce) {
        if (ff.method == SOFAB200_TET_SMALL) SB_TRY((tet_launch_mode<R, TM_DF_SMALL>(ff, d, in, ep)));
        else if (ff.split_j) SB_TRY((tet_launch_mode<R, TM_DF_COROT_JS>(ff, d, in, ep)));
        else SB_TRY((tet_launch_mode<R, TM_DF_COROT>(ff, d, in, ep)));
    } else {
        switch (ff.method) {
        case SOFAB200_TET_SMALL: SB_TRY((tet_launch_mode<R, TM_F_SMALL>(ff, d, in, ep))); break;
        case SOFAB200_TET_LARGE:
            if (ff.split_j) SB_TRY((tet_launch_mode<R, TM_F_LARGE_JS>(ff, d, in, ep)));
            else SB_TRY((tet_launch_mode<R, TM_F_LARGE>(ff, d, in, ep)));
            break;
        case SOFAB200_TET_POLAR: SB_TRY((tet_launch_mode<R, TM_F_POLAR>(ff, d, in, ep))); break;
        default: SB_TRY((tet_launch_mode<R, TM_F_SVD>(ff, d, in, ep))); break;
        }
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
template int tet_run<float>(sofab200_tetfem*, bool, const float*, float, NodeEpilogue<float>, bool);
template int tet_run<double>(sofab200_tetfem*, bool, const double*, double, NodeEpilogue<double>, bool);

int tet_real(sofab200_tetfem* ff) { return ff->real; }
bool tet_is_fast(sofab200_tetfem* ff) { return ff->kind == 1; }
// the plan's table of shared nodes (chunks of kGatherChunk, 0xFFFFFFFF = padding), for the multi-GPU set-up
const std::vector<uint32_t>& tet_shared_node_table(sofab200_tetfem* base) {
    static const std::vector<uint32_t> none;
    if (base->kind == 1) return none;
    if (base->real == SOFAB200_F32) return static_cast<TetFF<float>*>(base)->h.plan.sh_nodes;
    return static_cast<TetFF<double>*>(base)->h.plan.sh_nodes;
}
size_t tet_tile_node_count(sofab200_tetfem* base) {
    if (base->kind == 1) return fast_tile_node_count(base);
    if (base->real == SOFAB200_F32) return static_cast<TetFF<float>*>(base)->h.plan.tile_nodes.size();
    return static_cast<TetFF<double>*>(base)->h.plan.tile_nodes.size();
}
size_t tet_nodes(sofab200_tetfem* ff) { return ff->n_nodes; }
int tet_partial_count(sofab200_tetfem* base) {
    if (base->kind == 1) return fast_partial_count(base);
    if (base->real == SOFAB200_F32) { auto& ff = *static_cast<TetFF<float>*>(base); return ff.h.plan.n_tiles + ff.h.plan.n_chunks; }
    auto& ff = *static_cast<TetFF<double>*>(base); return ff.h.plan.n_tiles + ff.h.plan.n_chunks;
}

template <class R> static int tet_create(sofab200_ctx* ctx, size_t n_nodes, const void* rest, size_t n_tets, const uint32_t* tets,
                                         const sofab200_tetfem_desc* desc, sofab200_tetfem** out) {
    std::unique_ptr<TetFF<R>> ff(new TetFF<R>());
    ff->ctx = ctx; ff->real = sizeof(R) == 4 ? SOFAB200_F32 : SOFAB200_F64; ff->method = desc->method;
    ff->n_nodes = n_nodes; ff->n_tets = n_tets;
    const std::string err = tet_host_build(ff->h, n_nodes, static_cast<const R*>(rest), n_tets, tets, desc, kGatherChunk, ctx->sm_count);
    if (!err.empty()) return fail(SOFAB200_ERR_INVALID, err);
    ff->threads = (ff->h.plan.tile_e >= 2048 && sizeof(R) == 4) ? 512 : 256;
    if (const char* env = getenv("SOFAB200_TILE_THREADS")) { const int v = atoi(env); if (v >= 64 && v <= 1024 && v % 32 == 0) ff->threads = v; }
    if (const char* env = getenv("SOFAB200_PREFETCH")) ff->prefetch = atoi(env) != 0;
    ff->von_mises = desc->compute_von_mises;
    ff->sibling = desc->tetrahedral_corotational != 0;
    ff->update_j = desc->update_stiffness_matrix != 0 && desc->method != SOFAB200_TET_SMALL;      // (accumulateForceSmall never reads the flag)
    ff->split_j = ff->update_j && desc->method == SOFAB200_TET_LARGE && !desc->tetrahedral_corotational;
    ff->plastic[0] = desc->plastic_max_threshold; ff->plastic[1] = desc->plastic_yield_threshold; ff->plastic[2] = desc->plastic_creep;
    SB_TRY(tet_upload(*ff));
    *out = ff.release();
    return SOFAB200_OK;
}

template <class R> static int tet_get(TetFF<R>& ff, const std::string& what, void* out) {
    auto cp = [&](const std::vector<R>& v) { std::memcpy(out, v.data(), v.size() * sizeof(R)); return SOFAB200_OK; };
    if (what == "initialRotations") return cp(ff.h.h_R0t);
    if (what == "strainDisplacements") return cp(ff.h.h_J);
    if (what == "materialsStiffnesses") return cp(ff.h.h_K);
    if (what == "rotatedInitialElements") return cp(ff.h.h_X0);
    if (what == "initialTransformation") return cp(ff.h.h_A0inv);
    if (what == "rotations") {
        if (ff.method == SOFAB200_TET_SMALL) return fail(SOFAB200_ERR_UNSUPPORTED, "method small keeps no rotations");
        SB_TRY(ff.rot_export.alloc(9 * ff.n_tets));
        const size_t NS = size_t(ff.h.plan.n_tiles) * ff.h.plan.tile_e;
        tet_export_rotations_kernel<R><<<unsigned((NS + 255) / 256), 256, 0, ff.ctx->stream>>>(ff.dev(), ff.orig.p, ff.rot_export.p);
        ff.ctx->launches++;
        SB_CUDA(cudaGetLastError());
        SB_CUDA(cudaMemcpyAsync(out, ff.rot_export.p, 9 * ff.n_tets * sizeof(R), cudaMemcpyDeviceToHost, ff.ctx->stream));
        SB_CUDA(cudaStreamSynchronize(ff.ctx->stream));
        return SOFAB200_OK;
    }
    if (what == "plasticStrains") {
        if (!ff.pl0.p) return fail(SOFAB200_ERR_UNSUPPORTED, "no plastic strains: plasticMaxThreshold <= 0");
        const size_t NS = size_t(ff.h.plan.n_tiles) * ff.h.plan.tile_e;
        std::vector<Quad<R>> a(NS), b(NS);
        SB_CUDA(cudaMemcpyAsync(a.data(), ff.pl0.p, NS * sizeof(Quad<R>), cudaMemcpyDeviceToHost, ff.ctx->stream));
        SB_CUDA(cudaMemcpyAsync(b.data(), ff.pl1.p, NS * sizeof(Quad<R>), cudaMemcpyDeviceToHost, ff.ctx->stream));
        SB_CUDA(cudaStreamSynchronize(ff.ctx->stream));
        R* o = static_cast<R*>(out);
        for (size_t es = 0; es < NS; ++es) {
            const uint32_t e = ff.h.plan.order[es];
            if (e == 0xFFFFFFFFu) continue;
            R* q = o + 6 * size_t(e);
            q[0] = a[es].a; q[1] = a[es].b; q[2] = a[es].c; q[3] = a[es].d; q[4] = b[es].a; q[5] = b[es].b;
        }
        return SOFAB200_OK;
    }
    return fail(SOFAB200_ERR_INVALID, "unknown array name: " + what);
}

template <class R> static int tet_build_incidence(TetFF<R>& ff);
// getRotations(VecReal&): 9 Reals per node into a device array
template <class R> static int tet_node_rotations(TetFF<R>& ff, R* out_dev) {
    cudaStream_t s = ff.ctx->stream;
    if (ff.method == SOFAB200_TET_SMALL && !ff.sibling) {   // :783-791: identity (and a warning) when no rotation is computed
        std::vector<R> id(9 * ff.n_nodes, R(0));
        for (size_t n = 0; n < ff.n_nodes; ++n) id[9 * n] = id[9 * n + 4] = id[9 * n + 8] = R(1);
        SB_CUDA(cudaMemcpyAsync(out_dev, id.data(), id.size() * sizeof(R), cudaMemcpyHostToDevice, s));
        SB_CUDA(cudaStreamSynchronize(s));
        return SOFAB200_OK;
    }
    SB_TRY(tet_build_incidence(ff));
    int sibling = 0;
    if (ff.sibling) {
        if (ff.method == SOFAB200_TET_SMALL) return fail(SOFAB200_ERR_UNSUPPORTED, "TetrahedralCorotationalFEMForceField keeps no rotations with method small");
        sibling = ff.method == SOFAB200_TET_POLAR ? 2 : 1;
        if (sibling == 2 && !ff.a0_el.p) { SB_TRY(ff.a0_el.upload(ff.h.h_A0, s)); SB_CUDA(cudaStreamSynchronize(s)); }
    }
    tet_node_rotations_kernel<R><<<unsigned((ff.n_nodes + 127) / 128), 128, 0, s>>>(ff.dev(), ff.inc_off.p, ff.inc_es.p, ff.inc_e.p, sibling == 2 ? ff.a0_el.p : ff.r0t_el.p, ff.es_of_first,
                                                                                      out_dev, sibling);
    ff.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}
// node -> incident elements in ascending element index (TetrahedraAroundVertex order), built at the first use
template <class R> static int tet_build_incidence(TetFF<R>& ff) {
    cudaStream_t s = ff.ctx->stream;
    if (!ff.inc_off.p) {
        const HostPlan& P = ff.h.plan;
        const size_t NS = size_t(P.n_tiles) * P.tile_e, T = ff.n_tets;
        std::vector<uint32_t> es_of(T, 0), node_of(4 * T);
        for (size_t es = 0; es < NS; ++es) if (P.order[es] != 0xFFFFFFFFu) es_of[P.order[es]] = uint32_t(es);
        // the corner nodes of element e in original order: local ids of its tile -> node ids
        std::vector<uint32_t> off(ff.n_nodes + 1, 0);
        for (size_t es = 0; es < NS; ++es) {
            const uint32_t e = P.order[es];
            if (e == 0xFFFFFFFFu) continue;
            const size_t tile = es / size_t(P.tile_e);
            for (int k = 0; k < 4; ++k) { const uint32_t g = P.tile_nodes[P.tile_node_off[tile] + P.lnode[4 * es + k]]; node_of[4 * size_t(e) + k] = g; ++off[g + 1]; }
        }
        for (size_t n = 0; n < ff.n_nodes; ++n) off[n + 1] += off[n];
        std::vector<uint32_t> fill(off.begin(), off.end() - 1), ies(4 * T), ie(4 * T);
        for (size_t e = 0; e < T; ++e) for (int k = 0; k < 4; ++k) { const uint32_t at = fill[node_of[4 * e + k]]++; ies[at] = es_of[e]; ie[at] = uint32_t(e); }   // ascending element index
        ff.es_of_first = T ? es_of[0] : 0;
        SB_TRY(ff.inc_off.upload(off, s)); SB_TRY(ff.inc_es.upload(ies, s)); SB_TRY(ff.inc_e.upload(ie, s)); SB_TRY(ff.r0t_el.upload(ff.h.h_R0t, s));
        SB_CUDA(cudaStreamSynchronize(s));
    }
    return SOFAB200_OK;
}
// computeVonMisesStress(): per element (original order) and per node
template <class R> static int tet_von_mises(TetFF<R>& ff, const R* x, R* per_element, R* per_node) {
    if (!ff.von_mises) return fail(SOFAB200_ERR_UNSUPPORTED, "computeVonMisesStress was 0 when the force field was created");
    if (ff.von_mises == 1 && ff.method == SOFAB200_TET_SMALL) return fail(SOFAB200_ERR_UNSUPPORTED, "computeVonMisesStress=1 needs a corotational method (the reference reads rotations it does not have with method small)");
    cudaStream_t s = ff.ctx->stream;
    if (!ff.vm_shf.p) {
        SB_TRY(ff.vm_shf.upload(ff.h.h_shf, s)); SB_TRY(ff.vm_lambda.upload(ff.h.h_lambda, s)); SB_TRY(ff.vm_mu.upload(ff.h.h_mu, s)); SB_TRY(ff.vm_rest.upload(ff.h.h_rest, s));
        SB_TRY(ff.vm_elem.alloc(ff.n_tets));
        SB_CUDA(cudaStreamSynchronize(s));
    }
    R* vme = per_element ? per_element : ff.vm_elem.p;
    const size_t NS = size_t(ff.h.plan.n_tiles) * ff.h.plan.tile_e;
    if (NS) tet_von_mises_kernel<R><<<unsigned((NS + 127) / 128), 128, 0, s>>>(ff.dev(), ff.orig.p, x, ff.vm_rest.p, ff.vm_shf.p, ff.vm_lambda.p, ff.vm_mu.p, ff.von_mises,
                                                                             ff.method == SOFAB200_TET_LARGE ? 1 : 0, vme);
    ff.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    if (per_node && ff.n_nodes) {
        SB_TRY(tet_build_incidence(ff));
        tet_von_mises_nodes_kernel<R><<<unsigned((ff.n_nodes + 127) / 128), 128, 0, s>>>(ff.n_nodes, ff.inc_off.p, ff.inc_e.p, vme, per_node);
        ff.ctx->launches++;
        SB_CUDA(cudaGetLastError());
    }
    return SOFAB200_OK;
}

}  // namespace sb

extern "C" {

int sofab200_tetfem_create(sofab200_ctx* ctx, sofab200_real real, size_t n_nodes, const void* rest_position_host, size_t n_tets,
                           const uint32_t* tets_host, const sofab200_tetfem_desc* desc, sofab200_tetfem** out) {
    SB_CHECK(ctx && out && desc && rest_position_host && (tets_host || n_tets == 0), "null argument");
    if (desc->fast_corotational) {
        SB_CHECK(desc->n_young > 0 && desc->young && desc->n_poisson > 0 && desc->poisson, "youngModulus / poissonRatio are required");
        SB_CHECK(n_nodes < 0xFFFFFFFFull && n_tets < 0x0FFFFFFFull, "mesh too large for 32-bit indices");
        SB_CUDA(cudaSetDevice(ctx->device));
        return fast_create(ctx, int(real), n_nodes, rest_position_host, n_tets, tets_host, desc, out);
    }
    SB_CHECK(desc->method >= 0 && desc->method <= 3, "method must be small, large, polar or svd");
    if (desc->tetrahedral_corotational) {
        SB_CHECK(desc->method != SOFAB200_TET_SVD, "TetrahedralCorotationalFEMForceField has no svd method");
        if (desc->plastic_max_threshold > 0 || desc->compute_von_mises)
            return fail(SOFAB200_ERR_UNSUPPORTED, "plasticity and computeVonMisesStress belong to TetrahedronFEMForceField; TetrahedralCorotationalFEMForceField's own von Mises routine is not provided");
    }
    SB_CHECK(desc->compute_von_mises >= 0 && desc->compute_von_mises <= 2, "computeVonMisesStress must be 0, 1 or 2");
    SB_CHECK(desc->n_young > 0 && desc->young && desc->n_poisson > 0 && desc->poisson, "youngModulus / poissonRatio are required");
    SB_CHECK(n_nodes < 0xFFFFFFFFull && n_tets < 0x3FFFFFFFull, "mesh too large for 32-bit indices");
    SB_CUDA(cudaSetDevice(ctx->device));
    if (real == SOFAB200_F32) return tet_create<float>(ctx, n_nodes, rest_position_host, n_tets, tets_host, desc, out);
    return tet_create<double>(ctx, n_nodes, rest_position_host, n_tets, tets_host, desc, out);
}
int sofab200_tetfem_destroy(sofab200_tetfem* ff) { delete ff; return SOFAB200_OK; }

int sofab200_tetfem_add_force(sofab200_tetfem* ff, void* f_dev, const void* x_dev) {
    SB_CHECK(ff && f_dev && x_dev, "null argument");
    if (ff->real == SOFAB200_F32) {
        NodeEpilogue<float> ep{}; ep.init_src = static_cast<float*>(f_dev); ep.out = static_cast<float*>(f_dev); ep.sign = +1;
        return tet_run<float>(ff, false, static_cast<const float*>(x_dev), 0.f, ep, false);
    }
    NodeEpilogue<double> ep{}; ep.init_src = static_cast<double*>(f_dev); ep.out = static_cast<double*>(f_dev); ep.sign = +1;
    return tet_run<double>(ff, false, static_cast<const double*>(x_dev), 0.0, ep, false);
}
int sofab200_tetfem_add_dforce(sofab200_tetfem* ff, void* df_dev, const void* dx_dev, double k_factor) {
    SB_CHECK(ff && df_dev && dx_dev, "null argument");
    SB_CHECK(df_dev != dx_dev, "df and dx must be distinct vectors");
    if (ff->real == SOFAB200_F32) {
        NodeEpilogue<float> ep{}; ep.init_src = static_cast<float*>(df_dev); ep.out = static_cast<float*>(df_dev); ep.sign = -1;
        return tet_run<float>(ff, true, static_cast<const float*>(dx_dev), float(k_factor), ep, false);
    }
    NodeEpilogue<double> ep{}; ep.init_src = static_cast<double*>(df_dev); ep.out = static_cast<double*>(df_dev); ep.sign = -1;
    return tet_run<double>(ff, true, static_cast<const double*>(dx_dev), k_factor, ep, false);
}
int sofab200_tetfem_get(sofab200_tetfem* ff, const char* what, void* out_host) {
    SB_CHECK(ff && what && out_host, "null argument");
    if (ff->kind == 1) return fast_get(ff, what, out_host);
    if (ff->real == SOFAB200_F32) return tet_get(*static_cast<TetFF<float>*>(ff), what, out_host);
    return tet_get(*static_cast<TetFF<double>*>(ff), what, out_host);
}
int sofab200_tetfem_compute_von_mises(sofab200_tetfem* ff, const void* x_dev, void* per_element_dev, void* per_node_dev) {
    SB_CHECK(ff && x_dev, "null argument");
    if (ff->kind == 1) return fail(SOFAB200_ERR_UNSUPPORTED, "FastTetrahedralCorotationalForceField has no computeVonMisesStress");
    if (ff->real == SOFAB200_F32) return tet_von_mises(*static_cast<TetFF<float>*>(ff), static_cast<const float*>(x_dev), static_cast<float*>(per_element_dev), static_cast<float*>(per_node_dev));
    return tet_von_mises(*static_cast<TetFF<double>*>(ff), static_cast<const double*>(x_dev), static_cast<double*>(per_element_dev), static_cast<double*>(per_node_dev));
}
int sofab200_tetfem_reset(sofab200_tetfem* ff) {
    SB_CHECK(ff, "null argument");
    if (ff->kind == 1) return SOFAB200_OK;      // (no plastic strain to clear)
    if (ff->real == SOFAB200_F32) { auto* f = static_cast<TetFF<float>*>(ff); SB_TRY(f->pl0.zero(ff->ctx->stream)); SB_TRY(f->pl1.zero(ff->ctx->stream)); }
    else { auto* f = static_cast<TetFF<double>*>(ff); SB_TRY(f->pl0.zero(ff->ctx->stream)); SB_TRY(f->pl1.zero(ff->ctx->stream)); }
    return SOFAB200_OK;
}
int sofab200_tetfem_get_rotations(sofab200_tetfem* ff, void* vecR_dev) {
    SB_CHECK(ff && vecR_dev, "null argument");
    if (ff->kind == 1) return fail(SOFAB200_ERR_UNSUPPORTED, "FastTetrahedralCorotationalForceField has no getRotations");
    if (ff->real == SOFAB200_F32) return tet_node_rotations(*static_cast<TetFF<float>*>(ff), static_cast<float*>(vecR_dev));
    return tet_node_rotations(*static_cast<TetFF<double>*>(ff), static_cast<double*>(vecR_dev));
}
int sofab200_tetfem_stats(const sofab200_tetfem* ff, uint64_t out[8]) {
    SB_CHECK(ff && out, "null argument");
    if (ff->kind == 1) return fast_stats(ff, out);
    const HostPlan* P; size_t smem;
    if (ff->real == SOFAB200_F32) { auto* f = static_cast<const TetFF<float>*>(ff); P = &f->h.plan; smem = f->h.smem_bytes; }
    else { auto* f = static_cast<const TetFF<double>*>(ff); P = &f->h.plan; smem = f->h.smem_bytes; }
    out[0] = P->n_tiles; out[1] = P->tile_e; out[2] = P->n_interior; out[3] = P->n_shared; out[4] = P->n_staged_corners;
    out[5] = smem; out[6] = P->maxval; out[7] = ff->n_tets;
    return SOFAB200_OK;
}

}  // extern "C"
