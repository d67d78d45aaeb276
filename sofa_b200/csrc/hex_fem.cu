// HexahedronFEMForceField<B200Vec3Types>: device upload + launches + C ABI.
#include <cstring>
#include <memory>

#include "hex_host.h"

using namespace sb;

struct sofab200_hexfem {
    virtual ~sofab200_hexfem() {}
    sofab200_ctx* ctx = nullptr;
    int real = 0, method = 0;
    size_t n_nodes = 0, n_hexas = 0;
};

namespace sb {

template <class R> struct HexFF : sofab200_hexfem {
    HostHex<R> h;
    DevBuf<uint4> lnode, slot_a, slot_b;
    DevBuf<uint32_t> orig, kidx, tile_kuniq;
    bool coop = true;      // addDForce: eight lanes per hexahedron
    DevBuf<Quad<R>> r0, r1, r2, x0;
    DevBuf<R> ktab;
    DevBuf<uint32_t> tile_node_off, tile_nodes, tile_shslot, tile_nint, sh_nodes, sh_base;
    DevBuf<uint16_t> tile_val, tile_jds, sh_val;
    DevBuf<Quad<R>> stage;
    DevBuf<R> rot_export;
    DevBuf<uint32_t> inc_off, inc_es, inc_e; DevBuf<R> rot0_el;   // getRotations: node -> incident elements (built at the first call)
    size_t n_unique = 0;
    HexDev<R> dev() {
        const HostPlan& plan = h.plan;
        HexDev<R> d;
        d.t.n_nodes = int(n_nodes); d.t.n_elems = int(n_hexas); d.t.n_tiles = plan.n_tiles; d.t.tile_e = plan.tile_e; d.t.maxval = plan.maxval;
        d.t.tile_node_off = tile_node_off.p; d.t.tile_nodes = tile_nodes.p; d.t.tile_shslot = tile_shslot.p; d.t.tile_nint = tile_nint.p; d.t.tile_nb = nullptr; d.t.tile_val = tile_val.p; d.t.tile_jds = tile_jds.p;
        d.t.n_shared = plan.n_shared; d.t.n_chunks = plan.n_chunks; d.t.sh_nodes = sh_nodes.p; d.t.sh_val = sh_val.p; d.t.sh_base = sh_base.p;
        d.t.stage = stage.p; d.t.stage_n = plan.stage_n;
        d.lnode = lnode.p; d.slot_a = slot_a.p; d.slot_b = slot_b.p; d.r0 = r0.p; d.r1 = r1.p; d.r2 = r2.p;
        d.kidx = kidx.p; d.ktab = ktab.p; d.tile_kuniq = tile_kuniq.p; d.x0 = x0.p; d.n_slots = size_t(plan.n_tiles) * plan.tile_e;
        d.k_factor = R(0);
        return d;
    }
};

template <class R> static int hex_upload(HexFF<R>& ff) {
    HostHex<R>& H = ff.h;
    const HostPlan& P = H.plan;
    cudaStream_t s = ff.ctx->stream;
    SB_TRY(ff.lnode.upload(H.lnode, s)); SB_TRY(ff.slot_a.upload(H.slot_a, s)); SB_TRY(ff.slot_b.upload(H.slot_b, s)); SB_TRY(ff.orig.upload(P.order, s));
    SB_TRY(ff.r0.upload(H.r0, s)); SB_TRY(ff.r1.upload(H.r1, s)); SB_TRY(ff.r2.upload(H.r2, s)); SB_TRY(ff.x0.upload(H.x0, s));
    SB_TRY(ff.kidx.upload(H.kidx, s)); SB_TRY(ff.tile_kuniq.upload(H.tile_kuniq, s)); SB_TRY(ff.ktab.upload(H.ktab, s));
    SB_TRY(ff.tile_node_off.upload(P.tile_node_off, s)); SB_TRY(ff.tile_nodes.upload(P.tile_nodes, s)); SB_TRY(ff.tile_shslot.upload(P.tile_shslot, s)); SB_TRY(ff.tile_nint.upload(P.tile_nint, s));
    SB_TRY(ff.tile_val.upload(P.tile_val, s)); SB_TRY(ff.tile_jds.upload(P.tile_jds, s));
    SB_TRY(ff.sh_nodes.upload(P.sh_nodes, s)); SB_TRY(ff.sh_val.upload(P.sh_val, s)); SB_TRY(ff.sh_base.upload(P.sh_base, s));
    SB_TRY(ff.stage.alloc(P.stage_n)); SB_TRY(ff.stage.zero(s));
    SB_CUDA(cudaStreamSynchronize(s));
    ff.n_unique = H.ktab.size() / 576;
    for (auto* v : {&H.r0, &H.r1, &H.r2, &H.x0}) { v->clear(); v->shrink_to_fit(); }
    H.lnode.clear(); H.slot_a.clear(); H.slot_b.clear(); H.kidx.clear();
    return SOFAB200_OK;
}

// addDForce with eight lanes per hexahedron (hex_tile_df_coop_kernel); SOFAB200_HEX_COOP=0 selects the one-thread-per-hexahedron pass
template <class R> static int hex_launch_df_coop(HexFF<R>& ff, const HexDev<R>& d, const R* in, const NodeEpilogue<R>& ep) {
    auto kern = hex_tile_df_coop_kernel<R>;
    if (ff.h.smem_bytes > 48 * 1024) SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ff.h.smem_bytes)));
    ff.ctx->prof_start(0);
    kern<<<ff.h.plan.n_tiles, 256, ff.h.smem_bytes, ff.ctx->stream>>>(d, in, ep, ff.h.plan.max_touched, ff.h.plan.max_slots);
    ff.ctx->prof_stop(0);
    ff.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}
template <class R, int MODE> static int hex_launch_mode(HexFF<R>& ff, const HexDev<R>& d, const R* in, const NodeEpilogue<R>& ep) {
    if (MODE == HM_DF && ff.coop && sizeof(R) == 4) return hex_launch_df_coop<R>(ff, d, in, ep);
    auto kern = hex_tile_kernel<R, MODE>;
    // (per device and context, not per thread: set for the current device on every launch)
    if (ff.h.smem_bytes > 48 * 1024) SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ff.h.smem_bytes)));
    const int cls = MODE == HM_DF ? 0 : 2;
    ff.ctx->prof_start(cls);
    kern<<<ff.h.plan.n_tiles, 256, ff.h.smem_bytes, ff.ctx->stream>>>(d, in, ep, ff.h.plan.max_touched, ff.h.plan.max_slots);
    ff.ctx->prof_stop(cls);
    ff.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}

// persistent CG kernel (see tet_fem.cu: tet_persist_variant); hexahedral tiles, 256 threads
template <class R> int hex_cg_persistent(sofab200_hexfem* base, R k_factor, PersistCG<R> a, size_t sync_capacity, bool dry_run) {
    HexFF<R>& ff = *static_cast<HexFF<R>*>(base);
    HexDev<R> d = ff.dev();
    d.k_factor = k_factor;
    auto kern = hex_cg_persistent_kernel<R>;
    const HostPlan& P = ff.h.plan;
    const int threads = 256, groups = threads / kGatherChunk;
    cudaFuncAttributes fa;
    SB_CUDA(cudaFuncGetAttributes(&fa, kern));
    int dev_smem_optin = 0;
    SB_CUDA(cudaDeviceGetAttribute(&dev_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ff.ctx->device));
    for (int tiles_per_cta = 1; tiles_per_cta <= 2; ++tiles_per_cta) {
        const PersistLayout L = persist_layout<R>(tiles_per_cta, P.max_touched, P.max_slots, P.max_int, P.max_shtouch, P.maxval, sizeof(R) * 576 * kHexSmemMatrices);
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
// fused CG kernel (cg_fused.cuh) around the hexahedral element pass: cached when the CTA's tiles fit, else streamed (any number of tiles)
template <class R, bool CACHED> static int hex_fused_launch(HexFF<R>& ff, HexDev<R> d, FusedCG<R> a, const FusedLayout& L, int grid, bool dry_run, int* info) {
    constexpr int ET = sizeof(R) == 4 ? 512 : 256;     // eight lanes per hexahedron at <= 128 registers (Vec3f); Vec3d needs 255
    auto kern = fused_cg_kernel<R, HexPass<R>, ET, 0, CACHED>;
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
    a.lay = L;
    int ded_share = 0;
    void* args[] = {&d, &a, &ded_share};
    ff.ctx->prof_start(4);
    SB_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(ET), args, L.total, ff.ctx->stream));
    ff.ctx->prof_stop(4);
    ff.ctx->launches++;
    return SOFAB200_OK;
}
template <class R> int hex_cg_fused(sofab200_hexfem* base, R k_factor, FusedCG<R> a, size_t sync_capacity, bool dry_run, int* info) {
    HexFF<R>& ff = *static_cast<HexFF<R>*>(base);
    HexDev<R> d = ff.dev();
    d.k_factor = k_factor;
    const HostPlan& P = ff.h.plan;
    const int grid = std::max(1, std::min(ff.ctx->sm_count, P.n_tiles));
    const int tiles_per_cta = (P.n_tiles + grid - 1) / grid;
    const int n_units = P.n_chunks * (kGatherChunk / kUnit);
    const int units_per_cta = (n_units + grid - 1) / grid;
    if (fused_sync_words(grid) > sync_capacity) return fail(SOFAB200_ERR_INVALID, "sync buffer too small for the fused CG kernel");
    const size_t extra = sizeof(R) * kHexKPadded * kHexSmemMatrices;
    const FusedLayout Lc = fused_layout<R>(true, tiles_per_cta, units_per_cta, P.max_touched, P.max_slots, P.max_int, P.max_shtouch, P.maxval, extra);
    if (tiles_per_cta <= 2 && Lc.total + 2048 <= 200 * 1024) {
        const int rc = hex_fused_launch<R, true>(ff, d, a, Lc, grid, dry_run, info);
        if (rc != kPersistNotEligible) return rc;
    }
    // streamed tiles: measured on C3 (491 520 hexahedra, 4 tiles per CTA) the one-thread-per-hexahedron pass at 255 registers runs slower inside the
    // persistent kernel (3.44 ms per step; 4.62 ms with the eight-lane pass, which spills next to the kernel's other phases) than as plain tile launches
    // (2.30 ms), so big hexahedral meshes keep the multi-kernel loop unless asked (SOFAB200_HEX_FUSED_STREAMED=1)
    { const char* env = getenv("SOFAB200_HEX_FUSED_STREAMED"); if (!env || atoi(env) == 0) return kPersistNotEligible; }
    const FusedLayout Ls = fused_layout<R>(false, tiles_per_cta, units_per_cta, P.max_touched, P.max_slots, P.max_int, P.max_shtouch, P.maxval, extra);
    return hex_fused_launch<R, false>(ff, d, a, Ls, grid, dry_run, info);
}
template int hex_cg_fused<float>(sofab200_hexfem*, float, FusedCG<float>, size_t, bool, int*);
template int hex_cg_fused<double>(sofab200_hexfem*, double, FusedCG<double>, size_t, bool, int*);
size_t hex_shared_slot_count(sofab200_hexfem* base) {
    if (base->real == SOFAB200_F32) return size_t(static_cast<HexFF<float>*>(base)->h.plan.n_chunks) * kGatherChunk;
    return size_t(static_cast<HexFF<double>*>(base)->h.plan.n_chunks) * kGatherChunk;
}
template int hex_cg_persistent<float>(sofab200_hexfem*, float, PersistCG<float>, size_t, bool);
template int hex_cg_persistent<double>(sofab200_hexfem*, double, PersistCG<double>, size_t, bool);
const std::vector<uint32_t>& hex_shared_node_table(sofab200_hexfem* base) {
    if (base->real == SOFAB200_F32) return static_cast<HexFF<float>*>(base)->h.plan.sh_nodes;
    return static_cast<HexFF<double>*>(base)->h.plan.sh_nodes;
}
size_t hex_tile_node_count(sofab200_hexfem* base) {
    if (base->real == SOFAB200_F32) return static_cast<HexFF<float>*>(base)->h.plan.tile_nodes.size();
    return static_cast<HexFF<double>*>(base)->h.plan.tile_nodes.size();
}

template <class R> TileDev<R> hex_tiledev(sofab200_hexfem* base) { return static_cast<HexFF<R>*>(base)->dev().t; }
template TileDev<float> hex_tiledev<float>(sofab200_hexfem*);
template TileDev<double> hex_tiledev<double>(sofab200_hexfem*);

// skip_gather: the caller sums the shared nodes itself (fused CG tail kernel)
template <class R> int hex_run(sofab200_hexfem* base, bool dforce, const R* in, R k_factor, NodeEpilogue<R> ep, bool skip_gather) {
    HexFF<R>& ff = *static_cast<HexFF<R>*>(base);
    const HostPlan& plan = ff.h.plan;
    SB_CHECK((!ep.mdx_src || ep.mdx_src == in) && (!ep.dot_with || ep.dot_with == in), "mass / dot operands must be the pass's input vector");
    HexDev<R> d = ff.dev();
    d.k_factor = k_factor;
    ep.partial_base = 0;
    ep.partial_total = plan.n_tiles + plan.n_chunks;
    if (dforce) SB_TRY((hex_launch_mode<R, HM_DF>(ff, d, in, ep)));
    else if (ff.method == SOFAB200_HEX_SMALL) SB_TRY((hex_launch_mode<R, HM_F_SMALL>(ff, d, in, ep)));
    else if (ff.method == SOFAB200_HEX_LARGE) SB_TRY((hex_launch_mode<R, HM_F_LARGE>(ff, d, in, ep)));
    else SB_TRY((hex_launch_mode<R, HM_F_POLAR>(ff, d, in, ep)));
    if (skip_gather) return SOFAB200_OK;
    ep.partial_base = plan.n_tiles;
    ff.ctx->prof_start(1);
    gather_shared_kernel<R><<<plan.n_chunks, kGatherChunk, 0, ff.ctx->stream>>>(d.t, ep);
    ff.ctx->prof_stop(1);
    ff.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}
template int hex_run<float>(sofab200_hexfem*, bool, const float*, float, NodeEpilogue<float>, bool);
template int hex_run<double>(sofab200_hexfem*, bool, const double*, double, NodeEpilogue<double>, bool);

int hex_real(sofab200_hexfem* ff) { return ff->real; }
size_t hex_nodes(sofab200_hexfem* ff) { return ff->n_nodes; }
int hex_partial_count(sofab200_hexfem* base) {
    if (base->real == SOFAB200_F32) { auto& ff = *static_cast<HexFF<float>*>(base); return ff.h.plan.n_tiles + ff.h.plan.n_chunks; }
    auto& ff = *static_cast<HexFF<double>*>(base); return ff.h.plan.n_tiles + ff.h.plan.n_chunks;
}

template <class R> static int hex_create(sofab200_ctx* ctx, size_t n_nodes, const void* rest, size_t n_hexas, const uint32_t* hexas,
                                         const sofab200_hexfem_desc* desc, sofab200_hexfem** out) {
    std::unique_ptr<HexFF<R>> ff(new HexFF<R>());
    ff->ctx = ctx; ff->real = sizeof(R) == 4 ? SOFAB200_F32 : SOFAB200_F64; ff->method = desc->method;
    if (const char* env = getenv("SOFAB200_HEX_COOP")) ff->coop = atoi(env) != 0;
    ff->n_nodes = n_nodes; ff->n_hexas = n_hexas;
    const std::string err = hex_host_build(ff->h, n_nodes, static_cast<const R*>(rest), n_hexas, hexas, desc, kGatherChunk, ctx->sm_count);
    if (!err.empty()) return fail(SOFAB200_ERR_INVALID, err);
    SB_TRY(hex_upload(*ff));
    *out = ff.release();
    return SOFAB200_OK;
}

template <class R> static int hex_get(HexFF<R>& ff, const std::string& what, void* out) {
    if (what == "rotatedInitialElements") { std::memcpy(out, ff.h.h_X0.data(), ff.h.h_X0.size() * sizeof(R)); return SOFAB200_OK; }
    if (what == "elementStiffnesses") {
        R* o = static_cast<R*>(out);
        for (size_t e = 0; e < ff.n_hexas; ++e) std::memcpy(o + 576 * e, ff.h.ktab.data() + 576 * size_t(ff.h.h_kidx[e]), 576 * sizeof(R));
        return SOFAB200_OK;
    }
    if (what == "rotations") {
        SB_TRY(ff.rot_export.alloc(9 * ff.n_hexas));
        const HexDev<R> d = ff.dev();
        hex_export_rotations_kernel<R><<<unsigned((d.n_slots + 255) / 256), 256, 0, ff.ctx->stream>>>(d, ff.orig.p, ff.rot_export.p);
        ff.ctx->launches++;
        SB_CUDA(cudaGetLastError());
        SB_CUDA(cudaMemcpyAsync(out, ff.rot_export.p, 9 * ff.n_hexas * sizeof(R), cudaMemcpyDeviceToHost, ff.ctx->stream));
        SB_CUDA(cudaStreamSynchronize(ff.ctx->stream));
        return SOFAB200_OK;
    }
    return fail(SOFAB200_ERR_INVALID, "unknown array name: " + what);
}

// getRotations: getNodeRotation for every node, 9 Reals per node into a device array
template <class R> static int hex_node_rotations(HexFF<R>& ff, R* out_dev) {
    cudaStream_t s = ff.ctx->stream;
    if (!ff.inc_off.p) {
        const HostPlan& P = ff.h.plan;
        const size_t NS = size_t(P.n_tiles) * P.tile_e, H = ff.n_hexas;
        std::vector<uint32_t> es_of(H, 0), node_of(8 * H), off(ff.n_nodes + 1, 0);
        for (size_t es = 0; es < NS; ++es) {
            const uint32_t e = P.order[es];
            if (e == 0xFFFFFFFFu) continue;
            es_of[e] = uint32_t(es);
            const size_t tile = es / size_t(P.tile_e);
            for (int k = 0; k < 8; ++k) { const uint32_t g = P.tile_nodes[P.tile_node_off[tile] + P.lnode[8 * es + k]]; node_of[8 * size_t(e) + k] = g; ++off[g + 1]; }
        }
        for (size_t n = 0; n < ff.n_nodes; ++n) off[n + 1] += off[n];
        std::vector<uint32_t> fill(off.begin(), off.end() - 1), ies(8 * H), ie(8 * H);
        for (size_t e = 0; e < H; ++e) for (int k = 0; k < 8; ++k) { const uint32_t at = fill[node_of[8 * e + k]]++; ies[at] = es_of[e]; ie[at] = uint32_t(e); }   // ascending element index
        SB_TRY(ff.inc_off.upload(off, s)); SB_TRY(ff.inc_es.upload(ies, s)); SB_TRY(ff.inc_e.upload(ie, s)); SB_TRY(ff.rot0_el.upload(ff.h.h_rot0, s));
        SB_CUDA(cudaStreamSynchronize(s));
    }
    if (!ff.n_nodes) return SOFAB200_OK;
    hex_node_rotations_kernel<R><<<unsigned((ff.n_nodes + 127) / 128), 128, 0, s>>>(ff.dev(), ff.inc_off.p, ff.inc_es.p, ff.inc_e.p, ff.rot0_el.p, out_dev);
    ff.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}

}  // namespace sb

extern "C" {
int sofab200_hexfem_get_rotations(sofab200_hexfem* ff, void* vecR_dev) {
    SB_CHECK(ff && vecR_dev, "null argument");
    if (ff->real == SOFAB200_F32) return hex_node_rotations(*static_cast<HexFF<float>*>(ff), static_cast<float*>(vecR_dev));
    return hex_node_rotations(*static_cast<HexFF<double>*>(ff), static_cast<double*>(vecR_dev));
}


int sofab200_hexfem_create(sofab200_ctx* ctx, sofab200_real real, size_t n_nodes, const void* rest_position_host, size_t n_hexas,
                           const uint32_t* hexas_host, const sofab200_hexfem_desc* desc, sofab200_hexfem** out) {
    SB_CHECK(ctx && out && desc && rest_position_host && (hexas_host || n_hexas == 0), "null argument");
    SB_CHECK(desc->method >= 0 && desc->method <= 2, "method must be large, polar or small");
    SB_CHECK(desc->n_young > 0 && desc->young && desc->n_poisson > 0 && desc->poisson, "youngModulus / poissonRatio are required");
    SB_CHECK(n_nodes < 0xFFFFFFFFull && n_hexas < 0x1FFFFFFFull, "mesh too large for 32-bit indices");
    SB_CUDA(cudaSetDevice(ctx->device));
    if (real == SOFAB200_F32) return hex_create<float>(ctx, n_nodes, rest_position_host, n_hexas, hexas_host, desc, out);
    return hex_create<double>(ctx, n_nodes, rest_position_host, n_hexas, hexas_host, desc, out);
}
int sofab200_hexfem_destroy(sofab200_hexfem* ff) { delete ff; return SOFAB200_OK; }
int sofab200_hexfem_add_force(sofab200_hexfem* ff, void* f_dev, const void* x_dev) {
    SB_CHECK(ff && f_dev && x_dev, "null argument");
    if (ff->real == SOFAB200_F32) {
        NodeEpilogue<float> ep{}; ep.init_src = static_cast<float*>(f_dev); ep.out = static_cast<float*>(f_dev); ep.sign = +1;
        return hex_run<float>(ff, false, static_cast<const float*>(x_dev), 0.f, ep, false);
    }
    NodeEpilogue<double> ep{}; ep.init_src = static_cast<double*>(f_dev); ep.out = static_cast<double*>(f_dev); ep.sign = +1;
    return hex_run<double>(ff, false, static_cast<const double*>(x_dev), 0.0, ep, false);
}
int sofab200_hexfem_add_dforce(sofab200_hexfem* ff, void* df_dev, const void* dx_dev, double k_factor) {
    SB_CHECK(ff && df_dev && dx_dev, "null argument");
    SB_CHECK(df_dev != dx_dev, "df and dx must be distinct vectors");
    if (ff->real == SOFAB200_F32) {
        NodeEpilogue<float> ep{}; ep.init_src = static_cast<float*>(df_dev); ep.out = static_cast<float*>(df_dev); ep.sign = -1;
        return hex_run<float>(ff, true, static_cast<const float*>(dx_dev), float(k_factor), ep, false);
    }
    NodeEpilogue<double> ep{}; ep.init_src = static_cast<double*>(df_dev); ep.out = static_cast<double*>(df_dev); ep.sign = -1;
    return hex_run<double>(ff, true, static_cast<const double*>(dx_dev), k_factor, ep, false);
}
int sofab200_hexfem_get(sofab200_hexfem* ff, const char* what, void* out_host) {
    SB_CHECK(ff && what && out_host, "null argument");
    if (ff->real == SOFAB200_F32) return hex_get(*static_cast<HexFF<float>*>(ff), what, out_host);
    return hex_get(*static_cast<HexFF<double>*>(ff), what, out_host);
}
int sofab200_hexfem_stats(const sofab200_hexfem* ff, uint64_t out[8]) {
    SB_CHECK(ff && out, "null argument");
    const HostPlan* P; size_t smem, uniq;
    if (ff->real == SOFAB200_F32) { auto* f = static_cast<const HexFF<float>*>(ff); P = &f->h.plan; smem = f->h.smem_bytes; uniq = f->n_unique; }
    else { auto* f = static_cast<const HexFF<double>*>(ff); P = &f->h.plan; smem = f->h.smem_bytes; uniq = f->n_unique; }
    out[0] = P->n_tiles; out[1] = P->tile_e; out[2] = P->n_interior; out[3] = P->n_shared; out[4] = P->n_staged_corners;
    out[5] = smem; out[6] = uniq; out[7] = ff->n_hexas;
    return SOFAB200_OK;
}

}  // extern "C"
