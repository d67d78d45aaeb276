// TEST INFRASTRUCTURE: executes the product's tile/gather plan and its host+device (HD) element and
// epilogue functions on the CPU, thread by thread, so that the layout logic (tiles, slots, jagged
// diagonals, summation order) can be checked against the oracle without a GPU (-m "not gpu" tests).
// It is NOT part of libsofa_b200.so and is not reachable from the product API: the product has no CPU path.
#include <memory>

#include "tet_host.h"

using namespace sb;

namespace {
template <class R> struct Emu {
    HostTet<R> h;
    std::vector<Quad<R>> stage;
    TetDev<R> dev() {
        const HostPlan& P = h.plan;
        TetDev<R> d{};
        d.t.n_nodes = int(h.n_nodes); d.t.n_elems = int(h.n_tets); d.t.n_tiles = P.n_tiles; d.t.tile_e = P.tile_e; d.t.maxval = P.maxval;
        d.t.tile_node_off = P.tile_node_off.data(); d.t.tile_nodes = P.tile_nodes.data(); d.t.tile_nint = P.tile_nint.data();
        d.t.tile_val = P.tile_val.data(); d.t.tile_jds = P.tile_jds.data();
        d.t.n_shared = P.n_shared; d.t.n_chunks = P.n_chunks; d.t.sh_nodes = P.sh_nodes.data(); d.t.sh_val = P.sh_val.data();
        d.t.sh_base = P.sh_base.data(); d.t.stage = stage.data(); d.t.stage_n = P.stage_n;
        d.lnode = h.lnode.data(); d.slot = h.slot.data(); d.js0 = nullptr; d.js1 = nullptr; d.js2 = nullptr;
        d.rk0 = h.rk0.data(); d.rk1 = h.rk1.data(); d.rk2 = h.rk2.data(); d.j0 = h.j0.data(); d.j1 = h.j1.data(); d.j2 = h.j2.data();
        d.x0a = h.x0a.data(); d.x0b = h.x0b.data(); d.x0c = h.x0c.data();
        d.sv0 = h.sv[0].data(); d.sv1 = h.sv[1].data(); d.sv2 = h.sv[2].data(); d.sv3 = h.sv[3].data(); d.sv4 = h.sv[4].data();
        return d;
    }
};

template <class R, int MODE> double run_mode(Emu<R>& E, const R* in, R kf, NodeEpilogue<R> ep) {
    TetDev<R> d = E.dev();
    d.k_factor = kf;
    const TileDev<R>& t = d.t;
    const HostPlan& P = E.h.plan;
    double dot = 0.0;
    std::vector<V3<R>> s_in(P.max_touched);
    std::vector<R> s_slot(3 * size_t(P.max_slots));
    for (int tile = 0; tile < t.n_tiles; ++tile) {
        const uint32_t node_off = t.tile_node_off[tile];
        const int n_touched = int(t.tile_node_off[tile + 1] - node_off), n_int = int(t.tile_nint[tile]);
        for (int k = 0; k < n_touched; ++k) { const uint32_t g = t.tile_nodes[node_off + k]; s_in[k] = mk3<R>(in[3 * size_t(g)], in[3 * size_t(g) + 1], in[3 * size_t(g) + 2]); }
        const uint16_t* jds = t.tile_jds + size_t(tile) * (t.maxval + 1);
        std::fill(s_slot.begin(), s_slot.end(), R(12345));  // poison: every slot read must have been written
        for (int le = 0; le < t.tile_e; ++le) {
            const size_t es = size_t(tile) * t.tile_e + le;
            const ushort4 ln = d.lnode[es];
            if (ln.x == 0xFFFFu) continue;
            const uint4 sl = d.slot[es];
            const V3<R> Pn[4] = {s_in[ln.x], s_in[ln.y], s_in[ln.z], s_in[ln.w]};
            V3<R> C[4];
            tet_element<R, MODE>(d, es, tet_load_rec(d, es), Pn, C);
            const unsigned s4[4] = {sl.x, sl.y, sl.z, sl.w};
            for (int n = 0; n < 4; ++n) {
                const unsigned s = s4[n];
                if (s & kStageFlag) stage_store(t.stage + (s & ~kStageFlag), C[n].x, C[n].y, C[n].z, 0);
                else { s_slot[s] = C[n].x; s_slot[P.max_slots + s] = C[n].y; s_slot[2 * size_t(P.max_slots) + s] = C[n].z; }
            }
        }
        for (int k = 0; k < n_int; ++k) {
            const uint32_t g = t.tile_nodes[node_off + k];
            const int val = t.tile_val[node_off + k];
            R ax, ay, az;
            node_pre(ep, g, ax, ay, az);
            node_mass(ep, ep.pre_kind, g, ax, ay, az);
            for (int jj = 0; jj < val; ++jj) {
                const int s = jds[jj] + k;
                if (ep.sign > 0) { ax += s_slot[s]; ay += s_slot[P.max_slots + s]; az += s_slot[2 * size_t(P.max_slots) + s]; }
                else { ax -= s_slot[s]; ay -= s_slot[P.max_slots + s]; az -= s_slot[2 * size_t(P.max_slots) + s]; }
            }
            dot += node_post(ep, g, ax, ay, az);
        }
    }
    for (int chunk = 0; chunk < t.n_chunks; ++chunk) {
        for (int k = 0; k < kGatherChunk; ++k) {
            const uint32_t g = t.sh_nodes[size_t(chunk) * kGatherChunk + k];
            if (g == 0xFFFFFFFFu) continue;
            const int val = t.sh_val[size_t(chunk) * kGatherChunk + k];
            const size_t base = t.sh_base[chunk];
            R ax, ay, az;
            node_pre(ep, g, ax, ay, az);
            node_mass(ep, ep.pre_kind, g, ax, ay, az);
            for (int j = 0; j < val; ++j) {
                const Quad<R> v = stage_load(t.stage + (base + size_t(j) * kGatherChunk + k), 0);
                if (ep.sign > 0) { ax += v.a; ay += v.b; az += v.c; } else { ax -= v.a; ay -= v.b; az -= v.c; }
            }
            dot += node_post(ep, g, ax, ay, az);
        }
    }
    return dot;
}

template <class R> double run(Emu<R>& E, int dforce, const R* in, double kf, NodeEpilogue<R> ep) {
    if (dforce) return E.h.method == SOFAB200_TET_SMALL ? run_mode<R, TM_DF_SMALL>(E, in, R(kf), ep) : run_mode<R, TM_DF_COROT>(E, in, R(kf), ep);
    switch (E.h.method) {
    case SOFAB200_TET_SMALL: return run_mode<R, TM_F_SMALL>(E, in, R(0), ep);
    case SOFAB200_TET_LARGE: return run_mode<R, TM_F_LARGE>(E, in, R(0), ep);
    case SOFAB200_TET_POLAR: return run_mode<R, TM_F_POLAR>(E, in, R(0), ep);
    default: return run_mode<R, TM_F_SVD>(E, in, R(0), ep);
    }
}
struct EmuAny { int real; Emu<float> f; Emu<double> d; std::string err; };
}  // namespace

// flat epilogue description for ctypes
struct EmuEpilogue {
    const void* init_src; void* out; int sign; int pre_kind; int post_kind; const void* mass; const void* mdx_src;
    double mass_factor; double gravity[3]; int has_scale; double scale; const unsigned char* fixed; const void* dot_with;
};

extern "C" {
void* emu_tet_create(int real, size_t n_nodes, const void* rest, size_t n_tets, const uint32_t* tets, int method, size_t ny, const double* young,
                     size_t np, const double* poisson, int tile_e, const unsigned char* shared_nodes) {
    EmuAny* e = new EmuAny(); e->real = real;
    sofab200_tetfem_desc d{}; d.method = method; d.n_young = ny; d.young = young; d.n_poisson = np; d.poisson = poisson; d.tile_elems = tile_e; d.shared_nodes = shared_nodes;
    if (real == 0) { e->err = tet_host_build(e->f.h, n_nodes, (const float*)rest, n_tets, tets, &d, kGatherChunk); e->f.stage.assign(e->f.h.plan.stage_n, Quad<float>{777.f, 777.f, 777.f, 0.f}); }
    else { e->err = tet_host_build(e->d.h, n_nodes, (const double*)rest, n_tets, tets, &d, kGatherChunk); e->d.stage.assign(e->d.h.plan.stage_n, Quad<double>{777.0, 777.0, 777.0, 0.0}); }
    return e;
}
const char* emu_tet_error(void* h) { return static_cast<EmuAny*>(h)->err.c_str(); }
void emu_tet_destroy(void* h) { delete static_cast<EmuAny*>(h); }
void emu_tet_stats(void* h, uint64_t* out) {
    EmuAny* e = static_cast<EmuAny*>(h);
    const HostPlan& P = e->real == 0 ? e->f.h.plan : e->d.h.plan;
    out[0] = P.n_tiles; out[1] = P.tile_e; out[2] = P.n_interior; out[3] = P.n_shared; out[4] = P.n_staged_corners;
    out[5] = e->real == 0 ? e->f.h.smem_bytes : e->d.h.smem_bytes; out[6] = P.maxval; out[7] = P.n_elems;
}
double emu_tet_run(void* h, int dforce, const void* in, double kf, const EmuEpilogue* q) {
    EmuAny* e = static_cast<EmuAny*>(h);
    auto fill = [&](auto& ep, auto tag) {
        typedef decltype(tag) R;
        ep.init_src = (const R*)q->init_src; ep.out = (R*)q->out; ep.sign = q->sign; ep.pre_kind = q->pre_kind; ep.post_kind = q->post_kind;
        ep.mass = (const R*)q->mass; ep.mdx_src = (const R*)q->mdx_src; ep.mass_factor = R(q->mass_factor); ep.mass_factor_is_one = q->mass_factor == 1.0;
        ep.gx = R(q->gravity[0]); ep.gy = R(q->gravity[1]); ep.gz = R(q->gravity[2]); ep.has_scale = q->has_scale; ep.scale = R(q->scale);
        ep.fixed = q->fixed; ep.dot_kind = q->dot_with ? DOT_STORE : DOT_NONE; ep.dot_with = (const R*)q->dot_with;
    };
    if (e->real == 0) { NodeEpilogue<float> ep{}; fill(ep, float()); return run<float>(e->f, dforce, (const float*)in, kf, ep); }
    NodeEpilogue<double> ep{}; fill(ep, double()); return run<double>(e->d, dforce, (const double*)in, kf, ep);
}
// rotations[e] in original element order
void emu_tet_rotations(void* h, void* out) {
    EmuAny* e = static_cast<EmuAny*>(h);
    auto go = [&](auto& E, auto tag) {
        typedef decltype(tag) R;
        const HostPlan& P = E.h.plan;
        R* o = (R*)out;
        for (size_t es = 0; es < size_t(P.n_tiles) * P.tile_e; ++es) {
            const uint32_t el = P.order[es];
            if (el == 0xFFFFFFFFu) continue;
            const Quad<R> q0 = E.h.rk0[es], q1 = E.h.rk1[es], q2 = E.h.rk2[es];
            R* r = o + 9 * size_t(el);
            r[0] = q0.a; r[1] = q0.b; r[2] = q0.c; r[3] = q0.d; r[4] = q1.a; r[5] = q1.b; r[6] = q1.c; r[7] = q1.d; r[8] = q2.a;
        }
    };
    if (e->real == 0) go(e->f, float()); else go(e->d, double());
}
}
