// Host-side init of TetrahedronFEMForceField<B200Vec3Types>: the reference's reinit() arithmetic, the tile /
// gather plan, and the element planes in tile order.  Pure host code (also used by tests/emu).
#pragma once
#include <cstring>
#include <limits>

#include "math3.cuh"
#include "plan.h"
#include "tet_kernels.cuh"

namespace sb {

template <class R> struct HostTet {
    int method = 1;
    size_t n_nodes = 0, n_tets = 0;
    HostPlan plan;
    // element-ordered values, as the reference class keeps them
    std::vector<R> h_K, h_J, h_X0, h_R0t, h_A0inv;
    std::vector<R> h_A0;   // TetrahedralCorotationalFEMForceField, polar: initialTransformation = the rest edge matrix (its getRotation multiplies by it)
    std::vector<R> h_shf, h_lambda, h_mu, h_rest;   // computeVonMisesStress != 0: elemShapeFun rows 1..3 (12 per element), elemLambda, elemMu, d_initialPoints
    // element planes in tile order
    std::vector<ushort4> lnode; std::vector<uint4> slot;
    std::vector<Quad<R>> rk0, rk1, rk2, j0, j1, j2, x0a, x0b, x0c, sv[5];
    size_t smem_bytes = 0;
};

template <class R> static void strain_displacement(R* j, const V3<R>& a, const V3<R>& b, const V3<R>& c, const V3<R>& d) { tet_strain_displacement(j, a, b, c, d); }   // tet_kernels.cuh

// invertMatrix, general case (Sofa/framework/Type/src/sofa/type/Mat.h:1103-1166): Gauss-Jordan with full pivoting, S = 4
template <class R> static bool invert4(R dest[4][4], const R from[4][4]) {
    const int S = 4;
    int r[S] = {0, 0, 0, 0}, c[S] = {0, 0, 0, 0}, row[S] = {0, 0, 0, 0}, col[S] = {0, 0, 0, 0};
    R m1[S][S], m2[S][S];
    for (int i = 0; i < S; ++i) for (int j = 0; j < S; ++j) { m1[i][j] = from[i][j]; m2[i][j] = i == j ? R(1) : R(0); dest[i][j] = R(0); }
    for (int k = 0; k < S; k++) {
        R pivot = 0;
        for (int i = 0; i < S; i++) {
            if (row[i]) continue;
            for (int j = 0; j < S; j++) {
                if (col[j]) continue;
                R t = m1[i][j]; if (t < 0) t = -t;
                if (t > pivot) { pivot = t; r[k] = i; c[k] = j; }
            }
        }
        if (std::abs(pivot) <= std::numeric_limits<R>::epsilon()) return false;
        row[r[k]] = col[c[k]] = 1;
        pivot = m1[r[k]][c[k]];
        for (int j = 0; j < S; ++j) m1[r[k]][j] /= pivot;
        m1[r[k]][c[k]] = 1;
        for (int j = 0; j < S; ++j) m2[r[k]][j] /= pivot;
        for (int i = 0; i < S; i++) {
            if (i != r[k]) {
                const R f = m1[i][c[k]];
                for (int j = 0; j < S; ++j) m1[i][j] -= m1[r[k]][j] * f;
                m1[i][c[k]] = 0;
                for (int j = 0; j < S; ++j) m2[i][j] -= m2[r[k]][j] * f;
            }
        }
    }
    for (int i = 0; i < S; i++) for (int j = 0; j < S; j++) if (c[j] == i) row[i] = r[j];
    for (int i = 0; i < S; i++) for (int j = 0; j < S; ++j) dest[i][j] = m2[row[i]][j];
    return true;
}

// reinit(): TetrahedronFEMForceField.inl:1390-1505 with computeMaterialStiffness :255-291,
// initSmall :526-532, initLarge :834-868, initPolar :992-1023, initSVD :1086-1117
template <class R> static int tet_init_elements(HostTet<R>& ff, const R* x0, const uint32_t* tets, const sofab200_tetfem_desc* desc) {
    const size_t T = ff.n_tets;
    std::vector<R> young(desc->n_young), poisson(desc->n_poisson), lsf(desc->n_local_stiffness);
    for (size_t i = 0; i < young.size(); ++i) young[i] = R(desc->young[i]);
    for (size_t i = 0; i < poisson.size(); ++i) poisson[i] = R(desc->poisson[i]);
    for (size_t i = 0; i < lsf.size(); ++i) lsf[i] = R(desc->local_stiffness[i]);
    ff.h_K.assign(3 * T, 0); ff.h_J.assign(12 * T, 0); ff.h_X0.assign(12 * T, 0); ff.h_R0t.assign(9 * T, 0);
    if (ff.method == SOFAB200_TET_SVD) ff.h_A0inv.assign(9 * T, 0);
    auto P = [&](uint32_t n) { return mk3<R>(x0[3 * size_t(n)], x0[3 * size_t(n) + 1], x0[3 * size_t(n) + 2]); };
    if (desc->compute_von_mises) {
        // elemShapeFun :1521-1541: inverse of the matrix whose rows are (1, x0, y0, z0) of the 4 corners; only rows 1..3 are ever read
        ff.h_lambda.assign(T, 0); ff.h_mu.assign(T, 0); ff.h_shf.assign(12 * T, 0); ff.h_rest.assign(x0, x0 + 3 * ff.n_nodes);
        for (size_t i = 0; i < T; ++i) {
            R mv[4][4], inv[4][4];
            for (int k = 0; k < 4; ++k) { const size_t ix = tets[4 * i + k]; mv[k][0] = R(1.0); for (int l = 1; l < 4; ++l) mv[k][l] = x0[3 * ix + l - 1]; }
            invert4(inv, mv);
            for (int l = 1; l < 4; ++l) for (int m = 0; m < 4; ++m) ff.h_shf[12 * i + 4 * (l - 1) + m] = inv[l][m];
        }
    }
    for (size_t i = 0; i < T; ++i) {
        const uint32_t ia = tets[4 * i], ib = tets[4 * i + 1], ic = tets[4 * i + 2], id = tets[4 * i + 3];
        const V3<R> a = P(ia), b = P(ib), c = P(ic), d = P(id);
        // material stiffness
        const R E_el = young.size() > i ? young[i] : young[0];
        const R E = (lsf.empty() ? 1.0f : lsf[i * lsf.size() / T]) * E_el;
        const R nu = poisson.size() > i ? poisson[i] : poisson[0];
        R k00 = 1, k01 = nu / (1 - nu), k33 = (1 - 2 * nu) / (2 * (1 - nu));
        const R s = (E * (1 - nu)) / ((1 + nu) * (1 - 2 * nu));
        k00 *= s; k01 *= s; k33 *= s;
        if (desc->compute_von_mises) { ff.h_lambda[i] = k01; ff.h_mu[i] = k33; }   // elemLambda / elemMu, :278-282 (before the division by 36 V)
        const R vol = std::abs(dot3(cross3(b - a, c - a), d - a) / R(6));  // geometry::Tetrahedron::volume
        const R div = vol * 36;
        ff.h_K[3 * i] = k00 / div; ff.h_K[3 * i + 1] = k01 / div; ff.h_K[3 * i + 2] = k33 / div;
        R* X0 = &ff.h_X0[12 * i];
        R* R0t = &ff.h_R0t[9 * i];
        if (ff.method == SOFAB200_TET_SMALL) {
            strain_displacement(&ff.h_J[12 * i], a, b, c, d);
            const V3<R> q[4] = {a, b, c, d};
            for (int n = 0; n < 4; ++n) { X0[3 * n] = q[n].x; X0[3 * n + 1] = q[n].y; X0[3 * n + 2] = q[n].z; }
            R0t[0] = R0t[4] = R0t[8] = 1;
            continue;
        }
        M3<R> R01;
        if (ff.method == SOFAB200_TET_LARGE) {
            V3<R> ex = b - a; normalize3(ex);
            V3<R> ey = c - a;
            V3<R> ez = cross3(ex, ey); normalize3(ez);
            ey = cross3(ez, ex);
            set_row(R01, 0, ex); set_row(R01, 1, ey); set_row(R01, 2, ez);
        } else {
            M3<R> A;
            set_row(A, 0, b - a); set_row(A, 1, c - a); set_row(A, 2, d - a);
            if (desc->tetrahedral_corotational) { if (ff.h_A0.empty()) ff.h_A0.assign(9 * T, 0); for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) ff.h_A0[9 * i + 3 * r + cc] = A.m[r][cc]; }
            if (ff.method == SOFAB200_TET_SVD) {
                M3<R> Ai;
                std::memset(&Ai, 0, sizeof(Ai));
                invert3(Ai, A);
                for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) ff.h_A0inv[9 * i + 3 * r + cc] = Ai.m[r][cc];
            }
            polar_decomposition(A, R01);
        }
        for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) R0t[3 * r + cc] = R01.m[cc][r];
        V3<R> q[4] = {mul(R01, a), mul(R01, b), mul(R01, c), mul(R01, d)};
        if (ff.method == SOFAB200_TET_LARGE) {
            q[1] = q[1] - q[0]; q[2] = q[2] - q[0]; q[3] = q[3] - q[0];
            q[0] = mk3<R>(0, 0, 0);
        }
        for (int n = 0; n < 4; ++n) { X0[3 * n] = q[n].x; X0[3 * n + 1] = q[n].y; X0[3 * n + 2] = q[n].z; }
        strain_displacement(&ff.h_J[12 * i], q[0], q[1], q[2], q[3]);
    }
    return SOFAB200_OK;
}


// element planes in tile order
template <class R> static void tet_fill_planes(HostTet<R>& ff) {
    const HostPlan& P = ff.plan;
    const size_t NS = size_t(P.n_tiles) * P.tile_e;
    const Quad<R> z{0, 0, 0, 0};
    ff.lnode.resize(NS); ff.slot.resize(NS);
    for (auto* v : {&ff.rk0, &ff.rk1, &ff.rk2, &ff.j0, &ff.j1, &ff.j2, &ff.x0a, &ff.x0b, &ff.x0c}) v->assign(NS, z);
    const bool svd = ff.method == SOFAB200_TET_SVD;
    if (svd) for (auto& v : ff.sv) v.assign(NS, z);
    for (size_t es = 0; es < NS; ++es) {
        const uint32_t e = P.order[es];
        ff.lnode[es] = make_ushort4(P.lnode[4 * es], P.lnode[4 * es + 1], P.lnode[4 * es + 2], P.lnode[4 * es + 3]);
        ff.slot[es] = make_uint4(P.slot[4 * es], P.slot[4 * es + 1], P.slot[4 * es + 2], P.slot[4 * es + 3]);
        if (e == 0xFFFFFFFFu) continue;
        const R* r = &ff.h_R0t[9 * size_t(e)]; const R* k = &ff.h_K[3 * size_t(e)];
        const R* j = &ff.h_J[12 * size_t(e)]; const R* x = &ff.h_X0[12 * size_t(e)];
        ff.rk0[es] = Quad<R>{r[0], r[1], r[2], r[3]}; ff.rk1[es] = Quad<R>{r[4], r[5], r[6], r[7]}; ff.rk2[es] = Quad<R>{r[8], k[0], k[1], k[2]};
        ff.j0[es] = Quad<R>{j[0], j[1], j[2], j[3]}; ff.j1[es] = Quad<R>{j[4], j[5], j[6], j[7]}; ff.j2[es] = Quad<R>{j[8], j[9], j[10], j[11]};
        ff.x0a[es] = Quad<R>{x[0], x[1], x[2], x[3]}; ff.x0b[es] = Quad<R>{x[4], x[5], x[6], x[7]}; ff.x0c[es] = Quad<R>{x[8], x[9], x[10], x[11]};
        if (svd) {
            const R* a = &ff.h_A0inv[9 * size_t(e)];
            ff.sv[0][es] = Quad<R>{a[0], a[1], a[2], a[3]}; ff.sv[1][es] = Quad<R>{a[4], a[5], a[6], a[7]};
            ff.sv[2][es] = Quad<R>{a[8], r[0], r[1], r[2]}; ff.sv[3][es] = Quad<R>{r[3], r[4], r[5], r[6]}; ff.sv[4][es] = Quad<R>{r[7], r[8], 0, 0};
        }
    }
}

// init()+reinit()+layout.  Returns "" or an error text.
template <class R> static std::string tet_host_build(HostTet<R>& ff, size_t n_nodes, const R* x0, size_t n_tets, const uint32_t* tets,
                                                     const sofab200_tetfem_desc* desc, int chunk, int sm_count = 148) {
    ff.method = desc->method; ff.n_nodes = n_nodes; ff.n_tets = n_tets;
    for (size_t i = 0; i < 4 * n_tets; ++i) if (tets[i] >= n_nodes) return "tetrahedron refers to a node index out of range";
    tet_init_elements(ff, x0, tets, desc);
    std::vector<double> pos(3 * n_nodes);
    for (size_t i = 0; i < 3 * n_nodes; ++i) pos[i] = double(x0[i]);
    // Tile size.  One CTA streams one tile, so the number of tiles should be a whole multiple of the SM count (no partial
    // last wave), and the fewer waves the fewer times the per-tile phases (nodal staging, ordered sums) are paid: default is
    // the largest tile that cuts the mesh into k * sm_count equal parts and still keeps at least half of the corner
    // contributions in shared memory (build_plan demotes interior nodes beyond the budget to the L2 staging path).
    // desc->tile_elems or SOFAB200_TILE_ELEMS override; SOFAB200_SMEM_KB sets the budget.
    const bool fixed_tile = desc->tile_elems > 0 || getenv("SOFAB200_TILE_ELEMS");
    int tile_e = desc->tile_elems;
    if (const char* env = getenv("SOFAB200_TILE_ELEMS")) { const int v = atoi(env); if (v > 0) tile_e = v; }
    size_t smem_limit = 150 * 1024;   // what is left of the 256 KB L1/shared array holds the in-flight element records: a larger carve-out throttles the stream
    if (const char* env = getenv("SOFAB200_SMEM_KB")) { const int v = atoi(env); if (v >= 16 && v <= 224) smem_limit = size_t(v) * 1024; }
    typedef typename SVec<R>::T SV;
    // first guess: the tile whose corners would fit if 45 % of them were interior
    int k_waves = std::max<int>(1, int((double(n_tets) * 4 * 0.45 * 3 * sizeof(R)) / (double(sm_count) * smem_limit) + 0.999));
    if (const char* env = getenv("SOFAB200_TILE_WAVES")) { const int v = atoi(env); if (v > 0) k_waves = v; }
    auto tile_for = [&](int k) { return std::max(32, (int((n_tets + size_t(sm_count) * k - 1) / (size_t(sm_count) * k)) + 31) / 32 * 32); };
    if (tile_e <= 0) tile_e = tile_for(k_waves);
    tile_e = std::max(32, (tile_e + 31) / 32 * 32);
    int persist_tries = 0;
    for (;;) {
        // the persistent CG kernel gives CTA b the tiles b and b + ceil(n_tiles / 2) (or just b when there are at most sm_count):
        // the tile it processes last lists its elements that feed shared nodes first (see build_plan)
        const int n_tiles_guess = std::max(1, int((n_tets + size_t(tile_e) - 1) / size_t(tile_e)));
        // (tile b + c * grid goes to CTA b: the last tile of every CTA is among the final `sm_count` tiles)
        const int reorder_from = n_tiles_guess <= sm_count ? 0 : (n_tiles_guess <= 2 * sm_count ? (n_tiles_guess + 1) / 2 : n_tiles_guess - sm_count);
        const std::string err = build_plan(ff.plan, int(n_nodes), int(n_tets), 4, tets, pos.data(), tile_e, chunk, kStageFlag, smem_limit, sizeof(SV), 3 * sizeof(R), desc->shared_nodes,
                                           reorder_from);
        ff.smem_bytes = tile_smem_bytes<R>(ff.plan.max_touched, ff.plan.max_slots);
        const bool too_big = ff.smem_bytes > smem_limit || err.find("use a smaller tile") != std::string::npos;
        // demotion is meant to shave a tile that is a little too large, not to push a third of the mesh through the staging path:
        // when more than 2 % of the nodes had to be demoted, smaller tiles (one more wave) are the better layout
        const bool too_staged = err.empty() && ff.plan.n_demoted * 50 > n_nodes;
        if ((too_big || too_staged) && !fixed_tile && tile_e > 32) { tile_e = tile_for(++k_waves); continue; }
        if (!err.empty()) return err;
        if (too_big) return "tile does not fit in shared memory; use a smaller tile_elems";
        // The persistent CG kernel keeps the node tables of two tiles next to the slots.  196 KB is the largest shared-memory
        // carve-out that still leaves the L1 enough room for the in-flight element records (measured: the next step, 228 KB,
        // slows the stream by 15 %), so if its layout is a little over, give the slots a smaller budget and plan again.
        if (!fixed_tile && ff.plan.n_tiles <= 2 * sm_count && persist_tries < 3) {
            const int tpc = ff.plan.n_tiles > sm_count ? 2 : 1;
            const PersistLayout L = persist_layout<R>(tpc, ff.plan.max_touched, ff.plan.max_slots, ff.plan.max_int, ff.plan.max_shtouch, ff.plan.maxval);
            const size_t need = L.total + persist_static_smem<R>(512) + 1024, target = 196 * 1024;
            if (need > target && need - target < 24 * 1024) {
                const size_t tile_bytes = tile_smem_bytes<R>(ff.plan.max_touched, ff.plan.max_slots);
                smem_limit = std::min(smem_limit, tile_bytes) - (need - target) - 256;
                ++persist_tries;
                continue;
            }
        }
        break;
    }
    if (getenv("SOFAB200_VERBOSE"))
        fprintf(stderr, "[sofa_b200] tet plan: %d tiles x %d elems, max_touched %d, max_slots %d, max_int %d, shared %d in %d chunks, staged %zu, demoted %zu, smem %zu B\n",
                ff.plan.n_tiles, ff.plan.tile_e, ff.plan.max_touched, ff.plan.max_slots, ff.plan.max_int, ff.plan.n_shared, ff.plan.n_chunks, ff.plan.n_staged_corners,
                ff.plan.n_demoted, ff.smem_bytes);
    tet_fill_planes(ff);
    return "";
}

}  // namespace sb
