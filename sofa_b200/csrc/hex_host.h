// Host-side init of HexahedronFEMForceField<B200Vec3Types>: reinit() of the reference (material stiffness, rest rotation,
// rotated rest shape, 24x24 element stiffness by 2x2x2 Gauss quadrature), de-duplication of bit-identical K_e, tile plan.
#pragma once
#include <cstring>
#include <string>
#include <unordered_map>

#include "hex_kernels.cuh"
#include "plan.h"

namespace sb {

template <class R> struct HostHex {
    int method = 0;
    size_t n_nodes = 0, n_hexas = 0;
    HostPlan plan;
    std::vector<R> h_rot0, h_X0;          // element order: _initialrotations (9), _rotatedInitialElements (24)
    std::vector<uint32_t> h_kidx;         // element order
    std::vector<R> ktab;                  // unique K_e, 576 each
    std::vector<uint4> lnode, slot_a, slot_b;
    std::vector<Quad<R>> r0, r1, r2, x0;  // tile order (x0: 6 planes)
    std::vector<uint32_t> kidx, tile_kuniq;
    size_t smem_bytes = 0;
};

// computeElementStiffness, HexahedronFEMForceField.inl:306-536 (GENERIC_STIFFNESS_MATRIX, MAT_STIFFNESS_USE_W, DN_USE_J);
// the integrateStiffness() block (:457-502) is overwritten by `K = K1` (:504-506) and therefore omitted.
template <class R> static void hex_element_stiffness(R* K, R U, R V, R W, const V3<R>* nodes, double stiffnessFactor) {
    static const int coef[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};   // _coef, :66-89
    for (int i = 0; i < 576; ++i) K[i] = 0;
    M3<R> J, J_1;
    std::memset(&J, 0, sizeof(J)); std::memset(&J_1, 0, sizeof(J_1));
    R detJ = R(1.0);
    auto nrm = [](const V3<R>& v) { R n2 = v.x * v.x; n2 += v.y * v.y; n2 += v.z * v.z; return sqrt_r(n2); };
    auto comp = [](const V3<R>& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : v.z); };
    const V3<R> lx = nodes[1] - nodes[0], ly = nodes[3] - nodes[0], lz = nodes[4] - nodes[0];
    bool isParallel = false;
    if (nrm(nodes[3] + lx - nodes[2]) < nrm(lx) * 0.001 && nrm(nodes[0] + lz - nodes[4]) < nrm(lz) * 0.001 && nrm(nodes[1] + lz - nodes[5]) < nrm(lz) * 0.001 &&
        nrm(nodes[2] + lz - nodes[6]) < nrm(lz) * 0.001 && nrm(nodes[3] + lz - nodes[7]) < nrm(lz) * 0.001) {
        isParallel = true;
        for (int c = 0; c < 3; ++c) { J.m[c][0] = comp(lx, c) / 2; J.m[c][1] = comp(ly, c) / 2; J.m[c][2] = comp(lz, c) / 2; }
        detJ = det3(J);
        invert3(J_1, J);
    }
    const double inv_sqrt3 = 1.0 / std::sqrt(3.0);
    for (int gx1 = -1; gx1 <= 1; gx1 += 2) for (int gx2 = -1; gx2 <= 1; gx2 += 2) for (int gx3 = -1; gx3 <= 1; gx3 += 2) {
        const double x1 = gx1 * inv_sqrt3, x2 = gx2 * inv_sqrt3, x3 = gx3 * inv_sqrt3;
        if (!isParallel) {
            for (int c = 0; c < 3; ++c) {
                const R n0 = comp(nodes[0], c), n1 = comp(nodes[1], c), n2 = comp(nodes[2], c), n3 = comp(nodes[3], c), n4 = comp(nodes[4], c), n5 = comp(nodes[5], c), n6 = comp(nodes[6], c), n7 = comp(nodes[7], c);
                J.m[c][0] = (R)((n1 - n0) * (1 - x2) * (1 - x3) / 8 + (n2 - n3) * (1 + x2) * (1 - x3) / 8 + (n5 - n4) * (1 - x2) * (1 + x3) / 8 + (n6 - n7) * (1 + x2) * (1 + x3) / 8);
                J.m[c][1] = (R)((n3 - n0) * (1 - x1) * (1 - x3) / 8 + (n2 - n1) * (1 + x1) * (1 - x3) / 8 + (n7 - n4) * (1 - x1) * (1 + x3) / 8 + (n6 - n5) * (1 + x1) * (1 + x3) / 8);
                J.m[c][2] = (R)((n4 - n0) * (1 - x1) * (1 - x2) / 8 + (n5 - n1) * (1 + x1) * (1 - x2) / 8 + (n6 - n2) * (1 + x1) * (1 + x2) / 8 + (n7 - n3) * (1 - x1) * (1 + x2) / 8);
            }
            detJ = det3(J);
            invert3(J_1, J);
        }
        R qx[8], qy[8], qz[8];
        for (int i = 0; i < 8; ++i) {
            const R d1 = (R)((coef[i][0]) * (1 + coef[i][1] * x2) * (1 + coef[i][2] * x3) / 8.0);
            const R d2 = (R)((1 + coef[i][0] * x1) * (coef[i][1]) * (1 + coef[i][2] * x3) / 8.0);
            const R d3 = (R)((1 + coef[i][0] * x1) * (1 + coef[i][1] * x2) * (coef[i][2]) / 8.0);
            qx[i] = d1 * J_1.m[0][0] + d2 * J_1.m[1][0] + d3 * J_1.m[2][0];
            qy[i] = d1 * J_1.m[0][1] + d2 * J_1.m[1][1] + d3 * J_1.m[2][1];
            qz[i] = d1 * J_1.m[0][2] + d2 * J_1.m[1][2] + d3 * J_1.m[2][2];
        }
        for (int i = 0; i < 8; ++i) {
            R MB[6][3];
            MB[0][0] = U * qx[i]; MB[0][1] = V * qy[i]; MB[0][2] = V * qz[i];
            MB[1][0] = V * qx[i]; MB[1][1] = U * qy[i]; MB[1][2] = V * qz[i];
            MB[2][0] = V * qx[i]; MB[2][1] = V * qy[i]; MB[2][2] = U * qz[i];
            MB[3][0] = W * qy[i]; MB[3][1] = W * qx[i]; MB[3][2] = (R)0;
            MB[4][0] = (R)0;      MB[4][1] = W * qz[i]; MB[4][2] = W * qy[i];
            MB[5][0] = W * qz[i]; MB[5][1] = (R)0;      MB[5][2] = W * qx[i];
            for (int j = i; j < 8; ++j) {
                R k[3][3];
                k[0][0] = qx[j] * MB[0][0] + qy[j] * MB[3][0] + qz[j] * MB[5][0];
                k[0][1] = qx[j] * MB[0][1] + qy[j] * MB[3][1];
                k[0][2] = qx[j] * MB[0][2] + qz[j] * MB[5][2];
                k[1][0] = qy[j] * MB[1][0] + qx[j] * MB[3][0];
                k[1][1] = qy[j] * MB[1][1] + qx[j] * MB[3][1] + qz[j] * MB[4][1];
                k[1][2] = qy[j] * MB[1][2] + qz[j] * MB[4][2];
                k[2][0] = qz[j] * MB[2][0] + qx[j] * MB[5][0];
                k[2][1] = qz[j] * MB[2][1] + qy[j] * MB[4][1];
                k[2][2] = qz[j] * MB[2][2] + qy[j] * MB[4][2] + qx[j] * MB[5][2];
                for (int m = 0; m < 3; ++m) for (int l = 0; l < 3; ++l) { k[l][m] *= detJ; }
                for (int m = 0; m < 3; ++m) for (int l = 0; l < 3; ++l) K[(i * 3 + m) * 24 + (j * 3 + l)] += k[l][m];
            }
        }
    }
    for (int i = 0; i < 24; ++i) for (int j = i + 1; j < 24; ++j) K[j * 24 + i] = K[i * 24 + j];
    const R sf = (R)stiffnessFactor;
    for (int i = 0; i < 576; ++i) K[i] *= sf;
}

// reinit(): HexahedronFEMForceField.inl:125-180 with computeMaterialStiffness :683-709, initLarge :788-814, initPolar :886-916, initSmall :722-738
template <class R> static std::string hex_host_build(HostHex<R>& ff, size_t n_nodes, const R* x0, size_t n_hexas, const uint32_t* hexas,
                                                     const sofab200_hexfem_desc* desc, int chunk, int sm_count = 148) {
    ff.method = desc->method; ff.n_nodes = n_nodes; ff.n_hexas = n_hexas;
    for (size_t i = 0; i < 8 * n_hexas; ++i) if (hexas[i] >= n_nodes) return "hexahedron refers to a node index out of range";
    std::vector<R> young(desc->n_young), poisson(desc->n_poisson);
    for (size_t i = 0; i < young.size(); ++i) young[i] = R(desc->young[i]);
    for (size_t i = 0; i < poisson.size(); ++i) poisson[i] = R(desc->poisson[i]);
    ff.h_rot0.assign(9 * n_hexas, 0); ff.h_X0.assign(24 * n_hexas, 0); ff.h_kidx.assign(n_hexas, 0); ff.ktab.clear();
    std::unordered_map<std::string, uint32_t> uniq;
    std::vector<R> K(576);
    for (size_t i = 0; i < n_hexas; ++i) {
        const R nu = poisson.size() > i ? poisson[i] : poisson[0];
        const R E = young.size() > i ? young[i] : young[0];
        R U = 1, V = nu / (1 - nu), W = (1 - 2 * nu) / (2 * (1 - nu));
        const R s = (E * (1 - nu)) / ((1 + nu) * (1 - 2 * nu));
        U *= s; V *= s; W *= s;
        V3<R> nodes[8], rn[8];
        for (int w = 0; w < 8; ++w) { const size_t n = hexas[8 * i + w]; nodes[w] = mk3<R>(x0[3 * n], x0[3 * n + 1], x0[3 * n + 2]); }
        M3<R> rot;
        if (ff.method == SOFAB200_HEX_SMALL) { std::memset(&rot, 0, sizeof(rot)); rot.m[0][0] = rot.m[1][1] = rot.m[2][2] = 1; }
        else hex_rotation(rot, nodes, ff.method == SOFAB200_HEX_POLAR);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) ff.h_rot0[9 * i + 3 * r + c] = rot.m[r][c];
        for (int w = 0; w < 8; ++w) { rn[w] = mul(rot, nodes[w]); ff.h_X0[24 * i + 3 * w] = rn[w].x; ff.h_X0[24 * i + 3 * w + 1] = rn[w].y; ff.h_X0[24 * i + 3 * w + 2] = rn[w].z; }
        hex_element_stiffness(K.data(), U, V, W, rn, 1.0);
        const std::string key(reinterpret_cast<const char*>(K.data()), 576 * sizeof(R));
        auto it = uniq.find(key);
        if (it == uniq.end()) { it = uniq.emplace(key, uint32_t(ff.ktab.size() / 576)).first; ff.ktab.insert(ff.ktab.end(), K.begin(), K.end()); }
        ff.h_kidx[i] = it->second;
    }
    std::vector<double> pos(3 * n_nodes);
    for (size_t i = 0; i < 3 * n_nodes; ++i) pos[i] = double(x0[i]);
    const bool fixed_tile = desc->tile_elems > 0;
    const int cap = sizeof(R) == 4 ? 1024 : 512;
    int k_waves = std::max<int>(1, int((n_hexas + size_t(sm_count) * cap - 1) / (size_t(sm_count) * cap)));
    int tile_e = desc->tile_elems > 0 ? desc->tile_elems : int((n_hexas + size_t(sm_count) * k_waves - 1) / (size_t(sm_count) * k_waves));
    tile_e = std::max(32, (tile_e + 31) / 32 * 32);
    for (;;) {
        const std::string err = build_plan(ff.plan, int(n_nodes), int(n_hexas), 8, hexas, pos.data(), tile_e, chunk, kStageFlag);
        ff.smem_bytes = hex_smem_bytes<R>(ff.plan.max_touched, ff.plan.max_slots);
        const bool too_big = ff.smem_bytes > 200 * 1024 || err.find("use a smaller tile") != std::string::npos;
        if (too_big && !fixed_tile && tile_e > 32) { ++k_waves; tile_e = std::max(32, (int((n_hexas + size_t(sm_count) * k_waves - 1) / (size_t(sm_count) * k_waves)) + 31) / 32 * 32); continue; }
        if (!err.empty()) return err;
        if (too_big) return "tile does not fit in shared memory; use a smaller tile_elems";
        break;
    }
    // planes in tile order
    const HostPlan& P = ff.plan;
    const size_t NS = size_t(P.n_tiles) * P.tile_e;
    const Quad<R> z{0, 0, 0, 0};
    ff.lnode.assign(NS, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
    ff.slot_a.assign(NS, make_uint4(0, 0, 0, 0)); ff.slot_b.assign(NS, make_uint4(0, 0, 0, 0));
    ff.r0.assign(NS, z); ff.r1.assign(NS, z); ff.r2.assign(NS, z); ff.x0.assign(6 * NS, z); ff.kidx.assign(NS, 0);
    ff.tile_kuniq.assign(size_t(P.n_tiles) * (kHexSmemMatrices + 1), 0);
    for (int t = 0; t < P.n_tiles; ++t) {
        std::vector<uint32_t> u;
        bool many = false;
        for (size_t es = size_t(t) * P.tile_e; es < size_t(t + 1) * P.tile_e; ++es) {
            const uint32_t e = P.order[es];
            if (e == 0xFFFFFFFFu) continue;
            const uint32_t ki = ff.h_kidx[e];
            if (std::find(u.begin(), u.end(), ki) == u.end()) { if (int(u.size()) == kHexSmemMatrices) { many = true; break; } u.push_back(ki); }
        }
        uint32_t* dst = &ff.tile_kuniq[size_t(t) * (kHexSmemMatrices + 1)];
        if (!many) { dst[0] = uint32_t(u.size()); for (size_t i = 0; i < u.size(); ++i) dst[1 + i] = u[i]; }
    }
    for (size_t es = 0; es < NS; ++es) {
        const uint32_t e = P.order[es];
        if (e == 0xFFFFFFFFu) continue;
        const uint16_t* l = &P.lnode[8 * es]; const uint32_t* s = &P.slot[8 * es];
        ff.lnode[es] = make_uint4(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16), l[4] | (uint32_t(l[5]) << 16), l[6] | (uint32_t(l[7]) << 16));
        ff.slot_a[es] = make_uint4(s[0], s[1], s[2], s[3]); ff.slot_b[es] = make_uint4(s[4], s[5], s[6], s[7]);
        const R* r = &ff.h_rot0[9 * size_t(e)];
        ff.r0[es] = Quad<R>{r[0], r[1], r[2], r[3]}; ff.r1[es] = Quad<R>{r[4], r[5], r[6], r[7]}; ff.r2[es] = Quad<R>{r[8], 0, 0, 0};
        const R* x = &ff.h_X0[24 * size_t(e)];
        for (int q = 0; q < 6; ++q) ff.x0[size_t(q) * NS + es] = Quad<R>{x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]};
        ff.kidx[es] = ff.h_kidx[e];
    }
    return "";
}

}  // namespace sb
