// TetrahedronFEMForceField on the device: one CTA per tile of elements.
//   phase 1  stage the tile's nodal input vectors (x or dx) in shared memory
//   phase 2  one thread per element: rotation (addForce) or cached rotation (addDForce), F = J (K (J^T D)),
//            rotate back; scatter the 4 corner contributions to their slot (shared memory for interior
//            nodes, HBM staging for shared nodes)
//   phase 3  one thread per interior node: sequential sum of its slots in element order + fused epilogue
// Arithmetic follows the reference statement by statement (file:line cited per function); compiled with
// -fmad=false so no multiply-add is contracted.
#pragma once
#include "cg_fused.cuh"
#include "math3.cuh"

namespace sb {

enum TetMode { TM_DF_COROT = 0, TM_DF_SMALL = 1, TM_F_SMALL = 2, TM_F_LARGE = 3, TM_F_POLAR = 4, TM_F_SVD = 5,
               // TetrahedronFEMForceField, method large with updateStiffnessMatrix: addForce rewrites ONE of the three copies of nine cofactors
               // (the normal-strain columns J(.,0..2), TetrahedronFEMForceField.inl:908-922); the shear columns J(.,3..5) keep their initial values,
               // so the element carries a second set of 12 cofactors (planes js0..2) for them.
               TM_DF_COROT_JS = 6, TM_F_LARGE_JS = 7 };

// peudo_determinant_for_coef, TetrahedronFEMForceField.inl:204-208
template <class R> HD R tet_pdet(R m00, R m01, R m02, R m10, R m11, R m12) {
    return m01 * m12 - m11 * m02 - m00 * m12 + m10 * m02 + m00 * m11 - m10 * m01;
}
// computeStrainDisplacement, :134-202 -- the 12 distinct cofactors
template <class R> HD void tet_strain_displacement(R* j, const V3<R>& a, const V3<R>& b, const V3<R>& c, const V3<R>& d) {
    j[0] = -tet_pdet(b.y, c.y, d.y, b.z, c.z, d.z);
    j[1] = tet_pdet(b.x, c.x, d.x, b.z, c.z, d.z);
    j[2] = -tet_pdet(b.x, c.x, d.x, b.y, c.y, d.y);
    j[3] = tet_pdet(c.y, d.y, a.y, c.z, d.z, a.z);
    j[4] = -tet_pdet(c.x, d.x, a.x, c.z, d.z, a.z);
    j[5] = tet_pdet(c.x, d.x, a.x, c.y, d.y, a.y);
    j[6] = -tet_pdet(d.y, a.y, b.y, d.z, a.z, b.z);
    j[7] = tet_pdet(d.x, a.x, b.x, d.z, a.z, b.z);
    j[8] = -tet_pdet(d.x, a.x, b.x, d.y, a.y, b.y);
    j[9] = tet_pdet(a.y, b.y, c.y, a.z, b.z, c.z);
    j[10] = -tet_pdet(a.x, b.x, c.x, a.z, b.z, c.z);
    j[11] = tet_pdet(a.x, b.x, c.x, a.y, b.y, c.y);
}

template <class R> struct TetDev {
    TileDev<R> t;
    const ushort4* lnode;      // [n_tiles*tile_e] local node index of the 4 corners (0xFFFF = padding element); read as one 64-bit word
    const uint4* slot;         // destination slot of each corner's contribution
    Quad<R>* rk0; Quad<R>* rk1; Quad<R>* rk2;            // rotations[e] (9, row-major) + {K00, K01, K33}
    const Quad<R>* j0; const Quad<R>* j1; const Quad<R>* j2;     // 12 strain-displacement cofactors
    const Quad<R>* js0; const Quad<R>* js1; const Quad<R>* js2;   // TM_*_JS: cofactors of the shear columns of J (initial values)
    Quad<R>* j0w; Quad<R>* j1w; Quad<R>* j2w;                    // the same planes, writable: non-null when updateStiffnessMatrix is set (polar / svd addForce rewrites them)
    const Quad<R>* x0a; const Quad<R>* x0b; const Quad<R>* x0c;  // _rotatedInitialElements (small: rest positions)
    const Quad<R>* sv0; const Quad<R>* sv1; const Quad<R>* sv2; const Quad<R>* sv3; const Quad<R>* sv4;  // svd: A0^-1 (9) + R0^T (9)
    R k_factor;                // addDForce: (Real)kFactorIncludingRayleighDamping
    Quad<R>* pl0; Quad<R>* pl1;  // _plasticStrains[e] (6 Voigt components: pl0 = 0..3, pl1.a/b = 4..5); null unless plasticMaxThreshold > 0
    R plastic_max, plastic_yield, plastic_creep;
};

// computeForce (TetrahedronFEMForceField.inl:293-415 without plasticity, :417-521 with `fact`).
// j[3n..3n+2] = (jx,jy,jz) of node n: J(3n,0)=J(3n+1,3)=J(3n+2,5)=jx, J(3n,3)=J(3n+1,1)=J(3n+2,4)=jy,
// J(3n,5)=J(3n+1,4)=J(3n+2,2)=jz; K has three distinct values k0=K(i,i) i<3, k1=K(i,j) i!=j<3, k2=K(i,i) i>=3.
// `ps` (addForce with plasticMaxThreshold > 0): the element's plastic strain, updated as :357-371 do; pp = {max, yield, creep}.
template <class R, bool USE_FACT, bool PLASTIC = false> HD void tet_compute_force(R F[12], const R D[12], const R j[12], const R js[12], R k0, R k1, R k2, R fact, R* ps = nullptr, const R* pp = nullptr) {
    // j: the copies of the cofactors in the normal-strain columns J(.,0..2); js: those in the shear columns J(.,3..5) (the same array unless TM_*_JS)
    R s0 = j[0] * D[0] + j[3] * D[3] + j[6] * D[6] + j[9] * D[9];
    R s1 = j[1] * D[1] + j[4] * D[4] + j[7] * D[7] + j[10] * D[10];
    R s2 = j[2] * D[2] + j[5] * D[5] + j[8] * D[8] + j[11] * D[11];
    R s3 = js[1] * D[0] + js[0] * D[1] + js[4] * D[3] + js[3] * D[4] + js[7] * D[6] + js[6] * D[7] + js[10] * D[9] + js[9] * D[10];
    R s4 = js[2] * D[1] + js[1] * D[2] + js[5] * D[4] + js[4] * D[5] + js[8] * D[7] + js[7] * D[8] + js[11] * D[10] + js[10] * D[11];
    R s5 = js[2] * D[0] + js[0] * D[2] + js[5] * D[3] + js[3] * D[5] + js[8] * D[6] + js[6] * D[8] + js[11] * D[9] + js[9] * D[11];
    if (PLASTIC) {
        // elasticStrain = JtD - plasticStrain; creep when |elastic|^2 > yield^2; clamp |plastic| to max; JtD -= plasticStrain
        const R el[6] = {s0 - ps[0], s1 - ps[1], s2 - ps[2], s3 - ps[3], s4 - ps[4], s5 - ps[5]};
        R n2 = el[0] * el[0];
#pragma unroll
        for (int i = 1; i < 6; ++i) n2 += el[i] * el[i];
        if (n2 > pp[1] * pp[1]) {
#pragma unroll
            for (int i = 0; i < 6; ++i) ps[i] += pp[2] * el[i];
        }
        R pn2 = ps[0] * ps[0];
#pragma unroll
        for (int i = 1; i < 6; ++i) pn2 += ps[i] * ps[i];
        if (pn2 > pp[0] * pp[0]) {
            const R sc = pp[0] / R(sqrt(pn2));   // helper::rsqrt is the square root (rmath.h:125-138)
#pragma unroll
            for (int i = 0; i < 6; ++i) ps[i] *= sc;
        }
        s0 -= ps[0]; s1 -= ps[1]; s2 -= ps[2]; s3 -= ps[3]; s4 -= ps[4]; s5 -= ps[5];
    }
    R t0 = k0 * s0 + k1 * s1 + k1 * s2;
    R t1 = k1 * s0 + k0 * s1 + k1 * s2;
    R t2 = k1 * s0 + k1 * s1 + k0 * s2;
    R t3 = k2 * s3, t4 = k2 * s4, t5 = k2 * s5;
    if (USE_FACT) { t0 *= fact; t1 *= fact; t2 *= fact; t3 *= fact; t4 *= fact; t5 *= fact; }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const R jx = j[3 * n], jy = j[3 * n + 1], jz = j[3 * n + 2], sx = js[3 * n], sy = js[3 * n + 1], sz = js[3 * n + 2];
        F[3 * n + 0] = jx * t0 + sy * t3 + sz * t5;
        F[3 * n + 1] = jy * t1 + sx * t3 + sz * t4;
        F[3 * n + 2] = jz * t2 + sy * t4 + sx * t5;
    }
}

// One element: inputs are the 4 nodal vectors P (positions for addForce, dx for addDForce); outputs the 4 corner
// contributions C in the form the reference adds (addForce) or subtracts (addDForce) them.  addForce modes also
// store rotations[e].  Host-callable so that tests/emu can execute the very same statements on the CPU.
template <class R> struct TetRec { Quad<R> q0, q1, q2, ja, jb, jc; };   // rotations[e]+K and the 12 cofactors of one element
template <class R> HD TetRec<R> tet_load_rec(const TetDev<R>& d, size_t es, uint64_t pol = 0) {
    TetRec<R> r;
    r.q0 = rec_load(d.rk0 + es, pol); r.q1 = rec_load(d.rk1 + es, pol); r.q2 = rec_load(d.rk2 + es, pol);
    r.ja = rec_load(d.j0 + es, pol); r.jb = rec_load(d.j1 + es, pol); r.jc = rec_load(d.j2 + es, pol);
    return r;
}
// computeForce(F, D, _plasticStrains[e], K, J) as addForce calls it (:560,927,1071,1180)
template <class R> HD void tet_compute_force_addforce(const TetDev<R>& d, size_t es, R F[12], const R D[12], const R j[12], const R js[12], R k0, R k1, R k2) {
    if (d.pl0) {
        const Quad<R> a = d.pl0[es], b = d.pl1[es];
        R ps[6] = {a.a, a.b, a.c, a.d, b.a, b.b};
        const R pp[3] = {d.plastic_max, d.plastic_yield, d.plastic_creep};
        tet_compute_force<R, false, true>(F, D, j, js, k0, k1, k2, R(0), ps, pp);
        d.pl0[es] = Quad<R>{ps[0], ps[1], ps[2], ps[3]}; d.pl1[es] = Quad<R>{ps[4], ps[5], R(0), R(0)};
    } else tet_compute_force<R, false>(F, D, j, js, k0, k1, k2, R(0));
}
template <class R, int MODE> HD void tet_element(const TetDev<R>& d, size_t es, const TetRec<R>& rec, const V3<R> P[4], V3<R> C[4]) {
    const Quad<R> q0 = rec.q0, q1 = rec.q1, q2 = rec.q2;
    const Quad<R> ja = rec.ja, jb = rec.jb, jc = rec.jc;
    R j[12] = {ja.a, ja.b, ja.c, ja.d, jb.a, jb.b, jb.c, jb.d, jc.a, jc.b, jc.c, jc.d};
    const R k0 = q2.b, k1 = q2.c, k2 = q2.d;
    R js_own[12];
    const R* js = j;
    if (MODE == TM_DF_COROT_JS || MODE == TM_F_LARGE_JS) {
        const Quad<R> sa = d.js0[es], sb = d.js1[es], sc = d.js2[es];
        js_own[0] = sa.a; js_own[1] = sa.b; js_own[2] = sa.c; js_own[3] = sa.d; js_own[4] = sb.a; js_own[5] = sb.b; js_own[6] = sb.c; js_own[7] = sb.d;
        js_own[8] = sc.a; js_own[9] = sc.b; js_own[10] = sc.c; js_own[11] = sc.d;
        js = js_own;
    }
    R F[12];
    if (MODE == TM_DF_COROT || MODE == TM_DF_COROT_JS) {
        // applyStiffnessCorotational, TetrahedronFEMForceField.inl:1192-1237 (rot = rotations[e])
        const R r00 = q0.a, r01 = q0.b, r02 = q0.c, r10 = q0.d, r11 = q1.a, r12 = q1.b, r20 = q1.c, r21 = q1.d, r22 = q2.a;
        R X[12];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            X[3 * n + 0] = r00 * P[n].x + r10 * P[n].y + r20 * P[n].z;
            X[3 * n + 1] = r01 * P[n].x + r11 * P[n].y + r21 * P[n].z;
            X[3 * n + 2] = r02 * P[n].x + r12 * P[n].y + r22 * P[n].z;
        }
        tet_compute_force<R, true>(F, X, j, js, k0, k1, k2, d.k_factor);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            C[n].x = r00 * F[3 * n] + r01 * F[3 * n + 1] + r02 * F[3 * n + 2];
            C[n].y = r10 * F[3 * n] + r11 * F[3 * n + 1] + r12 * F[3 * n + 2];
            C[n].z = r20 * F[3 * n] + r21 * F[3 * n + 1] + r22 * F[3 * n + 2];
        }
    } else if (MODE == TM_DF_SMALL) {
        // applyStiffnessSmall, :724-748
        R X[12];
#pragma unroll
        for (int n = 0; n < 4; ++n) { X[3 * n] = P[n].x; X[3 * n + 1] = P[n].y; X[3 * n + 2] = P[n].z; }
        tet_compute_force<R, true>(F, X, j, js, k0, k1, k2, d.k_factor);
#pragma unroll
        for (int n = 0; n < 4; ++n) C[n] = mk3<R>(F[3 * n], F[3 * n + 1], F[3 * n + 2]);
    } else {
        const Quad<R> xa = d.x0a[es], xb = d.x0b[es], xc = d.x0c[es];
        const R X0[12] = {xa.a, xa.b, xa.c, xa.d, xb.a, xb.b, xb.c, xb.d, xc.a, xc.b, xc.c, xc.d};
        R D[12];
        if (MODE == TM_F_SMALL) {
            // accumulateForceSmall, :534-556 (X0 planes hold the rest positions of the 4 nodes)
            D[0] = 0; D[1] = 0; D[2] = 0;
#pragma unroll
            for (int n = 1; n < 4; ++n) {
                D[3 * n + 0] = X0[3 * n + 0] - X0[0] - P[n].x + P[0].x;
                D[3 * n + 1] = X0[3 * n + 1] - X0[1] - P[n].y + P[0].y;
                D[3 * n + 2] = X0[3 * n + 2] - X0[2] - P[n].z + P[0].z;
            }
            tet_compute_force_addforce<R>(d, es, F, D, j, js, k0, k1, k2);
#pragma unroll
            for (int n = 0; n < 4; ++n) C[n] = mk3<R>(F[3 * n], F[3 * n + 1], F[3 * n + 2]);
        } else {
            M3<R> R02;  // R_0_2: rows = element frame axes
            if (MODE == TM_F_LARGE || MODE == TM_F_LARGE_JS) {
                // computeRotationLarge, :754-778
                V3<R> ex = P[1] - P[0];
                normalize3(ex);
                V3<R> ey = P[2] - P[0];
                V3<R> ez = cross3(ex, ey);
                normalize3(ez);
                ey = cross3(ez, ex);
                set_row(R02, 0, ex); set_row(R02, 1, ey); set_row(R02, 2, ez);
            } else {
                M3<R> A;
                set_row(A, 0, P[1] - P[0]); set_row(A, 1, P[2] - P[0]); set_row(A, 2, P[3] - P[0]);
                if (MODE == TM_F_SVD) {
                    // accumulateForceSVD, :1122-1152
                    const Quad<R> v0 = d.sv0[es], v1 = d.sv1[es], v2 = d.sv2[es], v3 = d.sv3[es], v4 = d.sv4[es];
                    M3<R> A0i, R0t;
                    A0i.m[0][0] = v0.a; A0i.m[0][1] = v0.b; A0i.m[0][2] = v0.c; A0i.m[1][0] = v0.d;
                    A0i.m[1][1] = v1.a; A0i.m[1][2] = v1.b; A0i.m[2][0] = v1.c; A0i.m[2][1] = v1.d; A0i.m[2][2] = v2.a;
                    R0t.m[0][0] = v2.b; R0t.m[0][1] = v2.c; R0t.m[0][2] = v2.d; R0t.m[1][0] = v3.a;
                    R0t.m[1][1] = v3.b; R0t.m[1][2] = v3.c; R0t.m[2][0] = v3.d; R0t.m[2][1] = v4.a; R0t.m[2][2] = v4.b;
                    const M3<R> Fm = mul(A, A0i);
                    if (double(det3(Fm)) < 1e-6) {
                        polar_decomposition_stable(Fm, R02);
                        R02 = mul_abt(R02, R0t);  // R_0_2.multTransposed(_initialRotations[e])
                    } else polar_decomposition(A, R02);
                } else polar_decomposition(A, R02);  // accumulateForcePolar, :1025-1038
            }
            const M3<R> rot = transpose(R02);  // rotations[e] = R_0_2^T
            d.rk0[es] = Quad<R>{rot.m[0][0], rot.m[0][1], rot.m[0][2], rot.m[1][0]};
            d.rk1[es] = Quad<R>{rot.m[1][1], rot.m[1][2], rot.m[2][0], rot.m[2][1]};
            d.rk2[es] = Quad<R>{rot.m[2][2], k0, k1, k2};
            V3<R> def[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) def[n] = mul(R02, P[n]);
            if (MODE == TM_F_LARGE || MODE == TM_F_LARGE_JS) {
                // accumulateForceLarge, :870-905
                def[1].x -= def[0].x;
                def[2].x -= def[0].x;
                def[2].y -= def[0].y;
                def[3].x -= def[0].x; def[3].y -= def[0].y; def[3].z -= def[0].z;
                D[0] = 0; D[1] = 0; D[2] = 0;
                D[3] = X0[3] - def[1].x; D[4] = 0; D[5] = 0;
                D[6] = X0[6] - def[2].x; D[7] = X0[7] - def[2].y; D[8] = 0;
                D[9] = X0[9] - def[3].x; D[10] = X0[10] - def[3].y; D[11] = X0[11] - def[3].z;
                if (d.j0w) {   // d_updateStiffnessMatrix: TetrahedralCorotationalFEMForceField.inl:920-937 rewrites all three copies of the nine cofactors (TM_F_LARGE, one set);
                               // TetrahedronFEMForceField.inl:908-922 only the copy in the normal-strain column (TM_F_LARGE_JS: `j` here, the shear copies stay in js)
                    j[0] = -def[2].y * def[3].z;
                    j[1] = def[2].x * def[3].z - def[1].x * def[3].z;
                    j[2] = def[2].y * def[3].x - def[2].x * def[3].y + def[1].x * def[3].y - def[1].x * def[2].y;
                    j[3] = def[2].y * def[3].z;
                    j[4] = -def[2].x * def[3].z;
                    j[5] = -def[2].y * def[3].x + def[2].x * def[3].y;
                    j[7] = def[1].x * def[3].z;
                    j[8] = -def[1].x * def[3].y;
                    j[11] = def[1].x * def[2].y;
                    d.j0w[es] = Quad<R>{j[0], j[1], j[2], j[3]}; d.j1w[es] = Quad<R>{j[4], j[5], j[6], j[7]}; d.j2w[es] = Quad<R>{j[8], j[9], j[10], j[11]};
                }
            } else {
                // :1047-1058 / :1160-1171
#pragma unroll
                for (int n = 0; n < 4; ++n) { D[3 * n] = X0[3 * n] - def[n].x; D[3 * n + 1] = X0[3 * n + 1] - def[n].y; D[3 * n + 2] = X0[3 * n + 2] - def[n].z; }
                if (d.j0w) {   // d_updateStiffnessMatrix: shape functions from the deformed element, :1063-1067 / :1174-1177
                    tet_strain_displacement(j, def[0], def[1], def[2], def[3]);
                    d.j0w[es] = Quad<R>{j[0], j[1], j[2], j[3]}; d.j1w[es] = Quad<R>{j[4], j[5], j[6], j[7]}; d.j2w[es] = Quad<R>{j[8], j[9], j[10], j[11]};
                }
            }
            tet_compute_force_addforce<R>(d, es, F, D, j, js, k0, k1, k2);
            // f[index[i/3]] += rotations[e] * Deriv(F[i],F[i+1],F[i+2]), :928-929
#pragma unroll
            for (int n = 0; n < 4; ++n) C[n] = mul(rot, mk3<R>(F[3 * n], F[3 * n + 1], F[3 * n + 2]));
        }
    }
}

// phase 2 of a tile: one thread per element, 4 corner contributions scattered to their slots.
// PF: software-prefetch the next element record (the record of the thread's NEXT element is requested before the current
// one is processed, so one element's worth of HBM latency is always overlapped with ~500 instructions of arithmetic).
// arrive_slots (persistent CG kernel, the CTA's last tile): once the elements [0, arrive_at) are done -- the plan puts the elements
// that feed shared nodes first -- the CTA arrives at the "staged contributions complete" grid barrier and goes on with the
// rest of the tile; arrive_at is a multiple of blockDim.x with arrive_at + blockDim.x <= tile_e (every thread passes there).
// NTHR: number of threads that share the tile (0: the whole CTA); on_boundary: called by every one of them, at the same trip count, once the
// elements [0, arrive_at) are done (arrive_at < 0: never).
struct NoBoundaryCall { __device__ __forceinline__ void operator()() const {} };
// The records of the FIRST round of a tile (one element per thread): the persistent CG kernel asks the L2 for them before the phase that precedes
// the tile (interior sums of the previous tile, x/r/p update of the previous iteration), so that the stream does not start from an idle memory
// system.  L2 prefetches only (no registers held across the phase; holding the record itself made ptxas spill it, which waits for the data).
struct TetFirst {};
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
template <class R> __device__ __forceinline__ void tet_prefetch_first(const TetDev<R>& d, int tile) {
    const TileDev<R>& t = d.t;
    if (int(threadIdx.x) < t.tile_e && (threadIdx.x & 7) == 0) {      // one request per 128-byte line of the 16-byte planes
        const size_t es = size_t(tile) * t.tile_e + threadIdx.x;
        prefetch_l2(d.slot + es); prefetch_l2(d.rk0 + es); prefetch_l2(d.rk1 + es); prefetch_l2(d.rk2 + es);
        prefetch_l2(d.j0 + es); prefetch_l2(d.j1 + es); prefetch_l2(d.j2 + es);
        if ((threadIdx.x & 15) == 0) prefetch_l2(d.lnode + es);
    }
}
template <class R, int MODE, bool PF, int NTHR = 0, class OnBoundary = NoBoundaryCall>
__device__ __forceinline__ void tet_tile_elements(const TetDev<R>& d, int tile, const typename SVec<R>::T* s_in, R* s_slot, int max_slots,
                                                  int arrive_at = -1, OnBoundary on_boundary = OnBoundary()) {
    typedef typename SVec<R>::T SV;
    const TileDev<R>& t = d.t;
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const int nthr = NTHR > 0 ? NTHR : int(blockDim.x);
    int le = threadIdx.x;
    uint2 lnw = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); uint4 sl = make_uint4(0, 0, 0, 0);
    TetRec<R> rec;
    if (le < t.tile_e) {
        const size_t es = size_t(tile) * t.tile_e + le;
        lnw = idx_load(reinterpret_cast<const uint2*>(d.lnode + es), pol_stream); sl = idx_load(d.slot + es, pol_stream); rec = tet_load_rec(d, es, pol_stream);
    }
    while (le < t.tile_e) {
        if (le - int(threadIdx.x) == arrive_at) on_boundary();
        const size_t es = size_t(tile) * t.tile_e + le;
        const int nle = le + nthr;
        uint2 n_lnw = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); uint4 n_sl = make_uint4(0, 0, 0, 0);
        TetRec<R> n_rec;
        if (PF && nle < t.tile_e) {
            const size_t nes = size_t(tile) * t.tile_e + nle;
            n_lnw = idx_load(reinterpret_cast<const uint2*>(d.lnode + nes), pol_stream); n_sl = idx_load(d.slot + nes, pol_stream); n_rec = tet_load_rec(d, nes, pol_stream);
        }
        if ((lnw.x & 0xFFFFu) != 0xFFFFu) {
            const SV pa = s_in[lnw.x & 0xFFFFu], pb = s_in[lnw.x >> 16], pc = s_in[lnw.y & 0xFFFFu], pd = s_in[lnw.y >> 16];
            const V3<R> P[4] = {mk3<R>(pa.x, pa.y, pa.z), mk3<R>(pb.x, pb.y, pb.z), mk3<R>(pc.x, pc.y, pc.z), mk3<R>(pd.x, pd.y, pd.z)};
            V3<R> C[4];
            tet_element<R, MODE>(d, es, rec, P, C);
            const unsigned s4[4] = {sl.x, sl.y, sl.z, sl.w};
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                tile_scatter<R>(t, s4[n], C[n].x, C[n].y, C[n].z, s_slot, max_slots, pol_keep);
            }
        }
        le = nle;
        if (PF) { lnw = n_lnw; sl = n_sl; rec = n_rec; }
        else if (le < t.tile_e) {
            const size_t nes = size_t(tile) * t.tile_e + le;
            lnw = idx_load(reinterpret_cast<const uint2*>(d.lnode + nes), pol_stream); sl = idx_load(d.slot + nes, pol_stream); rec = tet_load_rec(d, nes, pol_stream);
        }
    }
}

// MAXT: CTA size the kernel is compiled for (register budget 65536/MAXT)
template <class R, int MODE, int MAXT, bool PF>
__global__ void __launch_bounds__(MAXT) tet_tile_kernel(TetDev<R> d, const R* __restrict__ in, NodeEpilogue<R> ep, int max_touched, int max_slots) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ uint16_t s_jds[1024];
    typedef typename SVec<R>::T SV;
    if (ep.cg && ep.cg->done) return;
    SV* s_in = reinterpret_cast<SV*>(smem_raw);
    size_t off = (sizeof(SV) * size_t(max_touched) + 15) & ~size_t(15);
    R* s_slot = reinterpret_cast<R*>(smem_raw + off);  // 3 planes of max_slots

    const TileDev<R>& t = d.t;
    const int tile = blockIdx.x;
    trace_mark(ep.trace, 0, 0);
    // ---- phase 1: stage input vectors of the touched nodes
    tile_phase1<R>(t, tile, in, s_in, s_jds);
    trace_mark(ep.trace, 0, 1);
    // ---- phase 2: elements
    tet_tile_elements<R, MODE, PF>(d, tile, s_in, s_slot, max_slots);
    trace_mark(ep.trace, 0, 2);
    __syncthreads();
    trace_mark(ep.trace, 0, 3);

    // ---- phase 3: interior nodes, sequential sum in element order + fused epilogue
    const double part = tile_phase3<R>(t, tile, ep, s_in, s_slot, max_slots, s_jds);
    if (ep.dot_kind != DOT_NONE) {
        const double tot = block_sum(part, red);
        finish_dot(ep, tot, red, false);
    }
    trace_mark(ep.trace, 0, 4);
}

// ---- the whole CG loop of CGLinearSolver::solve in ONE persistent cooperative kernel ----------------------------------
// One CTA per SM keeps its (one or two) tiles for the whole solve; the three global dependencies of an iteration (staged
// contributions of the shared nodes, den = p.q, rho = r.r) are crossed with a grid-wide sync-and-sum instead of a kernel
// boundary, and everything static (node ids, valences, masses, fixed flags) is read from HBM once per solve:
//   [A] p = p*beta + r for the nodes of the CTA's tiles; per tile: element pass, interior nodes -> q, partial p.q
//   [B] shared nodes (ordered sum of the staged contributions) -> q, p                        -> den, alpha
//   [C] x += alpha p ; r -= alpha q                                                            -> rho, beta
// Every CTA sums the same partials in the same order, so all CTAs take the same branches.
template <class R, int MODE, int MAXT, bool PF>
__global__ void __launch_bounds__(MAXT) tet_cg_persistent_kernel(TetDev<R> d, PersistCG<R> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ double bcast;
    __shared__ GRec<R> s_grec[MAXT];
    typedef typename SVec<R>::T SV;
    CGDev* cg = a.cg;
    if (cg->done) return;
    const TileDev<R>& t = d.t;
    const PersistLayout& L = a.lay;
    SV* s_in = reinterpret_cast<SV*>(smem_raw);
    R* s_slot = reinterpret_cast<R*>(smem_raw + L.off_slot);
    PersistState<R> st(a);
    trace_mark(a.ep.trace, kTraceTail, 8);
    persist_load_tables<R>(t, a, smem_raw, s_grec);
    trace_mark(a.ep.trace, kTraceTail, 9);
    if (!persist_init<R>(a, st, red, &bcast)) { persist_finish<R>(t, a, st, smem_raw, s_grec); return; }
    trace_mark(a.ep.trace, kTraceTail, 10);
    for (;;) {
        trace_mark(a.ep.trace, kTraceTail, 0);
        // ---- [A]
        persist_phase1<R>(t, a, st, smem_raw);
        trace_mark(a.ep.trace, kTraceTail, 12);
        double part = 0.0;
        bool arrived = false;
        for (int c = 0; c < L.tiles_cached; ++c) {
            const int tile = blockIdx.x + c * gridDim.x;
            if (tile >= t.n_tiles) break;
            // the CTA's last tile: its elements that feed shared nodes come first; after them the staged contributions of this
            // CTA are complete, and the rest of the tile (plus its interior sums) runs while the other CTAs arrive
            const bool last = c + 1 == L.tiles_cached || tile + int(gridDim.x) >= t.n_tiles;
            int arrive_at = -1;
            if (last && t.tile_nb && t.tile_nb[tile] != 0xFFFFFFFFu) {
                const int nb_up = (int(t.tile_nb[tile]) + int(blockDim.x) - 1) / int(blockDim.x) * int(blockDim.x);
                if (nb_up + int(blockDim.x) <= t.tile_e) arrive_at = nb_up;
            }
            tet_tile_elements<R, MODE, PF>(d, tile, s_in + c * L.max_touched, s_slot, L.max_slots, arrive_at,
                                           [&]() { grid_arrive(a.sync, 0u, blockDim.x - 32u); });   // the last warp has the shortest tail of the tile
            if (arrive_at >= 0) arrived = true;
            __syncthreads();
            if (c == 0) trace_mark(a.ep.trace, kTraceTail, 13);
            part += persist_phase3<R>(t, tile, c, a, smem_raw);
            __syncthreads();
            if (c == 0) trace_mark(a.ep.trace, kTraceTail, 14);
        }
        if (!persist_rest<R>(t, a, st, part, red, &bcast, smem_raw, s_grec, arrived)) break;
    }
    persist_finish<R>(t, a, st, smem_raw, s_grec);
    trace_mark(a.ep.trace, kTraceTail, 11);
}

// element policy of the fused CG kernel (cg_fused.cuh)
template <class R, int MODE, bool PF> struct TetPass {
    typedef TetDev<R> Dev;
    static __device__ __forceinline__ const TileDev<R>& tiles(const Dev& d) { return d.t; }
    typedef TetFirst First;
    static __device__ __forceinline__ void prefetch(const Dev& d, int tile, First&) { tet_prefetch_first<R>(d, tile); }
    template <int ET, class OnBoundary>
    static __device__ __forceinline__ void elements(const Dev& d, int tile, const typename SVec<R>::T* s_in, R* s_slot, int max_slots, unsigned char*, int arrive_at, OnBoundary f,
                                                    const First&) {
        tet_tile_elements<R, MODE, PF, ET>(d, tile, s_in, s_slot, max_slots, arrive_at, f);
    }
};

// rotations[e] back in ORIGINAL element order (getRotations-style accessors, parity checks)
template <class R> __global__ void tet_export_rotations_kernel(TetDev<R> d, const uint32_t* __restrict__ orig, R* __restrict__ out) {
    const size_t es = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (es >= size_t(d.t.n_tiles) * d.t.tile_e) return;
    const uint32_t e = orig[es];
    if (e == 0xFFFFFFFFu) return;
    const Quad<R> q0 = d.rk0[es], q1 = d.rk1[es], q2 = d.rk2[es];
    R* o = out + 9 * size_t(e);
    o[0] = q0.a; o[1] = q0.b; o[2] = q0.c; o[3] = q0.d; o[4] = q1.a; o[5] = q1.b; o[6] = q1.c; o[7] = q1.d; o[8] = q2.a;
}

// computeVonMisesStress, TetrahedronFEMForceField.inl:2196-2360: one thread per element (tile order).  how = 1: strain of the corotational
// displacement D (rotation recomputed from x and written to rotations[e], as the reference does); how = 2: Green-Lagrange strain of
// U = x - x0.  shf = rows 1..3 of elemShapeFun in ORIGINAL element order, out = d_vonMisesPerElement in original order.
template <class R> __global__ void tet_von_mises_kernel(TetDev<R> d, const uint32_t* __restrict__ orig, const R* __restrict__ x, const R* __restrict__ rest,
                                                        const R* __restrict__ shf_all, const R* __restrict__ lambda, const R* __restrict__ mu, int how, int large,
                                                        R* __restrict__ out) {
    const size_t es = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (es >= size_t(d.t.n_tiles) * d.t.tile_e) return;
    const uint32_t e = orig[es];
    if (e == 0xFFFFFFFFu) return;
    const size_t tile = es / size_t(d.t.tile_e);
    const ushort4 ln = d.lnode[es];
    const uint32_t* tn = d.t.tile_nodes + d.t.tile_node_off[tile];
    const uint32_t g[4] = {tn[ln.x], tn[ln.y], tn[ln.z], tn[ln.w]};
    V3<R> P[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) P[n] = mk3<R>(x[3 * size_t(g[n])], x[3 * size_t(g[n]) + 1], x[3 * size_t(g[n]) + 2]);
    const R* shf = shf_all + 12 * size_t(e);
    R D[12];
    if (how == 2) {
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            D[3 * n] = P[n].x - rest[3 * size_t(g[n])]; D[3 * n + 1] = P[n].y - rest[3 * size_t(g[n]) + 1]; D[3 * n + 2] = P[n].z - rest[3 * size_t(g[n]) + 2];
        }
    } else {
        const Quad<R> xa = d.x0a[es], xb = d.x0b[es], xc = d.x0c[es];
        const R X0[12] = {xa.a, xa.b, xa.c, xa.d, xb.a, xb.b, xb.c, xb.d, xc.a, xc.b, xc.c, xc.d};
        M3<R> R02;
        if (large) {
            V3<R> ex = P[1] - P[0];
            normalize3(ex);
            V3<R> ey = P[2] - P[0];
            V3<R> ez = cross3(ex, ey);
            normalize3(ez);
            ey = cross3(ez, ex);
            set_row(R02, 0, ex); set_row(R02, 1, ey); set_row(R02, 2, ez);
        } else {
            M3<R> A;
            set_row(A, 0, P[1] - P[0]); set_row(A, 1, P[2] - P[0]); set_row(A, 2, P[3] - P[0]);
            polar_decomposition(A, R02);
        }
        const M3<R> rot = transpose(R02);
        const Quad<R> q2 = d.rk2[es];
        d.rk0[es] = Quad<R>{rot.m[0][0], rot.m[0][1], rot.m[0][2], rot.m[1][0]};
        d.rk1[es] = Quad<R>{rot.m[1][1], rot.m[1][2], rot.m[2][0], rot.m[2][1]};
        d.rk2[es] = Quad<R>{rot.m[2][2], q2.b, q2.c, q2.d};
        V3<R> def[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) def[n] = mul(R02, P[n]);
        if (large) {
            def[1].x -= def[0].x;
            def[2].x -= def[0].x;
            def[2].y -= def[0].y;
            def[3].x -= def[0].x; def[3].y -= def[0].y; def[3].z -= def[0].z;
            D[0] = 0; D[1] = 0; D[2] = 0;
            D[3] = X0[3] - def[1].x; D[4] = 0; D[5] = 0;
            D[6] = X0[6] - def[2].x; D[7] = X0[7] - def[2].y; D[8] = 0;
            D[9] = X0[9] - def[3].x; D[10] = X0[10] - def[3].y; D[11] = X0[11] - def[3].z;
        } else {
#pragma unroll
            for (int n = 0; n < 4; ++n) { D[3 * n] = X0[3 * n] - def[n].x; D[3 * n + 1] = X0[3 * n + 1] - def[n].y; D[3 * n + 2] = X0[3 * n + 2] - def[n].z; }
        }
    }
    // gradU(k,l) = sum_m shf(l+1,m) * D[3m+k], accumulated from 0.0 (:2236-2241, 2319-2324)
    R gu[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            R a = R(0);
#pragma unroll
            for (int m = 0; m < 4; ++m) a += shf[4 * l + m] * D[3 * m + k];
            gu[k][l] = a;
        }
    R st[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            R v = gu[i][j] + gu[j][i];
            if (how == 2) v = v + (gu[0][i] * gu[0][j] + gu[1][i] * gu[1][j] + gu[2][i] * gu[2][j]);   // + gradU^T gradU (Mat.h 3x3 product)
            st[i][j] = v * R(0.5);
        }
    const R vs[6] = {st[0][0], st[1][1], st[2][2], st[1][2], st[0][2], st[0][1]};
    const R lam = lambda[e], m_ = mu[e];
    R s[6];
    R tr = R(0);
#pragma unroll
    for (int k = 0; k < 3; ++k) { tr += vs[k]; s[k] = vs[k] * 2 * m_; }
#pragma unroll
    for (int k = 3; k < 6; ++k) s[k] = vs[k] * 2 * m_;
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] += lam * tr;
    R v = R(sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2] - s[0] * s[1] - s[1] * s[2] - s[2] * s[0] + 3 * s[3] * s[3] + 3 * s[4] * s[4] + 3 * s[5] * s[5]));
    if (double(v) < 1e-10) v = R(0);
    out[e] = v;
}
// d_vonMisesPerNode :2363-2372: mean of the incident elements' values, ascending element index
template <class R> __global__ void tet_von_mises_nodes_kernel(size_t n, const uint32_t* __restrict__ inc_off, const uint32_t* __restrict__ inc_e, const R* __restrict__ vme, R* __restrict__ out) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = inc_off[i], e = inc_off[i + 1];
    R a = R(0);
    for (uint32_t k = b; k < e; ++k) a += vme[inc_e[k]];
    if (e > b) a /= R(e - b);
    out[i] = a;
}

// getRotation / getRotations(VecReal&), TetrahedronFEMForceField.inl:781-833,2033-2042: per node, the mean of rotations[t] * R0(t) over the
// tetrahedra around it in ascending index, made orthogonal by polarDecomposition.  One thread per node; `inc` lists the tile-order slot
// of the incident elements and `r0t` holds _initialRotations in ORIGINAL element order.  A node without tetrahedra takes element
// _rotationIdx[node] = 0 (the array is zero-filled by resize, :1471).
// sibling != 0: TetrahedralCorotationalFEMForceField::getRotation (TetrahedralCorotationalFEMForceField.inl:779-820) instead -- the sum of
// rotation * initialTransformation (1: large, initialTransformation = R_0_1 = r0t transposed; 2: polar, = the rest edge matrix, r0t then holds it
// as is), scaled by Real(1.0f / n), and made orthogonal by normalising rows 0 and 1 and two cross products (no polar decomposition).
template <class R> __global__ void tet_node_rotations_kernel(TetDev<R> d, const uint32_t* __restrict__ inc_off, const uint32_t* __restrict__ inc_es,
                                                             const uint32_t* __restrict__ inc_e, const R* __restrict__ r0t, uint32_t es_of_first, R* __restrict__ out,
                                                             int sibling = 0) {
    const size_t n = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (n >= size_t(d.t.n_nodes)) return;
    if (sibling) {
        M3<R> acc;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc.m[i][j] = R(0);
        const uint32_t b = inc_off[n], e = inc_off[n + 1];
        for (uint32_t k = b; k < e; ++k) {
            const uint32_t es = inc_es[k];
            const Quad<R> q0 = d.rk0[es], q1 = d.rk1[es], q2 = d.rk2[es];
            M3<R> rot, r01;
            rot.m[0][0] = q0.a; rot.m[0][1] = q0.b; rot.m[0][2] = q0.c; rot.m[1][0] = q0.d; rot.m[1][1] = q1.a; rot.m[1][2] = q1.b;
            rot.m[2][0] = q1.c; rot.m[2][1] = q1.d; rot.m[2][2] = q2.a;
            const R* p = r0t + 9 * size_t(inc_e[k]);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) r01.m[i][j] = sibling == 1 ? p[3 * j + i] : p[3 * i + j];
            const M3<R> pr = mul(rot, r01);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) acc.m[i][j] += pr.m[i][j];
        }
        const R sc = R(1.0f / float(int(e - b)));
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc.m[i][j] = acc.m[i][j] * sc;
        V3<R> ex = row(acc, 0), ey = row(acc, 1);
        normalize3(ex);
        normalize3(ey);
        V3<R> ez = cross3(ex, ey);
        normalize3(ez);
        ey = cross3(ez, ex);
        normalize3(ey);
        R* o = out + 9 * n;
        o[0] = ex.x; o[1] = ex.y; o[2] = ex.z; o[3] = ey.x; o[4] = ey.y; o[5] = ey.z; o[6] = ez.x; o[7] = ez.y; o[8] = ez.z;
        return;
    }
    auto elem_product = [&](uint32_t es, uint32_t e) {
        const Quad<R> q0 = d.rk0[es], q1 = d.rk1[es], q2 = d.rk2[es];
        M3<R> rot, r0;
        rot.m[0][0] = q0.a; rot.m[0][1] = q0.b; rot.m[0][2] = q0.c; rot.m[1][0] = q0.d; rot.m[1][1] = q1.a; rot.m[1][2] = q1.b;
        rot.m[2][0] = q1.c; rot.m[2][1] = q1.d; rot.m[2][2] = q2.a;
        const R* p = r0t + 9 * size_t(e);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) r0.m[i][j] = p[3 * j + i];   // R0t.transpose(_initialRotations[t])
        return mul(rot, r0);
    };
    const uint32_t b = inc_off[n], e = inc_off[n + 1];
    M3<R> acc;
    if (b == e) {
        if (d.t.n_elems > 0) acc = elem_product(es_of_first, 0u);
        else {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) acc.m[i][j] = i == j ? R(1) : R(0);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc.m[i][j] = R(0);
        for (uint32_t k = b; k < e; ++k) {
            const M3<R> pr = elem_product(inc_es[k], inc_e[k]);
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
        acc = q;
    }
    R* o = out + 9 * n;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o[3 * i + j] = acc.m[i][j];
}


}  // namespace sb
