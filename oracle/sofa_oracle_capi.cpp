// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  C ABI over oracle/sofa_oracle.hpp for ctypes
// (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference only).
// `real`: 0 = float (Vec3f build of the reference), 1 = double (Vec3d).  State arrays are
// AoS Vec3 of that Real; indices are uint32 (sofa::Index).
#include "sofa_oracle.hpp"

#include <atomic>
#include <mutex>
#include <string>
#include <thread>

using namespace orc;

namespace {
template <class R> std::vector<Vec3<R>> toVec(const void* p, size_t n) {
    std::vector<Vec3<R>> v(n);
    if (p && n) std::memcpy(v.data(), p, n * sizeof(Vec3<R>));
    return v;
}
template <class R> void fromVec(const std::vector<Vec3<R>>& v, void* p) { if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(Vec3<R>)); }

struct SceneAny {
    int real;
    Scene<float> f;
    Scene<double> d;
};
template <class R> Scene<R>& S(SceneAny* s);
template <> Scene<float>& S<float>(SceneAny* s) { return s->f; }
template <> Scene<double>& S<double>(SceneAny* s) { return s->d; }

#define DISPATCH(h, ...)                         \
    do {                                         \
        SceneAny* s_ = static_cast<SceneAny*>(h);\
        if (s_->real == 0) { typedef float R; Scene<R>& sc = s_->f; (void)sc; __VA_ARGS__; } \
        else { typedef double R; Scene<R>& sc = s_->d; (void)sc; __VA_ARGS__; }            \
    } while (0)

void fillReal(std::vector<float>& dst, const double* src, int n) { dst.resize(n); for (int i = 0; i < n; ++i) dst[i] = float(src[i]); }
void fillReal(std::vector<double>& dst, const double* src, int n) { dst.assign(src, src + n); }

template <class R> void copyOut(const std::vector<R>& v, void* out) { if (!v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(R)); }
template <class R> void copyMats(const std::vector<Mat3<R>>& v, void* out) { if (!v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(Mat3<R>)); }

}  // namespace

// PlaneForceField stand-alone: prm = {nx, ny, nz, d (both before setPlane's normalisation), stiffness, damping, maxForce, bilateral}
// addForce: f += ..., contacts_out[i] = 1 for the nodes in contact; addDForce uses the contacts of the given flags
namespace {
template <class R> void planeRun(size_t n, const double* prm, void* f, const void* x, const void* v, unsigned char* contacts, const void* dx, double kf, int dforce) {
    PlaneForceField<R> pf;
    pf.stiffness = R(prm[4]); pf.damping = R(prm[5]); pf.maxForce = R(prm[6]); pf.bilateral = prm[7] != 0;
    pf.setPlane(Vec3<R>(R(prm[0]), R(prm[1]), R(prm[2])), R(prm[3]));
    VecDeriv<R> F = toVec<R>(f, n);
    if (!dforce) {
        pf.addForce(F, toVec<R>(x, n), toVec<R>(v, n));
        std::memset(contacts, 0, n);
        for (uint32_t i : pf.contacts) contacts[i] = 1;
    } else {
        for (size_t i = 0; i < n; ++i) if (contacts[i]) pf.contacts.push_back(uint32_t(i));
        pf.addDForce(F, toVec<R>(dx, n), kf);
    }
    fromVec(F, f);
}
}  // namespace
extern "C" {

// ---- mesh generation ------------------------------------------------------------------------------------------
void orc_grid(int nx, int ny, int nz, const double* mn, const double* mx, double* pos_out, uint32_t* hex_out) {
    GridMesh g = regularGrid(nx, ny, nz, mn, mx);
    std::memcpy(pos_out, g.pos.data(), g.pos.size() * sizeof(double));
    if (hex_out && !g.hexas.empty()) std::memcpy(hex_out, g.hexas.data(), g.hexas.size() * sizeof(uint32_t));
}
void orc_hexas_to_tetras(int nx, int ny, int nz, int mode, uint32_t* tets_out) {
    const double z[3] = {0, 0, 0}, o[3] = {1, 1, 1};
    GridMesh g = regularGrid(nx, ny, nz, z, o);
    std::vector<uint32_t> t = hexasToTetras(g, mode);
    std::memcpy(tets_out, t.data(), t.size() * sizeof(uint32_t));
}

// ---- restated sofa::type / helper::Decompose math, for bit-for-bit comparison with oracle/_ref ----------------
#define MATH(SFX, R)                                                                                                  \
    R orc_polar_##SFX(const R* M, R* Q) { Mat3<R> m, q; std::memcpy(m.m, M, sizeof(m.m)); R det = Decompose<R>::polarDecomposition(m, q); std::memcpy(Q, q.m, sizeof(q.m)); return det; } \
    int orc_polar_stable_##SFX(const R* M, R* Q) { Mat3<R> m, q; std::memcpy(m.m, M, sizeof(m.m)); bool d = Decompose<R>::polarDecomposition_stable(m, q); std::memcpy(Q, q.m, sizeof(q.m)); return d; } \
    int orc_svd_stable_##SFX(const R* F, R* U, R* Sd, R* V) { Mat3<R> f, u, v; Vec3<R> s; std::memcpy(f.m, F, sizeof(f.m)); bool d = Decompose<R>::SVD_stable(f, u, s, v); std::memcpy(U, u.m, sizeof(u.m)); std::memcpy(V, v.m, sizeof(v.m)); std::memcpy(Sd, s.v, sizeof(s.v)); return d; } \
    int orc_mat3_invert_##SFX(const R* A, R* Ai) { Mat3<R> a, i; std::memcpy(a.m, A, sizeof(a.m)); bool ok = invertMatrix(i, a); std::memcpy(Ai, i.m, sizeof(i.m)); return ok; } \
    R orc_mat3_det_##SFX(const R* A) { Mat3<R> a; std::memcpy(a.m, A, sizeof(a.m)); return determinant(a); }           \
    void orc_mat3_mul_##SFX(const R* A, const R* B, R* C) { Mat3<R> a, b; std::memcpy(a.m, A, sizeof(a.m)); std::memcpy(b.m, B, sizeof(b.m)); Mat3<R> c = a * b; std::memcpy(C, c.m, sizeof(c.m)); } \
    void orc_mat3_mul_transposed_##SFX(const R* A, const R* B, R* C) { Mat3<R> a, b; std::memcpy(a.m, A, sizeof(a.m)); std::memcpy(b.m, B, sizeof(b.m)); Mat3<R> c = a.multTransposed(b); std::memcpy(C, c.m, sizeof(c.m)); } \
    void orc_mat3_vec_##SFX(const R* A, const R* v, R* r) { Mat3<R> a; std::memcpy(a.m, A, sizeof(a.m)); Vec3<R> x(v[0], v[1], v[2]); Vec3<R> y = a * x; std::memcpy(r, y.v, sizeof(y.v)); } \
    void orc_mat3_tvec_##SFX(const R* A, const R* v, R* r) { Mat3<R> a; std::memcpy(a.m, A, sizeof(a.m)); Vec3<R> x(v[0], v[1], v[2]); Vec3<R> y = a.multTranspose(x); std::memcpy(r, y.v, sizeof(y.v)); } \
    void orc_frame_large_##SFX(const R* a, const R* b, const R* c, R* Ro) { std::vector<Vec3<R>> p = {Vec3<R>(a[0], a[1], a[2]), Vec3<R>(b[0], b[1], b[2]), Vec3<R>(c[0], c[1], c[2])}; Mat3<R> r; TetFEM<R>::computeRotationLarge(r, p, 0, 1, 2); std::memcpy(Ro, r.m, sizeof(r.m)); } \
    R orc_tet_volume_##SFX(const R* a, const R* b, const R* c, const R* d) { return DiagonalMass<R>::tetVol(Vec3<R>(a[0], a[1], a[2]), Vec3<R>(b[0], b[1], b[2]), Vec3<R>(c[0], c[1], c[2]), Vec3<R>(d[0], d[1], d[2])); }
MATH(f, float)
MATH(d, double)

// ---- MechanicalObject vector ops on raw arrays (aliasing by pointer identity, null = no operand) ---------------
void orc_plane(int real, size_t n, const double* prm, void* f, const void* x, const void* v, unsigned char* contacts, const void* dx, double kf, int dforce) {
    if (real == 0) planeRun<float>(n, prm, f, x, v, contacts, dx, kf, dforce); else planeRun<double>(n, prm, f, x, v, contacts, dx, kf, dforce);
}
void orc_vop(int real, size_t n, void* r, const void* a, const void* b, double k) {
    auto run = [&](auto tag) {
        typedef decltype(tag) R;
        VecDeriv<R> vr = toVec<R>(r, n), va, vb;
        const VecDeriv<R>* pa = nullptr; const VecDeriv<R>* pb = nullptr;
        if (a) { if (a == r) pa = &vr; else { va = toVec<R>(a, n); pa = &va; } }
        if (b) { if (b == r) pb = &vr; else if (b == a && pa != &vr) pb = pa; else { vb = toVec<R>(b, n); pb = &vb; } }
        VOps<R>::vOp(&vr, pa, pb, k);
        fromVec(vr, r);
    };
    if (real == 0) run(float()); else run(double());
}
double orc_vdot(int real, size_t n, const void* a, const void* b) {
    if (real == 0) return VOps<float>::dot(toVec<float>(a, n), toVec<float>(b, n));
    return VOps<double>::dot(toVec<double>(a, n), toVec<double>(b, n));
}

// ---- scene ------------------------------------------------------------------------------------------------------
void* orc_scene_create(int real) { SceneAny* s = new SceneAny(); s->real = real; return s; }
void orc_scene_destroy(void* h) { delete static_cast<SceneAny*>(h); }
void orc_scene_set_dot_double(void* h, int on) { DISPATCH(h, { sc.dotDouble = (on & 1) != 0; sc.dotReverse = (on & 2) != 0; }); }
void orc_scene_set_threads(void* h, int n) { DISPATCH(h, { sc.threads = n < 1 ? 1 : n; }); }

// positions (also the rest positions) and velocities
void orc_scene_set_state(void* h, size_t n, const void* x, const void* v) {
    DISPATCH(h, { sc.x = toVec<R>(x, n); sc.x0 = sc.x; sc.v = toVec<R>(v, n); if (!v) sc.v.assign(n, Vec3<R>()); sc.f.assign(n, Vec3<R>()); sc.dx.assign(n, Vec3<R>()); });
}
void orc_scene_set_external_force(void* h, const void* f) { DISPATCH(h, { if (f) sc.externalForce = toVec<R>(f, sc.x.size()); else sc.externalForce.clear(); }); }
void orc_scene_set_x(void* h, const void* x) { DISPATCH(h, { sc.x = toVec<R>(x, sc.x.size()); }); }
void orc_scene_set_v(void* h, const void* v) { DISPATCH(h, { sc.v = toVec<R>(v, sc.v.size()); }); }


// TetrahedronFEMForceField: method 0 small, 1 large, 2 polar, 3 svd; reinit() on the rest positions
void orc_scene_set_tets(void* h, size_t T, const uint32_t* tets, int method, int ny, const double* young, int np, const double* poisson,
                        int nlsf, const double* lsf) {
    DISPATCH(h, {
        sc.hasTet = true; sc.tet.method = method; sc.tet.tets.assign(tets, tets + 4 * T);
        fillReal(sc.tet.young, young, ny); fillReal(sc.tet.poisson, poisson, np);
        if (nlsf > 0) fillReal(sc.tet.localStiffnessFactor, lsf, nlsf); else sc.tet.localStiffnessFactor.clear();
        sc.tet.reinit(sc.x0);
    });
}
// FastTetrahedralCorotationalForceField: method 0 polar, 1 qr, 2 polar2, 3 none; E > 0: the topology's own edge list
void orc_scene_set_fast_tets(void* h, size_t T, const uint32_t* tets, int method, int ny, const double* young, int np, const double* poisson, size_t E, const uint32_t* edges) {
    DISPATCH(h, {
        sc.hasFast = true; sc.fast.method = method; sc.fast.tets.assign(tets, tets + 4 * T);
        fillReal(sc.fast.young, young, ny); fillReal(sc.fast.poisson, poisson, np);
        std::vector<uint32_t> given; if (E > 0) given.assign(edges, edges + 2 * E);
        sc.fast.init(sc.x0, E > 0 ? &given : nullptr);
    });
}
// HexahedronFEMForceField: method 0 large, 1 polar, 2 small
void orc_scene_set_hexas(void* h, size_t H, const uint32_t* hexas, int method, int ny, const double* young, int np, const double* poisson) {
    DISPATCH(h, {
        sc.hasHex = true; sc.hex.method = method; sc.hex.hexas.assign(hexas, hexas + 8 * H);
        fillReal(sc.hex.young, young, ny); fillReal(sc.hex.poisson, poisson, np);
        sc.hex.reinit(sc.x0);
    });
}
// DiagonalMass: kind 0 = massDensity, 1 = totalMass, lumped over `elems` (elemSize 4 or 8); kind 2 = explicit vertexMass array;
// UniformMass: kind 3 = vertexMass, kind 4 = totalMass
void orc_scene_set_mass(void* h, int kind, double value, size_t nelems, const uint32_t* elems, int elemSize, const void* vertexMass) {
    DISPATCH(h, {
        sc.hasMass = true;
        sc.mass.uniform = false;
        if (kind == 3) sc.mass.initUniformFromVertexMass(R(value), sc.x.size());
        else if (kind == 4) sc.mass.initUniformFromTotalMass(value, sc.x.size());
        else if (kind == 2) { sc.mass.vertexMass.resize(sc.x.size()); std::memcpy(sc.mass.vertexMass.data(), vertexMass, sc.x.size() * sizeof(R)); }
        else {
            std::vector<uint32_t> e(elems, elems + nelems * elemSize);
            if (kind == 0) sc.mass.initFromMassDensity(R(value), sc.x0, e, elemSize);
            else sc.mass.initFromTotalMass(R(value), sc.x0, e, elemSize);
        }
    });
}
// PlaneForceField of the node: prm as for orc_plane, + its rayleighStiffness
// the node's mass becomes a MeshMatrixMass over the tetrahedra (massDensity, lumping)
void orc_scene_set_mesh_mass(void* h, size_t T, const uint32_t* tets, double density, int lumping) {
    DISPATCH(h, {
        std::vector<uint32_t> t(tets, tets + 4 * T);
        sc.hasMass = true; sc.hasMeshMass = true;
        sc.meshMass.initFromMassDensityTets(R(density), sc.x0, t, lumping != 0);
    });
}
void orc_scene_set_plane(void* h, const double* prm, double rayleigh) {
    DISPATCH(h, {
        sc.hasPlane = true; sc.planeRayleighStiffness = rayleigh;
        sc.plane.stiffness = R(prm[4]); sc.plane.damping = R(prm[5]); sc.plane.maxForce = R(prm[6]); sc.plane.bilateral = prm[7] != 0;
        sc.plane.setPlane(Vec3<R>(R(prm[0]), R(prm[1]), R(prm[2])), R(prm[3]));
    });
}
void orc_scene_set_fixed(void* h, size_t n, const uint32_t* idx, int fixAll) {
    DISPATCH(h, { sc.fixed.assign(idx, idx + n); sc.fixAll = fixAll != 0; });
}
// params: [0..2] gravity, 3 dt, 4 rayleighStiffness, 5 rayleighMass, 6 vdamping, 7 firstOrder, 8 trapezoidal,
//         9 iterations, 10 tolerance, 11 threshold, 12 warmStart, 13 massFirst, 14 ff.rayleighStiffness, 15 mass.rayleighMass
void orc_scene_set_params(void* h, const double* p) {
    DISPATCH(h, {
        sc.gravity[0] = p[0]; sc.gravity[1] = p[1]; sc.gravity[2] = p[2];
        sc.dt = p[3]; sc.rayleighStiffness = p[4]; sc.rayleighMass = p[5]; sc.vdamping = p[6];
        sc.firstOrder = p[7] != 0; sc.trapezoidal = p[8] != 0;
        sc.maxIter = unsigned(p[9]); sc.tolerance = p[10]; sc.threshold = p[11]; sc.warmStart = p[12] != 0;
        sc.massFirst = p[13] != 0; sc.ffRayleighStiffness = p[14]; sc.massRayleighMass = p[15];
    });
}
// One EulerImplicitSolver::solve.  Returns the "CG iterations" value (nb_iter).
int orc_scene_step(void* h) {
    SceneAny* s = static_cast<SceneAny*>(h);
    int it = 0;
    DISPATCH(h, { sc.step(); it = int(sc.lastIter); });
    (void)s;
    return it;
}
int orc_scene_end_condition(void* h) { int e = 0; DISPATCH(h, { e = sc.endCond; }); return e; }
void orc_scene_reset_timestep_count(void* h) { DISPATCH(h, { sc.timeStepCount = 0; }); }

// f += addForce(x) for the FEM force field alone (f in/out)
void orc_scene_fem_add_force(void* h, void* f_inout, const void* x) {
    DISPATCH(h, {
        const size_t n = sc.x.size();
        std::vector<Vec3<R>> saved = sc.x; sc.x = toVec<R>(x, n);
        VecDeriv<R> F = toVec<R>(f_inout, n);
        sc.femAddForce(F);
        fromVec(F, f_inout); sc.x = saved;
    });
}
// df += addDForce(dx) with kFactorIncludingRayleighDamping = kFactor (df in/out)
void orc_scene_fem_add_dforce(void* h, void* df_inout, const void* dx, double kFactor) {
    DISPATCH(h, {
        const size_t n = sc.x.size();
        VecDeriv<R> df = toVec<R>(df_inout, n), d = toVec<R>(dx, n);
        sc.femAddDForce(df, d, kFactor);
        fromVec(df, df_inout);
    });
}
// mop.computeForce: f = sum of addForce of the node's force fields at the scene's current x
void orc_scene_compute_force(void* h, void* f_out) { DISPATCH(h, { VecDeriv<R> F; sc.computeForce(F); fromVec(F, f_out); }); }
// q = project( (m M + b B + k K) p )  -- GraphScatteredMatrix::apply
void orc_scene_apply(void* h, void* q_out, const void* p, double m, double b, double k) {
    DISPATCH(h, {
        const size_t n = sc.x.size();
        sc.mFact = m; sc.bFact = b; sc.kFact = k;
        VecDeriv<R> q, pp = toVec<R>(p, n);
        sc.applyA(q, pp);
        fromVec(q, q_out);
    });
}
// CG on the system with factors (m,b,k); returns nb_iter.  x_inout is the initial guess when warmStart.
int orc_scene_cg(void* h, void* x_inout, const void* b, double m, double bf, double k) {
    int it = 0;
    DISPATCH(h, {
        const size_t n = sc.x.size();
        sc.mFact = m; sc.bFact = bf; sc.kFact = k;
        VecDeriv<R> X = toVec<R>(x_inout, n), B = toVec<R>(b, n);
        sc.cgSolve(X, B);
        fromVec(X, x_inout); it = int(sc.lastIter);
    });
    return it;
}
size_t orc_scene_graph(void* h, int which, double* out, size_t cap) {
    size_t n = 0;
    DISPATCH(h, { const std::vector<double>& g = which == 0 ? sc.graphError : sc.graphDen; n = g.size(); for (size_t i = 0; i < n && i < cap; ++i) out[i] = g[i]; });
    return n;
}
// Named array getter.  Returns the number of Real (or uint32) scalars written; call with out = nullptr for the count.
size_t orc_scene_get(void* h, const char* what, void* out) {
    const std::string w(what);
    size_t cnt = 0;
    DISPATCH(h, {
        auto vec3 = [&](const std::vector<Vec3<R>>& v) { cnt = 3 * v.size(); if (out) fromVec(v, out); };
        auto mats = [&](const std::vector<Mat3<R>>& v) { cnt = 9 * v.size(); if (out) copyMats(v, out); };
        auto reals = [&](const std::vector<R>& v) { cnt = v.size(); if (out) copyOut(v, out); };
        if (w == "x") vec3(sc.x); else if (w == "v") vec3(sc.v); else if (w == "f") vec3(sc.lastForce);
        else if (w == "b") vec3(sc.lastB); else if (w == "sol") vec3(sc.lastSol); else if (w == "x0") vec3(sc.x0);
        else if (w == "vertexMass") reals(sc.mass.vertexMass);
        else if (w == "tet.rotations") mats(sc.tet.rotations); else if (w == "tet.initialRotations") mats(sc.tet.initialRotations);
        else if (w == "tet.initialTransformation") mats(sc.tet.initialTransformation);
        else if (w == "tet.plasticStrains") reals(sc.tet.plasticStrains);
        else if (w == "tet.elemShapeFun") reals(sc.tet.elemShapeFun);
        else if (w == "tet.J") reals(sc.tet.J); else if (w == "tet.Jsh") reals(sc.tet.Jsh); else if (w == "tet.K") reals(sc.tet.K); else if (w == "tet.X0") vec3(sc.tet.X0);
        else if (w.rfind("fast.", 0) == 0) {
            std::vector<Mat3<R>> m; std::vector<Vec3<R>> v; std::vector<R> r;
            const auto& ti = sc.fast.tetrahedronInfo;
            if (w == "fast.rotations") { for (auto& t : ti) m.push_back(t.rotation); mats(m); }
            else if (w == "fast.restRotations") { for (auto& t : ti) m.push_back(t.restRotation); mats(m); }
            else if (w == "fast.linearDfDx") { for (auto& t : ti) for (int j = 0; j < 6; ++j) m.push_back(t.linearDfDx[j]); mats(m); }
            else if (w == "fast.linearDfDxDiag") { for (auto& t : ti) for (int j = 0; j < 4; ++j) m.push_back(t.linearDfDxDiag[j]); mats(m); }
            else if (w == "fast.shapeVectors") { for (auto& t : ti) for (int j = 0; j < 4; ++j) v.push_back(t.shapeVector[j]); vec3(v); }
            else if (w == "fast.restEdgeVectors") { for (auto& t : ti) for (int j = 0; j < 6; ++j) v.push_back(t.restEdgeVector[j]); vec3(v); }
            else if (w == "fast.edgeOrientations") { for (auto& t : ti) for (int j = 0; j < 6; ++j) r.push_back(t.edgeOrientation[j]); reals(r); }
            else if (w == "fast.edgeInfo") mats(sc.fast.edgeInfo);
            else if (w == "fast.edges") { for (uint32_t e : sc.fast.edges) r.push_back(R(e)); reals(r); }
        }
        else if (w == "hex.rotations") mats(sc.hex.rotations); else if (w == "hex.initialRotations") mats(sc.hex.initialRotations);
        else if (w == "hex.Ke") reals(sc.hex.Ke); else if (w == "hex.X0") vec3(sc.hex.X0); else if (w == "hex.Kmat") reals(sc.hex.Kmat);
    });
    return cnt;
}
// full reference-shaped matrices of one tetrahedron (for the golden vectors): J 12x6, K 6x6, row-major, as double
void orc_scene_tet_matrices(void* h, size_t e, double* J72, double* K36) {
    DISPATCH(h, {
        R j[72], k[36];
        sc.tet.strainDisplacementMatrix(e, j); sc.tet.materialStiffnessMatrix(e, k);
        for (int i = 0; i < 72; ++i) J72[i] = j[i];
        for (int i = 0; i < 36; ++i) K36[i] = k[i];
    });
}
// MeshMatrixMass on tetrahedra, stand-alone.  orc_meshmass_create returns a handle; arrays are copied out with the getters.
namespace {
struct MeshMassAny { int real; MeshMatrixMass<float> f; MeshMatrixMass<double> d; };
}
void* orc_meshmass_create(int real, size_t n, const void* pos, size_t T, const uint32_t* tets, double density, int lumping) {
    MeshMassAny* m = new MeshMassAny(); m->real = real;
    std::vector<uint32_t> t(tets, tets + 4 * T);
    if (real == 0) m->f.initFromMassDensityTets(float(density), toVec<float>(pos, n), t, lumping != 0);
    else m->d.initFromMassDensityTets(density, toVec<double>(pos, n), t, lumping != 0);
    return m;
}
void orc_meshmass_destroy(void* h) { delete static_cast<MeshMassAny*>(h); }
size_t orc_meshmass_n_edges(void* h) { MeshMassAny* m = static_cast<MeshMassAny*>(h); return (m->real == 0 ? m->f.edges.size() : m->d.edges.size()) / 2; }
double orc_meshmass_total(void* h) { MeshMassAny* m = static_cast<MeshMassAny*>(h); return m->real == 0 ? m->f.totalMass : m->d.totalMass; }
void orc_meshmass_arrays(void* h, uint32_t* edges, void* vertexMass, void* edgeMass) {
    MeshMassAny* m = static_cast<MeshMassAny*>(h);
    if (m->real == 0) { copyOut(m->f.edges, edges); copyOut(m->f.vertexMass, vertexMass); copyOut(m->f.edgeMass, edgeMass); }
    else { copyOut(m->d.edges, edges); copyOut(m->d.vertexMass, vertexMass); copyOut(m->d.edgeMass, edgeMass); }
}
// op 0: addMDx(res, dx, factor); 1: addForce(res, gravity = g); 2: accFromF(res = a, dx = f) -> returns 0 when refused
int orc_meshmass_op(void* h, int op, size_t n, void* res, const void* dx, double factor, const double* g) {
    MeshMassAny* m = static_cast<MeshMassAny*>(h);
    int ok = 1;
    auto run = [&](auto& mm, auto tag) {
        typedef decltype(tag) R;
        VecDeriv<R> r = toVec<R>(res, n);
        if (op == 0) mm.addMDx(r, toVec<R>(dx, n), factor);
        else if (op == 1) mm.addForce(r, g);
        else ok = mm.accFromF(r, toVec<R>(dx, n)) ? 1 : 0;
        fromVec(r, res);
    };
    if (m->real == 0) run(m->f, float()); else run(m->d, double());
    return ok;
}
// plasticMaxThreshold / plasticYieldThreshold / plasticCreep; call before or after set_tets (clears the plastic strains)
void orc_scene_tet_set_plastic(void* h, double mx, double yield, double creep) {
    DISPATCH(h, { sc.tet.plastic[0] = R(mx); sc.tet.plastic[1] = R(yield); sc.tet.plastic[2] = R(creep); sc.tet.reset(); });
}
void orc_scene_tet_set_update_stiffness(void* h, int on) { DISPATCH(h, { sc.tet.updateStiffnessMatrix = on != 0; }); }
void orc_scene_tet_set_sibling(void* h, int on) { DISPATCH(h, { sc.tet.tetrahedralCorotational = on != 0; }); }
void orc_scene_tet_reset(void* h) { DISPATCH(h, { sc.tet.reset(); }); }
// computeVonMisesStress at positions x (how = 1 or 2): per_element T Reals, per_node N Reals
void orc_scene_tet_von_mises(void* h, const void* x, int how, void* per_element, void* per_node) {
    DISPATCH(h, {
        sc.tet.computeVonMisesStress(toVec<R>(x, sc.x.size()), how);
        copyOut(sc.tet.vonMisesPerElement, per_element); copyOut(sc.tet.vonMisesPerNode, per_node);
    });
}
// TetrahedronFEMForceField::getRotations(VecReal&): out = 9 Reals per node
void orc_scene_tet_get_rotations(void* h, void* out) {
    DISPATCH(h, { std::vector<Mat3<R>> v; sc.tet.getRotations(v, sc.x.size()); copyMats(v, out); });
}
// HexahedronFEMForceField::getNodeRotation for every node: out = 9 Reals per node
void orc_scene_hex_get_rotations(void* h, void* out) {
    DISPATCH(h, { std::vector<Mat3<R>> v; sc.hex.getRotations(v, sc.x.size()); copyMats(v, out); });
}
double orc_scene_hex_potential_energy(void* h) { double e = 0; DISPATCH(h, { e = sc.hex.potentialEnergy; }); return e; }

}  // extern "C"
