// TEST INFRASTRUCTURE.  C ABI over the REFERENCE's own templates, compiled by
// oracle/build_ref.sh against the sources under /root/reference (never copied here).
// Used only by tests/ to pin the oracle's restated math (oracle/sofa_oracle.cpp)
// bit-for-bit against the reference's object code:
//   sofa::helper::Decompose<Real>  Sofa/framework/Helper/src/sofa/helper/decompose.inl:672-723,755-764,1663-1829
//   sofa::type::Mat / Vec          Sofa/framework/Type/src/sofa/type/Mat.h, Vec.h
//   sofa::geometry::Tetrahedron    Sofa/framework/Geometry/src/sofa/geometry/Tetrahedron.h:55-82
#include <sofa/type/Mat.h>
#include <sofa/type/Vec.h>
#include <sofa/helper/decompose.h>
#include <sofa/geometry/Tetrahedron.h>

namespace {
template <class R> using M3 = sofa::type::Mat<3, 3, R>;
template <class R> using V3 = sofa::type::Vec<3, R>;

template <class R> M3<R> ld(const R* p) {
    M3<R> m;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(i, j) = p[3 * i + j];
    return m;
}
template <class R> void st(const M3<R>& m, R* p) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) p[3 * i + j] = m(i, j);
}
template <class R> V3<R> ldv(const R* p) { return V3<R>(p[0], p[1], p[2]); }
template <class R> void stv(const V3<R>& v, R* p) { p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; }

template <class R> R polar(const R* M, R* Q) {
    M3<R> q;
    R det = sofa::helper::Decompose<R>::polarDecomposition(ld(M), q);
    st(q, Q);
    return det;
}
template <class R> int polar_stable(const R* M, R* Q) {
    M3<R> q;
    bool deg = sofa::helper::Decompose<R>::polarDecomposition_stable(ld(M), q);
    st(q, Q);
    return deg ? 1 : 0;
}
template <class R> int svd_stable(const R* F, R* U, R* S, R* V) {
    M3<R> u, v; V3<R> s;
    bool deg = sofa::helper::Decompose<R>::SVD_stable(ld(F), u, s, v);
    st(u, U); st(v, V); stv(s, S);
    return deg ? 1 : 0;
}
template <class R> int invert(const R* A, R* Ainv) {
    M3<R> inv;
    bool ok = inv.invert(ld(A));
    st(inv, Ainv);
    return ok ? 1 : 0;
}
template <class R> void frame_large(const R* a, const R* b, const R* c, R* Rout) {
    // same expression sequence as TetrahedronFEMForceField<DT>::computeRotationLarge
    // (TetrahedronFEMForceField.inl:755-778) written on the reference's Vec type.
    const V3<R> edgex = (ldv(b) - ldv(a)).normalized();
    V3<R> edgey = ldv(c) - ldv(a);
    const V3<R> edgez = cross(edgex, edgey).normalized();
    edgey = cross(edgez, edgex);
    for (int j = 0; j < 3; ++j) { Rout[j] = edgex[j]; Rout[3 + j] = edgey[j]; Rout[6 + j] = edgez[j]; }
}
}  // namespace

#define SHIM(SFX, R)                                                                              \
    extern "C" R ref_polar_##SFX(const R* M, R* Q) { return polar<R>(M, Q); }                   \
    extern "C" int ref_polar_stable_##SFX(const R* M, R* Q) { return polar_stable<R>(M, Q); }   \
    extern "C" int ref_svd_stable_##SFX(const R* F, R* U, R* S, R* V) { return svd_stable<R>(F, U, S, V); } \
    extern "C" int ref_mat3_invert_##SFX(const R* A, R* Ai) { return invert<R>(A, Ai); }         \
    extern "C" R ref_mat3_det_##SFX(const R* A) { return sofa::type::determinant(ld(A)); }        \
    extern "C" void ref_mat3_mul_##SFX(const R* A, const R* B, R* C) { st(ld(A) * ld(B), C); }    \
    extern "C" void ref_mat3_mul_transposed_##SFX(const R* A, const R* B, R* C) {                  \
        st(ld(A).multTransposed(ld(B)), C); }                                                     \
    extern "C" void ref_mat3_vec_##SFX(const R* A, const R* v, R* r) { stv(ld(A) * ldv(v), r); }  \
    extern "C" void ref_mat3_tvec_##SFX(const R* A, const R* v, R* r) {                            \
        stv(ld(A).multTranspose(ldv(v)), r); }                                                    \
    extern "C" void ref_frame_large_##SFX(const R* a, const R* b, const R* c, R* Ro) {             \
        frame_large<R>(a, b, c, Ro); }                                                            \
    extern "C" R ref_tet_volume_##SFX(const R* a, const R* b, const R* c, const R* d) {            \
        return sofa::geometry::Tetrahedron::volume(ldv(a), ldv(b), ldv(c), ldv(d)); }

SHIM(f, float)
SHIM(d, double)
