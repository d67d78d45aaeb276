// ============================================================================
// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// CPU restatement ("oracle") of SOFA's implicit-dynamics FEM hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, load or call this.  The product
// (sofa_b200/csrc) never includes or links anything in oracle/.
//
// Parity status: PINNED.  (1) the restated sofa::type / helper::Decompose math is
// checked bit-for-bit against the reference's own object code (oracle/_ref,
// built by oracle/build_ref.sh from /root/reference) in tests/test_oracle_ref.py;
// (2) the component restatement is checked against the golden vectors of the
// reference's own tests (tests/test_oracle_golden.py):
//   Sofa/Component/SolidMechanics/FEM/Elastic/tests/BaseTetrahedronFEMForceField_test.h:290-315,379-431
//   Sofa/Component/SolidMechanics/FEM/Elastic/tests/TetrahedronFEMForceField_stepTest.cpp:53-83
//   Sofa/Component/SolidMechanics/FEM/Elastic/tests/HexahedronFEMForceField_test.cpp:55-91
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference).  Arithmetic is written operation-for-operation in the
// reference's order, templated on Real (float = Vec3f build, double = Vec3d),
// and must be compiled WITHOUT fp contraction or fast-math (see Makefile).
// ============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>
#include <map>
#include <vector>

namespace orc {

typedef double SReal;  // Sofa/framework/Config/src/sofa/config.h.in (SOFA_FLOAT undefined)

// ---------------------------------------------------------------------------
// sofa::type::Vec<3,Real>   Sofa/framework/Type/src/sofa/type/Vec.h
// ---------------------------------------------------------------------------
template <class R> struct Vec3 {
    R v[3];
    Vec3() : v{0, 0, 0} {}
    Vec3(R a, R b, R c) : v{a, b, c} {}
    R& operator[](int i) { return v[i]; }
    const R& operator[](int i) const { return v[i]; }
    Vec3 operator+(const Vec3& o) const { return Vec3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }  // Vec.h:437-444
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }  // Vec.h:455-462
    Vec3 operator-() const { return Vec3(-v[0], -v[1], -v[2]); }
    void operator+=(const Vec3& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; }  // Vec.h:447-452
    void operator-=(const Vec3& o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; }  // Vec.h:465-470
    Vec3 operator*(R f) const { return Vec3(v[0] * f, v[1] * f, v[2] * f); }            // Vec.h:325-344 (scalar cast to Real first)
    void operator*=(R f) { v[0] *= f; v[1] *= f; v[2] *= f; }                          // Vec.h:347-363
    Vec3 operator/(R f) const { return Vec3(v[0] / f, v[1] / f, v[2] / f); }            // Vec.h:366-384
    void operator/=(R f) { v[0] /= f; v[1] /= f; v[2] /= f; }                          // Vec.h:387-403
    R norm2() const { R r = v[0] * v[0]; r += v[1] * v[1]; r += v[2] * v[2]; return r; }  // Vec.h:483-493
    // Vec.h:496-499.  sqrt(double) on a float then cast back == sqrtf (correctly rounded).
    R norm() const { return R(std::sqrt(norm2())); }
    bool normalize() {  // Vec.h:545-562: divide by the norm when it exceeds epsilon
        const R n = norm();
        if (n > std::numeric_limits<R>::epsilon()) { v[0] /= n; v[1] /= n; v[2] /= n; return true; }
        return false;
    }
    Vec3 normalized() const { Vec3 r(*this); r.normalize(); return r; }  // Vec.h:573-578
};
template <class R> inline R dot(const Vec3<R>& a, const Vec3<R>& b) {  // Vec.h:406-413
    R r = a[0] * b[0]; r += a[1] * b[1]; r += a[2] * b[2]; return r;
}
template <class R> inline Vec3<R> cross(const Vec3<R>& a, const Vec3<R>& b) {  // Vec.h:774-780
    return Vec3<R>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

// ---------------------------------------------------------------------------
// sofa::type::Mat<3,3,Real>   Sofa/framework/Type/src/sofa/type/Mat.h
// ---------------------------------------------------------------------------
template <class R> struct Mat3 {
    R m[3][3];
    Mat3() { for (auto& r : m) for (auto& x : r) x = 0; }
    R& operator()(int i, int j) { return m[i][j]; }
    const R& operator()(int i, int j) const { return m[i][j]; }
    Vec3<R> row(int i) const { return Vec3<R>(m[i][0], m[i][1], m[i][2]); }
    void setRow(int i, const Vec3<R>& v) { m[i][0] = v[0]; m[i][1] = v[1]; m[i][2] = v[2]; }
    void identity() { *this = Mat3(); m[0][0] = m[1][1] = m[2][2] = 1; }
    Mat3 transposed() const { Mat3 t; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t.m[i][j] = m[j][i]; return t; }
    Vec3<R> operator*(const Vec3<R>& v) const {  // Mat.h:577-587
        Vec3<R> r;
        for (int i = 0; i < 3; ++i) { r[i] = m[i][0] * v[0]; for (int j = 1; j < 3; ++j) r[i] += m[i][j] * v[j]; }
        return r;
    }
    Vec3<R> multTranspose(const Vec3<R>& v) const {  // Mat.h:601-611
        Vec3<R> r;
        for (int i = 0; i < 3; ++i) { r[i] = m[0][i] * v[0]; for (int j = 1; j < 3; ++j) r[i] += m[j][i] * v[j]; }
        return r;
    }
    Mat3 operator*(const Mat3& b) const {  // Mat.h:1443-1481 (3x3 specialisation)
        Mat3 r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][0] * b.m[0][j] + m[i][1] * b.m[1][j] + m[i][2] * b.m[2][j];
        return r;
    }
    Mat3 multTransposed(const Mat3& b) const {  // Mat.h:624-636  (this * b^T)
        Mat3 r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            r.m[i][j] = m[i][0] * b.m[j][0];
            for (int k = 1; k < 3; ++k) r.m[i][j] += m[i][k] * b.m[j][k];
        }
        return r;
    }
    Mat3 multTranspose(const Mat3& b) const {  // Mat.h:1501-1540 (this^T * b, 3x3 specialisation)
        Mat3 r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[0][i] * b.m[0][j] + m[1][i] * b.m[1][j] + m[2][i] * b.m[2][j];
        return r;
    }
    Mat3 multDiagonal(const Vec3<R>& d) const {  // Mat.h:591-598
        Mat3 r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][j] * d[j];
        return r;
    }
    Mat3 operator*(R f) const { Mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][j] * f; return r; }  // Mat.h:660-667
    Mat3 operator+(const Mat3& b) const { Mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][j] + b.m[i][j]; return r; }
    void operator-=(const Mat3& b) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] -= b.m[i][j]; }
    void operator*=(R f) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] *= f; }  // Mat.h:686-690
};
template <class R> inline R rabs(R r) { return (r >= 0) ? r : -r; }  // Sofa/framework/Helper/src/sofa/helper/rmath.h:100-104
template <class R> inline R determinant(const Mat3<R>& m) {             // Mat.h:988-997
    return m(0, 0) * m(1, 1) * m(2, 2) + m(1, 0) * m(2, 1) * m(0, 2) + m(2, 0) * m(0, 1) * m(1, 2)
         - m(0, 0) * m(2, 1) * m(1, 2) - m(1, 0) * m(0, 1) * m(2, 2) - m(2, 0) * m(1, 1) * m(0, 2);
}
template <class R> inline bool invertMatrix(Mat3<R>& dest, const Mat3<R>& from) {  // Mat.h:1171-1193
    const R det = determinant(from);
    if (rabs(det) <= std::numeric_limits<R>::epsilon()) return false;  // equalsZero, Mat.h:45-49
    dest(0, 0) = (from(1, 1) * from(2, 2) - from(2, 1) * from(1, 2)) / det;
    dest(1, 0) = (from(1, 2) * from(2, 0) - from(2, 2) * from(1, 0)) / det;
    dest(2, 0) = (from(1, 0) * from(2, 1) - from(2, 0) * from(1, 1)) / det;
    dest(0, 1) = (from(2, 1) * from(0, 2) - from(0, 1) * from(2, 2)) / det;
    dest(1, 1) = (from(2, 2) * from(0, 0) - from(0, 2) * from(2, 0)) / det;
    dest(2, 1) = (from(2, 0) * from(0, 1) - from(0, 0) * from(2, 1)) / det;
    dest(0, 2) = (from(0, 1) * from(1, 2) - from(1, 1) * from(0, 2)) / det;
    dest(1, 2) = (from(0, 2) * from(1, 0) - from(1, 2) * from(0, 0)) / det;
    dest(2, 2) = (from(0, 0) * from(1, 1) - from(1, 0) * from(0, 1)) / det;
    return true;
}
template <class R> inline R oneNorm(const Mat3<R>& A) {  // Mat.h:1055-1066
    R norm = 0;
    for (int i = 0; i < 3; ++i) { R s = rabs(A(0, i)) + rabs(A(1, i)) + rabs(A(2, i)); if (s > norm) norm = s; }
    return norm;
}
template <class R> inline R infNorm(const Mat3<R>& A) {  // Mat.h:1069-1080
    R norm = 0;
    for (int i = 0; i < 3; ++i) { R s = rabs(A(i, 0)) + rabs(A(i, 1)) + rabs(A(i, 2)); if (s > norm) norm = s; }
    return norm;
}

// ---------------------------------------------------------------------------
// sofa::helper::Decompose<Real>   Sofa/framework/Helper/src/sofa/helper/decompose.inl
// ---------------------------------------------------------------------------
template <class R> struct Decompose {
    static R zeroTolerance();  // decompose.h:377-387
    static R rsqrt(R a);       // rmath.h:124-137: sqrtf for float, sqrt for double

    // decompose.inl:672-723.  Scaled Newton iteration on M^T (Shoemake 1993 / Barbic, Vega).
    // `sqrt` / `fabs` there are the unqualified C names: in the reference's translation unit they
    // resolve to the double overloads, so for Real=float the gamma expression is evaluated in
    // double and rounded once to float (verified bit-for-bit against oracle/_ref).
    static R polarDecomposition(const Mat3<R>& M, Mat3<R>& Q) {
        Mat3<R> Mk = M.transposed(), Ek;
        R det, M_oneNorm = oneNorm(Mk), M_infNorm = infNorm(Mk), E_oneNorm;
        do {
            Mat3<R> MadjTk;
            MadjTk.setRow(0, cross(Mk.row(1), Mk.row(2)));
            MadjTk.setRow(1, cross(Mk.row(2), Mk.row(0)));
            MadjTk.setRow(2, cross(Mk.row(0), Mk.row(1)));
            det = Mk(0, 0) * MadjTk(0, 0) + Mk(0, 1) * MadjTk(0, 1) + Mk(0, 2) * MadjTk(0, 2);
            if (det == 0.0) break;
            const R MadjT_one = oneNorm(MadjTk), MadjT_inf = infNorm(MadjTk);
            const R gamma = gammaExpr((MadjT_one * MadjT_inf) / (M_oneNorm * M_infNorm), det);
            const R g1 = gamma * R(0.5);
            const R g2 = R(0.5) / (gamma * det);
            Ek = Mk;
            Mk = Mk * g1 + MadjTk * g2;
            Ek -= Mk;
            E_oneNorm = oneNorm(Ek);
            M_oneNorm = oneNorm(Mk);
            M_infNorm = infNorm(Mk);
        } while (E_oneNorm > M_oneNorm * zeroTolerance());
        Q = Mk.transposed();
        return det;
    }
    static R gammaExpr(R ratio, R det);

    // decompose.inl:1489-1557 (QL with implicit shifts on a 3x3 tridiagonal, from Numerical Recipes)
    static void QLAlgorithm(Vec3<R>& diag, Vec3<R>& subDiag, Mat3<R>& V) {
        const int iSize = 3, iMaxIter = 32;
        for (int i0 = 0; i0 < iSize; ++i0) {
            int i1;
            for (i1 = 0; i1 < iMaxIter; ++i1) {
                int i2;
                for (i2 = i0; i2 <= iSize - 2; ++i2) {
                    R fTmp = rabs(diag[i2]) + rabs(diag[i2 + 1]);
                    if (rabs(subDiag[i2]) + fTmp == fTmp) break;
                }
                if (i2 == i0) break;
                R fG = (diag[i0 + 1] - diag[i0]) / (R(2.0) * subDiag[i0]);
                R fR = rsqrt(fG * fG + R(1.0));
                if (fG < R(0.0)) fG = diag[i2] - diag[i0] + subDiag[i0] / (fG - fR);
                else             fG = diag[i2] - diag[i0] + subDiag[i0] / (fG + fR);
                R fSin = 1.0, fCos = 1.0, fP = 0.0;
                for (int i3 = i2 - 1; i3 >= i0; --i3) {
                    R fF = fSin * subDiag[i3];
                    R fB = fCos * subDiag[i3];
                    if (rabs(fF) >= rabs(fG)) {
                        fCos = fG / fF;
                        fR = rsqrt(fCos * fCos + R(1.0));
                        subDiag[i3 + 1] = fF * fR;
                        fSin = R(1.0) / fR;
                        fCos *= fSin;
                    } else {
                        fSin = fF / fG;
                        fR = rsqrt(fSin * fSin + R(1.0));
                        subDiag[i3 + 1] = fG * fR;
                        fCos = R(1.0) / fR;
                        fSin *= fCos;
                    }
                    fG = diag[i3 + 1] - fP;
                    fR = (diag[i3] - fG) * fSin + R(2.0) * fB * fCos;
                    fP = fSin * fR;
                    diag[i3 + 1] = fG + fP;
                    fG = fCos * fR - fB;
                    for (int i4 = 0; i4 < iSize; ++i4) {
                        fF = V(i4, i3 + 1);
                        V(i4, i3 + 1) = fSin * V(i4, i3) + fCos * fF;
                        V(i4, i3) = fCos * V(i4, i3) - fSin * fF;
                    }
                }
                diag[i0] -= fP;
                subDiag[i0] = fG;
                subDiag[i2] = R(0.0);
            }
            if (i1 == iMaxIter) return;
        }
    }

    // decompose.inl:1561-1608 (Householder tridiagonalisation then QL)
    static void eigenDecomposition_iterative(const Mat3<R>& M, Mat3<R>& V, Vec3<R>& diag) {
        Vec3<R> subDiag;
        const R fM00 = M(0, 0);
        R fM01 = M(0, 1), fM02 = M(0, 2);
        const R fM11 = M(1, 1), fM12 = M(1, 2), fM22 = M(2, 2);
        diag[0] = fM00;
        subDiag[2] = R(0.0);
        if (fM02 != R(0.0)) {
            R fLength = rsqrt(fM01 * fM01 + fM02 * fM02);
            R fInvLength = R(1.0) / fLength;
            fM01 *= fInvLength;
            fM02 *= fInvLength;
            R fQ = R(2.0) * fM01 * fM12 + fM02 * (fM22 - fM11);
            diag[1] = fM11 + fM02 * fQ;
            diag[2] = fM22 - fM02 * fQ;
            subDiag[0] = fLength;
            subDiag[1] = fM12 - fM01 * fQ;
            V(0, 0) = R(1.0); V(0, 1) = R(0.0); V(0, 2) = R(0.0);
            V(1, 0) = R(0.0); V(1, 1) = fM01;   V(1, 2) = fM02;
            V(2, 0) = R(0.0); V(2, 1) = fM02;   V(2, 2) = -fM01;
        } else {
            diag[1] = fM11;
            diag[2] = fM22;
            subDiag[0] = fM01;
            subDiag[1] = fM12;
            V.identity();
        }
        QLAlgorithm(diag, subDiag, V);
    }

    // decompose.inl:1662-1829
    static bool SVD_stable(const Mat3<R>& F, Mat3<R>& U, Vec3<R>& S, Mat3<R>& V) {
        Mat3<R> FtF = F.multTranspose(F);
        eigenDecomposition_iterative(FtF, V, S);
        if (determinant(V) < R(0)) for (int i = 0; i < 3; ++i) V(i, 0) = -V(i, 0);
        int degenerated = 0;
        Vec3<R> S_1;
        for (int i = 0; i < 3; ++i) {
            if (S[i] < zeroTolerance()) { degenerated++; S[i] = R(0); S_1[i] = R(1); }
            else { S[i] = rsqrt(S[i]); S_1[i] = R(1.) / S[i]; }
        }
        unsigned Sorder[3];
        if (S[0] < S[1]) {
            if (S[0] < S[2]) {
                Sorder[0] = 0;
                if (S[1] < S[2]) { Sorder[1] = 1; Sorder[2] = 2; } else { Sorder[1] = 2; Sorder[2] = 1; }
            } else { Sorder[0] = 2; Sorder[1] = 0; Sorder[2] = 1; }
        } else {
            if (S[1] < S[2]) {
                Sorder[0] = 1;
                if (S[0] < S[2]) { Sorder[1] = 0; Sorder[2] = 2; } else { Sorder[1] = 2; Sorder[2] = 0; }
            } else { Sorder[0] = 2; Sorder[1] = 1; Sorder[2] = 0; }
        }
        switch (degenerated) {
        case 0:
            U = F * V.multDiagonal(S_1);
            break;
        case 1: {
            U = F * V.multDiagonal(S_1);
            Vec3<R> c = cross(Vec3<R>(U(0, Sorder[1]), U(1, Sorder[1]), U(2, Sorder[1])),
                              Vec3<R>(U(0, Sorder[2]), U(1, Sorder[2]), U(2, Sorder[2])));
            U(0, Sorder[0]) = c[0]; U(1, Sorder[0]) = c[1]; U(2, Sorder[0]) = c[2];
            break;
        }
        case 2: {
            U = F * V.multDiagonal(S_1);
            // (sic) the reference mixes Sorder[2]/Sorder[0] here, decompose.inl:1767
            Vec3<R> edge0, edge1, edge2(U(0, Sorder[2]), U(1, Sorder[0]), U(2, Sorder[0]));
            R abs0 = rabs(edge2[0]), abs1 = rabs(edge2[1]), abs2 = rabs(edge2[2]);
            if (abs0 > abs1) {
                if (abs0 > abs2) { edge0 = Vec3<R>(0, 1, 0); } else { edge0 = Vec3<R>(1, 0, 0); }
            } else {
                if (abs1 > abs2) { edge0 = Vec3<R>(0, 0, 1); } else { edge0 = Vec3<R>(1, 0, 0); }
            }
            edge1 = cross(edge2, edge0);
            edge1.normalize();
            edge0 = cross(edge1, edge2);
            U(0, Sorder[0]) = edge0[0]; U(1, Sorder[0]) = edge0[1]; U(2, Sorder[0]) = edge0[2];
            U(0, Sorder[1]) = edge1[0]; U(1, Sorder[1]) = edge1[1]; U(2, Sorder[1]) = edge1[2];
            break;
        }
        case 3:
            U.identity();
            break;
        }
        const bool inverted = (determinant(U) < R(0));
        if (inverted) {
            U(0, Sorder[0]) *= R(-1); U(1, Sorder[0]) *= R(-1); U(2, Sorder[0]) *= R(-1);
            S[Sorder[0]] *= R(-1);
        }
        return degenerated || inverted;
    }

    // decompose.inl:754-764   Q = U * V^T
    static bool polarDecomposition_stable(const Mat3<R>& M, Mat3<R>& Q) {
        Mat3<R> U, V; Vec3<R> Sdiag;
        const bool degenerated = SVD_stable(M, U, Sdiag, V);
        Q = U.multTransposed(V);
        return degenerated;
    }
};
template <> inline float Decompose<float>::zeroTolerance() { return 1e-6f; }
template <> inline double Decompose<double>::zeroTolerance() { return 1e-8; }
template <> inline float Decompose<float>::rsqrt(float a) { return sqrtf(a); }
template <> inline double Decompose<double>::rsqrt(double a) { return std::sqrt(a); }
template <> inline double Decompose<double>::gammaExpr(double ratio, double det) {
    return std::sqrt(std::sqrt(ratio) / std::fabs(det));
}
#ifndef ORC_POLAR_FLOAT_VARIANT
#define ORC_POLAR_FLOAT_VARIANT 0
#endif
template <> inline float Decompose<float>::gammaExpr(float ratio, float det) {
#if ORC_POLAR_FLOAT_VARIANT == 0
    // ::sqrt(double)/::fabs(double): float operands widen, one final rounding to float.
    return float(std::sqrt(std::sqrt(double(ratio)) / std::fabs(double(det))));
#else
    return sqrtf(sqrtf(ratio) / fabsf(det));
#endif
}

// ---------------------------------------------------------------------------
// Mesh generation (init-time inputs of the path)
// ---------------------------------------------------------------------------
struct GridMesh {
    int nx, ny, nz;
    std::vector<double> pos;        // 3*N, SReal positions as RegularGridTopology produces them
    std::vector<uint32_t> hexas;    // 8*H
};
// Sofa/Component/Topology/Container/Grid/src/sofa/component/topology/container/grid/RegularGridTopology.cpp:124-168
// and GridTopology.cpp:352-398 (point index nx*(ny*k+j)+i; hexa corners).
inline GridMesh regularGrid(int nx, int ny, int nz, const double mn[3], const double mx[3]) {
    GridMesh g; g.nx = nx; g.ny = ny; g.nz = nz;
    double p0[3] = {mn[0], mn[1], mn[2]}, d[3];
    const int n[3] = {nx - 1, ny - 1, nz - 1};
    for (int c = 0; c < 3; ++c) {
        if (n[c] > 0) d[c] = (mx[c] - mn[c]) / n[c];
        else { d[c] = mx[c] - mn[c]; if (c < 2) p0[c] = (mx[c] + mn[c]) / 2; }
    }
    g.pos.resize(size_t(3) * nx * ny * nz);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        size_t id = size_t(nx) * (size_t(ny) * k + j) + i;
        // p0 + dx*i + dy*j + dz*k on Vec3 (SReal): each axis only gets its own term (+0 from the others)
        g.pos[3 * id + 0] = p0[0] + d[0] * i + 0.0 * j + 0.0 * k;
        g.pos[3 * id + 1] = p0[1] + 0.0 * i + d[1] * j + 0.0 * k;
        g.pos[3 * id + 2] = p0[2] + 0.0 * i + 0.0 * j + d[2] * k;
    }
    auto P = [&](int x, int y, int z) { return uint32_t(nx * (ny * z + y) + x); };
    for (int z = 0; z < nz - 1; ++z) for (int y = 0; y < ny - 1; ++y) for (int x = 0; x < nx - 1; ++x) {
        const uint32_t h[8] = {P(x, y, z), P(x + 1, y, z), P(x + 1, y + 1, z), P(x, y + 1, z),
                               P(x, y, z + 1), P(x + 1, y, z + 1), P(x + 1, y + 1, z + 1), P(x, y + 1, z + 1)};
        g.hexas.insert(g.hexas.end(), h, h + 8);
    }
    return g;
}
// Sofa/framework/Geometry/src/sofa/geometry/Hexahedron.h:67-88
static const int kXEdges[4][2] = {{0, 1}, {4, 5}, {3, 2}, {7, 6}};
static const int kYEdges[4][2] = {{4, 7}, {5, 6}, {1, 2}, {0, 3}};
static const int kZEdges[4][2] = {{4, 0}, {5, 1}, {6, 2}, {7, 3}};
// mode 0: Hexa2TetraTopologicalMapping swapping=false; mode 1: swapping=true
//   (Sofa/Component/Topology/Mapping/src/sofa/component/topology/mapping/Hexa2TetraTopologicalMapping.cpp:104-196)
// mode 2: tessellation built into TetrahedronFEMForceField::init when the topology has only hexahedra
//   (TetrahedronFEMForceField.inl:1317-1371: always swaps, always the first pattern)
inline std::vector<uint32_t> hexasToTetras(const GridMesh& g, int mode) {
    static const int nonSwapped[6][4] = {{0, 5, 1, 6}, {0, 1, 3, 6}, {1, 3, 6, 2}, {6, 3, 0, 7}, {6, 7, 0, 5}, {7, 5, 4, 0}};
    static const int swappedP[6][4] = {{0, 5, 6, 1}, {0, 1, 6, 3}, {1, 3, 2, 6}, {6, 3, 7, 0}, {6, 7, 5, 0}, {7, 5, 0, 4}};
    const size_t H = g.hexas.size() / 8;
    const int nx = g.nx - 1, ny = g.ny - 1;
    std::vector<uint32_t> tets; tets.reserve(H * 24);
    for (size_t i = 0; i < H; ++i) {
        uint32_t c[8]; for (int k = 0; k < 8; ++k) c[k] = g.hexas[8 * i + k];
        bool swapped = false;
        if (mode != 0) {
            if (!((i % nx) & 1)) { for (auto& e : kXEdges) std::swap(c[e[0]], c[e[1]]); swapped = !swapped; }
            if (((i / nx) % ny) & 1) { for (auto& e : kYEdges) std::swap(c[e[0]], c[e[1]]); swapped = !swapped; }
            if ((i / (size_t(nx) * ny)) & 1) { for (auto& e : kZEdges) std::swap(c[e[0]], c[e[1]]); swapped = !swapped; }
        }
        const int (*pat)[4] = (mode == 1 && swapped) ? swappedP : nonSwapped;
        for (int t = 0; t < 6; ++t) for (int k = 0; k < 4; ++k) tets.push_back(c[pat[t][k]]);
    }
    return tets;
}

// ---------------------------------------------------------------------------
// State vectors of a MechanicalObject<Vec3Types>: AoS Vec3 arrays
// ---------------------------------------------------------------------------
template <class R> using VecDeriv = std::vector<Vec3<R>>;

// MechanicalObject::vOp cases used on the path
// Sofa/Component/StateContainer/src/sofa/component/statecontainer/MechanicalObject.inl:1930-2203
template <class R> struct VOps {
    static void clear(VecDeriv<R>& v) { for (auto& x : v) x = Vec3<R>(); }                                    // r = 0          :2088-2098
    static void teq(VecDeriv<R>& v, SReal k) { const R f = R(k); for (auto& x : v) x *= f; }                    // r *= k         vOp_vf :1930-1943
    static void eq_bf(VecDeriv<R>& v, const VecDeriv<R>& b, SReal k) { const R f = R(k); v.resize(b.size()); for (size_t i = 0; i < v.size(); ++i) v[i] = b[i] * f; }  // r = b*k vOp_vbf :1945-1962
    static void eq(VecDeriv<R>& v, const VecDeriv<R>& a) { v = a; }                                             // r = a          vOp_va :1964-1984
    static void peq(VecDeriv<R>& v, const VecDeriv<R>& b) { for (size_t i = 0; i < v.size(); ++i) v[i] += b[i]; }  // r += b       vOp_vb :1986-2003
    static void avf(VecDeriv<R>& v, const VecDeriv<R>& a, SReal k) { const R f = R(k); for (size_t i = 0; i < v.size(); ++i) { v[i] *= f; v[i] += a[i]; } }  // r = a + r*k  vOp_avf :2005-2023
    static void peq_bf(VecDeriv<R>& v, const VecDeriv<R>& b, SReal k) { const R f = R(k); for (size_t i = 0; i < v.size(); ++i) v[i] += b[i] * f; }         // r += b*k     vOp_v_inc_bf :2025-2042
    static void eq_ab(VecDeriv<R>& v, const VecDeriv<R>& a, const VecDeriv<R>& b) { v.resize(b.size()); for (size_t i = 0; i < v.size(); ++i) v[i] = a[i] + b[i]; }  // vOp_vab
    static void eq_abf(VecDeriv<R>& v, const VecDeriv<R>& a, const VecDeriv<R>& b, SReal k) { const R f = R(k); v.resize(b.size()); for (size_t i = 0; i < v.size(); ++i) v[i] = a[i] + b[i] * f; }  // vOp_vabf
    // generic dispatcher with the null-id semantics of MechanicalObject::vOp (:2075-2203); null == nullptr
    static void vOp(VecDeriv<R>* r, const VecDeriv<R>* a, const VecDeriv<R>* b, SReal k) {
        if (!a) {
            if (!b) clear(*r);
            else if (r == b) teq(*r, k);
            else eq_bf(*r, *b, k);
        } else if (!b) eq(*r, *a);
        else if (r == a) { if (k == 1.0) peq(*r, *b); else peq_bf(*r, *b, k); }
        else if (r == b) { if (k == 1.0) peq(*r, *a); else avf(*r, *a, k); }
        else { if (k == 1.0) eq_ab(*r, *a, *b); else eq_abf(*r, *a, *b, k); }
    }
    // MechanicalObject::vDot :2333-2356 -- serial accumulation in Real, returned as SReal
    static SReal dot(const VecDeriv<R>& a, const VecDeriv<R>& b) {
        R r = 0.0;
        for (size_t i = 0; i < a.size(); ++i) r += orc::dot(a[i], b[i]);
        return r;
    }
    // MechanicalObject::vMultiOp integration fast path :2208-2241 (f_v_v == f_x_x == 1, f_v_a == 1)
    static void integrate(VecDeriv<R>& v, VecDeriv<R>& x, const VecDeriv<R>& a, SReal h) {
        const R f_x_v = R(h);
        for (size_t i = 0; i < x.size(); ++i) { v[i] += a[i]; x[i] += v[i] * f_x_v; }
    }
};

// ---------------------------------------------------------------------------
// TetrahedronFEMForceField<Vec3Types>
// Sofa/Component/SolidMechanics/FEM/Elastic/src/sofa/component/solidmechanics/fem/elastic/TetrahedronFEMForceField.inl
// ---------------------------------------------------------------------------
enum TetMethod { SMALL = 0, LARGE = 1, POLAR = 2, SVD = 3 };

template <class R> struct TetFEM {
    typedef Vec3<R> Coord;
    int method = LARGE;
    std::vector<uint32_t> tets;           // 4*T  (_indexedElements)
    std::vector<Coord> initialPoints;     // d_initialPoints (rest positions)
    std::vector<R> young, poisson;        // d_youngModulus / d_poissonRatio (VecReal)
    std::vector<R> localStiffnessFactor;  // d_localStiffnessFactor
    // per element state, as the reference class keeps it
    std::vector<R> K;                     // materialsStiffnesses: 3 distinct values per element {K00, K01, K33} (see :256-291)
    std::vector<R> J;                     // strainDisplacements: the 12 cofactors {jx,jy,jz} x 4 nodes (all other entries are copies/zeros)
    std::vector<Mat3<R>> rotations;       // rotations[e]         (= R^T of the current frame)
    std::vector<Mat3<R>> initialRotations;// _initialRotations[e] (= R0^T)
    std::vector<Coord> X0;                // _rotatedInitialElements[e][0..3]
    std::vector<Mat3<R>> initialTransformation;  // _initialTransformation (svd: A0^-1)
    std::vector<uint32_t> rotationIdx;    // _rotationIdx
    bool tetrahedralCorotational = false; // the sibling class TetrahedralCorotationalFEMForceField (same statements, TetrahedralCorotationalFEMForceField.inl:356-1175): only its
                                          // accumulateForceLarge differs, by rewriting all three copies of a cofactor under updateStiffnessMatrix (:920-937)
    bool updateStiffnessMatrix = false;   // d_updateStiffnessMatrix: polar / svd recompute the whole strain-displacement matrix (:1063-1067, :1174-1177); `large` rewrites nine
                                          // single entries, all in the normal-strain columns J(.,0..2) (:908-922), so the copies of the same cofactors in the shear columns
                                          // J(.,3..5) keep their initial values: J holds the normal-column copies, Jsh the shear-column copies (equal until that happens)
    std::vector<R> Jsh;
    std::vector<R> plasticStrains;        // _plasticStrains: 6 Voigt components per element
    R plastic[3] = {R(0), R(0.0001f), R(0.9f)};  // d_plasticMaxThreshold, d_plasticYieldThreshold, d_plasticCreep (defaults :51-53)
    R* plasticPtr(size_t e) { return plastic[0] > 0 ? &plasticStrains[6 * e] : nullptr; }
    void reset() { std::fill(plasticStrains.begin(), plasticStrains.end(), R(0)); }   // reset() :1380-1388
    double restVolume = 0;

    size_t nbTets() const { return tets.size() / 4; }
    R youngIn(size_t e) const { return young.size() > e ? young[e] : young[0]; }      // BaseLinearElasticityFEMForceField.inl:112-137
    R poissonIn(size_t e) const { return poisson.size() > e ? poisson[e] : poisson[0]; }

    static R pd(R m00, R m01, R m02, R m10, R m11, R m12) {  // peudo_determinant_for_coef :204-208
        return m01 * m12 - m11 * m02 - m00 * m12 + m10 * m02 + m00 * m11 - m10 * m01;
    }
    // computeStrainDisplacement :134-202.  Output: j[3*n+{0,1,2}] = the three distinct values of node n's rows
    // (J(3n,0)=J(3n+1,3)=J(3n+2,5)=jx, J(3n,3)=J(3n+1,1)=J(3n+2,4)=jy, J(3n,5)=J(3n+1,4)=J(3n+2,2)=jz).
    static void computeStrainDisplacement(R* j, Coord a, Coord b, Coord c, Coord d) {
        j[0] = -pd(b[1], c[1], d[1], b[2], c[2], d[2]);
        j[1] =  pd(b[0], c[0], d[0], b[2], c[2], d[2]);
        j[2] = -pd(b[0], c[0], d[0], b[1], c[1], d[1]);
        j[3] =  pd(c[1], d[1], a[1], c[2], d[2], a[2]);
        j[4] = -pd(c[0], d[0], a[0], c[2], d[2], a[2]);
        j[5] =  pd(c[0], d[0], a[0], c[1], d[1], a[1]);
        j[6] = -pd(d[1], a[1], b[1], d[2], a[2], b[2]);
        j[7] =  pd(d[0], a[0], b[0], d[2], a[2], b[2]);
        j[8] = -pd(d[0], a[0], b[0], d[1], a[1], b[1]);
        j[9]  =  pd(a[1], b[1], c[1], a[2], b[2], c[2]);
        j[10] = -pd(a[0], b[0], c[0], a[2], b[2], c[2]);
        j[11] =  pd(a[0], b[0], c[0], a[1], b[1], c[1]);
    }
    // full 12x6 matrix view, for the golden-vector tests
    void strainDisplacementMatrix(size_t e, R out[72]) const {
        std::memset(out, 0, sizeof(R) * 72);
        for (int n = 0; n < 4; ++n) {
            const R jx = J[12 * e + 3 * n], jy = J[12 * e + 3 * n + 1], jz = J[12 * e + 3 * n + 2];
            R* r0 = out + 6 * (3 * n); R* r1 = r0 + 6; R* r2 = r1 + 6;
            r0[0] = jx; r0[3] = jy; r0[5] = jz;
            r1[1] = jy; r1[3] = jx; r1[4] = jz;
            r2[2] = jz; r2[4] = jy; r2[5] = jx;
        }
    }
    void materialStiffnessMatrix(size_t e, R out[36]) const {
        std::memset(out, 0, sizeof(R) * 36);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out[6 * i + j] = (i == j) ? K[3 * e] : K[3 * e + 1];
        for (int i = 3; i < 6; ++i) out[6 * i + i] = K[3 * e + 2];
    }

    // computeMaterialStiffness :255-291 ; volume: Sofa/framework/Geometry/src/sofa/geometry/Tetrahedron.h:55-82
    void computeMaterialStiffness(size_t i, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        const R youngModulusElement = youngIn(i);
        const R youngModulus = (localStiffnessFactor.empty() ? 1.0f : localStiffnessFactor[i * localStiffnessFactor.size() / nbTets()]) * youngModulusElement;
        const R poissonRatio = poissonIn(i);
        R k00 = 1;
        R k01 = poissonRatio / (1 - poissonRatio);
        R k33 = (1 - 2 * poissonRatio) / (2 * (1 - poissonRatio));
        const R s = (youngModulus * (1 - poissonRatio)) / ((1 + poissonRatio) * (1 - 2 * poissonRatio));
        k00 *= s; k01 *= s; k33 *= s;
        elemLambda[i] = k01; elemMu[i] = k33;   // :278-282, before the division by 36 V
        const Coord A = initialPoints[b] - initialPoints[a], B = initialPoints[c] - initialPoints[a], C = initialPoints[d] - initialPoints[a];
        const R tetrahedronVolume = std::abs(dot(cross(A, B), C) / R(6));
        restVolume += tetrahedronVolume;
        const R div = tetrahedronVolume * 36;
        K[3 * i] = k00 / div; K[3 * i + 1] = k01 / div; K[3 * i + 2] = k33 / div;
    }
    // computeRotationLarge :754-778
    static void computeRotationLarge(Mat3<R>& r, const std::vector<Coord>& p, uint32_t a, uint32_t b, uint32_t c) {
        const Coord edgex = (p[b] - p[a]).normalized();
        Coord edgey = p[c] - p[a];
        const Coord edgez = cross(edgex, edgey).normalized();
        edgey = cross(edgez, edgex);
        r.setRow(0, edgex); r.setRow(1, edgey); r.setRow(2, edgez);
    }
    void initSmall(size_t i, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {  // :526-532
        computeStrainDisplacement(&J[12 * i], initialPoints[a], initialPoints[b], initialPoints[c], initialPoints[d]);
    }
    void initLarge(size_t i, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {  // :834-868
        Mat3<R> R_0_1;
        computeRotationLarge(R_0_1, initialPoints, a, b, c);
        initialRotations[i] = R_0_1.transposed();
        rotations[i] = initialRotations[i];
        rotationIdx[a] = rotationIdx[b] = rotationIdx[c] = rotationIdx[d] = uint32_t(i);
        Coord* x0 = &X0[4 * i];
        x0[0] = R_0_1 * initialPoints[a]; x0[1] = R_0_1 * initialPoints[b];
        x0[2] = R_0_1 * initialPoints[c]; x0[3] = R_0_1 * initialPoints[d];
        x0[1] -= x0[0]; x0[2] -= x0[0]; x0[3] -= x0[0];
        x0[0] = Coord(0, 0, 0);
        computeStrainDisplacement(&J[12 * i], x0[0], x0[1], x0[2], x0[3]);
    }
    void initPolarLike(size_t i, uint32_t a, uint32_t b, uint32_t c, uint32_t d, bool svd) {  // initPolar :992-1023, initSVD :1086-1117
        Mat3<R> A;
        A.setRow(0, initialPoints[b] - initialPoints[a]);
        A.setRow(1, initialPoints[c] - initialPoints[a]);
        A.setRow(2, initialPoints[d] - initialPoints[a]);
        if (svd) invertMatrix(initialTransformation[i], A);
        Mat3<R> R_0_1;
        Decompose<R>::polarDecomposition(A, R_0_1);
        initialRotations[i] = R_0_1.transposed();
        rotations[i] = initialRotations[i];
        rotationIdx[a] = rotationIdx[b] = rotationIdx[c] = rotationIdx[d] = uint32_t(i);
        Coord* x0 = &X0[4 * i];
        x0[0] = R_0_1 * initialPoints[a]; x0[1] = R_0_1 * initialPoints[b];
        x0[2] = R_0_1 * initialPoints[c]; x0[3] = R_0_1 * initialPoints[d];
        computeStrainDisplacement(&J[12 * i], x0[0], x0[1], x0[2], x0[3]);
    }
    // reinit :1390-1505
    void reinit(const std::vector<Coord>& restPosition) {
        initialPoints = restPosition;
        const size_t T = nbTets();
        K.assign(3 * T, 0); J.assign(12 * T, 0);
        plasticStrains.assign(6 * T, R(0));   // :1415
        elemLambda.assign(T, R(0)); elemMu.assign(T, R(0));   // :1424-1425
        restVolume = 0;
        if (method != SMALL) {
            rotations.assign(T, Mat3<R>()); initialRotations.assign(T, Mat3<R>());
            rotationIdx.assign(restPosition.size(), 0); X0.assign(4 * T, Coord());
            if (method == SVD) initialTransformation.assign(T, Mat3<R>());
        }
        for (size_t i = 0; i < T; ++i) {
            const uint32_t a = tets[4 * i], b = tets[4 * i + 1], c = tets[4 * i + 2], d = tets[4 * i + 3];
            computeMaterialStiffness(i, a, b, c, d);
            switch (method) {
            case SMALL: initSmall(i, a, b, c, d); break;
            case LARGE: initLarge(i, a, b, c, d); break;
            case POLAR: initPolarLike(i, a, b, c, d, false); break;
            case SVD:   initPolarLike(i, a, b, c, d, true); break;
            }
        }
        Jsh = J;
    }

    // computeForce :293-415 (plasticity off) and :417-521 (with `fact`); useFact selects `KJtD *= fact`.
    // plasticStrain / plastic = {max, yield, creep}: the plasticity branch :357-371 (only when d_plasticMaxThreshold > 0)
    static void computeForce(R F[12], const R D[12], const R* k, const R* j, const R* js, bool useFact, SReal fact, R* plasticStrain = nullptr, const R* plastic = nullptr) {
        // J(3n,0)=j[3n] J(3n+1,1)=j[3n+1] J(3n+2,2)=j[3n+2]; J(3n,3)=js[3n+1] J(3n+1,3)=js[3n];
        // J(3n+1,4)=js[3n+2] J(3n+2,4)=js[3n+1]; J(3n,5)=js[3n+2] J(3n+2,5)=js[3n]
        R JtD[6];
        JtD[0] = j[0] * D[0] + j[3] * D[3] + j[6] * D[6] + j[9] * D[9];
        JtD[1] = j[1] * D[1] + j[4] * D[4] + j[7] * D[7] + j[10] * D[10];
        JtD[2] = j[2] * D[2] + j[5] * D[5] + j[8] * D[8] + j[11] * D[11];
        JtD[3] = js[1] * D[0] + js[0] * D[1] + js[4] * D[3] + js[3] * D[4] + js[7] * D[6] + js[6] * D[7] + js[10] * D[9] + js[9] * D[10];
        JtD[4] = js[2] * D[1] + js[1] * D[2] + js[5] * D[4] + js[4] * D[5] + js[8] * D[7] + js[7] * D[8] + js[11] * D[10] + js[10] * D[11];
        JtD[5] = js[2] * D[0] + js[0] * D[2] + js[5] * D[3] + js[3] * D[5] + js[8] * D[6] + js[6] * D[8] + js[11] * D[9] + js[9] * D[11];
        if (plasticStrain && plastic[0] > 0) {
            R elasticStrain[6];
            for (int i = 0; i < 6; ++i) elasticStrain[i] = JtD[i] - plasticStrain[i];           // VoigtTensor elasticStrain = JtD; elasticStrain -= plasticStrain
            R n2 = elasticStrain[0] * elasticStrain[0];                                         // Vec::norm2, Vec.h:483-493
            for (int i = 1; i < 6; ++i) n2 += elasticStrain[i] * elasticStrain[i];
            if (n2 > plastic[1] * plastic[1]) for (int i = 0; i < 6; ++i) plasticStrain[i] += elasticStrain[i] * plastic[2];   // += creep * elasticStrain
            R plasticStrainNorm2 = plasticStrain[0] * plasticStrain[0];
            for (int i = 1; i < 6; ++i) plasticStrainNorm2 += plasticStrain[i] * plasticStrain[i];
            if (plasticStrainNorm2 > plastic[0] * plastic[0]) {
                const R sc = plastic[0] / R(std::sqrt(plasticStrainNorm2));                      // helper::rsqrt = square root (rmath.h:125-138)
                for (int i = 0; i < 6; ++i) plasticStrain[i] *= sc;
            }
            for (int i = 0; i < 6; ++i) JtD[i] -= plasticStrain[i];
        }
        R KJtD[6];
        KJtD[0] = k[0] * JtD[0] + k[1] * JtD[1] + k[1] * JtD[2];
        KJtD[1] = k[1] * JtD[0] + k[0] * JtD[1] + k[1] * JtD[2];
        KJtD[2] = k[1] * JtD[0] + k[1] * JtD[1] + k[0] * JtD[2];
        KJtD[3] = k[2] * JtD[3];
        KJtD[4] = k[2] * JtD[4];
        KJtD[5] = k[2] * JtD[5];
        if (useFact) { const R f = R(fact); for (int i = 0; i < 6; ++i) KJtD[i] *= f; }
        for (int n = 0; n < 4; ++n) {
            const R jx = j[3 * n], jy = j[3 * n + 1], jz = j[3 * n + 2], sx = js[3 * n], sy = js[3 * n + 1], sz = js[3 * n + 2];
            F[3 * n + 0] = jx * KJtD[0] + sy * KJtD[3] + sz * KJtD[5];
            F[3 * n + 1] = jy * KJtD[1] + sx * KJtD[3] + sz * KJtD[4];
            F[3 * n + 2] = jz * KJtD[2] + sy * KJtD[4] + sx * KJtD[5];
        }
    }

    void accumulateForceSmall(VecDeriv<R>& f, const std::vector<Coord>& p, size_t e) {  // :534-620
        const uint32_t a = tets[4 * e], b = tets[4 * e + 1], c = tets[4 * e + 2], d = tets[4 * e + 3];
        const std::vector<Coord>& ip = initialPoints;
        R D[12];
        D[0] = 0; D[1] = 0; D[2] = 0;
        const uint32_t idx[3] = {b, c, d};
        for (int n = 0; n < 3; ++n) for (int k = 0; k < 3; ++k)
            D[3 * (n + 1) + k] = ip[idx[n]][k] - ip[a][k] - p[idx[n]][k] + p[a][k];
        R F[12];
        computeForce(F, D, &K[3 * e], &J[12 * e], &Jsh[12 * e], false, 0, plasticPtr(e), plastic);
        f[a] += Coord(F[0], F[1], F[2]); f[b] += Coord(F[3], F[4], F[5]);
        f[c] += Coord(F[6], F[7], F[8]); f[d] += Coord(F[9], F[10], F[11]);
    }
    void accumulateForceLarge(VecDeriv<R>& f, const std::vector<Coord>& p, size_t e) {  // :870-986
        const uint32_t* index = &tets[4 * e];
        Mat3<R> R_0_2;
        computeRotationLarge(R_0_2, p, index[0], index[1], index[2]);
        rotations[e] = R_0_2.transposed();
        Coord deforme[4];
        for (int i = 0; i < 4; ++i) deforme[i] = R_0_2 * p[index[i]];
        deforme[1][0] -= deforme[0][0];
        deforme[2][0] -= deforme[0][0];
        deforme[2][1] -= deforme[0][1];
        deforme[3] -= deforme[0];
        const Coord* x0 = &X0[4 * e];
        R D[12];
        D[0] = 0; D[1] = 0; D[2] = 0;
        D[3] = x0[1][0] - deforme[1][0]; D[4] = 0; D[5] = 0;
        D[6] = x0[2][0] - deforme[2][0]; D[7] = x0[2][1] - deforme[2][1]; D[8] = 0;
        D[9] = x0[3][0] - deforme[3][0]; D[10] = x0[3][1] - deforme[3][1]; D[11] = x0[3][2] - deforme[3][2];
        if (updateStiffnessMatrix) {   // TetrahedronFEMForceField.inl:908-922: J(0,0) J(1,1) J(2,2) J(3,0) J(4,1) J(5,2) J(7,1) J(8,2) J(11,2) (jx_n = j[3n], jy_n = j[3n+1], jz_n = j[3n+2])
            R* j = &J[12 * e];
            j[0] = ( - deforme[2][1]*deforme[3][2] );
            j[1] = ( deforme[2][0]*deforme[3][2] - deforme[1][0]*deforme[3][2] );
            j[2] = ( deforme[2][1]*deforme[3][0] - deforme[2][0]*deforme[3][1] + deforme[1][0]*deforme[3][1] - deforme[1][0]*deforme[2][1] );
            j[3] = ( deforme[2][1]*deforme[3][2] );
            j[4] = ( - deforme[2][0]*deforme[3][2] );
            j[5] = ( - deforme[2][1]*deforme[3][0] + deforme[2][0]*deforme[3][1] );
            j[7] = ( deforme[1][0]*deforme[3][2] );
            j[8] = ( - deforme[1][0]*deforme[3][1] );
            j[11] = ( deforme[1][0]*deforme[2][1] );
            if (tetrahedralCorotational)   // TetrahedralCorotationalFEMForceField.inl:920-937 assigns all three copies of each of the nine cofactors
                for (int q : {0, 1, 2, 3, 4, 5, 7, 8, 11}) Jsh[12 * e + q] = j[q];
        }
        R F[12];
        computeForce(F, D, &K[3 * e], &J[12 * e], &Jsh[12 * e], false, 0, plasticPtr(e), plastic);
        for (int i = 0; i < 12; i += 3) f[index[i / 3]] += rotations[e] * Coord(F[i], F[i + 1], F[i + 2]);
    }
    void accumulateForcePolarLike(VecDeriv<R>& f, const std::vector<Coord>& p, size_t e, bool svd) {  // :1025-1079, :1122-1185
        const uint32_t* index = &tets[4 * e];
        Mat3<R> A;
        A.setRow(0, p[index[1]] - p[index[0]]);
        A.setRow(1, p[index[2]] - p[index[0]]);
        A.setRow(2, p[index[3]] - p[index[0]]);
        Mat3<R> R_0_2;
        if (svd) {
            Mat3<R> Fm = A * initialTransformation[e];
            if (determinant(Fm) < 1e-6) {  // compared as double, :1140
                Decompose<R>::polarDecomposition_stable(Fm, R_0_2);
                R_0_2 = R_0_2.multTransposed(initialRotations[e]);
            } else Decompose<R>::polarDecomposition(A, R_0_2);
        } else Decompose<R>::polarDecomposition(A, R_0_2);
        rotations[e] = R_0_2.transposed();
        Coord deforme[4];
        for (int i = 0; i < 4; ++i) deforme[i] = R_0_2 * p[index[i]];
        const Coord* x0 = &X0[4 * e];
        R D[12];
        for (int n = 0; n < 4; ++n) for (int k = 0; k < 3; ++k) D[3 * n + k] = x0[n][k] - deforme[n][k];
        if (updateStiffnessMatrix) {   // :1063-1067 / :1174-1177
            computeStrainDisplacement(&J[12 * e], deforme[0], deforme[1], deforme[2], deforme[3]);
            for (int q = 0; q < 12; ++q) Jsh[12 * e + q] = J[12 * e + q];
        }
        R F[12];
        computeForce(F, D, &K[3 * e], &J[12 * e], &Jsh[12 * e], false, 0, plasticPtr(e), plastic);
        for (int i = 0; i < 12; i += 3) f[index[i / 3]] += rotations[e] * Coord(F[i], F[i + 1], F[i + 2]);
    }
    // addForce :1547-1604
    void addForce(VecDeriv<R>& f, const std::vector<Coord>& p) {
        f.resize(p.size());
        const size_t T = nbTets();
        for (size_t e = 0; e < T; ++e) {
            switch (method) {
            case SMALL: accumulateForceSmall(f, p, e); break;
            case LARGE: accumulateForceLarge(f, p, e); break;
            case POLAR: accumulateForcePolarLike(f, p, e, false); break;
            case SVD:   accumulateForcePolarLike(f, p, e, true); break;
            }
        }
    }
    void applyStiffnessSmall(VecDeriv<R>& f, const VecDeriv<R>& x, size_t i, SReal fact) {  // :724-748
        const uint32_t* t = &tets[4 * i];
        R X[12], F[12];
        for (int n = 0; n < 4; ++n) for (int k = 0; k < 3; ++k) X[3 * n + k] = x[t[n]][k];
        computeForce(F, X, &K[3 * i], &J[12 * i], &Jsh[12 * i], true, fact);
        for (int n = 0; n < 4; ++n) f[t[n]] += Coord(-F[3 * n], -F[3 * n + 1], -F[3 * n + 2]);
    }
    void applyStiffnessCorotational(VecDeriv<R>& f, const VecDeriv<R>& x, size_t i, SReal fact) {  // :1192-1237
        const uint32_t* t = &tets[4 * i];
        const Mat3<R>& rot = rotations[i];
        R X[12], F[12];
        for (int n = 0; n < 4; ++n) {
            const Coord& xn = x[t[n]];
            X[3 * n + 0] = rot(0, 0) * xn[0] + rot(1, 0) * xn[1] + rot(2, 0) * xn[2];
            X[3 * n + 1] = rot(0, 1) * xn[0] + rot(1, 1) * xn[1] + rot(2, 1) * xn[2];
            X[3 * n + 2] = rot(0, 2) * xn[0] + rot(1, 2) * xn[1] + rot(2, 2) * xn[2];
        }
        computeForce(F, X, &K[3 * i], &J[12 * i], &Jsh[12 * i], true, fact);
        for (int n = 0; n < 4; ++n) {
            Coord& fn = f[t[n]];
            fn[0] -= rot(0, 0) * F[3 * n] + rot(0, 1) * F[3 * n + 1] + rot(0, 2) * F[3 * n + 2];
            fn[1] -= rot(1, 0) * F[3 * n] + rot(1, 1) * F[3 * n + 1] + rot(1, 2) * F[3 * n + 2];
            fn[2] -= rot(2, 0) * F[3 * n] + rot(2, 1) * F[3 * n + 1] + rot(2, 2) * F[3 * n + 2];
        }
    }
    // ---- von Mises stress.  PARITY UNPINNED by reference vectors (no KAT in the reference tree).
    // invertMatrix, general case (Sofa/framework/Type/src/sofa/type/Mat.h:1103-1166): Gauss-Jordan with full pivoting, S = 4
    static bool invertMatrix4(R dest[4][4], const R from[4][4]) {
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
            if (rabs(pivot) <= std::numeric_limits<R>::epsilon()) return false;
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
    std::vector<R> elemLambda, elemMu;    // Lame coefficients of the element's material (:278-282)
    std::vector<R> elemShapeFun;          // elemShapeFun[e]: inverse of [1 x0 y0 z0] rows, 16 per element (:1521-1541)
    std::vector<R> vonMisesPerElement, vonMisesPerNode;
    void initVonMises() {
        const size_t T = nbTets();
        elemShapeFun.assign(16 * T, R(0));
        for (size_t i = 0; i < T; ++i) {
            R matVert[4][4], inv[4][4];
            for (int k = 0; k < 4; k++) { const uint32_t ix = tets[4 * i + k]; matVert[k][0] = R(1.0); for (int l = 1; l < 4; l++) matVert[k][l] = initialPoints[ix][l - 1]; }
            for (auto& rw : inv) for (auto& v : rw) v = R(0);
            invertMatrix4(inv, matVert);
            for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) elemShapeFun[16 * i + 4 * a + b] = inv[a][b];
        }
    }
    // computeVonMisesStress :2196-2416 (values only; the colour map is display code).  how = d_computeVonMisesStress (1 or 2).
    void computeVonMisesStress(const std::vector<Coord>& X, int how) {
        const size_t T = nbTets();
        if (elemShapeFun.size() != 16 * T) initVonMises();
        vonMisesPerElement.assign(T, R(0));
        for (size_t el = 0; el < T; ++el) {
            const uint32_t* index = &tets[4 * el];
            const R* shf = &elemShapeFun[16 * el];
            R vStrain[6];
            Mat3<R> gradU;
            if (how == 2) {
                Coord U[4];
                for (int m = 0; m < 4; ++m) U[m] = X[index[m]] - initialPoints[index[m]];
                for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) {
                    gradU(k, l) = 0.0;
                    for (int m = 0; m < 4; m++) gradU(k, l) += shf[4 * (l + 1) + m] * U[m][k];
                }
                const Mat3<R> gT = gradU.transposed();
                const Mat3<R> strain = ((gradU + gT) + gT * gradU) * R(0.5);     // (Real)0.5 * (gradU + gradU^T + gradU^T gradU)
                for (int i = 0; i < 3; i++) vStrain[i] = strain(i, i);
                vStrain[3] = strain(1, 2); vStrain[4] = strain(0, 2); vStrain[5] = strain(0, 1);
            } else {
                Mat3<R> R_0_2;
                R D[12];
                const Coord* x0 = &X0[4 * el];
                if (method == LARGE) {
                    computeRotationLarge(R_0_2, X, index[0], index[1], index[2]);
                    rotations[el] = R_0_2.transposed();
                    Coord deforme[4];
                    for (int i = 0; i < 4; ++i) deforme[i] = R_0_2 * X[index[i]];
                    deforme[1][0] -= deforme[0][0];
                    deforme[2][0] -= deforme[0][0];
                    deforme[2][1] -= deforme[0][1];
                    deforme[3] -= deforme[0];
                    D[0] = 0; D[1] = 0; D[2] = 0;
                    D[3] = x0[1][0] - deforme[1][0]; D[4] = 0; D[5] = 0;
                    D[6] = x0[2][0] - deforme[2][0]; D[7] = x0[2][1] - deforme[2][1]; D[8] = 0;
                    D[9] = x0[3][0] - deforme[3][0]; D[10] = x0[3][1] - deforme[3][1]; D[11] = x0[3][2] - deforme[3][2];
                } else {  // POLAR / SVD: always the plain polar decomposition here (:2291-2299)
                    Mat3<R> A;
                    A.setRow(0, X[index[1]] - X[index[0]]); A.setRow(1, X[index[2]] - X[index[0]]); A.setRow(2, X[index[3]] - X[index[0]]);
                    Decompose<R>::polarDecomposition(A, R_0_2);
                    rotations[el] = R_0_2.transposed();
                    Coord deforme[4];
                    for (int i = 0; i < 4; ++i) deforme[i] = R_0_2 * X[index[i]];
                    for (int n = 0; n < 4; ++n) for (int k = 0; k < 3; ++k) D[3 * n + k] = x0[n][k] - deforme[n][k];
                }
                for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) {
                    gradU(k, l) = 0.0;
                    for (int m = 0; m < 4; m++) gradU(k, l) += shf[4 * (l + 1) + m] * D[3 * m + k];
                }
                const Mat3<R> strain = (gradU + gradU.transposed()) * R(0.5);
                for (int i = 0; i < 3; i++) vStrain[i] = strain(i, i);
                vStrain[3] = strain(1, 2); vStrain[4] = strain(0, 2); vStrain[5] = strain(0, 1);
            }
            const R lambda = elemLambda[el], mu = elemMu[el];
            R s[6];
            R traceStrain = 0.0;
            for (int k = 0; k < 3; k++) { traceStrain += vStrain[k]; s[k] = vStrain[k] * 2 * mu; }
            for (int k = 3; k < 6; k++) s[k] = vStrain[k] * 2 * mu;
            for (int k = 0; k < 3; k++) s[k] += lambda * traceStrain;
            R v = Decompose<R>::rsqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2] - s[0] * s[1] - s[1] * s[2] - s[2] * s[0] + 3 * s[3] * s[3] + 3 * s[4] * s[4] + 3 * s[5] * s[5]);
            if (v < 1e-10) v = 0.0;
            vonMisesPerElement[el] = v;
        }
        const size_t N = X.size();
        std::vector<std::vector<uint32_t>> around(N);
        for (size_t t = 0; t < T; ++t) for (int k = 0; k < 4; ++k) around[tets[4 * t + k]].push_back(uint32_t(t));
        vonMisesPerNode.assign(N, R(0));
        for (size_t dof = 0; dof < N; dof++) {
            R a = 0.0;
            for (size_t at = 0; at < around[dof].size(); at++) a += vonMisesPerElement[around[dof][at]];
            if (!around[dof].empty()) a /= R(around[dof].size());
            vonMisesPerNode[dof] = a;
        }
    }

    // getRotation :781-833 (what WarpPreconditioner / RotationMatrix consumers read): the mean of rotations[t] * R0(t) over the
    // tetrahedra around the node (TetrahedraAroundVertex lists them in ascending index), made orthogonal by polarDecomposition.
    // A node without tetrahedra takes the element _rotationIdx names (:803-808; the array is zero-filled, :850-853 overwrite it).
    void getRotation(Mat3<R>& Rn, uint32_t nodeIdx, const std::vector<std::vector<uint32_t>>& tetrahedraAroundVertex) const {
        if (method == SMALL) { Rn.identity(); return; }
        const std::vector<uint32_t>& liste = tetrahedraAroundVertex[nodeIdx];
        Rn = Mat3<R>();
        const std::size_t numTetra = liste.size();
        if (numTetra == 0) {
            if (!rotationIdx.empty()) Rn = rotations[rotationIdx[nodeIdx]] * initialRotations[rotationIdx[nodeIdx]].transposed();
            else Rn.identity();
            return;
        }
        for (std::size_t ti = 0; ti < numTetra; ++ti) {
            const Mat3<R> prod = rotations[liste[ti]] * initialRotations[liste[ti]].transposed();
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rn.m[i][j] += prod.m[i][j];   // Mat.h operator+=
        }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rn.m[i][j] = Rn.m[i][j] / numTetra;   // Real / size_t, as written at :823-825
        Mat3<R> Rmoy;
        Decompose<R>::polarDecomposition(Rn, Rmoy);
        Rn = Rmoy;
    }
    // TetrahedralCorotationalFEMForceField::getRotation, TetrahedralCorotationalFEMForceField.inl:779-820 (large / polar)
    void getRotationSibling(Mat3<R>& Rn, const std::vector<uint32_t>& liste) const {
        const int numNeiTetra = int(liste.size());
        Mat3<R> r;
        for (int i = 0; i < numNeiTetra; i++) {
            const uint32_t e = liste[i];
            Mat3<R> r01;
            if (method == POLAR) {   // initPolar :1046-1070: initialTransformation = A
                r01.setRow(0, initialPoints[tets[4 * e + 1]] - initialPoints[tets[4 * e]]);
                r01.setRow(1, initialPoints[tets[4 * e + 2]] - initialPoints[tets[4 * e]]);
                r01.setRow(2, initialPoints[tets[4 * e + 3]] - initialPoints[tets[4 * e]]);
            } else r01 = initialRotations[e].transposed();   // initLarge :840-881: initialTransformation = R_0_1
            const Mat3<R> r21 = rotations[e] * r01;
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) r.m[a][b] += r21.m[a][b];
        }
        Rn = r * R(1.0f / numNeiTetra);
        Vec3<R> ex = Rn.row(0), ey = Rn.row(1);
        ex.normalize(); ey.normalize();
        Vec3<R> ez = cross(ex, ey); ez.normalize();
        ey = cross(ez, ex); ey.normalize();
        Rn.setRow(0, ex); Rn.setRow(1, ey); Rn.setRow(2, ez);
    }
    // getRotations(VecReal&) :2033-2042: 9 Reals per node, row-major
    void getRotations(std::vector<Mat3<R>>& vecR, size_t nbdof) const {
        std::vector<std::vector<uint32_t>> around(nbdof);
        for (size_t t = 0; t < nbTets(); ++t) for (int k = 0; k < 4; ++k) around[tets[4 * t + k]].push_back(uint32_t(t));   // TetrahedronSetTopologyContainer::createTetrahedraAroundVertexArray
        vecR.assign(nbdof, Mat3<R>());
        for (uint32_t i = 0; i < nbdof; ++i) { if (tetrahedralCorotational) getRotationSibling(vecR[i], around[i]); else getRotation(vecR[i], i, around); }
    }
    // addDForce :1606-1636.  kFactor already includes the Rayleigh term
    // (MechanicalParams.h:62) and is narrowed to Real there (:1615).
    void addDForce(VecDeriv<R>& df, const VecDeriv<R>& dx, SReal kFactorIncludingRayleigh) {
        df.resize(dx.size());
        const R kFactor = R(kFactorIncludingRayleigh);
        const size_t T = nbTets();
        if (method == SMALL) for (size_t i = 0; i < T; ++i) applyStiffnessSmall(df, dx, i, kFactor);
        else                 for (size_t i = 0; i < T; ++i) applyStiffnessCorotational(df, dx, i, kFactor);
    }
};

// ---------------------------------------------------------------------------
// HexahedronFEMForceField<Vec3Types>
// Sofa/Component/SolidMechanics/FEM/Elastic/src/sofa/component/solidmechanics/fem/elastic/HexahedronFEMForceField.inl
// (methods: 0 = large, 1 = polar, 2 = small, as setMethod() numbers them)
// ---------------------------------------------------------------------------
enum HexMethod { HEX_LARGE = 0, HEX_POLAR = 1, HEX_SMALL = 2 };

template <class R> struct HexaFEM {
    typedef Vec3<R> Coord;
    int method = HEX_LARGE;
    std::vector<uint32_t> hexas;            // 8*H
    std::vector<Coord> initialPoints;
    std::vector<R> young, poisson;
    std::vector<R> Kmat;                    // _materialsStiffnesses: {U, V, W} per element (:684-709)
    std::vector<R> Ke;                      // d_elementStiffnesses: 576 per element, row-major 24x24
    std::vector<Mat3<R>> rotations;         // _rotations[e]  (= R, NOT transposed -- opposite of the tetra class)
    std::vector<Mat3<R>> initialRotations;  // _initialrotations
    std::vector<Coord> X0;                  // _rotatedInitialElements: 8 per element
    double potentialEnergy = 0;

    size_t nbHexas() const { return hexas.size() / 8; }
    R youngIn(size_t e) const { return young.size() > e ? young[e] : young[0]; }
    R poissonIn(size_t e) const { return poisson.size() > e ? poisson[e] : poisson[0]; }
    static int coef(int i, int c) {  // _coef, :66-89
        static const int t[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
        return t[i][c];
    }
    void computeMaterialStiffness(size_t i) {  // :683-709
        const R poissonRatio = poissonIn(i);
        R U = 1, V = poissonRatio / (1 - poissonRatio), W = (1 - 2 * poissonRatio) / (2 * (1 - poissonRatio));
        const R s = (youngIn(i) * (1 - poissonRatio)) / ((1 + poissonRatio) * (1 - 2 * poissonRatio));
        Kmat[3 * i] = U * s; Kmat[3 * i + 1] = V * s; Kmat[3 * i + 2] = W * s;
    }
    // computeElementStiffness :306-536, GENERIC_STIFFNESS_MATRIX + MAT_STIFFNESS_USE_W + DN_USE_J path
    // (the integrateStiffness() result is discarded by `K=K1`, :504-506, so it is not restated).
    void computeElementStiffness(R* K, const R* M, const Coord* nodes, double stiffnessFactor) const {
        for (int i = 0; i < 576; ++i) K[i] = 0;
        Mat3<R> J, J_1;
        R detJ = R(1.0);
        Coord lx = nodes[1] - nodes[0], ly = nodes[3] - nodes[0], lz = nodes[4] - nodes[0];
        bool isParallel = false;
        if ((nodes[3] + lx - nodes[2]).norm() < lx.norm() * 0.001 && (nodes[0] + lz - nodes[4]).norm() < lz.norm() * 0.001 &&
            (nodes[1] + lz - nodes[5]).norm() < lz.norm() * 0.001 && (nodes[2] + lz - nodes[6]).norm() < lz.norm() * 0.001 &&
            (nodes[3] + lz - nodes[7]).norm() < lz.norm() * 0.001) {
            isParallel = true;
            for (int c = 0; c < 3; ++c) { J(c, 0) = lx[c] / 2; J(c, 1) = ly[c] / 2; J(c, 2) = lz[c] / 2; }
            detJ = determinant(J);
            invertMatrix(J_1, J);
        }
        const R U = M[0], V = M[1], W = M[2];
        const double inv_sqrt3 = 1.0 / std::sqrt(3.0);
        for (int gx1 = -1; gx1 <= 1; gx1 += 2) for (int gx2 = -1; gx2 <= 1; gx2 += 2) for (int gx3 = -1; gx3 <= 1; gx3 += 2) {
            const double x1 = gx1 * inv_sqrt3, x2 = gx2 * inv_sqrt3, x3 = gx3 * inv_sqrt3;
            if (!isParallel) {
                for (int c = 0; c < 3; ++c) {
                    J(c, 0) = (R)((nodes[1][c] - nodes[0][c]) * (1 - x2) * (1 - x3) / 8 + (nodes[2][c] - nodes[3][c]) * (1 + x2) * (1 - x3) / 8 + (nodes[5][c] - nodes[4][c]) * (1 - x2) * (1 + x3) / 8 + (nodes[6][c] - nodes[7][c]) * (1 + x2) * (1 + x3) / 8);
                    J(c, 1) = (R)((nodes[3][c] - nodes[0][c]) * (1 - x1) * (1 - x3) / 8 + (nodes[2][c] - nodes[1][c]) * (1 + x1) * (1 - x3) / 8 + (nodes[7][c] - nodes[4][c]) * (1 - x1) * (1 + x3) / 8 + (nodes[6][c] - nodes[5][c]) * (1 + x1) * (1 + x3) / 8);
                    J(c, 2) = (R)((nodes[4][c] - nodes[0][c]) * (1 - x1) * (1 - x2) / 8 + (nodes[5][c] - nodes[1][c]) * (1 + x1) * (1 - x2) / 8 + (nodes[6][c] - nodes[2][c]) * (1 + x1) * (1 + x2) / 8 + (nodes[7][c] - nodes[3][c]) * (1 - x1) * (1 + x2) / 8);
                }
                detJ = determinant(J);
                invertMatrix(J_1, J);
            }
            R qx[8], qy[8], qz[8];
            for (int i = 0; i < 8; ++i) {
                R dNi_dx1 = (R)((coef(i, 0)) * (1 + coef(i, 1) * x2) * (1 + coef(i, 2) * x3) / 8.0);
                R dNi_dx2 = (R)((1 + coef(i, 0) * x1) * (coef(i, 1)) * (1 + coef(i, 2) * x3) / 8.0);
                R dNi_dx3 = (R)((1 + coef(i, 0) * x1) * (1 + coef(i, 1) * x2) * (coef(i, 2)) / 8.0);
                qx[i] = dNi_dx1 * J_1(0, 0) + dNi_dx2 * J_1(1, 0) + dNi_dx3 * J_1(2, 0);
                qy[i] = dNi_dx1 * J_1(0, 1) + dNi_dx2 * J_1(1, 1) + dNi_dx3 * J_1(2, 1);
                qz[i] = dNi_dx1 * J_1(0, 2) + dNi_dx2 * J_1(1, 2) + dNi_dx3 * J_1(2, 2);
            }
            for (int i = 0; i < 8; ++i) {
                R MBi[6][3];
                MBi[0][0] = U * qx[i]; MBi[0][1] = V * qy[i]; MBi[0][2] = V * qz[i];
                MBi[1][0] = V * qx[i]; MBi[1][1] = U * qy[i]; MBi[1][2] = V * qz[i];
                MBi[2][0] = V * qx[i]; MBi[2][1] = V * qy[i]; MBi[2][2] = U * qz[i];
                MBi[3][0] = W * qy[i]; MBi[3][1] = W * qx[i]; MBi[3][2] = (R)0;
                MBi[4][0] = (R)0;      MBi[4][1] = W * qz[i]; MBi[4][2] = W * qy[i];
                MBi[5][0] = W * qz[i]; MBi[5][1] = (R)0;      MBi[5][2] = W * qx[i];
                for (int j = i; j < 8; ++j) {
                    Mat3<R> k;
                    k(0, 0) = qx[j] * MBi[0][0] + qy[j] * MBi[3][0] + qz[j] * MBi[5][0];
                    k(0, 1) = qx[j] * MBi[0][1] + qy[j] * MBi[3][1];
                    k(0, 2) = qx[j] * MBi[0][2] + qz[j] * MBi[5][2];
                    k(1, 0) = qy[j] * MBi[1][0] + qx[j] * MBi[3][0];
                    k(1, 1) = qy[j] * MBi[1][1] + qx[j] * MBi[3][1] + qz[j] * MBi[4][1];
                    k(1, 2) = qy[j] * MBi[1][2] + qz[j] * MBi[4][2];
                    k(2, 0) = qz[j] * MBi[2][0] + qx[j] * MBi[5][0];
                    k(2, 1) = qz[j] * MBi[2][1] + qy[j] * MBi[4][1];
                    k(2, 2) = qz[j] * MBi[2][2] + qy[j] * MBi[4][2] + qx[j] * MBi[5][2];
                    k *= detJ;
                    for (int m = 0; m < 3; ++m) for (int l = 0; l < 3; ++l) K[(i * 3 + m) * 24 + (j * 3 + l)] += k(l, m);
                }
            }
        }
        for (int i = 0; i < 24; ++i) for (int j = i + 1; j < 24; ++j) K[j * 24 + i] = K[i * 24 + j];
        const R sf = (R)stiffnessFactor;
        for (int i = 0; i < 576; ++i) K[i] *= sf;
    }
    static void computeRotationLarge(Mat3<R>& r, Coord& edgex, Coord& edgey) {  // :816-834
        edgex.normalize();
        const Coord edgez = cross(edgex, edgey).normalized();
        edgey = cross(edgez, edgex);
        r.setRow(0, edgex); r.setRow(1, edgey); r.setRow(2, edgez);
    }
    static void meanEdges(const Coord* n, Coord& ex, Coord& ey, Coord& ez) {  // :797-801, :920-931
        ex = (n[1] - n[0] + n[2] - n[3] + n[5] - n[4] + n[6] - n[7]) * R(.25);
        ey = (n[3] - n[0] + n[2] - n[1] + n[7] - n[4] + n[6] - n[5]) * R(.25);
        ez = (n[4] - n[0] + n[5] - n[1] + n[7] - n[3] + n[6] - n[2]) * R(.25);
    }
    void computeRotation(Mat3<R>& r, const Coord* n) const {
        Coord ex, ey, ez; meanEdges(n, ex, ey, ez);
        if (method == HEX_LARGE) computeRotationLarge(r, ex, ey);
        else if (method == HEX_POLAR) {  // computeRotationPolar :918-943
            Mat3<R> A; A.setRow(0, ex); A.setRow(1, ey); A.setRow(2, ez);
            Decompose<R>::polarDecomposition(A, r);
        } else r.identity();
    }
    void reinit(const std::vector<Coord>& restPosition) {  // :125-180, initLarge :788-814, initPolar :886-916, initSmall :722-738
        initialPoints = restPosition;
        const size_t H = nbHexas();
        Kmat.assign(3 * H, 0); Ke.assign(576 * H, 0);
        rotations.assign(H, Mat3<R>()); initialRotations.assign(H, Mat3<R>()); X0.assign(8 * H, Coord());
        for (size_t i = 0; i < H; ++i) {
            computeMaterialStiffness(i);
            Coord nodes[8];
            for (int w = 0; w < 8; ++w) nodes[w] = initialPoints[hexas[8 * i + w]];
            computeRotation(rotations[i], nodes);
            initialRotations[i] = rotations[i];
            for (int w = 0; w < 8; ++w) X0[8 * i + w] = rotations[i] * nodes[w];
            computeElementStiffness(&Ke[576 * i], &Kmat[3 * i], &X0[8 * i], 1.0);
        }
    }
    static void computeForce(R F[24], const R D[24], const R* K) {  // :711-715  F = K*Depl (Mat.h:577-587)
        for (int i = 0; i < 24; ++i) { F[i] = K[i * 24] * D[0]; for (int j = 1; j < 24; ++j) F[i] += K[i * 24 + j] * D[j]; }
    }
    // addForce :194-246 ; accumulateForce{Small,Large,Polar} :740-786, :836-884, :1027-1076
    void addForce(VecDeriv<R>& f, const std::vector<Coord>& p) {
        f.resize(p.size());
        potentialEnergy = 0;
        const size_t H = nbHexas();
        for (size_t i = 0; i < H; ++i) {
            const uint32_t* elem = &hexas[8 * i];
            Coord nodes[8], deformed[8];
            for (int w = 0; w < 8; ++w) nodes[w] = p[elem[w]];
            if (method == HEX_SMALL) { for (int w = 0; w < 8; ++w) deformed[w] = nodes[w]; }
            else {
                computeRotation(rotations[i], nodes);
                for (int w = 0; w < 8; ++w) deformed[w] = rotations[i] * nodes[w];
            }
            R D[24], F[24];
            for (int k = 0; k < 8; ++k) for (int j = 0; j < 3; ++j) D[3 * k + j] = X0[8 * i + k][j] - deformed[k][j];
            computeForce(F, D, &Ke[576 * i]);
            for (int w = 0; w < 8; ++w) {
                const Coord Fw(F[3 * w], F[3 * w + 1], F[3 * w + 2]);
                if (method == HEX_SMALL) f[elem[w]] += Fw;
                else f[elem[w]] += rotations[i].multTranspose(Fw);
            }
            for (int w = 0; w < 8; ++w)
                potentialEnergy += dot(Coord(F[3 * w], F[3 * w + 1], F[3 * w + 2]), -Coord(D[3 * w], D[3 * w + 1], D[3 * w + 2]));
        }
        potentialEnergy /= -2.0;
    }
    // addDForce :248-286
    // getNodeRotation :946-974: starts from the IDENTITY (not zero), adds _rotations[h] * _initialrotations[h]^T over the hexahedra around
    // the node (ascending index), divides by their number and takes the polar decomposition.  getRotations :976-1023 stores it row-major.
    void getNodeRotation(Mat3<R>& Rn, const std::vector<uint32_t>& liste_hexa) const {
        Rn.identity();
        const std::size_t numHexa = liste_hexa.size();
        for (std::size_t ti = 0; ti < numHexa; ti++) {
            const Mat3<R> prod = rotations[liste_hexa[ti]] * initialRotations[liste_hexa[ti]].transposed();
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rn.m[i][j] += prod.m[i][j];
        }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rn.m[i][j] = Rn.m[i][j] / numHexa;
        Mat3<R> Rmoy;
        Decompose<R>::polarDecomposition(Rn, Rmoy);
        Rn = Rmoy;
    }
    void getRotations(std::vector<Mat3<R>>& vecR, size_t nbdof) const {
        std::vector<std::vector<uint32_t>> around(nbdof);
        for (size_t h = 0; h < hexas.size() / 8; ++h) for (int k = 0; k < 8; ++k) around[hexas[8 * h + k]].push_back(uint32_t(h));
        vecR.assign(nbdof, Mat3<R>());
        for (size_t i = 0; i < nbdof; ++i) getNodeRotation(vecR[i], around[i]);
    }
    void addDForce(VecDeriv<R>& df, const VecDeriv<R>& dx, SReal kFactorIncludingRayleigh) {
        const R kFactor = (R)kFactorIncludingRayleigh;
        if (df.size() != dx.size()) df.resize(dx.size());
        const size_t H = nbHexas();
        for (size_t i = 0; i < H; ++i) {
            const uint32_t* elem = &hexas[8 * i];
            R X[24], F[24];
            for (int w = 0; w < 8; ++w) {
                const Coord x_2 = rotations[i] * dx[elem[w]];
                X[3 * w] = x_2[0]; X[3 * w + 1] = x_2[1]; X[3 * w + 2] = x_2[2];
            }
            computeForce(F, X, &Ke[576 * i]);
            for (int w = 0; w < 8; ++w)
                df[elem[w]] -= rotations[i].multTranspose(Coord(F[3 * w], F[3 * w + 1], F[3 * w + 2])) * kFactor;
        }
    }
};

// ---------------------------------------------------------------------------
// DiagonalMass<Vec3Types>   Sofa/Component/Mass/src/sofa/component/mass/DiagonalMass.inl
// ---------------------------------------------------------------------------
template <class R> struct DiagonalMass {
    std::vector<R> vertexMass;  // d_vertexMass
    R massDensity = 1;
    R totalMass = 0;
    // UniformMass (Sofa/Component/Mass/src/sofa/component/mass/UniformMass.inl): one MassType for every node.  Kept in the same
    // struct because the solver node holds exactly one mass component.  PARITY UNPINNED by reference vectors (no numerical KAT
    // for its addMDx / addForce in the reference tree).
    bool uniform = false;
    R uniformVertexMass = 0;    // d_vertexMass of UniformMass
    void initUniformFromVertexMass(R m, size_t n) { uniform = true; uniformVertexMass = m; vertexMass.assign(n, m); totalMass = R(double(m) * double(n)); }   // :300-309
    void initUniformFromTotalMass(double tm, size_t n) {   // :329-345: *m = d_totalMass.getValue() / Real(size)
        uniform = true; totalMass = R(tm);
        uniformVertexMass = n > 0 ? R(tm / R(n)) : R(0);
        vertexMass.assign(n, uniformVertexMass);
    }
    // computeVertexMass :1000-1100 (tetrahedra branch :1061-1079, hexahedra branch :1080-1100)
    R computeVertexMassTets(R density, const std::vector<Vec3<R>>& pos, const std::vector<uint32_t>& tets) {
        vertexMass.assign(pos.size(), R(0));
        R total_mass = R(0);
        for (size_t i = 0; i < tets.size() / 4; ++i) {
            const uint32_t* t = &tets[4 * i];
            const Vec3<R> a = pos[t[1]] - pos[t[0]], b = pos[t[2]] - pos[t[0]], c = pos[t[3]] - pos[t[0]];
            const R tetraVolume = std::abs(dot(cross(a, b), c) / R(6));
            const R mass = (density * tetraVolume) / R(4.0);
            for (int j = 0; j < 4; ++j) { vertexMass[t[j]] += mass; total_mass += mass; }
        }
        return total_mass;
    }
    static R tetVol(const Vec3<R>& n0, const Vec3<R>& n1, const Vec3<R>& n2, const Vec3<R>& n3) {  // Tetrahedron.h:55-82
        return std::abs(dot(cross(n1 - n0, n2 - n0), n3 - n0) / R(6));
    }
    // hexahedra branch :1080-1110; volume = sum of the 6 inner tetrahedra, Sofa/framework/Geometry/src/sofa/geometry/Hexahedron.h:242-254
    R computeVertexMassHexas(R density, const std::vector<Vec3<R>>& pos, const std::vector<uint32_t>& hexas) {
        vertexMass.assign(pos.size(), R(0));
        R total_mass = R(0);
        for (size_t i = 0; i < hexas.size() / 8; ++i) {
            const uint32_t* h = &hexas[8 * i];
            const Vec3<R>&n0 = pos[h[0]], &n1 = pos[h[1]], &n2 = pos[h[2]], &n3 = pos[h[3]], &n4 = pos[h[4]], &n5 = pos[h[5]], &n6 = pos[h[6]], &n7 = pos[h[7]];
            const R hexaVolume = tetVol(n0, n5, n1, n6) + tetVol(n0, n1, n3, n6) + tetVol(n1, n3, n6, n2) + tetVol(n6, n3, n0, n7) + tetVol(n6, n7, n0, n5) + tetVol(n7, n5, n4, n0);
            const R mass = (density * hexaVolume) / R(8.0);
            for (int j = 0; j < 8; ++j) { vertexMass[h[j]] += mass; total_mass += mass; }
        }
        return total_mass;
    }
    // elemSize: 4 = tetrahedra topology, 8 = hexahedra topology (m_massTopologyType)
    R computeVertexMass(R density, const std::vector<Vec3<R>>& pos, const std::vector<uint32_t>& elems, int elemSize) {
        return elemSize == 8 ? computeVertexMassHexas(density, pos, elems) : computeVertexMassTets(density, pos, elems);
    }
    void initFromMassDensity(R md, const std::vector<Vec3<R>>& pos, const std::vector<uint32_t>& elems, int elemSize = 4) {  // :1218-1230
        massDensity = md;
        totalMass = computeVertexMass(md, pos, elems, elemSize);
    }
    void initFromTotalMass(R tm, const std::vector<Vec3<R>>& pos, const std::vector<uint32_t>& elems, int elemSize = 4) {  // :1233-1258
        totalMass = tm;
        const R sumMass = computeVertexMass(R(1.0), pos, elems, elemSize);
        if (sumMass < std::numeric_limits<R>::epsilon()) massDensity = R(1.0);
        else massDensity = R(totalMass / sumMass);
        for (auto& vm : vertexMass) vm *= massDensity;
    }
    // addMDx :535-559
    void addMDx(VecDeriv<R>& res, const VecDeriv<R>& dx, SReal factor) const {
        size_t n = vertexMass.size();
        if (dx.size() < n) n = dx.size();
        if (res.size() < n) n = res.size();
        if (uniform) {   // UniformMass::addMDx, UniformMass.inl:403-420
            R m = uniformVertexMass;
            if (factor != 1.0) m *= R(factor);
            for (size_t i = 0; i < n; ++i) res[i] += dx[i] * m;
            return;
        }
        if (factor == 1.0) for (size_t i = 0; i < n; ++i) res[i] += dx[i] * vertexMass[i];
        else for (size_t i = 0; i < n; ++i) res[i] += (dx[i] * vertexMass[i]) * R(factor);
    }
    // addForce :1392-1413 (gravity as Vec3d narrowed to Deriv)
    void addForce(VecDeriv<R>& f, const double g[3]) const {
        const Vec3<R> theGravity((R)g[0], (R)g[1], (R)g[2]);
        if (uniform) {   // UniformMass::addForce, UniformMass.inl:469-496: mg = theGravity * m once, then f[i] += mg
            const Vec3<R> mg = theGravity * uniformVertexMass;
            for (size_t i = 0; i < vertexMass.size(); ++i) f[i] += mg;
            return;
        }
        for (size_t i = 0; i < vertexMass.size(); ++i) f[i] += theGravity * vertexMass[i];
    }
};

// ---------------------------------------------------------------------------
// MeshMatrixMass<Vec3Types> on a tetrahedral topology
// Sofa/Component/Mass/src/sofa/component/mass/MeshMatrixMass.inl
// PARITY UNPINNED by reference vectors: the reference's MeshMatrixMass_test.cpp checks totals and vertex/edge masses of tiny meshes
// (reproduced in tests/test_oracle_golden.py), not addMDx outputs.
// ---------------------------------------------------------------------------
template <class R> struct MeshMatrixMass {
    std::vector<R> vertexMass, edgeMass;      // d_vertexMass / d_edgeMass
    std::vector<uint32_t> edges;              // 2*E, l_topology->getEdges()
    bool lumping = false;                     // d_lumping
    R massLumpingCoeff = R(0);                // m_massLumpingCoeff (:994; 2.5 on tetrahedra whether lumped or not, :1484)
    double totalMass = 0;
    // TetrahedronSetTopologyContainer::createEdgeSetArray (Topology/Container/Dynamic/.../TetrahedronSetTopologyContainer.cpp): edges in order of
    // first appearance over the tetrahedra, local edges {0,1},{0,2},{0,3},{1,2},{1,3},{2,3} (core/topology/Topology.cpp:44), vertices sorted.
    // edgesInTet: 6 edge ids per tetrahedron (m_edgesInTetrahedron).
    static void createEdgeSetArray(const std::vector<uint32_t>& tets, std::vector<uint32_t>& edgesOut, std::vector<uint32_t>& edgesInTet) {
        static const int L[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
        std::map<std::pair<uint32_t, uint32_t>, uint32_t> edgeMap;
        edgesOut.clear(); edgesInTet.assign(tets.size() / 4 * 6, 0);
        for (size_t i = 0; i < tets.size() / 4; ++i)
            for (int j = 0; j < 6; ++j) {
                const uint32_t v1 = tets[4 * i + L[j][0]], v2 = tets[4 * i + L[j][1]];
                const std::pair<uint32_t, uint32_t> e = v1 < v2 ? std::make_pair(v1, v2) : std::make_pair(v2, v1);
                auto it = edgeMap.find(e);
                if (it == edgeMap.end()) { it = edgeMap.emplace(e, uint32_t(edgeMap.size())).first; edgesOut.push_back(e.first); edgesOut.push_back(e.second); }
                edgesInTet[6 * i + j] = it->second;
            }
    }
    // massInitialization on tetrahedra :1476-1490 with applyVertexMassTetrahedronCreation :547-603 and applyEdgeMassTetrahedronCreation :606-665,
    // uniform d_massDensity
    void initFromMassDensityTets(R density, const std::vector<Vec3<R>>& pos, const std::vector<uint32_t>& tets, bool lumped) {
        lumping = lumped;
        massLumpingCoeff = R(2.5);
        std::vector<uint32_t> eit;
        createEdgeSetArray(tets, edges, eit);
        vertexMass.assign(pos.size(), R(0)); edgeMass.assign(edges.size() / 2, R(0));
        totalMass = 0;
        auto volume = [&](const uint32_t* t) {   // sofa::geometry::Tetrahedron::volume (Geometry/.../Tetrahedron.h:55-82)
            const Vec3<R> a = pos[t[1]] - pos[t[0]], b = pos[t[2]] - pos[t[0]], c = pos[t[3]] - pos[t[0]];
            return R(std::abs(dot(cross(a, b), c) / R(6)));
        };
        if (!lumping)
            for (size_t i = 0; i < tets.size() / 4; ++i) {
                const R mass = (density * volume(&tets[4 * i])) / R(20.0);
                for (int j = 0; j < 6; ++j) edgeMass[eit[6 * i + j]] += mass;
                totalMass += 6.0 * mass * 2.0;
            }
        for (size_t i = 0; i < tets.size() / 4; ++i) {
            const R mass = (density * volume(&tets[4 * i])) / R(10.0);
            for (int j = 0; j < 4; ++j) vertexMass[tets[4 * i + j]] += mass;
            totalMass += !lumping ? 4.0 * mass : 4.0 * mass * massLumpingCoeff;
        }
    }
    // addMDx :1987-2048
    void addMDx(VecDeriv<R>& res, const VecDeriv<R>& dx, SReal factor) const {
        if (lumping) {
            for (size_t i = 0; i < dx.size(); i++) res[i] += dx[i] * vertexMass[i] * massLumpingCoeff * R(factor);
        } else {
            for (size_t i = 0; i < dx.size(); i++) res[i] += dx[i] * vertexMass[i] * R(factor);
            for (size_t j = 0; j < edges.size() / 2; ++j) {
                const R tempMass = edgeMass[j] * R(factor);
                res[edges[2 * j]] += dx[edges[2 * j + 1]] * tempMass;
                res[edges[2 * j + 1]] += dx[edges[2 * j]] * tempMass;
            }
        }
    }
    // accFromF :2050-2069 (lumped only; the reference refuses otherwise)
    bool accFromF(VecDeriv<R>& a, const VecDeriv<R>& f) const {
        if (!lumping) return false;
        for (size_t i = 0; i < vertexMass.size(); i++) a[i] = f[i] / (vertexMass[i] * massLumpingCoeff);
        return true;
    }
    // addForce :2072-2092
    void addForce(VecDeriv<R>& f, const double g[3]) const {
        const Vec3<R> theGravity = Vec3<R>(R(g[0]), R(g[1]), R(g[2]));
        for (size_t i = 0; i < f.size(); ++i) f[i] += theGravity * vertexMass[i] * massLumpingCoeff;
    }
};


// PlaneForceField  Sofa/Component/MechanicalLoad/src/sofa/component/mechanicalload/PlaneForceField.inl
// (present in every SofaCUDA FEM benchmark scene).  setPlane :139-145, addForce :158-205, addDForce :208-226.
// PARITY UNPINNED by reference vectors: the reference tree holds no numerical KAT for this class, only the behaviour test
// MechanicalLoad/tests/PlaneForceField_test.cpp (reproduced in tests/test_oracle_golden.py).
// ---------------------------------------------------------------------------
// FastTetrahedralCorotationalForceField<DataTypes>   
// Sofa/Component/SolidMechanics/FEM/Elastic/src/sofa/component/solidmechanics/fem/elastic/FastTetrahedralCorotationalForceField.{h,inl}
// Per tetrahedron: four shape vectors, six 3x3 edge blocks of the linear stiffness (linearDfDx), the rest edge vectors and the rotation of the
// last addForce; addDForce runs over the EDGES of the topology with one 3x3 matrix per edge, assembled from the tetrahedra at the first call after
// each addForce.  Edge numbering: TetrahedronSetTopologyContainer::createEdgesInTetrahedronArray (see MeshMatrixMass::createEdgeSetArray above),
// or the caller's edge list when the topology already holds one.
// ---------------------------------------------------------------------------
enum FastMethod { FAST_POLAR = 0, FAST_QR = 1, FAST_POLAR2 = 2, FAST_LINEAR = 3 };   // RotationDecompositionMethod, .h:70-76 (d_method "polar" / "qr","large" / "polar2" / "none","linear","small", .inl:191-203)
template <class R> struct FastTetFEM {
    typedef Vec3<R> Coord;
    std::vector<uint32_t> tets;           // 4 * T
    std::vector<uint32_t> edges;          // 2 * E  (l_topology->getEdges())
    std::vector<uint32_t> edgesInTet;     // 6 * T  (getEdgesInTetrahedron)
    std::vector<R> young, poisson;
    int method = FAST_QR;
    struct TetrahedronRestInformation {   // .h:83-106
        Coord shapeVector[4]; R restVolume; Coord restEdgeVector[6]; Mat3<R> linearDfDxDiag[4]; Mat3<R> linearDfDx[6]; Mat3<R> rotation; Mat3<R> restRotation; R edgeOrientation[6];
    };
    std::vector<TetrahedronRestInformation> tetrahedronInfo;   // d_tetrahedronInfo
    std::vector<Mat3<R>> edgeInfo;                             // d_edgeInfo
    bool updateMatrix = true;
    static constexpr int L[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};   // edgesInTetrahedronArray, core/topology/Topology.cpp:44

    size_t nbTets() const { return tets.size() / 4; }
    R youngIn(size_t e) const { return young.size() > e ? young[e] : young[0]; }      // BaseLinearElasticityFEMForceField.inl:112-137
    R poissonIn(size_t e) const { return poisson.size() > e ? poisson[e] : poisson[0]; }

    static void computeQRRotation(Mat3<R>& r, const Coord* dp) {   // .inl:272-294
        const Coord edgex = dp[0].normalized();
        Coord edgey = dp[1];
        const Coord edgez = cross(edgex, edgey).normalized();
        edgey = cross(edgez, edgex);
        r.setRow(0, edgex); r.setRow(1, edgey); r.setRow(2, edgez);
    }
    void createTetrahedronRestInformation(size_t i, TetrahedronRestInformation& my_tinfo, const std::vector<Coord>& restPosition) {   // .inl:38-150
        const R youngModulusElement = youngIn(i), poissonRatioElement = poissonIn(i);
        // toLameParameters<3, Real>, impl/LameParameters.h:57-65
        R mu = youngModulusElement / (2 * (1 + poissonRatioElement));
        R lambda = youngModulusElement * poissonRatioElement / ((1 + poissonRatioElement) * (1 - (3 - 1) * poissonRatioElement));
        Coord point[4];
        const uint32_t* t = &tets[4 * i];
        for (int j = 0; j < 4; ++j) point[j] = restPosition[t[j]];
        // geometry::Tetrahedron::signedVolume, Geometry/src/sofa/geometry/Tetrahedron.h:73-83
        const R tetrahedronVolume = -(dot(cross(point[1] - point[0], point[2] - point[0]), point[3] - point[0]) / R(6));
        my_tinfo.restVolume = tetrahedronVolume;
        mu *= std::fabs(tetrahedronVolume);
        lambda *= std::fabs(tetrahedronVolume);
        for (int j = 0; j < 4; ++j) {
            const Coord c = cross(point[(j + 2) % 4] - point[(j + 1) % 4], point[(j + 3) % 4] - point[(j + 1) % 4]);
            if ((j % 2) == 0) my_tinfo.shapeVector[j] = c / (tetrahedronVolume * 6);
            else my_tinfo.shapeVector[j] = (-c) / (tetrahedronVolume * 6);
        }
        for (int j = 0; j < 4; ++j) {
            const R val = mu * dot(my_tinfo.shapeVector[j], my_tinfo.shapeVector[j]);
            for (int m = 0; m < 3; ++m)
                for (int n = m; n < 3; ++n) {
                    my_tinfo.linearDfDxDiag[j](m, n) = lambda * my_tinfo.shapeVector[j][n] * my_tinfo.shapeVector[j][m] + mu * my_tinfo.shapeVector[j][n] * my_tinfo.shapeVector[j][m];
                    if (m == n) my_tinfo.linearDfDxDiag[j](m, m) += R(val);
                    else my_tinfo.linearDfDxDiag[j](n, m) = my_tinfo.linearDfDxDiag[j](m, n);
                }
        }
        for (int j = 0; j < 6; ++j) {
            const int k = L[j][0], l = L[j][1];
            my_tinfo.restEdgeVector[j] = point[l] - point[k];
            const R val = mu * dot(my_tinfo.shapeVector[l], my_tinfo.shapeVector[k]);
            for (int m = 0; m < 3; ++m)
                for (int n = 0; n < 3; ++n) {
                    my_tinfo.linearDfDx[j](m, n) = lambda * my_tinfo.shapeVector[k][n] * my_tinfo.shapeVector[l][m] + mu * my_tinfo.shapeVector[l][n] * my_tinfo.shapeVector[k][m];
                    if (m == n) my_tinfo.linearDfDx[j](m, m) += R(val);
                }
        }
        if (method == FAST_QR) computeQRRotation(my_tinfo.restRotation, my_tinfo.restEdgeVector);
        else if (method == FAST_POLAR2) {
            Mat3<R> Transformation;
            Transformation.setRow(0, point[1] - point[0]); Transformation.setRow(1, point[2] - point[0]); Transformation.setRow(2, point[3] - point[0]);
            Decompose<R>::polarDecomposition(Transformation, my_tinfo.restRotation);
        }
    }
    // init :176-246 + updateTopologyInformation :249-270.  `givenEdges`: the topology's own edge list when it has one (TetrahedronSetTopologyContainer.cpp:163-205)
    void init(const std::vector<Coord>& restPosition, const std::vector<uint32_t>* givenEdges = nullptr) {
        const size_t T = nbTets();
        if (givenEdges && !givenEdges->empty()) {
            edges = *givenEdges;
            std::map<std::pair<uint32_t, uint32_t>, uint32_t> idx;
            for (size_t e = 0; e < edges.size() / 2; ++e) idx.emplace(std::minmax(edges[2 * e], edges[2 * e + 1]), uint32_t(e));
            edgesInTet.assign(6 * T, 0);
            for (size_t i = 0; i < T; ++i) for (int j = 0; j < 6; ++j) edgesInTet[6 * i + j] = idx.at(std::minmax(tets[4 * i + L[j][0]], tets[4 * i + L[j][1]]));
        } else MeshMatrixMass<R>::createEdgeSetArray(tets, edges, edgesInTet);
        edgeInfo.assign(edges.size() / 2, Mat3<R>());
        tetrahedronInfo.assign(T, TetrahedronRestInformation());
        for (size_t i = 0; i < T; ++i) createTetrahedronRestInformation(i, tetrahedronInfo[i], restPosition);
        for (size_t i = 0; i < T; ++i)
            for (int j = 0; j < 6; ++j)
                tetrahedronInfo[i].edgeOrientation[j] = (tets[4 * i + L[j][0]] == edges[2 * edgesInTet[6 * i + j]]) ? R(1) : R(-1);
        updateMatrix = true;
    }
    void addForce(VecDeriv<R>& f, const std::vector<Coord>& x) {   // .inl:296-399
        for (size_t i = 0; i < nbTets(); ++i) {
            TetrahedronRestInformation& tetraInfo = tetrahedronInfo[i];
            const uint32_t* tetra = &tets[4 * i];
            Coord tetraVertex[4], displ[6];
            for (int j = 0; j < 4; ++j) tetraVertex[j] = x[tetra[j]];
            for (int j = 0; j < 6; ++j) displ[j] = tetraVertex[L[j][1]] - tetraVertex[L[j][0]];
            Mat3<R> deformationGradient, S, Rm;
            if (method == FAST_POLAR) {
                Coord sv = tetraInfo.shapeVector[1];
                for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) deformationGradient(k, l) = displ[0][k] * sv[l];
                for (int j = 1; j < 3; ++j) {
                    sv = tetraInfo.shapeVector[j + 1];
                    for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) deformationGradient(k, l) += displ[j][k] * sv[l];
                }
                Decompose<R>::polarDecomposition(deformationGradient, Rm);
            } else if (method == FAST_QR) {
                computeQRRotation(S, displ);
                Rm = S.multTranspose(tetraInfo.restRotation);
            } else if (method == FAST_POLAR2) {
                S.setRow(0, displ[0]); S.setRow(1, displ[1]); S.setRow(2, displ[2]);
                Decompose<R>::polarDecomposition(S, Rm);
                Rm = Rm.transposed() * tetraInfo.restRotation;
            } else Rm.identity();
            tetraInfo.rotation = Rm.transposed();
            Coord force[4];
            for (int j = 0; j < 6; ++j) {
                displ[j] = tetraInfo.rotation * displ[j] - tetraInfo.restEdgeVector[j];
                force[L[j][1]] += tetraInfo.linearDfDx[j] * displ[j];
                force[L[j][0]] -= tetraInfo.linearDfDx[j].multTranspose(displ[j]);
            }
            for (int j = 0; j < 4; ++j) f[tetra[j]] += Rm * force[j];
        }
        updateMatrix = true;
    }
    void assembleEdgeMatrices() {   // .inl:414-450
        for (auto& m : edgeInfo) m = Mat3<R>();
        for (size_t i = 0; i < nbTets(); ++i) {
            const TetrahedronRestInformation& tetinfo = tetrahedronInfo[i];
            for (int j = 0; j < 6; ++j) {
                const uint32_t edgeID = edgesInTet[6 * i + j];
                const Mat3<R> tmp = tetinfo.linearDfDx[j] * tetinfo.rotation;
                const Mat3<R> add = tetinfo.edgeOrientation[j] == 1 ? tetinfo.rotation.multTranspose(tmp) : tmp.multTranspose(tetinfo.rotation);
                edgeInfo[edgeID] = edgeInfo[edgeID] + add;
            }
        }
    }
    void addDForce(VecDeriv<R>& df, const VecDeriv<R>& dx, SReal kFactorIncludingRayleigh) {   // .inl:402-470
        const R kFactor = R(kFactorIncludingRayleigh);
        if (updateMatrix) { updateMatrix = false; assembleEdgeMatrices(); }
        for (size_t i = 0; i < edges.size() / 2; ++i) {
            const uint32_t e0 = edges[2 * i], e1 = edges[2 * i + 1];
            const Coord deltax = (dx[e1] - dx[e0]) * kFactor;
            df[e1] += edgeInfo[i] * deltax;
            df[e0] -= edgeInfo[i].multTranspose(deltax);
        }
    }
};

template <class R> struct PlaneForceField {
    Vec3<R> planeNormal = Vec3<R>(0, 1, 0);
    R planeD = 0, stiffness = 500, damping = 5, maxForce = 0;
    bool bilateral = false;
    std::vector<uint32_t> contacts;      // m_contacts
    void setPlane(const Vec3<R>& normal, R d) { const R n = normal.norm(); planeNormal = normal / n; planeD = d / n; }
    void addForce(VecDeriv<R>& f, const std::vector<Vec3<R>>& p, const VecDeriv<R>& v) {
        contacts.clear();
        R limit = maxForce;
        limit *= limit;
        const R stiff = stiffness, damp = damping;
        const Vec3<R> planeN = planeNormal;
        for (size_t i = 0; i < p.size(); ++i) {
            const R d = dot(p[i], planeN) - planeD;
            if (bilateral || d < 0) {
                const R forceIntensity = -stiff * d;
                const R dampingIntensity = -damp * d;
                Vec3<R> force = planeN * forceIntensity - v[i] * dampingIntensity;
                const R amplitude = force.norm2();
                if (limit > 0 && amplitude > limit) force *= std::sqrt(limit / amplitude);
                f[i] += force;
                contacts.push_back(uint32_t(i));
            }
        }
    }
    void addDForce(VecDeriv<R>& df, const VecDeriv<R>& dx, double kFactorIncludingRayleighDamping) const {
        const R fact = (R)(-stiffness * kFactorIncludingRayleighDamping);
        for (uint32_t p : contacts) df[p] = df[p] + planeNormal * (fact * dot(dx[p], planeNormal));
    }
};

// FixedProjectiveConstraint::projectResponse
// Sofa/Component/Constraint/Projective/src/sofa/component/constraint/projective/FixedProjectiveConstraint.inl:183-206
template <class R> inline void projectResponse(VecDeriv<R>& res, const std::vector<uint32_t>& indices, bool fixAll) {
    if (fixAll) for (auto& x : res) x = Vec3<R>();
    else for (uint32_t i : indices) res[i] = Vec3<R>();
}

// ---------------------------------------------------------------------------
// One solver node: MechanicalObject + DiagonalMass + {Tetrahedron,Hexahedron}FEMForceField +
// FixedProjectiveConstraint under EulerImplicitSolver + CGLinearSolver<GraphScattered>.
// ---------------------------------------------------------------------------
template <class R> struct Scene {
    typedef Vec3<R> Coord;
    std::vector<Coord> x, x0;
    VecDeriv<R> v, f, dx;
    DiagonalMass<R> mass;
    MeshMatrixMass<R> meshMass; bool hasMeshMass = false;   // when set, the node's mass component is a MeshMatrixMass instead
    bool hasMass = true;
    double massRayleighMass = 0;      // Mass::rayleighMass Data of the mass component (default 0)
    TetFEM<R> tet; bool hasTet = false;
    HexaFEM<R> hex; bool hasHex = false;
    FastTetFEM<R> fast; bool hasFast = false;
    double ffRayleighStiffness = 0;   // BaseForceField::rayleighStiffness of the FEM component (default 0)
    bool massFirst = true;            // scene order of the two force fields (mass before FEM in every reference scene)
    PlaneForceField<R> plane; bool hasPlane = false;   // last force field of the node (as in the SofaCUDA benchmark scenes)
    double planeRayleighStiffness = 0;
    std::vector<uint32_t> fixed; bool fixAll = false;
    double gravity[3] = {0, -9.81, 0};
    // EulerImplicitSolver Data (EulerImplicitSolver.cpp:40-50)
    double dt = 0.01, rayleighStiffness = 0, rayleighMass = 0, vdamping = 0;
    bool firstOrder = false, trapezoidal = false;
    // CGLinearSolver Data (CGLinearSolver.inl:35-45)
    unsigned maxIter = 25; double tolerance = 1e-5, threshold = 1e-5; bool warmStart = false;
    unsigned timeStepCount = 0;
    // TEST KNOB (not a reference Data): accumulate the CG dot products in double instead of the reference's serial
    // Real accumulation (MechanicalObject.inl:2333-2356).  Used to separate "the dot product is summed in another
    // order / precision" (inherent to any parallel implementation) from every other source of difference.
    // dotReverse additionally walks the vector backwards: the same numbers summed in another order, to measure how
    // strongly the trajectory of the reference itself depends on the (unspecified) summation order of vDot.
    bool dotDouble = false, dotReverse = false;
    SReal sdot(const VecDeriv<R>& a, const VecDeriv<R>& b) const {
        if (!dotDouble && !dotReverse) return VOps<R>::dot(a, b);
        double r = 0.0;
        const size_t n = a.size();
        for (size_t k = 0; k < n; ++k) {
            const size_t i = dotReverse ? n - 1 - k : k;
            r += double(a[i][0]) * double(b[i][0]) + double(a[i][1]) * double(b[i][1]) + double(a[i][2]) * double(b[i][2]);
        }
        return r;
    }
    // outputs of the last solve
    unsigned lastIter = 0; int endCond = 0;  // 0 iterations, 1 tolerance, 2 threshold, 3 den==0, 4 b==0
    std::vector<double> graphError, graphDen;
    VecDeriv<R> lastB, lastSol, lastForce;
    // MechanicalParams factors of the system being applied
    double mFact = 0, bFact = 0, kFact = 0;

    void femAddForce(VecDeriv<R>& F) { if (hasTet) tet.addForce(F, x); if (hasHex) hex.addForce(F, x); if (hasFast) fast.addForce(F, x); }
    // threads > 1: the MultiThreading plugin's ParallelTetrahedronFEMForceField::addDForce
    // (applications/plugins/MultiThreading/src/MultiThreading/component/solidmechanics/fem/elastic/ParallelTetrahedronFEMForceField.inl:67-99):
    // element ranges over threads, thread-local df, merged under a mutex.  A TIMING variant for the CPU baseline
    // (its summation order differs from the sequential loop exactly as it does in the reference plugin).
    int threads = 1;
    void parallelTetAddDForce(VecDeriv<R>& df, const VecDeriv<R>& d, double kf) {
        df.resize(d.size());
        const size_t T = tet.nbTets();
        const R kFactor = R(kf);
        std::mutex mtx;
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) {
            pool.emplace_back([&, t]() {
                const size_t lo = T * t / threads, hi = T * (t + 1) / threads;
                VecDeriv<R> local(d.size());
                if (tet.method == SMALL) for (size_t i = lo; i < hi; ++i) tet.applyStiffnessSmall(local, d, i, kFactor);
                else for (size_t i = lo; i < hi; ++i) tet.applyStiffnessCorotational(local, d, i, kFactor);
                std::lock_guard<std::mutex> g(mtx);
                for (size_t i = 0; i < df.size(); ++i) df[i] += local[i];
            });
        }
        for (auto& th : pool) th.join();
    }
    // threads > 1: ParallelHexahedronFEMForceField::addDForce
    // (applications/plugins/MultiThreading/src/MultiThreading/component/solidmechanics/fem/elastic/ParallelHexahedronFEMForceField.inl:201-268):
    // pass 1, element ranges over threads: every hexahedron's 8 corner terms df_w = -(R^T F_w) kFactor into m_elementsDf; pass 2, vertex
    // ranges over threads: df[v] += m_elementsDf[h][corner of v in h] over the hexahedra around v (m_around, ascending hexahedron index,
    // built once as in the class's initStiffnessMatrices :147-172).  A TIMING variant for the CPU baseline.
    std::vector<std::vector<uint32_t>> hexAround, hexAroundCorner;
    std::vector<Coord> hexElementsDf;
    void parallelHexAddDForce(VecDeriv<R>& df, const VecDeriv<R>& d, double kf) {
        if (df.size() != d.size()) df.resize(d.size());
        const size_t H = hex.nbHexas(), N = d.size();
        const R kFactor = R(kf);
        if (hexAround.size() != N) {
            hexAround.assign(N, {}); hexAroundCorner.assign(N, {});
            for (size_t h = 0; h < H; ++h) for (int w = 0; w < 8; ++w) { hexAround[hex.hexas[8 * h + w]].push_back(uint32_t(h)); hexAroundCorner[hex.hexas[8 * h + w]].push_back(uint32_t(w)); }
        }
        hexElementsDf.resize(8 * H);
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t)
            pool.emplace_back([&, t]() {
                for (size_t i = H * t / threads; i < H * (t + 1) / threads; ++i) {
                    const uint32_t* elem = &hex.hexas[8 * i];
                    R X[24], F[24];
                    for (int w = 0; w < 8; ++w) { const Coord x_2 = hex.rotations[i] * d[elem[w]]; X[3 * w] = x_2[0]; X[3 * w + 1] = x_2[1]; X[3 * w + 2] = x_2[2]; }
                    HexaFEM<R>::computeForce(F, X, &hex.Ke[576 * i]);
                    for (int w = 0; w < 8; ++w) hexElementsDf[8 * i + w] = -hex.rotations[i].multTranspose(Coord(F[3 * w], F[3 * w + 1], F[3 * w + 2])) * kFactor;
                }
            });
        for (auto& th : pool) th.join();
        pool.clear();
        for (int t = 0; t < threads; ++t)
            pool.emplace_back([&, t]() {
                for (size_t v = N * t / threads; v < N * (t + 1) / threads; ++v)
                    for (size_t a = 0; a < hexAround[v].size(); ++a) df[v] += hexElementsDf[8 * size_t(hexAround[v][a]) + hexAroundCorner[v][a]];
            });
        for (auto& th : pool) th.join();
    }
    void femAddDForce(VecDeriv<R>& df, const VecDeriv<R>& d, double kf) {
        if (hasTet) { if (threads > 1) parallelTetAddDForce(df, d, kf); else tet.addDForce(df, d, kf); }
        if (hasHex) { if (threads > 1) parallelHexAddDForce(df, d, kf); else hex.addDForce(df, d, kf); }
        if (hasFast) fast.addDForce(df, d, kf);
    }
    // mop.computeForce: resetForce, accumulateForce (no external force), every force field's addForce in scene order
    // Sofa/framework/Simulation/Core/src/sofa/simulation/MappingGraphMechanicalOperations.cpp:41-93
    VecDeriv<R> externalForce;   // MechanicalObject Data `externalForce` (empty = none)
    void computeForce(VecDeriv<R>& F) {
        F.assign(x.size(), Coord());
        // MechanicalObject::accumulateForce, Sofa/Component/StateContainer/src/sofa/component/statecontainer/MechanicalObject.inl:1356-1375
        for (size_t i = 0; i < externalForce.size(); ++i)
            if (!(externalForce[i][0] == R(0) && externalForce[i][1] == R(0) && externalForce[i][2] == R(0))) F[i] += externalForce[i];
        auto massForce = [&]() { if (hasMeshMass) meshMass.addForce(F, gravity); else mass.addForce(F, gravity); };
        if (massFirst && hasMass) massForce();
        femAddForce(F);
        if (!massFirst && hasMass) massForce();
        if (hasPlane) plane.addForce(F, x, v);
    }
    // addMBKdx over the node's force fields: BaseForceField::addMBKdx (BaseForceField.cpp:38-47) and
    // Mass::addMBKdx (Sofa/framework/Core/src/sofa/core/behavior/Mass.inl:93-105); factors per MechanicalParams.h:62-64
    void addMBKdx(VecDeriv<R>& df, const VecDeriv<R>& d, double m, double b, double k) {
        auto massPart = [&]() {
            if (!hasMass) return;
            const double mf = m - b * massRayleighMass;
            if (mf != 0.0) { if (hasMeshMass) meshMass.addMDx(df, d, mf); else mass.addMDx(df, d, mf); }
        };
        auto femPart = [&]() {
            const double kf = k + b * ffRayleighStiffness;
            if (kf != 0.0 || b != 0.0) femAddDForce(df, d, kf);
        };
        if (massFirst) { massPart(); femPart(); } else { femPart(); massPart(); }
        if (hasPlane) {   // BaseForceField::addMBKdx again, for the plane
            const double kf = k + b * planeRayleighStiffness;
            if (kf != 0.0 || b != 0.0) plane.addDForce(df, d, kf);
        }
    }
    // GraphScatteredMatrix::apply  Sofa/Component/LinearSolver/Iterative/src/sofa/component/linearsolver/iterative/GraphScatteredTypes.cpp:33-46
    void applyA(VecDeriv<R>& res, const VecDeriv<R>& p) {
        res.assign(p.size(), Coord());
        addMBKdx(res, p, mFact, bFact, kFact);
        projectResponse(res, fixed, fixAll);
    }
    // CGLinearSolver<GraphScattered>::solve  CGLinearSolver.inl:73-315, cgstep_* CGLinearSolver.cpp:41-66
    void cgSolve(VecDeriv<R>& X, const VecDeriv<R>& b) {
        VecDeriv<R> p(b.size()), q(b.size()), r(b.size());
        SReal rho, rho_1 = 0, alpha, beta;
        if (warmStart) { applyA(r, X); VOps<R>::avf(r, b, -1.0); /* r = b + r*(-1): eq(b,r,-1) -> vOp(r,b,r,-1) */ }
        else { VOps<R>::clear(X); r = b; }
        const SReal normb = std::sqrt(sdot(b, b));
        graphError.clear(); graphError.push_back(1); graphDen.clear();
        unsigned nb_iter = 0; endCond = 0;
        if (normb != 0.0) {
            for (nb_iter = 1; nb_iter <= maxIter; nb_iter++) {
                rho = sdot(r, r);
                const SReal normr = std::sqrt(rho);
                const SReal err = normr / normb;
                graphError.push_back(err);
                if (err <= tolerance) {
                    if (nb_iter == 1 && timeStepCount == 0) { /* warning only */ }
                    else { endCond = 1; break; }
                }
                if (nb_iter == 1) p = r;
                else { beta = rho / rho_1; VOps<R>::avf(p, r, beta); }
                applyA(q, p);
                const SReal den = sdot(p, q);
                graphDen.push_back(den);
                if (den != 0.0) {
                    if (std::fabs(den) <= threshold) {
                        if (nb_iter == 1 && timeStepCount == 0) { /* warning only */ }
                        else { endCond = 2; break; }
                    }
                    alpha = rho / den;
                    // cgstep_alpha -> BaseMechanicalState::vMultiOp fallback (BaseMechanicalState.cpp:42-79):
                    // vOp(x,x,p,alpha); vOp(r,r,q,-alpha)
                    VOps<R>::vOp(&X, &X, &p, alpha);
                    VOps<R>::vOp(&r, &r, &q, -alpha);
                } else { endCond = 3; break; }
                rho_1 = rho;
            }
        } else endCond = 4;
        timeStepCount++;
        lastIter = nb_iter;
    }
    // EulerImplicitSolver::solve  Sofa/Component/ODESolver/Backward/src/sofa/component/odesolver/backward/EulerImplicitSolver.cpp:83-341
    void step() {
        const SReal h = dt;
        const SReal tr = trapezoidal ? 0.5 : 1.0;
        computeForce(f);
        lastForce = f;
        VecDeriv<R> b = f;
        if (!firstOrder) {
            // mop.addMBKv(b, M(-rM), B(0), K(h*tr + rK))  (dx := v)
            addMBKdx(b, v, -rayleighMass, 0.0, h * tr + rayleighStiffness);
            VOps<R>::teq(b, h);
        }
        projectResponse(b, fixed, fixAll);
        mFact = firstOrder ? 1 : 1 + tr * h * rayleighMass;
        bFact = firstOrder ? 0 : -tr * h;
        kFact = firstOrder ? -h * tr : -tr * h * (tr * h + rayleighStiffness);
        dx.resize(x.size());
        cgSolve(dx, b);
        lastB = b; lastSol = dx;
        if (firstOrder) { v = dx; for (size_t i = 0; i < x.size(); ++i) x[i] += v[i] * R(h); }  // newVel.eq(x); newPos.eq(pos,newVel,h) -> vOp(x,x,v,h)
        else VOps<R>::integrate(v, x, dx, h);
        if (vdamping != 0.0) VOps<R>::teq(v, std::exp(-h * vdamping));
    }
};

}  // namespace orc
