// 3x3 / 3-vector arithmetic shared by the host-side init and the device kernels.
//
// Every routine performs the reference's floating-point operations in the reference's order
// (sofa::type::Vec / Mat and sofa::helper::Decompose), so that with contraction disabled
// (nvcc -fmad=false, host -ffp-contract=off) results are bit-identical to a stock CPU build of SOFA:
//   Vec  : Sofa/framework/Type/src/sofa/type/Vec.h:406-413 (dot), :483-499 (norm), :545-562 (normalize), :774-780 (cross)
//   Mat  : Sofa/framework/Type/src/sofa/type/Mat.h:577-611 (M*v, M^T*v), :624-636, :988-997 (det), :1055-1080 (norms),
//          :1171-1193 (inverse), :1443-1540 (3x3 products)
//   Decompose : Sofa/framework/Helper/src/sofa/helper/decompose.inl:672-723 (polar), :755-764, :1489-1608, :1662-1829 (stable SVD)
#pragma once
#include <cfloat>
#include <cmath>

#include "common.cuh"

namespace sb {

template <class R> struct V3 { R x, y, z; };
template <class R> struct M3 { R m[3][3]; };

template <class R> HD V3<R> mk3(R x, R y, R z) { V3<R> v; v.x = x; v.y = y; v.z = z; return v; }
template <class R> HD V3<R> operator-(const V3<R>& a, const V3<R>& b) { return mk3<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class R> HD V3<R> operator+(const V3<R>& a, const V3<R>& b) { return mk3<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class R> HD V3<R> operator*(const V3<R>& a, R f) { return mk3<R>(a.x * f, a.y * f, a.z * f); }
template <class R> HD R dot3(const V3<R>& a, const V3<R>& b) { R r = a.x * b.x; r += a.y * b.y; r += a.z * b.z; return r; }
template <class R> HD V3<R> cross3(const V3<R>& a, const V3<R>& b) {
    return mk3<R>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
HD float sqrt_r(float a) { return sqrtf(a); }  // == float(sqrt(double(a))): sqrt is correctly rounded either way
HD double sqrt_r(double a) { return sqrt(a); }
template <class R> HD R eps_r();
template <> HD float eps_r<float>() { return FLT_EPSILON; }
template <> HD double eps_r<double>() { return DBL_EPSILON; }
template <class R> HD R abs_r(R r) { return (r >= 0) ? r : -r; }
template <class R> HD void normalize3(V3<R>& v) {
    R n2 = v.x * v.x; n2 += v.y * v.y; n2 += v.z * v.z;
    const R n = sqrt_r(n2);
    if (n > eps_r<R>()) { v.x /= n; v.y /= n; v.z /= n; }
}
template <class R> HD V3<R> row(const M3<R>& a, int i) { return mk3<R>(a.m[i][0], a.m[i][1], a.m[i][2]); }
template <class R> HD void set_row(M3<R>& a, int i, const V3<R>& v) { a.m[i][0] = v.x; a.m[i][1] = v.y; a.m[i][2] = v.z; }
template <class R> HD M3<R> transpose(const M3<R>& a) {
    M3<R> t;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) t.m[i][j] = a.m[j][i];
    return t;
}
// M * v
template <class R> HD V3<R> mul(const M3<R>& a, const V3<R>& v) {
    V3<R> r;
    r.x = a.m[0][0] * v.x; r.x += a.m[0][1] * v.y; r.x += a.m[0][2] * v.z;
    r.y = a.m[1][0] * v.x; r.y += a.m[1][1] * v.y; r.y += a.m[1][2] * v.z;
    r.z = a.m[2][0] * v.x; r.z += a.m[2][1] * v.y; r.z += a.m[2][2] * v.z;
    return r;
}
// M^T * v
template <class R> HD V3<R> mul_t(const M3<R>& a, const V3<R>& v) {
    V3<R> r;
    r.x = a.m[0][0] * v.x; r.x += a.m[1][0] * v.y; r.x += a.m[2][0] * v.z;
    r.y = a.m[0][1] * v.x; r.y += a.m[1][1] * v.y; r.y += a.m[2][1] * v.z;
    r.z = a.m[0][2] * v.x; r.z += a.m[1][2] * v.y; r.z += a.m[2][2] * v.z;
    return r;
}
// A * B
template <class R> HD M3<R> mul(const M3<R>& a, const M3<R>& b) {
    M3<R> r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
// A * B^T
template <class R> HD M3<R> mul_abt(const M3<R>& a, const M3<R>& b) {
    M3<R> r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            R s = a.m[i][0] * b.m[j][0]; s += a.m[i][1] * b.m[j][1]; s += a.m[i][2] * b.m[j][2];
            r.m[i][j] = s;
        }
    return r;
}
// A^T * B
template <class R> HD M3<R> mul_atb(const M3<R>& a, const M3<R>& b) {
    M3<R> r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[0][i] * b.m[0][j] + a.m[1][i] * b.m[1][j] + a.m[2][i] * b.m[2][j];
    return r;
}
template <class R> HD R det3(const M3<R>& m) {
    return m.m[0][0] * m.m[1][1] * m.m[2][2] + m.m[1][0] * m.m[2][1] * m.m[0][2] + m.m[2][0] * m.m[0][1] * m.m[1][2]
         - m.m[0][0] * m.m[2][1] * m.m[1][2] - m.m[1][0] * m.m[0][1] * m.m[2][2] - m.m[2][0] * m.m[1][1] * m.m[0][2];
}
template <class R> HD bool invert3(M3<R>& d, const M3<R>& f) {
    const R det = det3(f);
    if (abs_r(det) <= eps_r<R>()) return false;
    d.m[0][0] = (f.m[1][1] * f.m[2][2] - f.m[2][1] * f.m[1][2]) / det;
    d.m[1][0] = (f.m[1][2] * f.m[2][0] - f.m[2][2] * f.m[1][0]) / det;
    d.m[2][0] = (f.m[1][0] * f.m[2][1] - f.m[2][0] * f.m[1][1]) / det;
    d.m[0][1] = (f.m[2][1] * f.m[0][2] - f.m[0][1] * f.m[2][2]) / det;
    d.m[1][1] = (f.m[2][2] * f.m[0][0] - f.m[0][2] * f.m[2][0]) / det;
    d.m[2][1] = (f.m[2][0] * f.m[0][1] - f.m[0][0] * f.m[2][1]) / det;
    d.m[0][2] = (f.m[0][1] * f.m[1][2] - f.m[1][1] * f.m[0][2]) / det;
    d.m[1][2] = (f.m[0][2] * f.m[1][0] - f.m[1][2] * f.m[0][0]) / det;
    d.m[2][2] = (f.m[0][0] * f.m[1][1] - f.m[1][0] * f.m[0][1]) / det;
    return true;
}
template <class R> HD R one_norm(const M3<R>& a) {
    R n = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { const R s = abs_r(a.m[0][i]) + abs_r(a.m[1][i]) + abs_r(a.m[2][i]); if (s > n) n = s; }
    return n;
}
template <class R> HD R inf_norm(const M3<R>& a) {
    R n = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { const R s = abs_r(a.m[i][0]) + abs_r(a.m[i][1]) + abs_r(a.m[i][2]); if (s > n) n = s; }
    return n;
}
template <class R> HD R zero_tol();
template <> HD float zero_tol<float>() { return 1e-6f; }   // decompose.h:377-381
template <> HD double zero_tol<double>() { return 1e-8; }  // decompose.h:383-387
// gamma of the scaled Newton step.  The reference calls the C `sqrt`/`fabs` (double) on Real operands, so the
// float build evaluates this expression in double and rounds once (checked against the reference's object code).
HD float polar_gamma(float ratio, float det) { return float(sqrt(sqrt(double(ratio)) / fabs(double(det)))); }
HD double polar_gamma(double ratio, double det) { return sqrt(sqrt(ratio) / fabs(det)); }

// Decompose<R>::polarDecomposition(M, Q): Q = rotation factor of M.
template <class R> HD void polar_decomposition(const M3<R>& M, M3<R>& Q) {
    M3<R> Mk = transpose(M);
    R M_one = one_norm(Mk), M_inf = inf_norm(Mk), E_one;
    do {
        M3<R> adj;
        set_row(adj, 0, cross3(row(Mk, 1), row(Mk, 2)));
        set_row(adj, 1, cross3(row(Mk, 2), row(Mk, 0)));
        set_row(adj, 2, cross3(row(Mk, 0), row(Mk, 1)));
        const R det = Mk.m[0][0] * adj.m[0][0] + Mk.m[0][1] * adj.m[0][1] + Mk.m[0][2] * adj.m[0][2];
        if (det == R(0)) break;
        const R adj_one = one_norm(adj), adj_inf = inf_norm(adj);
        const R gamma = polar_gamma((adj_one * adj_inf) / (M_one * M_inf), det);
        const R g1 = gamma * R(0.5);
        const R g2 = R(0.5) / (gamma * det);
        M3<R> Ek = Mk;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                Mk.m[i][j] = Mk.m[i][j] * g1 + adj.m[i][j] * g2;
                Ek.m[i][j] -= Mk.m[i][j];
            }
        E_one = one_norm(Ek);
        M_one = one_norm(Mk);
        M_inf = inf_norm(Mk);
    } while (E_one > M_one * zero_tol<R>());
    Q = transpose(Mk);
}

// Decompose<R>::QLAlgorithm<3> (decompose.inl:1489-1557)
template <class R> HD void ql_algorithm3(R diag[3], R sub[3], M3<R>& V) {
    for (int i0 = 0; i0 < 3; ++i0) {
        int i1;
        for (i1 = 0; i1 < 32; ++i1) {
            int i2;
            for (i2 = i0; i2 <= 1; ++i2) {
                const R t = abs_r(diag[i2]) + abs_r(diag[i2 + 1]);
                if (abs_r(sub[i2]) + t == t) break;
            }
            if (i2 == i0) break;
            R fG = (diag[i0 + 1] - diag[i0]) / (R(2.0) * sub[i0]);
            R fR = sqrt_r(fG * fG + R(1.0));
            if (fG < R(0.0)) fG = diag[i2] - diag[i0] + sub[i0] / (fG - fR);
            else             fG = diag[i2] - diag[i0] + sub[i0] / (fG + fR);
            R fSin = 1.0, fCos = 1.0, fP = 0.0;
            for (int i3 = i2 - 1; i3 >= i0; --i3) {
                R fF = fSin * sub[i3];
                const R fB = fCos * sub[i3];
                if (abs_r(fF) >= abs_r(fG)) {
                    fCos = fG / fF;
                    fR = sqrt_r(fCos * fCos + R(1.0));
                    sub[i3 + 1] = fF * fR;
                    fSin = R(1.0) / fR;
                    fCos *= fSin;
                } else {
                    fSin = fF / fG;
                    fR = sqrt_r(fSin * fSin + R(1.0));
                    sub[i3 + 1] = fG * fR;
                    fCos = R(1.0) / fR;
                    fSin *= fCos;
                }
                fG = diag[i3 + 1] - fP;
                fR = (diag[i3] - fG) * fSin + R(2.0) * fB * fCos;
                fP = fSin * fR;
                diag[i3 + 1] = fG + fP;
                fG = fCos * fR - fB;
                for (int i4 = 0; i4 < 3; ++i4) {
                    fF = V.m[i4][i3 + 1];
                    V.m[i4][i3 + 1] = fSin * V.m[i4][i3] + fCos * fF;
                    V.m[i4][i3] = fCos * V.m[i4][i3] - fSin * fF;
                }
            }
            diag[i0] -= fP;
            sub[i0] = fG;
            sub[i2] = R(0.0);
        }
        if (i1 == 32) return;
    }
}
// Decompose<R>::eigenDecomposition_iterative for 3x3 (decompose.inl:1561-1608)
template <class R> HD void eigen_iterative3(const M3<R>& M, M3<R>& V, R diag[3]) {
    R sub[3];
    const R m00 = M.m[0][0], m11 = M.m[1][1], m12 = M.m[1][2], m22 = M.m[2][2];
    R m01 = M.m[0][1], m02 = M.m[0][2];
    diag[0] = m00;
    sub[2] = R(0.0);
    if (m02 != R(0.0)) {
        const R len = sqrt_r(m01 * m01 + m02 * m02);
        const R inv = R(1.0) / len;
        m01 *= inv;
        m02 *= inv;
        const R q = R(2.0) * m01 * m12 + m02 * (m22 - m11);
        diag[1] = m11 + m02 * q;
        diag[2] = m22 - m02 * q;
        sub[0] = len;
        sub[1] = m12 - m01 * q;
        V.m[0][0] = R(1.0); V.m[0][1] = R(0.0); V.m[0][2] = R(0.0);
        V.m[1][0] = R(0.0); V.m[1][1] = m01;    V.m[1][2] = m02;
        V.m[2][0] = R(0.0); V.m[2][1] = m02;    V.m[2][2] = -m01;
    } else {
        diag[1] = m11;
        diag[2] = m22;
        sub[0] = m01;
        sub[1] = m12;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) V.m[i][j] = (i == j) ? R(1) : R(0);
    }
    ql_algorithm3(diag, sub, V);
}
// Decompose<R>::polarDecomposition_stable(M, Q) = U V^T from SVD_stable (decompose.inl:755-764, 1662-1829)
template <class R> HD void polar_decomposition_stable(const M3<R>& F, M3<R>& Q) {
    M3<R> V, U;
    R S[3], S_1[3];
    const M3<R> FtF = mul_atb(F, F);
    eigen_iterative3(FtF, V, S);
    if (det3(V) < R(0)) for (int i = 0; i < 3; ++i) V.m[i][0] = -V.m[i][0];
    int degenerated = 0;
    for (int i = 0; i < 3; ++i) {
        if (S[i] < zero_tol<R>()) { degenerated++; S[i] = R(0); S_1[i] = R(1); }
        else { S[i] = sqrt_r(S[i]); S_1[i] = R(1.) / S[i]; }
    }
    int o0, o1, o2;  // Sorder
    if (S[0] < S[1]) {
        if (S[0] < S[2]) { o0 = 0; if (S[1] < S[2]) { o1 = 1; o2 = 2; } else { o1 = 2; o2 = 1; } }
        else { o0 = 2; o1 = 0; o2 = 1; }
    } else {
        if (S[1] < S[2]) { o0 = 1; if (S[0] < S[2]) { o1 = 0; o2 = 2; } else { o1 = 2; o2 = 0; } }
        else { o0 = 2; o1 = 1; o2 = 0; }
    }
    if (degenerated == 3) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) U.m[i][j] = (i == j) ? R(1) : R(0);
    } else {
        M3<R> VS;  // V.multDiagonal(S_1)
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) VS.m[i][j] = V.m[i][j] * S_1[j];
        U = mul(F, VS);
        if (degenerated == 1) {
            const V3<R> c = cross3(mk3<R>(U.m[0][o1], U.m[1][o1], U.m[2][o1]), mk3<R>(U.m[0][o2], U.m[1][o2], U.m[2][o2]));
            U.m[0][o0] = c.x; U.m[1][o0] = c.y; U.m[2][o0] = c.z;
        } else if (degenerated == 2) {
            const V3<R> edge2 = mk3<R>(U.m[0][o2], U.m[1][o0], U.m[2][o0]);  // (sic) decompose.inl:1767
            V3<R> edge0, edge1;
            const R a0 = abs_r(edge2.x), a1 = abs_r(edge2.y), a2 = abs_r(edge2.z);
            if (a0 > a1) { if (a0 > a2) edge0 = mk3<R>(0, 1, 0); else edge0 = mk3<R>(1, 0, 0); }
            else         { if (a1 > a2) edge0 = mk3<R>(0, 0, 1); else edge0 = mk3<R>(1, 0, 0); }
            edge1 = cross3(edge2, edge0);
            normalize3(edge1);
            edge0 = cross3(edge1, edge2);
            U.m[0][o0] = edge0.x; U.m[1][o0] = edge0.y; U.m[2][o0] = edge0.z;
            U.m[0][o1] = edge1.x; U.m[1][o1] = edge1.y; U.m[2][o1] = edge1.z;
        }
    }
    if (det3(U) < R(0)) { U.m[0][o0] *= R(-1); U.m[1][o0] *= R(-1); U.m[2][o0] *= R(-1); }
    Q = mul_abt(U, V);
}

}  // namespace sb
