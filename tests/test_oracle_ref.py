"""Pin the oracle's restated sofa::type / helper::Decompose math bit-for-bit against the
reference's own object code (oracle/_ref/libsofa_ref.so, built from /root/reference by
oracle/build_ref.sh).  Skipped when oracle/_ref was never built (it ships prebuilt to the GPU box)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref not built (no reference tree)")

DT = [("f", np.float32), ("d", np.float64)]


def _mats(rng, n, dt, kind):
    for i in range(n):
        if kind == "near_rot":
            M = np.eye(3) + 0.3 * rng.standard_normal((3, 3))
        elif kind == "wild":
            M = rng.standard_normal((3, 3)) * 10.0 ** rng.integers(-3, 3)
        elif kind == "flat":
            M = rng.standard_normal((3, 3)); M[:, 2] = M[:, 0] * rng.standard_normal() + 1e-7 * rng.standard_normal(3)
        elif kind == "inverted":
            M = np.eye(3) + 0.3 * rng.standard_normal((3, 3)); M[:, 0] *= -1
        elif kind == "edge":
            u = rng.standard_normal(3); M = np.outer(rng.standard_normal(3), u)
        else:
            M = np.zeros((3, 3))
        yield np.ascontiguousarray(M, dt)


@pytest.mark.parametrize("sfx,dt", DT)
@pytest.mark.parametrize("kind", ["near_rot", "wild", "inverted"])
def test_polar_decomposition_bit_exact(sfx, dt, kind):
    L, R = O.lib(), O.ref_lib()
    rng = np.random.default_rng(1)
    for M in _mats(rng, 3000, dt, kind):
        q1 = np.empty((3, 3), dt); q2 = np.empty((3, 3), dt)
        d1 = getattr(L, "orc_polar_" + sfx)(O._ptr(M), O._ptr(q1))
        d2 = getattr(R, "ref_polar_" + sfx)(O._ptr(M), O._ptr(q2))
        assert d1 == d2 and q1.tobytes() == q2.tobytes(), M


@pytest.mark.parametrize("sfx,dt", DT)
@pytest.mark.parametrize("kind", ["near_rot", "wild", "flat", "inverted", "edge", "zero"])
def test_polar_stable_and_svd_bit_exact(sfx, dt, kind):
    L, R = O.lib(), O.ref_lib()
    rng = np.random.default_rng(2)
    for M in _mats(rng, 2000 if kind != "zero" else 1, dt, kind):
        q1 = np.empty((3, 3), dt); q2 = np.empty((3, 3), dt)
        a = getattr(L, "orc_polar_stable_" + sfx)(O._ptr(M), O._ptr(q1))
        b = getattr(R, "ref_polar_stable_" + sfx)(O._ptr(M), O._ptr(q2))
        assert a == b and q1.tobytes() == q2.tobytes(), M
        U1, V1, S1 = np.empty((3, 3), dt), np.empty((3, 3), dt), np.empty(3, dt)
        U2, V2, S2 = np.empty((3, 3), dt), np.empty((3, 3), dt), np.empty(3, dt)
        a = getattr(L, "orc_svd_stable_" + sfx)(O._ptr(M), O._ptr(U1), O._ptr(S1), O._ptr(V1))
        b = getattr(R, "ref_svd_stable_" + sfx)(O._ptr(M), O._ptr(U2), O._ptr(S2), O._ptr(V2))
        assert a == b and U1.tobytes() == U2.tobytes() and V1.tobytes() == V2.tobytes() and S1.tobytes() == S2.tobytes()


@pytest.mark.parametrize("sfx,dt", DT)
def test_mat_vec_helpers_bit_exact(sfx, dt):
    L, R = O.lib(), O.ref_lib()
    rng = np.random.default_rng(3)
    for _ in range(2000):
        A = rng.standard_normal((3, 3)).astype(dt); B = rng.standard_normal((3, 3)).astype(dt)
        v = rng.standard_normal(3).astype(dt)
        for name in ("mat3_mul", "mat3_mul_transposed"):
            c1 = np.empty((3, 3), dt); c2 = np.empty((3, 3), dt)
            getattr(L, f"orc_{name}_{sfx}")(O._ptr(A), O._ptr(B), O._ptr(c1))
            getattr(R, f"ref_{name}_{sfx}")(O._ptr(A), O._ptr(B), O._ptr(c2))
            assert c1.tobytes() == c2.tobytes()
        for name in ("mat3_vec", "mat3_tvec"):
            r1 = np.empty(3, dt); r2 = np.empty(3, dt)
            getattr(L, f"orc_{name}_{sfx}")(O._ptr(A), O._ptr(v), O._ptr(r1))
            getattr(R, f"ref_{name}_{sfx}")(O._ptr(A), O._ptr(v), O._ptr(r2))
            assert r1.tobytes() == r2.tobytes()
        i1 = np.zeros((3, 3), dt); i2 = np.zeros((3, 3), dt)
        ok1 = getattr(L, f"orc_mat3_invert_{sfx}")(O._ptr(A), O._ptr(i1))
        ok2 = getattr(R, f"ref_mat3_invert_{sfx}")(O._ptr(A), O._ptr(i2))
        assert ok1 == ok2 and i1.tobytes() == i2.tobytes()
        assert getattr(L, f"orc_mat3_det_{sfx}")(O._ptr(A)) == getattr(R, f"ref_mat3_det_{sfx}")(O._ptr(A))
        a, b, c, d = (rng.standard_normal(3).astype(dt) * 5 for _ in range(4))
        f1 = np.empty((3, 3), dt); f2 = np.empty((3, 3), dt)
        getattr(L, f"orc_frame_large_{sfx}")(O._ptr(a), O._ptr(b), O._ptr(c), O._ptr(f1))
        getattr(R, f"ref_frame_large_{sfx}")(O._ptr(a), O._ptr(b), O._ptr(c), O._ptr(f2))
        assert f1.tobytes() == f2.tobytes()
        assert getattr(L, f"orc_tet_volume_{sfx}")(*(O._ptr(p) for p in (a, b, c, d))) == \
               getattr(R, f"ref_tet_volume_{sfx}")(*(O._ptr(p) for p in (a, b, c, d)))
