"""The single-reduction CG recurrence of the fused persistent kernel (sofa_b200/csrc/cg_fused.cuh), restated in numpy around the
oracle's A*p and compared with the oracle's CGLinearSolver restatement (CGLinearSolver.inl:130-272):

    per iteration ONE reduction of r.r, p.q, r.q, q.q;  alpha = (r.r)/(p.q) with the MEASURED r.r;
    rho' predicted = r.r + 2 malpha r.q + malpha^2 q.q  (exact identity for r' = r + q malpha), used for beta and the tolerance test only.

Checked here without a GPU: same iteration counts as the reference loop, and in Vec3f a solution at least as close to the double-dot
oracle as the reference's own serial-float dot products are (the yardstick the GPU tests use)."""
import numpy as np
import pytest

import gpu_common as G


def fused_cg(A, b, dtype, max_iter, tol, thr, tsc=1):
    R = np.dtype(dtype).type
    d = np.float64
    x = np.zeros_like(b); r = b.copy(); p = None
    normb = np.sqrt(np.sum(b.astype(d) ** 2))
    errs = [np.sqrt(np.sum(r.astype(d) ** 2)) / normb]
    if errs[-1] <= tol and tsc != 0:
        return x, 0, errs
    it, beta = 1, R(0)
    while True:
        p = r.copy() if it == 1 else (p * beta + r).astype(dtype)
        q = A(p)
        rr = np.sum(r.astype(d) ** 2); pq = np.sum(p.astype(d) * q.astype(d)); rq = np.sum(r.astype(d) * q.astype(d)); qq = np.sum(q.astype(d) ** 2)
        errs[-1] = np.sqrt(rr) / normb            # the measured value replaces the prediction
        if pq == 0 or (abs(pq) <= thr and not (it == 1 and tsc == 0)):
            return x, it, errs
        alpha_d = rr / pq
        alpha, malpha = R(alpha_d), R(-alpha_d)
        ma = float(malpha)
        rho_new = max(rr + 2 * ma * rq + ma * ma * qq, 0.0)
        x = (x + p * alpha).astype(dtype); r = (r + q * malpha).astype(dtype)
        it += 1
        if it > max_iter:
            return x, it, errs
        errs.append(np.sqrt(rho_new) / normb)
        if errs[-1] <= tol and not (it == 1 and tsc == 0):
            return x, it, errs
        beta = R(rho_new / rr)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("cfg,iters,tol", [("C1", 25, 1e-9), ("C1", 100, 1e-5), ("GRID_TEST", 20, 1e-5), ("C2_SMALL", 60, 1e-4)])
def test_single_reduction_recurrence_matches_the_reference_loop(dtype, cfg, iters, tol):
    s = G.oracle_scene(cfg, dtype)
    s.set_params(iterations=iters, tolerance=tol, threshold=1e-12)
    for _ in range(2):
        s.step()                                   # a deformed state: the cached rotations are not the identity
    c = G.CONFIGS[cfg]
    h, rK, rM = c["dt"], c["rK"], c["rM"]
    b = (s.compute_force() * h).astype(dtype); b[G.mesh(cfg)[4]] = 0
    m, bf, k = 1 + h * rM, -h, -h * (h + rK)
    s.set_dot_double(True)
    x_dd, it_dd = s.cg(b, m, bf, k); g_dd = s.graph("Error")[1:]
    s.set_dot_double(False)
    x_ref, it_ref = s.cg(b, m, bf, k)
    x_f, it_f, errs = fused_cg(lambda p: s.apply(p, m, bf, k), b, dtype, iters, tol, 1e-12)
    assert abs(it_f - it_ref) <= 1 and abs(it_f - it_dd) <= 1
    n = min(len(errs), len(g_dd), 26)
    assert np.allclose(errs[:n], g_dd[:n], rtol=1e-12 if dtype == np.float64 else 5e-6)
    if it_f == it_dd:
        yard = max(G.rel_err(x_ref, x_dd), 1e-15)          # the reference's own sensitivity to the order of its dot products
        assert G.rel_err(x_f, x_dd) <= max(2 * yard, 1e-13 if dtype == np.float64 else 2e-6), (G.rel_err(x_f, x_dd), yard)
