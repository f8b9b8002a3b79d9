"""CPU-side statement of the recurrence the persistent PCG kernel uses (stark_b200/csrc/pcg.cu): the Chronopoulos-Gear form
    p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = M^-1 r ; w = A u
    gamma' = r.u ; delta = w.u ; beta = gamma'/gamma ; alpha = gamma' / (delta - beta gamma'/alpha)
against the textbook preconditioned CG the reference runs (bsm/solve_pcg.h:83-232): same iterates in exact arithmetic, and
p^T A p -- the reference's indefiniteness test -- equals the denominator delta - beta gamma'/alpha."""
import numpy as np


def block_jacobi_inverse(A):
    n = A.shape[0]
    Minv = np.zeros_like(A)
    for b in range(0, n, 3):
        Minv[b:b + 3, b:b + 3] = np.linalg.inv(A[b:b + 3, b:b + 3])
    return Minv


def textbook_pcg(A, Minv, b, iters):
    x = np.zeros_like(b); r = b.copy(); z = Minv @ r; p = z.copy(); rz = r @ z
    xs, pAps = [], []
    for _ in range(iters):
        Ap = A @ p
        pAp = p @ Ap
        pAps.append(pAp)
        alpha = rz / pAp
        x = x + alpha * p
        r = r - alpha * Ap
        z = Minv @ r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
        xs.append(x.copy())
    return xs, pAps


def chronopoulos_gear_pcg(A, Minv, b, iters):
    x = np.zeros_like(b); r = b.copy(); u = Minv @ r
    p = np.zeros_like(b); s = np.zeros_like(b)
    gamma = r @ u
    w = A @ u
    delta = w @ u
    alpha, beta = gamma / delta, 0.0
    xs, pAps = [], [delta]
    for it in range(iters):
        p = u + beta * p
        s = w + beta * s
        x = x + alpha * p
        r = r - alpha * s
        u = Minv @ r
        gamma_new = r @ u
        w = A @ u
        delta = w @ u
        beta = gamma_new / gamma
        pAp = delta - beta * gamma_new / alpha
        gamma = gamma_new
        alpha = gamma / pAp
        xs.append(x.copy())
        pAps.append(pAp)
    return xs, pAps[:-1]


def test_same_iterates_and_same_curvature_test():
    rng = np.random.default_rng(11)
    n = 60
    B = rng.normal(size=(n, n))
    A = B @ B.T + 0.5 * np.eye(n)
    Minv = block_jacobi_inverse(A)
    b = rng.normal(size=n)
    xs_t, pAp_t = textbook_pcg(A, Minv, b, 12)
    xs_c, pAp_c = chronopoulos_gear_pcg(A, Minv, b, 12)
    for xt, xc in zip(xs_t, xs_c):
        assert np.abs(xt - xc).max() <= 1e-9 * np.abs(xt).max()
    for a, c in zip(pAp_t, pAp_c):
        assert abs(a - c) <= 1e-8 * abs(a)


def test_indefinite_matrix_is_detected_at_the_same_iteration():
    rng = np.random.default_rng(12)
    n = 30
    B = rng.normal(size=(n, n))
    A = B + B.T                      # indefinite
    A[np.arange(n), np.arange(n)] += 3.0
    Minv = block_jacobi_inverse(A)
    b = rng.normal(size=n)
    _, pAp_t = textbook_pcg(A, Minv, b, 10)
    _, pAp_c = chronopoulos_gear_pcg(A, Minv, b, 10)
    first_t = next((i for i, v in enumerate(pAp_t) if v <= 0.0), None)
    first_c = next((i for i, v in enumerate(pAp_c) if v <= 0.0), None)
    assert first_t is not None and first_t == first_c
