"""CPU-side statement of the two identities the projection kernel (stark_b200/csrc/project.cu) relies on, checked with numpy
against the reference's definition of the projection (project_to_PD.cpp:13-82: symmetric eigen-decomposition, eigenvalues
below eps clamped to eps or mirrored, matrix rebuilt):

1. translation deflation: for an element whose energy depends on differences of its nodes' positions only, H has the three
   rigid translations in its null space; with Q = W (x) I3, W the Helmert basis of the nn nodes, the projection of H equals
   Q proj(Q^T H Q) Q^T + eps (I - Q Q^T), and I - Q Q^T = (1/nn) ones (x) I3;
2. the division-free Jacobi rotation: c = sqrt((1 + |a|/h)/2), s = sign(a) b / (2 h c) with a = aqq - app, b = 2 apq,
   h = hypot(a, b) annihilates apq and is the small-angle rotation of the classical formulas."""
import numpy as np


def project_reference(H, eps, mirror=False):
    w, V = np.linalg.eigh(0.5 * (H + H.T))
    w2 = np.where(w < eps, -w if mirror else eps, w)
    return (V * w2) @ V.T


def helmert(nn):
    W = np.zeros((nn, nn - 1))
    for j in range(nn - 1):
        s = 1.0 / np.sqrt((j + 1) * (j + 2))
        W[: j + 1, j] = s
        W[j + 1, j] = -(j + 1) * s
    return W


def test_helmert_basis_is_orthonormal_and_orthogonal_to_translations():
    for nn in (2, 3, 4, 5, 8):
        W = helmert(nn)
        assert np.abs(W.T @ W - np.eye(nn - 1)).max() < 1e-15
        assert np.abs(W.sum(axis=0)).max() < 1e-15
        assert np.abs(np.eye(nn) - W @ W.T - np.full((nn, nn), 1.0 / nn)).max() < 1e-15


def test_deflated_projection_equals_full_projection():
    rng = np.random.default_rng(3)
    eps = 1e-10
    for nn in (2, 3, 4):
        n = 3 * nn
        Q = np.kron(helmert(nn), np.eye(3))
        for trial in range(20):
            # translation-invariant symmetric indefinite matrix: H = D^T S D with D the difference operator to node 0
            D = np.kron(np.hstack([-np.ones((nn - 1, 1)), np.eye(nn - 1)]), np.eye(3))
            S = rng.normal(size=(n - 3, n - 3))
            S = S + S.T
            H = D.T @ S @ D
            assert np.abs(H @ np.kron(np.ones((nn, 1)), np.eye(3))).max() < 1e-12
            full = project_reference(H, eps)
            M = Q.T @ H @ Q
            defl = Q @ project_reference(M, eps) @ Q.T + eps * np.kron(np.full((nn, nn), 1.0 / nn), np.eye(3))
            assert np.abs(full - defl).max() <= 1e-10 * np.abs(H).max()
            # mirrored variant: the null eigenvalues (+-1e-16 |H| in the reference) are below eps and mirror to ~0
            full_m = project_reference(H, eps, mirror=True)
            defl_m = Q @ project_reference(M, eps, mirror=True) @ Q.T
            assert np.abs(full_m - defl_m).max() <= 1e-10 * np.abs(H).max()


def test_division_free_rotation_annihilates_and_is_the_small_angle():
    rng = np.random.default_rng(5)
    for _ in range(200):
        app, aqq, apq = rng.normal(size=3) * 10.0 ** rng.integers(-6, 7)
        if apq == 0.0:
            continue
        a, b = aqq - app, 2.0 * apq
        h = np.hypot(a, b)
        c = np.sqrt(0.5 + 0.5 * abs(a) / h)
        s = (1.0 if a >= 0 else -1.0) * 0.5 * b / (h * c)
        assert abs(c * c + s * s - 1.0) < 1e-14
        assert c >= np.sqrt(0.5) - 1e-15                                   # |theta| <= pi/4
        # classical: tau = a / b, t = sign(tau) / (|tau| + sqrt(1 + tau^2)), c = 1 / sqrt(1 + t^2), s = t c
        tau = a / b
        t = (1.0 if tau >= 0 else -1.0) / (abs(tau) + np.sqrt(1.0 + tau * tau))
        c0 = 1.0 / np.sqrt(1.0 + t * t)
        s0 = t * c0
        if a != 0.0:
            assert abs(c - c0) < 1e-12 and abs(s - s0) < 1e-12
        # J^T A J with J = [[c, s], [-s, c]] in the kernel's convention: new apq = (c^2 - s^2) apq + c s (app - aqq)
        new_apq = (c * c - s * s) * apq + c * s * (app - aqq)
        assert abs(new_apq) <= 1e-13 * max(abs(app), abs(aqq), abs(apq))
