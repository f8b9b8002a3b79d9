"""CPU test of the fill-reducing ordering behind sb_solve_llt (stark_b200/csrc/direct.cu: reverse Cuthill-McKee on the
block graph, dense rows last) through the C-ABI's host-only entry point sb_llt_order -- no GPU involved."""
import ctypes as C

import numpy as np

from stark_b200 import capi


def _grid_graph(nx, ny, nz, extra_dense=False):
    """27-point stencil on an nx x ny x nz node grid, numbered x-fastest (so the natural bandwidth is ~nx*ny), as BCSR
    block rows / first scalar columns; optionally one extra node coupled to every node of the top face (a rigid body)."""
    idx = np.arange(nx * ny * nz).reshape(nz, ny, nx)
    n = idx.size + (1 if extra_dense else 0)
    adj = [set([i]) for i in range(n)]
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                a = idx[max(0, -dz):nz - max(0, dz), max(0, -dy):ny - max(0, dy), max(0, -dx):nx - max(0, dx)]
                b = idx[max(0, dz):nz - max(0, -dz), max(0, dy):ny - max(0, -dy), max(0, dx):nx - max(0, -dx)]
                for i, j in zip(a.ravel(), b.ravel()):
                    adj[i].add(int(j))
    if extra_dense:
        for i in idx[-1].ravel():
            adj[n - 1].add(int(i)); adj[int(i)].add(n - 1)
    rows = np.zeros(n + 1, dtype=np.uint64)
    cols = []
    for i in range(n):
        c = sorted(adj[i])
        rows[i + 1] = rows[i] + len(c)
        cols += [3 * j for j in c]
    return n, rows, np.asarray(cols, dtype=np.int32)


def _order(n, rows, cols):
    lib = capi.load()
    perm = np.empty(n, dtype=np.int32)
    rc = lib.sb_llt_order(C.c_int(n), rows.ctypes.data_as(C.POINTER(C.c_ulonglong)), cols.ctypes.data_as(C.POINTER(C.c_int32)),
                          perm.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return perm


def _bandwidth(n, rows, cols, perm, skip=()):
    bw = 0
    for i in range(n):
        if i in skip:
            continue
        for j in cols[int(rows[i]):int(rows[i + 1])] // 3:
            if int(j) in skip:
                continue
            bw = max(bw, abs(int(perm[i]) - int(perm[j])))
    return bw


def test_order_is_a_permutation_and_narrows_a_bar():
    # a 5 x 5 x 40 bar numbered along its LONG axis last is already good; number it badly by shuffling, then order
    n, rows, cols = _grid_graph(5, 5, 40)
    rng = np.random.default_rng(0)
    shuffle = rng.permutation(n)              # new id of old node
    inv = np.argsort(shuffle)
    # relabel the graph
    r2 = np.zeros(n + 1, dtype=np.uint64)
    c2 = []
    for new in range(n):
        old = inv[new]
        cc = sorted(int(shuffle[j]) for j in cols[int(rows[old]):int(rows[old + 1])] // 3)
        r2[new + 1] = r2[new] + len(cc)
        c2 += [3 * j for j in cc]
    c2 = np.asarray(c2, dtype=np.int32)
    perm = _order(n, r2, c2)
    assert sorted(perm.tolist()) == list(range(n))
    assert _bandwidth(n, r2, c2, np.arange(n)) > 500          # the shuffled numbering has no band
    assert _bandwidth(n, r2, c2, perm) <= 2 * 25 + 12         # about two cross-sections


def test_dense_rows_are_ordered_last():
    n, rows, cols = _grid_graph(16, 16, 6, extra_dense=True)   # node n-1 touches the 256 nodes of the top face
    perm = _order(n, rows, cols)
    assert sorted(perm.tolist()) == list(range(n))
    assert perm[n - 1] == n - 1
    # the rest keeps a band of about two 16 x 6 cross-sections although the dense row couples a whole face
    assert _bandwidth(n, rows, cols, perm, skip={n - 1}) <= 2 * 96 + 32


def test_disconnected_components_and_empty():
    n, rows, cols = _grid_graph(3, 3, 3)
    # two copies, no coupling
    rows2 = np.concatenate([rows, rows[1:] + rows[-1]]).astype(np.uint64)
    cols2 = np.concatenate([cols, cols + 3 * n]).astype(np.int32)
    perm = _order(2 * n, rows2, cols2)
    assert sorted(perm.tolist()) == list(range(2 * n))
    assert _order(0, np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=np.int32)).size == 0


def test_cholesky_fill_stays_inside_the_tile_row_envelope():
    """The factor of sb_solve_llt is stored as a row envelope of 64 x 64 tiles (per tile row the run F(I) .. I, F from the
    pattern of the permuted matrix): on a fixture's real pattern -- tet block under hinged rigid boxes, i.e. with dense rigid-body
    rows -- the dense numpy Cholesky factor of the permuted matrix has no entry outside that envelope, and the envelope is much
    smaller than the dense lower triangle."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from golden_util import Golden
    NB = 64
    for name, max_fraction in (("tetbar_n2", 0.9), ("tetchain_n3", 0.95), ("tetdrop_n5", 0.9)):
        g = Golden(name)
        rows, cols, vals = g["bcsr_rows"].astype(np.uint64), g["bcsr_cols"].astype(np.int32), g["bcsr_vals"].astype(np.float64)
        nbr = len(rows) - 1
        perm = _order(nbr, rows, cols)
        n = 3 * nbr
        A = np.zeros((n, n))
        blocks = vals.reshape(-1, 3, 3).transpose(0, 2, 1)
        for i in range(nbr):
            for j in range(int(rows[i]), int(rows[i + 1])):
                c = int(cols[j]) // 3
                A[3 * perm[i]:3 * perm[i] + 3, 3 * perm[c]:3 * perm[c] + 3] = blocks[j]
        A = np.tril(A) + np.tril(A, -1).T
        pattern = A != 0.0
        A = A + (np.abs(A).sum(axis=1).max() + 1.0) * np.eye(n)           # diagonally dominant: positive definite, same pattern
        L = np.linalg.cholesky(A)
        nt = (n + NB - 1) // NB
        F = np.arange(nt)
        ii, jj = np.nonzero(np.tril(pattern))
        np.minimum.at(F, ii // NB, jj // NB)
        li, lj = np.nonzero(np.abs(L) > 1e-13 * np.abs(L).max())
        assert np.all(lj // NB >= F[li // NB]), name
        tiles = int(np.sum(np.arange(nt) - F + 1))
        assert tiles <= max_fraction * nt * (nt + 1) // 2 or nt <= 3, (name, tiles, nt)
