"""N > 1 path of the benchmark plumbing on CPU: two processes over gloo (the replicas themselves need GPUs; what is shared
between ranks -- barrier, MAX of the timings, SUM of the work -- is exercised here)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from stark_b200 import dist as sbdist
    sbdist.barrier()
    # rank r: 10 + 5 r ms for 46 + r Newton iterations, 20 ms for 100 evaluations
    t, w = sbdist.aggregate([10.0 + 5.0 * rank, 20.0], [46.0 + rank, 100.0])
    v = sbdist.throughput([10.0 + 5.0 * rank, 20.0], [46.0 + rank, 100.0])
    out[rank] = (t, w, v)
    sbdist.barrier()
    dist.destroy_process_group()


def test_two_rank_aggregation_over_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert set(out.keys()) == {0, 1}
    for r in (0, 1):
        t, w, v = out[r]
        assert t == [15.0, 20.0]              # MAX over ranks
        assert w == [93.0, 200.0]             # SUM over ranks
        assert abs(v[0] - 93.0 / 0.015) < 1e-9 and abs(v[1] - 200.0 / 0.020) < 1e-9
    assert out[0] == out[1]


def test_single_process_is_identity():
    from stark_b200 import dist as sbdist
    assert sbdist.world() == 1
    t, w = sbdist.aggregate([3.0], [7.0])
    assert t == [3.0] and w == [7.0]
    assert sbdist.throughput([2.0], [10.0]) == [5000.0]


# ---- the distributed linear solve's host logic: row partition + halo send lists (sb_dist_plan), two ranks over gloo ----

def _band_matrix(nbr, half_band, seed, dense_row=True):
    """Symmetric positive definite 3x3-block matrix with a band pattern (+ one block row coupled to many others: a rigid body)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    pairs = set()
    for i in range(nbr):
        for j in range(max(0, i - half_band), min(nbr, i + half_band + 1)):
            if i == j or rng.random() < 0.6:
                pairs.add((i, j)); pairs.add((j, i))
    if dense_row:
        for j in range(0, nbr, 3):
            pairs.add((nbr - 1, j)); pairs.add((j, nbr - 1))
    rows = np.zeros(nbr + 1, dtype=np.uint64)
    cols, vals = [], []
    blocks = {}
    for (i, j) in pairs:
        if (j, i) in blocks:
            blocks[(i, j)] = blocks[(j, i)].T
        else:
            b = rng.standard_normal((3, 3)) * 0.1
            blocks[(i, j)] = (b + b.T) * 0.5 if i == j else b
    for i in range(nbr):
        cs = sorted(j for (a, j) in pairs if a == i)
        rows[i + 1] = rows[i] + len(cs)
        for j in cs:
            b = blocks[(i, j)] + (np.eye(3) * (2.0 + 0.3 * len(cs)) if i == j else 0.0)
            cols.append(3 * j); vals.append(b)
    return rows, np.asarray(cols, dtype=np.int32), np.asarray(vals)


def _spmv_rows(rows, cols, vals, u, lo, hi):
    import numpy as np
    y = np.zeros(3 * (hi - lo))
    for i in range(lo, hi):
        for j in range(int(rows[i]), int(rows[i + 1])):
            y[3 * (i - lo):3 * (i - lo) + 3] += vals[j] @ u[cols[j]:cols[j] + 3]
    return y


def _solver_worker(rank, world_size, port, out):
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from stark_b200 import dist as sbdist
    nbr, grid = 240, 5
    rows, cols, vals = _band_matrix(nbr, 6, seed=3)          # replicated: every rank builds the same matrix
    bounds, mask = sbdist.plan(rows, cols, world_size, grid, rank)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    # every rank owns a copy of u in which only ITS rows are current; the owners push the rows the masks name (here: gloo
    # send / recv of (row, value) lists; on the GPUs: NVLink stores inside the kernel)
    u_true = np.random.default_rng(7).standard_normal(3 * nbr)
    u_mine = np.full(3 * nbr, np.nan)
    u_mine[3 * lo:3 * hi] = u_true[3 * lo:3 * hi]
    peer = 1 - rank
    send_rows = np.nonzero(mask & (1 << peer))[0]
    assert np.all((send_rows >= lo) & (send_rows < hi)) and not np.any(mask & (1 << rank))
    payload = torch.tensor(np.concatenate([[len(send_rows)], send_rows, u_true.reshape(-1, 3)[send_rows].ravel()]), dtype=torch.float64)
    size = torch.tensor([payload.numel()], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world_size)]
    dist.all_gather(sizes, size)
    recv = torch.zeros(int(sizes[peer].item()), dtype=torch.float64)
    if rank == 0:
        dist.send(payload, dst=1); dist.recv(recv, src=1)
    else:
        dist.recv(recv, src=0); dist.send(payload, dst=0)
    k = int(recv[0].item())
    rr = recv[1:1 + k].numpy().astype(np.int64)
    u_mine.reshape(-1, 3)[rr] = recv[1 + k:].numpy().reshape(-1, 3)
    # the product of the own rows only touches own rows and received halo rows (no NaN left), and matches the global product
    y = _spmv_rows(rows, cols, vals, u_mine, lo, hi)
    y_ref = _spmv_rows(rows, cols, vals, u_true, lo, hi)
    ok = bool(np.all(np.isfinite(y)) and np.allclose(y, y_ref, rtol=0, atol=0))
    # the dot-product partial of this rank; the all-reduce (on the GPUs: partials pushed to every rank, same order everywhere)
    part = torch.tensor([float(y @ u_true[3 * lo:3 * hi])], dtype=torch.float64)
    parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world_size)]
    dist.all_gather(parts, part)
    total = sum(p.item() for p in parts)                      # fixed order: identical on both ranks
    out[rank] = (bounds.tolist(), ok, total, len(send_rows), len(rr))
    dist.destroy_process_group()


def test_two_rank_solver_partition_and_halo_over_gloo():
    import numpy as np
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_solver_worker, args=(2, port, out), nprocs=2, join=True)
    b0, ok0, t0, ns0, nr0 = out[0]
    b1, ok1, t1, ns1, nr1 = out[1]
    assert b0 == b1 and b0[0] == 0 and b0[-1] == 240 and 0 < b0[1] < 240     # the same partition everywhere, both ranks own rows
    assert ok0 and ok1                                                       # halo lists are sufficient and the products exact
    assert t0 == t1                                                          # identical reduction on both ranks
    assert ns0 == nr1 and ns1 == nr0 and ns0 > 0 and ns1 > 0                 # what one sends is what the other receives
    # against the undistributed product
    rows, cols, vals = _band_matrix(240, 6, seed=3)
    u = np.random.default_rng(7).standard_normal(720)
    y = _spmv_rows(rows, cols, vals, u, 0, 240)
    assert abs(t0 - float(y @ u)) <= 1e-12 * abs(float(y @ u))


def test_plan_partition_matches_virtual_grid_rule():
    """bounds[q] is where CTA (q, 0) of the virtual grid starts: first row r with rows[r] + 4 r >= cost * q * grid / (world * grid)."""
    import numpy as np
    from stark_b200 import dist as sbdist
    rows, cols, _ = _band_matrix(100, 4, seed=1, dense_row=False)
    for w in (1, 2, 4, 8):
        bounds, mask = sbdist.plan(rows, cols, w, 148, 0)
        cost = int(rows[-1]) + 4 * 100
        for q in range(w):
            t = (cost * q * 148) // (w * 148)
            r = next(r for r in range(101) if r == 100 or int(rows[r]) + 4 * r >= t)
            assert bounds[q] == r
        assert bounds[w] == 100
        if w == 1:
            assert not mask.any()


# ---- the whole distributed solve in numpy over gloo: Chronopoulos-Gear PCG, row slabs, halo pushes, shared partials ----

def _cg_worker(rank, world_size, port, out):
    """What the DIST instances of k_pcg_solve do (stark_b200/csrc/pcg.cu), restated per rank: every rank owns the rows
    [bounds[rank], bounds[rank + 1]) of the replicated matrix, keeps x, r, p, s, w only for them, pushes the rows of u = M^-1 r its
    peer reads (needmask) before every product, and all ranks sum the same per-rank partials in the same order."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from stark_b200 import dist as sbdist
    nbr = 180
    rows, cols, blocks = _band_matrix(nbr, 5, seed=11)
    vals = np.stack([b.T.reshape(9) for b in blocks]).astype(np.float32).reshape(-1)   # BCSR: float, column-major blocks
    rows_i = rows.astype(np.int64)
    b = np.random.default_rng(3).standard_normal(3 * nbr)
    bounds, mask = sbdist.plan(rows, cols, world_size, 7, rank)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    peer = 1 - rank
    send_rows = np.nonzero(mask & (1 << peer))[0]
    n_send = torch.tensor([len(send_rows)], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world_size)]
    dist.all_gather(counts, n_send)
    recv_rows_t = torch.zeros(int(counts[peer].item()), dtype=torch.int64)
    if rank == 0:
        dist.send(torch.tensor(send_rows, dtype=torch.int64), dst=1); dist.recv(recv_rows_t, src=1)
    else:
        dist.recv(recv_rows_t, src=0); dist.send(torch.tensor(send_rows, dtype=torch.int64), dst=0)
    recv_rows = recv_rows_t.numpy()
    dinv = oracle.block_jacobi(rows_i, cols, vals).astype(np.float64).reshape(-1, 3, 3)

    def own(v):
        return v[3 * lo:3 * hi]

    def prec_own(r_own):
        return np.einsum("kij,kj->ki", dinv[lo:hi], r_own.reshape(-1, 3)).reshape(-1)

    def exchange(u_full):
        """halo push: my rows the peer reads -> its copy of u; its rows I read -> mine (NaN everywhere else stays NaN)"""
        payload = torch.tensor(u_full.reshape(-1, 3)[send_rows].ravel(), dtype=torch.float64)
        got = torch.zeros(3 * len(recv_rows), dtype=torch.float64)
        if rank == 0:
            dist.send(payload, dst=1); dist.recv(got, src=1)
        else:
            dist.recv(got, src=0); dist.send(payload, dst=0)
        u_full.reshape(-1, 3)[recv_rows] = got.numpy().reshape(-1, 3)

    def allsum(*partials):
        t = torch.tensor(list(partials), dtype=torch.float64)
        parts = [torch.zeros_like(t) for _ in range(world_size)]
        dist.all_gather(parts, t)
        return [sum(p[k].item() for p in parts) for k in range(len(partials))]     # rank order: identical everywhere

    blocks64 = vals.reshape(-1, 3, 3).transpose(0, 2, 1).astype(np.float64)   # the float-stored blocks, row-major

    def spmv_own(u_full):
        return _spmv_rows(rows_i, cols, blocks64, u_full, lo, hi)

    abs_tol, rel_tol, max_iter = 1e-10, 1e-12, 500
    x = np.zeros(3 * (hi - lo)); r = own(b).copy(); p = np.zeros_like(x); s = np.zeros_like(x)
    u_full = np.full(3 * nbr, np.nan)
    u_full[3 * lo:3 * hi] = prec_own(r)
    bb, gamma = allsum(float(r @ r), float(r @ own(u_full)))
    exchange(u_full)
    w = spmv_own(u_full)
    assert np.all(np.isfinite(w))                     # the halo lists cover every column the own rows read
    (delta,) = allsum(float(w @ own(u_full)))
    alpha, beta, it, ok = gamma / delta, 0.0, 0, False
    while it < max_iter:
        it += 1
        u_own = own(u_full).copy()
        p = u_own + beta * p
        s = w + beta * s
        x += alpha * p
        r -= alpha * s
        u_full[:] = np.nan
        u_full[3 * lo:3 * hi] = prec_own(r)
        rr, gamma_new = allsum(float(r @ r), float(r @ own(u_full)))
        if np.sqrt(rr / bb) < abs_tol or np.sqrt(rr / bb) < rel_tol:
            ok = True
            break
        exchange(u_full)
        w = spmv_own(u_full)
        (delta,) = allsum(float(w @ own(u_full)))
        beta = gamma_new / gamma
        pAp = delta - beta * gamma_new / alpha
        gamma = gamma_new
        assert pAp > 0.0
        alpha = gamma / pAp
    # du: every rank's slice to everybody
    mine = torch.zeros(3 * nbr, dtype=torch.float64)
    mine[3 * lo:3 * hi] = torch.tensor(x)
    dist.all_reduce(mine)
    out[rank] = (it, ok, mine.numpy().tolist())
    dist.destroy_process_group()


def test_two_rank_distributed_pcg_over_gloo_matches_the_oracle():
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    mgr = mp.Manager()
    out = mgr.dict()
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_cg_worker, args=(2, port, out), nprocs=2, join=True)
    it0, ok0, x0 = out[0]
    it1, ok1, x1 = out[1]
    assert ok0 and ok1 and it0 == it1 and x0 == x1            # same decisions, bit-identical solution on both ranks
    rows, cols, blocks = _band_matrix(180, 5, seed=11)
    vals = np.stack([b.T.reshape(9) for b in blocks]).astype(np.float32).reshape(-1)
    b = np.random.default_rng(3).standard_normal(540)
    x_ref, it_ref, ok_ref = oracle.solve_pcg(rows.astype(np.int64), cols, vals, b, 1e-10, 1e-12, 500)
    assert ok_ref and abs(it_ref - it0) <= 1                  # Chronopoulos-Gear = textbook PCG in exact arithmetic
    assert np.abs(np.array(x0) - x_ref).max() <= 1e-8 * np.abs(x_ref).max()


def test_plan_send_lists_are_consistent_for_eight_ranks():
    """For every pair of ranks: what r pushes to q (needmask of r, bit q) is exactly the set of columns in r's rows that q's rows
    reference -- checked in one process on a band matrix with a dense row, for the 8-rank partition of one NVSwitch domain."""
    import numpy as np
    from stark_b200 import dist as sbdist
    nbr, W, grid = 400, 8, 3
    rows, cols, _ = _band_matrix(nbr, 7, seed=5)
    plans = [sbdist.plan(rows, cols, W, grid, r) for r in range(W)]
    bounds = plans[0][0]
    assert all(np.array_equal(bounds, p[0]) for p in plans) and bounds[0] == 0 and bounds[W] == nbr and np.all(np.diff(bounds) > 0)
    owner = np.searchsorted(bounds, np.arange(nbr), side="right") - 1
    for r in range(W):
        mask = plans[r][1]
        assert not np.any(mask[owner != r]) and not np.any(mask & (1 << r))
        for q in range(W):
            if q == r:
                continue
            need = set()
            for i in range(int(bounds[q]), int(bounds[q + 1])):
                for c in cols[int(rows[i]):int(rows[i + 1])] // 3:
                    if owner[c] == r:
                        need.add(int(c))
            assert need == set(np.nonzero(mask & (1 << q))[0].tolist()), (r, q)
    # the dense last row (a rigid body) makes the last rank read rows of every other rank
    assert all(np.any(plans[r][1] & (1 << (W - 1))) for r in range(W - 1))
