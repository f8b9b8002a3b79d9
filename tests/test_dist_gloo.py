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
