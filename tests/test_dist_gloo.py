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
