"""Multi-GPU plumbing of the benchmark: one process per GPU, scenes are independent replicas (DESIGN.md "Multi-GPU"), so the
only communication is the barrier around the timed region and the reduction of the per-rank timings / work counters.
Works with the `nccl` backend on GPUs and with `gloo` on CPU tensors (tests)."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def barrier(device=None):
    if world() > 1:
        dist.barrier()
    if device is not None and torch.cuda.is_available():
        torch.cuda.synchronize(device)


def aggregate(timings_ms, work, device="cpu"):
    """timings_ms: list of per-rank times (the job time is the MAX over ranks); work: list of per-rank counters (the job's
    work is the SUM over ranks).  Returns (max timings, summed work) as Python lists, identical on every rank."""
    t = torch.tensor(list(timings_ms), dtype=torch.float64, device=device)
    w = torch.tensor(list(work), dtype=torch.float64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return t.tolist(), w.tolist()


def throughput(timings_ms, work, device="cpu"):
    """Whole-job units per second for every (time, work) pair: sum of the work of all ranks / max time over ranks."""
    t, w = aggregate(timings_ms, work, device)
    return [wi / (ti * 1e-3) if ti > 0 else 0.0 for ti, wi in zip(t, w)]
