"""Multi-GPU plumbing: one process per GPU.  torch.distributed carries only the set-up and the bookkeeping -- the exchange of
the CUDA IPC handles of the ranks' peer buffers (connect_solver), the barrier around the timed region and the reduction of the
per-rank timings / work counters.  The data plane of the distributed linear solve (halo of u, dot-product partials, barrier,
du) is NVLink loads / stores inside the persistent PCG kernel (stark_b200/csrc/pcg.cu), not a collective call.
Works with the `nccl` backend on GPUs and with `gloo` on CPU tensors (tests)."""
import ctypes as C

import numpy as np
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


def connect_solver(lib, ctx_handle, max_dofs, device=None):
    """Make every later block-Jacobi PCG solve of this context ONE solve shared by all ranks of the process group
    (sb_dist_init + all-gather of the IPC handles + sb_dist_connect).  No-op for a single rank.  Returns the world size."""
    w = world()
    if w <= 1:
        return 1
    rank = dist.get_rank()
    handle = (C.c_ubyte * 64)()
    rc = lib.sb_dist_init(ctx_handle, rank, w, int(max_dofs), handle)
    if rc != 0:
        raise RuntimeError(f"sb_dist_init failed ({rc}): {lib.sb_last_error(ctx_handle).decode()}")
    mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    gathered = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(gathered, mine)
    blob = np.concatenate([g.cpu().numpy() for g in gathered]).astype(np.uint8)
    rc = lib.sb_dist_connect(ctx_handle, blob.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"sb_dist_connect failed ({rc}): {lib.sb_last_error(ctx_handle).decode()}")
    dist.barrier()   # every rank has mapped every peer buffer before the first solve pushes into them
    return w


def plan(rows, cols, world_size, grid, rank):
    """Host restatement of the solve's row partition and halo send lists (sb_dist_plan): (bounds[world + 1], needmask[nbr])."""
    from . import capi
    lib = capi.load()
    rows = np.ascontiguousarray(rows, dtype=np.uint64)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    nbr = len(rows) - 1
    bounds = np.zeros(world_size + 1, dtype=np.int32)
    mask = np.zeros(max(nbr, 1), dtype=np.uint8)
    rc = lib.sb_dist_plan(nbr, rows.ctypes.data_as(C.POINTER(C.c_ulonglong)), cols.ctypes.data_as(C.POINTER(C.c_int32)), int(world_size), int(grid), int(rank),
                          bounds.ctypes.data_as(C.POINTER(C.c_int32)), mask.ctypes.data_as(C.POINTER(C.c_ubyte)))
    if rc != 0:
        raise RuntimeError(f"sb_dist_plan failed ({rc})")
    return bounds, mask[:nbr]
