#!/usr/bin/env python
"""Distributed block-Jacobi PCG over peer memory, checked against the single-GPU solve of the same system.

    python tools/dist_selftest.py [--world 2] [--fixture tetdrop_n5]            # W "ranks" as threads on ONE GPU (SB_PCG_GRID small)
    torchrun --nproc-per-node N tools/dist_selftest.py --multiprocess [...]      # one process per GPU, peer buffers through CUDA IPC

Prints one JSON line: iterations of the reference / distributed solves, max |du_dist - du_ref| / |du_ref|, whether every rank
returned bit-identical du.  Exit code 0 only if all of it holds (tests/test_dist_gpu.py runs the one-GPU form)."""
import argparse
import ctypes as C
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=2)
ap.add_argument("--fixture", default="tetdrop_n5")
ap.add_argument("--multiprocess", action="store_true")
ap.add_argument("--solves", type=int, default=3)
args = ap.parse_args()

if not args.multiprocess:
    os.environ.setdefault("SB_PCG_GRID", "16")           # W kernels of 16 CTAs share the GPU
    os.environ.setdefault("SB_DIST_TIMEOUT_S", "20")
os.environ.setdefault("SB_DIST_POLICY", "always")        # the fixtures are tiny: the automatic policy would keep their solves local

import numpy as np  # noqa: E402
from golden_util import Golden, bind  # noqa: E402
from stark_b200 import capi  # noqa: E402


def prepared_context(device):
    g = Golden(args.fixture)
    ctx = capi.Context(device)
    bind(ctx, g, set(capi.kernel_names()))
    ctx.eval("PGH")
    ctx.project_to_pd(0.0)
    ctx.assemble()
    return g, ctx


def solve(ctx, g):
    return ctx.solve_pcg(1e-9, 1e-12, 10000, True)


if args.multiprocess:
    import torch
    import torch.distributed as dist
    from stark_b200 import dist as sbdist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g, ctx = prepared_context(local)
    ref = solve(ctx, g)                       # single-GPU solve first (not yet connected)
    du_ref = ctx.du()
    sbdist.connect_solver(ctx.lib, ctx.h, ctx.ndofs(), device=torch.device("cuda", local))
    outs, dus = [], []
    for _ in range(args.solves):
        outs.append(solve(ctx, g))
        dus.append(ctx.du())
    mine = torch.tensor(dus[-1], device="cuda")
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    identical = all(torch.equal(allv[0], v) for v in allv)
    err = float(np.abs(dus[-1] - du_ref).max() / np.abs(du_ref).max())
    ok = bool(identical and all(o["ok"] for o in outs) and err < 1e-6 and all(abs(o["iterations"] - ref["iterations"]) <= 2 for o in outs)
              and all(np.array_equal(dus[0], d) for d in dus))
    if rank == 0:
        print(json.dumps({"mode": "multiprocess", "world": world, "fixture": args.fixture, "ref_iterations": ref["iterations"],
                          "dist_iterations": [o["iterations"] for o in outs], "rel_err": err, "identical_across_ranks": identical, "ok": ok}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)

# ---- one GPU, W contexts, W threads ----
W = args.world
g, ref_ctx = prepared_context(0)
ref = solve(ref_ctx, g)
du_ref = ref_ctx.du()
ctxs = [prepared_context(0)[1] for _ in range(W)]
lib = ctxs[0].lib
for r, c in enumerate(ctxs):
    rc = lib.sb_dist_init(c.h, r, W, c.ndofs(), None)
    assert rc == 0, lib.sb_last_error(c.h)
bases = (C.c_void_p * W)(*[lib.sb_dist_local_base(c.h) for c in ctxs])
for c in ctxs:
    assert lib.sb_dist_connect_ptrs(c.h, bases) == 0
results = [[None] * W for _ in range(args.solves)]
dus = [[None] * W for _ in range(args.solves)]
errors = []
for k in range(args.solves):
    def work(r, k=k):
        try:
            results[k][r] = solve(ctxs[r], g)
            dus[k][r] = ctxs[r].du()
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=work, args=(r,)) for r in range(W)]
    for t in th:
        t.start()
    for t in th:
        t.join(120)
    if errors or any(t.is_alive() for t in th):
        print(json.dumps({"mode": "threads", "ok": False, "errors": errors, "hung": any(t.is_alive() for t in th)}))
        os._exit(1)
# an evaluation on connected contexts ends with rank 0's gradient / energy on every rank (dist_bcast_from_root): perturb the
# ranks' states differently in the last bits first, so that their own gradients differ
grads, energies = [None] * W, [None] * W
for r, c in enumerate(ctxs):
    u = c.dofs_get()
    c.dofs_set(u * (1.0 + r * 2.0 ** -50))
def eval_work(r):
    try:
        energies[r], _ = ctxs[r].eval("PGH")
        grads[r] = ctxs[r].grad()
    except Exception as e:   # noqa: BLE001
        errors.append(repr(e))
th = [threading.Thread(target=eval_work, args=(r,)) for r in range(W)]
for t in th:
    t.start()
for t in th:
    t.join(120)
if errors or any(t.is_alive() for t in th):
    print(json.dumps({"mode": "threads", "ok": False, "errors": errors, "hung": any(t.is_alive() for t in th)}))
    os._exit(1)
bcast_ok = all(np.array_equal(grads[0], grads[r]) and energies[0] == energies[r] for r in range(W))
identical = all(np.array_equal(dus[k][0], dus[k][r]) for k in range(args.solves) for r in range(W))
repeat = all(np.array_equal(dus[0][0], dus[k][0]) for k in range(args.solves))
err = float(np.abs(dus[-1][0] - du_ref).max() / np.abs(du_ref).max())
its = [results[k][0]["iterations"] for k in range(args.solves)]
ok = bool(identical and repeat and bcast_ok and err < 1e-6 and all(results[k][r]["ok"] for k in range(args.solves) for r in range(W))
          and all(abs(i - ref["iterations"]) <= 2 for i in its))
st = (C.c_double * 4)()
lib.sb_dist_stats(ctxs[0].h, None, None, st)
print(json.dumps({"mode": "threads", "world": W, "fixture": args.fixture, "ref_iterations": ref["iterations"], "dist_iterations": its, "rel_err": err,
                  "identical_across_ranks": identical, "identical_across_solves": repeat, "rank0_gradient_everywhere": bcast_ok, "barriers": st[0], "ok": ok}))
sys.exit(0 if ok else 1)
