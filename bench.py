#!/usr/bin/env python
"""Benchmark of the STARK Newton hot path (BASELINE.json metric: Newton iterations / second, incl. contact detection,
element evaluation, PD projection, assembly, linear solve and line search).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one simulation time step (stark::Simulation::run_one_time_step): the Newton solve of that step is the hot
path.  Workload at N = 1: BASELINE.json configs[1] (C2) -- 26^3 tet grid (210,912 tets, Soft_Rubber stable Neo-Hookean)
dropped on a fixed rigid floor with IPC contact and friction; `--config C1|C3|C4|C5` runs the other BASELINE configurations
(both arms).  For N > 1 the N ranks work on ONE scene (strong scaling): every block-Jacobi PCG solve is shared by all ranks --
row slabs of the matrix, each rank keeping 1/N of it in shared memory, halo / partial sums / barrier over NVLink peer memory
inside the persistent kernel (DESIGN.md "Multi-GPU"); each rank assembles only its own rows, element evaluation and projection are
replicated with rank 0's gradient / energy broadcast to every rank.  The default configuration at N > 1 is C5 (the 1M-tet bar
BASELINE.json names for the 1/2/4/8-GPU sweep); the line carries the same run's single-GPU rate.  `--replicas` runs N independent
scenes instead (weak scaling, no data-plane exchange).

Printed keys (one JSON line on rank 0):
  value   Newton iterations / s with the state resident in HBM: sum(iterations) / sum(device time of the solves, CUDA events)
  e2e     the same metric through the host API (stark_b200::Simulation::run_one_time_step): per-step host->device upload
          of the state arrays and device->host read of the solution inside the timed region
  roofline / cpu_baseline   see DESIGN.md "Measurement"
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "newton_iterations_per_second"
UNIT = "Newton it/s"
TET_BYTES = 1632       # algorithmic bytes of one EnergyTetStrain element in PGH mode

# BASELINE.json configs (SURVEY.md section 8, table of concrete instantiations).  The driver's default is C2 (the configuration
# the metric is quoted on); the others are run by hand (`--config C5`) and recorded under profiles/.
CONFIGS = {
    "C1": dict(scene="cloth", n=32, ny=-1, nz=-1, steps=40, warmup=5,
               workload="C1 cloth: 32x32 Cotton_Fabric surface grid (triangle strain + flat bending) over a scripted fixed rigid box, IPC contact d=2mm, dt=10ms, PPN+BDPCG defaults"),
    "C2": dict(scene="tetdrop", n=26, ny=-1, nz=-1, steps=30, warmup=5,
               workload="C2 tetdrop: 26^3 Soft_Rubber tet grid (12 tets/hex) on a fixed rigid floor, IPC contact d=1mm k_min=1e8 mu=0.5, dt=10ms, PPN+BDPCG defaults"),
    "C3": dict(scene="cloth_shells", n=256, ny=-1, nz=-1, steps=4, warmup=3,
               workload="C3 cloth: 256x256 Cotton_Fabric grid (0.4 m, edges 1.56 mm) with discrete-shell hinges over a scripted fixed rigid box, IPC contact d=0.47mm (0.3 edge lengths), friction mu=0.3, dt=10ms, PPN+BDPCG defaults"),
    "C4": dict(scene="tetchain", n=16, ny=10, nz=-1, steps=20, warmup=5,
               workload="C4 tetchain: 16^3 tet grid (49,152 tets, bottom face prescribed) under a chain of 10 hinged rigid boxes, IPC contact + friction mu=0.3, dt=10ms, PPN+BDPCG defaults"),
    "C5": dict(scene="tetbar", n=22, ny=22, nz=172, steps=10, warmup=3,
               workload="C5 tetbar: 22x22x172 Soft_Rubber tet bar (998,976 tets), end caps prescribed, one cap turning 90 deg/s, no contact, dt=10ms, PPN+BDPCG defaults"),
}


def workload_config(n_gpus, name, cfg, grid=None, replicas=False):
    n = grid if grid else cfg["n"]
    reduced = grid is not None and grid != cfg["n"]
    if n_gpus == 1:
        par = "single GPU"
    elif replicas:
        par = f"{n_gpus} independent replicas (one scene per GPU)"
    else:
        par = (f"slab decomposition of the linear solve over {n_gpus} GPUs: contiguous block-row ranges of the 3x3-BCSR per rank (1/{n_gpus} of the matrix in each "
               "rank's shared memory), halo of u + dot-product partials + barrier as NVLink peer-memory stores inside the persistent PCG kernel; each rank "
               "assembles only its own rows; element evaluation and projection replicated on every rank (rank 0's gradient / energy broadcast)")
    return {"workload": cfg["workload"] if not reduced else f"{cfg['scene']} at a reduced grid {n} (NOT the benchmark configuration)",
            "name": name, "scene": cfg["scene"], "grid": n, "dt": 0.01,
            "parallelism": par,
            "l2_policy": "the element outputs written and read by every evaluation (1,152 B per tet / hinge) exceed the 126 MB L2 at the C2, C3 and C5 sizes; C1 and C4 are L2-resident (latency-bound scenes)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): the sampler runs from
    before the warm-up, every sample is stamped on arrival and only those inside [t_begin, t_end] are summarised."""

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        if self.proc:
            self.proc.terminate()
        inside = [s for t, s in self.samples if t_begin is None or (t_begin <= t <= t_end)]
        if not inside and self.samples:   # very short region: the sample closest to it
            mid = 0.5 * ((t_begin or 0.0) + (t_end or 0.0))
            inside = [min(self.samples, key=lambda ts: abs(ts[0] - mid))[1]]
        sm = sorted(int(s[0]) for s in inside if s and s[0].isdigit())
        mx = [int(s[1]) for s in inside if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in inside if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def reference_cmd(cfg, steps, warmup, grid=None, llt=False):
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    codegen = f"/tmp/stark_ref_codegen_{os.getuid()}"   # JIT cache of the reference on THIS machine (its kernels use -march=native)
    cmd = [driver, "--scene", cfg["scene"], "--n", str(grid if grid else cfg["n"]), "--bench", "--steps", str(steps), "--warmup", str(warmup), "--codegen", codegen,
           "--threads", str(os.cpu_count() or 1)]
    if cfg["ny"] > 0:
        cmd += ["--ny", str(cfg["ny"])]
    if cfg["nz"] > 0:
        cmd += ["--nz", str(cfg["nz"])]
    if llt:
        cmd += ["--llt"]
    return driver, cmd


def run_reference(cfg, steps, warmup, grid=None, llt=False):
    """The UNMODIFIED reference (oracle/_ref/ref_driver, built from /root/reference by oracle/Makefile.ref) on the host cores."""
    driver, cmd = reference_cmd(cfg, steps, warmup, grid, llt)
    if not os.path.exists(driver):
        return None
    env = dict(os.environ, CXX="/usr/bin/g++")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=3000)
    for line in out.stdout.splitlines()[::-1]:
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("reference driver produced no result: " + out.stdout[-400:] + out.stderr[-400:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="stark_b200")
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration; default: C2 at --gpus 1 (the 1xB200 configuration the metric is quoted on), C5 at --gpus N > 1 (the "
                         "1M-tet bar BASELINE.json names for the 1/2/4/8-GPU sweep)")
    ap.add_argument("--grid", type=int, default=None, help="override the grid size (diagnostic: NOT the benchmark configuration)")
    ap.add_argument("--llt", action="store_true", help="DirectLLT instead of the default BDPCG (both arms; reference: Eigen SimplicialLLT, ours: sparse tile-envelope Cholesky with DMMA tile products)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", action="store_true", help="N > 1: N independent scenes (weak scaling) instead of one scene with the distributed solve")
    ap.add_argument("--stage-steps", type=int, default=4, help="extra (untimed) steps run with stage profiling on after the timed region; 0 = off")
    args = ap.parse_args()
    if args.config is None:
        args.config = "C2" if args.gpus <= 1 else "C5"
    cfg = CONFIGS[args.config]
    steps = args.steps if args.steps is not None else cfg["steps"]
    warmup = max(args.warmup if args.warmup is not None else cfg["warmup"], 3)
    grid = args.grid if args.grid else cfg["n"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the same window of time steps as the stark_b200 arm (same warm-up rule), bounded so that the run ends within minutes
        ref_steps = max(1, min(steps, 200))
        r = run_reference(cfg, ref_steps, warmup, args.grid, args.llt)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver has not been built (oracle/Makefile.ref)"}))
            return 0
        v = r["newton_it_per_s"]
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": ref_steps, "warmup": warmup,
                "ms_per_step": 1e3 * r["wall_s"] / max(1, r["steps"]), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(args.gpus, args.config, cfg, args.grid, args.replicas or args.llt), "newton_iterations": r["newton_iterations"],
                "accepted_steps": r.get("accepted_steps"), "cg_iterations": r.get("cg_iterations"), "linear_solver": "DirectLLT" if args.llt else "BDPCG",
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["threads"], "kind": "reference",
                                 "sample": f"{ref_steps} time steps of the same scene after {warmup} warm-up steps (unmodified reference, all host threads)"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import ctypes as C
    import torch
    import torch.distributed as dist
    from stark_b200 import capi, scenes
    from stark_b200 import dist as sbdist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the stark_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # (NCCL prints its version banner on stdout at the first collective: keep stdout for the one JSON line)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    stream = torch.cuda.Stream()
    sc = scenes.Scene(cfg["scene"], n=grid, ny=cfg["ny"], nz=cfg["nz"], dt=0.01, drop=0.003, device=local_rank, stream=stream.cuda_stream, llt=args.llt)
    distributed = world > 1 and not args.replicas and not args.llt
    if distributed:
        # one scene, N ranks: every PCG solve from here on is shared by all ranks (peer buffers mapped through CUDA IPC).
        # Connected BEFORE the first step, so that the replicas never differ (rank 0's gradient / energy are everybody's from the
        # first evaluation on); the scene registers its degrees of freedom at its first step: nodes + up to 1,024 rigid bodies
        warmup_left = warmup
        sbdist.connect_solver(capi.load(), C.c_void_p(sc.lib.sbh_scene_context(sc.h)), 3 * int(sc.totals()["nodes"]) + 6 * 1024, device=torch.device("cuda", local_rank))

    def barrier():
        sbdist.barrier(torch.device("cuda", local_rank))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(warmup_left if distributed else warmup):
        sc.step()
    t0 = sc.totals()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    wall0 = time.perf_counter()
    its = evals = cg = accepted = 0
    solve_gpu_ms = 0.0
    for _ in range(steps):
        s = sc.step()
        its += int(s["newton_iterations"]); evals += int(s["evaluations"]); cg += int(s["cg_iterations"]); accepted += int(s["accepted"])
        solve_gpu_ms += s["solve_gpu_ms"]
    sc.sync()   # the read-back of the last step's positions / velocities into the host mirrors ends inside the timed region
    ev1.record(stream)
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    wall_s = time.perf_counter() - wall0
    clocks = sampler.stop(wall0, wall0 + wall_s) if rank == 0 else None
    t1 = sc.totals()

    # max over ranks of the timed regions, sum over ranks of the work
    (e2e_ms, solve_gpu_ms), (its_all, evals_all, cg_all) = sbdist.aggregate([e2e_ms, solve_gpu_ms], [float(its), float(evals), float(cg)], device="cuda")
    if distributed:   # ONE scene: the work is rank 0's, not the sum over the ranks that shared it
        its_all, evals_all, cg_all = float(its), float(evals), float(cg)
    ctx_handle = C.c_void_p(sc.lib.sbh_scene_context(sc.h))
    lib = capi.load()
    # ---- per-stage breakdown (diagnostic, outside the timed region): extra steps with a stream sync at every stage boundary ----
    # (with the distributed solve every rank takes these steps: a solve needs all of them)
    stages = None
    if args.stage_steps > 0 and (rank == 0 or distributed):
        try:
            lib.sb_profile_stages(ctx_handle, 1)
            it_s = 0
            for _ in range(args.stage_steps):
                it_s += int(sc.step()["newton_iterations"])
            rep = lib.sb_profile_report(ctx_handle).decode()
            lib.sb_profile_stages(ctx_handle, 0)
            stages = {"steps": args.stage_steps, "newton_iterations": it_s}
            for ln in rep.splitlines():
                name, ms, calls = ln.split()
                stages[name] = {"ms": round(float(ms), 4), "calls": int(calls)}
        except Exception as e:
            stages = {"error": repr(e)}
            if distributed:
                raise
    # ---- the single-GPU rate of the SAME scene in the SAME run (outside the timed region): sharing switched off on every rank,
    #      each rank then solves its own replica locally; rank 0's rate is the strong-scaling baseline of this line.  Last thing
    #      that steps the scene: the replicas are no longer kept identical once the sharing is off ----
    single = None
    if distributed:
        lib.sb_dist_set_enabled(ctx_handle, 0)
        n_single = max(2, min(steps, 6))
        sc.step()   # (first local solve: kernel attributes of the local instance)
        its1, ms1 = 0, 0.0
        for _ in range(n_single):
            s1 = sc.step()
            its1 += int(s1["newton_iterations"]); ms1 += s1["solve_gpu_ms"]
        barrier()   # (the sharing stays off: nothing after this point needs the other ranks)
        single = {"value": its1 / (ms1 * 1e-3) if ms1 > 0 else None, "unit": UNIT, "steps": n_single,
                  "note": "same scene, same process, the steps right after the timed region with the sharing of solves switched off (every rank solves locally); device time as `value`"}
    if rank != 0:
        if world > 1:
            dist.barrier()   # rank 0 finishes its single-GPU diagnostics before the group is torn down
            dist.destroy_process_group()
        return 0

    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
    peak = peaks["hbm_gbs"] if peaks else 6650.0
    peak_source = "MEASURED_PEAKS.json hbm_gbs (burst: kernels timed alone / inside their own launch)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    # ---- rooflines, both measured live in this run ----
    # (1) element evaluation: the EnergyTetStrain P + grad + Hessian kernel timed alone with CUDA events on the context stream
    # (2) the PCG kernel: in-kernel timers of the iteration loop (global timer, read in the kernel; stage steps above) over the
    #     algorithmic bytes of one iteration (nnzb 40 + 8 (nbr + 1) + 156 ndofs)
    kernels = []
    try:
        n_tets = int(t1["tets"])
        if n_tets > 0 and not args.llt:
            E, g = C.c_double(), C.c_double()
            lib.sb_eval(ctx_handle, 2, C.byref(E), C.byref(g))
            pot = int(sc.lib.sbh_scene_potential(sc.h, b"EnergyTetStrain"))
            ms = C.c_double()
            lib.sb_profile_potential(ctx_handle, pot, 2, 3, C.byref(ms))       # warm-up
            lib.sb_profile_potential(ctx_handle, pot, 2, 20, C.byref(ms))
            achieved = TET_BYTES * n_tets / (ms.value * 1e-3) / 1e9
            kernels.append({"kernel": "k_tet_analytic: EnergyTetStrain element evaluation (P + grad + dense 12x12 Hessian)", "bound": "hbm", "achieved": achieved, "peak": peak,
                            "peak_source": peak_source, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                            "traffic_note": "dram__bytes of this kernel are not measurable inside a bench run; the ncu --set full capture of the same launch is profiles/r2_ncu_kernels.csv",
                            "launch_ms": ms.value, "algorithmic_bytes_per_launch": TET_BYTES * n_tets,
                            "how": "20 launches alone on the context stream between two CUDA events; working set per launch exceeds L2 at the C2 / C5 sizes"})
    except Exception as e:   # the line must still print
        kernels.append({"kernel": "k_tet_analytic", "error": repr(e)})
    try:
        if stages and "cg_iterations" in stages and stages["cg_iterations"]["calls"] > 0:
            nbr, nnzb = C.c_int(), C.c_int64()
            lib.sb_bcsr_info(ctx_handle, C.byref(nbr), C.byref(nnzb))
            ndofs = int(t1["ndofs"])
            bytes_it = 40 * nnzb.value + 8 * (nbr.value + 1) + 156 * ndofs
            us_it = 1e3 * stages["cg_iterations"]["ms"] / stages["cg_iterations"]["calls"]
            achieved = bytes_it / (us_it * 1e-6) / 1e9
            peak_all = peak * (world if distributed else 1)   # a distributed iteration runs on all ranks' GPUs
            kernels.append({"kernel": "k_pcg_solve: one block-Jacobi PCG iteration (SpMV from the resident / streamed 3x3-BCSR + vector phase + 2 grid barriers)", "bound": "hbm",
                            "achieved": achieved, "peak": peak_all, "peak_source": peak_source + (f" x {world} GPUs" if distributed else ""), "unit": "GB/s", "frac": achieved / peak_all, "traffic": None,
                            "us_per_iteration": us_it, "algorithmic_bytes_per_iteration": bytes_it, "iterations_timed": stages["cg_iterations"]["calls"],
                            "how": "the kernel's own %globaltimer around its iteration loop, summed over the stage-profiled steps (one launch per solve)"})
    except Exception as e:
        kernels.append({"kernel": "k_pcg_solve", "error": repr(e)})
    roof = dict(kernels[0]) if kernels and "error" not in kernels[0] else {"error": "no roofline kernel in this configuration"}
    roof["kernels"] = kernels

    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample of the same window ----
    cpu = None
    if not args.no_cpu_baseline and args.gpus == 1:
        try:
            n_cpu = max(1, min(steps, 12))
            r = run_reference(cfg, n_cpu, warmup, args.grid, args.llt)
            if r is not None:
                cpu = {"value": r["newton_it_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                       "sample": f"the first {n_cpu} of the window's time steps after the same {warmup} warm-up steps (unmodified reference, all host threads; the full window is what `--impl reference` times)",
                       "newton_iterations": r["newton_iterations"], "accepted_steps": r.get("accepted_steps"), "wall_s": r["wall_s"]}
        except Exception as e:
            cpu = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": its_all / (solve_gpu_ms * 1e-3), "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": e2e_ms / steps, "higher_is_better": True, "scaling": "strong" if distributed else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, args.config, cfg, args.grid, args.replicas or args.llt),
        "value_note": "Newton iterations / device time of the steps' Newton path (CUDA events on the context stream from the start-of-step collision detection to the end of the solve), state resident in HBM",
        "e2e": {"value": its_all / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": (t1["h2d_bytes"] - t0["h2d_bytes"]) / steps, "d2h_bytes_per_step": (t1["d2h_bytes"] - t0["d2h_bytes"]) / steps,
                "note": "through stark_b200::Simulation::run_one_time_step: per-step uploads of whatever host state changed (scripted boundary data, rigid-body state), read-back of the new positions / velocities into the pinned host mirrors and of the rigid DoFs"},
        "gpu_launches": int(t1["launches"] - t0["launches"]),
        "newton_iterations": its_all, "evaluations": evals_all, "cg_iterations": cg_all, "accepted_steps": accepted, "wall_s_rank0": wall_s,
        "linear_solver": "DirectLLT" if args.llt else "BDPCG",
        "solve_gpu_ms_per_iteration": solve_gpu_ms / max(1.0, its_all if distributed else its_all / world),
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "stages": stages,
    }
    if distributed:
        st4 = (C.c_double * 4)()
        lib.sb_dist_stats(ctx_handle, None, None, st4)
        line["single_gpu_same_config"] = single
        line["distributed"] = {"solves_shared_by_all_ranks": int(st4[1]), "solves_kept_local_by_policy": int(st4[3]), "cross_gpu_barriers": int(st4[0]),
                               "peer_buffer_bytes": int(st4[2]), "policy": os.environ.get("SB_DIST_POLICY", "auto"),
                               "note": "policy auto: a matrix resident in one GPU's shared memory is solved locally by every rank (identical results); larger systems are shared"}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
