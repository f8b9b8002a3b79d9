#!/usr/bin/env python
"""Benchmark of the STARK Newton hot path (BASELINE.json metric: Newton iterations / second, incl. contact detection,
element evaluation, PD projection, assembly, linear solve and line search).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one simulation time step (stark::Simulation::run_one_time_step): the Newton solve of that step is the hot
path.  Workload at N = 1: BASELINE.json configs[1] -- 26^3 tet grid (210,912 tets, Soft_Rubber stable Neo-Hookean)
dropped on a fixed rigid floor with IPC contact and friction.  For N > 1 every rank runs its own replica of the scene
(the path does not shard below one scene at this size -- DESIGN.md "Multi-GPU") and `value` is the aggregate.

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
GRID_N = 26            # 26^3 hexahedra x 12 tets = 210,912 tets (BASELINE.json configs[1])
TET_BYTES = 1632       # algorithmic bytes of one EnergyTetStrain element in PGH mode
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_tet_analytic launch at GRID_N = 26 from the `ncu --set full` capture
# committed as profiles/r1_ncu_full_tet_assembly.csv (8.70 MB read + 199.07 MB written; gathers hit L2, the tail of the
# Hessian stream is still in L2 when the kernel ends)
TET_DRAM_TRAFFIC_NCU = 207.8e6


def workload_config(n_gpus, grid=GRID_N):
    name = "C2 tetdrop" if grid == GRID_N else "tetdrop (reduced grid: NOT the benchmark configuration)"
    return {"workload": f"{name}: {grid}^3 Soft_Rubber tet grid (12 tets/hex) on a fixed rigid floor, IPC contact d=1mm k_min=1e8 mu=0.5, dt=10ms, PPN+BDPCG defaults",
            "tets": 12 * grid ** 3, "grid": grid, "dt": 0.01,
            "parallelism": "single GPU" if n_gpus == 1 else f"{n_gpus} independent replicas (one scene per GPU)",
            "l2_policy": ("working set per evaluation (~350 MB element outputs) exceeds the 126 MB L2" if grid == GRID_N
                          else f"working set per evaluation ~{12 * grid ** 3 * 1632 / 1e6:.0f} MB")}


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


def reference_cmd(steps, warmup, grid=GRID_N):
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    codegen = f"/tmp/stark_ref_codegen_{os.getuid()}"   # JIT cache of the reference on THIS machine (its kernels use -march=native)
    return driver, [driver, "--scene", "tetdrop", "--n", str(grid), "--bench", "--steps", str(steps), "--warmup", str(warmup), "--codegen", codegen,
                    "--threads", str(os.cpu_count() or 1)]


def run_reference(steps, warmup, grid=GRID_N):
    """The UNMODIFIED reference (oracle/_ref/ref_driver, built from /root/reference by oracle/Makefile.ref) on the host cores."""
    driver, cmd = reference_cmd(steps, warmup, grid)
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
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="stark_b200")
    ap.add_argument("--grid", type=int, default=GRID_N)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage-steps", type=int, default=4, help="extra (untimed) steps run with stage profiling on after the timed region; 0 = off")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the same window of time steps as the stark_b200 arm (same warm-up rule), bounded so that the run ends within minutes
        ref_steps = max(1, min(args.steps, 200))
        ref_warmup = max(args.warmup, 3)
        r = run_reference(ref_steps, ref_warmup, args.grid)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver has not been built (oracle/Makefile.ref)"}))
            return 0
        v = r["newton_it_per_s"]
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": ref_steps, "warmup": ref_warmup,
                "ms_per_step": 1e3 * r["wall_s"] / max(1, r["steps"]), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(1, args.grid), "newton_iterations": r["newton_iterations"],
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["threads"], "kind": "reference", "sample": f"{ref_steps} time steps of the same scene after {ref_warmup} warm-up steps (unmodified reference, all host threads)"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from stark_b200 import capi, scenes
    from stark_b200 import dist as sbdist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the stark_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    sc = scenes.Scene("tetdrop", n=args.grid, dt=0.01, drop=0.003, device=local_rank, stream=stream.cuda_stream)

    def barrier():
        sbdist.barrier(torch.device("cuda", local_rank))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        sc.step()
    t0 = sc.totals()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    wall0 = time.perf_counter()
    its = evals = cg = accepted = 0
    solve_gpu_ms = 0.0
    for _ in range(args.steps):
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
    if rank != 0:
        if world > 1:
            dist.barrier()   # rank 0 finishes its single-GPU diagnostics before the group is torn down
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: EnergyTetStrain P+grad+Hessian element evaluation, timed alone (CUDA events) ----
    ctx_handle = sc.lib.sbh_scene_context(sc.h)
    roof = None
    try:
        import ctypes as C
        lib = capi.load()
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
        peak = peaks["hbm_gbs"] if peaks else 6650.0
        E, g = C.c_double(), C.c_double()
        lib.sb_eval(C.c_void_p(ctx_handle), 2, C.byref(E), C.byref(g))
        pot = int(sc.lib.sbh_scene_potential(sc.h, b"EnergyTetStrain"))
        ms = C.c_double()
        lib.sb_profile_potential(C.c_void_p(ctx_handle), pot, 2, 3, C.byref(ms))       # warm-up
        lib.sb_profile_potential(C.c_void_p(ctx_handle), pot, 2, 20, C.byref(ms))
        n_tets = int(t1["tets"])
        achieved = TET_BYTES * n_tets / (ms.value * 1e-3) / 1e9
        roof = {"kernel": "EnergyTetStrain element evaluation (P + grad + dense 12x12 Hessian)", "bound": "hbm", "achieved": achieved, "peak": peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)", "unit": "GB/s", "frac": achieved / peak,
                "traffic": TET_DRAM_TRAFFIC_NCU if args.grid == GRID_N else None, "traffic_source": "ncu --set full, profiles/r1_ncu_full_tet_assembly.csv", "launch_ms": ms.value, "algorithmic_bytes_per_launch": TET_BYTES * n_tets}
    except Exception as e:   # the line must still print
        roof = {"error": repr(e)}

    # ---- per-stage breakdown (diagnostic, outside the timed region): extra steps with a stream sync at every stage boundary ----
    stages = None
    if args.stage_steps > 0:
        try:
            import ctypes as C
            lib = capi.load()
            lib.sb_profile_stages(C.c_void_p(ctx_handle), 1)
            it_s = 0
            for _ in range(args.stage_steps):
                it_s += int(sc.step()["newton_iterations"])
            rep = lib.sb_profile_report(C.c_void_p(ctx_handle)).decode()
            lib.sb_profile_stages(C.c_void_p(ctx_handle), 0)
            stages = {"steps": args.stage_steps, "newton_iterations": it_s}
            for ln in rep.splitlines():
                name, ms, calls = ln.split()
                stages[name] = {"ms": round(float(ms), 4), "calls": int(calls)}
        except Exception as e:
            stages = {"error": repr(e)}

    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample ----
    cpu = None
    if not args.no_cpu_baseline and args.gpus == 1:
        try:
            n_cpu = max(1, min(args.steps, 12))
            r = run_reference(n_cpu, max(args.warmup, 3), args.grid)
            if r is not None:
                cpu = {"value": r["newton_it_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                       "sample": f"{n_cpu} time steps of the same scene after {max(args.warmup, 3)} warm-up steps (unmodified reference, all host threads)",
                       "newton_iterations": r["newton_iterations"], "wall_s": r["wall_s"]}
        except Exception as e:
            cpu = {"error": repr(e)}

    steps_total = args.steps
    line = {
        "metric": METRIC, "value": its_all / (solve_gpu_ms * 1e-3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": e2e_ms / steps_total, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, args.grid),
        "e2e": {"value": its_all / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": (t1["h2d_bytes"] - t0["h2d_bytes"]) / steps_total, "d2h_bytes_per_step": (t1["d2h_bytes"] - t0["d2h_bytes"]) / steps_total},
        "gpu_launches": int(t1["launches"] - t0["launches"]),
        "newton_iterations": its_all, "evaluations": evals_all, "cg_iterations": cg_all, "accepted_steps_rank0": accepted, "wall_s_rank0": wall_s,
        "solve_gpu_ms_per_iteration": solve_gpu_ms / max(1.0, its_all / world),
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "stages": stages,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
