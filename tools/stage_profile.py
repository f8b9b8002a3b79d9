#!/usr/bin/env python
"""Diagnostic: per-stage breakdown (sb_profile_stages) of a window of time steps of a bench scene.
    python tools/stage_profile.py [scene] [grid] [skip_steps] [steps]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stark_b200 import capi, scenes  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "tetdrop"
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 26
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 3
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
sc = scenes.Scene(scene, n=grid, dt=0.01, drop=0.003)
lib = capi.load()
ctx = C.c_void_p(sc.lib.sbh_scene_context(sc.h))
for _ in range(skip):
    sc.step()
lib.sb_profile_stages(ctx, 1)
its = 0
per_step = []
for _ in range(steps):
    s = sc.step()
    its += int(s["newton_iterations"])
    per_step.append((int(s["newton_iterations"]), int(s["cg_iterations"]), int(s["evaluations"]), round(s["solve_gpu_ms"], 3), int(s["accepted"])))
rep = lib.sb_profile_report(ctx).decode()
out = {"scene": scene, "grid": grid, "skip": skip, "steps": steps, "newton_iterations": its, "per_step(it,cg,evals,solve_ms,accepted)": per_step}
for ln in rep.splitlines():
    name, ms, calls = ln.split()
    out[name] = {"ms": round(float(ms), 3), "calls": int(calls)}
print(json.dumps(out))
