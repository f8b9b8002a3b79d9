#!/usr/bin/env python
"""Soak run: many time steps of a scene, reporting failures, device-memory growth and the iteration rate per window.
    python tools/soak.py [scene] [grid] [steps]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stark_b200 import scenes

scene = sys.argv[1] if len(sys.argv) > 1 else "tetdrop"
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 26
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
sc = scenes.Scene(scene, n=grid)
free0 = None
win = []
t0 = time.time(); its = 0; rejected = 0
for i in range(steps):
    s = sc.step()
    its += int(s["newton_iterations"]); rejected += 0 if s["accepted"] else 1
    if not s["keep_going"]:
        print(json.dumps({"stopped_at": i, "result": s["result"]})); break
    if i == 20: free0 = torch.cuda.mem_get_info()[0]
    if (i + 1) % 50 == 0:
        dt = time.time() - t0
        win.append({"steps": i + 1, "time": round(s["time"], 3), "its": its, "it_per_s": round(its / dt, 1), "rejected": rejected,
                    "mem_growth_mb": round((free0 - torch.cuda.mem_get_info()[0]) / 2**20, 1) if free0 else None})
        t0 = time.time(); its = 0
print(json.dumps(win))
