"""GPU / host timeline of Newton solves (SB_TIMELINE): python tools/timeline.py <scene> <n> <skip> <steps> [min_iterations]
Diagnostic only -- prints, per stage of sb_newton_solve, when the host issued it and when the GPU ran it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
scene, n, skip, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
os.environ["SB_TIMELINE"] = sys.argv[5] if len(sys.argv) > 5 else "1"
from stark_b200 import scenes  # noqa: E402

sc = scenes.Scene(scene, n=n)
for i in range(skip + steps):
    if i == skip:
        sys.stderr.write("TIMELINE ==== timed window begins ====\n")
    s = sc.step()
    sys.stderr.write(f"TIMELINE step {i}: iterations {int(s['newton_iterations'])} cg {int(s['cg_iterations'])} gpu_ms {s['solve_gpu_ms']:.3f} accepted {int(s['accepted'])}\n")
