"""The compiled drop-in (integration/symx_newton_shim.cpp): the reference's unmodified sources with symx::NewtonsMethod replaced
by the same class over the stark_b200 C-ABI (built by integration/Makefile.shim into oracle/_ref/shim/, shipped with the snapshot).
  * the reference's OWN Catch2 suite (tests/rb_constraints.cpp: 13 cases, DirectLLT, steady-state constraint forces to 1e-3)
    runs on the GPU path;
  * scenes built through the reference's public API (oracle/ref_driver.cpp) are stepped by both builds and their trajectories
    compared step by step (iteration counts +-1, positions to the tolerance of the inexact linear solves)."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "_ref", "shim")
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
CODEGEN = f"/tmp/stark_ref_codegen_{os.getuid()}"


def need(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, ROOT)} has not been built (integration/Makefile.shim needs /root/reference)")
    return path


@pytest.mark.gpu
def test_reference_catch2_suite_passes_on_the_gpu_path():
    exe = need(os.path.join(SHIM, "stark_tests_shim"))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=1500, cwd=tempfile.gettempdir())
    tail = out.stdout[-1500:] + out.stderr[-500:]
    assert out.returncode == 0, tail
    assert "All tests passed" in out.stdout and "13 test cases" in out.stdout, tail


def trace(exe, args, steps):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "trace.jsonl")
        cmd = [exe] + args + ["--steps", str(steps), "--trace", path, "--codegen", CODEGEN, "--threads", str(min(8, os.cpu_count() or 1))]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=dict(os.environ, CXX="/usr/bin/g++"))
        assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-1000:]
        return [json.loads(l) for l in open(path)]


@pytest.mark.gpu
@pytest.mark.parametrize("scene,args,steps", [
    ("joints", ["--scene", "joints", "--n", "1"], 12),                         # six rigid-body constraint types, no contact: no host callbacks inside the solve
    ("tetbar", ["--scene", "tetbar", "--n", "2", "--nz", "10"], 6),            # volume elements + scripted prescribed positions
    ("tetdrop", ["--scene", "tetdrop", "--n", "4", "--vz", "0.25"], 14),        # IPC contact + friction: the reference's host-side contact callbacks run around every evaluation
    ("cloth", ["--scene", "cloth", "--n", "8"], 14),                           # triangle strain + bending over a scripted rigid box
    ("magnet", ["--scene", "magnet", "--n", "3", "--vz", "0.25"], 12),         # a USER potential (add_potential with a lambda): kernel generated from its symx sequence by NVRTC
])
def test_scene_trajectories_match_the_unmodified_reference(scene, args, steps):
    ref = trace(need(REF_DRIVER), args, steps)
    gpu = trace(need(os.path.join(SHIM, "ref_driver_shim")), args, steps)
    assert len(ref) == len(gpu) == steps
    for a, b in zip(ref, gpu):
        assert a["accepted"] == b["accepted"] and abs(a["time"] - b["time"]) < 1e-12, (a, b)
        assert abs(a["newton_iterations"] - b["newton_iterations"]) <= 1, (a, b)
        assert a["ls_inv"] == b["ls_inv"], (a, b)
        scale = max(1.0, abs(a["sum_x2"]))
        assert abs(a["sum_x2"] - b["sum_x2"]) <= 1e-6 * scale, (a, b)
        assert abs(a["max_abs_x"] - b["max_abs_x"]) <= 1e-6 * max(1.0, a["max_abs_x"]), (a, b)
        assert abs(a["rigid_t"] - b["rigid_t"]) <= 1e-6 * max(1.0, abs(a["rigid_t"])), (a, b)
        assert abs(a["rigid_q"] - b["rigid_q"]) <= 1e-6 * max(1.0, abs(a["rigid_q"])), (a, b)
    assert sum(s["newton_iterations"] for s in gpu) > 0
