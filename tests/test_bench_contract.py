"""CPU-side check of bench.py's reference arm: the unmodified reference (oracle/_ref/ref_driver) runs a reduced grid on the
host cores and the line carries the keys the driver's contract names.  Skipped where the reference has not been built."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_driver")):
        pytest.skip("oracle/_ref/ref_driver not built (no /root/reference here)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3", "--grid", "3"],
                         capture_output=True, text=True, timeout=580)
    assert out.returncode == 0, out.stderr[-1500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "newton_iterations_per_second" and line["unit"] == "Newton it/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    assert line["config"]["grid"] == 3 and line["config"]["name"] == "C2" and line["config"]["scene"] == "tetdrop"
    assert line["accepted_steps"] is not None and 0 <= line["accepted_steps"] <= 2
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.timeout(600)
def test_reference_arm_runs_the_other_baseline_configurations():
    """`--config C5` (tet bar, no contact) through the same arm at a reduced grid: same contract keys."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_driver")):
        pytest.skip("oracle/_ref/ref_driver not built (no /root/reference here)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C5", "--steps", "2", "--warmup", "3", "--grid", "2"],
                         capture_output=True, text=True, timeout=580)
    assert out.returncode == 0, out.stderr[-1500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"]["name"] == "C5" and line["config"]["scene"] == "tetbar" and line["value"] > 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3", "--grid", "3"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.timeout(600)
def test_default_configuration_follows_the_gpu_count():
    """`--gpus 1` runs C2 (the 1xB200 configuration the metric is quoted on), `--gpus N > 1` the 1M-tet bar BASELINE.json names for
    the 1/2/4/8-GPU sweep (one scene shared by the ranks); both arms pick the same one."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_driver")):
        pytest.skip("oracle/_ref/ref_driver not built (no /root/reference here)")
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3", "--grid", "2"],
                         capture_output=True, text=True, timeout=580, env=env)
    assert out.returncode == 0, out.stderr[-1500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["config"]["name"] == "C5" and line["n_gpus"] == 2 and "slab decomposition" in line["config"]["parallelism"]
