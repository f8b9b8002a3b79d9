"""GPU test of the distributed block-Jacobi PCG (stark_b200/csrc/pcg.cu, DIST instances): W "ranks" as W contexts / threads on ONE
GPU (SB_PCG_GRID shrinks the persistent grids so that they are co-resident), peer buffers connected by pointer -- the same kernel,
barrier, halo pushes and partial-sum exchange that run between GPUs over NVLink (tools/dist_selftest.py --multiprocess under
torchrun is the multi-GPU form; profiles/ holds its runs).  Runs in a subprocess so that a hang cannot take the suite with it."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, fixture, extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dist_selftest.py"), "--world", str(world), "--fixture", fixture],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.stdout.strip(), out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0 and r["ok"], (r, out.stderr[-2000:])
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("world,fixture", [(2, "tetdrop_n5"), (4, "tetdrop_n5"), (2, "tetchain_n3"), (3, "cloth_shells_n8")])
def test_distributed_pcg_matches_single_gpu(world, fixture):
    """Same system, same stopping rule: the distributed solve returns the single-GPU direction (to CG tolerance; the partials are
    summed over a different partition), bit-identical on every rank and from solve to solve."""
    r = _run(world, fixture)
    assert r["identical_across_ranks"] and r["identical_across_solves"] and r["rel_err"] < 1e-6
    assert r["rank0_gradient_everywhere"]   # an evaluation on connected contexts ends with rank 0's gradient and energy on every rank


@pytest.mark.gpu
def test_distributed_pcg_streaming_slices():
    """The same with a shared-memory budget too small for the slices (the path of million-tet scenes on few GPUs)."""
    r = _run(2, "tetdrop_n5", {"SB_PCG_SMEM_LIMIT": "4096"})
    assert r["identical_across_ranks"]


@pytest.mark.gpu
def test_policy_keeps_resident_solves_local_and_adopts_rank0_du():
    """SB_DIST_POLICY=auto: a matrix that is resident in one GPU's shared memory is solved locally by every rank (no cross-GPU
    barrier is executed), and every rank then adopts rank 0's du (dist_bcast_from_root), so the replicas stay identical."""
    r = _run(2, "tetdrop_n5", {"SB_DIST_POLICY": "auto"})
    assert r["barriers"] == 0 and r["identical_across_ranks"] and r["rank0_gradient_everywhere"]
