"""Generate the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/ref_driver).

Runs only in the build container (needs /root/reference to have been compiled by oracle/Makefile.ref).
Each fixture is one Newton iteration on a deterministic scene: every potential's connectivity and bound
arrays, per-element [E | grad | hess] outputs, global E/grad, BCSR, PCG solution, proximity and
intersection lists (SURVEY.md section 8(c) parity protocol).

    python tests/golden/make_golden.py            # regenerate all fixtures
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
CODEGEN = os.path.join(ROOT, "oracle", "_ref", "codegen")

FIXTURES = {
    # name: driver args
    "tetdrop_n3": ["--scene", "tetdrop", "--n", "3", "--steps", "4"],
    "tetdrop_n5": ["--scene", "tetdrop", "--n", "5", "--steps", "5"],
    "tetbar_n2": ["--scene", "tetbar", "--n", "2", "--nz", "10", "--steps", "3"],
    "cloth_n8": ["--scene", "cloth", "--n", "8", "--steps", "12", "--dt", "0.01"],
    "cloth_shells_n8": ["--scene", "cloth_shells", "--n", "8", "--steps", "12", "--dt", "0.01"],
    # rigid-rigid contact + friction (all six + four tables), linear / angular velocity controllers
    "boxes": ["--scene", "boxes", "--n", "1", "--steps", "4", "--vz", "0.5"],
    # C4 at small size: tet cube under a chain of 4 hinged boxes (rb_constraint points + directions, rb_d contact)
    "tetchain_n3": ["--scene", "tetchain", "--n", "3", "--ny", "4", "--steps", "14"],
    # the five attachment potentials
    "attach_n6": ["--scene", "attach", "--n", "6", "--steps", "3"],
    # all ten deformable-deformable contact / friction tables (two-mesh and self contact), the point-side rigid-deformable
    # ones, the three _Elasticity_Only strain potentials and both rod potentials; state injected at v1 ~ 0 (contacts of x0)
    "zoo_n4": ["--scene", "zoo", "--n", "4", "--steps", "1", "--inject", "0"],
    # same scene, state injected at v1 = v0 (bodies separating and sliding: friction beyond the stick threshold)
    "zoo_slide_n4": ["--scene", "zoo", "--n", "4", "--steps", "1"],
    # slider, distance, distance-limit, angle-limit and damped-spring rigid-body constraints, all violated
    "joints": ["--scene", "joints", "--n", "1", "--steps", "6"],
    # a USER potential (examples/main.cpp:666-690 EnergyMagneticAttraction, added with a lambda through add_potential): besides
    # the usual dump, the symx operation sequences of [E] and [E | grad | hess] the reference's code generator prints
    "magnet_n2": ["--scene", "magnet", "--n", "2", "--steps", "3", "--ops", "EnergyMagneticAttraction"],
}


def rd(d, name, dtype):
    p = os.path.join(d, name + ".bin")
    if not os.path.exists(p):
        return np.zeros(0, dtype=dtype)
    return np.fromfile(p, dtype=dtype)


def pack(d):
    meta = json.load(open(os.path.join(d, "meta.json")))
    out = {}
    out["dofs"] = rd(d, "dofs", np.float64)
    out["grad"] = rd(d, "grad", np.float64)
    for i, a in enumerate(meta["arrays"]):
        out[f"array{i}"] = rd(d, f"array{i}", np.float64).reshape(a["n_elements"], a["stride"])
    for p, pot in enumerate(meta["potentials"]):
        if pot["n_elements"] == 0:
            continue
        out[f"pot{p}_conn"] = rd(d, f"pot{p}_conn", np.int32).reshape(pot["n_elements"], pot["conn_stride"])
        out[f"pot{p}_sol"] = rd(d, f"pot{p}_sol", np.float64).reshape(pot["n_elements"], pot["n_out"])
        out[f"pot{p}_active"] = rd(d, f"pot{p}_active", np.uint8)
        if pot.get("user_ops"):
            for tag in ("p", "pgh"):
                out[f"pot{p}_ops_{tag}"] = rd(d, f"pot{p}_ops_{tag}", np.int32).reshape(-1, 5)
                out[f"pot{p}_opsc_{tag}"] = rd(d, f"pot{p}_opsc_{tag}", np.float64)
            out[f"pot{p}_block_slots"] = rd(d, f"pot{p}_block_slots", np.int32)
    out["bcsr_rows"] = rd(d, "bcsr_rows", np.uint64).astype(np.int64)
    out["bcsr_cols"] = rd(d, "bcsr_cols", np.int32)
    out["bcsr_vals"] = rd(d, "bcsr_vals", np.float32)
    out["pcg_du"] = rd(d, "pcg_du", np.float64)
    out["projected_hessians"] = rd(d, "projected_hessians", np.float64)
    out["projected_sizes"] = rd(d, "projected_sizes", np.int32)
    out["element_block_rows"] = rd(d, "element_block_rows", np.int32)
    for g, mesh in enumerate(meta.get("meshes", [])):
        out[f"mesh{g}_vertices"] = rd(d, f"mesh{g}_vertices", np.float64).reshape(-1, 3)
        out[f"mesh{g}_triangles"] = rd(d, f"mesh{g}_triangles", np.int32).reshape(-1, 3)
        out[f"mesh{g}_edges"] = rd(d, f"mesh{g}_edges", np.int32).reshape(-1, 2)
        out[f"mesh{g}_ps_index"] = rd(d, f"mesh{g}_ps_index", np.int32)
    if "meshes" in meta:
        out["rigidbody_local_vertices"] = rd(d, "rigidbody_local_vertices", np.float64).reshape(-1, 3)
    for k in ["pt_pp", "pt_pe", "pt_pt", "ee_pp", "ee_pe", "ee_ee"]:
        if f"prox_{k}_n" in meta:
            w = meta[f"prox_{k}_width"]
            ids = rd(d, f"prox_{k}_ids", np.int32)
            out[f"prox_{k}_ids"] = ids.reshape(-1, w) if w else ids.reshape(0, 0)
            out[f"prox_{k}_dist"] = rd(d, f"prox_{k}_dist", np.float64)
    if "n_intersections" in meta:
        out["intersections"] = rd(d, "intersections", np.int32).reshape(-1, 4)
    res = os.path.join(d, "next_step_residuals.txt")
    if os.path.exists(res):
        out["next_step_residuals"] = np.loadtxt(res, ndmin=1)
        meta["next_step_stats"] = json.load(open(os.path.join(d, "next_step_stats.json")))
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def main():
    env = dict(os.environ, CXX="/usr/bin/g++")
    names = sys.argv[1:] or list(FIXTURES)
    for name in names:
        args = FIXTURES[name]
        with tempfile.TemporaryDirectory() as d:
            cmd = [DRIVER] + args + ["--dump", d, "--codegen", CODEGEN, "--threads", "4"]
            print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL)
            out = pack(d)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(f"  -> {path} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
