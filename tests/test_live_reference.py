"""Parity at BASELINE sizes against the UNMODIFIED reference run live on the GPU box's host cores (oracle/_ref/ref_driver
--dump, shipped with the snapshot; nothing here reads /root/reference).  One Newton iteration on the injected state of the
full-size scene: contact sets exact as sorted integer tuples (thousands of pairs through the tiled Morton broad phase against
the reference's octree), distances 1e-12, BCSR pattern exact, E and gradient 1e-10 relative to |g|_inf, every contact /
friction table's row count, and the block-Jacobi PCG solution.  SURVEY.md 8(c) stages 1, 2, 4 at C2 / C4 / C3 sizes."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from golden_util import Golden, bind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

KINDS = ["pt_pp", "pt_pe", "pt_pt", "ee_pp", "ee_pe", "ee_ee"]
X0, XREST, DT, RB_T0, RB_Q0 = 1, 13, 9, 59, 60   # array roles, as in test_contact_parity

CASES = {
    # name: (driver args, min number of proximity pairs expected -- the point of the test is a LARGE contact set)
    "C2_tetdrop_26": (["--scene", "tetdrop", "--n", "26", "--steps", "9"], 900),     # the whole bottom face (729 nodes) + its edges
    "C4_tetchain_16": (["--scene", "tetchain", "--n", "16", "--steps", "16"], 50),
    # cloth draped over the box corner: all six proximity types, ~950 pairs, and a state where the reference's PCG stops on
    # indefiniteness (pcg_converged = 0): the failure signal is compared as well
    "C3_cloth_shells_64": (["--scene", "cloth_shells", "--n", "64", "--steps", "40"], 200),   # (the reference's trajectory depends on its thread count: 945 pairs on 8 threads, fewer on other machines)
}


def live_dump(args):
    import make_golden
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/ref_driver has not been built")
    codegen = f"/tmp/stark_ref_codegen_{os.getuid()}"
    with tempfile.TemporaryDirectory() as d:
        cmd = [DRIVER] + args + ["--dump", d, "--codegen", codegen, "--threads", str(min(16, os.cpu_count() or 1))]
        subprocess.run(cmd, check=True, env=dict(os.environ, CXX="/usr/bin/g++"), stdout=subprocess.DEVNULL, timeout=1500)
        return Golden("live", data=make_golden.pack(d))


def as_set(ids):
    return sorted(map(tuple, ids.tolist()))


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_full_size_iteration_matches_live_reference(case):
    from stark_b200 import capi
    args, min_pairs = CASES[case]
    g = live_dump(args)
    ctx = capi.Context(0)
    skip = [p["name"] for _, p in g.potentials(False) if p["name"].startswith(("contact_", "friction_"))]
    handles = bind(ctx, g, set(capi.kernel_names()), skip=skip)
    ids = [a["id"] for a in g.meta["arrays"]]
    dof = [ids.index(d) for d in g.meta["dof_array_ids"]]
    ctx.contact_init(soft_v1=dof[0], soft_x0=X0, soft_X=XREST, rb_v1=dof[1], rb_w1=dof[2], rb_t0=RB_T0, rb_q0=RB_Q0, dt=DT)
    for k, m in enumerate(g.meta["meshes"]):
        tri, edg, psi = g[f"mesh{k}_triangles"], g[f"mesh{k}_edges"], g[f"mesh{k}_ps_index"]
        if m["ps"] == 0:
            ctx.contact_add_mesh(0, -1, psi, None, tri, edg, m["contact_thickness"])
        else:
            ctx.contact_add_mesh(1, m["idx_in_ps"], None, g["rigidbody_local_vertices"][psi], tri, edg, m["contact_thickness"])
    for a, b in g.meta["blacklist"]:
        ctx.contact_blacklist(a, b)
    for a, b, mu in g.meta["friction_pairs"]:
        ctx.contact_set_friction(a, b, mu)
    ctx.contact_set_params(g.meta["contact_stiffness"], g.meta["friction_stick_slide_threshold"])

    # ---- stage 4: proximity / intersection sets (device-side vertex update, tiled Morton broad phase, narrow phase) ----
    ctx.contact_detect(g.meta["proximity_enlargement"], True)
    total = 0
    for kind, name in enumerate(KINDS):
        out_ids, dist = ctx.contact_proximity(kind)
        ref_ids, ref_dist = g[f"prox_{name}_ids"], g[f"prox_{name}_dist"]
        total += len(ref_dist)
        assert as_set(out_ids) == as_set(ref_ids.reshape(-1, out_ids.shape[1]) if ref_ids.size else np.zeros((0, out_ids.shape[1]), int)), name
        if len(dist):
            o, ro = np.lexsort(out_ids.T[::-1]), np.lexsort(ref_ids.T[::-1])
            assert np.abs(dist[o] - ref_dist[ro]).max() <= 1e-12 * np.abs(ref_dist).max(), name
    assert total >= min_pairs, f"only {total} proximity pairs: the scene did not reach contact"
    out_ids, _ = ctx.contact_proximity(6)
    assert as_set(out_ids) == as_set(g["intersections"])

    # ---- stage 1: tables, energy, gradient ----
    ctx.contact_update_friction()
    ctx.contact_update()
    for i, p in g.potentials(False):
        if p["name"].startswith(("contact_", "friction_")):
            assert ctx.potential_info(ctx.contact_potential(p["name"]))[2] == p["n_elements"], p["name"]
    E, res = ctx.eval("PGH")
    grad = ctx.grad()
    gmax = np.abs(g["grad"]).max()
    assert np.abs(grad - g["grad"]).max() <= 1e-10 * gmax, np.abs(grad - g["grad"]).max() / gmax
    assert abs(res - g.meta["residual_inf"]) <= 1e-10 * g.meta["residual_inf"]
    assert abs(E - g.meta["E"]) <= 1e-10 * max(1.0, abs(g.meta["E"]))

    # ---- stage 2: BCSR pattern exact, values within float accumulation noise of the reference's ----
    ctx.assemble()
    rows, cols, vals = ctx.bcsr()
    assert np.array_equal(rows, g["bcsr_rows"])
    assert np.array_equal(cols, g["bcsr_cols"])
    ref = g["bcsr_vals"].astype(np.float64).reshape(-1, 9)
    ref_scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-7 * np.abs(ref).max())
    assert (np.abs(vals.astype(np.float64).reshape(-1, 9) - ref) / ref_scale).max() < 5e-5

    # ---- stage 3: the linear solve meets the reference's contract and lands on the reference's direction ----
    r = ctx.solve_pcg(g.meta["pcg_abs_tol"], g.meta["pcg_rel_tol"], 10000, True)
    assert r["ok"] == bool(g.meta["pcg_converged"])
    if r["ok"]:
        assert abs(r["iterations"] - g.meta["pcg_iterations"]) <= max(2, g.meta["pcg_iterations"] // 20), (r, g.meta["pcg_iterations"])
        du = ctx.du()
        assert r["du_dot_grad"] < 0
        assert np.abs(du - g["pcg_du"]).max() <= 2e-3 * np.abs(g["pcg_du"]).max()
    ctx.close()
