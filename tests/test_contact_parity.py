"""GPU parity of collision detection and contact / friction list construction against the unmodified reference
(SURVEY.md 8(c) stage 4): the six proximity lists and the intersection list as SORTED SETS of integer tuples (the
reference's own order depends on its thread count), distances, the contact / friction connectivity tables, and the
global energy / gradient evaluated from the device-built tables."""
import numpy as np
import pytest

from golden_util import Golden, bind

KINDS = ["pt_pp", "pt_pe", "pt_pt", "ee_pp", "ee_pe", "ee_ee"]
FIXTURES = ["tetdrop_n3", "tetdrop_n5", "cloth_n8", "cloth_shells_n8", "boxes", "tetchain_n3", "zoo_n4", "zoo_slide_n4"]
# array roles in the fixtures (docs/potential_layouts.txt): a1 x0, a13 X, a9 dt, a59 t0, a60 q0 (w,x,y,z)
X0, XREST, DT, RB_T0, RB_Q0 = 1, 13, 9, 59, 60


def make(fixture, with_potentials):
    from stark_b200 import capi
    ctx = capi.Context(0)
    g = Golden(fixture)
    skip = [p["name"] for _, p in g.potentials(False) if p["name"].startswith(("contact_", "friction_"))]
    handles = bind(ctx, g, set(capi.kernel_names()) if with_potentials else set(), skip=skip)
    ids = [a["id"] for a in g.meta["arrays"]]
    dof = [ids.index(d) for d in g.meta["dof_array_ids"]]
    ctx.contact_init(soft_v1=dof[0], soft_x0=X0, soft_X=XREST, rb_v1=dof[1], rb_w1=dof[2], rb_t0=RB_T0, rb_q0=RB_Q0, dt=DT)
    groups = []
    for k, m in enumerate(g.meta["meshes"]):
        tri, edg, psi = g[f"mesh{k}_triangles"], g[f"mesh{k}_edges"], g[f"mesh{k}_ps_index"]
        if m["ps"] == 0:
            groups.append(ctx.contact_add_mesh(0, -1, psi, None, tri, edg, m["contact_thickness"]))
        else:
            local = g["rigidbody_local_vertices"][psi]
            groups.append(ctx.contact_add_mesh(1, m["idx_in_ps"], None, local, tri, edg, m["contact_thickness"]))
    for a, b in g.meta["blacklist"]:
        ctx.contact_blacklist(a, b)
    for a, b, mu in g.meta["friction_pairs"]:
        ctx.contact_set_friction(a, b, mu)
    ctx.contact_set_params(g.meta["contact_stiffness"], g.meta["friction_stick_slide_threshold"])
    return capi, ctx, g, handles


def as_set(ids):
    return sorted(map(tuple, ids.tolist()))


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", FIXTURES)
def test_vertices_and_proximity_lists(fixture):
    capi, ctx, g, _ = make(fixture, False)
    ctx.contact_detect(g.meta["proximity_enlargement"], True)
    # collision vertices rebuilt on the device from x0, v1, t0, q0, w1 (EnergyFrictionalContact::_update_vertices)
    for k, m in enumerate(g.meta["meshes"]):
        x = ctx.contact_get_vertices(k, m["n_vertices"])
        assert np.abs(x - g[f"mesh{k}_vertices"]).max() <= 1e-14 * max(1.0, np.abs(g[f"mesh{k}_vertices"]).max())
    for kind, name in enumerate(KINDS):
        ids, dist = ctx.contact_proximity(kind)
        ref_ids, ref_dist = g[f"prox_{name}_ids"], g[f"prox_{name}_dist"]
        assert as_set(ids) == as_set(ref_ids.reshape(-1, ids.shape[1]) if ref_ids.size else np.zeros((0, ids.shape[1]), int)), name
        if len(dist):
            o, ro = np.lexsort(ids.T[::-1]), np.lexsort(ref_ids.T[::-1])
            assert np.abs(dist[o] - ref_dist[ro]).max() <= 1e-12 * np.abs(ref_dist).max(), name
    ids, _ = ctx.contact_proximity(6)
    assert as_set(ids) == as_set(g["intersections"])
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", FIXTURES)
def test_detection_on_reference_vertices_is_bit_exact(fixture):
    """Same check with the reference's own vertex positions injected (isolates detection from the vertex update)."""
    capi, ctx, g, _ = make(fixture, False)
    for k, m in enumerate(g.meta["meshes"]):
        ctx.contact_set_vertices(k, g[f"mesh{k}_vertices"])
    ctx.contact_detect(g.meta["proximity_enlargement"], True)
    for kind, name in enumerate(KINDS):
        ids, dist = ctx.contact_proximity(kind)
        ref_ids = g[f"prox_{name}_ids"]
        assert as_set(ids) == as_set(ref_ids.reshape(-1, ids.shape[1]) if ref_ids.size else np.zeros((0, ids.shape[1]), int)), name
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", FIXTURES)
def test_contact_tables_and_global_evaluation(fixture):
    capi, ctx, g, handles = make(fixture, True)
    ctx.contact_update_friction()   # friction lists are built at dt = 0 at the start of the step
    ctx.contact_update()
    for i, p in g.potentials(False):
        if not p["name"].startswith(("contact_", "friction_")):
            continue
        h = ctx.contact_potential(p["name"])
        assert h >= 0, p["name"]
        n = ctx.potential_info(h)[2]
        assert n == p["n_elements"], (p["name"], n, p["n_elements"])
    E, res = ctx.eval("PGH")
    # friction data in the fixture was built by the reference at the start of ITS step (before v1 was perturbed), i.e. from
    # the same x0: energies must agree
    assert abs(E - g.meta["E"]) <= 1e-9 * abs(g.meta["E"])
    grad = ctx.grad()
    assert np.abs(grad - g["grad"]).max() <= 1e-9 * np.abs(g["grad"]).max()
    # per-table element outputs as sorted multisets of rows
    for i, p in g.potentials():
        if not p["name"].startswith(("contact_", "friction_")):
            continue
        out = ctx.element_output(ctx.contact_potential(p["name"]))
        ref = g[f"pot{i}_sol"]
        scale = np.abs(ref).max()
        # (rows are matched one to one by nearest row, not by sorting on the energy: symmetric pairs -- the same two vertices
        #  found from both sides -- have equal energies and mirrored gradients)
        assert out.shape == ref.shape, p["name"]
        used = np.zeros(len(out), dtype=bool)
        for r in ref:
            dist = np.abs(out - r).max(axis=1)
            dist[used] = np.inf
            j = int(np.argmin(dist))
            assert dist[j] <= 1e-8 * scale, (p["name"], dist[j] / scale)
            used[j] = True
    ctx.close()
