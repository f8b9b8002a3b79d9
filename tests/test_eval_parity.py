"""GPU parity of the element evaluation kernels against the unmodified reference (golden fixtures):
per-element [E | grad | hess] of every potential, global energy and gradient (SURVEY.md 8(c) stage 1)."""
import numpy as np
import pytest

from golden_util import Golden, bind

FIXTURES = ["tetdrop_n3", "tetdrop_n5", "tetbar_n2", "cloth_n8", "cloth_shells_n8", "boxes", "attach_n6", "tetchain_n3", "zoo_n4", "zoo_slide_n4", "joints", "magnet_n2"]
RTOL = 1e-10  # north_star: 1e-10 relative on gradient / residual


def _ctx():
    from stark_b200 import capi
    return capi, capi.Context(0)


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", FIXTURES)
def test_element_outputs_match_reference(fixture):
    capi, ctx = _ctx()
    g = Golden(fixture)
    handles = bind(ctx, g, set(capi.kernel_names()))
    missing = [p["name"] for i, p in g.potentials() if i not in handles]
    assert not missing, f"no kernel for {missing}"
    E, res = ctx.eval("PGH")
    for i, h in handles.items():
        p = g.meta["potentials"][i]
        ref = g[f"pot{i}_sol"][g[f"pot{i}_active"].astype(bool)]
        out = ctx.element_output(h)
        n = p["n_dofs"]
        assert out.shape == ref.shape, p["name"]
        # energy
        # (bending energies of a nearly flat sheet are ~1e-14 and dominated by cancellation in both implementations: the
        #  absolute floor is RTOL of the scene's energy per element, far below anything the solver can see)
        e_floor = RTOL * max(np.abs(ref[:, 0]).max(), abs(g.meta["E"]) / max(1, len(ref)))
        np.testing.assert_allclose(out[:, 0], ref[:, 0], rtol=RTOL, atol=e_floor, err_msg=p["name"] + " E")
        # gradient: relative to the element's own gradient scale
        # (floor: elements whose whole gradient is below 1e-6 of the global gradient are compared on that scale)
        gs = np.maximum(np.abs(ref[:, 1:1 + n]).max(axis=1, keepdims=True), 1e-6 * np.abs(g["grad"]).max()) + 1e-300
        # The dihedral angle is acos((1 - 1e-12) n0.n1): on a nearly flat cloth d(acos)/dx ~ 1/sqrt(2e-12) amplifies the
        # last-bit differences between two compilers ~1e6 times, so two correct evaluations of ONE hinge agree only to
        # ~1e-9 of that hinge's (tiny) gradient; the global gradient below is still held to 1e-10.
        tol = 1e-8 if p["name"] == "EnergyDiscreteShells" else RTOL
        assert (np.abs(out[:, 1:1 + n] - ref[:, 1:1 + n]) / gs).max() < tol, p["name"] + " grad"
        # Hessian: relative to the element Hessian's Frobenius norm
        hs = np.linalg.norm(ref[:, 1 + n:], axis=1, keepdims=True) + 1e-300
        assert (np.abs(out[:, 1 + n:] - ref[:, 1 + n:]) / hs).max() < tol, p["name"] + " hess"
    # global energy and gradient
    assert abs(E - g.meta["E"]) <= RTOL * max(1.0, abs(g.meta["E"]))
    grad = ctx.grad()
    ref_grad = g["grad"]
    assert np.abs(grad - ref_grad).max() <= RTOL * np.abs(ref_grad).max()
    assert abs(res - g.meta["residual_inf"]) <= RTOL * g.meta["residual_inf"]
    # energy-only evaluation (Armijo path)
    E_only = ctx.eval("P")
    assert abs(E_only - g.meta["E_only"]) <= RTOL * max(1.0, abs(g.meta["E_only"]))
    ctx.close()


@pytest.mark.gpu
def test_block_rows_match_reference():
    capi, ctx = _ctx()
    g = Golden("tetdrop_n3")
    handles = bind(ctx, g, set(capi.kernel_names()))
    ctx.eval("PGH")
    offs = g.meta["dof_offsets"]
    for i, h in handles.items():
        p = g.meta["potentials"][i]
        conn = g[f"pot{i}_conn"][g[f"pot{i}_active"].astype(bool)]
        expect = np.stack([offs[s] // 3 + conn[:, c] for s, c in p["dof_in_conn"]], axis=1)
        np.testing.assert_array_equal(ctx.block_rows(h), expect, err_msg=p["name"])
    ctx.close()


@pytest.mark.gpu
def test_tet_ad_and_analytic_kernels_agree():
    capi, ctx = _ctx()
    g = Golden("tetdrop_n5")
    h1 = bind(ctx, g, set(capi.kernel_names()), skip=())
    ctx.eval("PGH")
    idx = [i for i, p in g.potentials() if p["name"] == "EnergyTetStrain"][0]
    a = ctx.element_output(h1[idx])
    ctx2 = capi.Context(0)
    h2 = bind(ctx2, g, set(capi.kernel_names()), rename={"EnergyTetStrain": "EnergyTetStrain_AD"})
    ctx2.eval("PGH")
    b = ctx2.element_output(h2[idx])
    scale = np.linalg.norm(b[:, 13:], axis=1, keepdims=True)
    assert (np.abs(a[:, 13:] - b[:, 13:]) / scale).max() < 1e-11
    ctx.close()
    ctx2.close()


def _tet_arrays(g, idx):
    """array index bound at each first_symbol of the tet potential"""
    return {m["first_symbol"]: m["array"] for m in g.meta["potentials"][idx]["maps"]}


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["damping", "limit", "damping+limit", "split_fetch"])
def test_tet_analytic_rare_terms_match_ad(case):
    """Strain-rate damping and strain-limit terms (EnergyTetStrain.cpp:62-72) are zero in the golden scenes; switch them
    on, deform the mesh, and hold the hand-derived kernel to the AD kernel (itself pinned to the reference above).
    'split_fetch' binds one node's x0 through a separate buffer, which takes the kernel's generic gather path."""
    capi, _ = _ctx()
    g = Golden("tetdrop_n5")
    idx = [i for i, p in g.potentials() if p["name"] == "EnergyTetStrain"][0]
    sym = _tet_arrays(g, idx)
    rng = np.random.default_rng(7)
    override = {}
    v1 = np.array(g[f"array{sym[0]}"], dtype=np.float64, copy=True)
    override[sym[0]] = v1 + rng.uniform(-3.0, 3.0, v1.shape)          # dt = 0.01 -> a few % of strain
    if "damping" in case:
        override[sym[41]] = np.full_like(g[f"array{sym[41]}"], 0.37)
    if "limit" in case:
        override[sym[39]] = np.full_like(g[f"array{sym[39]}"], -0.02)  # below the rest state: active in every element
        override[sym[40]] = np.full_like(g[f"array{sym[40]}"], 5.0e3)
    split = ("EnergyTetStrain", 18) if case == "split_fetch" else None
    out = []
    for rename in (None, {"EnergyTetStrain": "EnergyTetStrain_AD"}):
        ctx = capi.Context(0)
        h = bind(ctx, g, set(capi.kernel_names()), rename=rename, override=override, split_fetch=split if rename is None else None)
        E, res = ctx.eval("PGH")
        out.append((E, ctx.grad(), ctx.element_output(h[idx]), ctx.eval("P")))
        ctx.close()
    (E1, g1, a, P1), (E2, g2, b, P2) = out
    assert abs(E1 - E2) <= 1e-12 * abs(E2)
    # the energy-only kernel of the line search (k_tet_energy) against the AD energy, and against its own P+G+H kernel
    assert abs(P1 - P2) <= 1e-12 * abs(P2)
    assert abs(P1 - E1) <= 1e-13 * abs(E1)
    assert np.abs(g1 - g2).max() <= 1e-11 * np.abs(g2).max()
    np.testing.assert_allclose(a[:, 0], b[:, 0], rtol=1e-11, atol=1e-11 * np.abs(b[:, 0]).max())
    gs = np.abs(b[:, 1:13]).max(axis=1, keepdims=True)
    assert (np.abs(a[:, 1:13] - b[:, 1:13]) / gs).max() < 1e-11
    hs = np.linalg.norm(b[:, 13:], axis=1, keepdims=True)
    assert (np.abs(a[:, 13:] - b[:, 13:]) / hs).max() < 1e-11
