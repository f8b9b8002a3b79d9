"""GPU parity of assembly, PD projection, block-Jacobi PCG and the Newton driver (SURVEY.md 8(c) stages 2, 3, 5):
CUDA path (through the C-ABI) vs the numpy oracle on the same inputs and vs the reference's golden outputs."""
import os
import sys

import numpy as np
import pytest

from golden_util import Golden, bind

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle  # noqa: E402

FIXTURES = ["tetdrop_n3", "tetdrop_n5", "tetbar_n2", "cloth_n8", "cloth_shells_n8", "boxes", "attach_n6", "tetchain_n3", "zoo_n4", "joints", "magnet_n2"]


def setup(fixture):
    from stark_b200 import capi
    ctx = capi.Context(0)
    g = Golden(fixture)
    handles = bind(ctx, g, set(capi.kernel_names()))
    return capi, ctx, g, handles


def gpu_elements(ctx, handles):
    Hs, rows = [], []
    for i in sorted(handles):
        H = ctx.hessians(handles[i])
        r = ctx.block_rows(handles[i])
        Hs += list(H)
        rows += [list(x) for x in r]
    return Hs, rows


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", FIXTURES)
def test_assembly(fixture):
    capi, ctx, g, handles = setup(fixture)
    ctx.eval("PGH")
    ctx.assemble()
    rows, cols, vals = ctx.bcsr()
    # pattern: bit-exact against the reference
    assert np.array_equal(rows, g["bcsr_rows"])
    assert np.array_equal(cols, g["bcsr_cols"])
    # values: bit-exact against the oracle's float64-accumulated, float-rounded sum of the SAME element Hessians
    Hs, erows = gpu_elements(ctx, handles)
    rp, oc, ov = oracle.assemble_bcsr(Hs, erows, len(rows) - 1)
    assert np.array_equal(rp, rows) and np.array_equal(oc, cols)
    diff = np.abs(ov.astype(np.float64) - vals.astype(np.float64)).reshape(-1, 9)
    scale = np.abs(ov.astype(np.float64)).reshape(-1, 9).max(axis=1, keepdims=True) + 1e-30
    assert (diff / scale).max() < 2e-7   # one float ulp (summation order inside float64 may differ)
    # and within float-accumulation noise of the reference's values
    ref = g["bcsr_vals"].astype(np.float64).reshape(-1, 9)
    # (blocks whose contributions cancel -- e.g. the coupling of two hinged bodies -- are compared on the matrix scale)
    ref_scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-7 * np.abs(ref).max())
    assert (np.abs(vals.astype(np.float64).reshape(-1, 9) - ref) / ref_scale).max() < 5e-5
    # second assembly with an unchanged pattern reuses the symbolic phase and gives identical values
    ctx.assemble()
    _, _, vals2 = ctx.bcsr()
    assert np.array_equal(vals, vals2)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", FIXTURES)
def test_pcg(fixture):
    capi, ctx, g, handles = setup(fixture)
    ctx.eval("PGH")
    ctx.assemble()
    rows, cols, vals = ctx.bcsr()
    grad = ctx.grad()
    out = ctx.solve_pcg(g.meta["pcg_abs_tol"], g.meta["pcg_rel_tol"], 10000, True)
    du = ctx.du()
    # oracle on the same matrix / rhs
    x, it, ok = oracle.solve_pcg(rows, cols, vals, -grad, g.meta["pcg_abs_tol"], g.meta["pcg_rel_tol"], 10000)
    assert out["ok"] == ok == bool(g.meta["pcg_converged"])
    assert out["iterations"] == it
    assert abs(out["iterations"] - g.meta["pcg_iterations"]) <= 1
    # (joints: six stiff 1e6 N/m constraints on 42 DoFs -- the float-stored matrix is ill-conditioned, two correct float64 PCG
    #  runs with different summation orders agree to 2e-7)
    assert np.abs(du - x).max() <= (2e-7 if fixture == "joints" else 1e-7) * np.abs(x).max()
    if out["iterations"] == g.meta["pcg_iterations"]:
        assert np.abs(du - g["pcg_du"]).max() <= 1e-4 * np.abs(g["pcg_du"]).max()
    if not ok:
        # the indefiniteness signal (solve_pcg.h:183-192: p^T A p <= 0 stops the solve, x is returned as it is) -- magnet_n2: the
        # attraction's Hessian is indefinite and nothing is projected here; same verdict at the same iteration as the reference
        assert fixture == "magnet_n2" and out["iterations"] == g.meta["pcg_iterations"]
        ctx.close()
        return
    # contract of the inexact solve: residual below the forcing tolerance, descent direction
    r = oracle.bcsr_spmv(rows, cols, vals, du) + grad
    assert np.linalg.norm(r) / np.linalg.norm(grad) < max(g.meta["pcg_abs_tol"], g.meta["pcg_rel_tol"]) * 1.0000001
    assert out["du_dot_grad"] < 0
    assert abs(out["du_dot_grad"] - du @ grad) <= 1e-12 * abs(du @ grad)
    assert abs(out["du_inf"] - np.abs(du).max()) == 0.0
    assert abs(out["du_dot_grad"] - g.meta["du_dot_grad"]) <= 1e-4 * abs(g.meta["du_dot_grad"])
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["tetdrop_n3", "cloth_n8", "cloth_shells_n8", "boxes"])
def test_projection_all(fixture):
    capi, ctx, g, handles = setup(fixture)
    ctx.eval("PGH")
    before, _ = gpu_elements(ctx, handles)
    n_proj, n_hess, all_proj = ctx.project_to_pd(0.0)
    assert n_proj == n_hess == len(before) and all_proj
    after, _ = gpu_elements(ctx, handles)
    n_changed = 0
    for H, P in zip(before, after):
        ref, changed = oracle.project_to_pd(H)
        n_changed += changed
        assert np.abs(P - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
        if changed:
            assert np.linalg.eigvalsh(0.5 * (P + P.T)).min() > -1e-9 * np.abs(P).max()
    assert n_changed > 0
    # projecting again is idempotent (everything is flagged as projected)
    n_proj2, _, _ = ctx.project_to_pd(0.0)
    assert n_proj2 == n_proj
    ctx.close()


@pytest.mark.gpu
def test_projection_selective_ppn():
    capi, ctx, g, handles = setup("tetdrop_n3")
    E, res = ctx.eval("PGH")
    grad = ctx.grad()
    thr = 0.5 * res
    before, rows = gpu_elements(ctx, handles)
    n_proj, n_hess, all_proj = ctx.project_to_pd(thr)
    active = np.abs(grad.reshape(-1, 3)).max(axis=1) >= thr
    expect = sum(1 for r in rows if active[r].any())
    assert n_proj == expect and not all_proj and 0 < n_proj < n_hess
    after, _ = gpu_elements(ctx, handles)
    for H, P, r in zip(before, after, rows):
        if active[r].any():
            ref, _ = oracle.project_to_pd(H)
            assert np.abs(P - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
        else:
            assert np.array_equal(H, P)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["tetbar_n2", "tetdrop_n3"])
def test_newton_converges_from_injected_state(fixture):
    """Newton driver without the contact callbacks (fixed contact lists): residual history starts at the reference's
    residual for this state and drops below the tolerance with a monotone energy decrease."""
    capi, ctx, g, handles = setup(fixture)
    s = ctx.newton_default_settings()
    s.contact_enabled = 0
    s.max_iterations = 50
    E0 = ctx.eval("P")
    st = ctx.newton_solve(s)
    assert st.result == 0, st.result
    assert abs(st.residuals[0] - g.meta["residual_inf"]) <= 1e-10 * g.meta["residual_inf"]
    assert st.last_residual < 1e-6 or st.n_evaluations > 1
    assert st.last_energy < E0
    assert 1 <= st.newton_iterations <= 20
    ctx.close()


def _dense(rows, cols, vals):
    n = 3 * (len(rows) - 1)
    A = np.zeros((n, n))
    v = vals.astype(np.float64).reshape(-1, 3, 3)   # column-major inside the block
    for br in range(len(rows) - 1):
        for j in range(rows[br], rows[br + 1]):
            A[3 * br:3 * br + 3, cols[j]:cols[j] + 3] = v[j].T
    return A


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["tetdrop_n3", "tetdrop_n5", "tetbar_n2", "cloth_n8", "cloth_shells_n8", "attach_n6", "tetchain_n3", "boxes", "joints"])
def test_direct_llt(fixture):
    """DirectLLT branch (NewtonsMethod.cpp:395-418): the sparse tile-envelope Cholesky solves the assembled (float-stored) matrix
    like a FP64 direct solver does; an indefinite matrix is reported as a failed solve."""
    capi, ctx, g, handles = setup(fixture)
    ctx.eval("PGH")
    ctx.project_to_pd(0.0)          # every element PD -> the assembled matrix is PD
    ctx.assemble()
    rows, cols, vals = ctx.bcsr()
    grad = ctx.grad()
    out = ctx.solve_llt()
    assert out["ok"]
    du = ctx.du()
    A = _dense(rows, cols, vals)
    A = np.tril(A) + np.tril(A, -1).T   # the factorisation reads the lower triangle
    x = np.linalg.solve(A, -grad)
    assert np.abs(du - x).max() <= 1e-9 * np.abs(x).max()
    assert abs(out["du_dot_grad"] - du @ grad) <= 1e-12 * abs(du @ grad)
    assert out["du_inf"] == np.abs(du).max()
    assert out["du_dot_grad"] < 0
    # a second factorisation of the same pattern reuses ordering + envelope (no host analysis) and gives the same result
    st0 = ctx.llt_stats()
    assert st0["orderings"] == 1 and st0["analyses"] == 1 and st0["tiles"] <= st0["tile_rows"] * (st0["tile_rows"] + 1) // 2
    out2 = ctx.solve_llt()
    st1 = ctx.llt_stats()
    assert out2["ok"] and st1["orderings"] == 1 and st1["analyses"] == 1
    assert np.array_equal(ctx.du(), du)
    ctx.close()


@pytest.mark.gpu
def test_direct_llt_reports_indefinite_matrix():
    capi, ctx, g, handles = setup("tetdrop_n3")   # unprojected barrier Hessians: lambda_min = -4.85
    ctx.eval("PGH")
    ctx.assemble()
    rows, cols, vals = ctx.bcsr()
    A = _dense(rows, cols, vals)
    A = np.tril(A) + np.tril(A, -1).T
    assert np.linalg.eigvalsh(A).min() < 0
    assert not ctx.solve_llt()["ok"]
    ctx.close()


@pytest.mark.gpu
def test_newton_with_direct_llt():
    capi, ctx, g, handles = setup("tetbar_n2")
    s = ctx.newton_default_settings()
    s.contact_enabled = 0
    s.max_iterations = 50
    s.linear_solver = 0
    st = ctx.newton_solve(s)
    assert st.result == 0 and st.cg_iterations == 0
    assert st.last_residual < 1e-6 or st.n_evaluations > 1
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("limit", [1024, 4096, 16384])
def test_pcg_streaming_paths(limit):
    """The persistent PCG keeps matrix / vector slices in shared memory when they fit; SB_PCG_SMEM_LIMIT shrinks the budget so
    that this small fixture takes the paths of scenes that do not fit (slices streamed from global memory, partial or no
    window of the SpMV operand).  Same iterates as the resident path."""
    import subprocess, sys, json, textwrap
    code = textwrap.dedent("""
        import sys, json, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from golden_util import Golden, bind
        from stark_b200 import capi
        g = Golden("tetdrop_n5"); ctx = capi.Context(0)
        bind(ctx, g, set(capi.kernel_names()))
        ctx.eval("PGH"); ctx.assemble()
        out = ctx.solve_pcg(g.meta["pcg_abs_tol"], g.meta["pcg_rel_tol"], 10000, True)
        du = ctx.du()
        print(json.dumps({"it": out["iterations"], "ok": out["ok"], "dg": out["du_dot_grad"], "du": du.tolist()}))
    """) % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    res = []
    for lim in (None, limit):
        env = dict(os.environ)
        if lim is not None:
            env["SB_PCG_SMEM_LIMIT"] = str(lim)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        res.append(json.loads(out.stdout.strip().splitlines()[-1]))
    a, b = res
    assert a["ok"] and b["ok"] and a["it"] == b["it"]
    da, db = np.array(a["du"]), np.array(b["du"])
    assert np.abs(da - db).max() <= 1e-12 * np.abs(da).max()


@pytest.mark.gpu
def test_assembly_absorbs_changed_contact_tables_without_a_symbolic_phase():
    """Scatter mode of the assembly (assembly.cu): once the pattern holds the blocks of the contact neighbourhood, a detection that
    changes the contact tables is absorbed without sorting -- every dynamic source is looked up in the BCSR rows and added through
    an FP64 hash table.  After a run that has been through such steps, the matrix handed out by sb_bcsr_get must still be the
    exact assembly of the element Hessians (pattern with explicit zero blocks of vanished pairs dropped; values one float ulp from
    the float64-accumulated oracle), and the trajectory must equal the one of a run with SB_NO_SCATTER=1."""
    import json, subprocess, textwrap
    here = os.path.dirname(os.path.abspath(__file__))
    code = textwrap.dedent("""
        import sys, json
        sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
        import numpy as np
        from stark_b200 import scenes, capi
        import oracle
        sc = scenes.Scene("tetdrop", n=6, vz=0.6)
        log = []
        for _ in range(14):
            s = sc.step()
            log.append([s["accepted"], s["newton_iterations"], s["result"], s["first_residual"]])
        ctx = capi.Context.borrow(sc.lib.sbh_scene_context(sc.h))
        ctx.eval("PGH"); ctx.assemble()
        rows, cols, vals = ctx.bcsr()
        Hs, erows = [], []
        pot = 0
        while True:
            try:
                n_in, n, ne = ctx.potential_info(pot)
            except capi.SBError:
                break
            if ne > 0:
                Hs += list(ctx.hessians(pot)); erows += [list(x) for x in ctx.block_rows(pot)]
            pot += 1
        rp, oc, ov = oracle.assemble_bcsr(Hs, erows, len(rows) - 1)
        ok_pattern = bool(np.array_equal(rp, rows) and np.array_equal(oc, cols))
        err = 1.0
        if ok_pattern:
            d = np.abs(ov.astype(np.float64) - vals.astype(np.float64)).reshape(-1, 9)
            sc_ = np.abs(ov.astype(np.float64)).reshape(-1, 9).max(axis=1, keepdims=True) + 1e-30
            err = float((d / sc_).max())
        print(json.dumps({"log": log, "x": sc.positions()[::7].tolist(), "pattern": ok_pattern, "err": err}))
    """) % (here, os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"))
    out = []
    for no_scatter in (False, True):
        env = dict(os.environ, SB_ASM_DUMP="1")
        env.pop("SB_NO_SCATTER", None)
        if no_scatter:
            env["SB_NO_SCATTER"] = "1"
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out.append((json.loads(r.stdout.strip().splitlines()[-1]), r.stderr))
    (a, err_a), (b, err_b) = out
    # the scatter path was really taken (and not in the control run)
    import re
    hits = int(re.search(r"without a symbolic phase: (\d+)", err_a).group(1))
    assert hits > 0, err_a
    assert int(re.search(r"without a symbolic phase: (\d+)", err_b).group(1)) == 0
    for run in (a, b):
        assert run["pattern"], "BCSR pattern differs from the assembly of the element Hessians"
        assert run["err"] < 2e-7
    for sa, sb in zip(a["log"], b["log"]):
        assert sa[:3] == sb[:3], (sa, sb)
        assert abs(sa[3] - sb[3]) <= 1e-4 * abs(sa[3]), (sa, sb)
    assert np.abs(np.array(a["x"]) - np.array(b["x"])).max() <= 1e-6
