"""User potentials: symx operation sequence -> CUDA source -> NVRTC -> element kernel (stark_b200/csrc/user.cu), the replacement of
the reference's code generator + g++ JIT (symx/compile/Compilation.cpp:381-469).
CPU: the generator's output for a hand-written sequence and its NVRTC cross-compilation for sm_100a (no GPU needed).
GPU: the generated kernel against numpy on the same sequence, and against the reference's own outputs for the sequences the
reference itself produced for examples/main.cpp's EnergyMagneticAttraction (tests/golden/magnet_n2.npz)."""
import ctypes as C
import os

import numpy as np
import pytest

from stark_b200 import capi

ADD, SUB, MUL, RECIP, SQRT, CONST, OUT, BRANCH = 6, 7, 8, 9, 12, 4, 5, 2


def ops_array(ops):
    return capi.ops_array([o[:5] for o in ops], [o[5] for o in ops])


def spring_sequences():
    """E = k/2 |x - t|^2 if k > 0 else 0 (one 3-DoF block): in = [x(3) | t(3) | k]."""
    head = [(SUB, 7, 0, 3, 0, 0.0), (SUB, 8, 1, 4, 0, 0.0), (SUB, 9, 2, 5, 0, 0.0),
            (MUL, 10, 7, 7, 0, 0.0), (MUL, 11, 8, 8, 0, 0.0), (MUL, 12, 9, 9, 0, 0.0),
            (ADD, 13, 10, 11, 0, 0.0), (ADD, 14, 13, 12, 0, 0.0), (CONST, 15, -1, -1, 0, 0.5), (MUL, 16, 15, 6, 0, 0.0), (MUL, 17, 16, 14, 0, 0.0),
            (CONST, 18, -1, -1, 0, 0.0)]
    p = head + [(BRANCH, -1, 0, -1, 6, 0.0), (OUT, 0, 17, -1, 0, 0.0), (BRANCH, -1, 1, -1, 6, 0.0), (OUT, 0, 18, -1, 0, 0.0), (BRANCH, -1, -1, -1, -2, 0.0)]
    pgh = head + [(MUL, 19, 6, 7, 0, 0.0), (MUL, 20, 6, 8, 0, 0.0), (MUL, 21, 6, 9, 0, 0.0),
                  (BRANCH, -1, 0, -1, 6, 0.0), (OUT, 0, 17, -1, 0, 0.0), (OUT, 1, 19, -1, 0, 0.0), (OUT, 2, 20, -1, 0, 0.0), (OUT, 3, 21, -1, 0, 0.0)]
    for i in range(3):
        for j in range(3):
            pgh.append((OUT, 4 + 3 * i + j, 6 if i == j else 18, -1, 0, 0.0))
    pgh.append((BRANCH, -1, 1, -1, 6, 0.0))
    pgh += [(OUT, k, 18, -1, 0, 0.0) for k in range(13)]
    pgh.append((BRANCH, -1, -1, -1, -2, 0.0))
    return ops_array(p), ops_array(pgh)


def test_codegen_and_nvrtc_cross_compilation(tmp_path, monkeypatch):
    monkeypatch.setenv("SB_CACHE_DIR", str(tmp_path / "cache"))
    lib = capi.load()
    p, pgh = spring_sequences()
    n = C.c_longlong()
    assert lib.sb_user_codegen(b"Spring", 7, 1, p, len(p), pgh, len(pgh), None, 0, C.byref(n)) == 0
    buf = C.create_string_buffer(n.value + 1)
    assert lib.sb_user_codegen(b"Spring", 7, 1, p, len(p), pgh, len(pgh), buf, n.value + 1, None) == 0
    src = buf.value.decode()
    assert "const double v7 = in[0] - in[3];" in src and "if (in[6] > 0.0)" in src and "out[12] = in[6];" in src and 'extern "C" __global__' in src
    assert src.count("{") == src.count("}")
    # an unbalanced branch and an unknown operation are rejected, not emitted
    bad = ops_array([(BRANCH, -1, 0, -1, 6, 0.0), (OUT, 0, 6, -1, 0, 0.0)])
    assert lib.sb_user_codegen(b"Bad", 7, 1, bad, len(bad), pgh, len(pgh), None, 0, None) != 0
    bad = ops_array([(3, 7, 0, 0, 0, 0.0)])
    assert lib.sb_user_codegen(b"Bad", 7, 1, bad, len(bad), pgh, len(pgh), None, 0, None) != 0
    # NVRTC compiles the source for sm_100a without a GPU; the second call is served from the cache
    size, cached = C.c_longlong(), C.c_int()
    log = C.create_string_buffer(4096)
    rc = lib.sb_user_compile(buf.value, C.byref(size), C.byref(cached), log, 4096)
    if rc == -4:
        pytest.skip("NVRTC is not installed here: " + log.value.decode())
    assert rc == 0, log.value.decode()
    assert size.value > 1000 and cached.value == 0
    assert lib.sb_user_compile(buf.value, C.byref(size), C.byref(cached), log, 4096) == 0 and cached.value == 1
    assert any(f.endswith(".cubin") for f in os.listdir(tmp_path / "cache"))


def _fetch(entries):
    arr = (capi.Fetch * len(entries))()
    for i, (array, col, slot, stride) in enumerate(entries):
        arr[i] = capi.Fetch(array, col, slot, stride)
    return arr


@pytest.mark.gpu
def test_generated_kernel_against_numpy(tmp_path, monkeypatch):
    monkeypatch.setenv("SB_CACHE_DIR", str(tmp_path / "cache"))
    ctx = capi.Context(0)
    lib = ctx.lib
    rng = np.random.default_rng(5)
    n_nodes, n_elem = 50, 37
    x = rng.standard_normal((n_nodes, 3))
    t = rng.standard_normal((n_elem, 3))
    k = rng.uniform(-1.0, 3.0, (n_elem, 1))          # negative stiffness -> the else branch (zero energy)
    ax, at, ak = ctx.array("x", 3, x), ctx.array("t", 3, t), ctx.array("k", 1, k)
    ctx.dof_add(ax)
    conn = np.stack([np.arange(n_elem), rng.integers(0, n_nodes, n_elem)], axis=1).astype(np.int32)   # [element idx, node]
    p, pgh = spring_sequences()
    h = C.c_int()
    slots = (C.c_int32 * 1)(0)
    rc = lib.sb_potential_create_user(ctx.h, b"Spring", 2, _fetch([(ax, 1, 0, 3), (at, 0, 3, 3), (ak, 0, 6, 1)]), 3, 7, 1, slots, p, len(p), pgh, len(pgh), C.byref(h))
    assert rc == 0, lib.sb_last_error(ctx.h)
    ctx.set_connectivity(h.value, conn)
    E, res = ctx.eval("PGH")
    d = x[conn[:, 1]] - t
    on = (k[:, 0] > 0)
    E_ref = float(np.sum(0.5 * k[on, 0] * np.sum(d[on] ** 2, axis=1)))
    g_ref = np.zeros_like(x)
    np.add.at(g_ref, conn[on, 1], k[on] * d[on])
    assert abs(E - E_ref) <= 1e-13 * abs(E_ref)
    assert np.abs(ctx.grad() - g_ref.ravel()).max() <= 1e-13 * np.abs(g_ref).max()
    H = ctx.hessians(h.value)
    for e in range(n_elem):
        assert np.array_equal(H[e], (k[e, 0] if on[e] else 0.0) * np.eye(3))
    E_only = ctx.eval("P")
    assert E_only == E
    # the element Hessians assemble like any other potential's
    ctx.assemble()
    rows, cols, vals = ctx.bcsr()
    assert len(rows) == n_nodes + 1 and len(cols) == len(set(conn[:, 1]))   # one diagonal block per touched node
    ctx.close()


def test_reference_sequences_of_a_user_potential_compile(tmp_path, monkeypatch):
    """The sequences the reference itself produced for examples/main.cpp's EnergyMagneticAttraction (fixture magnet_n2: 20 / 112
    operations incl. Sqrt, Reciprocal, PowN) translate and cross-compile."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from golden_util import Golden
    monkeypatch.setenv("SB_CACHE_DIR", str(tmp_path / "cache"))
    g = Golden("magnet_n2")
    i, p = next((i, p) for i, p in g.potentials() if p.get("user_ops"))
    assert p["name"] == "EnergyMagneticAttraction" and p["n_dofs"] == 3
    lib = capi.load()
    ops_p = capi.ops_array(g[f"pot{i}_ops_p"], g[f"pot{i}_opsc_p"])
    ops_pgh = capi.ops_array(g[f"pot{i}_ops_pgh"], g[f"pot{i}_opsc_pgh"])
    n = C.c_longlong()
    assert lib.sb_user_codegen(p["name"].encode(), p["n_in"], 1, ops_p, len(ops_p), ops_pgh, len(ops_pgh), None, 0, C.byref(n)) == 0
    buf = C.create_string_buffer(n.value + 1)
    assert lib.sb_user_codegen(p["name"].encode(), p["n_in"], 1, ops_p, len(ops_p), ops_pgh, len(ops_pgh), buf, n.value + 1, None) == 0
    src = buf.value.decode()
    assert src.count("] = ") >= 13 + 1 and "sqrt(" in src
    size, cached = C.c_longlong(), C.c_int()
    log = C.create_string_buffer(4096)
    rc = lib.sb_user_compile(buf.value, C.byref(size), C.byref(cached), log, 4096)
    if rc == -4:
        pytest.skip("NVRTC is not installed here")
    assert rc == 0 and size.value > 1000, log.value.decode()


def test_generated_statements_evaluate_like_the_reference_on_the_host(tmp_path):
    """The statement list the generator prints for the reference's own sequences (magnet_n2) is plain C: compiled with gcc and run
    on the fixture's gathered inputs it reproduces the per-element [E | grad | hess] of the reference's JIT-compiled code --
    a check of the generator's text itself that needs no GPU."""
    import re
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle
    from golden_util import Golden
    g = Golden("magnet_n2")
    i, p = next((i, p) for i, p in g.potentials() if p.get("user_ops"))
    lib = capi.load()
    ops_p = capi.ops_array(g[f"pot{i}_ops_p"], g[f"pot{i}_opsc_p"])
    ops_pgh = capi.ops_array(g[f"pot{i}_ops_pgh"], g[f"pot{i}_opsc_pgh"])
    n = C.c_longlong()
    assert lib.sb_user_codegen(p["name"].encode(), p["n_in"], 1, ops_p, len(ops_p), ops_pgh, len(ops_pgh), None, 0, C.byref(n)) == 0
    buf = C.create_string_buffer(n.value + 1)
    assert lib.sb_user_codegen(p["name"].encode(), p["n_in"], 1, ops_p, len(ops_p), ops_pgh, len(ops_pgh), buf, n.value + 1, None) == 0
    src = buf.value.decode()
    m = re.search(r"void f_pgh\(const double\* __restrict__ in, double\* __restrict__ out\)\n\{\n(.*?)\n\}\n__device__", src, re.S)
    assert m, "f_pgh not found in the generated source"
    c_file = tmp_path / "f_pgh.c"
    c_file.write_text("#include <math.h>\n#define SB_INF INFINITY\nvoid f_pgh(const double* in, double* out)\n{\n" + m.group(1) + "\n}\n")
    so = tmp_path / "f_pgh.so"
    subprocess.run(["gcc", "-O1", "-ffp-contract=off", "-shared", "-fPIC", str(c_file), "-o", str(so), "-lm"], check=True)
    f = C.CDLL(str(so)).f_pgh
    arrays = {k: g[f"array{k}"] for k in range(len(g.meta["arrays"]))}
    maps = [(mm["array"], mm["conn_idx"], mm["first_symbol"], mm["stride"]) for mm in p["maps"]]
    conn = g[f"pot{i}_conn"][g[f"pot{i}_active"].astype(bool)]
    X = np.ascontiguousarray(oracle.gather_inputs(p["n_in"], conn, maps, arrays))
    ref = g[f"pot{i}_sol"][g[f"pot{i}_active"].astype(bool)]
    out = np.zeros_like(ref)
    for e in range(X.shape[0]):
        f(X[e].ctypes.data_as(C.POINTER(C.c_double)), out[e].ctypes.data_as(C.POINTER(C.c_double)))
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
