"""CPU-side checks: the C-ABI library loads and exports every symbol include/stark_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from stark_b200 import capi
    header = open(os.path.join(ROOT, "include", "stark_b200.h")).read()
    declared = set(re.findall(r"SB_API\s+[\w\s\*]+?\b(sb_\w+)\s*\(", header))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s


def test_kernel_registry_covers_reference_potentials():
    """Every potential the reference registers for the benchmark configs has a hand-written kernel."""
    from stark_b200 import capi
    from golden_util import Golden
    names = set(capi.kernel_names())
    for fx in ["tetdrop_n3", "tetbar_n2", "cloth_n8", "cloth_shells_n8"]:
        g = Golden(fx)
        for i, p in g.potentials():
            assert p["name"] in names, p["name"]


def test_every_kernel_has_reference_elements_in_a_fixture():
    """Each of the 63 potentials the reference registers -- i.e. every name in sb_kernel_names() except the AD cross-check
    twin of the analytic tet kernel -- has at least one ACTIVE element in a committed reference fixture, and the GPU parity
    tests (test_eval_parity.FIXTURES) run over all of those fixtures."""
    import glob
    from stark_b200 import capi
    from golden_util import Golden
    import test_eval_parity
    covered = {}
    for fx in test_eval_parity.FIXTURES:
        g = Golden(fx)
        for i, p in g.potentials():
            if g[f"pot{i}_active"].any():
                covered.setdefault(p["name"], []).append(fx)
    registered = [p["name"] for p in Golden("tetdrop_n3").meta["potentials"]]
    assert len(registered) == 63
    assert not [n for n in registered if n not in covered], [n for n in registered if n not in covered]
    kernels = set(capi.kernel_names()) - {"EnergyTetStrain_AD", "EnergyTetStrain_Elasticity_Only_AD"}
    assert kernels == set(registered), kernels ^ set(registered)
    assert sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))) == sorted(test_eval_parity.FIXTURES)


def test_create_fails_loudly_without_gpu():
    import torch
    from stark_b200 import capi
    if torch.cuda.is_available():
        return
    try:
        capi.Context(0)
    except capi.SBError:
        return
    raise AssertionError("sb_create must fail when no CUDA device is visible (no CPU fallback)")
