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
