"""stark_b200: B200-native Newton hot path of STARK behind a C-ABI (include/stark_b200.h).

Python here is only plumbing for tests and bench.py; the product is stark_b200/lib/libstark_b200.so."""
from . import capi  # noqa: F401
