"""turbo_b200 — B200-native dive-and-solve engine for Turbo's TNF interval propagation path.

The product is the C ABI in include/turbo_b200.h implemented by turbo_b200/libturbo_b200.so
(hand-written sm_100a kernels + a C++ front-end) and the `turbo` command-line driver.  This Python
package is a thin ctypes mirror of that ABI used by the tests and by bench.py.
"""
from . import abi  # noqa: F401

__version__ = "0.1.0"
