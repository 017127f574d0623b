"""flagger_b200: B200-native E-step for HMM-Flagger (mobinasri/flagger v1.2.0) behind a C-ABI.

The compute lives in flagger_b200/csrc (CUDA, sm_100a) and is reached only through
include/hfg.h; this package is the thin Python host side (ctypes binding, chunk/window
data formats, synthetic workloads) used by tests and bench.py.
"""
from . import _abi  # noqa: F401

__all__ = ["_abi"]
