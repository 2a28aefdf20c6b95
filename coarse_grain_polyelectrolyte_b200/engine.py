"""Product entry point: `Engine` drives libcgpe.so (hand-written sm_100a CUDA) through its C-ABI.

There is no CPU path: if the shared library is missing or no CUDA device is visible the
constructor raises.  Build the library with `python -m coarse_grain_polyelectrolyte_b200.build`
(or `__graft_entry__.build()`).
"""
import ctypes as C
import os

from . import _abi
from ._abi import CgpeError, make_config  # noqa: F401  (re-exported)

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcgpe.so")
_binding = None


def library_path():
    return _LIB_PATH


def load_library():
    """dlopen libcgpe.so and bind every symbol include/cgpe.h declares."""
    global _binding
    if _binding is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -m coarse_grain_polyelectrolyte_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        lib = C.CDLL(_LIB_PATH)
        b = _abi.Binding(lib, "cgpe_")
        if b.abi_version() != _abi.ABI_VERSION:
            raise RuntimeError(f"libcgpe.so ABI {b.abi_version()} != Python binding {_abi.ABI_VERSION}: rebuild")
        _binding = b
    return _binding


class Engine(_abi.EngineBase):
    """R replicas of one bead-spring polyelectrolyte system resident on one B200."""

    def __init__(self, cfg):
        super().__init__(load_library(), cfg)
