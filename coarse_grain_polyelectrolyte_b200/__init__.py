"""B200-native constant-pH Monte Carlo + Langevin MD for weak polyelectrolytes.

The hot path lives in libcgpe.so (csrc/*.cu, C-ABI in include/cgpe.h); this package holds the
ctypes binding (`engine.Engine`), the pint-free unit reduction, host-side system construction
and the `espressomd`-shaped facade the src/modules drop-in driver is written against.
"""
from ._abi import CgpeError, make_config  # noqa: F401

__all__ = ["CgpeError", "make_config"]
__version__ = "0.1.0"
