"""ctypes front end of the CPU oracle (oracle/cgpe_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs,
never by the product package.  PARITY UNPINNED against ESPResSo (see the C file's header).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from coarse_grain_polyelectrolyte_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "cgpe_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "cgpe.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


_binding = None


def binding():
    global _binding
    if _binding is None:
        lib = C.CDLL(build())
        _binding = _abi.Binding(lib, "orc_")
        lib.orc_get_state_f64.argtypes = [C.c_void_p] * 4
        lib.orc_forces_f64.argtypes = [C.c_void_p] * 2
        lib.orc_set_positions_f64.argtypes = [C.c_void_p] * 2
        lib.orc_mc_run_follow.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_double, C.POINTER(C.c_int32)]
        lib.orc_philox4x32_10.argtypes = [C.c_void_p] * 3
        lib.orc_philox4x32_10.restype = None
        lib.orc_set_threads.argtypes = [C.c_int32]
        lib.orc_set_list_mode.argtypes = [C.c_void_p, C.c_int32, C.c_double]
    return _binding


class OracleEngine(_abi.EngineBase):
    """FP64 CPU restatement behind the same interface as the CUDA Engine."""

    def __init__(self, cfg):
        super().__init__(binding(), cfg)

    def set_list_mode(self, on=True, skin=0.0):
        """Verlet-list evaluation (same sums, listed pairs only): the CPU-baseline configuration."""
        self._check(self._b.lib.orc_set_list_mode(self._h, int(on), float(skin)), "set_list_mode")

    def state_f64(self):
        x = np.empty((self.R, self.P, 3)); v = np.empty((self.R, self.P, 3)); q = np.empty((self.R, self.P))
        self._check(self._b.lib.orc_get_state_f64(self._h, _abi._ptr(x), _abi._ptr(v), _abi._ptr(q)), "get_state_f64")
        return x, v, q

    def set_positions_f64(self, x):
        x = np.ascontiguousarray(np.broadcast_to(np.asarray(x, np.float64), (self.R, self.P, 3)))
        self._check(self._b.lib.orc_set_positions_f64(self._h, _abi._ptr(x)), "set_positions_f64")

    def forces_f64(self):
        out = np.empty((self.R, self.P, 3))
        self._check(self._b.lib.orc_forces_f64(self._h, _abi._ptr(out)), "forces_f64")
        return out

    def mc_run_follow(self, reaction_id, device_trace, tie_tol=1e-4):
        n = device_trace.shape[1]
        rec = np.zeros((self.R, n), _abi.TRIAL_DTYPE)
        ties = C.c_int32()
        dev = np.ascontiguousarray(device_trace)
        self._check(self._b.lib.orc_mc_run_follow(self._h, reaction_id, n, _abi._ptr(rec), _abi._ptr(dev),
                                                  tie_tol, C.byref(ties)), "mc_run_follow")
        return rec, ties.value


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    binding().lib.orc_philox4x32_10(_abi._ptr(c), _abi._ptr(k), _abi._ptr(o))
    return o


def set_threads(n):
    """Number of OpenMP threads the oracle's replica loops use (CPU baseline timing)."""
    binding().lib.orc_set_threads(int(n))
