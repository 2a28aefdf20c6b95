"""electrostatics.DH / P3M with ESPResSo's constructor keywords (M.py:85-107)."""
import numpy as np


class DH:
    """Debye-Hueckel: U = prefactor q_i q_j exp(-kappa r)/r for r < r_cut (M.py:97-107).
    kappa and r_cut accept one value or one per replica."""

    def __init__(self, prefactor, kappa, r_cut):
        kappa, r_cut = np.asarray(kappa, dtype=np.float64), np.asarray(r_cut, dtype=np.float64)
        if prefactor < 0 or np.any(kappa < 0) or np.any(r_cut <= 0):
            raise ValueError("prefactor, kappa must be non-negative and r_cut positive")
        self.prefactor = float(prefactor)
        self.kappa = kappa if kappa.ndim else float(kappa)
        self.r_cut = r_cut if r_cut.ndim else float(r_cut)
        self._system = None

    def get_params(self):
        return {"prefactor": self.prefactor, "kappa": self.kappa, "r_cut": self.r_cut}


class P3M:
    """The reference can ask for P3M (M.py:85-95) but every configuration it ships, and the hot
    path this engine implements, uses Debye-Hueckel (USE_P3M False)."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("P3M electrostatics is outside the implemented hot path; use USE_P3M = False")
