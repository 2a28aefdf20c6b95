"""Bonded interaction objects with ESPResSo's constructor keywords (M.py:12-15,34-54)."""
import math


class BondedInteraction:
    n_partners = 1

    def __init__(self):
        self._id = None

    @property
    def params(self):
        return {k: v for k, v in vars(self).items() if not k.startswith("_")}


class HarmonicBond(BondedInteraction):
    """HarmonicBond(k=, r_0=): U = 1/2 k (r - r_0)^2 (M.py:39)."""

    def __init__(self, k, r_0, r_cut=None):
        super().__init__()
        if k < 0 or r_0 < 0:
            raise ValueError("k and r_0 must be non-negative")
        self.k, self.r_0 = float(k), float(r_0)


class FeneBond(BondedInteraction):
    """FeneBond(k=, d_r_max=, r_0=): U = -1/2 k d_r_max^2 ln(1 - ((r - r_0)/d_r_max)^2) (M.py:35-37)."""

    def __init__(self, k, d_r_max, r_0=0.0):
        super().__init__()
        if k < 0 or d_r_max <= 0 or r_0 < 0:
            raise ValueError("k, r_0 must be non-negative and d_r_max positive")
        self.k, self.d_r_max, self.r_0 = float(k), float(d_r_max), float(r_0)


class AngleHarmonic(BondedInteraction):
    """AngleHarmonic(bend=, phi0=): U = 1/2 bend (phi - phi0)^2 (M.py:43-45)."""
    n_partners = 2

    def __init__(self, bend, phi0):
        super().__init__()
        if bend < 0 or not 0.0 <= phi0 <= math.pi:
            raise ValueError("bend must be non-negative and phi0 within [0, pi]")
        self.bend, self.phi0 = float(bend), float(phi0)


class Dihedral(BondedInteraction):
    """Dihedral(bend=, mult=, phase=): U = bend [1 - cos(mult phi - phase)] (M.py:49-53)."""
    n_partners = 3

    def __init__(self, bend, mult, phase, sin_min=0.0):
        """sin_min is not an ESPResSo argument: the engine evaluates the torsion gradient with sin(bond angle) >= sin_min
        (cgpe.h, dihedral_sin_min).  0 = the library's default safeguard (0.02), negative = ESPResSo's own guard only
        (the term is dropped for |b1 x b2| <= 1e-4, nothing else)."""
        super().__init__()
        if int(mult) != mult or not 0 <= int(mult) <= 12:
            raise ValueError("mult must be an integer in [0, 12]")
        self.bend, self.mult, self.phase, self.sin_min = float(bend), int(mult), float(phase), float(sin_min)
