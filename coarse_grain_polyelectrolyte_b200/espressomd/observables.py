"""espressomd.observables — the observables the reference registers (Observables.py:27-76 of the
reference: RDF, ParticleDistances, BondDihedrals, ParticlePositions, ComPosition, ComVelocity),
same constructor keywords, evaluated on the configuration the engine holds.

These are analysis, not the hot path: `calculate()` reads the state back once (cached by the
System until the next engine call) and reduces it with numpy.  With a replica axis (R > 1) every
result gains a leading axis of length R; with R = 1 the shapes are ESPResSo's.
"""
import numpy as np


def _system(explicit=None):
    if explicit is not None:
        return explicit
    from . import _current_system
    if _current_system is None:
        raise RuntimeError("no espressomd.System has been created")
    return _current_system


class Observable:
    """Base class: subclasses implement _evaluate(pos [R][P][3], vel [R][P][3], system)."""

    def __init__(self, ids=None, system=None):
        self.ids = None if ids is None else np.asarray(list(ids), dtype=np.int64)
        if self.ids is not None and self.ids.ndim != 1:
            raise ValueError("ids must be a flat sequence of particle ids")
        self._bound = system

    def _state(self):
        s = _system(self._bound)
        xyzq, vel, _types = s._state()
        n = xyzq.shape[1]
        if self.ids is not None and self.ids.size and (self.ids.min() < 0 or self.ids.max() >= n):
            raise ValueError(f"particle id out of range (the system holds {n} particles)")
        return np.asarray(xyzq[..., :3], np.float64), np.asarray(vel[..., :3], np.float64), s

    def calculate(self):
        pos, vel, s = self._state()
        out = self._evaluate(pos, vel, s)
        return out[0] if s.n_replicas == 1 else out

    def shape(self):
        return tuple(np.shape(self.calculate()))


class ParticlePositions(Observable):           # O.py:63-65
    def _evaluate(self, pos, vel, s):
        return pos[:, self.ids]


class ParticleVelocities(Observable):
    def _evaluate(self, pos, vel, s):
        return vel[:, self.ids]


class ParticleDistances(Observable):           # O.py:55-57: distances between consecutive ids
    def _evaluate(self, pos, vel, s):
        d = np.diff(pos[:, self.ids], axis=1)
        L = float(s.box_l[0])
        d -= L * np.rint(d / L)
        return np.linalg.norm(d, axis=2)


class BondAngles(Observable):
    def _evaluate(self, pos, vel, s):
        x = pos[:, self.ids]
        a, b = x[:, :-2] - x[:, 1:-1], x[:, 2:] - x[:, 1:-1]
        c = np.einsum("rij,rij->ri", a, b) / (np.linalg.norm(a, axis=2) * np.linalg.norm(b, axis=2))
        return np.arccos(np.clip(c, -1.0, 1.0))


class BondDihedrals(Observable):               # O.py:59-61: torsion of consecutive quadruples, in (-pi, pi]
    def _evaluate(self, pos, vel, s):
        x = pos[:, self.ids]
        b1, b2, b3 = x[:, 1:-2] - x[:, :-3], x[:, 2:-1] - x[:, 1:-2], x[:, 3:] - x[:, 2:-1]
        n1, n2 = np.cross(b1, b2), np.cross(b2, b3)
        m = np.cross(n1, b2 / np.linalg.norm(b2, axis=2, keepdims=True))
        return np.arctan2(np.einsum("rij,rij->ri", m, n2), np.einsum("rij,rij->ri", n1, n2))


class ComPosition(Observable):                 # O.py:67-72 (all masses are 1)
    def _evaluate(self, pos, vel, s):
        return pos[:, self.ids].mean(axis=1)


class ComVelocity(Observable):                 # O.py:73-76
    def _evaluate(self, pos, vel, s):
        return vel[:, self.ids].mean(axis=1)


class RDF(Observable):
    """Radial distribution function between ids1 and ids2 (ids1 with itself when ids2 is None),
    minimum image, normalised by the ideal-gas pair count of each shell (O.py:27-29)."""

    def __init__(self, ids1, ids2=None, min_r=0.0, max_r=None, n_r_bins=100, system=None):
        super().__init__(ids=None, system=system)
        self.ids1 = np.asarray(list(ids1), dtype=np.int64)
        self.ids2 = None if ids2 is None else np.asarray(list(ids2), dtype=np.int64)
        if max_r is None or not (max_r > min_r >= 0.0):
            raise ValueError("RDF needs 0 <= min_r < max_r")
        if int(n_r_bins) < 1:
            raise ValueError("n_r_bins must be positive")
        self.min_r, self.max_r, self.n_r_bins = float(min_r), float(max_r), int(n_r_bins)

    def bin_edges(self):
        return np.linspace(self.min_r, self.max_r, self.n_r_bins + 1)

    def bin_centers(self):
        e = self.bin_edges()
        return 0.5 * (e[:-1] + e[1:])

    def _evaluate(self, pos, vel, s):
        L = float(s.box_l[0])
        a = pos[:, self.ids1]
        same = self.ids2 is None
        b = a if same else pos[:, self.ids2]
        edges = self.bin_edges()
        shell = 4.0 / 3.0 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
        n_pairs = a.shape[1] * (a.shape[1] - 1) / 2 if same else a.shape[1] * b.shape[1]
        out = np.zeros((pos.shape[0], self.n_r_bins))
        for r in range(pos.shape[0]):
            d = a[r][:, None, :] - b[r][None, :, :]
            d -= L * np.rint(d / L)
            dist = np.sqrt(np.einsum("ijk,ijk->ij", d, d))
            if same:
                dist = dist[np.triu_indices(a.shape[1], k=1)]
            counts, _ = np.histogram(dist.ravel(), bins=edges)
            out[r] = counts * (L ** 3) / (shell * max(n_pairs, 1))
        return out
