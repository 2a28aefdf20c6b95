"""espressomd.accumulators — TimeSeries, MeanVarianceCalculator and the multiple-tau Correlator the
reference builds in Observables.py:27-76, with ESPResSo 4.2's keywords.  An accumulator added to
`system.auto_update_accumulators` is updated every `delta_N` steps of `integrator.run`.

The correlator is a multiple-tau scheme: level 0 keeps the last tau_lin + 1 samples at the sampling
interval; each further level holds samples at twice the interval of the one below ("discard1"
keeps the newer sample of a pair, "discard2" the older, "linear" their mean) and contributes the
lags tau_lin/2 + 1 ... tau_lin of its own resolution.  (ESPResSo is not available here, so the
lag grid follows its documented layout; values at level 0 are the exact direct correlation.)
"""
import numpy as np


class _Accumulator:
    def __init__(self, delta_N=1):
        if int(delta_N) < 1:
            raise ValueError("delta_N must be positive")
        self.delta_N = int(delta_N)
        self._since_update = 0          # integration steps since the last automatic update


class TimeSeries(_Accumulator):
    def __init__(self, obs, delta_N=1):
        super().__init__(delta_N)
        self.obs = obs
        self._data = []

    def update(self):
        self._data.append(np.array(self.obs.calculate(), dtype=np.float64))

    def update_with(self, value, _second=None):
        """Feed one sample that was evaluated elsewhere (the engine's on-device frame recorder)."""
        self._data.append(np.array(value, dtype=np.float64))

    def time_series(self):
        return np.array(self._data)

    def clear(self):
        self._data = []


class MeanVarianceCalculator(_Accumulator):
    def __init__(self, obs, delta_N=1):
        super().__init__(delta_N)
        self.obs = obs
        self._n, self._mean, self._m2 = 0, None, None

    def update(self):
        self.update_with(self.obs.calculate())

    def update_with(self, value, _second=None):
        x = np.array(value, dtype=np.float64)
        if self._mean is None:
            self._mean, self._m2 = np.zeros_like(x), np.zeros_like(x)
        self._n += 1
        d = x - self._mean
        self._mean += d / self._n
        self._m2 += d * (x - self._mean)

    def mean(self):
        return self._mean

    def variance(self):
        return self._m2 / (self._n - 1) if self._n > 1 else np.full_like(self._mean, np.nan)

    def std_error(self):
        return np.sqrt(self.variance() / self._n)


_OPS = {
    "scalar_product": lambda old, new: np.sum(old * new, axis=-1, keepdims=True),
    "componentwise_product": lambda old, new: old * new,
    "square_distance_componentwise": lambda old, new: (new - old) ** 2,
}
_COMPRESS = {
    "discard1": lambda older, newer: newer,
    "discard2": lambda older, newer: older,
    "linear": lambda older, newer: 0.5 * (older + newer),
}


class Correlator(_Accumulator):
    def __init__(self, obs1, obs2=None, tau_lin=16, tau_max=None, delta_N=1, corr_operation="scalar_product",
                 compress1="discard1", compress2=None, dt=None):
        super().__init__(delta_N)
        if corr_operation not in _OPS:
            raise ValueError(f"corr_operation must be one of {sorted(_OPS)}")
        if compress1 not in _COMPRESS or (compress2 is not None and compress2 not in _COMPRESS):
            raise ValueError(f"compress1/compress2 must be one of {sorted(_COMPRESS)}")
        if int(tau_lin) < 2 or int(tau_lin) % 2:
            raise ValueError("tau_lin must be an even number >= 2")
        if tau_max is None or not tau_max > 0:
            raise ValueError("tau_max must be positive")
        self.obs1, self.obs2 = obs1, obs2
        self.tau_lin, self.tau_max = int(tau_lin), float(tau_max)
        self.corr_operation, self.compress1, self.compress2 = corr_operation, compress1, compress2 or compress1
        self._dt = dt                    # time between updates; default delta_N * system.time_step at first update
        self._levels = None
        self._finalized = False

    # -- layout ----------------------------------------------------------------------------------
    def _setup(self):
        if self._dt is None:
            from .observables import _system
            ts = _system(getattr(self.obs1, "_bound", None)).time_step
            if ts is None:
                raise RuntimeError("system.time_step must be set before a Correlator is updated")
            self._dt = self.delta_N * float(ts)
        depth = 1
        while self.tau_lin * (2 ** (depth - 1)) * self._dt < self.tau_max:
            depth += 1
        self._levels = [dict(a=[], b=[], n=0, corr={}, cnt={}) for _ in range(depth)]

    def _lags(self, level):
        return range(0, self.tau_lin + 1) if level == 0 else range(self.tau_lin // 2 + 1, self.tau_lin + 1)

    def _push(self, level, a, b):
        lv = self._levels[level]
        lv["a"].append(a); lv["b"].append(b)
        if len(lv["a"]) > self.tau_lin + 1:
            lv["a"].pop(0); lv["b"].pop(0)
        lv["n"] += 1
        op = _OPS[self.corr_operation]
        for j in self._lags(level):
            if j < len(lv["a"]):
                c = op(lv["a"][-1 - j], b)
                if j in lv["corr"]:
                    lv["corr"][j] += c; lv["cnt"][j] += 1
                else:
                    lv["corr"][j] = np.array(c, dtype=np.float64); lv["cnt"][j] = 1
        if lv["n"] % 2 == 0 and level + 1 < len(self._levels):
            self._push(level + 1, _COMPRESS[self.compress1](lv["a"][-2], lv["a"][-1]),
                       _COMPRESS[self.compress2](lv["b"][-2], lv["b"][-1]))

    # -- ESPResSo interface ------------------------------------------------------------------------
    def update(self):
        if self._finalized:
            raise RuntimeError("the correlator has been finalized")
        if self._levels is None:
            self._setup()
        a = np.array(self.obs1.calculate(), dtype=np.float64)
        b = a if self.obs2 is None else np.array(self.obs2.calculate(), dtype=np.float64)
        self._push(0, a, b)

    def update_with(self, value, second=None):
        """Feed one sample that was evaluated elsewhere (the engine's on-device frame recorder)."""
        if self._finalized:
            raise RuntimeError("the correlator has been finalized")
        if self._levels is None:
            self._setup()
        a = np.array(value, dtype=np.float64)
        self._push(0, a, a if second is None else np.array(second, dtype=np.float64))

    def finalize(self):
        self._finalized = True

    def _grid(self):
        if self._levels is None:
            return []
        return [(l, j) for l in range(len(self._levels)) for j in self._lags(l) if j in self._levels[l]["corr"]]

    def lag_times(self):
        return np.array([j * (2 ** l) * self._dt for l, j in self._grid()])

    def sample_sizes(self):
        return np.array([self._levels[l]["cnt"][j] for l, j in self._grid()])

    def result(self):
        """[n_lags, ...]: the correlation at every lag of lag_times() (a leading replica axis of the
        observable stays in place after the lag axis)."""
        return np.array([self._levels[l]["corr"][j] / self._levels[l]["cnt"][j] for l, j in self._grid()])


class AutoUpdateAccumulators:
    """system.auto_update_accumulators (O.py:13-19)."""

    def __init__(self):
        self._items = []

    def add(self, acc):
        if not hasattr(acc, "update") or not hasattr(acc, "delta_N"):
            raise TypeError("auto_update_accumulators.add expects an accumulator")
        if acc not in self._items:
            acc._since_update = 0
            self._items.append(acc)
        return acc

    def remove(self, acc):
        self._items.remove(acc)

    def clear(self):
        self._items = []

    def __len__(self):
        return len(self._items)

    def __iter__(self):
        return iter(self._items)

    def steps_until_next(self):
        return min((a.delta_N - a._since_update for a in self._items), default=None)

    def advance(self, steps):
        """`steps` integration steps were taken (never past an accumulator's next update)."""
        for a in self._items:
            a._since_update += steps
            if a._since_update >= a.delta_N:
                a._since_update = 0
                a.update()
