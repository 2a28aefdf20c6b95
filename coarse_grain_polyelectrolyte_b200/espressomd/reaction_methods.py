"""reaction_methods.ConstantpHEnsemble with ESPResSo 4.2's keywords (M.py:211-261, E.py:68-71,94-95)."""
import math

import numpy as np


class _Registered:
    """One HA <-> A + B reaction as the engine sees it."""

    def __init__(self, owner, index, gamma, type_prot, type_deprot, type_ion, charges):
        self.owner, self.index = owner, index
        self.gamma = float(gamma)
        self.type_prot, self.type_deprot, self.type_ion = int(type_prot), int(type_deprot), int(type_ion)
        self.charges = dict(charges)

    def as_dict(self):
        o = self.owner
        return dict(type_prot=self.type_prot, type_deprot=self.type_deprot, type_ion=self.type_ion, seed=o.seed,
                    q_prot=float(self.charges[self.type_prot]), q_deprot=float(self.charges[self.type_deprot]),
                    q_ion=float(self.charges.get(self.type_ion, 0.0)), pK=-math.log10(self.gamma), kT=o.kT,
                    exclusion_range=o.exclusion_range)


class ConstantpHEnsemble:
    """ConstantpHEnsemble(kT=, exclusion_range=, seed=, constant_pH=).  constant_pH accepts one
    value or one per replica."""

    def __init__(self, kT, exclusion_range, seed, constant_pH=None, **unsupported):
        if unsupported:
            raise NotImplementedError(f"arguments not supported: {sorted(unsupported)}")
        if kT <= 0:
            raise ValueError("kT must be positive")
        if exclusion_range < 0:
            raise ValueError("exclusion_range must be non-negative")
        from . import _current_system
        if _current_system is None:
            raise RuntimeError("create an espressomd.System first")
        self._system = _current_system
        self.kT, self.exclusion_range, self.seed = float(kT), float(exclusion_range), int(seed)
        self._pH = constant_pH
        self._reactions = []
        self.non_interacting_type = None

    # reactionX.constant_pH = pH (E.py:68-69)
    @property
    def constant_pH(self):
        return self._pH

    @constant_pH.setter
    def constant_pH(self, value):
        a = np.asarray(value, dtype=np.float64)
        if a.ndim > 1 or (a.ndim == 1 and a.size != self._system.n_replicas):
            raise ValueError("constant_pH takes one value or one per replica")
        self._pH = value

    def add_reaction(self, gamma, reactant_types, product_types, default_charges, **unsupported):
        """add_reaction(gamma=10**-pK, reactant_types=[HA], product_types=[A, B], default_charges=) (M.py:235-255)."""
        if unsupported:
            raise NotImplementedError(f"arguments not supported: {sorted(unsupported)}")
        if gamma <= 0:
            raise ValueError("gamma must be positive")
        if len(reactant_types) != 1 or len(product_types) != 2:
            raise ValueError("a constant-pH reaction has the form HA <-> A + B")
        if self._reactions:
            raise NotImplementedError("one reaction per ConstantpHEnsemble object (as the reference uses it)")
        for t in (reactant_types[0], *product_types):
            if t not in default_charges:
                raise ValueError(f"default_charges lacks type {t}")
        s = self._system
        reg = _Registered(self, len(s._reactions), gamma, reactant_types[0], product_types[0], product_types[1],
                          default_charges)
        self._reactions.append(reg)
        s._reactions.append(reg)
        s._structure_changed()

    def set_non_interacting_type(self, type):                     # M.py:256-261
        self.non_interacting_type = int(type)

    def _single_reaction(self):
        if not self._reactions:
            raise RuntimeError("no reaction has been added")
        return self._reactions[0]

    def _push_pH(self, eng):
        if self._pH is None:
            raise RuntimeError("constant_pH has not been set")
        eng.set_replica_params(pH=self._pH)

    def reaction(self, reaction_steps=1):                         # E.py:70-71,94-95
        n = int(reaction_steps)
        if n < 0:
            raise ValueError("reaction_steps must be non-negative")
        eng = self._system._ensure_engine()
        self._push_pH(eng)
        eng.mc_run(self._single_reaction().index, n)
        self._system._state_cache = None

    def get_acceptance_rate_reaction(self, reaction_id=0):
        c = self._system._ensure_engine().counters()
        m = self._single_reaction().index
        tried = np.maximum(c["trials"][:, m], 1)
        rate = c["accepted"][:, m] / tried
        return float(rate[0]) if self._system.n_replicas == 1 else rate
