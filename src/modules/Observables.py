"""Observables — registry of accumulators (reference: src/modules/Observables.py).

Same methods as the reference's class; the accumulators come from the espressomd-shaped facade of
this package (`espressomd.observables`, `espressomd.accumulators`) and are updated every
`delta_N = 100` steps of `integrator.run` once `update_accumulators(system)` has registered them.
`rdf_observable_corrected`, which in the reference lacks `self` and reads an undefined name
(O.py:31-52), is a working static method here.
"""
import numpy as np

from coarse_grain_polyelectrolyte_b200.espressomd import accumulators, observables


class Observables:
    def __init__(self):
        self.observables = {}

    def add_observable(self, key_name, observable_class, *args, **kwargs):      # O.py:9-11
        self.observables[key_name] = observable_class(*args, **kwargs)

    def update_accumulators(self, system):                                      # O.py:13-15
        for obs in self.observables.values():
            system.auto_update_accumulators.add(obs)

    def clear_accumulators(self, system):                                       # O.py:18-20
        system.auto_update_accumulators.clear()
        self.observables.clear()

    def rdf_observable(self, ids1, ids2, r_min, r_max, n_r_bins):               # O.py:27-29
        obs = observables.RDF(ids1=ids1, ids2=ids2, min_r=r_min, max_r=r_max, n_r_bins=n_r_bins)
        return accumulators.TimeSeries(obs=obs, delta_N=100)

    @staticmethod
    def rdf_observable_corrected(distances, nA, nB, box_size, bin_count=100):   # O.py:31-52
        """RDF of a set of pair distances: (bin centres, g(r)) with the shells up to box_size/2
        normalised by the density of the distances handed in."""
        distances = np.asarray(distances, dtype=np.float64).ravel()
        bin_edges = np.linspace(0, box_size / 2, bin_count + 1)
        bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2
        bin_volume = (4 / 3) * np.pi * (bin_edges[1:] ** 3 - bin_edges[:-1] ** 3)
        counts, _ = np.histogram(distances, bins=bin_edges)
        density = len(distances) / (box_size ** 3)
        return bin_centers, counts / (bin_volume * density)

    def particle_distance_observable(self, ids):                                # O.py:54-56
        return accumulators.TimeSeries(obs=observables.ParticleDistances(ids=ids), delta_N=100)

    def dihedral_observable(self, ids):                                         # O.py:58-60
        return accumulators.TimeSeries(obs=observables.BondDihedrals(ids=ids), delta_N=100)

    def position_observable(self, ids):                                         # O.py:62-64
        return accumulators.TimeSeries(obs=observables.ParticlePositions(ids=ids), delta_N=100)

    def com_pos_cor(self, pids_monomers, tau_max):                              # O.py:66-70
        com_pos = observables.ComPosition(ids=pids_monomers)
        return accumulators.Correlator(obs1=com_pos, tau_lin=16, tau_max=tau_max, delta_N=100,
                                       corr_operation="square_distance_componentwise", compress1="discard1")

    def com_vel_cor(self, pids_monomers, tau_max):                              # O.py:72-76
        com_vel = observables.ComVelocity(ids=pids_monomers)
        return accumulators.Correlator(obs1=com_vel, tau_lin=16, tau_max=tau_max, delta_N=100,
                                       corr_operation="scalar_product", compress1="discard1")
