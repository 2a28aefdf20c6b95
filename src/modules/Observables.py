"""Observables — registry of accumulators (reference: src/modules/Observables.py).

The reference's registry wraps ESPResSo accumulators (RDF, particle distances, dihedrals,
positions, COM correlators) and is used from its notebook only, never by EnsembleDynamics; it is
outside the hot path implemented here.  The class keeps the registry
interface so scripts that construct it keep working; asking for an accumulator says what is
missing instead of failing on an import.
"""


class Observables:
    def __init__(self):
        self.observables = {}

    def add_observable(self, key_name, observable_class, *args, **kwargs):
        self.observables[key_name] = observable_class(*args, **kwargs)

    def update_accumulators(self, system):
        if self.observables:
            raise NotImplementedError("auto-updated accumulators are not part of the B200 hot path yet")

    def clear_accumulators(self, system):
        self.observables.clear()

    def _missing(self, *_, **__):
        raise NotImplementedError("ESPResSo accumulators (RDF, distances, dihedrals, positions, correlators) "
                                  "are not part of the B200 hot path yet (SURVEY.md 8f)")

    rdf_observable = particle_distance_observable = dihedral_observable = position_observable = _missing
    com_pos_cor = com_vel_cor = _missing
