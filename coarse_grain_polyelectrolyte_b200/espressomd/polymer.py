"""espressomd.polymer.linear_polymer_positions (M.py:116-122)."""
from ..topology import linear_polymer_positions as _lpp


def linear_polymer_positions(n_polymers, beads_per_chain, bond_length, seed, min_distance=0.0,
                             max_tries=1000, start_positions=None, **unsupported):
    if unsupported:
        raise NotImplementedError(f"arguments not supported: {sorted(unsupported)}")
    from . import _current_system
    box = float(_current_system.box_l[0]) if _current_system is not None else None
    return _lpp(n_polymers, beads_per_chain, bond_length, seed, min_distance=min_distance, box_l=box,
                max_tries=max_tries, start_positions=start_positions)
