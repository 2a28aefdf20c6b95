"""Host-side system construction: start configuration, bead types, interaction tables.

These are the one-off setup pieces of the reference's Model (src/modules/Model.py) that do not
belong on the GPU: the self-avoiding random-walk start configuration that ESPResSo's
`polymer.linear_polymer_positions` provides (M.py:116-122), the titratable-site pattern
(M.py:126-155) and the 8x8 WCA table with its combination rules (M.py:58-76, utils.py:2-13).
"""
import numpy as np


def linear_polymer_positions(n_polymers, beads_per_chain, bond_length, seed, min_distance=0.0,
                             box_l=None, max_tries=1000, start_positions=None):
    """Random-walk chain start positions, shape [n_polymers][beads_per_chain][3], unfolded.

    Same contract as espressomd.polymer.linear_polymer_positions as the reference calls it
    (M.py:116-122): each bead sits `bond_length` from its predecessor in a uniformly random
    direction and no two beads are closer than `min_distance`; a bead is redrawn up to
    `max_tries` times before the whole chain is restarted.  The random stream is numpy's
    Philox seeded with `seed` (ESPResSo's own stream is not reproducible outside ESPResSo).
    """
    if n_polymers < 1 or beads_per_chain < 1:
        raise ValueError("n_polymers and beads_per_chain must be positive")
    if bond_length <= 0:
        raise ValueError("bond_length must be positive")
    rng = np.random.Generator(np.random.Philox(int(seed)))
    box = float(box_l) if box_l is not None else None
    if min_distance <= 0.0 and start_positions is None:   # free random walks: no rejection, vectorised
        steps = rng.normal(size=(n_polymers, beads_per_chain, 3))
        steps *= bond_length / np.linalg.norm(steps, axis=2, keepdims=True)
        steps[:, 0] = rng.random((n_polymers, 3)) * (box if box is not None else 1.0)
        return np.cumsum(steps, axis=1)
    out = np.empty((n_polymers, beads_per_chain, 3))
    placed = np.empty((n_polymers * beads_per_chain, 3))
    n_placed = 0
    min2 = float(min_distance) ** 2

    def too_close(p, upto):
        if min2 <= 0.0 or upto == 0:
            return False
        d = placed[:upto] - p
        if box is not None:
            d -= box * np.rint(d / box)
        return bool(np.any(np.einsum("ij,ij->i", d, d) < min2))

    for c in range(n_polymers):
        for _attempt in range(max_tries):
            base = n_placed
            ok = True
            if start_positions is not None:
                first = np.asarray(start_positions[c], dtype=float)
            else:
                for _ in range(max_tries):
                    first = rng.random(3) * (box if box is not None else 1.0)
                    if not too_close(first, base):
                        break
                else:
                    ok = False
            if ok:
                placed[base] = first
                k = 1
                while k < beads_per_chain:
                    for _ in range(max_tries):
                        v = rng.normal(size=3)
                        v *= bond_length / np.linalg.norm(v)
                        cand = placed[base + k - 1] + v
                        if not too_close(cand, base + k):
                            placed[base + k] = cand
                            break
                    else:
                        ok = False
                        break
                    k += 1
            if ok:
                out[c] = placed[base:base + beads_per_chain]
                n_placed = base + beads_per_chain
                break
        else:
            raise RuntimeError("could not place polymer %d without overlaps" % c)
    return out


def chain_site_pattern(n_mon):
    """Per-bead role along one chain (M.py:126-155): every third bead is titratable, the ones
    at either chain end belong to the second species.  Returns an int array with
    0 = inert backbone bead ("CH2"), 1 = site of reaction 1 ("N"), 2 = site of reaction 2 ("N2")."""
    role = np.zeros(n_mon, dtype=np.int64)
    idx = np.arange(n_mon)
    site = idx % 3 == 0
    role[site] = 1
    role[site & ((idx == 0) | (idx == n_mon - 1))] = 2
    return role


def combination_sigma(sig1, sig2):
    """Arithmetic mean (utils.py:9-11)."""
    return 0.5 * (sig1 + sig2)


def combination_epsilon(eps1, eps2):
    """Geometric mean (utils.py:2-4)."""
    return (eps1 * eps2) ** 0.5


def wca_tables(type_index, sigma, epsilon):
    """[T][T] epsilon and sigma tables exactly as the reference fills them (M.py:58-76):
    eps = sqrt(eps1 eps2), sigma = 0.5 * (sig1 + sig2)/2."""
    n = max(type_index.values()) + 1
    eps, sig = np.zeros((n, n)), np.zeros((n, n))
    for a, ia in type_index.items():
        for b, ib in type_index.items():
            eps[ia, ib] = combination_epsilon(epsilon[a], epsilon[b])
            sig[ia, ib] = 0.5 * combination_sigma(sigma[a], sigma[b])
    return eps, sig
