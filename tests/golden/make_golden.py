"""Generates the golden fixtures under tests/golden/ (committed next to this script).

The reference holds no golden vector for its hot path (its arithmetic lives in ESPResSo, which
cannot be run here), so these are the vectors SURVEY.md §8c asks the build to create:

  few_body.npz      2-, 3- and 4-particle configurations with U and F from closed-form numpy
                    expressions written here (independent of oracle/cgpe_oracle.c)
  chain_n52.npz     frozen N=52 chain (the notebook's size): FP64 oracle energy breakdown + forces
  chain_n100.npz    frozen N=100 chain (examples/run_test.py size): same
  chain_n100_np.npz the N=100 chain next to a fixed sphere of charge -20 (BASELINE config 4): same
  trials_n100.npz   10^4-trial proposal log of reaction 1 (site, direction, u, Delta-E, decision)
  titration.npz     ideal titration curve at 64 pH points (utils.py:15-16 closed form)

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from coarse_grain_polyelectrolyte_b200 import _abi  # noqa: E402
from helpers import load, make_workload, random_state  # noqa: E402
from oracle.oracle import OracleEngine  # noqa: E402


# ---- closed forms ----------------------------------------------------------------------------------
def wca(eps, sig, r):
    if r >= 2 ** (1 / 6) * sig:
        return 0.0, 0.0
    s6 = (sig / r) ** 6
    return 4 * eps * (s6 * s6 - s6 + 0.25), 24 * eps * (2 * s6 * s6 - s6) / r      # U, -dU/dr


def dh(pref, qi, qj, kappa, rc, r):
    if r >= rc:
        return 0.0, 0.0
    u = pref * qi * qj * np.exp(-kappa * r) / r
    return u, u * (kappa + 1 / r)


def few_body():
    out = {}
    eps, sig, pref, kappa, rc = 0.7, 1.1, 2.0, 0.3, 5.0
    L = 40.0
    # (a) two particles along x: WCA + DH + harmonic bond, all analytic
    r = 1.05
    k, r0 = 300.0, 0.9
    uw, fw = wca(eps, sig, r)
    ud, fd = dh(pref, 1.0, 1.0, kappa, rc, r)
    ub, fb = 0.5 * k * (r - r0) ** 2, -k * (r - r0)
    out["pair_x"] = np.array([[10.0, 10.0, 10.0, 1.0], [10.0 + r, 10.0, 10.0, 1.0]])
    out["pair_energy"] = np.array([uw + ud + ub, 0.0, ud, ub, uw])          # total, kin, coul, bonded, nb
    f = fw + fd + fb                                                         # on particle 1 along +x
    out["pair_force"] = np.array([[-f, 0, 0], [f, 0, 0]])
    out["pair_params"] = np.array([eps, sig, pref, kappa, rc, L, k, r0])
    # (b) the same pair across the periodic boundary (minimum image): identical energy
    out["pair_pbc_x"] = np.array([[0.3, 5.0, 5.0, 1.0], [L - (r - 0.3), 5.0, 5.0, 1.0]])
    # (c) three beads, harmonic angle only: right angle, phi0 = 2 pi/3
    kb, phi0 = 6.5, 2 * np.pi / 3
    a = 1.2
    out["angle_x"] = np.array([[a, 0, 0, 0], [0, 0, 0, 0], [0, a, 0, 0.0]]) + np.array([5, 5, 5, 0.0])
    dphi = np.pi / 2 - phi0
    out["angle_energy"] = 0.5 * kb * dphi ** 2
    # F_left = -dU/dphi * dphi/dr_left ; for a right angle dphi/dr_left = -e_y/a (moving the left bead along +y closes the angle)
    g = kb * dphi / a
    out["angle_force"] = np.array([[0, g, 0], [-g, -g, 0], [g, 0, 0]])
    out["angle_params"] = np.array([kb, phi0])
    # (d) four beads, dihedral only: planar cis (phi=0) and trans (phi=pi), U = k[1-cos(3 phi - pi)]
    kd = 2.0
    cis = np.array([[1, 1, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0.0]], float)     # placeholder, fixed below
    cis = np.array([[0, 1, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0], [1, 1, 0, 0.0]], float)
    trans = np.array([[0, 1, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0], [1, -1, 0, 0.0]], float)
    gauche = np.array([[0, 1, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0], [1, np.cos(np.pi / 3), np.sin(np.pi / 3), 0.0]], float)
    out["dih_cis_x"], out["dih_trans_x"], out["dih_gauche_x"] = cis + [5, 5, 5, 0], trans + [5, 5, 5, 0], gauche + [5, 5, 5, 0]
    out["dih_energy"] = np.array([kd * (1 - np.cos(0 - np.pi)), kd * (1 - np.cos(3 * np.pi - np.pi)), kd * (1 - np.cos(np.pi - np.pi))])
    out["dih_params"] = np.array([kd, 3, np.pi])
    return out


def chain(n_mon, seed, np_charge=None):
    """np_charge: BASELINE config 4, one fixed sphere (sigma = 10 sim_length) of that charge in contact with the chain."""
    kw = {} if np_charge is None else dict(nanoparticle=dict(charge=np_charge, clearance=0.9))
    w = make_workload(n_mon, pH=[8.0], **kw)
    state = random_state(w, seed=seed, jitter=0.03)
    o = load(OracleEngine(w.cfg), w, state)
    out = dict(n_mon=n_mon, seed=seed, xyzq=state[0], vel=state[1], type=state[2], energy=o.energy(),
               forces=o.forces_f64(), min_dist=o.min_dist(), rg=o.observables()[0], ree=o.observables()[1])
    if np_charge is not None:
        out["np_charge"] = np_charge
    return out


def trials(n_mon=100, n_trials=10000, seed=17):
    w = make_workload(n_mon, pH=[8.3])
    state = random_state(w, seed=seed, jitter=0.03)
    o = load(OracleEngine(w.cfg), w, state)
    rec = o.mc_run(0, n_trials, trace=True)
    _, _, typ = o.get_state()
    return dict(n_mon=n_mon, seed=seed, pH=8.3, xyzq=state[0], vel=state[1], type=state[2], trace=rec,
                final_type=typ, counts=o.observables()[2])


def titration():
    pH = np.linspace(4.0, 12.0, 64)
    return dict(pH=pH, alpha_pK1=1.0 / (1.0 + 10.0 ** (8.18 - pH)), alpha_pK2=1.0 / (1.0 + 10.0 ** (10.02 - pH)))


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "few_body.npz"), **few_body())
    np.savez_compressed(os.path.join(HERE, "chain_n52.npz"), **chain(52, 52))
    np.savez_compressed(os.path.join(HERE, "chain_n100.npz"), **chain(100, 100))
    np.savez_compressed(os.path.join(HERE, "chain_n100_np.npz"), **chain(100, 101, np_charge=-20.0))
    np.savez_compressed(os.path.join(HERE, "trials_n100.npz"), **trials())
    np.savez_compressed(os.path.join(HERE, "titration.npz"), **titration())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
