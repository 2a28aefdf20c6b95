"""Shared builders for the parity tests: seeded synthetic states fed identically to the CUDA
engine and to the CPU oracle."""
import copy

import numpy as np

from coarse_grain_polyelectrolyte_b200 import workloads


def make_workload(n_mon, pH, c_salt=2e-2, flags=None, replica_offset=0, seed=12, explicit_ions=False, n_poly=1,
                  nanoparticle=None, **cfg_overrides):
    params = workloads.example_parameters(n_mon, c_salt=c_salt, n_poly=n_poly)
    if flags:
        params = copy.deepcopy(params)
        params["Simulation configuration"].update(flags)
    w = workloads.Workload(params, pH=pH, replica_offset=replica_offset, seed=seed, explicit_ions=explicit_ions,
                           nanoparticle=nanoparticle)
    for k, v in cfg_overrides.items():
        setattr(w.cfg, k, v)
    return w


def random_state(w, seed, protonated_fraction=0.5, jitter=0.05, vel_sigma=1.0):
    """Per-replica perturbed copies of the workload's start state with a random protonation
    pattern (types and charges consistent) and Maxwell-Boltzmann-like velocities."""
    rng = np.random.default_rng(seed)
    R, P = w.R, w.xyzq.shape[0]
    tix, ru = w.ru.type_index, w.ru
    xyzq = np.broadcast_to(w.xyzq, (R, P, 4)).copy()
    xyzq[..., :3] += rng.normal(scale=jitter, size=(R, P, 3)).astype(np.float32)
    types = np.broadcast_to(w.type, (R, P)).copy()
    for a, b in (("N", "NH"), ("N2", "NH2")):
        sites = types == tix[a]
        flip = sites & (rng.random((R, P)) < protonated_fraction)
        types[flip] = tix[b]
    charge_of_type = np.zeros(len(tix) + 1, np.float32)
    for name, t in tix.items():
        charge_of_type[t] = ru.charge[name]
    xyzq[..., 3] = charge_of_type[types]
    vel = np.zeros((R, P, 4), np.float32)
    vel[..., :3] = rng.normal(scale=vel_sigma, size=(R, P, 3))
    return xyzq.astype(np.float32), vel, types.astype(np.uint8)


def load(engine, w, state):
    xyzq, vel, types = state
    engine.set_state(xyzq, vel, types)
    engine.set_replica_params(pH=w.pH, kappa=w.kappa, r_cut=w.r_cut)
    engine.set_thermostat_seed(1434)
    return engine
