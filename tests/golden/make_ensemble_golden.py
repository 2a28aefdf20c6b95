#!/usr/bin/env python
"""Ensemble observables of the CPU oracle at the BASELINE sizes -> tests/golden/ensemble_c2.npz, ensemble_c3.npz.

    C2: N = 100, 20 mM, dihedral ON (examples/run_test.py as shipped), 8 of the 64 pH points of linspace(4, 12, 64),
        the FULL protocol of E.py:75-119 (4571 samples: ~1.6 M MD steps + 0.79 M trial moves per replica),
        8 independent starts (chain random walks with seeds 100..107).
    C3: N = 1000, 1 mM and 20 mM, 4 of the 128 pH points of linspace(2, 12, 128), 200 samples of the protocol
        (a full run is ~1.6 G bead-steps per replica: hours on a CPU), 8 independent starts.

Stored per replica (config, salt, pH, start): mean and binning error (NB:727-731) of alpha = n_N / n_nh,
alpha_2 = n_N2 / n_nh2, Rg, Ree after dropping the first tenth of the samples.  tests/test_gpu_ensemble.py runs the
same replicas on the CUDA engine with other random streams and compares within the error bars.

Run here (CPU, minutes):  python tests/golden/make_ensemble_golden.py [c2|c3|all]
The oracle is test infrastructure; nothing here is imported by the product.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from coarse_grain_polyelectrolyte_b200 import workloads  # noqa: E402

N_STARTS = 8
C2 = dict(name="c2", n_mon=100, salts=(2e-2,), pH=np.linspace(4.0, 12.0, 64)[[4, 20, 28, 33, 38, 44, 52, 60]], n_samples=None, warm=4000)
C3 = dict(name="c3", n_mon=1000, salts=(1e-3, 2e-2), pH=np.linspace(2.0, 12.0, 128)[[13, 51, 79, 115]], n_samples=200, warm=4000)


def binning_sem(x, n_bins=16):
    """NB:727-731: std(bin means, ddof=1.5) / sqrt(n_bins), along the last axis."""
    n = (x.shape[-1] // n_bins) * n_bins
    means = x[..., :n].reshape(x.shape[:-1] + (n_bins, -1)).mean(axis=-1)
    return np.std(means, axis=-1, ddof=1.5) / np.sqrt(n_bins)


def build(spec, replica_offset):
    """One handle holding every (salt, pH, start) replica of the spec: (workload, state, labels)."""
    salts, pH = np.asarray(spec["salts"]), np.asarray(spec["pH"])
    grid = [(s, p, k) for s in salts for p in pH for k in range(N_STARTS)]
    w = workloads.Workload(workloads.example_parameters(spec["n_mon"], c_salt=float(salts[0])), pH=[g[1] for g in grid],
                           c_salt=[g[0] for g in grid], replica_offset=replica_offset)
    P = w.xyzq.shape[0]
    xyzq = np.empty((len(grid), P, 4), np.float32)
    starts = {}
    for k in range(N_STARTS):                      # start k: its own self-avoiding walk
        wk = workloads.Workload(workloads.example_parameters(spec["n_mon"], c_salt=float(salts[0])), pH=[7.0], seed=100 + k)
        starts[k] = wk.xyzq
    for r, (_, _, k) in enumerate(grid):
        xyzq[r] = starts[k]
    types = np.broadcast_to(w.type, (len(grid), P)).copy()
    return w, (xyzq, None, types), np.array(grid)


def run(engine, w, state, spec):
    engine.set_state(*state)
    engine.set_replica_params(pH=w.pH, kappa=w.kappa, r_cut=w.r_cut)
    engine.set_thermostat_seed(1434)
    engine.md_run(spec["warm"])
    n = spec["n_samples"] or w.num_samples
    obs, _ = engine.sample_loop(w.sample_params(engine, n))
    return obs


def summarise(obs, w):
    skip = obs.shape[1] // 10
    cols = {"alpha": obs[:, skip:, 0] / w.n_sites[0], "alpha2": obs[:, skip:, 1] / max(w.n_sites[1], 1),
            "rg": obs[:, skip:, 3], "ree": obs[:, skip:, 4]}
    return {k: np.stack([v.mean(axis=1), binning_sem(v)], axis=1) for k, v in cols.items()}


def main():
    from oracle import oracle
    oracle.build()
    oracle.set_threads(os.cpu_count() or 1)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    for spec in (C2, C3):
        if which not in ("all", spec["name"]):
            continue
        w, state, grid = build(spec, replica_offset=0)
        o = oracle.OracleEngine(w.cfg)
        o.set_list_mode(True, 0.8)
        t0 = time.time()
        obs = run(o, w, state, spec)
        stats = summarise(obs, w)
        path = os.path.join(HERE, f"ensemble_{spec['name']}.npz")
        np.savez_compressed(path, grid=grid, n_samples=obs.shape[1], **stats)
        print(f"{spec['name']}: {len(grid)} replicas x {obs.shape[1]} samples in {time.time() - t0:.0f} s -> {path}", flush=True)
        for r in range(0, len(grid), N_STARTS):
            a, rg = stats["alpha"][r:r + N_STARTS, 0], stats["rg"][r:r + N_STARTS, 0]
            print(f"   salt {grid[r][0]:g} pH {grid[r][1]:5.2f}: alpha {a.mean():.4f} +- {a.std(ddof=1) / np.sqrt(N_STARTS):.4f}  Rg {rg.mean():.2f} +- {rg.std(ddof=1) / np.sqrt(N_STARTS):.2f}")


if __name__ == "__main__":
    main()
