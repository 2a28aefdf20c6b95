"""Correctness part 3 of the north star at the BASELINE sizes: ensemble observables alpha(pH), <Rg>, <Ree> of the CUDA
engine agree with the CPU oracle's within statistical error bars.

Fixtures tests/golden/ensemble_c2.npz / ensemble_c3.npz come from tests/golden/make_ensemble_golden.py (oracle, CPU):
  C2  N = 100, 20 mM, dihedral ON, 8 pH points, the full 4571-sample protocol (E.py:75-119), 8 independent starts;
  C3  N = 1000, 1 mM and 20 mM, 4 pH points, 200 samples, 8 independent starts.
The CUDA engine runs the same replicas (same starts, same protocol) on OTHER Philox streams (replica_offset shifted),
so the comparison is statistical, not a trajectory match.  Error bars: binning analysis per replica (NB:727-731) and the
scatter over the independent starts; a deviation must stay below 4 combined standard errors:
  alpha, alpha_2   z = |mean_gpu - mean_cpu| / sqrt(sem_gpu^2 + sem_cpu^2), sem = max(scatter over starts, binning) / sqrt(8)
  Rg, Ree          paired over the starts (both sides start from the same 8 chains, whose torsional states are frozen on
                   the run's time scale -- 13 kT barriers -- so the start dominates the value): d_s = gpu_s - cpu_s,
                   z = |mean d| / sqrt(var(d)/8 + mean binning variances/8)
"""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_ensemble_golden", os.path.join(GOLD, "make_ensemble_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

Z_MAX = 4.0


@pytest.mark.parametrize("spec", [gen.C2, gen.C3], ids=["C2_N100_full_protocol", "C3_N1000"])
def test_ensemble_observables_at_baseline_sizes(spec):
    from coarse_grain_polyelectrolyte_b200.engine import Engine
    path = os.path.join(GOLD, f"ensemble_{spec['name']}.npz")
    gold = np.load(path)
    w, state, grid = gen.build(spec, replica_offset=100000)        # other random streams than the fixture's
    np.testing.assert_allclose(grid, gold["grid"])
    obs = gen.run(Engine(w.cfg), w, state, spec)
    assert obs.shape[1] == int(gold["n_samples"])
    if spec["name"] == "c2":
        assert obs.shape[1] == 4571                                 # the full protocol of examples/run_test.py
    np.testing.assert_array_equal(obs[..., 2], obs[..., 0] + obs[..., 1])   # N_B = N_N + N_N2 (NB:253-254)
    stats = gen.summarise(obs, w)
    S = gen.N_STARTS
    report, worst = [], 0.0
    for r in range(0, len(grid), S):
        salt, pH = grid[r][0], grid[r][1]
        for key in ("alpha", "alpha2"):
            g, c = stats[key][r:r + S], gold[key][r:r + S]
            sem = [max(x[:, 0].std(ddof=1), np.sqrt(np.mean(x[:, 1] ** 2))) / np.sqrt(S) for x in (g, c)]
            z = abs(g[:, 0].mean() - c[:, 0].mean()) / np.sqrt(sem[0] ** 2 + sem[1] ** 2 + 1e-8)
            report.append((salt, pH, key, g[:, 0].mean(), c[:, 0].mean(), z))
        for key in ("rg", "ree"):
            g, c = stats[key][r:r + S], gold[key][r:r + S]
            d = g[:, 0] - c[:, 0]
            err = np.sqrt(d.var(ddof=1) / S + np.mean(g[:, 1] ** 2 + c[:, 1] ** 2) / S)
            z = abs(d.mean()) / err
            report.append((salt, pH, key, g[:, 0].mean(), c[:, 0].mean(), z))
    text = "\n".join(f"salt {s:g} pH {p:5.2f} {k:6s} cuda {a:9.4f} oracle {b:9.4f} z {z:4.2f}" for s, p, k, a, b, z in report)
    print(text)
    worst = max(z for *_, z in report)
    assert worst < Z_MAX, text
    # the titration really is sampled: alpha spans the curve over the pH points
    a = np.array([x[3] for x in report if x[2] == "alpha"])
    assert a.min() < 0.2 and a.max() > 0.9
