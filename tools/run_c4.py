"""BASELINE config 4: the C2 chain (N=100, 64 pH points, 20 mM) next to one fixed charged sphere
(sigma = 10 sim_length) for sphere charges -20, 0, +20: titration curves, chain-sphere distance, throughput.
Engine API directly (implicit-ion reactions, sphere in a spectator slot, fixed_type_mask)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine

n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 4571
pH = np.linspace(4.0, 12.0, 64)
rows = {}
for q in (-20.0, 0.0, 20.0):
    w = workloads.config_c4(charge=q, n_ph=64)
    e = Engine(w.cfg); w.init_engine(e)
    e.steepest_descent(200, 0.0, 0.01, 0.02); e.md_run(20000)                     # relax + thermalise
    e.sample_loop(w.sample_params(e, 200))                                          # equilibrate the charges
    c0 = e.counters(); t0 = time.time()
    obs, _ = e.sample_loop(w.sample_params(e, n_samples))
    dt = time.time() - t0; c1 = e.counters()
    x = e.get_state()[0]
    d_com = np.linalg.norm(x[:, :100, :3].mean(1) - x[:, 100, :3], axis=1)
    rows[q] = (obs[..., 0].mean(1) / w.n_sites[0], obs[..., 3].mean(1), d_com)
    steps = (c1["md_steps"] - c0["md_steps"]).sum(); trials = (c1["trials"] - c0["trials"]).sum()
    print(f"q_NP = {q:+.0f}: {n_samples} samples x 64 replicas in {dt:.2f} s: {steps * 101 / dt:.3g} particle-steps/s, "
          f"{trials / sum(w.n_sites) / dt:.3g} MC sweeps/s", flush=True)
print("    pH | alpha(q=-20)  alpha(q=0)  alpha(q=+20)  ideal | Rg(-20)  Rg(0)  Rg(+20) | d_com(-20)  d_com(+20)")
for r in range(0, 64, 4):
    ideal = 1.0 / (1.0 + 10 ** (8.18 - pH[r]))
    print(f"{pH[r]:6.2f} | {rows[-20.0][0][r]:11.4f} {rows[0.0][0][r]:11.4f} {rows[20.0][0][r]:12.4f} {ideal:7.4f} | "
          f"{rows[-20.0][1][r]:7.2f} {rows[0.0][1][r]:6.2f} {rows[20.0][1][r]:8.2f} | {rows[-20.0][2][r]:9.2f} {rows[20.0][2][r]:11.2f}")
