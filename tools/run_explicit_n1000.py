import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "src"), ROOT]
os.environ.setdefault("CGPE_QUIET", "1")
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from modules.EnsembleDynamics import EnsembleDynamics
n_mon = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
params = workloads.example_parameters(n_mon, c_salt=2e-2, **{"Montecarlo properties": {"N_BLOCKS": 1, "DESIRED_BLOCK_SIZE": 35, "PROB_REACTION": 0.7}})
params["Simulation configuration"]["ION_MODEL"] = "explicit"
pH = np.linspace(5.0, 11.0, R)
t0 = time.time()
ens = EnsembleDynamics(params, replicas=dict(pH=pH, c_salt=[2e-2] * R))
t1 = time.time()
print(f"N={n_mon} explicit ions: {len(ens.system.part)} particles/replica, {R} replicas; setup+warm-up {t1-t0:.1f} s; min_dist {np.round(ens.system.analysis.min_dist(), 2)}", flush=True)
ens.equilibrate_pH(); ens.perform_sampling()
t2 = time.time()
c = ens.system._engine.counters()
print(f"{ens.num_samples} samples: {t2-t1:.1f} s = {(t2-t1)/ens.num_samples*1e3:.1f} ms/sample; MD steps/replica {c['md_steps'].mean():.0f}; alpha {np.round(ens.titration_curve()[0], 3)}")
print("N_B == N_N + N_N2:", bool(np.all(ens.num_B == ens.num_As + ens.num_As2)), "Rg", np.round(ens.rad_gyr.mean(-1), 1))
