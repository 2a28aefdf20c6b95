import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "src"), ROOT]
os.environ.setdefault("CGPE_QUIET", "1")
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from modules.EnsembleDynamics import EnsembleDynamics
for n in (1000, 300, 100):
    t0 = time.time()
    ens = EnsembleDynamics(workloads.example_parameters(n, c_salt=2e-2, **{"Montecarlo properties": {"N_BLOCKS": 1}}), replicas=dict(pH=[6.0, 8.0, 10.0], c_salt=[2e-2] * 3))
    e = ens.system.analysis.energy()
    print(n, "warm-up %.1f s" % (time.time() - t0), "min_dist", ens.system.analysis.min_dist(), "E bonded", np.round(e["bonded"], 0), "T", np.round(2 / 3 * e["kinetic"] / n, 2), flush=True)
    ens.equilibrate_pH(); ens.perform_sampling()
    print("   sampling ok: alpha", np.round(ens.titration_curve()[0], 3), "rg", np.round(ens.rad_gyr.mean(1), 1), "time %.1f s" % (time.time() - t0), flush=True)
