import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
c_salt = float(sys.argv[1]); pH = float(sys.argv[2]); R = 16
w = workloads.Workload(workloads.example_parameters(1000, c_salt=c_salt), pH=np.full(R, pH))
e = Engine(w.cfg); w.init_engine(e)
e.md_run(3000)
print("start", e.energy()[0], e.observables()[0][:3])
for it in range(6):
    obs, _ = e.sample_loop(w.sample_params(e, 4))
    en = e.energy()
    print(it, "E", np.round(en[0], 1), "rg", np.round(obs[:3, -1, 3], 2), "alpha", obs[0, -1, 0] / w.n_sites[0], "maxE", en[:, 0].max(), "reb", e.counters()["list_rebuilds"][:3])
