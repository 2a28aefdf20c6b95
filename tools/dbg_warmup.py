import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "src"), ROOT]
os.environ.setdefault("CGPE_QUIET", "1")
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from modules.Model import Model
params = workloads.example_parameters(1000, c_salt=2e-2)
m = Model(params, replicas=dict(pH=[8.0, 9.0], c_salt=[2e-2, 2e-2]))
s = m.system
def closest():
    x = s.part.all().pos[0]; t = s.part.all().type[0]
    d = x[:, None, :] - x[None, :, :]; r = np.sqrt((d ** 2).sum(-1)) + np.eye(len(x)) * 99
    i, j = np.unravel_index(np.argmin(r), r.shape)
    return r[i, j], i, j, t[i], t[j], np.sort(r[i])[:4]
print("after model setup: min_dist", s.analysis.min_dist(), closest(), "E", {k: np.round(v, 1) for k, v in s.analysis.energy().items()})
s.force_cap = 5
s.thermostat.set_langevin(kT=0.0, gamma=1.0, seed=4)
cap = 5
for rnd in range(12):
    for _ in range(30): s.integrator.run(12)
    cap += 1; s.force_cap = cap
    print(rnd, "min_dist", s.analysis.min_dist(), closest()[:5], "Ekin", np.round(s.analysis.energy()["kinetic"], 2))
