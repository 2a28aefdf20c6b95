import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from helpers import load, make_workload, random_state
from coarse_grain_polyelectrolyte_b200.engine import Engine
from oracle.oracle import OracleEngine
w = make_workload(100, pH=[8.0, 9.0])
state = random_state(w, seed=5, jitter=0.1, vel_sigma=0.0)
g = load(Engine(w.cfg), w, state); o = load(OracleEngine(w.cfg), w, state)
fg = g.md_forces(); fo = o.forces_f64()
print("force err", np.abs(fg-fo).max()/np.abs(fo).max())
for n in range(12):
    g.steepest_descent(1, 0.0, 1e-4, 0.01); o.steepest_descent(1, 0.0, 1e-4, 0.01)
    xg = g.get_state()[0][..., :3]; xo = o.state_f64()[0]
    d = np.abs(xg-xo); k = np.unravel_index(np.argmax(d), d.shape)
    fo = o.forces_f64()
    print(n, "maxdiff", d.max(), "at", k, "force there", fo[k], "E", o.energy()[:,0], g.energy()[:,0])
