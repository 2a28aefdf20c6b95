import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import load, make_workload, random_state
from coarse_grain_polyelectrolyte_b200.engine import Engine
from oracle.oracle import OracleEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w = make_workload(n, pH=[6.0, 8.0, 10.0]); state = random_state(w, seed=n, jitter=0.03)
g = load(Engine(w.cfg), w, state); o = load(OracleEngine(w.cfg), w, state)
fo = o.forces_f64(); fmax = np.abs(fo).max()
for rep in range(3):
    fg = g.md_forces(); d = np.abs(fg - fo); k = np.unravel_index(np.argmax(d), d.shape)
    print("rep", rep, "err", d.max() / fmax, "at", k, "gpu", fg[k[0], k[1]], "oracle", fo[k[0], k[1]], "type", state[2][k[0], k[1]], "q", state[0][k[0], k[1], 3])
fa = g.forces(); print("all-pairs err", np.abs(fa - fo).max() / fmax)
