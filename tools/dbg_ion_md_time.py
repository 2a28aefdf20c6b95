import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
R = int(os.environ.get('R', 128))
w = workloads.Workload(workloads.example_parameters(1000, c_salt=2e-2), pH=np.full(R, 8.0), explicit_ions=True)
t0 = time.time(); e = Engine(w.cfg); t1 = time.time(); w.init_engine(e); t2 = time.time()
print(f"create {t1-t0:.2f} s, init {t2-t1:.2f} s", flush=True)
for n in (10, 150, 150, 1000):
    t = time.time(); e.md_run(n); e.synchronize(); dt = time.time() - t
    print(n, "steps: wall", round(dt, 3), "s; last_call_ms", e.last_call_ms(), "rebuilds", e.counters()['list_rebuilds'].mean(), flush=True)
