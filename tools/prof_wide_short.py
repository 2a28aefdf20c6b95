import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
n_mon = int(os.environ.get('N_MON', 10000)); R = int(os.environ.get('R', 16))
w = workloads.Workload(workloads.example_parameters(n_mon, c_salt=2e-2), pH=np.full(R, 8.0), explicit_ions=bool(os.environ.get('IONS')))
if os.environ.get('SKIN'): w.cfg.skin = float(os.environ['SKIN'])
e = Engine(w.cfg); w.init_engine(e)
import time
for _ in range(2):
    t = time.time(); e.md_run(int(os.environ.get('STEPS', 120))); e.synchronize(); print('wall', round(time.time() - t, 3), 's; device', e.last_call_ms(), flush=True)
