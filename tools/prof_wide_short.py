import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
w = workloads.Workload(workloads.example_parameters(10000, c_salt=2e-2), pH=np.full(16, 8.0))
e = Engine(w.cfg); w.init_engine(e)
e.md_run(120); e.synchronize(); print(e.last_call_ms())
