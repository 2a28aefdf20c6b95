"""One short launch of each hot kernel on the bench workload (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
n_mon = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 128
pH = float(sys.argv[3]) if len(sys.argv) > 3 else 8.0
w = workloads.Workload(workloads.example_parameters(n_mon, c_salt=2e-2), pH=np.full(R, pH))
e = Engine(w.cfg); w.init_engine(e)
e.md_run(3000); e.mc_run(0, 20000); e.mc_run(1, 500); e.md_run(1000); e.synchronize()
e.md_run(300); e.synchronize(); print("md_run(300) ms", e.last_call_ms()[0])
e.mc_run(0, 5 * w.n_sites[0] + 1); e.synchronize(); print("mc_run ms", e.last_call_ms()[0])
e.sample_loop_async(w.sample_params(e, 4)); e.synchronize(); print("sample_loop(4) ms", e.last_call_ms()[0])
