#!/usr/bin/env python
"""Target of an ncu capture of the MD kernel alone: 128 replicas x N=1000 at one pH, 1 mM, solo engine; k_md_run launches:
4000 (equilibration), 20-sample loop, 500, then 300 steps (capture the last: -k regex:k_md_run -s 2 -c 1).
python tools/prof_ncu_md.py [pH]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
pH = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
w = workloads.Workload(workloads.example_parameters(1000, c_salt=1e-3), pH=np.full(128, pH))
e = Engine(w.cfg); e.set_engine("solo"); w.init_engine(e)
e.md_run(4000); e.sample_loop(w.sample_params(e, 20)); e.md_run(500); e.synchronize()
e.md_run(300); e.synchronize(); ms = e.last_call_ms()[0]
print(f"pH {pH}: {ms/300*1e3:.2f} us/step; list stats per replica {np.round(e.list_stats().mean(0))}")
