#!/usr/bin/env python
"""Target of an ncu capture of the trial loop alone: 128 replicas x N=1000 at one pH, 1 mM, k_mc_run launched twice
(capture the second: -k regex:k_mc_run -s 1 -c 1).  python tools/prof_ncu_mc.py [pH] [solo|duo]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
pH = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
w = workloads.Workload(workloads.example_parameters(1000, c_salt=1e-3), pH=np.full(128, pH))
e = Engine(w.cfg); e.set_engine(sys.argv[2] if len(sys.argv) > 2 else "solo"); w.init_engine(e)
e.md_run(3000); e.sample_loop(w.sample_params(e, 10))
nt = 1661 * 4
e.mc_run(0, nt); e.synchronize()
c0 = e.counters(); e.mc_run(0, nt); e.synchronize(); ms = e.last_call_ms()[0]; c1 = e.counters()
print(f"pH {pH}: {ms*1e3/nt:.3f} us/trial, acceptance {(c1['accepted']-c0['accepted']).sum()/(128.0*nt):.3f}")
