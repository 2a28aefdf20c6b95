"""Per-kernel times of MD steps of the global-memory engine: run under
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/prof_wide_launches.py
(40 steps: below the CUDA-graph threshold, so every kernel of a step is a launch of its own)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine

n_mon = int(os.environ.get("N_MON", 10000)); R = int(os.environ.get("R", 128))
if os.environ.get("N_POLY"):   # config 5: one box of several chains with explicit ions (tools/run_c5.py)
    w = workloads.Workload(workloads.example_parameters(n_mon, c_salt=float(os.environ.get("C_SALT", 1e-2)), n_poly=int(os.environ["N_POLY"])),
                           pH=np.linspace(7.0, 9.0, R), explicit_ions=True, min_distance=0.0)
    w.cfg.cap_wca = 200
    e = Engine(w.cfg); w.init_engine(e)
    for cap in (0.02, 0.05, 0.1):
        e.steepest_descent(300, gamma=0.05 / w.ru.k_bond, max_displacement=cap)
    e.set_force_cap(200.0); e.md_run(500); e.set_force_cap(0.0); e.md_run(1000); e.synchronize()
else:
    w = workloads.Workload(workloads.example_parameters(n_mon, c_salt=float(os.environ.get("C_SALT", 1e-3))), pH=np.linspace(2.0, 12.0, R))
    e = Engine(w.cfg); w.init_engine(e)
    e.md_run(500); e.sample_loop(w.sample_params(e, 2)); e.md_run(100); e.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e.md_run(int(os.environ.get("STEPS", 40))); e.synchronize()
if os.environ.get("MC"):
    e.mc_run(0, int(os.environ["MC"])); e.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("device ms", e.last_call_ms())
