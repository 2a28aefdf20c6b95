"""Where the end-to-end call sequence of bench.py spends its wall time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
w = workloads.config_c3_shard(0, 1)
e = Engine(w.cfg); w.init_engine(e)
e.md_run(4000); e.sample_loop(w.sample_params(e, 30))
sp = w.sample_params(e, 100)
for rep in range(3):
    t = [time.perf_counter()]
    e.sample_loop_async(sp); e.synchronize(); t.append(time.perf_counter()); k0 = e.last_call_ms()[0]
    st = e.get_state(); t.append(time.perf_counter())
    e.set_state(*st); t.append(time.perf_counter())
    e.set_replica_params(pH=w.pH, kappa=w.kappa, r_cut=w.r_cut); t.append(time.perf_counter())
    obs, _ = e.sample_loop(sp); t.append(time.perf_counter()); k1 = e.last_call_ms()[0]
    st = e.get_state(); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(f"resident loop {d[0]:.1f} ms (kernel {k0:.1f}) | get_state {d[1]:.1f} | set_state {d[2]:.1f} | set_replica_params {d[3]:.1f} | "
          f"loop after set_state {d[4]:.1f} ms (kernel {k1:.1f}) | get_state {d[5]:.1f}", flush=True)
