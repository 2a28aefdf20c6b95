"""Timing of the global-memory engine at large N."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
n_mon = int(sys.argv[1]); R = int(sys.argv[2]); pH = float(sys.argv[3]) if len(sys.argv) > 3 else 8.0
w = workloads.Workload(workloads.example_parameters(n_mon, c_salt=float(os.environ.get('C_SALT', 2e-2))), pH=np.full(R, pH), explicit_ions=bool(os.environ.get('IONS')))
if os.environ.get('SKIN'): w.cfg.skin = float(os.environ['SKIN'])
e = Engine(w.cfg); w.init_engine(e)
e.md_run(500); e.synchronize()
c0 = e.counters(); e.md_run(1000); e.synchronize(); ms, nl = e.last_call_ms(); c1 = e.counters()
print(f"N={n_mon} R={R}: md {ms:.1f} ms/1000 steps = {ms:.3f} us/step, {R*w.xyzq.shape[0]*1000/(ms*1e-3):.3g} bead-steps/s, launches {nl}, rebuilds {(c1['list_rebuilds']-c0['list_rebuilds']).mean():.1f}", flush=True)
nt = 5 * w.n_sites[0] + 1
e.mc_run(0, nt); e.synchronize(); ms, nl = e.last_call_ms()
print(f"  mc {ms/nt*1e3:.3f} us/trial ({nt} trials)", flush=True)
sp = w.sample_params(e, 3)
e.sample_loop_async(sp); e.synchronize(); ms, nl = e.last_call_ms(); obs, _ = e.fetch_samples()
print(f"  sample {ms/3:.2f} ms, launches {nl}; alpha {obs[...,0].mean()/w.n_sites[0]:.2f} rg {obs[...,3].mean():.1f} stats {e.list_stats().mean(0)/n_mon}", flush=True)
