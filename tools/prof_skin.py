import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
pH = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
base = workloads.Workload(workloads.example_parameters(1000, c_salt=2e-2), pH=np.full(128, pH))
e0 = Engine(base.cfg); base.init_engine(e0); e0.md_run(3000); e0.mc_run(0, 30000); e0.mc_run(1, 500); e0.md_run(2000)
state = e0.get_state()
for skin in (0.3, 0.5, 0.8, 1.1, 1.5):
    w = workloads.Workload(workloads.example_parameters(1000, c_salt=2e-2), pH=np.full(128, pH)); w.cfg.skin = skin
    e = Engine(w.cfg); e.set_state(*state); e.set_replica_params(pH=w.pH, kappa=w.kappa, r_cut=w.r_cut); e.set_thermostat_seed(1434)
    e.md_run(300); e.synchronize()
    c0 = e.counters(); e.md_run(3000); e.synchronize(); ms = e.last_call_ms()[0]; c1 = e.counters()
    st = e.list_stats().mean(0)
    print(f"skin {skin}: {ms/3:.3f} us/step, rebuilds/1000 {(c1['list_rebuilds']-c0['list_rebuilds']).mean()/3:.1f}, W/bead {st[0]/1000:.1f} D/site {st[2]/334:.1f}", flush=True)
