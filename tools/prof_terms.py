"""Marginal cost of each force-field term: md us/step with terms switched on one by one."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
n_mon = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 128
pH = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
base = dict(USE_WCA=False, USE_ELECTROSTATICS=False, USE_BENDING=False, USE_DIHEDRAL_POT=False)
steps = [("bond only", {}), ("+angle", dict(USE_BENDING=True)), ("+dihedral", dict(USE_DIHEDRAL_POT=True)),
         ("+WCA", dict(USE_WCA=True)), ("+DH", dict(USE_ELECTROSTATICS=True))]
# equilibrated start from the full force field
wf = workloads.Workload(workloads.example_parameters(n_mon, c_salt=2e-2), pH=np.full(R, pH))
ef = Engine(wf.cfg); wf.init_engine(ef); ef.md_run(3000); ef.mc_run(0, 30000); ef.mc_run(1, 500); ef.md_run(2000)
state = ef.get_state()
flags = dict(base)
for name, add in steps:
    flags.update(add)
    params = copy.deepcopy(wf.params); params["Simulation configuration"].update(flags)
    w = workloads.Workload(params, pH=np.full(R, pH))
    e = Engine(w.cfg); e.set_state(*state); e.set_replica_params(pH=w.pH, kappa=w.kappa, r_cut=w.r_cut); e.set_thermostat_seed(1434)
    e.md_run(200); e.synchronize()
    c0 = e.counters(); e.md_run(1000); e.synchronize(); ms = e.last_call_ms()[0]; c1 = e.counters()
    print(f"{name:12s} {ms:8.3f} us/step   rebuilds/1000 = {(c1['list_rebuilds']-c0['list_rebuilds']).mean():.1f}", flush=True)
