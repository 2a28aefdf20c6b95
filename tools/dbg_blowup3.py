import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
w = workloads.Workload(workloads.example_parameters(1000, c_salt=2e-3), pH=np.full(1, 3.0), replica_offset=12)
e = Engine(w.cfg); w.init_engine(e)
def rep(tag):
    rg = e.observables()[0]; en = e.energy()[0]; f = e.forces()[0]; fm = e.md_forces()[0]
    print(tag, "rg %.4g" % rg[0], "E tot/kin/coul/bond/nb", np.round(en, 1), "min_dist %.3f" % e.min_dist()[0],
          "max|F| %.4g" % np.abs(f).max(), "max|F_md-F_all| %.3g" % np.abs(f - fm).max(), "reb", e.counters()["list_rebuilds"][0], flush=True)
e.md_run(3000); rep("md3000")
e.sample_loop_async(w.sample_params(e, 20)); e.synchronize(); rep("sample20")
e.list_stats()
e.md_run(2000); e.synchronize(); rep("md2000")
nt = 5 * w.n_sites[0] + 1
e.mc_run(0, nt * 20); e.synchronize(); rep("mc")
for k in range(30):
    e.sample_loop_async(w.sample_params(e, 1)); e.synchronize(); rep("s%02d" % k)
