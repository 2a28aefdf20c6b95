import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
c_salt = float(sys.argv[1]); pH = float(sys.argv[2]); R = 128
w = workloads.Workload(workloads.example_parameters(1000, c_salt=c_salt), pH=np.full(R, pH))
e = Engine(w.cfg); w.init_engine(e)
def rep(tag):
    rg = e.observables()[0]; en = e.energy()
    bad = np.nonzero(~np.isfinite(rg) | (rg > 100))[0]
    print(tag, "rg max %.3g" % np.nanmax(rg), "Emax %.4g" % np.nanmax(en[:, 0]), "bad replicas", bad[:10], flush=True)
e.md_run(3000); rep("md3000")
e.sample_loop_async(w.sample_params(e, 20)); e.synchronize(); rep("sample20")
e.list_stats(); rep("stats")
e.md_run(2000); e.synchronize(); rep("md2000")
nt = 5 * w.n_sites[0] + 1
e.mc_run(0, nt * 20); e.synchronize(); rep("mc")
for k in range(5):
    e.sample_loop_async(w.sample_params(e, 10)); e.synchronize(); rep("sample10 #%d" % k)
