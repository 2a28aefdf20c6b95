"""Rg / E_coul / counts versus time for the notebook configuration on the CUDA engine (diagnostic for tests/test_notebook_anchor.py)."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "src"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ["CGPE_QUIET"] = "1"
import numpy as np
import test_notebook_anchor as T
from modules.EnsembleDynamics import EnsembleDynamics
ens = EnsembleDynamics(copy.deepcopy(T.NOTEBOOK_PARAMETERS), replicas=dict(pH=[4.0, 10.0, 4.0, 10.0]))
ens.equilibrate_pH()
ens.system.integrator.run(100000)
ens.num_samples = 43
ens.n_samp_iter = 16
for k in range(62):
    ens.perform_sampling()
    e = ens.system.analysis.energy()
    rg = ens.system.analysis.calc_rg(0, 1, 52)[0]; ree = ens.system.analysis.calc_re(0, 1, 52)[0]
    nb = ens.system.number_of_particles(type=ens.type_particles["B"])
    print(k, "t", round((k + 1) * 43 * 0.6 * 0.052, 1), "Rg", np.round(rg, 2), "Ree", np.round(ree, 1), "Ecoul", np.round(e["coulomb"], 2),
          "Ebond", np.round(e["bonded"], 1), "Ekin", np.round(e["kinetic"], 1), "nB", nb, "nN", ens.num_As[:, -1], flush=True)
