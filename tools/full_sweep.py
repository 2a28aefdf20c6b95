"""Full pH sweep (all num_samples of the protocol, E.py:67-119) through the drop-in driver; wall times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "src"), ROOT]
os.environ.setdefault("CGPE_QUIET", "1")
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from modules.EnsembleDynamics import EnsembleDynamics
from modules.utils import ideal_alpha

def sem(x, nb=16):
    n = (x.shape[-1] // nb) * nb
    m = x[..., :n].reshape(x.shape[:-1] + (nb, -1)).mean(-1)
    return m.std(-1, ddof=1) / np.sqrt(nb)

def run(label, n_mon, pH, c_salt):
    params = workloads.example_parameters(n_mon, c_salt=c_salt)
    params["Simulation configuration"]["ION_MODEL"] = "implicit"
    t0 = time.time()
    ens = EnsembleDynamics(params, replicas=dict(pH=pH, c_salt=[c_salt] * len(pH)))
    t1 = time.time()
    ens.equilibrate_pH()
    t2 = time.time()
    ens.perform_sampling()
    t3 = time.time()
    eng = ens.system._engine
    c = eng.counters()
    a1, a2 = ens.titration_curve()
    print(f"## {label}: {len(pH)} pH replicas x N={n_mon}, c_salt={c_salt} M, {ens.num_samples} samples per replica")
    print(f"setup+warm-up {t1-t0:.2f} s | equilibrate_pH {t2-t1:.2f} s | perform_sampling {t3-t2:.2f} s  <- full pH-sweep wall time")
    print(f"MD steps/replica {c['md_steps'].mean():.0f}, trials/replica {c['trials'].sum(1).mean():.0f}, list rebuilds/replica {c['list_rebuilds'].mean():.0f}; "
          f"{c['md_steps'].sum()*n_mon/(t3-t2):.3g} bead-steps/s, {c['trials'].sum()/ (ens.n_nh+ens.n_nh2)/(t3-t2):.3g} MC sweeps/s over the wall time")
    print("   pH  alpha_NH  +-sem   ideal | alpha_NH2  ideal |     Rg  +-sem     Ree")
    for r in range(0, len(pH), max(1, len(pH) // 16)):
        print(f"{pH[r]:6.2f} {a1[r]:8.4f} {sem(ens.num_As[r])/ens.n_nh:7.4f} {ideal_alpha(pH[r], ens.pK):7.4f} | {a2[r]:8.4f} {ideal_alpha(pH[r], ens.pK2):7.4f} |"
          f" {ens.rad_gyr[r].mean():7.3f} {sem(ens.rad_gyr[r]):6.3f} {ens.end_2end[r].mean():8.3f}")
    print(flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
if which == "c2": run("C2", 100, np.linspace(4.0, 12.0, 64), 2e-2)
if which == "c3": run("C3 shard (one ionic strength)", 1000, np.linspace(2.0, 12.0, 128), float(sys.argv[2]) if len(sys.argv) > 2 else 2e-2)
