"""Trial moves alone: time per trial and acceptance, by pH and salt (positions frozen, caches rebuilt per call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
for c_salt in (1e-3, 2e-2):
    for pH in (2.0, 5.0, 8.0, 10.0, 12.0):
        w = workloads.Workload(workloads.example_parameters(1000, c_salt=c_salt), pH=np.full(128, pH))
        e = Engine(w.cfg); w.init_engine(e)
        e.md_run(3000); e.sample_loop(w.sample_params(e, 20))
        nt = 1661 * 20
        e.mc_run(0, nt); e.synchronize()
        c0 = e.counters(); e.mc_run(0, nt); e.synchronize(); ms = e.last_call_ms()[0]; c1 = e.counters()
        acc = (c1["accepted"] - c0["accepted"]).sum() / (128.0 * nt)
        q = (e.get_state()[0][..., 3] != 0).sum(axis=1).mean()
        us = ms * 1e3 / nt
        print(f"c_salt {c_salt:g} pH {pH:4.1f}: {us:.3f} us/trial, acceptance {acc:.3f}, charged sites {q:.0f} -> "
              f"{(us - 0.165) / max(acc, 1e-9):.2f} us per accepted flip beyond a rejected trial's 0.165", flush=True)
