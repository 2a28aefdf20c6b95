"""Where a sample of the fused loop goes: full loop vs trial moves only vs MD only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
c_salt = float(os.environ.get("C_SALT", 1e-3)); n = 100
for pH in (2.0, 8.0, 12.0):
    w = workloads.Workload(workloads.example_parameters(1000, c_salt=c_salt), pH=np.full(128, pH))
    e = Engine(w.cfg); w.init_engine(e)
    e.md_run(4000); e.sample_loop(w.sample_params(e, 20))
    out = []
    for label, kw in (("full", {}), ("mc+obs", dict(use_md=False)), ("md+obs", dict(mc_blocks=()))):
        k = dict(w.sample_kwargs); k.update(kw)
        sp = e.sample_params(n, q_snapshot_rows=0, **{**k, "q_snapshot_every": 0})
        e.sample_loop_async(sp); e.synchronize(); ms = e.last_call_ms()[0]
        out.append(f"{label} {ms/n:.3f} ms")
    print(f"pH {pH}: " + " | ".join(out), flush=True)
