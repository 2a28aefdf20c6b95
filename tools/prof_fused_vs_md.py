"""MD step of the fused sampling kernel against the stand-alone MD kernel (same replicas, same state)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
for pH in (2.0, 12.0):
    w = workloads.Workload(workloads.example_parameters(1000, c_salt=1e-3), pH=np.full(128, pH))
    e = Engine(w.cfg); e.set_engine(os.environ.get('ENGINE', 'solo')); w.init_engine(e)
    e.md_run(4000); e.sample_loop(w.sample_params(e, 20))
    e.md_run(500); e.synchronize()
    e.md_run(5000); e.synchronize(); t_md = e.last_call_ms()[0] / 5000 * 1e3
    k = dict(w.sample_kwargs); k.update(mc_blocks=(), p_burst1=1.0, p_burst2=0.0, q_snapshot_every=0)
    sp = e.sample_params(10, q_snapshot_rows=0, **k)
    e.sample_loop_async(sp); e.synchronize(); t_f = e.last_call_ms()[0] / 5000 * 1e3
    k.update(mc_blocks=w.sample_kwargs["mc_blocks"])
    sp = e.sample_params(10, q_snapshot_rows=0, **k)
    e.sample_loop_async(sp); e.synchronize(); t_fm = e.last_call_ms()[0] / 10
    print(f"pH {pH}: stand-alone MD kernel {t_md:.2f} us/step | fused loop, MD only {t_f:.2f} us/step | fused with trial moves {t_fm:.3f} ms/sample (500 steps each)", flush=True)
