"""Plan of the cluster engine for the bench launch (CGPE_PLAN_DEBUG=1 prints it to stderr) next to the per-CTA timeline and
the list statistics (directed Debye-Hueckel pairs listed / charged and inside the cutoff)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine
w = workloads.config_c3_shard(0, 1)
e = Engine(w.cfg); e.set_engine("duo"); w.init_engine(e)
e.md_run(4000); e.sample_loop(w.sample_params(e, 5))
ls = e.list_stats()
q = (e.get_state()[0][..., 3] != 0).sum(axis=1)
os.environ["CGPE_PLAN_DEBUG"] = "1"
c0 = e.counters(); e.sample_loop(w.sample_params(e, 20)); ms, _ = e.last_call_ms(); c1 = e.counters()
os.environ.pop("CGPE_PLAN_DEBUG")
ns = e.engine_info()["cta_ns"].astype(np.float64)
dur = (ns[:, :, 1] - ns[:, :, 0]) * 1e-6
md = c1["md_steps"] - c0["md_steps"]
print(f"launch {ms:.2f} ms")
for r in np.argsort(-dur[:, 0])[:40]:
    print(f"replica {r:3d} pH {w.pH[r]:5.2f} steps {md[r]} cta ms {dur[r,0]:.2f} charged sites {q[r]} listed {ls[r,2]:.0f} charged+in-cutoff {ls[r,3]:.0f}")
