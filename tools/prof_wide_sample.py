"""Sampling protocol on the global-memory engine (128 pH replicas x N = 10^4, 1 mM): ms per sample, split into MD and trial moves."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine

n_mon = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 128
w = workloads.Workload(workloads.example_parameters(n_mon, c_salt=float(os.environ.get("C_SALT", 1e-3))), pH=np.linspace(2.0, 12.0, R))
e = Engine(w.cfg); w.init_engine(e)
e.md_run(1000); e.synchronize()
sp = w.sample_params(e, 3)
c0 = e.counters(); e.sample_loop(sp); e.synchronize(); ms, nl = e.last_call_ms(); c1 = e.counters()
steps = (c1["md_steps"] - c0["md_steps"]).sum()
print(f"N={n_mon} R={R}: sample loop {ms/3:.1f} ms/sample, {nl} launches, {n_mon*steps/(ms*1e-3):.3g} bead-steps/s; list stats/particle {np.round(e.list_stats().mean(0)/n_mon, 2)}")
e.md_run(500); e.synchronize(); ms_md, _ = e.last_call_ms()
print(f"md alone: {ms_md/500*1e3:.1f} us/step = {R*n_mon*500/(ms_md*1e-3):.3g} bead-steps/s")
for m, nt in ((0, 5 * w.n_sites[0] + 1),):
    c0 = e.counters(); e.mc_run(m, nt); e.synchronize(); ms_mc, _ = e.last_call_ms(); c1 = e.counters()
    acc = (c1["accepted"] - c0["accepted"]).sum() / (R * nt)
    print(f"mc block {m}: {nt} trials in {ms_mc:.1f} ms = {ms_mc/nt*1e3:.2f} us/trial (all replicas concurrently); acceptance {acc:.2f}")
