#!/usr/bin/env python
"""Load-balance picture of one fused sampling launch: start/end of every CTA (GPU global timer) per replica,
for the one-CTA-per-replica engine and the cluster engine.  python tools/prof_timeline.py [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coarse_grain_polyelectrolyte_b200 import workloads  # noqa: E402
from coarse_grain_polyelectrolyte_b200.engine import Engine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for engine in ("solo", "duo"):
    w = workloads.config_c3_shard(0, 1)
    e = Engine(w.cfg)
    e.set_engine(engine)
    w.init_engine(e)
    e.md_run(4000)
    e.sample_loop(w.sample_params(e, 5))
    c0 = e.counters()
    e.sample_loop(w.sample_params(e, steps))
    ms, _ = e.last_call_ms()
    c1 = e.counters()
    info = e.engine_info()
    ns = info["cta_ns"].astype(np.float64)
    t0 = ns[ns > 0].min()
    dur = (ns[:, :, 1] - ns[:, :, 0]) * 1e-6
    end = (ns[:, :, 1] - t0) * 1e-6
    md = (c1["md_steps"] - c0["md_steps"])
    print(f"== {info['engine']}: launch {ms:.2f} ms for {steps} samples; {md.sum() * 1000 / ms * 1e3:.3e} bead-steps/s")
    print("   replica pH   md_steps  cta0 ms (end)   cta1 ms (end)")
    for r in range(0, w.R, int(os.environ.get("EVERY", 4))):
        print(f"   {r:4d} {w.pH[r]:5.2f} {md[r]:8d}   {dur[r,0]:7.2f} ({end[r,0]:7.2f})  {dur[r,1]:7.2f} ({end[r,1]:7.2f})")
    busy = dur[dur > 0].sum()
    print(f"   sum of CTA times {busy:.1f} ms; launch x 148 SMs = {ms * 148:.1f} ms")
