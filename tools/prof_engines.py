#!/usr/bin/env python
"""One-CTA-per-replica engine vs cluster engine on several shapes of the same protocol.
python tools/prof_engines.py [samples]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coarse_grain_polyelectrolyte_b200 import workloads  # noqa: E402
from coarse_grain_polyelectrolyte_b200.engine import Engine  # noqa: E402

samples = int(sys.argv[1]) if len(sys.argv) > 1 else 20
shapes = [(100, 64, 2e-2, 4.0), (100, 128, 2e-2, 2.0), (100, 592, 2e-2, 2.0), (300, 128, 1e-3, 2.0), (1000, 128, 2e-2, 2.0),
          (1000, 128, 1e-3, 2.0), (1000, 64, 1e-3, 2.0), (1000, 256, 1e-3, 2.0), (2000, 64, 1e-3, 2.0), (2000, 128, 1e-3, 2.0)]
for n_mon, R, salt, ph_lo in shapes:
    out = []
    for engine in ("solo", "duo"):
        w = workloads.Workload(workloads.example_parameters(n_mon, c_salt=salt), pH=np.linspace(ph_lo, 12.0, R))
        e = Engine(w.cfg)
        e.set_engine(engine)
        w.init_engine(e)
        try:
            e.md_run(2000)
            e.sample_loop(w.sample_params(e, 3))
            c0 = e.counters()
            e.sample_loop(w.sample_params(e, samples))
            ms, _ = e.last_call_ms()
            md = (e.counters()["md_steps"] - c0["md_steps"]).sum()
            out.append((e.engine_info()["engine"], ms / samples, md * n_mon / ms * 1e3))
        except Exception as exc:   # e.g. the one-CTA engine forced onto N = 2000: 5 WCA row entries per particle, overflow reported
            out.append((engine + " (" + str(exc)[:60] + ")", float("nan"), float("nan")))
        e.close()
    print(f"N={n_mon:5d} R={R:4d} salt={salt:g}: " + "  ".join(f"{n}: {ms:7.3f} ms/sample {v:.3e} bead-steps/s" for n, ms, v in out)
          + f"  duo/solo {out[1][2] / out[0][2]:.2f}", flush=True)
