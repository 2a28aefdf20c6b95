#!/usr/bin/env python
"""Target of the ncu captures: the bench workload (C3 shard, 128 x N=1000, 1 mM) on one engine, two launches of the
fused sampling loop (the second one is the one to capture: -k regex:sample_loop -s 1 -c 1).
python tools/prof_ncu_target.py solo|duo [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coarse_grain_polyelectrolyte_b200 import workloads  # noqa: E402
from coarse_grain_polyelectrolyte_b200.engine import Engine  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "duo"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w = workloads.config_c3_shard(0, 1)
e = Engine(w.cfg)
e.set_engine(engine)
w.init_engine(e)
e.md_run(4000)
e.sample_loop(w.sample_params(e, 5))
e.sample_loop(w.sample_params(e, steps))
ms, n = e.last_call_ms()
print(f"{e.engine_info()['engine']}: {steps} samples in {ms:.2f} ms")
