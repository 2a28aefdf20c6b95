"""BASELINE configs[4] at its lower end: a multi-chain solution with explicit counterions, 10^5 beads
in one box (one replica), on the global-memory engine.  Start: free random walks + random ions, overlaps
removed by capped steepest descent; then timed MD and trial moves."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine

n_poly = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_mon = int(sys.argv[2]) if len(sys.argv) > 2 else 500
R = int(sys.argv[3]) if len(sys.argv) > 3 else 1
c_salt = float(os.environ.get("C_SALT", 1e-2))
t0 = time.time()
w = workloads.Workload(workloads.example_parameters(n_mon, c_salt=c_salt, n_poly=n_poly), pH=np.linspace(7.0, 9.0, R),
                       explicit_ions=True, min_distance=0.0)
P = w.xyzq.shape[0]
print(f"{n_poly} chains x N={n_mon} = {n_poly*n_mon} beads + ions = {P} particles/replica, R={R}, box {w.ru.box_length:.1f}, "
      f"r_cut {w.r_cut[0]:.1f}; host setup {time.time()-t0:.1f} s", flush=True)
w.cfg.cap_wca = 200          # free random walks knot up before the descent has untied them
t0 = time.time()
e = Engine(w.cfg); w.init_engine(e); e.synchronize()
print(f"engine create + state upload: {time.time()-t0:.1f} s", flush=True)
t0 = time.time()
for cap in (0.02, 0.05, 0.1):
    e.steepest_descent(300, gamma=0.05 / w.ru.k_bond, max_displacement=cap)
e.synchronize()
t1 = time.time()
big = P > 600000   # the all-pairs diagnostics (min_dist, energy) are O(P^2): skipped for the largest boxes
md_ = float("nan") if big else e.min_dist()[0]
print(f"steepest descent 900 steps: {t1-t0:.2f} s; min_dist {md_:.2f} ({time.time()-t1:.1f} s)", flush=True)
t0 = time.time()
e.set_force_cap(200.0); e.md_run(500); e.set_force_cap(0.0); e.md_run(1000); e.synchronize()
print(f"warm-up 1500 steps: {time.time()-t0:.1f} s", flush=True)
c0 = e.counters(); e.md_run(1000); e.synchronize(); ms, nl = e.last_call_ms(); c1 = e.counters()
print(f"md: {ms/1000*1e3:.1f} us/step = {R*P*1000/(ms*1e-3):.3g} particle-steps/s; rebuilds/1000 steps "
      f"{(c1['list_rebuilds']-c0['list_rebuilds']).mean():.0f}; 64 B x that = {64*R*P*1000/(ms*1e-3)/1e9:.0f} GB/s", flush=True)
t0 = time.time(); ls = e.list_stats(); print(f"list stats/particle {np.round(ls.mean(0)/P, 2)} ({time.time()-t0:.1f} s)", flush=True)
nt = int(os.environ.get('TRIALS', 20000))
t0 = time.time()
e.mc_run(0, nt); e.synchronize(); ms, _ = e.last_call_ms()
print(f"mc wall {time.time()-t0:.1f} s", flush=True)
rg, ree, cnt = e.observables()
print(f"mc: {nt} trials in {ms:.1f} ms = {ms/nt*1e3:.2f} us/trial; counts {cnt[0]} (N_B == N_N + N_N2: {bool(cnt[0][2] == cnt[0][0] + cnt[0][1])}); Rg(chain 0) {rg[0]:.1f}")
if big:
    sys.exit(0)
t0 = time.time()
en = e.energy()[0]
print(f"energy: {time.time()-t0:.1f} s")
print("energy total/kin/coul/bonded/nonbonded per particle:", np.round(en / P, 3))
