"""Phase timing of the chain engine (device time from the library's CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coarse_grain_polyelectrolyte_b200 import workloads
from coarse_grain_polyelectrolyte_b200.engine import Engine

def run(n_mon, R, c_salt, pH, label, md_steps=2000, skin=0.0):
    w = workloads.Workload(workloads.example_parameters(n_mon, c_salt=c_salt), pH=np.full(R, pH))
    if skin: w.cfg.skin = skin
    e = Engine(w.cfg); w.init_engine(e)
    e.md_run(3000); e.sample_loop_async(w.sample_params(e, 20)); e.synchronize()
    st = e.list_stats().mean(0)
    c0 = e.counters()
    e.md_run(md_steps); e.synchronize(); t_md = e.last_call_ms()[0]
    c1 = e.counters()
    nt = 5 * w.n_sites[0] + 1
    e.mc_run(0, nt * 20); e.synchronize(); t_mc = e.last_call_ms()[0]
    sp = w.sample_params(e, 50)
    e.sample_loop_async(sp); e.synchronize(); t_s = e.last_call_ms()[0]
    c2 = e.counters()
    obs, _ = e.fetch_samples()
    reb = (c1["list_rebuilds"] - c0["list_rebuilds"]).mean()
    print(f"{label}: N={n_mon} R={R} c_salt={c_salt} pH={pH} | md {t_md/md_steps*1e3:.2f} us/step ({reb:.1f} rebuilds/{md_steps}) | "
          f"mc {t_mc/(nt*20)*1e3:.3f} us/trial | sample {t_s/50:.3f} ms ({(c2['md_steps']-c1['md_steps']).mean()/50:.0f} md steps) | "
          f"lists W {st[0]/n_mon:.1f}/bead in {st[1]/n_mon:.2f}; D {st[2]/max(sum(w.n_sites),1):.1f}/site in {st[3]/max(sum(w.n_sites),1):.2f} | "
          f"alpha {obs[...,0].mean()/w.n_sites[0]:.2f} rg {obs[...,3].mean():.1f}", flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "a"): run(1000, 128, 2e-2, 8.0, "20mM")
    if which in ("all", "b"): run(1000, 128, 1e-3, 8.0, "1mM")
    if which in ("all", "c"): run(100, 64, 2e-2, 8.0, "C2")
    if which in ("all", "d"): run(1000, 128, 2e-2, 4.0, "20mM charged")
    if which in ("all", "e"): run(1000, 128, 2e-2, 12.0, "20mM neutral")
    if which in ("all", "f"): run(1000, 1, 2e-2, 8.0, "single replica")
    if which in ("g",): run(1000, 128, 1e-3, 3.0, "1mM charged")
    if which in ("h",): run(1000, 128, 2e-3, 3.0, "2mM charged")
    if which in ("i",): run(1000, 128, 1e-3, 12.5, "1mM neutral")
    if which in ("j",):
        for ph in (2.0, 6.0, 8.0, 9.0, 10.0, 12.0): run(1000, 128, 1e-3, ph, f"1mM pH {ph}")
