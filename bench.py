#!/usr/bin/env python
"""bench.py — headline measurement of the hot path.

A "step" is one pass of the reference's sampling-loop body (EnsembleDynamics.perform_sampling,
E.py:88-106) over all replicas of this rank: the Bernoulli-scheduled bursts of 500 Langevin MD
steps, the two blocks of constant-pH trial moves (5*n_nh+1, 5*n_nh2+1) and the observables.
Workload at N GPUs: rank r holds 128 pH replicas of an N=1000 chain at one ionic strength
(BASELINE.json configs[2] sharded, one ionic strength per GPU: weak scaling).

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # CPU restatement on the host cores

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from coarse_grain_polyelectrolyte_b200 import workloads  # noqa: E402

METRIC = "md_bead_steps_per_s"
UNIT = "bead-steps/s"
SM_COUNT_B200 = 148
FP32_LANES_PER_SM = 128
MUFU_PER_SM = 16
# algorithmic work per unit
FLOP_PAIR_TEST, FLOP_PAIR_IN, FLOP_BEAD, FLOP_MC_TERM = 8.0, 24.0, 250.0, 14.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--n-mon", type=int, default=1000)
    ap.add_argument("--replicas-per-gpu", type=int, default=128)
    ap.add_argument("--equil-steps", type=int, default=4000, help="untimed MD steps before warm-up")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--skin", type=float, default=0.0)
    ap.add_argument("--scaling", choices=("weak", "strong"), default="weak",
                    help="weak: 128 pH replicas of one ionic strength per GPU (default); strong: the whole 128 x 8 grid "
                         "(1024 replicas) divided over the GPUs")
    ap.add_argument("--engine", choices=("auto", "solo", "duo"), default="auto")
    ap.add_argument("--no-secondary", action="store_true", help="skip the N=100 / N=10^4 side measurements (N=1 only)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _ncu_summary(engine):
    name = {"duo": "r2_sample_loop_duo_ncu_summary.json", "solo": "r2_sample_loop_solo_ncu_summary.json"}.get(engine)
    return os.path.join(ROOT, "profiles", name) if name else None


def ncu_traffic(engine="solo"):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of the engine in force
    (profiles/r2_sample_loop_{solo,duo}_ncu_summary.json; the capture command is recorded there; a 10-sample launch of the
    same 128 x N=1000 shard: the traffic is the state in and out, it does not grow with the samples).  None if absent."""
    path = _ncu_summary(engine)
    try:
        with open(path) as f:
            d = json.load(f)
        return {"dram_bytes_per_launch": d["derived"]["dram_bytes_per_launch"],
                "source": f"ncu --set full capture profiles/{os.path.basename(path)} (10-sample launch of the same shard)"}
    except (OSError, ValueError, KeyError, TypeError):
        return None


def ncu_issue(engine="solo"):
    """Instruction-issue figures of the dominant kernel from the committed ncu --set full capture of the engine in force:
    the roof that actually binds this kernel.  Measured under the profiler on the same workload, reported beside the live
    flop count, never instead of it."""
    path = _ncu_summary(engine)
    try:
        with open(path) as f:
            m = json.load(f)["metrics"]
        ipc = float(m["sm__inst_executed.avg.per_cycle_elapsed"][0])
        return {"warp_inst_per_cycle_per_sm": ipc, "peak": 4.0, "frac": ipc / 4.0,
                "issue_slots_busy_of_active_cycles": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]) / 100.0,
                "sm_active_of_elapsed_cycles": float(m["sm__cycles_active.avg"][0]) / float(m["sm__cycles_elapsed.avg"][0]),
                "lanes_per_warp_instruction": float(m["smsp__thread_inst_executed_per_inst_executed.ratio"][0]),
                "source": f"ncu --set full capture profiles/{os.path.basename(path)} (10-sample launch, under the profiler)"}
    except (OSError, ValueError, KeyError):
        return None


def algorithmic_flops(n_mon, md_steps, trials, stats):
    """SURVEY §8d convention on the work the kernel's algorithm performs (unique listed pairs)."""
    w_all, w_in, d_all, d_in = (float(x) / 2.0 for x in stats)   # directed -> unique
    per_step = (w_all + d_all) * FLOP_PAIR_TEST + (w_in + d_in) * FLOP_PAIR_IN + n_mon * FLOP_BEAD
    # a trial visits the site's two list rows (directed entries / particles that own a row)
    return md_steps * per_step, per_step


def cpu_replica_subset(w, n):
    """n replicas spread evenly over the shard's replica axis (every pH region, every salt of the shard), not the
    first n: the low-pH replicas are the most expensive ones."""
    n = min(n, w.R)
    return np.unique(np.linspace(0, w.R - 1, n).round().astype(int))


def run_cpu_sample(world_rank_workload, threads, n_replicas=None, list_mode=True):
    """The oracle on a stratified subset of the workload's replicas, one per host thread."""
    from oracle import oracle
    oracle.build()
    w = world_rank_workload
    idx = cpu_replica_subset(w, n_replicas or threads)
    sub = workloads.Workload(w.params, pH=w.pH[idx], c_salt=None, replica_offset=int(w.cfg.replica_offset))
    sub.kappa, sub.r_cut = w.kappa[idx], w.r_cut[idx]
    oracle.set_threads(threads)
    o = oracle.OracleEngine(sub.cfg)
    sub.init_engine(o)
    if list_mode:
        o.set_list_mode(True, 0.8)
    return o, sub, idx


def _cpu_timed(o, sub, budget_s, warm_samples):
    o.md_run(200)                                    # leave the random-walk start
    o.sample_loop(sub.sample_params(o, warm_samples))
    t0 = time.perf_counter(); o.sample_loop(sub.sample_params(o, 3)); dt3 = time.perf_counter() - t0
    n = int(max(3, min(2000, budget_s / max(dt3 / 3.0, 1e-6))))
    c0 = o.counters()
    t0 = time.perf_counter(); o.sample_loop(sub.sample_params(o, n)); dt = time.perf_counter() - t0
    c1 = o.counters()
    steps = int((c1["md_steps"] - c0["md_steps"]).sum())
    trials = int((c1["trials"] - c0["trials"]).sum())
    return steps, trials, dt, n


def cpu_measure(w, threads, budget_s, warm_samples=2, list_mode=True):
    """cpu_baseline: (i) all host threads, one replica each, replicas stratified over the shard; (ii) ONE core, replicas
    one after the other -- how the reference actually runs (one process, pH points in a serial loop, NB:276)."""
    o, sub, idx = run_cpu_sample(w, threads, list_mode=list_mode)
    steps, trials, dt, n = _cpu_timed(o, sub, 0.67 * budget_s, warm_samples)
    n_acid = sum(sub.n_sites)
    out = {
        "value": steps * sub.cfg.n_beads / dt, "unit": UNIT, "cores": threads, "kind": "port", "same_config": True,
        "sample": f"{len(idx)} of the {w.R} replicas, evenly spaced over the shard (pH {sub.pH[0]:.2f}..{sub.pH[-1]:.2f}), one per host "
                  f"thread x {n} samples of the same protocol, N={sub.cfg.n_beads}, FP64 C oracle with Verlet lists (skin 0.8); {dt:.1f} s",
        "mc_sweeps_per_s": trials / n_acid / dt, "samples_per_s": len(idx) * n / dt, "seconds": dt, "n_samples": n,
    }
    out["value_per_core"] = out["value"] / threads
    try:
        o1, sub1, idx1 = run_cpu_sample(w, 1, n_replicas=4, list_mode=list_mode)
        s1, t1, dt1, n1 = _cpu_timed(o1, sub1, 0.33 * budget_s, 1)
        out["single_core"] = {"value": s1 * sub1.cfg.n_beads / dt1, "unit": UNIT, "cores": 1,
                              "sample": f"{len(idx1)} replicas evenly spaced over the shard, one after the other on one core x {n1} samples; {dt1:.1f} s"}
    except Exception as exc:   # noqa: BLE001
        out["single_core"] = {"value": None, "sample": f"failed: {exc!r}"}
    return out, n


def main_reference(args, rank, world):
    """CPU arm: the restatement of the reference's path on the host cores (ESPResSo itself is not
    installable here).  Same workload as the CUDA arm; each step = one sample of the protocol over a
    bounded, stratified subset of its replicas (one per host thread)."""
    if rank != 0:
        return
    w = make_workload(args, 0, max(world, 1))
    threads = os.cpu_count() or 1
    o, sub, idx = run_cpu_sample(w, threads)
    o.md_run(200)
    for _ in range(args.warmup):
        o.sample_loop(sub.sample_params(o, 1))
    c0 = o.counters()
    t0 = time.perf_counter()
    o.sample_loop(sub.sample_params(o, args.steps))
    dt = time.perf_counter() - t0
    c1 = o.counters()
    steps = int((c1["md_steps"] - c0["md_steps"]).sum())
    trials = int((c1["trials"] - c0["trials"]).sum())
    value = steps * sub.cfg.n_beads / dt
    sample = (f"{len(idx)} of the {w.R} replicas of rank 0's shard, evenly spaced over it (pH {sub.pH[0]:.2f}..{sub.pH[-1]:.2f}), "
              f"one per host thread x {args.steps} samples, FP64 C oracle with Verlet lists")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, w, world) + f"; CPU arm: {len(idx)} replicas evenly spaced over the shard",
                   "replicas": len(idx), "n_mon": args.n_mon},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "same_config": True,
                         "value_per_core": value / threads, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mc_sweeps_per_s": trials / sum(sub.n_sites) / dt,
        "gpu_launches": 0,
    }
    emit(line)


def make_workload(args, rank, world, device=-1):
    if args.scaling == "strong":     # the whole pH x ionic-strength grid, divided over the ranks
        return workloads.config_c3_shard(rank, world, n_ph=args.replicas_per_gpu, n_mon=args.n_mon, device=device, all_salts=True)
    return workloads.config_c3_shard(rank, world, n_ph=args.replicas_per_gpu, n_mon=args.n_mon, device=device)


def workload_name(args, w, world):
    mc = w.sample_kwargs["mc_blocks"]
    body = ("step = perform_sampling loop body (Bernoulli bursts of 500 Langevin MD steps + "
            f"{mc[0][1]}+{mc[1][1]} constant-pH trials + Rg/Ree/counts)")
    if args.scaling == "strong":
        return (f"BASELINE configs[2], whole grid: {args.replicas_per_gpu} pH x 8 ionic strengths = {args.replicas_per_gpu * 8} replicas of an "
                f"N={args.n_mon} chain divided over {world} GPU(s) ({w.R} on each); " + body)
    return (f"BASELINE configs[2] shard: {w.R} pH replicas per GPU x N={args.n_mon} chain, one ionic strength per GPU; " + body)


_REAL_STDOUT = None


def claim_stdout():
    """Everything any library writes to fd 1 (NCCL prints its version banner there at every debug level)
    goes to stderr from here on; emit() alone writes to the real stdout: ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def measure_gpu(eng, w, steps, warmup, equil_steps, torch, stream, barrier=None):
    """W warm-up steps, then K steps of the protocol in ONE fused launch, CUDA events on the launching stream."""
    eng.md_run(equil_steps)
    eng.sample_loop_async(w.sample_params(eng, max(warmup, 3)))
    eng.synchronize()
    stats = eng.list_stats().mean(axis=0)
    c0 = eng.counters()
    sp = w.sample_params(eng, steps)
    if barrier:
        barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    eng.sample_loop_async(sp)
    ev1.record(stream)
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    c1 = eng.counters()
    return ms, stats, c0, c1, sp


def secondary_workload(name, w, torch, stream, steps, warmup, equil_steps, peaks, sm_count):
    """One side measurement (same protocol, other chain length) with its own roofline entry."""
    from coarse_grain_polyelectrolyte_b200.engine import Engine
    eng = Engine(w.cfg)
    eng.set_stream(stream.cuda_stream)
    w.init_engine(eng)
    ms, stats, c0, c1, _ = measure_gpu(eng, w, steps, warmup, equil_steps, torch, stream)
    n_mon = w.cfg.n_beads
    md = int((c1["md_steps"] - c0["md_steps"]).sum())
    trials = int((c1["trials"] - c0["trials"]).sum())
    info = eng.engine_info()
    value = md * n_mon / (ms * 1e-3)
    out = {"workload": name, "replicas": w.R, "n_mon": n_mon, "engine": info["engine"], "steps": steps,
           "value": value, "unit": UNIT, "ms_per_step": ms / steps, "mc_sweeps_per_s": trials / sum(w.n_sites) / (ms * 1e-3),
           "gpu_launches": eng.last_call_ms()[1]}
    ns = info["cta_ns"].astype(np.float64)
    busy = (ns[:, :, 1] - ns[:, :, 0]) * 1e-6
    if info["engine"] != "wide" and (busy > 0).any():
        out["sm_time_used_frac"] = float(busy[busy > 0].sum()) / (ms * sm_count) / (2.0 if info["engine"] == "duo" else 1.0)
    if info["engine"] == "wide":
        # global-memory engine: every MD step streams positions and velocities (float4 each, read + write) = 64 B per bead-step
        gbs = 64.0 * value / 1e9
        out["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                           "traffic": None, "note": "algorithmic 64 B per bead-step; the working set of this workload is L2-resident, "
                                                     "the step is bound by dependent L2 gathers of partner positions, not by HBM"}
    else:
        fp32_peak = sm_count * FP32_LANES_PER_SM * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
        flops, _ = algorithmic_flops(n_mon, md, trials, stats)
        out["roofline"] = {"bound": "fp32", "achieved": flops / (ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                           "frac": flops / (ms * 1e-3) / 1e12 / fp32_peak, "traffic": None}
    eng.close()
    return out


def secondary_c5(torch, stream, peaks, device, n_poly=200, n_mon=500, md_steps=1000, trials=2000):
    """BASELINE configs[4]: one box of n_poly chains with explicit counter-ions and salt (one replica of 4 x 10^5 particles),
    global-memory engine.  Start as tools/run_c5.py: random walks + random ions, overlaps removed by capped steepest descent;
    then timed MD (particle-steps/s against the HBM roof: 64 B per particle-step) and a block of trial moves."""
    from coarse_grain_polyelectrolyte_b200.engine import Engine
    w = workloads.Workload(workloads.example_parameters(n_mon, c_salt=1e-2, n_poly=n_poly), pH=np.array([8.0]), explicit_ions=True,
                           min_distance=0.0, device=device)
    P = int(w.xyzq.shape[0])
    w.cfg.cap_wca = 200
    eng = Engine(w.cfg)
    eng.set_stream(stream.cuda_stream)
    w.init_engine(eng)
    for cap in (0.02, 0.05, 0.1):
        eng.steepest_descent(300, gamma=0.05 / w.ru.k_bond, max_displacement=cap)
    eng.set_force_cap(200.0); eng.md_run(500); eng.set_force_cap(0.0); eng.md_run(1000)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream); eng.md_run(md_steps); ev1.record(stream); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = eng.last_call_ms()[1]
    value = P * md_steps / (ms * 1e-3)
    ev0.record(stream); eng.mc_run(0, trials); ev1.record(stream); torch.cuda.synchronize()
    ms_mc = ev0.elapsed_time(ev1)
    _, _, cnt = eng.observables()
    gbs = 64.0 * value / 1e9
    out = {"workload": f"BASELINE configs[4]: one box, {n_poly} chains x N={n_mon} + explicit ions = {P} particles, 10 mM", "particles": P,
           "engine": eng.engine_info()["engine"], "md_steps": md_steps, "value": value, "unit": "particle-steps/s",
           "us_per_md_step": ms / md_steps * 1e3, "us_per_trial_move": ms_mc / trials * 1e3, "gpu_launches": launches,
           "n_b_equals_n_n_plus_n_n2": bool(cnt[0][2] == cnt[0][0] + cnt[0][1]),
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                        "note": "algorithmic 64 B per particle-step (positions and velocities, read + write); the step is five short kernels "
                                "with L2 gathers of partner positions, list rebuilds every ~70 steps included"}}
    eng.close()
    return out


def main():
    args = parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return main_reference(args, rank, world)

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"      # NCCL's version banner goes to stdout; this script prints ONE line there
    import torch
    import torch.distributed as dist
    from coarse_grain_polyelectrolyte_b200 import sharding
    from coarse_grain_polyelectrolyte_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w = make_workload(args, rank, world, device=local_rank)
    if args.skin > 0:
        w.cfg.skin = args.skin
    eng = Engine(w.cfg)
    if args.engine != "auto":
        eng.set_engine(args.engine)
    # a real (non-legacy) stream: torch's default stream has handle 0, which the C-ABI reads as
    # "use the handle's own stream", and torch events would then bracket nothing
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    w.init_engine(eng)
    info = eng.device_info()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # L2 flush: the replicas live in shared memory; flush anyway so nothing is served from a warm L2
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    flush.fill_(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # untimed: leave the synthetic random-walk start, W warm-up steps; then ONE launch of the fused loop: K steps for all replicas
    ms, stats, c0, c1, sp = measure_gpu(eng, w, args.steps, args.warmup, args.equil_steps, torch, stream, barrier)
    clocks = sampler.stop()
    kernel_ms, launches = eng.last_call_ms()
    engine_info = eng.engine_info()
    obs, _ = eng.fetch_samples()
    md_steps = int((c1["md_steps"] - c0["md_steps"]).sum())
    trials = int((c1["trials"] - c0["trials"]).sum())
    rebuilds = int((c1["list_rebuilds"] - c0["list_rebuilds"]).sum())
    n_acid = sum(w.n_sites)
    # load balance of the launch: busy time of every CTA against the launch
    ns = engine_info["cta_ns"].astype(np.float64)
    cta_ms = (ns[:, :, 1] - ns[:, :, 0]) * 1e-6
    cta_ms = cta_ms[cta_ms > 0]

    # end to end through the public call with HOST buffers: upload the state, run K steps, read the observables and
    # the final state back -- and, with several GPUs, gather every rank's observable block over NCCL (the one
    # collective of the path); copies and the collective inside the timed region
    xyzq, vel, typ = eng.get_state()
    pin = [torch.from_numpy(a).pin_memory().numpy() for a in (xyzq, vel, typ)]
    n_total = w.R * world
    barrier()
    t0 = time.perf_counter()
    eng.set_state(*pin)
    t_a = time.perf_counter()
    eng.set_replica_params(pH=w.pH, kappa=w.kappa, r_cut=w.r_cut)
    t_b = time.perf_counter()
    obs2, _ = eng.sample_loop(sp)
    t_c = time.perf_counter()
    out_state = eng.get_state()
    t_d = time.perf_counter()
    obs_all = sharding.gather_observables(obs2, n_total, device=torch.device("cuda", local_rank)) if world > 1 else obs2
    t_e = time.perf_counter()
    barrier()
    e2e_s = time.perf_counter() - t0
    assert obs_all.shape[0] == n_total
    e2e_parts = {"set_state": t_a - t0, "set_replica_params": t_b - t_a, "sample_loop": t_c - t_b,
                 "get_state": t_d - t_c, "nccl_gather": t_e - t_d, "barrier": e2e_s - (t_e - t0), "kernel_ms": eng.last_call_ms()[0]}
    ce = eng.counters()
    md_steps_e2e = int((ce["md_steps"] - c1["md_steps"]).sum())
    h2d = sum(a.nbytes for a in pin) + 3 * 8 * w.R
    d2h = obs2.nbytes + sum(a.nbytes for a in out_state)

    # the trial moves alone (one launch of cgpe_mc_run per reaction, the protocol's block sizes): MUFU roofline of SURVEY 8d
    mc_ms, mc_trials = 0.0, 0
    for rx, n_tr in w.sample_kwargs["mc_blocks"]:
        eng.mc_run(rx, n_tr)
        eng.synchronize()
        mc_ms += eng.last_call_ms()[0]
        mc_trials += n_tr * w.R

    # max over ranks of the device time; sums over ranks of the work
    t_ms, t_e2e = ms, e2e_s
    tot = torch.tensor([md_steps, trials, md_steps_e2e, rebuilds], dtype=torch.float64, device="cuda")
    if world > 1:
        tt = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        t_ms, t_e2e = float(tt[0]), float(tt[1])
        # nothing below needs the other ranks: let them go instead of spinning in a barrier while rank 0 times the CPU arm
        dist.barrier()
        dist.destroy_process_group()
    md_all, tr_all, md_e2e_all, reb_all = (float(x) for x in tot.tolist())
    n_mon = args.n_mon
    value = md_all * n_mon / (t_ms * 1e-3)
    e2e_value = md_e2e_all * n_mon / t_e2e
    if rank != 0:
        return

    wide = engine_info["engine"] == "wide"          # replicas beyond the shared-memory engines run on the global-memory one
    peaks, peak_kind = measured_peaks()
    sm_count = info["sm_count"] or SM_COUNT_B200
    fp32_peak = sm_count * FP32_LANES_PER_SM * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12      # TFLOP/s
    mufu_peak = sm_count * MUFU_PER_SM * peaks["sm_max_mhz"] * 1e6                       # op/s (measured: tools/ubench_issue.cu, 15.9 op/clk/SM)
    md_flops, per_step = algorithmic_flops(n_mon, md_steps, trials, stats)
    terms_per_trial = (stats[0] / n_mon + stats[2] / max(n_acid, 1))                      # rows visited by a trial
    mc_flops = trials * terms_per_trial * FLOP_MC_TERM
    achieved = (md_flops + mc_flops) / (ms * 1e-3) / 1e12
    # the same step count charged as the SURVEY's all-pairs algorithm would be (context only)
    allpairs_equiv = md_steps * (n_mon * (n_mon - 1) / 2 * FLOP_PAIR_TEST + (stats[1] + stats[3]) / 2 * FLOP_PAIR_IN
                                 + n_mon * FLOP_BEAD) / (ms * 1e-3) / 1e12
    # SURVEY 8d, MC row: a Mode-I trial is (N-1) partner terms of 2 MUFU each; here a trial reads a cached potential and
    # only an accepted flip visits its Debye-Hueckel row, so both figures are given
    mc_terms_alg = mc_trials * (n_mon - 1)
    mc_terms_done = mc_trials * float(stats[2]) / max(n_acid, 1)
    engine_names = {"solo": "shared-memory (one persistent CTA per replica)",
                    "duo": "shared-memory cluster engine (one thread-block cluster of two CTAs per replica, positions mirrored through DSMEM)",
                    "wide": "global-memory (cell-grid lists, CUDA-graph steps)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_ms / max(args.steps, 1), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload_name(args, w, world),
            "replicas_per_gpu": w.R, "n_mon": n_mon, "n_titratable": n_acid, "ion_model": "implicit (Debye-Hueckel)",
            "engine": engine_names[engine_info["engine"]],
            "l2": ("replica state and Verlet rows live in global memory and are streamed by every kernel of a step; "
                   "256 MiB L2 flush before the timed region" if wide else
                   "replica state is shared-memory resident inside one launch; 256 MiB L2 flush before the timed region"),
            "equil_md_steps": args.equil_steps, "skin": float(w.cfg.skin) or 0.8,
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / max(args.steps, 1),
                "d2h_bytes_per_step": d2h / max(args.steps, 1), "seconds": t_e2e,
                "rank0_parts_s": {k: round(v, 4) for k, v in e2e_parts.items()},
                "call": "Engine.set_state + set_replica_params + sample_loop + get_state on host (pinned) buffers"
                        + (" + sharding.gather_observables (NCCL all_gather of every rank's observable block)" if world > 1 else "")},
        "gpu_launches": launches,
        "clocks": clocks,
        "mc_sweeps_per_s": tr_all / n_acid / (t_ms * 1e-3),
        "samples_per_s": w.R * world * args.steps / (t_ms * 1e-3),
        "md_steps_per_replica_sample": md_all / (w.R * world * max(args.steps, 1)),
        "list_rebuilds_per_1k_md_steps": 1e3 * reb_all / max(md_all, 1.0),
        "ph_sweep_wall_s_projected": w.num_samples * t_ms * 1e-3 / max(args.steps, 1),
        "kernel_ms_cgpe_events": kernel_ms,
        "load_balance": {"cta_busy_ms_sum": float(cta_ms.sum()), "launch_ms_x_sms": ms * sm_count,
                         "sm_time_used_frac": float(cta_ms.sum()) / (ms * sm_count) / (2.0 if engine_info["engine"] == "duo" else 1.0),
                         "slowest_cta_ms": float(cta_ms.max()) if cta_ms.size else None, "mean_cta_ms": float(cta_ms.mean()) if cta_ms.size else None,
                         "note": "busy time of every CTA of the timed launch (GPU global timer) against launch x SMs; two CTAs share an SM in the cluster engine"},
        "roofline": {
            "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
            "traffic": None if wide else (ncu_traffic(engine_info["engine"]) or {}).get("dram_bytes_per_launch"),
            "traffic_source": "no ncu --set full capture of the global-memory engine's kernels committed" if wide
                              else (ncu_traffic(engine_info["engine"]) or {}).get("source", "no ncu capture committed"), "peak_source": f"148 SM x 128 lanes x 2 x sm_max_mhz ({peak_kind} clock; no FP32 entry in MEASURED_PEAKS.json)",
            "kernel": "wk_forces + wk_dh_rows (per MD step) / wk_mc_block (trials)" if wide else ("k_duo_sample_loop" if engine_info["engine"] == "duo" else "k_sample_loop"),
            "algorithmic_flop_per_md_step": per_step, "unique_listed_pairs": [float(s) / 2 for s in stats],
            "note": ("Verlet-list algorithm: work counted on the pairs the lists hold (SURVEY 8d formula) over the whole step "
                     "(all kernels); the force kernel is bound by L2 gathers of partner positions, the trial loop by its "
                     "chain of dependent reads" if wide else
                     "Verlet-list algorithm: work counted on the pairs the lists hold (SURVEY 8d formula); the kernel is "
                     "bound by instruction issue (25 SASS instructions per listed Debye-Hueckel pair against 32 counted flop; "
                     "the packed FFMA2 form takes two issue slots, tools/ubench_issue.cu) and by the launch lasting as long as "
                     "its slowest replica, not by a pipe"),
            "allpairs_equivalent_tflops": allpairs_equiv,
            "instruction_issue": None if wide else ncu_issue(engine_info["engine"]),
        },
        "roofline_mc": {
            "bound": "mufu", "unit": "partner-terms/s x 2 MUFU", "peak": mufu_peak,
            "achieved_algorithmic": 2.0 * mc_terms_alg / (mc_ms * 1e-3) if mc_ms > 0 else None,
            "frac_algorithmic": 2.0 * mc_terms_alg / (mc_ms * 1e-3) / mufu_peak if mc_ms > 0 else None,
            "achieved_listed": 2.0 * mc_terms_done / (mc_ms * 1e-3) if mc_ms > 0 else None,
            "frac_listed": 2.0 * mc_terms_done / (mc_ms * 1e-3) / mufu_peak if mc_ms > 0 else None,
            "trials": mc_trials, "ms": mc_ms, "us_per_trial_per_replica": mc_ms * 1e3 / max(mc_trials / w.R, 1),
            "note": "trial moves timed alone (cgpe_mc_run, the protocol's two blocks, list build + per-site caches included). "
                    "algorithmic: SURVEY 8d's Mode-I cost of (N-1) partner terms per trial; listed: the partner terms of the site's "
                    "Debye-Hueckel row.  A trial reads a cached per-site potential, only accepted flips walk a row: the loop is a chain "
                    "of dependent shared-memory reads on one warp, far from any pipe's roof",
        },
        "observables_check": {"alpha_mean": float(obs[..., 0].mean() / max(w.n_sites[0], 1)),
                              "rg_mean": float(obs[..., 3].mean()), "ree_mean": float(obs[..., 4].mean())},
        "device": info["name"],
    }
    eng.close()
    if world == 1 and not args.no_secondary and args.scaling == "weak" and args.n_mon == 1000:
        sec = {}
        for name, label, mk, st in (
                ("N100", "128 pH replicas x N=100 chain, same protocol",
                 lambda: workloads.Workload(workloads.example_parameters(100, c_salt=2e-2), pH=np.linspace(2.0, 12.0, 128), device=local_rank), 60),
                ("N10000", "128 pH replicas x N=10000 chain, same protocol",
                 lambda: workloads.Workload(workloads.example_parameters(10000, c_salt=1e-3), pH=np.linspace(2.0, 12.0, 128), device=local_rank), 3),
                ("C3_whole_grid", "BASELINE configs[2] whole on this one GPU: 128 pH x 8 ionic strengths = 1024 replicas x N=1000 "
                                  "(what --scaling strong times at N=1)",
                 lambda: workloads.config_c3_shard(0, 1, n_ph=args.replicas_per_gpu, n_mon=1000, device=local_rank, all_salts=True), args.steps)):
            try:
                sec[name] = secondary_workload(label, mk(), torch, stream, st, 3, min(args.equil_steps, 1000), peaks, sm_count)
            except Exception as exc:  # noqa: BLE001  (the headline must still print)
                sec[name] = {"error": repr(exc)}
        for name, kw in (("C5_box", dict(n_poly=200, n_mon=500, md_steps=1000, trials=2000)),           # 10^5 beads, 4.0e5 particles
                         ("C5_box_1e6_beads", dict(n_poly=2000, n_mon=500, md_steps=300, trials=1000))):   # 10^6 beads, 4.0e6 particles
            try:
                sec[name] = secondary_c5(torch, stream, peaks, local_rank, **kw)
            except Exception as exc:  # noqa: BLE001
                sec[name] = {"error": repr(exc)}
        line["secondary"] = sec
    if not args.no_cpu_baseline:
        try:
            cpu, _ = cpu_measure(w, os.cpu_count() or 1, args.cpu_seconds)
            line["cpu_baseline"] = cpu
        except Exception as exc:  # the GPU line must still print
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {exc!r}"}
    emit(line)


if __name__ == "__main__":
    main()
