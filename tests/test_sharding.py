"""Multi-GPU host logic on CPU: replica sharding, the end-of-run gather (gloo, world size 2) and
pH-label exchange.  The compute stand-in is the CPU oracle; the invariant checked is the one the
GPU path relies on: results are independent of how replicas shard."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from coarse_grain_polyelectrolyte_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    for n, w in ((1024, 8), (10, 3), (5, 8), (128, 1)):
        blocks = [sharding.shard_bounds(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 3, 3)


def test_replica_grid_is_ph_major():
    pH, cs = sharding.replica_grid(np.linspace(2, 12, 128), [1e-3, 1e-2])
    assert pH.shape == (256,) and pH[0] == 2 and pH[127] == 12 and pH[128] == 2
    assert cs[0] == 1e-3 and cs[127] == 1e-3 and cs[128] == 1e-2


def test_label_exchange_conserves_labels_and_prefers_the_right_order():
    rng = np.random.default_rng(0)
    pH = np.linspace(6, 10, 8)
    n_prot = np.array([30, 28, 25, 20, 14, 8, 3, 1.0])          # fewer protons at higher pH: equilibrium order
    out, acc, prop = sharding.exchange_pH_labels(pH, n_prot, rng, sweep=0)
    assert sorted(out) == sorted(pH) and prop == 4
    # swapping two equilibrium neighbours costs weight: (pH_a - pH_b)(n_a - n_b) < 0 -> rarely accepted
    assert acc <= 2
    out, acc, prop = sharding.exchange_pH_labels(pH, n_prot[::-1].copy(), rng, sweep=0)
    assert acc == 4                                               # inverted order: every proposed swap is downhill


def test_label_exchange_stays_inside_an_ionic_strength():
    """On the pH x salt grid of replica_grid a swap happens between pH neighbours of the SAME ionic strength only: every
    salt keeps its multiset of pH labels; pairs of equal pH are not proposed."""
    pH, cs = sharding.replica_grid(np.linspace(4, 11, 8), [1e-3, 2e-2, 2e-1])
    rng = np.random.default_rng(3)
    n_prot = rng.integers(0, 34, pH.size).astype(float)
    cur = pH.copy()
    moved = 0
    for sweep in range(6):
        new, acc, prop = sharding.exchange_pH_labels(cur, n_prot, rng, sweep, group_key=cs)
        assert prop == 3 * (4 if sweep % 2 == 0 else 3)
        for salt in np.unique(cs):
            assert sorted(new[cs == salt]) == sorted(pH[cs == salt])
        moved += acc
        cur = new
    assert moved > 0
    # without a key the grid is one group: equal pH values of different salts are neighbours and are skipped
    _, acc, prop = sharding.exchange_pH_labels(pH, n_prot, np.random.default_rng(1), 0)
    assert prop < pH.size // 2


def test_sample_with_exchange_on_the_oracle(oracle_mod):
    """The driver on one process (no process group): labels stay a permutation, observables come back per round."""
    from helpers import load, make_workload, random_state
    pH, cs = sharding.replica_grid(np.linspace(7.0, 10.0, 4), [2e-2, 1e-1])
    w = workloads_grid(pH, cs)
    e = load(oracle_mod.OracleEngine(w.cfg), w, random_state(w, seed=2))
    mk = lambda: e.sample_params(5, md_steps_per_burst=4, use_md=True, mc_blocks=((0, 71), (1, 11)), p_burst1=0.7, p_burst2=0.1)
    obs, labels, (acc, prop) = sharding.sample_with_exchange(e, mk, pH, cs, n_protonatable=sum(w.n_sites), n_rounds=6)
    assert obs.shape == (8, 30, 6) and labels.shape == (8, 6) and prop == 2 * (3 * 2 + 3 * 1)
    for rnd in range(6):
        for salt in np.unique(cs):
            assert sorted(labels[cs == salt, rnd]) == sorted(pH[cs == salt])


def workloads_grid(pH, cs):
    from coarse_grain_polyelectrolyte_b200 import workloads
    return workloads.Workload(workloads.example_parameters(40, c_salt=float(cs[0])), pH=pH, c_salt=cs)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from helpers import load, make_workload, random_state
    from oracle.oracle import OracleEngine
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n_total = 6
    pH = np.linspace(6.0, 11.0, n_total)
    w_all = make_workload(40, pH=pH)
    state = random_state(w_all, seed=5)
    lo, hi = sharding.shard_bounds(n_total, rank, world)
    w = make_workload(40, pH=pH[lo:hi], replica_offset=lo)
    e = load(OracleEngine(w.cfg), w, tuple(a[lo:hi] for a in state))
    sp = e.sample_params(8, md_steps_per_burst=8, use_md=True, mc_blocks=((0, 41), (1, 11)), p_burst1=0.7, p_burst2=0.2)
    obs, _ = e.sample_loop(sp)
    full = sharding.gather_observables(obs, n_total)
    np.save(os.path.join(out_dir, f"gathered_{rank}.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_matches_single_process(tmp_path, oracle_mod):
    from helpers import load, make_workload, random_state
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = np.load(tmp_path / "gathered_0.npy"), np.load(tmp_path / "gathered_1.npy")
    np.testing.assert_array_equal(g0, g1)
    pH = np.linspace(6.0, 11.0, 6)
    w_all = make_workload(40, pH=pH)
    e = load(oracle_mod.OracleEngine(w_all.cfg), w_all, random_state(w_all, seed=5))
    sp = e.sample_params(8, md_steps_per_burst=8, use_md=True, mc_blocks=((0, 41), (1, 11)), p_burst1=0.7, p_burst2=0.2)
    ref, _ = e.sample_loop(sp)
    np.testing.assert_array_equal(g0, ref)


def _exchange_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from helpers import load, random_state
    from oracle.oracle import OracleEngine
    from coarse_grain_polyelectrolyte_b200 import workloads
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pH, cs = sharding.replica_grid(np.linspace(7.0, 10.0, 4), [2e-2, 1e-1])
    n_total = pH.size
    w_all = workloads.Workload(workloads.example_parameters(40, c_salt=float(cs[0])), pH=pH, c_salt=cs)
    state = random_state(w_all, seed=2)
    lo, hi = sharding.shard_bounds(n_total, rank, world)
    w = workloads.Workload(workloads.example_parameters(40, c_salt=float(cs[lo])), pH=pH[lo:hi], c_salt=cs[lo:hi], replica_offset=lo)
    e = load(OracleEngine(w.cfg), w, tuple(a[lo:hi] for a in state))
    mk = lambda: e.sample_params(5, md_steps_per_burst=4, use_md=True, mc_blocks=((0, 71), (1, 11)), p_burst1=0.7, p_burst2=0.1)  # noqa: E731
    obs, labels, counts = sharding.sample_with_exchange(e, mk, pH[lo:hi], cs[lo:hi], n_protonatable=sum(w.n_sites), n_rounds=6,
                                                        n_total=n_total, lo=lo)
    np.save(os.path.join(out_dir, f"labels_{rank}.npy"), sharding.all_gather_array(labels, n_total))
    np.save(os.path.join(out_dir, f"obs_{rank}.npy"), sharding.gather_observables(obs, n_total))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_replica_exchange_matches_single_process(tmp_path, oracle_mod):
    """Two ranks, half of the pH x salt grid each: the all-gathered counts lead both ranks to the same swaps, and the run
    equals the one-process run bit for bit (global replica index in every RNG key, shared exchange generator)."""
    from helpers import load, random_state
    port = _free_port()
    mp.spawn(_exchange_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    l0, l1 = np.load(tmp_path / "labels_0.npy"), np.load(tmp_path / "labels_1.npy")
    np.testing.assert_array_equal(l0, l1)
    pH, cs = sharding.replica_grid(np.linspace(7.0, 10.0, 4), [2e-2, 1e-1])
    w = workloads_grid(pH, cs)
    e = load(oracle_mod.OracleEngine(w.cfg), w, random_state(w, seed=2))
    mk = lambda: e.sample_params(5, md_steps_per_burst=4, use_md=True, mc_blocks=((0, 71), (1, 11)), p_burst1=0.7, p_burst2=0.1)  # noqa: E731
    obs, labels, _ = sharding.sample_with_exchange(e, mk, pH, cs, n_protonatable=sum(w.n_sites), n_rounds=6)
    np.testing.assert_array_equal(l0, labels)
    np.testing.assert_array_equal(np.load(tmp_path / "obs_0.npy"), obs)
    assert (labels != pH[:, None]).any()            # labels really moved


def test_strong_scaling_cells_are_a_world_independent_balanced_convention():
    """config_c3_shard(all_salts=True): which (pH, ionic strength) cell a global replica index names does not depend on the
    number of ranks (the index is the Philox key), all 1024 cells appear once, and every rank's contiguous block covers every
    ionic strength and the whole pH axis evenly."""
    from coarse_grain_polyelectrolyte_b200 import workloads
    ref = None
    for world in (1, 2, 4, 8):
        cells = []
        for r in range(world):
            w = workloads.config_c3_shard(r, world, n_ph=128, n_mon=30, all_salts=True)
            off = int(w.cfg.replica_offset)
            cells += [(off + k, round(float(w.pH[k]), 6), round(float(w.kappa[k]), 6)) for k in range(w.R)]
            assert abs(w.pH.mean() - 7.0) < 0.35 and len(set(np.round(w.kappa, 6))) == 8
        cells.sort()
        assert len(set(c[1:] for c in cells)) == 1024
        ref = ref or cells
        assert cells == ref
