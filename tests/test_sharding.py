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
    out, acc = sharding.exchange_pH_labels(pH, n_prot, rng, sweep=0)
    assert sorted(out) == sorted(pH)
    # swapping two equilibrium neighbours costs weight: (pH_a - pH_b)(n_a - n_b) < 0 -> rarely accepted
    assert acc <= 2
    out, acc = sharding.exchange_pH_labels(pH, n_prot[::-1].copy(), rng, sweep=0)
    assert acc == 4                                               # inverted order: every proposed swap is downhill


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
