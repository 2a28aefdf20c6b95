"""Replica sharding over the GPUs of one node.

Replicas r = (i_pH, i_I) are independent Markov chains: rank k of W holds the contiguous block
[k R/W, (k+1) R/W) and there is NO collective in the sampling loop.  torch.distributed is used
once per run to gather the per-replica observable blocks (NCCL over NVLink on the GPUs, gloo in
the CPU tests), and optionally for pH replica exchange, where only the pH LABELS move between
replicas: an all_gather of the per-replica protonation counts, then every rank takes the same
Philox-free deterministic decisions from a shared numpy generator.
"""
import numpy as np


def shard_bounds(n_total, rank, world):
    """[lo, hi) of the contiguous replica block of `rank`; blocks differ by at most one replica."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def replica_grid(pH_values, c_salt_values):
    """pH-major flattening of the pH x ionic-strength grid: replica r -> (pH[r % n_pH], c_salt[r // n_pH])."""
    pH_values, c_salt_values = np.asarray(pH_values, float), np.asarray(c_salt_values, float)
    r = np.arange(pH_values.size * c_salt_values.size)
    return pH_values[r % pH_values.size], c_salt_values[r // pH_values.size]


def gather_observables(local, n_total, group=None, device=None):
    """Concatenate the per-rank observable blocks [R_local, ...] into [n_total, ...] on every rank.
    One all_gather of padded blocks; no-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local)
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    counts = [shard_bounds(n_total, k, world)[1] - shard_bounds(n_total, k, world)[0] for k in range(world)]
    pad = max(counts)
    buf = np.zeros((pad,) + local.shape[1:], dtype=local.dtype)
    buf[: local.shape[0]] = local
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(out, counts)], axis=0)


def all_gather_array(local, n_total, group=None, device=None):
    """gather_observables for any per-replica array (same padding scheme)."""
    return gather_observables(np.asarray(local), n_total, group=group, device=device)


def exchange_pH_labels(pH, n_protonated, rng, sweep, group_key=None):
    """One round of pH replica exchange on the full (gathered) replica set.  Inside every group of replicas that
    share `group_key` (the ionic strength: a swap across ionic strengths would tear the pH x salt grid apart)
    neighbouring pH labels a < b swap with probability min(1, 10^{(pH_a - pH_b)(n_a - n_b)}), n the number of
    protonated sites (the constant-pH weight of a configuration is 10^{-pH n}); even pairs on even sweeps, odd
    pairs on odd sweeps.  Pairs of equal pH are skipped and not counted.  Only labels move, never configurations.
    `rng` must be seeded identically on every rank so that all ranks take the same decisions (one draw per proposed
    pair, in group order).  Returns (new labels, accepted swaps, proposed swaps)."""
    pH = np.array(pH, dtype=np.float64, copy=True)
    n = np.asarray(n_protonated, dtype=np.float64)
    key = np.zeros(pH.size) if group_key is None else np.asarray(group_key, dtype=np.float64)
    accepted = proposed = 0
    for g in np.unique(key):
        members = np.flatnonzero(key == g)
        order = members[np.argsort(pH[members], kind="stable")]
        for k in range(sweep % 2, len(order) - 1, 2):
            a, b = order[k], order[k + 1]
            if pH[a] == pH[b]:
                continue
            proposed += 1
            log10_w = (pH[a] - pH[b]) * (n[a] - n[b])
            u = rng.random()                      # drawn for every proposed pair: the stream does not depend on the outcome
            if log10_w >= 0 or u < 10.0 ** log10_w:
                pH[a], pH[b] = pH[b], pH[a]
                accepted += 1
    return pH, accepted, proposed


def sample_with_exchange(engine, make_sample_params, pH, group_key, n_protonatable, n_rounds, n_total=None, lo=0,
                         seed=2024, dist_group=None, device=None):
    """The fused sampling loop in `n_rounds` launches with one round of pH-label exchange between them (north star:
    "optional pH replica exchange").  Per round and rank: one launch (make_sample_params() -> cgpe_sample_loop),
    then the protonated-site counts of the local replicas and their current labels are all-gathered (a few KB over NCCL, or
    nothing at all on one GPU), every rank takes the same decisions from the same generator, and the local labels go back
    into the engine (cgpe_set_replica_params).  Configurations never move.

    pH, group_key: local arrays [R_local]; lo: global index of local replica 0; n_total: replicas over all ranks.
    Returns obs [R_local][n_rounds * n_samples][6], labels [R_local][n_rounds] (the pH each local replica carried in each
    round) and (accepted, proposed) swap counts."""
    pH = np.array(pH, dtype=np.float64, copy=True)
    group_key = np.asarray(group_key, dtype=np.float64)
    R = pH.size
    n_total = R if n_total is None else int(n_total)
    rng = np.random.default_rng(seed)
    obs_rounds, labels = [], np.empty((R, n_rounds))
    acc = prop = 0
    for rnd in range(n_rounds):
        engine.set_replica_params(pH=pH)
        obs, _ = engine.sample_loop(make_sample_params())
        obs_rounds.append(obs)
        labels[:, rnd] = pH
        n_prot = n_protonatable - (obs[:, -1, 0] + obs[:, -1, 1])      # protonated sites after the round's last sample
        pH_all = all_gather_array(pH, n_total, group=dist_group, device=device)
        n_all = all_gather_array(n_prot, n_total, group=dist_group, device=device)
        key_all = all_gather_array(group_key, n_total, group=dist_group, device=device)
        new_all, a, b = exchange_pH_labels(pH_all, n_all, rng, sweep=rnd, group_key=key_all)
        acc, prop = acc + a, prop + b
        pH = np.array(new_all[lo:lo + R])
    engine.set_replica_params(pH=pH)
    return np.concatenate(obs_rounds, axis=1), labels, (acc, prop)
