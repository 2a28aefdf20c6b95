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


def exchange_pH_labels(pH, n_protonated, rng, sweep):
    """One round of pH replica exchange on the full (gathered) replica set.  Neighbouring pH labels
    a < b of the same ionic strength swap with probability min(1, 10^{(pH_a - pH_b)(n_a - n_b)})
    where n is the number of protonated sites (the constant-pH weight is 10^{-pH n}); even pairs
    on even sweeps, odd pairs on odd sweeps.  Only labels move, never configurations.  `rng` must
    be seeded identically on every rank so that all ranks take the same decisions."""
    pH = np.array(pH, dtype=np.float64, copy=True)
    n = np.asarray(n_protonated, dtype=np.float64)
    order = np.argsort(pH, kind="stable")
    accepted = 0
    for k in range(sweep % 2, len(order) - 1, 2):
        a, b = order[k], order[k + 1]
        log10_w = (pH[a] - pH[b]) * (n[a] - n[b])
        if log10_w >= 0 or rng.random() < 10.0 ** log10_w:
            pH[a], pH[b] = pH[b], pH[a]
            accepted += 1
    return pH, accepted
