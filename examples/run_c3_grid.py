#!/usr/bin/env python
"""BASELINE configs[2] end to end: the 128 pH x 8 ionic-strength grid of N = 1000 chains (1024 replicas) through the
drop-in driver, sharded over the GPUs of one node.  The reference runs its pH sweep as a serial notebook loop
(run_test-checkpoint.ipynb, cell 4 = NB:276-326); here every (pH, salt) point is a replica, every rank holds a contiguous
block of them, nothing is communicated inside the sampling loop, and ONE NCCL all_gather at the end brings the observable
blocks together.  Rank 0 prints the titration table.

  python examples/run_c3_grid.py --samples 200                               # one GPU: all 1024 replicas
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
         examples/run_c3_grid.py                                             # eight GPUs: 128 replicas each, full protocol
  ... --exchange-every 50                                                    # pH replica exchange between launches

Options: --n-ph, --n-mon, --samples (default: the protocol's 4571), --exchange-every K (samples per launch; 0 = off).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "src")):
    sys.path.insert(0, p)
os.environ.setdefault("CGPE_QUIET", "1")

from coarse_grain_polyelectrolyte_b200 import sharding, workloads  # noqa: E402
from modules.EnsembleDynamics import EnsembleDynamics  # noqa: E402
from modules.utils import ideal_alpha  # noqa: E402


def binning_sem(x, n_bins=16):
    """NB:727-731."""
    n = (x.shape[-1] // n_bins) * n_bins
    m = x[..., :n].reshape(x.shape[:-1] + (n_bins, -1)).mean(-1)
    return m.std(-1, ddof=1.5) / np.sqrt(n_bins)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-ph", type=int, default=128)
    ap.add_argument("--n-mon", type=int, default=1000)
    ap.add_argument("--samples", type=int, default=0, help="samples per replica (0: the protocol's num_samples, 4571)")
    ap.add_argument("--exchange-every", type=int, default=0)
    ap.add_argument("--salts", type=int, default=8, help="first n of the eight ionic strengths")
    args = ap.parse_args()

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)

    pH_all, cs_all = sharding.replica_grid(np.linspace(2.0, 12.0, args.n_ph), workloads.C3_SALT_MOLAR[:args.salts])
    n_total = pH_all.size
    lo, hi = sharding.shard_bounds(n_total, rank, world)
    params = workloads.example_parameters(args.n_mon, c_salt=float(cs_all[lo]))
    params["Simulation configuration"]["ION_MODEL"] = "implicit"      # the north star's move; "explicit" is the reference's own model
    t0 = time.time()
    ens = EnsembleDynamics(params, replicas=dict(pH=pH_all[lo:hi], c_salt=cs_all[lo:hi]), replica_offset=lo, device=local)
    t_setup = time.time() - t0
    ens.equilibrate_pH()
    if args.samples > 0:
        ens.num_samples = args.samples
        ens.n_samp_iter = max(1, (ens.num_samples - 10) // 2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    swaps = None
    if args.exchange_every > 0:
        eng = ens.system._ensure_engine()
        rounds = max(1, ens.num_samples // args.exchange_every)
        p = ens.probability_reaction

        def make_sp():
            return eng.sample_params(args.exchange_every, md_steps_per_burst=ens.number_integration_steps, use_md=ens.USE_WCA,
                                     mc_blocks=((0, 5 * ens.n_nh + 1), (1, 5 * ens.n_nh2 + 1)), p_burst1=p,
                                     p_burst2=p * 2.0 / (ens.n_mon - 2), schedule_seed=10)
        obs, labels, swaps = sharding.sample_with_exchange(eng, make_sp, pH_all[lo:hi], cs_all[lo:hi], ens.n_nh + ens.n_nh2, rounds,
                                                           n_total=n_total, lo=lo, device=device)
        block = np.concatenate([obs[..., [0, 1, 3, 4]], np.repeat(labels, args.exchange_every, axis=1)[..., None]], axis=-1)
    else:
        ens.perform_sampling()
        block = np.stack([ens.num_As, ens.num_As2, ens.rad_gyr, ens.end_2end,
                          np.broadcast_to(pH_all[lo:hi, None], np.shape(ens.num_As))], axis=-1)
    torch.cuda.synchronize()
    t_local = time.time() - t0
    full = sharding.gather_observables(np.ascontiguousarray(block), n_total, device=device)     # the ONE collective of the path
    if world > 1:
        dist.barrier()
    t_all = time.time() - t0
    if rank == 0:
        n_samples = full.shape[1]
        skip = n_samples // 10
        print(f"{n_total} replicas ({args.n_ph} pH x {args.salts} ionic strengths) x N={args.n_mon} on {world} GPU(s), {n_samples} samples per "
              f"replica; setup+warm-up {t_setup:.1f} s, sampling {t_local:.2f} s on rank 0, {t_all:.2f} s incl. the NCCL gather "
              f"({full.nbytes / 1e6:.1f} MB)" + (f"; label swaps accepted {swaps[0]} of {swaps[1]}" if swaps else ""))
        md_steps = n_total * n_samples * ens.number_integration_steps * ens.probability_reaction * (1 + 2.0 / (ens.n_mon - 2))
        print(f"~{md_steps * args.n_mon / t_all:.3g} MD bead-steps/s over the wall time (expected number of steps of the schedule)")
        for s, salt in enumerate(workloads.C3_SALT_MOLAR[:args.salts]):
            rows = full[s * args.n_ph:(s + 1) * args.n_ph, skip:, :]
            print(f"\n# c_salt = {salt:g} M\n#   pH   alpha_NH  +-sem   ideal |  alpha_NH2 |     Rg  +-sem     Ree")
            if args.exchange_every > 0:     # samples follow their pH label
                labels_flat, a_flat = rows[..., 4].ravel(), rows[..., 0].ravel() / ens.n_nh
                for pH in np.unique(labels_flat)[::max(1, args.n_ph // 16)]:
                    m = labels_flat == pH
                    print(f"{pH:6.2f} {a_flat[m].mean():9.4f}    -    {ideal_alpha(pH, ens.pK):7.4f} |")
            else:
                for r in range(0, args.n_ph, max(1, args.n_ph // 16)):
                    a, a2 = rows[r, :, 0] / ens.n_nh, rows[r, :, 1] / ens.n_nh2
                    print(f"{rows[r, 0, 4]:6.2f} {a.mean():9.4f} {binning_sem(a):7.4f} {ideal_alpha(rows[r, 0, 4], ens.pK):7.4f} | {a2.mean():9.4f}  |"
                          f" {rows[r, :, 2].mean():7.2f} {binning_sem(rows[r, :, 2]):6.2f} {rows[r, :, 3].mean():8.2f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
