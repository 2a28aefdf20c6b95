#!/usr/bin/env python
"""Titration curve of a weak polyelectrolyte chain on one B200.

The reference's examples/run_test.py builds the N=100 system, equilibrates at pH 10 and integrates;
its notebook then loops serially over pH values (equilibrate_pH + perform_sampling per point).  Here
the whole pH axis runs as one batch of replicas: one equilibration call, one sampling call.

    python examples/run_titration.py --n-ph 64 --n-mon 100 [--explicit-ions] [--blocks 16]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "src"), ROOT]
os.environ.setdefault("CGPE_QUIET", "1")

from coarse_grain_polyelectrolyte_b200 import workloads  # noqa: E402
from modules.EnsembleDynamics import EnsembleDynamics  # noqa: E402
from modules.utils import ideal_alpha  # noqa: E402


def binning_sem(x, n_bins=16):
    """Standard error of the mean from a binning analysis (the notebook's error bar, NB:727-731)."""
    n = (len(x) // n_bins) * n_bins
    means = np.asarray(x[:n]).reshape(n_bins, -1).mean(axis=1)
    return np.std(means, ddof=1.5) / np.sqrt(n_bins)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-mon", type=int, default=100)
    ap.add_argument("--n-ph", type=int, default=64)
    ap.add_argument("--ph-min", type=float, default=4.0)
    ap.add_argument("--ph-max", type=float, default=12.0)
    ap.add_argument("--c-salt", type=float, default=2e-2)
    ap.add_argument("--blocks", type=int, default=16, help="N_BLOCKS (examples/run_test.py: 16)")
    ap.add_argument("--explicit-ions", action="store_true", help="the reference's B+/Na+/Cl- particles instead of implicit ions")
    args = ap.parse_args()

    params = workloads.example_parameters(args.n_mon, c_salt=args.c_salt,
                                          **{"Montecarlo properties": {"N_BLOCKS": args.blocks}})
    # the drop-in's default is the reference's explicit-ion model; this script shows the batched implicit-ion path by default
    params["Simulation configuration"]["ION_MODEL"] = "explicit" if args.explicit_ions else "implicit"
    pH = np.linspace(args.ph_min, args.ph_max, args.n_ph)
    t0 = time.time()
    ens = EnsembleDynamics(params, replicas=dict(pH=pH, c_salt=[args.c_salt] * args.n_ph))
    t1 = time.time()
    ens.equilibrate_pH()
    ens.perform_sampling()
    t2 = time.time()
    a1, a2 = ens.titration_curve()
    print(f"# N={args.n_mon}, {args.n_ph} pH replicas, {ens.num_samples} samples each; setup+warm-up {t1 - t0:.1f} s, "
          f"equilibration+sampling {t2 - t1:.1f} s")
    print("#   pH   alpha_NH   +-sem   ideal   alpha_NH2  ideal      Rg   +-sem     Ree")
    for r in range(args.n_ph):
        print(f"{pH[r]:6.2f}  {a1[r]:8.4f} {binning_sem(ens.num_As[r]) / ens.n_nh:7.4f} {ideal_alpha(pH[r], ens.pK):7.4f}"
              f"  {a2[r]:8.4f} {ideal_alpha(pH[r], ens.pK2):7.4f}  {ens.rad_gyr[r].mean():7.3f} {binning_sem(ens.rad_gyr[r]):7.3f}"
              f" {ens.end_2end[r].mean():8.3f}")


if __name__ == "__main__":
    main()
