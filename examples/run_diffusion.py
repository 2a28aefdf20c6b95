#!/usr/bin/env python
"""Centre-of-mass diffusion of the chain from the accumulators the reference registers (Observables.py:66-76), analysed the
way its notebook does (run_test-checkpoint.ipynb cells 23-29 = NB:727-936):

  * MSD(tau) of the COM from the multiple-tau correlator (square_distance_componentwise, summed over x, y, z), weighted
    linear fit over the diffusive window, D = slope / 6                                   (NB: "Diffusion = 0.00933", N = 52);
  * velocity autocorrelation of the COM, exponential fit a exp(-b tau)                    (NB: a = 0.0608, b = 0.980);
  * Green-Kubo: D(tau_int) = 1/3 int_0^tau_int <v(0) v(tau)> dtau, mean over the plateau.

The accumulators are fed by frames the MD kernel records on the device every delta_N = 100 steps, so the sampling loop of
perform_sampling stays ONE launch (the reference makes one engine call per update).  For a Langevin chain with the same
friction on every bead the COM is an Ornstein-Uhlenbeck particle: <v_c^2> = 3 kT / N, decay rate gamma, D = kT / (N gamma).

  python examples/run_diffusion.py [--n-mon 52] [--n-ph 4] [--samples 1333]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "src")):
    sys.path.insert(0, p)
os.environ.setdefault("CGPE_QUIET", "1")

from coarse_grain_polyelectrolyte_b200 import workloads  # noqa: E402
from modules.EnsembleDynamics import EnsembleDynamics  # noqa: E402
from modules.Observables import Observables  # noqa: E402


def diffusion_from_msd(tau, msd, weights, i_lo, i_hi):
    """NB cell 25: weighted linear fit of MSD(tau) over [i_lo, i_hi); D = slope / 6."""
    a, b = np.polyfit(tau[i_lo:i_hi], msd[i_lo:i_hi], 1, w=weights[i_lo:i_hi])
    return a / 6.0, (a, b)


def fit_vacf(tau, acf, n=85):
    """NB cell 28: a exp(-b tau) fitted to the first n lags by non-linear least squares (scipy's curve_fit, as the notebook);
    without scipy: log-linear fit over the leading run above a fifth of the zero-lag value."""
    t, y = np.asarray(tau[:n], float), np.asarray(acf[:n], float)
    lead = int(np.argmax(y < 0.2 * y[0])) or len(y)
    slope, icpt = np.polyfit(t[:lead], np.log(y[:lead]), 1)
    a0, b0 = float(np.exp(icpt)), float(-slope)
    try:
        from scipy.optimize import curve_fit
        (a, b), _ = curve_fit(lambda x, a, b: a * np.exp(-b * x), t, y, p0=(a0, b0))
        return float(a), float(b)
    except Exception:      # noqa: BLE001
        return a0, b0


def green_kubo(tau, acf, tau_lo=3.0, tau_hi=6.0):
    """NB cell 29: running integral D(tau_int) = 1/3 int_0^tau_int <v(0)v(tau)> dtau; the plateau value is its mean over
    tau_int in [tau_lo, tau_hi] (a few decay times: beyond, the integral only collects the noise of the correlator's tail)."""
    tau, acf = np.asarray(tau, float), np.asarray(acf, float)
    y = np.array([np.trapezoid(acf[:j], tau[:j]) / 3.0 for j in range(2, len(tau) + 1)])
    t_int = tau[1:]
    sel = (t_int >= tau_lo) & (t_int <= tau_hi)
    return float(np.mean(y[sel])) if sel.any() else float(y[-1]), y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-mon", type=int, default=52)
    ap.add_argument("--n-ph", type=int, default=4)
    ap.add_argument("--samples", type=int, default=1333)
    ap.add_argument("--tau-max", type=float, default=1000.0)
    args = ap.parse_args()
    params = workloads.example_parameters(args.n_mon, c_salt=1e-2)
    params["Simulation configuration"]["ION_MODEL"] = "implicit"
    pH = np.linspace(4.0, 10.0, args.n_ph)
    ens = EnsembleDynamics(params, replicas=dict(pH=pH, c_salt=[1e-2] * args.n_ph))
    ids = list(range(ens.n_mon))
    reg = Observables()
    reg.add_observable("com_pos_cor", reg.com_pos_cor, ids, args.tau_max)      # NB cell 7
    reg.add_observable("com_vel_cor", reg.com_vel_cor, ids, args.tau_max)
    reg.update_accumulators(ens.system)
    ens.equilibrate_pH()
    ens.num_samples = args.samples
    ens.n_samp_iter = max(1, (ens.num_samples - 10) // 2)
    ens.perform_sampling()
    launches = ens.system._engine.last_call_ms()[1]
    msd_c, vel_c = reg.observables["com_pos_cor"], reg.observables["com_vel_cor"]
    tau = msd_c.lag_times()
    msd = msd_c.result().sum(axis=-1)                    # NB cell 9: sum over x, y, z  -> [lags][R]
    acf = vel_c.result()[..., 0]                         # [lags][R]
    w = msd_c.sample_sizes()
    if msd.ndim == 1:
        msd, acf = msd[:, None], acf[:, None]
    print(f"N = {ens.n_mon}, {args.n_ph} pH replicas, {args.samples} samples each in {launches} launch(es); "
          f"{w[0]} frames per replica; lags {tau[1]:.3g} .. {tau[-1]:.3g}")
    d_rouse = 1.0 / ens.n_mon
    print(f"theory (Rouse / OU): <v_c^2> = 3 kT/N = {3 * d_rouse:.4f}, decay rate gamma = 1, D = kT/(N gamma) = {d_rouse:.5f}")
    print("   pH    D_MSD     D_GK    <v_c^2>(fit a)  rate(fit b)")
    n = len(tau)
    i_lo, i_hi = min(40, n // 2), min(75, n)
    for r, p in enumerate(pH):
        d_msd, _ = diffusion_from_msd(tau, msd[:, r], w, i_lo, i_hi)
        a, b = fit_vacf(vel_c.lag_times(), acf[:, r])
        d_gk, _ = green_kubo(vel_c.lag_times(), acf[:, r])
        print(f"{p:6.2f} {d_msd:9.5f} {d_gk:9.5f} {a:12.4f} {b:12.3f}")
    print("notebook (ESPResSo, N = 52 with explicit ions): D_MSD = 0.00933, a = 0.0608, b = 0.980")


if __name__ == "__main__":
    main()
