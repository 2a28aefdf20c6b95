"""espressomd.observables / accumulators of the facade and the Observables registry (reference:
src/modules/Observables.py).  CPU tests use the host-side description of a System (no engine);
the automatic updates during integrator.run are covered by the GPU test at the end."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "src")]

from coarse_grain_polyelectrolyte_b200 import espressomd  # noqa: E402
from coarse_grain_polyelectrolyte_b200.espressomd import accumulators, observables  # noqa: E402
from modules.Observables import Observables  # noqa: E402


def _system(n=12, L=20.0, seed=1):
    rng = np.random.default_rng(seed)
    s = espressomd.System(box_l=[L, L, L])
    s.time_step = 0.01
    pos = rng.random((n, 3)) * L
    s.part.add(pos=pos, type=[0] * n, q=[0.0] * n, v=rng.normal(size=(n, 3)))
    return s, pos


def test_geometric_observables_match_direct_formulas():
    s, pos = _system()
    ids = [1, 2, 3, 4, 7]
    np.testing.assert_allclose(observables.ParticlePositions(ids=ids).calculate(), pos[ids], rtol=1e-6)
    d = np.diff(pos[ids], axis=0)
    d -= 20.0 * np.rint(d / 20.0)
    np.testing.assert_allclose(observables.ParticleDistances(ids=ids).calculate(), np.linalg.norm(d, axis=1), rtol=1e-5)
    np.testing.assert_allclose(observables.ComPosition(ids=ids).calculate(), pos[ids].mean(0), rtol=1e-6)
    assert observables.ComVelocity(ids=ids).calculate().shape == (3,)
    # torsion of a planar zig-zag is pi (trans), of a U-shaped quadruple 0 (cis)
    s2 = espressomd.System(box_l=[50.0] * 3)
    s2.part.add(pos=[[0, 0, 0], [1, 1, 0], [2, 0, 0], [3, 1, 0], [3, 2, 0], [2, 2, 0]], type=[0] * 6, q=[0.0] * 6)
    phi = observables.BondDihedrals(ids=range(6)).calculate()
    assert phi.shape == (3,)
    np.testing.assert_allclose(np.abs(phi[0]), np.pi, atol=1e-6)
    np.testing.assert_allclose(phi[2], 0.0, atol=1e-6)
    np.testing.assert_allclose(observables.BondAngles(ids=range(3)).calculate(), [np.pi / 2], atol=1e-6)
    with pytest.raises(ValueError):
        observables.ParticlePositions(ids=[99]).calculate()


def test_rdf_of_an_ideal_gas_is_one_and_counts_every_pair():
    s, pos = _system(n=400, L=10.0, seed=3)
    rdf = observables.RDF(ids1=range(400), min_r=0.0, max_r=5.0, n_r_bins=10)
    g = rdf.calculate()
    assert g.shape == (10,) and rdf.bin_centers().shape == (10,)
    assert abs(g[3:].mean() - 1.0) < 0.05
    # pair count recovered from the normalisation
    edges = rdf.bin_edges()
    shell = 4 / 3 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
    n_pairs = 400 * 399 / 2
    d = pos[:, None] - pos[None]
    d -= 10.0 * np.rint(d / 10.0)
    dist = np.sqrt((d ** 2).sum(-1))[np.triu_indices(400, 1)]
    np.testing.assert_allclose(g * shell * n_pairs / 1000.0, np.histogram(dist, edges)[0], atol=1e-6)
    cross = observables.RDF(ids1=range(100), ids2=range(100, 400), min_r=0.5, max_r=4.0, n_r_bins=7).calculate()
    assert cross.shape == (7,) and abs(cross.mean() - 1.0) < 0.1
    with pytest.raises(ValueError):
        observables.RDF(ids1=[0, 1], min_r=2.0, max_r=1.0)


class _Signal:
    """Observable stand-in that replays a prepared series."""

    def __init__(self, series):
        self.series, self.k = np.asarray(series, dtype=np.float64), 0

    def calculate(self):
        self.k += 1
        return self.series[self.k - 1]


def test_time_series_and_mean_variance():
    x = np.random.default_rng(0).normal(size=(50, 3))
    ts = accumulators.TimeSeries(obs=_Signal(x), delta_N=5)
    mv = accumulators.MeanVarianceCalculator(obs=_Signal(x), delta_N=5)
    for _ in range(50):
        ts.update(); mv.update()
    np.testing.assert_array_equal(ts.time_series(), x)
    np.testing.assert_allclose(mv.mean(), x.mean(0))
    np.testing.assert_allclose(mv.variance(), x.var(0, ddof=1))
    ts.clear()
    assert ts.time_series().shape == (0,)


@pytest.mark.parametrize("op", ["scalar_product", "square_distance_componentwise", "componentwise_product"])
def test_multiple_tau_correlator_against_direct_correlation(op):
    rng = np.random.default_rng(5)
    x = np.cumsum(rng.normal(size=(700, 3)), axis=0)
    c = accumulators.Correlator(obs1=_Signal(x), tau_lin=8, tau_max=60.0, delta_N=1, corr_operation=op, dt=0.5)
    for _ in range(len(x)):
        c.update()
    c.finalize()
    lags, res, cnt = c.lag_times(), c.result(), c.sample_sizes()
    assert res.shape == (len(lags), 1 if op == "scalar_product" else 3) and np.all(np.diff(lags) > 0)
    assert lags[-1] >= 60.0 and lags[1] == 0.5

    def direct(series, j):
        a, b = series[:len(series) - j], series[j:]
        if op == "scalar_product":
            return (a * b).sum(-1, keepdims=True).mean(0), len(a)
        if op == "componentwise_product":
            return (a * b).mean(0), len(a)
        return ((b - a) ** 2).mean(0), len(a)

    # level 0 is the exact direct correlation; level l (discard1) that of every 2^l-th sample
    k = 0
    for level in range(10):
        stride = 2 ** level
        sub = x[stride - 1::stride]
        for j in (range(0, 9) if level == 0 else range(5, 9)):
            if k >= len(lags):
                break
            assert lags[k] == j * stride * 0.5
            want, n = direct(sub, j)
            np.testing.assert_allclose(res[k], want, rtol=1e-10)
            assert cnt[k] == n
            k += 1
    assert k == len(lags)
    with pytest.raises(RuntimeError):
        c.update()
    with pytest.raises(ValueError):
        accumulators.Correlator(obs1=_Signal(x), tau_lin=7, tau_max=1.0)


def test_registry_mirrors_the_reference_interface():
    s, _ = _system()
    reg = Observables()
    reg.add_observable("rdf", reg.rdf_observable, range(6), range(6, 12), 0.0, 5.0, 20)
    reg.add_observable("dist", reg.particle_distance_observable, range(12))
    reg.add_observable("dih", reg.dihedral_observable, range(12))
    reg.add_observable("pos", reg.position_observable, range(12))
    reg.add_observable("msd", reg.com_pos_cor, range(12), 10.0)
    reg.add_observable("vacf", reg.com_vel_cor, range(12), 10.0)
    assert all(o.delta_N == 100 for o in reg.observables.values())
    reg.update_accumulators(s)
    assert len(s.auto_update_accumulators) == 6
    for o in reg.observables.values():
        o.update()
    assert reg.observables["rdf"].time_series().shape == (1, 20)
    assert reg.observables["dist"].time_series().shape == (1, 11)
    assert reg.observables["dih"].time_series().shape == (1, 9)
    assert reg.observables["msd"].result().shape == (1, 3) and reg.observables["vacf"].result().shape == (1, 1)
    reg.clear_accumulators(s)
    assert len(s.auto_update_accumulators) == 0 and not reg.observables
    centers, g = Observables.rdf_observable_corrected(np.random.default_rng(1).random(5000) * 5.0, 10, 10, 10.0, bin_count=25)
    assert centers.shape == g.shape == (25,)


@pytest.mark.gpu
def test_accumulators_follow_integrator_run():
    from coarse_grain_polyelectrolyte_b200 import workloads
    from modules.EnsembleDynamics import EnsembleDynamics
    params = workloads.example_parameters(40, **{"Montecarlo properties": {"N_BLOCKS": 1, "DESIRED_BLOCK_SIZE": 14},
                                                 "Simulation configuration": {"ION_MODEL": "implicit"}})
    ens = EnsembleDynamics(params, replicas=dict(pH=[7.0, 9.0], c_salt=[2e-2, 2e-2]))
    s = ens.system
    reg = Observables()
    ids = list(range(ens.n_mon))
    reg.add_observable("bonds", reg.particle_distance_observable, ids)
    reg.add_observable("msd", reg.com_pos_cor, ids, 2.0)
    reg.add_observable("fast", lambda: espressomd.accumulators.TimeSeries(obs=espressomd.observables.ComPosition(ids=ids), delta_N=30))
    reg.update_accumulators(s)
    s.integrator.run(250)
    s.integrator.run(150)          # 400 steps: delta_N=100 -> 4 updates, delta_N=30 -> 13
    bonds = reg.observables["bonds"].time_series()
    assert bonds.shape == (4, 2, ens.n_mon - 1)
    assert reg.observables["fast"].time_series().shape == (13, 2, 3)
    np.testing.assert_allclose(bonds[-1], espressomd.observables.ParticleDistances(ids=ids).calculate(), rtol=1e-6)
    r0 = ens.bond_length_reduced_units.magnitude
    assert np.all(np.abs(bonds - r0) < 0.35 * r0)
    msd = reg.observables["msd"]
    assert msd.result().shape[1:] == (2, 3) and msd.sample_sizes()[0] == 4 and msd.lag_times()[1] == pytest.approx(100 * s.time_step)
    assert np.all(msd.result()[0] == 0.0) and np.all(msd.result()[1] > 0.0)
    # with these accumulators registered perform_sampling stays on the fused loop (the engine records their frames on the
    # device every gcd(100, 30) = 10 steps), and the accumulators keep filling
    n_before = len(reg.observables["fast"].time_series())
    ens.equilibrate_pH()
    ens.perform_sampling()
    assert ens.num_As.shape == (2, ens.num_samples) and np.all(ens.num_B == ens.num_As + ens.num_As2)
    assert np.all(ens.rad_gyr > 0) and len(reg.observables["fast"].time_series()) > n_before
    alpha = ens.titration_curve()[0]
    assert alpha[0] < alpha[1]
    reg.clear_accumulators(s)
    assert len(s.auto_update_accumulators) == 0
    ens.perform_sampling()         # back on the fused loop
    assert ens.num_As.shape == (2, ens.num_samples)


@pytest.mark.gpu
def test_device_frame_recorder_feeds_the_accumulators_like_the_host_path():
    """The accumulators of Observables.py:54-76 (bond lengths, torsions, COM position / velocity correlators) registered on
    two identical systems: one is updated the reference's way -- integrator.run stops every delta_N steps and the observables
    are evaluated on a state read-back -- the other by the frames the MD kernel records on the device inside ONE launch
    (cgpe_frames_*).  Same seeds, and a horizon short enough (60 steps) for the two FP32 trajectories to stay together
    (they part like any two roundings of a chaotic system: 150 ulp after 40 steps, test_md_steps_match_oracle), so the
    accumulated data agree to round-off.  delta_N = 10 and 5 instead of the registry's 100 for that reason."""
    from coarse_grain_polyelectrolyte_b200 import workloads
    from coarse_grain_polyelectrolyte_b200.espressomd import accumulators as A, observables as O
    from modules.EnsembleDynamics import EnsembleDynamics

    def build(device_path):
        params = workloads.example_parameters(40, **{"Montecarlo properties": {"N_BLOCKS": 1, "DESIRED_BLOCK_SIZE": 14},
                                                     "Simulation configuration": {"ION_MODEL": "implicit"}})
        ens = EnsembleDynamics(params, replicas=dict(pH=[7.0, 9.0], c_salt=[2e-2, 2e-2]))
        ids = list(range(ens.n_mon))
        sy = ens.system       # bound explicitly: an unbound observable reads the most recently created System
        accs = {
            "bonds": A.TimeSeries(obs=O.ParticleDistances(ids=ids, system=sy), delta_N=10),
            "dih": A.TimeSeries(obs=O.BondDihedrals(ids=ids, system=sy), delta_N=10),
            "msd": A.Correlator(obs1=O.ComPosition(ids=ids, system=sy), tau_lin=4, tau_max=0.05, delta_N=10,
                                corr_operation="square_distance_componentwise", compress1="discard1"),
            "vacf": A.Correlator(obs1=O.ComVelocity(ids=ids, system=sy), tau_lin=4, tau_max=0.05, delta_N=10,
                                 corr_operation="scalar_product", compress1="discard1"),
            "fast": A.TimeSeries(obs=O.ComVelocity(ids=ids, system=sy), delta_N=5),
        }
        for a in accs.values():
            ens.system.auto_update_accumulators.add(a)
        if not device_path:
            ens.system._frames_plan = lambda: None          # force the step-by-step host path
        return ens, accs

    host, dev = build(False), build(True)
    for ens, _ in (host, dev):
        ens.system.integrator.run(25)
        ens.system.integrator.run(35)
    assert dev[0].system._frames_state is not None and host[0].system._frames_state is None
    assert dev[0].system._engine.last_call_ms()[1] == 1          # the 35 steps were one launch
    for name in ("bonds", "dih", "fast"):
        a, b = host[1][name].time_series(), dev[1][name].time_series()
        assert a.shape == b.shape and a.shape[0] == (12 if name == "fast" else 6), (name, a.shape, b.shape)
        if name == "dih":
            assert np.abs(np.angle(np.exp(1j * (a - b)))).max() < 2e-3
        else:
            np.testing.assert_allclose(b, a, rtol=5e-4, atol=5e-4)
    for name in ("msd", "vacf"):
        ch, cd = host[1][name], dev[1][name]
        np.testing.assert_array_equal(ch.lag_times(), cd.lag_times())
        np.testing.assert_array_equal(ch.sample_sizes(), cd.sample_sizes())
        np.testing.assert_allclose(cd.result(), ch.result(), rtol=5e-3, atol=1e-6)


@pytest.mark.gpu
def test_registered_observables_keep_the_fused_sampling_loop():
    """perform_sampling with the chain observables of Observables.py registered stays ONE launch; the accumulators receive the
    frames recorded every delta_N steps of every replica's own bursts.  COM diffusion: MSD(tau) of the recorded COM positions
    grows linearly, 6 D tau with D = kT / (N gamma) (Rouse: all beads carry the same friction)."""
    from coarse_grain_polyelectrolyte_b200 import workloads
    from modules.EnsembleDynamics import EnsembleDynamics
    params = workloads.example_parameters(40, **{"Montecarlo properties": {"N_BLOCKS": 8, "DESIRED_BLOCK_SIZE": 35},
                                                 "Simulation configuration": {"ION_MODEL": "implicit"}})
    ens = EnsembleDynamics(params, replicas=dict(pH=[7.0, 8.0, 9.0, 10.0], c_salt=[2e-2] * 4))
    ids = list(range(ens.n_mon))
    reg = Observables()
    reg.add_observable("msd", reg.com_pos_cor, ids, 8.0)
    reg.add_observable("bonds", reg.particle_distance_observable, ids)
    reg.update_accumulators(ens.system)
    ens.equilibrate_pH()
    ens.perform_sampling()
    eng = ens.system._engine
    assert eng.last_call_ms()[1] == 1 and eng.engine_info()["engine"] == "solo"
    assert ens.num_As.shape == (4, ens.num_samples)
    msd = reg.observables["msd"]
    n_frames = msd.sample_sizes()[0]
    md_steps = np.rint(ens.times[:, -1] / ens.system.time_step)               # MD steps of the sampling loop, per replica
    assert abs(n_frames - (int(md_steps.min() // 100) + 10)) <= 1             # + the 1000 steps of equilibrate_pH
    assert n_frames > 0.5 * ens.num_samples * 5 * 0.7
    bonds = reg.observables["bonds"].time_series()
    assert bonds.shape == (n_frames, 4, ens.n_mon - 1)
    assert abs(bonds.mean() - ens.bond_length_reduced_units.magnitude) < 0.1
    tau, res = msd.lag_times(), msd.result().sum(axis=-1).mean(axis=1)       # sum over x, y, z; mean over replicas
    sel = (tau > 0.5) & (tau < 4.0)
    slope = np.polyfit(tau[sel], res[sel], 1)[0]
    d_rouse = 1.0 / ens.n_mon
    assert 0.5 * 6 * d_rouse < slope < 2.0 * 6 * d_rouse, (slope, 6 * d_rouse)
