"""Parity of the CUDA hot path (through the C-ABI of libcgpe.so) with the FP64 CPU oracle on the
same seeded inputs.

Tolerances (BASELINE.json north_star): energies and forces within 1e-5 relative; integer
charge-state bookkeeping bit-exact for identical proposal sequences.  Relative force error is
measured against the largest force in the configuration (a per-component ratio is meaningless
where a component passes through zero).
"""
import os

import numpy as np
import pytest

from coarse_grain_polyelectrolyte_b200 import _abi
from helpers import load, make_workload, random_state

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5


@pytest.fixture(scope="module", params=["duo", "solo"])
def Engine(request):
    """Every test of this module runs twice: on the cluster engine (two CTAs per replica) wherever it can serve the
    configuration -- otherwise on whatever the library chooses, e.g. the global-memory engine -- and pinned to the
    one-CTA-per-replica engine.  bench.py's headline (128 x N=1000 at 1 mM) runs on the cluster engine."""
    from coarse_grain_polyelectrolyte_b200.engine import Engine as E
    from coarse_grain_polyelectrolyte_b200._abi import CgpeError

    class Pinned(E):
        def set_state(self, *a, **k):
            super().set_state(*a, **k)
            try:
                self.set_engine(request.param)
            except CgpeError:            # the cluster engine does not serve this configuration (ions, spectators, size)
                self.set_engine("auto")
    return Pinned


def _pair(Engine, oracle_mod, w, state):
    g = load(Engine(w.cfg), w, state)
    o = load(oracle_mod.OracleEngine(w.cfg), w, state)
    return g, o


@pytest.mark.parametrize("n_mon,flags", [
    (50, None),
    (100, None),
    (100, {"USE_FENE": True}),
    (100, {"USE_BENDING": False, "USE_DIHEDRAL_POT": False}),
    (100, {"USE_ELECTROSTATICS": False}),
    (1000, None),
    (1500, None),
    (2048, None),      # the largest replica of the shared-memory engine (two particles per thread)
    (2049, None),      # the first one of the global-memory (wide) engine
    (2500, None),
])
def test_energy_and_forces_match_oracle(Engine, oracle_mod, n_mon, flags):
    w = make_workload(n_mon, pH=[6.0, 8.0, 10.0], flags=flags)
    state = random_state(w, seed=n_mon, jitter=0.03)
    g, o = _pair(Engine, oracle_mod, w, state)
    eg, eo = g.energy(), o.energy()
    scale = np.maximum(np.abs(eo), 1e-3 * np.abs(eo).max(axis=1, keepdims=True))
    assert np.max(np.abs(eg - eo) / scale) < REL_TOL, (eg, eo)
    fo = o.forces_f64()
    fmax = np.abs(fo).max()
    for name, fg in (("all-pairs", g.forces()), ("md-lists", g.md_forces())):
        err = np.abs(fg - fo).max() / fmax
        assert err < REL_TOL, f"{name}: max |dF|/max|F| = {err:.3g}"
    # analysis helpers
    np.testing.assert_allclose(g.min_dist(), o.min_dist(), rtol=1e-6)
    rg_g, ree_g, cnt_g = g.observables()
    rg_o, ree_o, cnt_o = o.observables()
    np.testing.assert_allclose(rg_g, rg_o, rtol=1e-6)
    np.testing.assert_allclose(ree_g, ree_o, rtol=1e-6)
    np.testing.assert_array_equal(cnt_g, cnt_o)


@pytest.mark.parametrize("n_mon,n_trials,c_salt", [(50, 4000, 2e-2), (100, 6000, 2e-2), (1000, 20000, 2e-2),
                                                   (1000, 12000, 1e-3),   # long Debye-Hueckel rows: helper warps update the potentials
                                                   (2500, 6000, 2e-2)])
def test_mc_bookkeeping_bit_exact(Engine, oracle_mod, n_mon, n_trials, c_salt):
    """Same Philox proposal stream on both sides -> identical sites, directions, decisions,
    types and charges; Delta-E within 1e-5 of the FP64 value.  A decision may differ only on a
    near-tie |u - bf| <= 1e-4 bf, which the oracle then follows and the test counts."""
    w = make_workload(n_mon, pH=[7.0, 8.18, 9.0, 10.5], c_salt=c_salt)
    state = random_state(w, seed=7 + n_mon, jitter=0.03)
    g, o = _pair(Engine, oracle_mod, w, state)
    for rx in (0, 1, 0):
        tg = g.mc_run(rx, n_trials, trace=True)
        to, ties = o.mc_run_follow(rx, tg, tie_tol=1e-4)
        assert ties <= max(2, n_trials * w.R // 2000), f"{ties} near-ties"
        np.testing.assert_array_equal(tg["particle"], to["particle"])
        np.testing.assert_array_equal(tg["direction"], to["direction"])
        np.testing.assert_array_equal(tg["accepted"], to["accepted"])
        np.testing.assert_array_equal(tg["u"], to["u"])
        tried = to["particle"] >= 0
        # Delta-E enters the decision as exp(-dE/kT): what matters is its absolute error in kT
        # (a WCA type swap is a difference of two O(1) terms, so |dE| itself can be tiny)
        err = np.abs(tg["delta_e"][tried] - to["delta_e"][tried]) / np.maximum(np.abs(to["delta_e"][tried]), 1.0)
        k = int(np.argmax(err))
        assert err[k] < REL_TOL, (err[k], tg[tried][k], to[tried][k])
        assert tg["accepted"].sum() > 0
    xg, _, tyg = g.get_state()
    xo, _, tyo = o.get_state()
    np.testing.assert_array_equal(tyg, tyo)
    np.testing.assert_array_equal(xg[..., 3], xo[..., 3])
    cg, co = g.counters(), o.counters()
    np.testing.assert_array_equal(cg["trials"], co["trials"])
    np.testing.assert_array_equal(cg["accepted"], co["accepted"])
    np.testing.assert_array_equal(g.observables()[2], o.observables()[2])


@pytest.mark.parametrize("n_mon", [50, 100, 1000, 1500, 2500])   # 1500: two particles per thread; 2500: global-memory engine
def test_md_steps_match_oracle(Engine, oracle_mod, n_mon):
    """One velocity-Verlet + Langevin step is a deterministic map given the Philox noise; FP32
    vs FP64 differ by rounding only.  The gap then grows with the Lyapunov exponent."""
    w = make_workload(n_mon, pH=[8.0, 9.0])
    state = random_state(w, seed=3 + n_mon, jitter=0.02)
    g, o = _pair(Engine, oracle_mod, w, state)
    x0 = state[0][..., :3].astype(np.float64)
    ulp = float(np.spacing(np.float32(np.abs(x0).max())))   # FP32 grid of the stored coordinates
    for n_steps, tol in ((1, 0.75 * ulp), (9, 8 * ulp), (40, 150 * ulp)):
        g.md_run(n_steps)
        o.md_run(n_steps)
        xg, vg, _ = g.get_state()
        xo, vo, _ = o.state_f64()[0], o.state_f64()[1], None
        dx = np.abs(xg[..., :3] - xo).max()
        dv = np.abs(vg[..., :3] - vo).max()
        assert dx < tol and dv < tol * 1e3, f"after +{n_steps} steps: dx={dx:.3g} dv={dv:.3g}"
    assert np.abs(xo - x0).max() > 1e-3     # the system actually moved
    np.testing.assert_array_equal(g.counters()["md_steps"], o.counters()["md_steps"])
    np.testing.assert_allclose(g.get_time(), o.get_time(), rtol=1e-12)


def test_md_then_mc_then_md_reuses_state(Engine, oracle_mod):
    """Interleaving as in equilibrate_pH / perform_sampling: forces are recomputed after a
    reaction block, kept across back-to-back integrator calls."""
    w = make_workload(100, pH=[8.5])
    state = random_state(w, seed=99, jitter=0.02)
    g, o = _pair(Engine, oracle_mod, w, state)
    for e in (g, o):
        e.md_run(5)
        e.md_run(5)
        e.mc_run(0, 200)
        e.mc_run(1, 20)
        e.md_run(5)
    xg, vg, tg = g.get_state()
    xo, vo, _ = o.state_f64()
    _, _, to = o.get_state()
    np.testing.assert_array_equal(tg, to)
    assert np.abs(xg[..., :3] - xo).max() < 5e-5
    assert np.abs(vg[..., :3] - vo).max() < 5e-2


def test_steepest_descent_matches_oracle(Engine, oracle_mod):
    """integrator.set_steepest_descent (M.py:78-79).  With the reference's gamma=0.1 and
    k_bond ~ 2.4e3 the clamped update is an expanding map (gamma*k >> 1), so FP32 and FP64 can
    only be compared over a single step there; the contracting regime is compared over 40."""
    w = make_workload(100, pH=[8.0, 9.0])
    state = random_state(w, seed=5, jitter=0.1, vel_sigma=0.0)
    g, o = _pair(Engine, oracle_mod, w, state)
    e0 = o.energy()[:, 0].copy()
    for e in (g, o):
        e.steepest_descent(40, f_max=0.0, gamma=1e-4, max_displacement=0.01)
    xg = g.get_state()[0][..., :3]
    xo = o.state_f64()[0]
    # every clamped step adds the same rounding error of x += 0.01 on the FP32 grid (~2e-6 at x~36)
    assert np.abs(xg - xo).max() < 1.5e-4
    assert np.all(o.energy()[:, 0] < e0)
    g1, o1 = _pair(Engine, oracle_mod, w, state)
    for e in (g1, o1):
        e.steepest_descent(1, f_max=0.0, gamma=0.1, max_displacement=0.1)
    # unclamped components move by gamma*f: FP32 force error 3e-7*|F|max ~ 1e-4 -> 1e-5 in x
    assert np.abs(g1.get_state()[0][..., :3] - o1.state_f64()[0]).max() < 5e-5
    # f_max above every force: converged after the first displacement, no further motion
    g2, o2 = _pair(Engine, oracle_mod, w, state)
    for e in (g2, o2):
        e.steepest_descent(25, f_max=1e9, gamma=1e-4, max_displacement=0.01)
    assert np.abs(g2.get_state()[0][..., :3] - o2.state_f64()[0]).max() < 1e-5


def test_force_cap_and_zero_temperature(Engine, oracle_mod):
    """Warm-up mode (E.py:24-52): capped forces, kT = 0 thermostat."""
    w = make_workload(50, pH=[8.0])
    state = random_state(w, seed=11, jitter=0.2, vel_sigma=0.0)
    g, o = _pair(Engine, oracle_mod, w, state)
    for e in (g, o):
        e.set_force_cap(5.0)
        e.set_replica_params(kT=0.0, gamma=1.0)
        e.set_thermostat_seed(4)
        e.md_run(12)
    xg, vg, _ = g.get_state()
    xo, vo, _ = o.state_f64()
    assert np.abs(xg[..., :3] - xo).max() < 5e-5
    assert np.abs(vg[..., :3] - vo).max() < 5e-3


def test_sample_loop_matches_oracle(Engine, oracle_mod):
    """The fused perform_sampling loop against the oracle's: integer columns identical, Rg/Ree
    and time to rounding.  Short bursts keep the trajectories in the linear-error regime."""
    w = make_workload(60, pH=[7.5, 8.5, 9.5])
    state = random_state(w, seed=21, jitter=0.02)
    g, o = _pair(Engine, oracle_mod, w, state)
    sp = g.sample_params(40, md_steps_per_burst=6, use_md=True, mc_blocks=((0, 31), (1, 11)),
                         q_snapshot_every=15, q_snapshot_rows=3, schedule_seed=10, p_burst1=0.7, p_burst2=0.3)
    og, qg = g.sample_loop(sp)
    oo, qo = o.sample_loop(sp)
    np.testing.assert_array_equal(og[..., :3], oo[..., :3])
    np.testing.assert_allclose(og[..., 3:5], oo[..., 3:5], rtol=2e-4)
    np.testing.assert_allclose(og[..., 5], oo[..., 5], rtol=1e-9)
    np.testing.assert_array_equal(qg, qo)
    assert og[..., 5].max() > 0 and len(np.unique(og[..., 0])) > 1
    cg, co = g.counters(), o.counters()
    for k in ("md_steps", "trials", "accepted"):
        np.testing.assert_array_equal(cg[k], co[k])


def test_results_do_not_depend_on_sharding(Engine):
    """RNG streams are keyed on the global replica index: 8 replicas in one handle == 2 handles
    of 4 with replica_offset 0 and 4, bit for bit."""
    pH = np.linspace(6.0, 11.0, 8)
    w_all = make_workload(100, pH=pH)
    state = random_state(w_all, seed=31)
    sp_kw = dict(md_steps_per_burst=25, use_md=True, mc_blocks=((0, 51), (1, 11)), schedule_seed=10,
                 p_burst1=0.7, p_burst2=0.2)
    g = load(Engine(w_all.cfg), w_all, state)
    ref, _ = g.sample_loop(g.sample_params(30, **sp_kw))
    parts = []
    for lo in (0, 4):
        w = make_workload(100, pH=pH[lo:lo + 4], replica_offset=lo)
        sub = tuple(a[lo:lo + 4] for a in state)
        e = load(Engine(w.cfg), w, sub)
        parts.append(e.sample_loop(e.sample_params(30, **sp_kw))[0])
    np.testing.assert_array_equal(ref, np.concatenate(parts, axis=0))


def test_ideal_titration_curve(Engine):
    """Interactions off: alpha(pH) must follow 1/(1+10^(pK-pH)) (utils.py:15-16 of the
    reference) within statistical error."""
    pH = np.linspace(6.0, 12.0, 16)
    w = make_workload(100, pH=pH, flags={"USE_WCA": False, "USE_ELECTROSTATICS": False})
    e = load(Engine(w.cfg), w, (np.broadcast_to(w.xyzq, (16, 100, 4)).copy(), None, np.broadcast_to(w.type, (16, 100)).copy()))
    n1, n2 = w.n_sites
    sp = e.sample_params(3000, use_md=False, mc_blocks=((0, 5 * n1 + 1), (1, 5 * n2 + 1)))
    obs, _ = e.sample_loop(sp)
    obs = obs[:, 200:]
    for col, n, pK in ((0, n1, 8.18), (1, n2, 10.02)):
        alpha = obs[..., col].mean(axis=1) / n
        ideal = 1.0 / (1.0 + 10.0 ** (pK - pH))
        sem = np.sqrt(ideal * (1 - ideal) / (n * obs.shape[1] / 4.0)) + 2e-3
        assert np.all(np.abs(alpha - ideal) < 5 * sem), (alpha, ideal)
    np.testing.assert_array_equal(obs[..., 2], obs[..., 0] + obs[..., 1])


def test_capacity_overflow_is_reported(Engine):
    w = make_workload(100, pH=[8.0], cap_wca=1)
    e = load(Engine(w.cfg), w, random_state(w, seed=1))
    with pytest.raises(_abi.CgpeError) as ei:
        e.md_forces()
    assert ei.value.status == _abi.ERR_CAPACITY


def test_invalid_calls_fail_loudly(Engine):
    w = make_workload(50, pH=[8.0])
    e = Engine(w.cfg)
    with pytest.raises(_abi.CgpeError) as ei:
        e.md_run(1)
    assert ei.value.status == _abi.ERR_STATE
    load(e, w, random_state(w, seed=2))
    with pytest.raises(_abi.CgpeError):
        e.mc_run(5, 10)
    with pytest.raises(_abi.CgpeError):
        e.set_replica_params(r_cut=1e6)
    w.cfg.abi_version = 1
    with pytest.raises(_abi.CgpeError):
        Engine(w.cfg)


def test_dihedral_near_straight_angle_matches_oracle_and_divergence_is_reported(Engine, oracle_mod):
    cfg = _abi.make_config(n_replicas=1, n_beads=4, n_types=1, box_l=40.0, time_step=1e-3, use_dihedral=1,
                           dihedral_k=6.5, dihedral_mult=3, dihedral_phase=np.pi)
    x = np.array([[-np.cos(1e-3), np.sin(1e-3), 0, 0], [0, 0, 0, 0], [1, 0, 0, 0],
                  [1.5, 0.6 * np.cos(0.7), 0.6 * np.sin(0.7), 0]], np.float32) + np.float32([5, 5, 5, 0])
    g, o = Engine(cfg), oracle_mod.OracleEngine(cfg)
    for e in (g, o):
        e.set_state(x, None, np.zeros(4, np.uint8))
    fo = o.forces_f64()[0]
    assert np.abs(fo).max() < 1.2e3
    for fg in (g.forces()[0], g.md_forces()[0]):
        assert np.abs(fg - fo).max() < 2e-4 * np.abs(fo).max()      # FP32 cross products of nearly parallel bonds
    # a replica whose velocities ran away is reported, not returned silently
    v = np.zeros((4, 4), np.float32)
    v[0, 0] = 5e3
    g.set_state(x, v, np.zeros(4, np.uint8))
    g.set_thermostat_seed(1)
    g.md_run(1)
    with pytest.raises(_abi.CgpeError) as ei:
        g.synchronize()
    assert ei.value.status == _abi.ERR_NUMERIC and "diverged" in str(ei.value)


@pytest.mark.parametrize("c_salt", [2e-2, 1e-3])   # 1 mM: partner rows of several hundred sites, accepted flips shared over the trial CTA's warps
def test_wide_engine_sample_loop_and_steepest_descent(Engine, oracle_mod, c_salt):
    """Replicas above 2048 particles run on the global-memory engine: same Markov chain."""
    w = make_workload(2100, pH=[8.0, 9.5], c_salt=c_salt)
    state = random_state(w, seed=77, jitter=0.02)
    g, o = _pair(Engine, oracle_mod, w, state)
    if c_salt < 1e-2:
        assert g.list_stats()[:, 2].min() / sum(w.n_sites) > 96   # mean row length: the helper path is the one that runs
    o.set_list_mode(True, 0.8)
    sp = g.sample_params(6, md_steps_per_burst=5, use_md=True, mc_blocks=((0, 301), (1, 11)), q_snapshot_every=4,
                         q_snapshot_rows=2, p_burst1=0.7, p_burst2=0.3)
    og, qg = g.sample_loop(sp)
    oo, qo = o.sample_loop(sp)
    np.testing.assert_array_equal(og[..., :3], oo[..., :3])
    np.testing.assert_allclose(og[..., 3:5], oo[..., 3:5], rtol=2e-4)
    np.testing.assert_allclose(og[..., 5], oo[..., 5], rtol=1e-9)
    np.testing.assert_array_equal(qg, qo)
    cg, co = g.counters(), o.counters()
    for k in ("md_steps", "trials", "accepted"):
        np.testing.assert_array_equal(cg[k], co[k])
    g2, o2 = _pair(Engine, oracle_mod, w, state)
    for e in (g2, o2):
        e.steepest_descent(10, f_max=0.0, gamma=1e-4, max_displacement=0.01)
    dx = np.abs(g2.get_state()[0][..., :3] - o2.state_f64()[0])
    # FP32 coordinates at |x| ~ 100 (ulp 7.6e-6), ten clamped steps; a bead next to a nearly straight
    # bond angle is more sensitive, hence the looser bound on the maximum
    assert np.median(dx) < 2e-5 and np.percentile(dx, 99.9) < 1e-4 and dx.max() < 2e-3
    s = g.list_stats()
    assert np.all(s[:, 0] > 0) and np.all(s[:, 1] <= s[:, 0]) and np.all(s[:, 3] <= s[:, 2])


# ---- explicit ions (the reference's particle model, M.py:173-198) ---------------
def _ion_state(w, seed, jitter=0.02):
    """Start state with explicit ions: chain jittered, velocities thermal, sites as built (all
    deprotonated, so N_B = N_N + N_N2 holds from the start)."""
    rng = np.random.default_rng(seed)
    R, P = w.R, w.xyzq.shape[0]
    xyzq = np.broadcast_to(w.xyzq, (R, P, 4)).copy()
    nc = w.cfg.n_chains * w.cfg.n_beads
    xyzq[:, :nc, :3] += rng.normal(scale=jitter, size=(R, nc, 3)).astype(np.float32)
    vel = np.zeros((R, P, 4), np.float32)
    vel[..., :3] = rng.normal(size=(R, P, 3))
    return xyzq.astype(np.float32), vel, np.broadcast_to(w.type, (R, P)).copy()


def _wrap_ions(x, w):
    x = np.array(x, dtype=np.float64)
    L, nc = w.cfg.box_l, w.cfg.n_chains * w.cfg.n_beads
    x[:, nc:] -= L * np.floor(x[:, nc:] / L)
    return x


def _ion_dx(xg, xo, w):
    """max |dx| with ions compared modulo the box (the engine folds free particles, the oracle does not)"""
    d = np.asarray(xg, np.float64) - xo
    L, nc = w.cfg.box_l, w.cfg.n_chains * w.cfg.n_beads
    d[:, nc:] -= L * np.rint(d[:, nc:] / L)
    return np.abs(d).max()


def test_explicit_ions_energy_forces_md(Engine, oracle_mod):
    w = make_workload(100, pH=[7.5, 9.0], explicit_ions=True)
    assert w.cfg.n_ion_slots == 2 * 34 + 2 * 200 + 34
    state = _ion_state(w, seed=3)
    g, o = _pair(Engine, oracle_mod, w, state)
    eg, eo = g.energy(), o.energy()
    scale = np.maximum(np.abs(eo), 1e-3 * np.abs(eo).max(axis=1, keepdims=True))
    assert np.max(np.abs(eg - eo) / scale) < REL_TOL, (eg, eo)
    fo = o.forces_f64()
    for fg in (g.forces(), g.md_forces()):
        assert np.abs(fg - fo).max() < REL_TOL * np.abs(fo).max()
    np.testing.assert_allclose(g.min_dist(), o.min_dist(), rtol=1e-6)
    for e in (g, o):
        e.md_run(12)
    xg, vg, tg = g.get_state()
    xo, vo, _ = o.state_f64()
    assert _ion_dx(xg[..., :3], xo, w) < 5e-5
    hidden = tg >= w.cfg.n_types
    assert np.all(vg[hidden][:, :3] == 0) and np.abs(vg[..., :3] - vo).max() < 5e-3


def test_explicit_ions_trial_moves_bit_exact(Engine, oracle_mod):
    """Insertion / deletion of H+ with exclusion range: same sites, ions, decisions and final state."""
    w = make_workload(100, pH=[7.0, 8.2, 9.5], explicit_ions=True)
    state = _ion_state(w, seed=5)
    g, o = _pair(Engine, oracle_mod, w, state)
    n_excl = 0
    for rx, n in ((0, 3000), (1, 400), (0, 1500)):
        tg = g.mc_run(rx, n, trace=True)
        to, ties = o.mc_run_follow(rx, tg, tie_tol=1e-4)
        assert ties <= 2
        for f in ("particle", "ion_slot", "direction", "accepted", "excluded", "u"):
            np.testing.assert_array_equal(tg[f], to[f], err_msg=f)
        tried = (to["particle"] >= 0) & (to["excluded"] == 0)
        err = np.abs(tg["delta_e"][tried] - to["delta_e"][tried]) / np.maximum(np.abs(to["delta_e"][tried]), 1.0)
        assert err.max() < REL_TOL
        n_excl += int(to["excluded"].sum())
        assert tg["accepted"].sum() > 0
    assert n_excl > 0
    xg, vg, tyg = g.get_state()
    xo, vo, qo = o.state_f64()
    _, _, tyo = o.get_state()
    np.testing.assert_array_equal(tyg, tyo)
    np.testing.assert_array_equal(xg[..., 3], qo.astype(np.float32))
    assert _ion_dx(xg[..., :3], xo, w) < 1e-5                    # inserted ions sit where the oracle put them
    assert np.abs(vg[..., :3] - vo).max() < 1e-5                  # with the same Maxwell-Boltzmann velocity
    cg, co = g.observables()[2], o.observables()[2]
    np.testing.assert_array_equal(cg, co)
    np.testing.assert_array_equal(cg[:, 2], cg[:, 0] + cg[:, 1])  # N_B = N_N + N_N2 (NB:253-254)


def test_explicit_ions_sample_loop(Engine, oracle_mod):
    w = make_workload(60, pH=[8.0, 9.5], explicit_ions=True)
    state = _ion_state(w, seed=9)
    g, o = _pair(Engine, oracle_mod, w, state)
    sp = g.sample_params(16, md_steps_per_burst=5, use_md=True, mc_blocks=((0, 61), (1, 11)), p_burst1=0.7, p_burst2=0.3)
    og, _ = g.sample_loop(sp)
    oo, _ = o.sample_loop(sp)
    np.testing.assert_array_equal(og[..., :3], oo[..., :3])
    np.testing.assert_allclose(og[..., 3:5], oo[..., 3:5], rtol=2e-4)
    np.testing.assert_array_equal(og[..., 2], og[..., 0] + og[..., 1])
    assert len(np.unique(og[..., 2])) > 1


def test_explicit_ion_configs_that_cannot_be_served_are_refused(Engine):
    w = make_workload(100, pH=[8.0], explicit_ions=True)
    w.cfg.reactions[0].exclusion_range = 0.1           # below the H+ WCA cutoff
    with pytest.raises(_abi.CgpeError) as ei:
        Engine(w.cfg)
    assert ei.value.status == _abi.ERR_UNSUPPORTED


def _binning_sem(x, n_bins=10):
    n = (x.shape[-1] // n_bins) * n_bins
    means = x[..., :n].reshape(x.shape[:-1] + (n_bins, -1)).mean(axis=-1)
    return means.std(axis=-1, ddof=1) / np.sqrt(n_bins)


def test_ensemble_observables_agree_within_error_bars(Engine, oracle_mod):
    """Third correctness criterion of the north star: alpha(pH), <Rg>, <Ree> of the interacting chain
    from the CUDA path and from the FP64 oracle, run with DIFFERENT random streams (independent Markov
    chains), agree within their binning-analysis error bars."""
    pH = np.array([6.5, 7.5, 8.2, 9.0, 10.5])
    oracle_mod.set_threads(os.cpu_count() or 1)
    results = {}
    # a short chain (Rouse time ~ 20 time units) and 0.25 time units of MD per sample: the 1600 samples
    # span ~ 400 time units, so Rg / Ree decorrelate many times and the binning error bars are honest
    for name, seed_shift in (("gpu", 0), ("oracle", 1000)):
        # no dihedral term here: its 13 kT barriers (k = 6.57 kT) freeze the torsions over any affordable
        # run, so <Rg> would depend on the initial torsion sequence of each independent chain
        w = make_workload(24, pH=pH, replica_offset=seed_shift, flags={"USE_DIHEDRAL_POT": False})
        e = Engine(w.cfg) if name == "gpu" else oracle_mod.OracleEngine(w.cfg)
        w.init_engine(e)
        n1, n2 = w.n_sites
        kw = dict(md_steps_per_burst=250, use_md=True, mc_blocks=((0, 5 * n1 + 1), (1, 5 * n2 + 1)), p_burst1=1.0, p_burst2=0.0)
        e.sample_loop(e.sample_params(200, **kw))                      # equilibration
        obs, _ = e.sample_loop(e.sample_params(1600, **kw))
        results[name] = (obs, n1)
    oracle_mod.set_threads(1)
    (og, n1), (oo, _) = results["gpu"], results["oracle"]
    for col, scale, label in ((0, n1, "alpha"), (3, 1.0, "Rg"), (4, 1.0, "Ree")):
        mg, mo = og[..., col].mean(axis=1) / scale, oo[..., col].mean(axis=1) / scale
        sg, so = _binning_sem(og[..., col]) / scale, _binning_sem(oo[..., col]) / scale
        sigma = np.sqrt(sg ** 2 + so ** 2) + 2e-3 * np.abs(mo)
        z = np.abs(mg - mo) / sigma
        assert np.all(z < 4.5), (label, mg, mo, sigma, z)
    # and the physics is the expected one: titration curve monotone, charged chain more extended
    alpha = og[..., 0].mean(axis=1) / n1
    assert np.all(np.diff(alpha) > 0) and alpha[0] < 0.4 and alpha[-1] > 0.97


def test_explicit_ions_on_the_wide_engine(Engine, oracle_mod):
    """More than 2048 particles per replica (chain + ions): the global-memory engine serves the same
    explicit-ion model.  N = 400 -> 400 + 2*134 + 2*800 + 134 spare = 2402 particles."""
    w = make_workload(400, pH=[7.5, 9.0], explicit_ions=True)
    assert w.xyzq.shape[0] == 400 + 3 * sum(w.n_sites) + 2 * w.ru.n_salt > 2048
    state = _ion_state(w, seed=13)
    g, o = _pair(Engine, oracle_mod, w, state)
    eg, eo = g.energy(), o.energy()
    scale = np.maximum(np.abs(eo), 1e-3 * np.abs(eo).max(axis=1, keepdims=True))
    assert np.max(np.abs(eg - eo) / scale) < REL_TOL
    fo = o.forces_f64()
    assert np.abs(g.md_forces() - fo).max() < REL_TOL * np.abs(fo).max()
    for rx, n in ((0, 2500), (1, 200)):
        tg = g.mc_run(rx, n, trace=True)
        to, ties = o.mc_run_follow(rx, tg, tie_tol=1e-4)
        assert ties <= 2
        for f in ("particle", "ion_slot", "direction", "accepted", "excluded", "u"):
            np.testing.assert_array_equal(tg[f], to[f], err_msg=f)
        tried = (to["particle"] >= 0) & (to["excluded"] == 0)
        err = np.abs(tg["delta_e"][tried] - to["delta_e"][tried]) / np.maximum(np.abs(to["delta_e"][tried]), 1.0)
        assert err.max() < REL_TOL and tg["accepted"].sum() > 0
    o.set_list_mode(True, 0.8)
    for e in (g, o):
        e.md_run(8)
    xg, vg, tyg = g.get_state()
    xo, vo, qo = o.state_f64()
    np.testing.assert_array_equal(tyg, o.get_state()[2])
    assert _ion_dx(xg[..., :3], xo, w) < 1e-4 and np.abs(vg[..., :3] - vo).max() < 5e-3
    sp = g.sample_params(5, md_steps_per_burst=4, use_md=True, mc_blocks=((0, 201), (1, 11)), p_burst1=0.7, p_burst2=0.3)
    og, _ = g.sample_loop(sp)
    oo, _ = o.sample_loop(sp)
    np.testing.assert_array_equal(og[..., :3], oo[..., :3])
    np.testing.assert_array_equal(og[..., 2], og[..., 0] + og[..., 1])
    np.testing.assert_allclose(og[..., 3:5], oo[..., 3:5], rtol=2e-4)


@pytest.mark.parametrize("n_poly,n_mon,ions", [
    (3, 40, False),     # shared-memory engine
    (3, 40, True),
    (6, 400, False),    # 2400 beads: global-memory engine, cell grids over several chains
    (4, 200, True),     # 800 beads + ions > 2048 particles
])
def test_several_chains_per_box(Engine, oracle_mod, n_poly, n_mon, ions):
    """"Number of polymers" > 1 (P.py:384, M.py:116-170): chains share the box, interact through WCA and
    Debye-Hueckel across chains, and every chain carries its own titratable sites; Rg / Ree are those of
    the first chain (E.py:100-101 analyses chain_start=0, number_of_chains=1)."""
    w = make_workload(n_mon, pH=[7.5, 9.0], explicit_ions=ions, n_poly=n_poly)
    nc = n_poly * n_mon
    assert w.cfg.n_chains == n_poly and w.xyzq.shape[0] == nc + w.cfg.n_ion_slots
    assert (w.xyzq.shape[0] > 2048) == (n_mon >= 200)
    state = _ion_state(w, seed=5) if ions else random_state(w, seed=5, jitter=0.03)
    g, o = _pair(Engine, oracle_mod, w, state)
    eg, eo = g.energy(), o.energy()
    scale = np.maximum(np.abs(eo), 1e-3 * np.abs(eo).max(axis=1, keepdims=True))
    assert np.max(np.abs(eg - eo) / scale) < REL_TOL, (eg, eo)
    fo = o.forces_f64()
    for name, fg in (("all-pairs", g.forces()), ("md-lists", g.md_forces())):
        assert np.abs(fg - fo).max() < REL_TOL * np.abs(fo).max(), name
    rg_g, ree_g, cnt_g = g.observables()
    rg_o, ree_o, cnt_o = o.observables()
    np.testing.assert_allclose(rg_g, rg_o, rtol=1e-6)
    np.testing.assert_allclose(ree_g, ree_o, rtol=1e-6)
    np.testing.assert_array_equal(cnt_g, cnt_o)
    for rx, n in ((0, 1500), (1, 300)):
        tg = g.mc_run(rx, n, trace=True)
        to, ties = o.mc_run_follow(rx, tg, tie_tol=1e-4)
        assert ties <= 2
        for f in ("particle", "ion_slot", "direction", "accepted", "excluded", "u"):
            np.testing.assert_array_equal(tg[f], to[f], err_msg=f)
        tried = (to["particle"] >= 0) & (to["excluded"] == 0)
        err = np.abs(tg["delta_e"][tried] - to["delta_e"][tried]) / np.maximum(np.abs(to["delta_e"][tried]), 1.0)
        assert err.max() < REL_TOL and tg["accepted"].sum() > 0
        assert len(np.unique(tg["particle"][tg["particle"] >= 0] // n_mon)) == n_poly   # sites of every chain are drawn
    o.set_list_mode(True, 0.8)
    for e in (g, o):
        e.md_run(8)
    xg, vg, tyg = g.get_state()
    xo, vo, qo = o.state_f64()
    np.testing.assert_array_equal(tyg, o.get_state()[2])
    assert _ion_dx(xg[..., :3], xo, w) < 1e-4 and np.abs(vg[..., :3] - vo).max() < 5e-3


@pytest.mark.parametrize("c_salt,build", [(2e-2, "one CTA"), (1e-3, "one CTA"), (1e-3, "cell walk"), (2e-2, "many CTAs"), (1e-3, "many CTAs")])
def test_cell_grid_rows_hold_every_pair_for_any_table_size(Engine, c_salt, build, monkeypatch):
    """The list build of the global-memory engine folds the cells of the box onto a table of A^3 buckets.
    Whatever A is (1 = a single bucket, 3 = every bucket is its own neighbour's neighbour, sizes that
    trigger the wrap adjustment nf mod A in {1, 2}, the default), the forces through the rows must equal
    the all-pairs forces, and the rows must hold the same pairs."""
    # builds: the default (at 1 mM the Debye-Hueckel rows come from the shared-memory staged brute force, "tile mode"),
    # the cell walk forced for long rows too (replicas whose chargeable set exceeds shared memory), the multi-CTA cell sort
    # of the 10^5..10^6-particle boxes forced at this size
    if build == "cell walk":
        monkeypatch.setenv("CGPE_NO_DH_TILE", "1")
    if build == "many CTAs":
        monkeypatch.setenv("CGPE_CELLS_MULTI_MIN", "1")
        monkeypatch.setenv("CGPE_NO_DH_TILE", "1")
    w = make_workload(2500, pH=[6.0, 9.5], c_salt=c_salt)
    state = random_state(w, seed=77, jitter=0.03)
    ref_forces, ref_stats = None, None
    for nc_max in (None, 1, 3, 4, 5, 6, 7, 9, 13):
        if nc_max is None:
            monkeypatch.delenv("CGPE_GRID_NC_MAX", raising=False)
        else:
            monkeypatch.setenv("CGPE_GRID_NC_MAX", str(nc_max))
        g = load(Engine(w.cfg), w, state)
        f_rows, stats = g.md_forces(), g.list_stats()
        if ref_forces is None:
            ref_forces, ref_stats = g.forces(), stats
            assert stats[:, 0].min() > 0 and stats[:, 2].min() > 0
        assert np.abs(f_rows - ref_forces).max() < REL_TOL * np.abs(ref_forces).max(), nc_max
        np.testing.assert_array_equal(stats, ref_stats, err_msg=f"nc_max={nc_max}")


@pytest.mark.parametrize("variant", ["mirrors", "all-particle sums", "cell grids"])
def test_ion_moves_of_the_global_memory_engine_agree_in_every_variant(Engine, oracle_mod, variant, monkeypatch):
    """The trial loop of the global-memory engine has three ways to get the sums of an ion insertion /
    deletion: shared-memory mirrors (replicas up to ~7000 particles), all-particle sums from global memory,
    and the cell grids in chunks of 256 trials with a journal of the ions inserted since the grids were built
    (large boxes).  All three must reproduce the oracle's decisions, also over several blocks in a row so that
    slots are deleted and re-used."""
    if variant != "mirrors":
        monkeypatch.setenv("CGPE_MC_STAGE_LIMIT", "0")
        monkeypatch.setenv("CGPE_MC_GRID", "1" if variant == "cell grids" else "0")
    w = make_workload(400, pH=[7.0, 8.5, 10.0], explicit_ions=True)
    state = _ion_state(w, seed=21)
    g, o = _pair(Engine, oracle_mod, w, state)
    n_acc = 0
    for rx, n in ((0, 1500), (1, 150), (0, 1100), (1, 90)):
        tg = g.mc_run(rx, n, trace=True)
        to, ties = o.mc_run_follow(rx, tg, tie_tol=1e-4)
        assert ties <= 2
        for f in ("particle", "ion_slot", "direction", "accepted", "excluded", "u"):
            np.testing.assert_array_equal(tg[f], to[f], err_msg=f"{variant}: {f}")
        tried = (to["particle"] >= 0) & (to["excluded"] == 0)
        err = np.abs(tg["delta_e"][tried] - to["delta_e"][tried]) / np.maximum(np.abs(to["delta_e"][tried]), 1.0)
        assert err.max() < REL_TOL, (variant, err.max())
        n_acc += int(tg["accepted"].sum())
        # move the ions: the next block sees new positions and fresh rows.  The oracle then takes the device's
        # coordinates (FP32 and FP64 trajectories part by ~1e-6, which Delta-E would see at the 1e-5 level).
        g.md_run(3); o.md_run(3)
        o.set_positions_f64(np.asarray(g.get_state()[0][..., :3], np.float64))
    assert n_acc > 300
    xg, vg, tyg = g.get_state()
    np.testing.assert_array_equal(tyg, o.get_state()[2])
    np.testing.assert_array_equal(g.observables()[2], o.observables()[2])


def test_compact_partner_lists_follow_charges_and_rebuilds(Engine, oracle_mod):
    """The MD loop of the shared-memory engine walks per-thread lists of the listed AND charged partners,
    written out at the start of a burst and after every list build.  Interleave bursts long enough to rebuild
    the lists with reaction blocks that change which sites are charged, at weak screening (long rows) and with
    explicit ions (every ion is a partner), and compare with the oracle step by step."""
    # n_mon = 300: 100 chargeable particles -- the one-CTA engine keeps the bit walk below 97 (chain_engine.cu,
    # build_dh_compact); the cluster engine always takes the compact lists
    for n_mon, ions, c_salt in ((300, False, 1e-3), (60, True, 2e-2)):
        w = make_workload(n_mon, pH=[6.0, 8.5, 11.0], c_salt=c_salt, explicit_ions=ions)
        state = _ion_state(w, seed=9) if ions else random_state(w, seed=9, jitter=0.02)
        g, o = _pair(Engine, oracle_mod, w, state)
        o.set_list_mode(True, 0.8)
        builds0 = g.engine_info()["compact_builds"].copy()
        for burst in range(4):
            g.md_run(60); o.md_run(60)
            xg = g.get_state()[0]
            assert _ion_dx(xg[..., :3], o.state_f64()[0], w) < 2e-3, (n_mon, burst)
            o.set_positions_f64(np.asarray(xg[..., :3], np.float64))     # keep the trajectories together
            for rx, n in ((0, 300), (1, 40)):
                tg = g.mc_run(rx, n, trace=True)
                to, ties = o.mc_run_follow(rx, tg, tie_tol=1e-4)
                np.testing.assert_array_equal(tg["accepted"], to["accepted"])
            fo = o.forces_f64()
            assert np.abs(g.md_forces() - fo).max() < REL_TOL * np.abs(fo).max()
        assert g.counters()["list_rebuilds"].min() >= 4
        # the compact partner lists were really the path taken: written at every burst start, list build and md_forces call
        assert (g.engine_info()["compact_builds"] - builds0).min() >= 8, g.engine_info()


@pytest.mark.parametrize("wrap", [False, True])
def test_md_of_the_headline_regime_matches_oracle_across_list_rebuilds(Engine, oracle_mod, wrap):
    """The code bench.py times: N = 1000 at 1 mM (r_cut 27.6: Debye-Hueckel rows of ~150 partners per site, compact
    partner lists), pH 2 / 4 / 8, i.e. every site charged down to a third of them.  Bursts of 40 steps from a hot
    start so that at least one Verlet-list rebuild falls inside; the oracle is re-seated on the device positions
    between bursts (the gap of a 40-step burst is what is compared).  wrap = True shrinks the box until the chain
    reaches through the periodic boundary, which switches the kernels to their minimum-image instantiation."""
    w = make_workload(1000, pH=[2.0, 4.0, 8.0], c_salt=1e-3, skin=0.2)   # thin skin: a rebuild every ~40 thermal steps
    if wrap:
        ext = float(np.ptp(w.xyzq[:, :3], axis=0).max())
        w.cfg.box_l = max(2.0 * (float(w.r_cut.max()) + 0.2) + 1.0, 0.9 * ext)   # r_cut + skin <= L/2, extent + list radius > L
    state = random_state(w, seed=41, jitter=0.02)
    xyzq, vel, types = state
    tix, ru = w.ru.type_index, w.ru
    for r, frac in enumerate((1.0, 1.0, 0.35)):                     # protonation as the titration has it at 1 mM
        sites = np.isin(types[r], [tix["N"], tix["NH"]])
        ends = np.isin(types[r], [tix["N2"], tix["NH2"]])
        rng = np.random.default_rng(100 + r)
        prot = rng.random(types.shape[1]) < frac
        types[r, sites] = np.where(prot[sites], tix["NH"], tix["N"])
        types[r, ends] = np.where(prot[ends], tix["NH2"], tix["N2"])
    q_of = np.zeros(len(tix) + 1, np.float32)
    for name, t in tix.items():
        q_of[t] = ru.charge[name]
    xyzq[..., 3] = q_of[types]
    g, o = _pair(Engine, oracle_mod, w, (xyzq, vel, types))
    o.set_list_mode(True, 0.2)
    ulp = float(np.spacing(np.float32(np.abs(xyzq[..., :3]).max())))
    builds0 = g.engine_info()["compact_builds"].copy()
    in_burst = np.zeros(w.R, np.int64)
    for burst in range(3):
        r0 = g.counters()["list_rebuilds"].copy()
        g.md_run(60); o.md_run(60)
        in_burst += g.counters()["list_rebuilds"] - r0 - 1          # one build opens the call, the rest fell inside the burst
        xg, vg, _ = g.get_state()
        xo, vo, _ = o.state_f64()
        dx, dv = np.abs(xg[..., :3] - xo).max(), np.abs(vg[..., :3] - vo).max()
        assert dx < 600 * ulp and dv < 0.5, f"burst {burst}: dx={dx:.3g} ({dx / ulp:.0f} ulp) dv={dv:.3g}"
        o.set_positions_f64(np.asarray(xg[..., :3], np.float64))
        fo = o.forces_f64()
        assert np.abs(g.md_forces() - fo).max() < REL_TOL * np.abs(fo).max()
    assert in_burst.min() >= 1, in_burst                           # every replica rebuilt its lists inside a burst
    assert (g.engine_info()["compact_builds"] - builds0).min() >= 6 + in_burst.min(), g.engine_info()
    stats = g.list_stats()
    assert stats[0, 2] / 334 > 100                 # the fully charged replica really has the long rows (directed entries per site)
    np.testing.assert_array_equal(g.counters()["md_steps"], o.counters()["md_steps"])


@pytest.mark.parametrize("explicit_ions", [False, True])
def test_config_c1_as_specified(Engine, oracle_mod, explicit_ions):
    """BASELINE configs[0] exactly as SURVEY 8d states it: N = 50 (16 N + 1 N2 sites), Debye-Hueckel at 10 mM
    (kappa 0.11452, r_cut 8.732), pH = pK1 = 8.18, one replica, implicit ions and the reference's explicit ions."""
    w = make_workload(50, pH=[8.18], c_salt=1e-2, explicit_ions=explicit_ions)
    assert w.n_sites == (16, 1) and abs(w.kappa[0] - 0.11452) < 2e-5 and abs(w.r_cut[0] - 8.732) < 2e-3
    state = _ion_state(w, seed=5) if explicit_ions else random_state(w, seed=5, jitter=0.03)
    g, o = _pair(Engine, oracle_mod, w, state)
    eg, eo = g.energy(), o.energy()
    scale = np.maximum(np.abs(eo), 1e-3 * np.abs(eo).max(axis=1, keepdims=True))
    assert np.max(np.abs(eg - eo) / scale) < REL_TOL, (eg, eo)
    fo = o.forces_f64()
    for fg in (g.forces(), g.md_forces()):
        assert np.abs(fg - fo).max() < REL_TOL * np.abs(fo).max()
    for rx, n in ((0, 2000), (1, 400)):
        tg = g.mc_run(rx, n, trace=True)
        to, ties = o.mc_run_follow(rx, tg, tie_tol=1e-4)
        assert ties <= 2
        for col in ("particle", "direction", "accepted", "u"):
            np.testing.assert_array_equal(tg[col], to[col])
        tried = to["particle"] >= 0
        err = np.abs(tg["delta_e"][tried] - to["delta_e"][tried]) / np.maximum(np.abs(to["delta_e"][tried]), 1.0)
        assert err.max() < REL_TOL
        if rx == 0:
            assert 0.05 < tg["accepted"].mean() < 0.95      # pH = pK1: both directions are taken
    np.testing.assert_array_equal(g.get_state()[2], o.get_state()[2])
    g.md_run(20); o.md_run(20)
    assert _ion_dx(g.get_state()[0][..., :3], o.state_f64()[0], w) < 2e-4
    sp = g.sample_params(6, md_steps_per_burst=20, use_md=True, mc_blocks=((0, 81), (1, 6)), p_burst1=0.7, p_burst2=0.05)
    og, _ = g.sample_loop(sp)
    oo, _ = o.sample_loop(sp)
    np.testing.assert_array_equal(og[..., :3], oo[..., :3])
    np.testing.assert_allclose(og[..., 3:5], oo[..., 3:5], rtol=2e-3)


@pytest.mark.parametrize("charge,n_mon", [(-20.0, 100), (20.0, 100), (-20.0, 1000)])
def test_charged_nanoparticle_config_c4(Engine, oracle_mod, charge, n_mon):
    """BASELINE config 4: the chain next to one fixed charged sphere (type "NP", sigma = 10 sim_length,
    fixed_type_mask) in a spectator slot, implicit-ion reactions.  Energies, forces (all-pairs and
    list path), trial moves, MD steps and the fused loop against the oracle; the sphere never moves."""
    w = make_workload(n_mon, pH=[6.0, 8.18, 10.0], nanoparticle=dict(charge=charge, clearance=0.9))
    state = random_state(w, seed=n_mon + int(charge), jitter=0.03, protonated_fraction=0.6)
    g, o = _pair(Engine, oracle_mod, w, state)
    i_np = n_mon
    x_np = g.get_state()[0][:, i_np, :3].copy()
    d = np.linalg.norm(state[0][0, :n_mon, :3] - state[0][0, i_np, :3], axis=1)
    assert (d < w.r_cut[0]).sum() >= 3
    eg, eo = g.energy(), o.energy()
    scale = np.maximum(np.abs(eo), 1e-3 * np.abs(eo).max(axis=1, keepdims=True))
    assert np.max(np.abs(eg - eo) / scale) < REL_TOL, (eg, eo)
    fo = o.forces_f64()
    fmax = np.abs(fo).max()
    for name, fg in (("all-pairs", g.forces()), ("md-lists", g.md_forces())):
        assert np.abs(fg[..., :3] - fo).max() / fmax < REL_TOL, name
        assert np.all(fg[:, i_np, :3] == 0.0), name
    # trial moves: same stream, same decisions, Delta-E with the sphere's potential in it
    for rx in (0, 1, 0):
        tg = g.mc_run(rx, 3000, trace=True)
        to, ties = o.mc_run_follow(rx, tg, tie_tol=1e-4)
        assert ties <= 4
        for k in ("particle", "direction", "accepted", "u"):
            np.testing.assert_array_equal(tg[k], to[k])
        tried = to["particle"] >= 0
        err = np.abs(tg["delta_e"][tried] - to["delta_e"][tried]) / np.maximum(np.abs(to["delta_e"][tried]), 1.0)
        assert err.max() < REL_TOL
        assert tg["accepted"].sum() > 0
    np.testing.assert_array_equal(g.get_state()[2], o.get_state()[2])
    # MD steps (the state now carries the flipped charges)
    ulp = float(np.spacing(np.float32(np.abs(state[0][..., :3]).max())))
    for n_steps, tol in ((1, 0.75 * ulp), (9, 8 * ulp)):
        g.md_run(n_steps); o.md_run(n_steps)
        xg, vg, _ = g.get_state()
        xo, vo, _ = o.state_f64()
        assert np.abs(xg[..., :3] - xo).max() < tol and np.abs(vg[..., :3] - vo).max() < tol * 1e3
    # the fused loop, with steepest descent in between
    g.steepest_descent(3, 0.0, 0.01, 0.01); o.steepest_descent(3, 0.0, 0.01, 0.01)
    sp = g.sample_params(20, md_steps_per_burst=5, use_md=True, mc_blocks=((0, 31), (1, 11)),
                         q_snapshot_every=0, q_snapshot_rows=0, schedule_seed=10, p_burst1=0.7, p_burst2=0.3)
    og, _ = g.sample_loop(sp)
    oo, _ = o.sample_loop(sp)
    np.testing.assert_array_equal(og[..., :3], oo[..., :3])
    np.testing.assert_allclose(og[..., 3:5], oo[..., 3:5], rtol=2e-4)
    np.testing.assert_array_equal(og[..., 2], og[..., 0] + og[..., 1])      # implicit ions: N_B = N_N + N_N2
    xg, vg, _ = g.get_state()
    np.testing.assert_array_equal(xg[:, i_np, :3], x_np)
    assert np.all(vg[:, i_np, :3] == 0.0)


def test_spectator_particles_beyond_the_shared_memory_engine_are_refused(Engine):
    w = make_workload(2500, pH=[8.0], nanoparticle=dict(charge=-20.0))
    e = Engine(w.cfg)
    with pytest.raises(Exception) as ei:
        e.set_state(w.xyzq, None, w.type)
    assert "spectator" in str(ei.value)


def test_full_size_shard_properties(Engine):
    """BASELINE config 3 at full size -- one shard, 128 pH replicas x N=1000 at 1 mM, the bench workload -- is
    beyond what the oracle finishes in seconds, so it is held to size-independent properties: Newton's third
    law, list path == all-pairs path, translation invariance, energy conservation without thermostat, and the
    integer bookkeeping identities of the fused loop."""
    from coarse_grain_polyelectrolyte_b200 import workloads
    w = workloads.config_c3_shard(0, 8)
    assert w.R == 128 and w.xyzq.shape[0] == 1000
    e = w.init_engine(Engine(w.cfg))
    e.md_run(400)
    e.sample_loop(w.sample_params(e, 2))            # off the synthetic start; low-pH replicas are charged now
    # forces: sum to zero per replica, and the Verlet-list path agrees with the all-pairs kernel
    f_md, f_all = e.md_forces().astype(np.float64), e.forces().astype(np.float64)
    for f in (f_md, f_all):
        assert np.all(np.abs(f.sum(axis=1)).max(axis=1) < 1e-5 * np.abs(f).sum(axis=1).max(axis=1))
    assert np.abs(f_md - f_all).max() < 1e-5 * np.abs(f_all).max()
    # rigid translation (the chains cross no box face at 1 mM): energies unchanged
    xyzq, vel, typ = e.get_state()
    e0 = e.energy()
    moved = xyzq.copy(); moved[..., :3] += np.float32([3.25, -1.5, 0.75])
    e.set_state(moved, vel, typ); e.set_replica_params(pH=w.pH, kappa=w.kappa, r_cut=w.r_cut)
    e1 = e.energy()
    np.testing.assert_allclose(e1[:, [0, 2, 3, 4]], e0[:, [0, 2, 3, 4]], rtol=2e-4, atol=1e-3)
    # charges follow types on every particle
    tix, ru = w.ru.type_index, w.ru
    q_of = np.zeros(len(tix) + 1, np.float32)
    for name, t in tix.items():
        q_of[t] = ru.charge[name]
    np.testing.assert_array_equal(xyzq[..., 3], q_of[typ])
    assert (xyzq[:32, :, 3] != 0).sum() > 32 * 100          # the acidic third of the replicas is well charged
    # the fused loop: bookkeeping identities at full size
    e.set_state(xyzq, vel, typ); e.set_replica_params(pH=w.pH, kappa=w.kappa, r_cut=w.r_cut)
    c0 = e.counters()
    n_samples = 6
    obs, _ = e.sample_loop(w.sample_params(e, n_samples))
    c1 = e.counters()
    n1, n2 = w.n_sites
    assert np.all(np.isfinite(obs)) and np.all(obs[..., 3] > 0) and np.all(obs[..., 4] > 0)
    np.testing.assert_array_equal(obs[..., 2], obs[..., 0] + obs[..., 1])            # N_B = N_N + N_N2 (NB:253-254)
    assert np.all((obs[..., 0] >= 0) & (obs[..., 0] <= n1) & (obs[..., 1] >= 0) & (obs[..., 1] <= n2))
    typ1 = e.get_state()[2]
    np.testing.assert_array_equal((typ1 == tix["N"]).sum(axis=1), obs[:, -1, 0].astype(np.int64))
    np.testing.assert_array_equal((typ1 == tix["N2"]).sum(axis=1), obs[:, -1, 1].astype(np.int64))
    np.testing.assert_array_equal((c1["trials"] - c0["trials"]).sum(axis=1), np.full(128, n_samples * (5 * n1 + 1 + 5 * n2 + 1)))
    assert np.all((c1["md_steps"] - c0["md_steps"]) % 500 == 0)                       # whole bursts (P.py:369)
    # deprotonation rises with pH across the shard
    alpha = obs[..., 0].mean(axis=1) / n1
    assert alpha[-8:].mean() > 0.95 and alpha[:8].mean() < alpha[60:68].mean() < alpha[-8:].mean()
    # no thermostat, no noise: the total energy is conserved.  Without the torsion term: its gradient is
    # regularised near straight bond angles, which a chain fresh from the synthetic start
    # visits, and is not conservative there -- the FP64 oracle drifts by the same 0.5 % with it.  What is left is
    # the jump of the unshifted Debye-Hueckel potential when a charged pair crosses r_cut (M.py:97-107).
    import copy
    params = copy.deepcopy(w.params)
    params["Simulation configuration"]["USE_DIHEDRAL_POT"] = False
    w2 = workloads.Workload(params, pH=w.pH, c_salt=1e-3)
    e2 = Engine(w2.cfg)
    e2.set_state(*e.get_state()); e2.set_replica_params(pH=w2.pH, kappa=w2.kappa, r_cut=w2.r_cut, kT=0.0, gamma=0.0)
    e2.md_run(200)
    ea = e2.energy()[:, 0]
    e2.md_run(1500)
    eb = e2.energy()[:, 0]
    assert np.all(np.abs(eb - ea) < 3e-3 * np.abs(ea)), np.abs((eb - ea) / ea).max()


@pytest.mark.parametrize("protonated_fraction", [1.0, 0.7, 0.3])
def test_both_shared_memory_engines_follow_the_same_trajectory(protonated_fraction):
    """State in HBM is engine-agnostic and the two shared-memory engines split every Debye-Hueckel partner list the same way
    (packed work items of the charged sites, same sub-threads per site), so an MD burst with list rebuilds followed by trial
    moves and another burst ends in the same charge state and, to a few ulp of the FP32 velocities (one force evaluation in
    a few hundred rounds differently: measured 0 ulp on the positions, <= 4 ulp on a velocity), in the same coordinates --
    which is what lets the library pick the engine per launch (replicas per GPU, salt) without changing the physics between
    a 1-GPU and an 8-GPU run."""
    from coarse_grain_polyelectrolyte_b200.engine import Engine as E
    w = make_workload(1000, pH=[2.0, 5.0, 8.0], c_salt=1e-3)
    state = random_state(w, seed=5, protonated_fraction=protonated_fraction, jitter=0.02)
    out = {}
    for name in ("solo", "duo"):
        e = load(E(w.cfg), w, state)
        e.set_engine(name)
        e.md_run(150)
        e.mc_run(0, 500)
        e.md_run(60)
        assert e.engine_info()["engine"] == name
        out[name] = e.get_state() + (e.counters()["accepted"],)
    np.testing.assert_array_equal(out["solo"][2], out["duo"][2])            # types
    np.testing.assert_array_equal(out["solo"][3], out["duo"][3])            # accepted trial moves
    np.testing.assert_array_equal(out["solo"][0][..., 3], out["duo"][0][..., 3])   # charges
    assert np.abs(out["solo"][0][..., :3] - out["duo"][0][..., :3]).max() < 1e-5
    assert np.abs(out["solo"][1] - out["duo"][1]).max() < 1e-5


def test_replicas_near_the_shared_memory_limit_get_an_engine_with_room_for_their_rows():
    """At ~2000 particles the one-CTA engine's fixed arrays leave 5 WCA row entries per particle: rows overflow, pairs are
    dropped, the replica blows up.  Left to itself the library serves such a replica with the cluster engine (each CTA holds
    the rows of half the particles); forced onto the one-CTA engine it reports the overflow, not the blow-up that follows."""
    from coarse_grain_polyelectrolyte_b200 import _abi
    from coarse_grain_polyelectrolyte_b200._abi import CgpeError
    from coarse_grain_polyelectrolyte_b200.engine import Engine as E
    w = make_workload(2000, pH=[2.0, 7.0], c_salt=1e-3)
    e = E(w.cfg)
    w.init_engine(e)
    assert e.engine_info()["engine"] == "duo"
    e.md_run(600)
    e.mc_run(0, 1000)
    e.md_run(150)
    x, v, _ = e.get_state()
    assert np.isfinite(x).all() and np.abs(v).max() < 20.0
    forced = E(w.cfg)
    w.init_engine(forced)
    forced.set_engine("solo")
    with pytest.raises(CgpeError) as ei:
        forced.md_run(2000)
        forced.get_state()
    assert ei.value.status == _abi.ERR_CAPACITY and "WCA" in str(ei.value)


def test_wave_order_of_the_one_cta_engine_does_not_change_results(monkeypatch):
    """With more replicas than SMs the fused sampling launch of the one-CTA engine starts the most expensive replicas first
    (predicted from the burst schedule, the charged pairs and the acceptance).  Replicas are independent: same results."""
    from coarse_grain_polyelectrolyte_b200.engine import Engine as E
    w = make_workload(100, pH=np.linspace(2.0, 12.0, 200), c_salt=2e-2)
    state = random_state(w, seed=9)
    out = []
    for planned in (True, False):
        if planned:
            monkeypatch.delenv("CGPE_NO_WAVE_PLAN", raising=False)
        else:
            monkeypatch.setenv("CGPE_NO_WAVE_PLAN", "1")
        e = load(E(w.cfg), w, state)
        e.set_engine("solo")
        sp = e.sample_params(4, md_steps_per_burst=40, use_md=True, mc_blocks=((0, 171), (1, 11)), p_burst1=0.7, p_burst2=0.2)
        obs, _ = e.sample_loop(sp)
        out.append((obs, e.get_state(), e.last_call_ms()[1]))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        np.testing.assert_array_equal(a, b)
    assert out[0][2] == 3 and out[1][2] == 1   # pair count + plan + loop against the loop alone


@pytest.mark.parametrize("R", [74, 75, 147, 148, 149, 200])
def test_cluster_plan_covers_every_replica_for_any_replica_count(R):
    """The launch plan of the cluster engine has three regimes on a 148-SM device -- every cluster alone (R <= 74), some alone
    and the others sharing in pairs (74 < R <= 148; at exactly 148 none alone), waves (R > 148): in each of them every replica
    must be run exactly once, i.e. the results equal the one-CTA engine's, replica by replica."""
    from coarse_grain_polyelectrolyte_b200.engine import Engine as E
    w = make_workload(100, pH=np.linspace(2.0, 12.0, R), c_salt=2e-2)
    state = random_state(w, seed=11)
    out = []
    for name in ("duo", "solo"):
        e = load(E(w.cfg), w, state)
        e.set_engine(name)
        sp = e.sample_params(3, md_steps_per_burst=30, use_md=True, mc_blocks=((0, 171), (1, 11)), p_burst1=0.7, p_burst2=0.2)
        obs, _ = e.sample_loop(sp)
        out.append((obs, e.counters()))
    np.testing.assert_array_equal(out[0][0][..., :3], out[1][0][..., :3])
    np.testing.assert_allclose(out[0][0][..., 3:5], out[1][0][..., 3:5], rtol=1e-5)
    for k in ("md_steps", "trials", "accepted"):
        np.testing.assert_array_equal(out[0][1][k], out[1][1][k])


def test_fused_launch_issued_on_both_engines_runs_on_exactly_one():
    """Left to itself the library issues the fused sampling launch of a long-chain shard on both shared-memory engines and
    the plan picks one on the device: the results must be those of that engine (not twice the samples, not none), and
    engine_info names it."""
    from coarse_grain_polyelectrolyte_b200.engine import Engine as E
    w = make_workload(1000, pH=np.linspace(2.0, 12.0, 100), c_salt=1e-3)
    state = random_state(w, seed=21, jitter=0.02)
    runs = {}
    for name in ("auto", "solo", "duo"):
        e = load(E(w.cfg), w, state)
        if name != "auto":
            e.set_engine(name)
        sp = e.sample_params(3, md_steps_per_burst=20, use_md=True, mc_blocks=((0, 501), (1, 11)), p_burst1=0.7, p_burst2=0.2)
        obs, _ = e.sample_loop(sp)
        runs[name] = (obs, e.counters(), e.last_call_ms()[1], e.engine_info()["engine"])
    picked = runs["auto"][3]
    assert picked in ("solo", "duo") and runs["auto"][2] == 4   # pair count, plan, two loop kernels (one of them returns at once)
    np.testing.assert_array_equal(runs["auto"][0][..., :3], runs[picked][0][..., :3])
    np.testing.assert_allclose(runs["auto"][0][..., 3:5], runs[picked][0][..., 3:5], rtol=1e-6)
    for k in ("md_steps", "trials", "accepted"):
        np.testing.assert_array_equal(runs["auto"][1][k], runs[picked][1][k])
    assert runs["auto"][1]["md_steps"].max() <= 3 * 40
