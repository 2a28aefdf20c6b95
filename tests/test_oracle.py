"""Pins the CPU oracle: Philox known answers, closed-form few-body energies and forces, finite
differences, the ideal titration curve, list-mode == all-pairs, and the committed golden vectors.

The reference holds no golden vector for this path (its arithmetic is ESPResSo's); these are the
anchors SURVEY.md §8c lists.  PARITY UNPINNED against ESPResSo itself.
"""
import os

import numpy as np
import pytest

from coarse_grain_polyelectrolyte_b200 import _abi
from helpers import load, make_workload, random_state

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def few_body_cfg(n, p, **kw):
    eps, sig, pref, kappa, rc, L = p[:6]
    base = dict(n_replicas=1, n_beads=n, n_types=1, box_l=L, time_step=1e-3, kappa=kappa, r_cut=rc,
                dh_prefactor=pref, wca_eps=[[eps]], wca_sig=[[sig]])
    base.update(kw)
    return _abi.make_config(**base)


def test_philox_known_answers(oracle_mod):
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(x) for x in oracle_mod.philox4x32_10(ctr, key)] == want


def test_pair_closed_form(oracle_mod):
    g = np.load(os.path.join(GOLD, "few_body.npz"))
    p = g["pair_params"]
    for key in ("pair_x", "pair_pbc_x"):
        cfg = few_body_cfg(2, p, use_wca=1, use_dh=1, bond_kind=_abi.BOND_HARMONIC, bond_k=p[6], bond_r0=p[7])
        o = oracle_mod.OracleEngine(cfg)
        x = g[key].copy()
        if key == "pair_pbc_x":      # bonds use unfolded coordinates: unfold the partner first
            cfg2 = few_body_cfg(2, p, use_wca=1, use_dh=1)
            o = oracle_mod.OracleEngine(cfg2)
            o.set_state(x.astype(np.float32), None, np.zeros(2, np.uint8))
            o.set_positions_f64(x[None, :, :3])
            e = o.energy()[0]
            np.testing.assert_allclose(e[2], g["pair_energy"][2], rtol=2e-6)
            np.testing.assert_allclose(e[4], g["pair_energy"][4], rtol=2e-5)
            continue
        o.set_state(x.astype(np.float32), None, np.zeros(2, np.uint8))
        o.set_positions_f64(x[None, :, :3])      # exact coordinates (set_state takes float32)
        np.testing.assert_allclose(o.energy()[0], g["pair_energy"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(o.forces_f64()[0], g["pair_force"], rtol=1e-5, atol=1e-7)


def test_angle_and_dihedral_closed_form(oracle_mod):
    g = np.load(os.path.join(GOLD, "few_body.npz"))
    kb, phi0 = g["angle_params"]
    cfg = _abi.make_config(n_replicas=1, n_beads=3, n_types=1, box_l=40.0, time_step=1e-3, use_angle=1,
                           angle_k=kb, angle_phi0=phi0)
    o = oracle_mod.OracleEngine(cfg)
    o.set_state(g["angle_x"].astype(np.float32), None, np.zeros(3, np.uint8))
    o.set_positions_f64(g["angle_x"][None, :, :3])
    np.testing.assert_allclose(o.energy()[0, 3], g["angle_energy"], rtol=1e-6)
    np.testing.assert_allclose(o.forces_f64()[0], g["angle_force"], rtol=1e-6, atol=1e-9)
    kd, mult, phase = g["dih_params"]
    cfg = _abi.make_config(n_replicas=1, n_beads=4, n_types=1, box_l=40.0, time_step=1e-3, use_dihedral=1,
                           dihedral_k=kd, dihedral_mult=int(mult), dihedral_phase=phase)
    for key, want in zip(("dih_cis_x", "dih_trans_x", "dih_gauche_x"), g["dih_energy"]):
        o = oracle_mod.OracleEngine(cfg)
        o.set_state(g[key].astype(np.float32), None, np.zeros(4, np.uint8))
        o.set_positions_f64(g[key][None, :, :3])
        np.testing.assert_allclose(o.energy()[0, 3], want, atol=1e-6)
        f = o.forces_f64()[0]
        np.testing.assert_allclose(f.sum(axis=0), 0.0, atol=1e-9)     # momentum
        np.testing.assert_allclose(f, 0.0, atol=1e-5)                 # cis/trans/gauche are stationary for mult=3, phase=pi


@pytest.mark.parametrize("flags", [None, {"USE_FENE": True}, {"USE_DIHEDRAL_POT": False}])
def test_forces_are_minus_grad_energy(oracle_mod, flags):
    w = make_workload(40, pH=[8.0], flags=flags)
    state = random_state(w, seed=4, jitter=0.03)
    o = load(oracle_mod.OracleEngine(w.cfg), w, state)
    x0 = o.state_f64()[0][0].copy()
    F = o.forces_f64()[0]
    h, worst = 1e-5, 0.0
    for i in range(0, 40, 3):
        for k in range(3):
            xp, xm = x0.copy(), x0.copy()
            xp[i, k] += h
            xm[i, k] -= h
            o.set_positions_f64(xp[None]); ep = o.energy()[0, 0]
            o.set_positions_f64(xm[None]); em = o.energy()[0, 0]
            worst = max(worst, abs(-(ep - em) / (2 * h) - F[i, k]) / max(1.0, abs(F[i, k])))
    assert worst < 1e-6
    assert np.abs(F.sum(axis=0)).max() < 1e-9 * np.abs(F).max() * 40


@pytest.mark.parametrize("name", ["chain_n52.npz", "chain_n100.npz", "chain_n100_np.npz"])
def test_golden_chain(oracle_mod, name):
    g = np.load(os.path.join(GOLD, name))
    kw = dict(nanoparticle=dict(charge=float(g["np_charge"]), clearance=0.9)) if "np_charge" in g else {}
    w = make_workload(int(g["n_mon"]), pH=[8.0], **kw)
    o = load(oracle_mod.OracleEngine(w.cfg), w, (g["xyzq"], g["vel"], g["type"]))
    np.testing.assert_allclose(o.energy(), g["energy"], rtol=1e-12)
    np.testing.assert_allclose(o.forces_f64(), g["forces"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(o.min_dist(), g["min_dist"], rtol=1e-12)
    rg, ree, _ = o.observables()
    np.testing.assert_allclose(rg, g["rg"], rtol=1e-12)
    np.testing.assert_allclose(ree, g["ree"], rtol=1e-12)


def test_golden_trial_log(oracle_mod):
    g = np.load(os.path.join(GOLD, "trials_n100.npz"))
    w = make_workload(int(g["n_mon"]), pH=[float(g["pH"])])
    o = load(oracle_mod.OracleEngine(w.cfg), w, (g["xyzq"], g["vel"], g["type"]))
    rec = o.mc_run(0, g["trace"].shape[1], trace=True)
    for f in ("particle", "direction", "accepted", "u"):
        np.testing.assert_array_equal(rec[f], g["trace"][f])
    np.testing.assert_allclose(rec["delta_e"], g["trace"]["delta_e"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(o.get_state()[2], g["final_type"])
    assert 0.05 < rec["accepted"].mean() < 0.95


def test_ideal_titration_curve(oracle_mod):
    """Interactions off -> alpha(pH) = 1/(1+10^(pK-pH)), the only closed form the reference
    carries (utils.py:15-16)."""
    g = np.load(os.path.join(GOLD, "titration.npz"))
    pH = g["pH"][::4]
    w = make_workload(100, pH=pH, flags={"USE_WCA": False, "USE_ELECTROSTATICS": False})
    o = w.init_engine(oracle_mod.OracleEngine(w.cfg))
    n1, n2 = w.n_sites
    sp = o.sample_params(1500, use_md=False, mc_blocks=((0, 5 * n1 + 1), (1, 5 * n2 + 1)))
    obs, _ = o.sample_loop(sp)
    obs = obs[:, 100:]
    alpha = obs[..., 0].mean(axis=1) / n1
    ideal = g["alpha_pK1"][::4]
    sem = np.sqrt(ideal * (1 - ideal) / (n1 * obs.shape[1] / 4.0)) + 2e-3
    assert np.all(np.abs(alpha - ideal) < 5 * sem), (alpha, ideal)
    np.testing.assert_array_equal(obs[..., 2], obs[..., 0] + obs[..., 1])


def test_list_mode_equals_all_pairs(oracle_mod):
    w = make_workload(120, pH=[7.5, 9.0])
    state = random_state(w, seed=8, jitter=0.02)
    a = load(oracle_mod.OracleEngine(w.cfg), w, state)
    b = load(oracle_mod.OracleEngine(w.cfg), w, state)
    b.set_list_mode(True, 0.6)
    np.testing.assert_allclose(a.energy(), b.energy(), rtol=1e-12)
    np.testing.assert_allclose(a.forces_f64(), b.forces_f64(), rtol=1e-10, atol=1e-10)
    sp = a.sample_params(15, md_steps_per_burst=40, use_md=True, mc_blocks=((0, 61), (1, 11)), p_burst1=0.7, p_burst2=0.3)
    oa, _ = a.sample_loop(sp)
    ob, _ = b.sample_loop(sp)
    np.testing.assert_array_equal(oa[..., :3], ob[..., :3])
    np.testing.assert_allclose(oa[..., 3:], ob[..., 3:], rtol=1e-9)
    assert b.counters()["list_rebuilds"].min() >= 1
    sa, sb = a.list_stats(), b.list_stats()
    np.testing.assert_array_equal(sa[:, [1, 3]], sb[:, [1, 3]])   # in-cutoff counts do not depend on the skin
    assert np.all(sb[:, 0] <= sa[:, 0])                          # skin 0.6 lists are shorter than skin 0.8 ones


def test_langevin_thermostat_temperature(oracle_mod):
    """<T> = 2 E_kin / (3 N) -> kT = 1 (the notebook's T_inst check, NB:234-248)."""
    w = make_workload(60, pH=[12.0], flags={"USE_ELECTROSTATICS": False})
    o = w.init_engine(oracle_mod.OracleEngine(w.cfg))
    o.set_list_mode(True, 0.8)
    o.md_run(3000)
    temps = []
    for _ in range(60):
        o.md_run(100)
        temps.append(2.0 * o.energy()[0, 1] / (3 * 60))
    assert abs(np.mean(temps) - 1.0) < 0.06, np.mean(temps)


def test_energy_conservation_without_thermostat(oracle_mod):
    w = make_workload(40, pH=[12.0])
    state = random_state(w, seed=2, jitter=0.01, vel_sigma=0.5)
    o = load(oracle_mod.OracleEngine(w.cfg), w, state)
    o.set_replica_params(kT=0.0, gamma=0.0)
    o.md_run(200)                     # let the jitter relax into the stiff bonds first
    e0 = o.energy()[0, 0]
    o.md_run(2000)
    e1 = o.energy()[0, 0]
    assert abs(e1 - e0) < 2e-3 * abs(e0)


def test_sharding_invariance(oracle_mod):
    pH = np.linspace(6.0, 11.0, 6)
    w_all = make_workload(50, pH=pH)
    state = random_state(w_all, seed=31)
    kw = dict(md_steps_per_burst=10, use_md=True, mc_blocks=((0, 41), (1, 11)), p_burst1=0.7, p_burst2=0.2)
    e = load(oracle_mod.OracleEngine(w_all.cfg), w_all, state)
    ref, _ = e.sample_loop(e.sample_params(12, **kw))
    parts = []
    for lo in (0, 3):
        w = make_workload(50, pH=pH[lo:lo + 3], replica_offset=lo)
        s = load(oracle_mod.OracleEngine(w.cfg), w, tuple(a[lo:lo + 3] for a in state))
        parts.append(s.sample_loop(s.sample_params(12, **kw))[0])
    np.testing.assert_array_equal(ref, np.concatenate(parts, axis=0))


def test_dihedral_gradient_is_bounded_near_straight_bond_angles(oracle_mod):
    """The torsion gradient grows like 1/sin(theta); it is evaluated with sin(theta) >= 0.02 so a
    thermally reachable straight angle cannot blow the integration up (cgpe.h, dihedral_sin_min).
    Away from the 1.1 degree cone nothing changes."""
    def forces(eps_angle, sin_min):
        cfg = _abi.make_config(n_replicas=1, n_beads=4, n_types=1, box_l=40.0, time_step=1e-3, use_dihedral=1,
                               dihedral_k=6.5, dihedral_mult=3, dihedral_phase=np.pi, dihedral_sin_min=sin_min)
        o = oracle_mod.OracleEngine(cfg)
        # p1-p2-p3 almost collinear (angle pi - eps_angle), p4 out of plane at an oblique torsion
        x = np.array([[-np.cos(eps_angle), np.sin(eps_angle), 0, 0], [0, 0, 0, 0], [1, 0, 0, 0],
                      [1.5, 0.6 * np.cos(0.7), 0.6 * np.sin(0.7), 0]], float) + [5, 5, 5, 0]
        o.set_state(x.astype(np.float32), None, np.zeros(4, np.uint8))
        o.set_positions_f64(x[None, :, :3])
        return o.forces_f64()[0]
    wild, tame = forces(1e-3, -1.0), forces(1e-3, 0.0)
    assert np.abs(wild).max() > 5e3                      # ESPResSo's expression: ~3k/(l sin theta)
    assert np.abs(tame).max() < 1.2e3                    # regularised
    np.testing.assert_allclose(tame.sum(axis=0), 0.0, atol=1e-9)
    np.testing.assert_allclose(forces(0.3, -1.0), forces(0.3, 0.0), rtol=1e-13)   # outside the cone: identical


@pytest.mark.parametrize("charge", [-20.0, 20.0])
def test_fixed_charged_nanoparticle(oracle_mod, charge):
    """BASELINE config 4: one fixed charged sphere (type "NP", fixed_type_mask) in a spectator slot next
    to the chain, implicit-ion reactions.  The sphere acts on the beads (forces = -grad E with the
    sphere held), feels nothing itself, never moves, and shifts the protonation equilibrium."""
    w = make_workload(40, pH=[8.0], nanoparticle=dict(charge=charge, clearance=0.9))
    assert w.cfg.n_ion_slots == 1 and w.cfg.fixed_type_mask == 1 << w.ru.type_index["NP"]
    assert all(w.cfg.reactions[m].type_ion < 0 for m in range(w.cfg.n_reactions))
    state = random_state(w, seed=4, jitter=0.03, protonated_fraction=0.7)
    o = load(oracle_mod.OracleEngine(w.cfg), w, state)
    i_np = 40
    x0 = o.state_f64()[0][0].copy()
    d = np.linalg.norm(x0[:40] - x0[i_np], axis=1)
    assert (d < w.r_cut[0]).sum() >= 3                       # the sphere is felt
    F = o.forces_f64()[0]
    assert np.all(F[i_np] == 0.0) and np.all(o.state_f64()[1][0, i_np] == 0.0)
    near = np.argsort(d)[:6]
    h, worst = 1e-5, 0.0
    for i in near:
        for k in range(3):
            xp, xm = x0.copy(), x0.copy()
            xp[i, k] += h
            xm[i, k] -= h
            o.set_positions_f64(xp[None]); ep = o.energy()[0, 0]
            o.set_positions_f64(xm[None]); em = o.energy()[0, 0]
            worst = max(worst, abs(-(ep - em) / (2 * h) - F[i, k]) / max(1.0, abs(F[i, k])))
    assert worst < 1e-6
    o.set_positions_f64(x0[None])
    # the same chain without the sphere: energies differ (it interacts), bead forces differ
    w0 = make_workload(40, pH=[8.0])
    s0 = tuple(a[:, :40] for a in state)
    o0 = load(oracle_mod.OracleEngine(w0.cfg), w0, s0)
    assert abs(o.energy()[0, 0] - o0.energy()[0, 0]) > 1e-3
    assert np.abs(F[:40] - o0.forces_f64()[0]).max() > 1e-3
    # dynamics and trial moves: the sphere stays put, implicit-ion bookkeeping (N_B = N_N + N_N2)
    o.md_run(300); o.mc_run(0, 400); o.mc_run(1, 40); o.steepest_descent(5, 0.0, 0.1, 0.1); o.md_run(50)
    x1, v1, _ = o.state_f64()
    assert np.all(x1[0, i_np] == x0[i_np]) and np.all(v1[0, i_np] == 0.0)
    assert np.abs(x1[0, :40] - x0[:40]).max() > 1e-2
    cnt = o.observables()[2][0]
    assert cnt[2] == cnt[0] + cnt[1]
    # list mode agrees with all pairs with the large sphere in the rows
    b = load(oracle_mod.OracleEngine(w.cfg), w, state)
    b.set_list_mode(True, 0.6)
    a = load(oracle_mod.OracleEngine(w.cfg), w, state)
    np.testing.assert_allclose(a.energy(), b.energy(), rtol=1e-12)
    np.testing.assert_allclose(a.forces_f64(), b.forces_f64(), rtol=1e-10, atol=1e-10)


def test_nanoparticle_shifts_the_titration_curve(oracle_mod):
    """Sites are cationic when protonated (NH+ <-> N): a negative sphere in contact stabilises the
    protonated state (alpha, the deprotonated fraction, drops), a positive one destabilises it."""
    alpha = {}
    for q in (-20.0, 0.0, 20.0):
        w = make_workload(40, pH=[8.18] * 4, flags={"USE_DIHEDRAL_POT": False},
                          nanoparticle=dict(charge=q, clearance=0.9))
        o = w.init_engine(oracle_mod.OracleEngine(w.cfg))
        sp = o.sample_params(260, md_steps_per_burst=20, use_md=True, mc_blocks=((0, 66), (1, 6)), p_burst1=0.7, p_burst2=0.3)
        obs, _ = o.sample_loop(sp)
        alpha[q] = obs[:, 60:, 0].mean() / w.n_sites[0]
    assert alpha[-20.0] < alpha[0.0] - 0.01 < alpha[20.0] - 0.02, alpha


def test_oracle_invariances_hold_for_random_states(oracle_mod):
    """Size-independent properties of the restatement itself (hypothesis-driven seeds): rigid translations leave every
    energy component unchanged, total = kinetic + coulomb + bonded + non_bonded, forces sum to zero, and replicas do
    not see each other (permuting them permutes the results)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=8, deadline=None)
    @given(seed=st.integers(0, 10_000), shift=st.tuples(*[st.floats(-3.0, 3.0)] * 3), frac=st.floats(0.0, 1.0))
    def check(seed, shift, frac):
        w = make_workload(30, pH=[7.0, 9.0, 11.0])
        state = random_state(w, seed=seed, jitter=0.03, protonated_fraction=frac)
        o = load(oracle_mod.OracleEngine(w.cfg), w, state)
        e0, f0 = o.energy(), o.forces_f64()
        np.testing.assert_allclose(e0[:, 0], e0[:, 1:].sum(axis=1), rtol=1e-12)
        assert np.abs(f0.sum(axis=1)).max() < 1e-9 * max(1.0, np.abs(f0).max()) * 30
        x = o.state_f64()[0]
        o.set_positions_f64(x + np.asarray(shift))
        np.testing.assert_allclose(o.energy()[:, [2, 3, 4]], e0[:, [2, 3, 4]], rtol=1e-9, atol=1e-9)
        perm = [2, 0, 1]
        w2 = make_workload(30, pH=[11.0, 7.0, 9.0])
        o2 = load(oracle_mod.OracleEngine(w2.cfg), w2, tuple(a[perm] for a in state))
        np.testing.assert_allclose(o2.energy(), e0[perm], rtol=1e-12)
        np.testing.assert_allclose(o2.forces_f64(), f0[perm], rtol=1e-12, atol=1e-12)

    check()


@pytest.mark.parametrize("explicit_ions", [False, True])
def test_local_delta_e_equals_the_whole_system_energy_difference(oracle_mod, explicit_ions):
    """ESPResSo's reaction engine accepts on E_pot(whole system, after) - E_pot(whole system, before)
    (E.py:70-71,94-95 call it).  The oracle -- like the CUDA kernels -- evaluates only the terms a
    trial changes.  One trial per call, the full potential energy around it: for every accepted trial the
    difference of the totals equals the trace's delta_e; a rejected trial leaves the total unchanged."""
    w = make_workload(50, pH=[8.18], c_salt=1e-2, explicit_ions=explicit_ions)   # C1's chain and salt
    state = random_state(w, seed=23, jitter=0.03)
    if explicit_ions:   # random_state flips site types without touching the H+ pool: keep the workload's own protonation
        xyzq, vel, types = state
        n = w.cfg.n_chains * w.cfg.n_beads
        types[:, :n] = w.type[:n]; xyzq[:, :n, 3] = w.q[:n]
        types[:, n:] = w.type[n:]; xyzq[:, n:, 3] = w.q[n:]
        state = (xyzq, vel, types)
    o = load(oracle_mod.OracleEngine(w.cfg), w, state)
    o.md_run(50)                                       # off the synthetic start

    def e_pot():
        e = o.energy()[0]
        return e[2] + e[3] + e[4]                      # coulomb + bonded + non_bonded (an inserted H+ brings kinetic energy)

    n_acc = n_rej = n_excl = 0
    worst = 0.0
    for k in range(400):
        rx = k % 2
        before = e_pot()
        tr = o.mc_run(rx, 1, trace=True)[0, 0]
        after = e_pot()
        if tr["accepted"]:
            n_acc += 1
            err = abs((after - before) - tr["delta_e"]) / max(abs(tr["delta_e"]), 1.0)
            worst = max(worst, err)
        else:
            n_rej += 1
            n_excl += int(tr["excluded"])
            assert after == before
    assert n_acc > 40 and n_rej > 40, (n_acc, n_rej)
    assert worst < 2e-6, worst                          # delta_e travels through the trace as float32
    if explicit_ions:
        assert n_excl >= 0
