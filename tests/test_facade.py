"""Host-side logic of the espressomd-shaped facade: call surface, argument checking, topology
recognition, config assembly.  No GPU needed: nothing here asks for arithmetic."""
import math
import os

import numpy as np
import pytest

os.environ.setdefault("CGPE_QUIET", "1")
from coarse_grain_polyelectrolyte_b200 import _abi, espressomd  # noqa: E402
from coarse_grain_polyelectrolyte_b200.espressomd import electrostatics, interactions, polymer, reaction_methods  # noqa: E402


def build_chain(n=12, n_replicas=1, angle=True, dihedral=True):
    s = espressomd.System(box_l=[30.0] * 3, n_replicas=n_replicas)
    s.time_step = 1e-3
    bond = interactions.HarmonicBond(k=100.0, r_0=1.0)
    s.bonded_inter.add(bond)
    bend = interactions.AngleHarmonic(bend=5.0, phi0=2 * math.pi / 3)
    dih = interactions.Dihedral(bend=1.0, mult=3, phase=math.pi)
    if angle:
        s.bonded_inter.add(bend)
    if dihedral:
        s.bonded_inter.add(dih)
    pos = polymer.linear_polymer_positions(n_polymers=1, beads_per_chain=n, bond_length=1.0, seed=3, min_distance=0.8)[0]
    for i, p in enumerate(pos):
        s.part.add(id=i, pos=p, type=1 if i % 3 == 0 else 3, q=0.0)
        if i > 0:
            s.part.by_id(i).add_bond((bond, i - 1))
            if angle and i < n - 1:
                s.part.by_id(i).add_bond((bend, i - 1, i + 1))
            if dihedral and i < n - 2:
                s.part.by_id(i).add_bond((dih, i - 1, i + 1, i + 2))
    return s


def test_call_surface_of_the_reference_exists():
    """Every name of SURVEY.md §8b that src/modules uses."""
    espressomd.assert_features(["WCA", "ELECTROSTATICS"])
    with pytest.raises(RuntimeError):
        espressomd.assert_features(["P3M_GPU"])
    s = espressomd.System(box_l=[10.0, 10.0, 10.0])
    for name in ("time_step", "cell_system", "periodicity", "time", "force_cap", "part", "bonded_inter",
                 "non_bonded_inter", "actors", "integrator", "thermostat", "analysis", "number_of_particles"):
        assert hasattr(s, name), name
    for name in ("run", "set_vv", "set_steepest_descent"):
        assert hasattr(s.integrator, name)
    for name in ("min_dist", "energy", "calc_rg", "calc_re"):
        assert hasattr(s.analysis, name)
    assert hasattr(s.thermostat, "set_langevin") and hasattr(s.cell_system, "set_n_square")
    s.cell_system.skin = 0.1
    s.cell_system.set_n_square()
    rx = reaction_methods.ConstantpHEnsemble(kT=1.0, exclusion_range=1.0, seed=77, constant_pH=4.0)
    for name in ("add_reaction", "set_non_interacting_type", "reaction", "constant_pH"):
        assert hasattr(rx, name)
    with pytest.raises(NotImplementedError):
        electrostatics.P3M(prefactor=2.0, accuracy=1e-5)


def test_particles_and_topology():
    s = build_chain(12)
    assert len(s.part) == 12
    assert s.number_of_particles(type=1) == 4 and s.number_of_particles(type=3) == 8
    np.testing.assert_array_equal(s.part.by_ids(np.arange(12)).q, np.zeros(12))
    assert s.part.by_id(3).type == 1
    n_chains, n_beads, bond, ang, dih = s._topology()
    assert (n_chains, n_beads) == (1, 12) and bond is not None and ang is not None and dih is not None
    cfg = s._build_config()
    assert cfg.n_beads == 12 and cfg.bond_kind == _abi.BOND_HARMONIC and cfg.use_angle == 1 and cfg.use_dihedral == 1
    assert cfg.dihedral_mult == 3 and cfg.angle_phi0 == pytest.approx(2 * math.pi / 3)
    # vectorised add of free particles = explicit ions: they go to ion slots behind the chain
    s.part.add(pos=np.random.random((5, 3)) * 30, type=[2] * 5, q=[1.0] * 5)
    assert len(s.part) == 17
    cfg = s._build_config()
    assert cfg.n_ion_slots == 5 and cfg.n_beads == 12


def test_unsupported_topologies_are_refused():
    s = espressomd.System(box_l=[30.0] * 3)
    s.time_step = 1e-3
    bond = interactions.HarmonicBond(k=1.0, r_0=1.0)
    s.bonded_inter.add(bond)
    for i in range(4):
        s.part.add(id=i, pos=[i, 0, 0], type=0)
    s.part.by_id(2).add_bond((bond, 0))            # not the predecessor
    with pytest.raises(NotImplementedError):
        s._topology()
    with pytest.raises(ValueError):
        s.part.by_id(1).add_bond((interactions.HarmonicBond(k=1.0, r_0=1.0), 0))   # never added to bonded_inter
    with pytest.raises(NotImplementedError):
        s.part.add(id=10, pos=[0, 0, 0])            # ids must be consecutive
    with pytest.raises(NotImplementedError):
        espressomd.System(box_l=[10.0, 20.0, 10.0])


def test_wca_table_and_dh_and_reactions_reach_the_config():
    s = build_chain(9, n_replicas=3, dihedral=False)
    s.non_bonded_inter[1, 3].wca.set_params(epsilon=0.2, sigma=0.9)
    s.non_bonded_inter[3, 3].wca.set_params(epsilon=0.3, sigma=1.1)
    assert s.non_bonded_inter[3, 1].wca.get_params() == {"epsilon": 0.2, "sigma": 0.9}
    dh = electrostatics.DH(prefactor=2.0, kappa=[0.1, 0.2, 0.3], r_cut=[10.0, 5.0, 3.3])
    s.actors.add(dh)
    with pytest.raises(RuntimeError):
        s.actors.add(electrostatics.DH(prefactor=2.0, kappa=0.1, r_cut=5.0))
    rx = reaction_methods.ConstantpHEnsemble(kT=1.0, exclusion_range=1.4, seed=77, constant_pH=4.0)
    rx.add_reaction(gamma=10 ** -8.18, reactant_types=[0], product_types=[1, 2], default_charges={0: 1.0, 1: 0.0, 2: 1.0})
    rx.set_non_interacting_type(type=8)
    rx.constant_pH = [5.0, 6.0, 7.0]
    with pytest.raises(ValueError):
        rx.constant_pH = [5.0, 6.0]
    with pytest.raises(NotImplementedError):
        rx.add_reaction(gamma=1.0, reactant_types=[6], product_types=[7, 2], default_charges={6: 1, 7: 0, 2: 1})
    cfg = s._build_config()
    assert cfg.n_replicas == 3 and cfg.use_wca == 1 and cfg.use_dh == 1 and cfg.n_reactions == 1
    assert cfg.wca_eps[1 * _abi.MAX_TYPES + 3] == 0.2 and cfg.wca_sig[3 * _abi.MAX_TYPES + 3] == 1.1
    assert cfg.r_cut == 10.0 and cfg.dh_prefactor == 2.0
    r0 = cfg.reactions[0]
    assert (r0.type_prot, r0.type_deprot, r0.type_ion, r0.seed) == (0, 1, 2, 77)
    assert r0.pK == pytest.approx(8.18) and r0.q_prot == 1.0 and r0.exclusion_range == 1.4


def test_thermostat_and_integrator_argument_checks():
    s = build_chain(6)
    with pytest.raises(ValueError, match="seed"):
        s.thermostat.set_langevin(kT=1.0, gamma=1.0)           # ESPResSo demands a seed the first time
    s.thermostat.set_langevin(kT=0.0, gamma=1.0, seed=4)
    s.thermostat.set_langevin(kT=2.0, gamma=1.0)                # later calls keep the seed (E.py:57)
    assert s.thermostat.seed == 4
    with pytest.raises(ValueError):
        s.thermostat.set_langevin(kT=-1.0, gamma=1.0)
    with pytest.raises(ValueError):
        s.integrator.set_steepest_descent(f_max=-1, gamma=0.1, max_displacement=0.1)
    with pytest.raises(ValueError):
        s.integrator.run(-5)
    s.force_cap = 5
    with pytest.raises(ValueError):
        s.force_cap = -1
    s.time = 0.0
    assert s.time == 0.0


def test_linear_polymer_positions_contract():
    espressomd.System(box_l=[50.0] * 3)
    p = polymer.linear_polymer_positions(n_polymers=2, beads_per_chain=40, bond_length=1.5, min_distance=0.5, seed=12)
    assert p.shape == (2, 40, 3)
    np.testing.assert_allclose(np.linalg.norm(np.diff(p, axis=1), axis=2), 1.5, rtol=1e-12)
    flat = p.reshape(-1, 3)
    d = flat[:, None, :] - flat[None, :, :]
    d -= 50.0 * np.rint(d / 50.0)
    r = np.sqrt((d ** 2).sum(-1)) + np.eye(80) * 10
    assert r.min() >= 0.5
    q = polymer.linear_polymer_positions(n_polymers=2, beads_per_chain=40, bond_length=1.5, min_distance=0.5, seed=12)
    np.testing.assert_array_equal(p, q)                          # seeded: reproducible


def test_model_dictionary_to_topology():
    """Model's site pattern (M.py:126-155): every third bead titratable, chain ends species 2."""
    from coarse_grain_polyelectrolyte_b200 import topology
    role = topology.chain_site_pattern(100)
    assert (role == 1).sum() == 32 and (role == 2).sum() == 2 and role[0] == 2 and role[99] == 2
    role = topology.chain_site_pattern(50)           # index 49 is not a site: only one end is species 2
    assert (role == 1).sum() == 16 and (role == 2).sum() == 1
    role = topology.chain_site_pattern(1000)
    assert (role == 1).sum() == 332 and (role == 2).sum() == 2


def test_fixed_nanoparticle_becomes_a_spectator_slot():
    """part.add(..., fix=[True]*3) behind the chain (BASELINE config 4): a spectator slot of a fixed
    type; the reactions stay the implicit-ion ones (no H+ pool is created)."""
    s = build_chain(9, dihedral=False)
    s.non_bonded_inter[1, 5].wca.set_params(epsilon=0.2, sigma=3.0)
    s.non_bonded_inter[3, 5].wca.set_params(epsilon=0.2, sigma=3.1)
    s.actors.add(electrostatics.DH(prefactor=2.0, kappa=0.16, r_cut=6.0))
    rx = reaction_methods.ConstantpHEnsemble(kT=1.0, exclusion_range=1.4, seed=77, constant_pH=8.0)
    rx.add_reaction(gamma=10 ** -8.18, reactant_types=[0], product_types=[1, 2], default_charges={0: 1.0, 1: 0.0, 2: 1.0})
    with pytest.raises(NotImplementedError):
        s.part.add(pos=[1.0, 2.0, 3.0], type=5, q=-20.0, fix=[True, False, True])
    s.part.add(pos=[15.0, 15.0, 15.0], type=5, q=-20.0, fix=[True, True, True])
    cfg = s._build_config()
    assert cfg.n_ion_slots == 1 and cfg.fixed_type_mask == 1 << 5 and cfg.n_types >= 6
    assert cfg.reactions[0].type_ion == -1
    # a free particle of the same type contradicts the type-wide pin
    s.part.add(pos=[5.0, 5.0, 5.0], type=5, q=-20.0)
    with pytest.raises(NotImplementedError):
        s._build_config()
