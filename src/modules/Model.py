"""Model — topology, force field and reactions of one weak-polyelectrolyte system.

Drop-in for the reference's src/modules/Model.py: `Model(input_dict)` derives from Properties
and leaves the same attributes behind (`system`, `bond_pot`, `bend`, `dihedral`, `coulomb`,
`reaction1`, `reaction2`, `polymers`), built through the same espressomd-shaped calls, which
here configure the B200 engine.

Beyond the reference:
  * `replicas=dict(pH=[...], c_salt=[...])` batches R independent copies of the system over the
    pH x ionic-strength grid (the reference loops over pH values serially, notebook cell 4);
  * "Simulation configuration" may carry "ION_MODEL": "explicit" (default, as in the reference: B+/Na+/Cl-
    particles, a reaction inserts or deletes an H+, M.py:173-198) or "implicit" (the chain alone, small ions only
    through the Debye-Hueckel screening: the move BASELINE.json's north star describes and bench.py times; N_B is
    then reported through the identity N_B = N_N + N_N2).  The choice is recorded in `ion_model`;
  * "Simulation configuration" may carry "REFERENCE_FAITHFUL": True -- the two places where the defaults depart from
    what the reference computes are switched back: the torsion gradient keeps ESPResSo's guard only (no sin-theta
    floor, cgpe.h dihedral_sin_min < 0) and the warm-up is the reference's own loop (E.py:39-48: no contracting
    pre-relaxation, no bound on the rounds).  "DIHEDRAL_SIN_MIN": x sets the floor alone;
  * an optional section "Nanoparticle properties": {"charge": q, "sigma": [Angstrom], "epsilon": [kJ/mol],
    "clearance": 1.3} places one fixed charged sphere (type "NP", `fix=[True]*3`) beside the first chain
    (BASELINE config 4; the reference only announces it, README.md:14-15).  Implicit ions only.
"""
import copy

import numpy as np

from coarse_grain_polyelectrolyte_b200 import espressomd, units
from coarse_grain_polyelectrolyte_b200.espressomd import electrostatics, polymer, reaction_methods
from coarse_grain_polyelectrolyte_b200.espressomd.interactions import AngleHarmonic, Dihedral, FeneBond, HarmonicBond

try:                                   # imported as modules.Model (PYTHONPATH=src) ...
    from .Properties import Properties
    from .utils import combination_rule_epsilon, combination_rule_sigma
except ImportError:                    # ... or flat, as the reference's own modules do (PYTHONPATH=src/modules)
    from Properties import Properties
    from utils import combination_rule_epsilon, combination_rule_sigma

espressomd.assert_features(["WCA", "ELECTROSTATICS"])


class Model(Properties):
    def __init__(self, input_dict, replicas=None, replica_offset=0, device=-1):
        self.replica_pH, self.replica_c_salt = self._replica_axes(input_dict, replicas)
        self.nanoparticle = input_dict.get("Nanoparticle properties")
        if self.nanoparticle is not None:   # one more row of the particle table, reduced like the others (P.py:307-311)
            input_dict = copy.deepcopy(input_dict)
            table = input_dict["Particles properties"]
            table["NP"] = {"index": max(v["index"] for v in table.values()) + 1, "charge": self.nanoparticle["charge"],
                           "sigma": self.nanoparticle.get("sigma", 34.8), "epsilon": self.nanoparticle.get("epsilon", 0.3)}
        super().__init__(input_dict, n_replicas=len(self.replica_pH), replica_offset=replica_offset, device=device)
        sim = input_dict["Simulation configuration"]
        self.ion_model = sim.get("ION_MODEL", "explicit")
        if self.ion_model not in ("implicit", "explicit"):
            raise ValueError("ION_MODEL must be 'implicit' or 'explicit'")
        if self.nanoparticle is not None and "ION_MODEL" not in sim:
            self.ion_model = "implicit"     # the only model the fixed sphere is served with (see _add_nanoparticle)
        self.reference_faithful = bool(sim.get("REFERENCE_FAITHFUL", False))
        self.dihedral_sin_min = float(sim.get("DIHEDRAL_SIN_MIN", -1.0 if self.reference_faithful else 0.0))
        self._bonded_interactions()
        self._create_polymer()
        self._add_nanoparticle()
        self._add_salt_ions()
        self._non_bonded_interactions()
        self._long_range_interactions()
        self._add_reactions()

    @staticmethod
    def _replica_axes(input_dict, replicas):
        c0 = input_dict["System properties"]["Concentration salt particles"]
        if not replicas:
            return np.array([np.nan]), np.array([float(c0)])
        pH = np.atleast_1d(np.asarray(replicas.get("pH", [np.nan]), dtype=np.float64))
        cs = np.atleast_1d(np.asarray(replicas.get("c_salt", [c0]), dtype=np.float64))
        n = max(pH.size, cs.size)
        if pH.size not in (1, n) or cs.size not in (1, n):
            raise ValueError("replicas['pH'] and replicas['c_salt'] must have one entry or one per replica")
        if np.any(cs <= 0):
            raise ValueError("salt concentrations must be positive")
        return np.broadcast_to(pH, (n,)).copy(), np.broadcast_to(cs, (n,)).copy()

    # -- M.py:32-54 ----------------------------------------------------------------------------------
    def _bonded_interactions(self):
        k, r0 = self.k_bond_reduced_units.magnitude, self.bond_length_reduced_units.magnitude
        self.bond_pot = FeneBond(k=k, d_r_max=4.0, r_0=r0) if self.USE_FENE else HarmonicBond(k=k, r_0=r0)
        self.system.bonded_inter.add(self.bond_pot)
        if self.USE_BENDING:
            self.bend = AngleHarmonic(bend=self.k_bending_reduced_units.magnitude, phi0=np.pi * (2.0 / 3.0))
            self.system.bonded_inter.add(self.bend)
        if self.USE_DIHEDRAL_POT:
            self.dihedral = Dihedral(bend=self.k_dihedral_reduced_units.magnitude, mult=3, phase=np.pi,
                                     sin_min=self.dihedral_sin_min)
            self.system.bonded_inter.add(self.dihedral)

    # -- M.py:114-170 --------------------------------------------------------------------------------
    def _create_polymer(self):
        self.polymers = polymer.linear_polymer_positions(
            n_polymers=self.n_poly, beads_per_chain=self.n_mon, bond_length=1.5, min_distance=0.5, seed=12)
        part = self.system.part
        for chain in self.polymers:
            last = len(chain) - 1
            for index, position in enumerate(chain):
                pid = len(part)
                # every third bead is titratable, chain ends belong to the second species; all
                # sites start deprotonated (M.py:128-147)
                if index % 3 == 0:
                    name = "N2" if index in (0, self.n_mon - 1) else "N"
                else:
                    name = "CH2"
                part.add(id=pid, pos=position, type=self.type_particles[name], q=self.charge_reduced_units[name])
                if index > 0:
                    me = part.by_id(pid)
                    me.add_bond((self.bond_pot, pid - 1))
                    if self.USE_BENDING and index < last:
                        me.add_bond((self.bend, pid - 1, pid + 1))
                    if self.USE_DIHEDRAL_POT and index < last - 1:
                        me.add_bond((self.dihedral, pid - 1, pid + 1, pid + 2))

    def _add_nanoparticle(self):
        if self.nanoparticle is None:
            return
        if self.ion_model != "implicit":
            raise NotImplementedError("the nanoparticle needs ION_MODEL 'implicit' (an inserted H+ would carry its WCA energy)")
        beads = np.asarray(self.polymers[0], dtype=np.float64)
        clear = self.nanoparticle.get("clearance", 1.3) * 0.25 * (self.sigma_reduced_units["NP"] + max(self.sigma_reduced_units.values()))
        pos = beads.mean(axis=0)
        while np.min(np.linalg.norm(np.asarray(self.polymers).reshape(-1, 3) - pos, axis=1)) < clear:
            pos[0] += 0.25
        self.nanoparticle_id = len(self.system.part)
        self.system.part.add(id=self.nanoparticle_id, pos=pos, type=self.type_particles["NP"],
                             q=self.charge_reduced_units["NP"], fix=[True, True, True])

    # -- M.py:173-198 --------------------------------------------------------------------------------
    def _add_salt_ions(self):
        if self.ion_model == "implicit":
            return
        L = self.box_length_reduced_units
        for name, count in (("B", self.n_acid), ("Cl", self.n_acid), ("Na", self.n_salt), ("Cl", self.n_salt)):
            if count > 0:
                self.system.part.add(pos=np.random.random((count, 3)) * L,
                                     type=[self.type_particles[name]] * count,
                                     q=[self.charge_reduced_units[name]] * count)

    # -- M.py:56-80 ----------------------------------------------------------------------------------
    def _non_bonded_interactions(self):
        if not self.USE_WCA:
            return
        names = list(self.type_particles)
        for a in names:
            for b in names:
                sig = combination_rule_sigma("Berthelot", self.sigma_reduced_units[a], self.sigma_reduced_units[b])
                eps = combination_rule_epsilon("Lorentz", self.epsilon_reduced_units[a], self.epsilon_reduced_units[b])
                self.system.non_bonded_inter[self.type_particles[a], self.type_particles[b]].wca.set_params(
                    epsilon=eps, sigma=0.5 * sig)
        # The reference relaxes overlaps with 100 steepest-descent steps after EVERY one of the
        # 64 type pairs (the call sits inside its double loop, M.py:77-80).  Once, with the full
        # table in place, does the same job.
        self.system.integrator.set_steepest_descent(f_max=0, gamma=0.1, max_displacement=0.1)
        self.system.integrator.run(100)
        self.system.integrator.set_vv()

    # -- M.py:82-112 ---------------------------------------------------------------------------------
    def _long_range_interactions(self):
        if not self.USE_ELECTROSTATICS:
            self.system.cell_system.set_n_square()
            return
        if self.USE_P3M:
            self.coulomb = electrostatics.P3M(prefactor=self.coulomb_prefactor_reduced_units, accuracy=1e-5)
        else:
            kappa = np.array([units.kappa_per_nm(c) for c in self.replica_c_salt]) * self.reduced.sim_length_nm
            self.replica_kappa = kappa
            self.coulomb = electrostatics.DH(prefactor=self.coulomb_prefactor_reduced_units,
                                             kappa=kappa if kappa.size > 1 else float(kappa[0]),
                                             r_cut=1.0 / kappa if kappa.size > 1 else float(1.0 / kappa[0]))
        self.system.actors.add(self.coulomb)

    # -- M.py:200-261 --------------------------------------------------------------------------------
    def _add_reactions(self):
        tp, q = self.type_particles, self.charge_reduced_units
        kT = self.KT.to("sim_energy").magnitude
        made = []
        for prot, deprot, pK in (("NH", "N", self.pK), ("NH2", "N2", self.pK2)):
            excl = combination_rule_sigma("Berthelot", self.sigma_reduced_units[deprot], self.sigma_reduced_units["B"])
            # the reference starts every reaction object at pH 4 (M.py:215,232); with a replica axis the replicas' own
            # pH values are in force from the start, so perform_sampling without equilibrate_pH cannot sample pH 4 by accident
            pH0 = self.replica_pH if np.all(np.isfinite(self.replica_pH)) else 4.0
            rx = reaction_methods.ConstantpHEnsemble(kT=kT, exclusion_range=excl, seed=77,
                                                     constant_pH=pH0 if np.size(pH0) > 1 else float(np.atleast_1d(pH0)[0]))
            rx.add_reaction(gamma=10 ** (-pK), reactant_types=[tp[prot]], product_types=[tp[deprot], tp["B"]],
                            default_charges={tp[prot]: q[prot], tp[deprot]: q[deprot], tp["B"]: q["B"]})
            rx.set_non_interacting_type(type=len(tp))
            made.append((rx, excl))
        (self.reaction1, self.exclusion_radius_reaction1), (self.reaction2, self.exclusion_radius_reaction2) = made
