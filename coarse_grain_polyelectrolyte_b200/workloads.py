"""Synthetic inputs for the BASELINE.json configurations.

Everything here is host-side setup: the parameter dictionary of examples/run_test.py:7-56 of the
reference, reduced with units.py, turned into a cgpe_config plus a start state.  Used by
bench.py, the tests and smoke(); the Model drop-in (src/modules/Model.py) builds the same
things through the espressomd-shaped facade instead.
"""
import copy
import math

import numpy as np

from . import _abi, topology, units

# examples/run_test.py:7-56 of the reference
EXAMPLE_PARAMETERS = {
    "Thermal properties": {"Temperature": 300},
    "System properties": {
        "Number of polymers": 1,
        "Concentration acid particles": 1e-2,
        "Concentration salt particles": 2 * 1e-2,
        "Number of monomers": 100,
        "NH monomers": 32,
        "NH2 monomers": 2,
        "bond length": 3.85,
        "bond constant": 200.0,
        "bending constant": 0.002,
        "dihedral constant": 0.002,
    },
    "Particles properties": {
        "NH": {"index": 0, "charge": +1, "sigma": 2 * 2.75, "epsilon": 0.32},
        "N": {"index": 1, "charge": 0, "sigma": 2 * 3.298, "epsilon": 0.305},
        "B": {"index": 2, "charge": +1, "sigma": 2 * 2.42, "epsilon": 0.20},
        "CH2": {"index": 3, "charge": 0, "sigma": 2 * 3.800, "epsilon": 0.47},
        "Na": {"index": 4, "charge": +1, "sigma": 2 * 2.21, "epsilon": 0.45},
        "Cl": {"index": 5, "charge": -1, "sigma": 2 * 3.550, "epsilon": 0.43},
        "NH2": {"index": 6, "charge": +1, "sigma": 2 * 2.99, "epsilon": 0.38},
        "N2": {"index": 7, "charge": 0, "sigma": 2 * 3.684, "epsilon": 0.31},
    },
    "pH properties": {"pK1": 8.18, "pK2": 10.02, "NUM_PHS": 12, "pHmin": 4.0, "pHmax": 12.0},
    "Montecarlo properties": {"N_BLOCKS": 16, "DESIRED_BLOCK_SIZE": 200, "PROB_REACTION": 0.7},
    "Simulation configuration": {
        "time step": 0.001,
        "USE_WCA": True,
        "USE_ELECTROSTATICS": True,
        "USE_FENE": False,
        "USE_BENDING": True,
        "USE_DIHEDRAL_POT": True,
        "USE_P3M": False,
    },
}


def example_parameters(n_mon=100, c_salt=2e-2, n_poly=1, **overrides):
    """The reference's example dictionary for `n_poly` chains of `n_mon` beads; the site counts
    (per chain, as Properties reads them) follow the every-third-bead pattern of M.py:126-155."""
    p = copy.deepcopy(EXAMPLE_PARAMETERS)
    role = topology.chain_site_pattern(n_mon)
    p["System properties"]["Number of polymers"] = int(n_poly)
    p["System properties"]["Number of monomers"] = int(n_mon)
    p["System properties"]["NH monomers"] = int(np.sum(role == 1))
    p["System properties"]["NH2 monomers"] = int(np.sum(role == 2))
    p["System properties"]["Concentration salt particles"] = float(c_salt)
    for section, values in overrides.items():
        p[section].update(values)
    return p


def reduced_units(params):
    s = params["System properties"]
    return units.reduce_parameters(
        params["Thermal properties"]["Temperature"], s["Concentration acid particles"],
        s["Concentration salt particles"], s["Number of polymers"], s["Number of monomers"],
        s["NH monomers"], s["NH2 monomers"], s["bond length"], s["bond constant"],
        s["bending constant"], s["dihedral constant"], params["Particles properties"])


def num_samples(params):
    """P.py:360-368."""
    mc = params["Montecarlo properties"]
    n = int(mc["N_BLOCKS"] * mc["DESIRED_BLOCK_SIZE"] / mc["PROB_REACTION"])
    return n, int((n - 10) / 2)


class Workload:
    """A cgpe_config + start state + sampling schedule for R replicas of a box with
    "Number of polymers" chains (+ explicit ions on request)."""

    def __init__(self, params, pH, c_salt=None, replica_offset=0, seed=12, device=-1, deprotonated=True,
                 explicit_ions=False, ion_seed=10, min_distance=None, nanoparticle=None):
        """nanoparticle: dict(charge=, sigma=[Angstrom, as in the particle table; 34.8 = 10 sim_length],
        epsilon=[kJ/mol], offset=(dx,dy,dz) [sim_length]) puts one fixed sphere of type "NP" at box centre +
        offset (BASELINE config 4; README.md:14-15 of the reference)."""
        self.params = params
        sp = params["System properties"]
        sc = params["Simulation configuration"]
        pH = np.atleast_1d(np.asarray(pH, dtype=np.float64))
        R = pH.size
        self.ru = ru = reduced_units(params)
        n_mon, n_poly = sp["Number of monomers"], sp["Number of polymers"]
        tix = ru.type_index
        c_salt_r = np.full(R, sp["Concentration salt particles"]) if c_salt is None else np.broadcast_to(
            np.asarray(c_salt, np.float64), (R,)).copy()
        if nanoparticle is not None:   # an extra row of the particle table, in the table's own units (Angstrom)
            params = copy.deepcopy(params)
            pp = params["Particles properties"]
            pp["NP"] = {"index": max(v["index"] for v in pp.values()) + 1, "charge": nanoparticle["charge"],
                        "sigma": nanoparticle.get("sigma", 34.8), "epsilon": nanoparticle.get("epsilon", 0.3)}
            self.params = params
            self.ru = ru = reduced_units(params)
            tix = ru.type_index
        self.pH = pH
        self.kappa = np.array([units.kappa_per_nm(c) for c in c_salt_r]) * ru.sim_length_nm
        self.r_cut = 1.0 / self.kappa                                   # M.py:106
        eps, sig = topology.wca_tables(tix, ru.sigma, ru.epsilon)
        reactions = [
            dict(type_prot=tix["NH"], type_deprot=tix["N"], type_ion=tix["B"], seed=77,
                 q_prot=ru.charge["NH"], q_deprot=ru.charge["N"], q_ion=ru.charge["B"],
                 pK=params["pH properties"]["pK1"], kT=1.0,
                 exclusion_range=topology.combination_sigma(ru.sigma["N"], ru.sigma["B"])),
            dict(type_prot=tix["NH2"], type_deprot=tix["N2"], type_ion=tix["B"], seed=77,
                 q_prot=ru.charge["NH2"], q_deprot=ru.charge["N2"], q_ion=ru.charge["B"],
                 pK=params["pH properties"]["pK2"], kT=1.0,
                 exclusion_range=topology.combination_sigma(ru.sigma["N2"], ru.sigma["B"])),
        ]
        # explicit ions (M.py:173-198): n_acid B+ and Cl-, n_salt Na+ and Cl-, plus one spare hidden
        # slot per site so that every site can release its H+
        n_acid = n_poly * (sp["NH monomers"] + sp["NH2 monomers"])
        ion_names = (["B"] * n_acid + ["Cl"] * n_acid + ["Na"] * ru.n_salt + ["Cl"] * ru.n_salt) if explicit_ions else []
        n_spare = n_acid if explicit_ions else 0
        if nanoparticle is not None:
            if explicit_ions:
                raise NotImplementedError("nanoparticle workload: implicit ions (an inserted H+ would need the "
                                          "sphere's WCA energy, which the explicit-ion move does not evaluate)")
            for rx in reactions:
                rx["type_ion"] = -1          # implicit-ion reactions; the slot holds a spectator
        self.n_ion_slots = len(ion_names) + n_spare + (1 if nanoparticle is not None else 0)
        self.cfg = _abi.make_config(
            n_replicas=R, replica_offset=replica_offset, n_chains=n_poly, n_beads=n_mon, n_ion_slots=self.n_ion_slots,
            n_types=len(tix), bond_kind=_abi.BOND_FENE if sc["USE_FENE"] else _abi.BOND_HARMONIC,
            use_angle=int(sc["USE_BENDING"]), use_dihedral=int(sc["USE_DIHEDRAL_POT"]), dihedral_mult=3,
            use_wca=int(sc["USE_WCA"]), use_dh=int(sc["USE_ELECTROSTATICS"]),
            bond_k=ru.k_bond, bond_r0=ru.bond_length, fene_drmax=4.0,
            angle_k=ru.k_bending, angle_phi0=math.pi * 2.0 / 3.0,
            dihedral_k=ru.k_dihedral, dihedral_phase=math.pi,
            dh_prefactor=ru.coulomb_prefactor, box_l=ru.box_length, time_step=sc["time step"],
            pH=float(pH[0]), kappa=float(self.kappa[0]), r_cut=float(self.r_cut[0]), kT=1.0, gamma=1.0,
            wca_eps=eps, wca_sig=sig, reactions=reactions, device=device,
            fixed_type_mask=(1 << tix["NP"]) if nanoparticle is not None else 0)
        # topology + start state: every site starts deprotonated (M.py:130-147)
        role = np.tile(topology.chain_site_pattern(n_mon), n_poly)
        if deprotonated:
            names = np.array(["CH2", "N", "N2"])[role]
        else:
            names = np.array(["CH2", "NH", "NH2"])[role]
        self.type = np.array([tix[n] for n in names], dtype=np.uint8)
        self.q = np.array([ru.charge[n] for n in names], dtype=np.float32)
        min_distance = 0.9 * ru.bond_length if min_distance is None else min_distance
        if n_poly == 1:
            pos = topology.linear_polymer_positions(1, n_mon, bond_length=ru.bond_length, seed=seed,
                                                    min_distance=min_distance, box_l=None)[0]
            pos += 0.5 * ru.box_length - pos.mean(axis=0)
        else:   # chains start at random points of the box (M.py:116-122), coordinates unfolded
            pos = topology.linear_polymer_positions(n_poly, n_mon, bond_length=ru.bond_length, seed=seed,
                                                    min_distance=min_distance, box_l=ru.box_length).reshape(-1, 3)
        self.xyzq = np.concatenate([pos, self.q[:, None]], axis=1).astype(np.float32)
        if explicit_ions:
            rng = np.random.RandomState(ion_seed)          # np.random.seed(10), P.py:376
            ion_pos = rng.random_sample((self.n_ion_slots, 3)) * ru.box_length
            ion_type = np.array([tix[n] for n in ion_names] + [len(tix)] * n_spare, dtype=np.uint8)
            ion_q = np.array([ru.charge[n] for n in ion_names] + [0.0] * n_spare, dtype=np.float32)
            self.xyzq = np.concatenate([self.xyzq, np.concatenate([ion_pos, ion_q[:, None]], axis=1).astype(np.float32)])
            self.type = np.concatenate([self.type, ion_type])
            self.q = np.concatenate([self.q, ion_q])
        if nanoparticle is not None:
            if "offset" in nanoparticle:
                np_pos = 0.5 * ru.box_length + np.asarray(nanoparticle["offset"], np.float64)
            else:   # beside the chain: the nearest point along +x that clears every bead's WCA core
                clear = nanoparticle.get("clearance", 1.3) * 0.25 * (ru.sigma["NP"] + max(ru.sigma.values()))
                np_pos = np.full(3, 0.5 * ru.box_length)
                while np.min(np.linalg.norm(self.xyzq[:, :3] - np_pos, axis=1)) < clear:
                    np_pos[0] += 0.25
            self.xyzq = np.concatenate([self.xyzq, np.array([[*np_pos, ru.charge["NP"]]], np.float32)])
            self.type = np.concatenate([self.type, np.array([tix["NP"]], np.uint8)])
            self.q = np.concatenate([self.q, np.array([ru.charge["NP"]], np.float32)])
        self.n_sites = (int(np.sum(role == 1)), int(np.sum(role == 2)))   # all chains
        ns, nsi = num_samples(params)
        p = params["Montecarlo properties"]["PROB_REACTION"]
        self.sample_kwargs = dict(
            md_steps_per_burst=500, use_md=bool(sc["USE_WCA"]),
            mc_blocks=((0, 5 * self.n_sites[0] + 1), (1, 5 * self.n_sites[1] + 1)),   # E.py:94-95
            q_snapshot_every=nsi, schedule_seed=10, p_burst1=p, p_burst2=p * 2.0 / (n_mon - 2))
        self.num_samples = ns

    @property
    def R(self):
        return self.cfg.n_replicas

    def init_engine(self, engine, thermostat_seed=1434):
        """Load the start state and per-replica parameters into a freshly created engine."""
        engine.set_state(self.xyzq, None, self.type)
        engine.set_replica_params(pH=self.pH, kappa=self.kappa, r_cut=self.r_cut)
        engine.set_thermostat_seed(thermostat_seed)                     # E.py:62
        return engine

    def sample_params(self, engine, n_samples):
        kw = dict(self.sample_kwargs)
        every = kw["q_snapshot_every"]
        rows = (n_samples + every - 1) // every if every > 0 else 0
        return engine.sample_params(n_samples, q_snapshot_rows=rows, **kw)


def config_c1(**kw):
    """C1: single chain N=50, Debye-Hueckel at 10 mM, pH = pK1."""
    return Workload(example_parameters(50, c_salt=1e-2), pH=[8.18], **kw)


def config_c2(**kw):
    """C2: 64 pH points x N=100 chain (examples/run_test.py scale)."""
    return Workload(example_parameters(100, c_salt=2e-2), pH=np.linspace(4.0, 12.0, 64), **kw)


C3_SALT_MOLAR = (1e-3, 2e-3, 5e-3, 1e-2, 2e-2, 5e-2, 1e-1, 2e-1)


def config_c4(charge=-20.0, n_ph=64, **kw):
    """C4: the C2 chain next to one fixed charged sphere (sigma = 10 sim_length) that interacts with
    every bead through WCA and Debye-Hueckel; the chain starts beside the sphere."""
    return Workload(example_parameters(100, c_salt=2e-2), pH=np.linspace(4.0, 12.0, n_ph),
                    nanoparticle=dict(charge=charge, sigma=34.8, epsilon=0.3), **kw)


def config_c3_shard(rank=0, world=8, n_ph=128, n_mon=1000, all_salts=False, **kw):
    """C3: 128 pH x 8 ionic strengths of N=1000 chains; rank `rank` of `world` holds the
    contiguous block of replicas [rank*R/world, (rank+1)*R/world) of the pH-major grid, i.e. with
    world=8 one ionic strength per GPU.  With world < 8 the grid shrinks to `world` ionic
    strengths (weak scaling: 128 replicas per GPU) unless all_salts (strong scaling: the whole
    grid of 8 x n_ph replicas whatever the number of GPUs)."""
    salts = C3_SALT_MOLAR if (all_salts or world > 8) else C3_SALT_MOLAR[:world]
    n_total = n_ph * len(salts)
    per = n_total // world
    lo = rank * per
    idx = np.arange(lo, lo + per)
    pH_grid = np.linspace(2.0, 12.0, n_ph)
    if all_salts and n_ph & (n_ph - 1) == 0 and n_ph > 1:
        # Strong scaling: which grid cell a global replica index (= Philox key) names is a convention, independent of the
        # number of GPUs: ionic strength = g mod 8, pH point = bit-reversed(g div 8).  Any contiguous block of indices a rank
        # holds then covers every ionic strength and the whole pH axis evenly, so the ranks cost the same (the salt-major order
        # gave rank 0 the four lowest ionic strengths, the expensive half: 1.72x on two GPUs).
        bits = n_ph.bit_length() - 1
        slot = idx // len(salts)
        rev = np.zeros_like(slot)
        for b in range(bits):
            rev |= ((slot >> b) & 1) << (bits - 1 - b)
        pH = pH_grid[rev]
        c_salt = np.array(salts)[idx % len(salts)]
    else:
        pH = pH_grid[idx % n_ph]
        c_salt = np.array(salts)[idx // n_ph]
    return Workload(example_parameters(n_mon, c_salt=float(c_salt[0])), pH=pH, c_salt=c_salt,
                    replica_offset=int(lo), **kw)
