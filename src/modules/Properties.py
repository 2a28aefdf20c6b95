"""Properties — validated parameters, unit reduction and the System object.

Drop-in for the reference's src/modules/Properties.py (class `Properties(input_dict)` with the
same six-section dictionary, attribute names, setter semantics and derived quantities), without
its pint / espressomd / matplotlib imports: the unit algebra is in
coarse_grain_polyelectrolyte_b200.units and `self.system` is the espressomd-shaped facade over
the B200 engine.

Setter contract pinned by the reference's tests_package/test_properties.py:
  * numbers: TypeError unless int/float as declared, ValueError unless > 0   (P.py:10-22)
  * n_nh / n_nh2: additionally ValueError when larger than n_mon              (P.py:107-123)
  * USE_* flags: TypeError unless bool                                        (P.py:25-32)
"""
import os

import numpy as np

from coarse_grain_polyelectrolyte_b200 import espressomd, units


class _Positive:
    """Data descriptor: a strictly positive number of the given type(s) (P.py:10-22)."""

    def __init__(self, kinds, at_most=None):
        self.kinds = kinds if isinstance(kinds, tuple) else (kinds,)
        self.at_most = at_most      # name of another attribute that bounds this one from above

    def __set_name__(self, owner, name):
        self.name, self.slot = name, "_" + name

    def __get__(self, obj, owner=None):
        return self if obj is None else getattr(obj, self.slot)

    def __set__(self, obj, value):
        if not isinstance(value, self.kinds):
            wanted = " or ".join(k.__name__ for k in self.kinds)
            raise TypeError(f"{self.name}: a {wanted} is required.")
        if value <= 0:
            raise ValueError(f"{self.name}: a positive number is required")
        if self.at_most is not None and value > getattr(obj, "_" + self.at_most):
            raise ValueError(f"{self.name} cannot be greater than {self.at_most}")
        setattr(obj, self.slot, value)


class _Flag:
    """Data descriptor: a bool (P.py:25-32)."""

    def __set_name__(self, owner, name):
        self.name, self.slot = name, "_" + name

    def __get__(self, obj, owner=None):
        return self if obj is None else getattr(obj, self.slot)

    def __set__(self, obj, value):
        if not isinstance(value, bool):
            raise TypeError(f"{self.name}: a boolean is required.")
        setattr(obj, self.slot, value)


class Quantity(float):
    """A float that remembers its unit and answers `.magnitude` / `.to(unit)` for the handful of
    conversions the reference performs on Properties attributes (it used pint quantities)."""
    _SCALES = {}

    def __new__(cls, value, unit, scales=None):
        q = super().__new__(cls, value)
        q.units, q._scales = unit, scales or {}
        return q

    @property
    def magnitude(self):
        return float(self)

    def to(self, unit):
        if unit == self.units:
            return self
        if unit not in self._scales:
            raise ValueError(f"cannot convert {self.units} to {unit}")
        k = self._scales[unit]
        value = float(self) / k[1] if isinstance(k, tuple) else float(self) * k
        return Quantity(value, unit, {})

    def __format__(self, spec):
        return f"{format(float(self), spec)} {self.units}"


_REAL = (float, int)


class Properties:
    temperature = _Positive(_REAL)          # K
    n_poly = _Positive(int)
    c_acid = _Positive(_REAL)               # mol/L
    c_salt = _Positive(_REAL)               # mol/L
    n_mon = _Positive(int)
    n_nh = _Positive(int, at_most="n_mon")
    n_nh2 = _Positive(int, at_most="n_mon")
    bond_l = _Positive(_REAL)               # Angstrom
    k_bond = _Positive(_REAL)               # kT / Angstrom^2
    k_bond_a = _Positive(_REAL)             # kT / degree^2
    k_dihedral = _Positive(_REAL)           # kT / degree^2
    n_blocks = _Positive(int)
    block_size = _Positive(int)
    probability_reaction = _Positive(_REAL)
    USE_WCA = _Flag()
    USE_ELECTROSTATICS = _Flag()
    USE_FENE = _Flag()
    USE_BENDING = _Flag()
    USE_DIHEDRAL_POT = _Flag()
    USE_P3M = _Flag()

    # (section, key) -> attribute, in the reference's reading order (P.py:378-433)
    _FIELDS = (
        ("Thermal properties", "Temperature", "temperature"),
        ("System properties", "Number of polymers", "n_poly"),
        ("System properties", "Concentration acid particles", "c_acid"),
        ("System properties", "Concentration salt particles", "c_salt"),
        ("System properties", "Number of monomers", "n_mon"),
        ("System properties", "NH monomers", "n_nh"),
        ("System properties", "NH2 monomers", "n_nh2"),
        ("System properties", "bond length", "bond_l"),
        ("System properties", "bond constant", "k_bond"),
        ("System properties", "bending constant", "k_bond_a"),
        ("System properties", "dihedral constant", "k_dihedral"),
        ("pH properties", "pK1", "pK"),
        ("pH properties", "pK2", "pK2"),
        ("pH properties", "NUM_PHS", "NUM_PHS"),
        ("pH properties", "pHmin", "pHmin"),
        ("pH properties", "pHmax", "pHmax"),
        ("Montecarlo properties", "N_BLOCKS", "n_blocks"),
        ("Montecarlo properties", "DESIRED_BLOCK_SIZE", "block_size"),
        ("Montecarlo properties", "PROB_REACTION", "probability_reaction"),
        ("Simulation configuration", "time step", "time_step"),
        ("Simulation configuration", "USE_WCA", "USE_WCA"),
        ("Simulation configuration", "USE_ELECTROSTATICS", "USE_ELECTROSTATICS"),
        ("Simulation configuration", "USE_FENE", "USE_FENE"),
        ("Simulation configuration", "USE_BENDING", "USE_BENDING"),
        ("Simulation configuration", "USE_DIHEDRAL_POT", "USE_DIHEDRAL_POT"),
        ("Simulation configuration", "USE_P3M", "USE_P3M"),
    )

    verbose = os.environ.get("CGPE_QUIET", "") == ""

    def __init__(self, input_dict, n_replicas=1, replica_offset=0, device=-1):
        self.input = input_dict
        self._n_replicas, self._replica_offset, self._device = int(n_replicas), int(replica_offset), device
        self._read_parameters()
        self._set_units()
        self._simulation_settings()
        self._system_definition()

    # -- P.py:378-433 ------------------------------------------------------------------------------
    def _read_parameters(self):
        for section, key, attr in self._FIELDS:
            setattr(self, attr, self.input[section][key])
        self.particle_properties = self.input["Particles properties"]

    # -- P.py:244-356 ------------------------------------------------------------------------------
    def _set_units(self):
        ru = units.reduce_parameters(
            self.temperature, self.c_acid, self.c_salt, self.n_poly, self.n_mon, self.n_nh, self.n_nh2,
            self.bond_l, self.k_bond, self.k_bond_a, self.k_dihedral, self.particle_properties)
        self.reduced = ru
        sl = ru.sim_length_nm
        length = lambda nm: {"nm": 1.0, "sim_length": 1.0 / sl, "angstrom": 10.0}          # noqa: E731
        self.KT = Quantity(ru.kT_joule, "joule", {"sim_energy": ("per", ru.kT_joule)})
        self.bjerrum_length = Quantity(ru.bjerrum_length_nm, "nm", length(1))
        self.n_acid = ru.n_acid
        self.box_volume = Quantity(ru.box_volume_nm3, "nm^3", {"sim_length^3": 1.0 / sl ** 3})
        self.box_length = Quantity(ru.box_length_nm, "nm", length(1))
        self.box_volume_reduced_units = Quantity(ru.box_volume, "sim_length^3")
        self.box_length_reduced_units = ru.box_length
        self.n_salt = ru.n_salt
        self.c_acid_unitless = float(self.c_acid)
        self.c_salt_unitless = float(self.c_salt)
        self.bond_length_reduced_units = Quantity(ru.bond_length, "sim_length", {"nm": sl, "angstrom": 10.0 * sl})
        self.k_bond_reduced_units = Quantity(ru.k_bond, "sim_energy/sim_length^2")
        self.k_bending_reduced_units = Quantity(ru.k_bending, "sim_energy/radian^2")
        self.k_dihedral_reduced_units = Quantity(ru.k_dihedral, "sim_energy/radian^2")
        self.kappa = Quantity(ru.kappa_per_nm, "1/nm", {"1/sim_length": sl})
        self.kappa_reduced_units = ru.kappa
        self.coulomb_prefactor_reduced_units = ru.coulomb_prefactor
        self.type_particles = dict(ru.type_index)
        self.charge_reduced_units = dict(ru.charge)
        self.sigma_reduced_units = dict(ru.sigma)
        self.epsilon_reduced_units = dict(ru.epsilon)
        if self.verbose:
            self._print_summary()

    def _print_summary(self):
        ru = self.reduced
        rows = [
            ("temperature", f"{self.temperature:.3f} K"), ("kT", f"{ru.kT_joule:.4g} J"),
            ("c_acid", f"{self.c_acid:.3g} mol/L"), ("c_salt", f"{self.c_salt:.3g} mol/L"),
            ("monomers per chain", self.n_mon), ("titratable sites", self.n_acid), ("salt ion pairs", self.n_salt),
            ("bond length", f"{ru.bond_length:.3f} sim_length"), ("bond constant", f"{ru.k_bond:.3f} kT/sim_length^2"),
            ("bending constant", f"{ru.k_bending:.3f} kT/rad^2"), ("dihedral constant", f"{ru.k_dihedral:.3f} kT/rad^2"),
            ("kappa", f"{ru.kappa_per_nm:.3f} 1/nm = {ru.kappa:.3f} 1/sim_length"),
            ("Debye length", f"{1.0 / ru.kappa_per_nm:.2f} nm = {1.0 / ru.kappa:.2f} sim_length"),
            ("Bjerrum length", f"{ru.bjerrum_length_nm:.3g} nm = {ru.bjerrum_length_nm / ru.sim_length_nm:.3g} sim_length"),
            ("box length", f"{ru.box_length_nm:.3g} nm = {ru.box_length:.3g} sim_length"),
        ]
        print("Simulation parameters:")
        for k, v in rows:
            print(f"\t{k}: {v}")
        print(f"\n{'type':<8}{'sigma':>12}{'epsilon':>12}{'charge':>10}   (reduced units)")
        for name in self.particle_properties:
            print(f"{name:<8}{ru.sigma[name]:>12.3f}{ru.epsilon[name]:>12.3f}{ru.charge[name]:>10.1f}")

    # -- P.py:358-369 ------------------------------------------------------------------------------
    def _simulation_settings(self):
        self.num_samples = int(self.n_blocks * self.block_size / self.probability_reaction)
        self.sample_iteration_size_capture = 2
        self.sample_begin_capture = 10
        self.n_samp_iter = int((self.num_samples - self.sample_begin_capture) / self.sample_iteration_size_capture)
        self.number_integration_steps = 500

    # -- P.py:371-376 ------------------------------------------------------------------------------
    def _system_definition(self):
        self.system = espressomd.System(box_l=[self.box_length_reduced_units] * 3, n_replicas=self._n_replicas,
                                        replica_offset=self._replica_offset, device=self._device)
        self.system.time_step = self.time_step
        self.system.cell_system.skin = 0.1
        self.system.periodicity = [True, True, True]
        np.random.seed(seed=10)
