"""Unit reduction of the reference's parameter dictionary without pint.

Follows src/modules/Properties.py:244-311 of the reference (`_set_units`): energies in k_B T,
lengths in half a Bjerrum length, charges in e, times in ps.  The reference does this with a
`pint.UnitRegistry`; the same arithmetic is written out here with the CODATA-2018 values pint
ships (exact SI-2019 constants plus the fine-structure constant for the vacuum permittivity).
"""
import math
from dataclasses import dataclass, field

# SI 2019 exact
BOLTZMANN = 1.380649e-23          # J/K
ELEMENTARY_CHARGE = 1.602176634e-19  # C
AVOGADRO = 6.02214076e23          # 1/mol
PLANCK = 6.62607015e-34           # J s
SPEED_OF_LIGHT = 299792458.0      # m/s
# CODATA 2018, as in pint's default registry
FINE_STRUCTURE = 7.2973525693e-3
VACUUM_PERMITTIVITY = ELEMENTARY_CHARGE ** 2 / (2.0 * FINE_STRUCTURE * PLANCK * SPEED_OF_LIGHT)  # F/m

WATER_PERMITTIVITY = 80           # P.py:251
NM = 1e-9
ANGSTROM = 1e-10
LITRE_IN_NM3 = 1e24
DEGREE = math.pi / 180.0


@dataclass
class ReducedUnits:
    """Everything `Properties._set_units` derives, as plain floats (P.py:250-311)."""
    kT_joule: float
    bjerrum_length_nm: float
    sim_length_nm: float
    box_volume_nm3: float
    box_length_nm: float
    box_length: float            # reduced
    box_volume: float            # reduced
    n_acid: int
    n_salt: int
    bond_length: float
    k_bond: float
    k_bending: float
    k_dihedral: float
    kappa_per_nm: float
    kappa: float                 # reduced
    coulomb_prefactor: float     # l_B kT / e^2 in sim units (M.py:98-104) == 2 by construction
    charge: dict = field(default_factory=dict)
    sigma: dict = field(default_factory=dict)
    epsilon: dict = field(default_factory=dict)
    type_index: dict = field(default_factory=dict)


def bjerrum_length_nm(temperature):
    """e^2 / (4 pi eps0 eps_r k_B T)   (P.py:252-258)."""
    kT = temperature * BOLTZMANN
    return ELEMENTARY_CHARGE ** 2 / (4.0 * math.pi * VACUUM_PERMITTIVITY * WATER_PERMITTIVITY * kT) / NM


def kappa_per_nm(c_salt_molar):
    """Inverse Debye length of a 1:1 electrolyte in water, sqrt(c)/0.304 nm^-1 (P.py:294)."""
    return math.sqrt(c_salt_molar) / 0.304


def reduce_parameters(temperature, c_acid, c_salt, n_poly, n_mon, n_nh, n_nh2, bond_l, k_bond,
                      k_bond_a, k_dihedral, particles):
    kT = temperature * BOLTZMANN
    l_b = bjerrum_length_nm(temperature)
    sim_length = 0.5 * l_b                                        # P.py:263
    # box volume so that n_poly*n_mon monomers give c_acid (P.py:271-279)
    box_volume_nm3 = n_poly * n_mon / (AVOGADRO * c_acid) * LITRE_IN_NM3
    box_length_nm = box_volume_nm3 ** (1.0 / 3.0)
    # P.py:280 truncates with int(), it does not round.  The exact value is c_salt/c_acid*n_mon,
    # often an integer; depending on the order of the unit factors the float lands a few ulp
    # below it (199.99999999999997) and truncation would then lose an ion pair, so values within
    # 1e-9 of an integer are snapped to it first.
    n_salt_float = c_salt * (box_volume_nm3 / LITRE_IN_NM3) * AVOGADRO
    n_salt = int(n_salt_float * (1.0 + 1e-9))
    kappa_nm = kappa_per_nm(c_salt)
    ru = ReducedUnits(
        kT_joule=kT,
        bjerrum_length_nm=l_b,
        sim_length_nm=sim_length,
        box_volume_nm3=box_volume_nm3,
        box_length_nm=box_length_nm,
        box_length=box_length_nm / sim_length,
        box_volume=box_volume_nm3 / sim_length ** 3,
        n_acid=n_nh + n_nh2,                                      # P.py:270
        n_salt=n_salt,
        bond_length=bond_l * ANGSTROM / NM / sim_length,          # P.py:284,290
        k_bond=k_bond * (sim_length * NM / ANGSTROM) ** 2,        # kT/A^2 -> kT/sim_length^2 (P.py:285,291)
        k_bending=k_bond_a / DEGREE ** 2,                         # kT/deg^2 -> kT/rad^2 (P.py:286,292)
        k_dihedral=k_dihedral / DEGREE ** 2,                      # P.py:287,293
        kappa_per_nm=kappa_nm,
        kappa=kappa_nm * sim_length,                              # P.py:295
        coulomb_prefactor=l_b / sim_length,                       # M.py:98-104
    )
    kj_per_mol_in_kT = 1e3 / AVOGADRO / kT
    for name, p in particles.items():                             # P.py:307-311
        ru.type_index[name] = p["index"]
        ru.charge[name] = float(p["charge"])
        ru.sigma[name] = p["sigma"] * ANGSTROM / NM / sim_length
        ru.epsilon[name] = p["epsilon"] * kj_per_mol_in_kT
    return ru
