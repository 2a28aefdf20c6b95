"""The Properties drop-in against the contract of the reference's tests_package/test_properties.py:
same fixture dictionary, same three checks per attribute (current value, TypeError on a string,
ValueError on a negative; n_nh / n_nh2 also ValueError above n_mon).

The reference's value assertions contradict its own fixture for seven of nine attributes
(e.g. it asserts n_poly == 10 for a fixture that sets 1); here every expected value
is the fixture's.  Runs without espressomd, pint or matplotlib.
"""
import os

import pytest

os.environ.setdefault("CGPE_QUIET", "1")
from modules.Properties import Properties  # noqa: E402  (same import path as the reference's test)


@pytest.fixture(scope="module")
def props():
    input_dict = {
        "Thermal properties": {"Temperature": 300},
        "System properties": {
            "Number of polymers": 1, "Concentration acid particles": 1e-3, "Concentration salt particles": 1e-3,
            "Number of monomers": 22, "NH monomers": 6, "NH2 monomers": 2, "bond length": 1.5, "bond constant": 300,
            "bending constant": 0.01, "dihedral constant": 100.0,
        },
        "Particles properties": {
            "HA": {"index": 0, "charge": +1, "sigma": 2 * 2.6, "epsilon": 231},
            "A": {"index": 1, "charge": 0, "sigma": 2 * 3.172, "epsilon": 231},
            "B": {"index": 2, "charge": +1, "sigma": 2 * 2.958, "epsilon": 0.0726},
            "N": {"index": 3, "charge": 0, "sigma": 2 * 3.93, "epsilon": 56.0},
            "Na": {"index": 4, "charge": +1, "sigma": 2 * 3.9624, "epsilon": 0.738},
            "Cl": {"index": 5, "charge": -1, "sigma": 2 * 3.915, "epsilon": 0.305},
            "HA2": {"index": 6, "charge": +1, "sigma": 2 * 2.6, "epsilon": 231},
            "A2": {"index": 7, "charge": 0, "sigma": 2 * 3.172, "epsilon": 231},
        },
        "pH properties": {"pK1": 8.18, "pK2": 10.02, "NUM_PHS": 8, "pHmin": 2.5, "pHmax": 12.0},
        "Montecarlo properties": {"N_BLOCKS": 16, "DESIRED_BLOCK_SIZE": 100, "PROB_REACTION": 0.5},
        "Simulation configuration": {
            "time step": 0.001, "USE_WCA": False, "USE_ELECTROSTATICS": False, "USE_FENE": False,
            "USE_BENDING": False, "USE_DIHEDRAL_POT": False, "USE_P3M": False,
        },
    }
    return Properties(input_dict)


CASES = [
    ("temperature", 300.0, "300.0", -300.0),
    ("n_poly", 1, "10", -10),
    ("c_acid", 1e-3, "0.05", -0.05),
    ("c_salt", 1e-3, "0.1", -0.1),
    ("n_mon", 22, "20", -20),
    ("n_nh", 6, "3", -3),
    ("n_nh2", 2, "2", -2),
    ("bond_l", 1.5, "0.97 A", -0.97),
    ("k_bond", 300, "10.0 kcal/mol/A^2", -10.0),
    ("k_bond_a", 0.01, "x", -1.0),
    ("k_dihedral", 100.0, "x", -1.0),
    ("n_blocks", 16, "16", -1),
    ("block_size", 100, "100", -1),
    ("probability_reaction", 0.5, "0.5", -0.5),
]


@pytest.mark.parametrize("name,value,wrong_type,negative", CASES)
def test_validated_attribute(props, name, value, wrong_type, negative):
    assert getattr(props, name) == value
    with pytest.raises(TypeError):
        setattr(props, name, wrong_type)
    with pytest.raises(ValueError):
        setattr(props, name, negative)
    assert getattr(props, name) == value          # a rejected assignment leaves the value alone


@pytest.mark.parametrize("name", ["n_nh", "n_nh2"])
def test_site_counts_bounded_by_chain_length(props, name):
    with pytest.raises(ValueError):
        setattr(props, name, 30)                   # 30 > n_mon = 22  (T.py:113, T.py:125)


def test_integer_attributes_reject_floats(props):
    for name in ("n_poly", "n_mon", "n_nh", "n_nh2", "n_blocks", "block_size"):
        with pytest.raises(TypeError):
            setattr(props, name, 2.0)


@pytest.mark.parametrize("name", ["USE_WCA", "USE_ELECTROSTATICS", "USE_FENE", "USE_BENDING", "USE_DIHEDRAL_POT", "USE_P3M"])
def test_flags_require_bool(props, name):
    assert getattr(props, name) is False
    with pytest.raises(TypeError):
        setattr(props, name, 1)
    setattr(props, name, True)
    setattr(props, name, False)


def test_derived_quantities_of_the_fixture(props):
    """SURVEY.md Appendix B, test-fixture column."""
    assert props.n_acid == 8 and props.n_salt == 22
    assert props.box_length_reduced_units == pytest.approx(95.3132, rel=1e-5)
    assert props.kappa_reduced_units == pytest.approx(0.03621, rel=2e-4)
    assert props.bond_length_reduced_units.magnitude == pytest.approx(0.43088, rel=1e-4)
    assert props.k_bond_reduced_units.magnitude == pytest.approx(3635.77, rel=1e-5)
    assert props.k_bending_reduced_units.magnitude == pytest.approx(32.83, rel=1e-3)
    assert props.k_dihedral_reduced_units.magnitude == pytest.approx(328281, rel=1e-5)
    assert (props.num_samples, props.n_samp_iter, props.number_integration_steps) == (3200, 1595, 500)
    assert props.KT.to("sim_energy").magnitude == 1.0
    assert props.bjerrum_length.to("sim_length").magnitude == pytest.approx(2.0)
    assert props.type_particles["Cl"] == 5 and props.charge_reduced_units["Cl"] == -1.0
    assert len(props.system.part) == 0 and props.system.time_step == 0.001
    assert list(props.system.box_l) == [props.box_length_reduced_units] * 3


def test_example_run_parameters():
    """SURVEY.md Appendix B, examples/run_test.py column."""
    from coarse_grain_polyelectrolyte_b200 import workloads
    p = Properties(workloads.EXAMPLE_PARAMETERS)
    assert p.box_length_reduced_units == pytest.approx(73.2845, rel=1e-6)
    assert p.kappa_reduced_units == pytest.approx(0.16195, rel=1e-4)
    assert 1.0 / p.kappa_reduced_units == pytest.approx(6.1748, rel=1e-4)
    assert p.bond_length_reduced_units.magnitude == pytest.approx(1.10592, rel=1e-5)
    assert p.k_bond_reduced_units.magnitude == pytest.approx(2423.85, rel=1e-5)
    assert p.k_bending_reduced_units.magnitude == pytest.approx(6.5656, rel=1e-4)
    assert p.coulomb_prefactor_reduced_units == pytest.approx(2.0)
    assert p.n_acid == 34 and p.n_salt == 200             # c_salt/c_acid * n_mon, truncated (P.py:280)
    assert (p.num_samples, p.n_samp_iter) == (4571, 2280)
    assert p.sigma_reduced_units["NH"] == pytest.approx(5.5 / 3.48127, rel=1e-5)


def test_missing_key_is_a_key_error():
    from coarse_grain_polyelectrolyte_b200 import workloads
    import copy
    d = copy.deepcopy(workloads.EXAMPLE_PARAMETERS)
    del d["System properties"]["NH monomers"]
    with pytest.raises(KeyError):
        Properties(d)


def test_flat_import_path_works(tmp_path):
    """The reference's modules import each other flat (PYTHONPATH=src/modules, M.py:18-19)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = "import Properties, Model, EnsembleDynamics, utils, Observables; print(utils.ideal_alpha(8.18, 8.18))"
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "src", "modules"), root]), CGPE_QUIET="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "0.5"
