"""Anchor to the only numbers ESPResSo itself produced for this path: the saved output of the reference's
notebook (/root/reference/src/modules/.ipynb_checkpoints/run_test-checkpoint.ipynb, raw JSON lines 232-254;
configuration lines 32-81; quoted below so that nothing is read from /root/reference at run time).

    N = 52 chain (16 NH + 2 NH2 sites), explicit B+/Na+/Cl- ions = 192 particles, 10 mM salt, Debye-Hueckel,
    harmonic bonds + bending, no dihedral, dt = 1e-4.  Notebook output (older revision of the loop: n_nh + 1
    trials per sample instead of 5 n_nh + 1 -- same equilibrium):

      pH  4: <n_N> 0.02 of 16 (ideal 0.00), <n_N2> 0.00 of 2; E_kin 259.9, E_coul +3.0, E_bonded 54.1, E_nb 0.004, T 0.996
      pH 10: <n_N> 15.61 of 16 (ideal 15.76), <n_N2> 1.04 of 2 (ideal 0.98), <n_B> 16.65 = <n_N> + <n_N2>;
             E_kin 287.1, E_coul -4.9, E_bonded 55.8, E_nb 0.31, T 1.007
      warm-up: final min distance 0.8212, final kT 0.99; T_inst between 0.90 and 1.01 over the printed frames

WHICH CODE PRODUCED THOSE NUMBERS.  Not the current src/modules: the saved output prints energies every 1000
samples, which only the older revision kept beside the notebook does (.ipynb_checkpoints/EnsembleDynamics-
checkpoint.py:97-123; its frames at time 0.05 and 31.15 are samples 0 and 1000 of 500-step bursts at dt = 1e-4,
p = 0.6: 1000 x 0.6 x 1.04 x 0.05 = 31.2).  That revision hard-codes the bond rest length r_0 = 1.5 sim_length
(Model-checkpoint.py:36-38; the notebook's "Final min distance warm up 0.8212" is impossible with the current
revision's r_0 = 1.5 Angstrom = 0.43: bonded neighbours would sit at ~0.65), reads the bond constant in
kcal/(mol A^2) and the bending constant in kcal/mol (Properties-checkpoint.py:250-256,295-303), and attempts
n_nh + 1 / n_nh2 + 1 trials per sample.  The dictionary below is the notebook's with those three entries
re-expressed in the current revision's units so that the *physical* parameters are the ones ESPResSo ran:
r_0 = 1.5 sim_length = 5.2219 A, k = 300 kcal/(mol A^2) = 503.22 kT/A^2, bend = 50 kcal/mol = 83.87 kT/rad^2 =
0.025548 kT/deg^2.  (The current revision's 5 n + 1 trials per sample sample the same equilibrium.)

What is asserted (tolerances stated per quantity): mean site counts within their binning error bar + 10 %;
the bookkeeping identity N_B = N_N + N_N2 sample by sample; E_kin = 3/2 kT per *present* particle (the
notebook's 259.9 vs 287.1 is exactly the 18 H+ that are absorbed at pH 4: 174 vs ~190 particles);
E_bonded = kT/2 per bond and angle (101 terms -> 50.5 +- 10 %; the notebook's four frames 54.1-56.3 must be
plausible draws of our frame distribution); sign of E_coul, and -- on the GPU, where the run is long enough -- its
size: the Debye-Hueckel energy of the charged chain follows its extension (the start is a random walk with 90 degree
angles that the bending term straightens over ~80 time units: Rg 5.7 -> 8.5, E_coul +24 -> +1..+7 kT, measured in
profiles/r2_notebook_anchor_trajectory.txt), so the notebook's +3.0 kT at Rg = 7.98 is compared with OUR frames of the
same extension (Rg 7.3..8.7).  Rg / Ree themselves depend on ESPResSo's own random start and are not asserted.

The same driver code (src/modules/EnsembleDynamics.py, ION_MODEL "explicit") runs on the CPU oracle
(not gpu: reduced sample count) and on the CUDA engine (gpu: the full 1333-sample protocol).
"""
import copy
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "src")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
os.environ.setdefault("CGPE_QUIET", "1")

# CODATA-2018 (what pint's default registry holds): l_B/2 at 300 K, eps_r = 80 (P.py:251-265), kcal/mol in kT
SIM_LENGTH_ANGSTROM = 3.4812713
KCAL_PER_MOL_IN_KT = 4184.0 / (6.02214076e23 * 1.380649e-23 * 300.0)

# notebook cell 0 (NB:32-81)
NOTEBOOK_PARAMETERS = {
    "Thermal properties": {"Temperature": 300},
    "System properties": {
        "Number of polymers": 1, "Concentration acid particles": 1e-2, "Concentration salt particles": 1e-2,
        "Number of monomers": 52, "NH monomers": 16, "NH2 monomers": 2,
        # older-revision physics in current-revision units (see the module docstring); notebook: 1.5, 300, 50.0, 100.
        "bond length": 1.5 * SIM_LENGTH_ANGSTROM, "bond constant": 300 * KCAL_PER_MOL_IN_KT,
        "bending constant": 50.0 * KCAL_PER_MOL_IN_KT * (np.pi / 180.0) ** 2, "dihedral constant": 100.0,
    },
    "Particles properties": {
        "NH": {"index": 0, "charge": +1, "sigma": 2 * 2.75, "epsilon": 0.0495},
        "N": {"index": 1, "charge": 0, "sigma": 2 * 3.298, "epsilon": 1.7504},
        "B": {"index": 2, "charge": +1, "sigma": 2 * 2.42, "epsilon": 0.0726},
        "CH2": {"index": 3, "charge": 0, "sigma": 2 * 3.800, "epsilon": 0.0663},
        "Na": {"index": 4, "charge": +1, "sigma": 2 * 2.21, "epsilon": 0.415},
        "Cl": {"index": 5, "charge": -1, "sigma": 2 * 3.550, "epsilon": 0.276},
        "NH2": {"index": 6, "charge": +1, "sigma": 2 * 2.99, "epsilon": 0.0389},
        "N2": {"index": 7, "charge": 0, "sigma": 2 * 3.684, "epsilon": 1.6398},
    },
    "pH properties": {"pK1": 8.18, "pK2": 10.02, "NUM_PHS": 3, "pHmin": 4.0, "pHmax": 10.0},
    "Montecarlo properties": {"N_BLOCKS": 16, "DESIRED_BLOCK_SIZE": 50, "PROB_REACTION": 0.6},
    "Simulation configuration": {
        "time step": 0.0001, "USE_WCA": True, "USE_ELECTROSTATICS": True, "USE_FENE": False,
        "USE_BENDING": True, "USE_DIHEDRAL_POT": False, "USE_P3M": False,
        "ION_MODEL": "explicit",
    },
}

# notebook output NB:232-254
NB = {
    4.0: dict(n_N=0.02, n_N2=0.00, e_kin=259.874, e_coul=2.999, e_bonded=54.144, e_nb=0.004, n_particles=174),
    10.0: dict(n_N=15.61, n_N2=1.04, n_B=16.65, e_kin=287.092, e_coul=-4.935, e_bonded=55.774, e_nb=0.312),
}


def binning_sem(x, n_bins=16):
    """NB:727-731: std(bin means, ddof=1.5) / sqrt(n_bins)."""
    x = np.asarray(x, float)
    n = (len(x) // n_bins) * n_bins
    means = x[:n].reshape(n_bins, -1).mean(axis=1)
    return float(np.std(means, ddof=1.5) / np.sqrt(n_bins))


def run_notebook_protocol(engine_factory, n_chunks, samples_per_chunk, equil_steps, n_streams=1):
    """The notebook's cells 0-4 for the pH 4 and pH 10 points as replicas of one batched run (n_streams independent
    copies of each).  The sampling loop runs in chunks with one frame (analysis.energy(), calc_rg as the older revision
    printed them) behind each."""
    from coarse_grain_polyelectrolyte_b200 import espressomd
    from modules.EnsembleDynamics import EnsembleDynamics
    old = espressomd.System.engine_factory
    espressomd.System.engine_factory = engine_factory
    try:
        ens = EnsembleDynamics(copy.deepcopy(NOTEBOOK_PARAMETERS), replicas=dict(pH=[4.0, 10.0] * n_streams))
        ens.verbose = False
        assert ens.n_acid == 18 and ens.n_salt == 52
        assert len(ens.system.part) >= 192                                # 52 + 2*18 + 2*52 (+ hidden spares)
        assert abs(ens.bond_length_reduced_units.magnitude - 1.5) < 1e-4 and abs(ens.k_bending_reduced_units.magnitude - 83.87) < 0.02
        ens.equilibrate_pH()
        ens.system.integrator.run(equil_steps)                             # NB cell 2 runs 1e5 steps here
        ens.num_samples = int(samples_per_chunk)
        ens.n_samp_iter = max(1, (ens.num_samples - 10) // 2)
        series = {k: [] for k in ("num_As", "num_As2", "num_B", "rad_gyr", "end_2end")}
        frames = []
        for _ in range(n_chunks):
            ens.perform_sampling()
            for k in series:
                series[k].append(np.asarray(getattr(ens, k)))
            e = ens.system.analysis.energy()
            n_free_ion = ens.system.number_of_particles(type=ens.type_particles["B"])
            rg = ens.system.analysis.calc_rg(chain_start=0, number_of_chains=1, chain_length=52)[0]
            frames.append([e["kinetic"], e["coulomb"], e["bonded"], e["non_bonded"], np.asarray(n_free_ion, float), np.asarray(rg, float)])
        return {k: np.concatenate(v, axis=1) for k, v in series.items()}, np.array(frames)
    finally:
        espressomd.System.engine_factory = old


def check_against_notebook(series, frames, coulomb_size):
    n_N, n_N2, n_B = series["num_As"], series["num_As2"], series["num_B"]
    np.testing.assert_array_equal(n_B, n_N + n_N2)                         # NB:253-254, sample by sample
    skip = n_N.shape[1] // 5
    for r in range(n_N.shape[0]):
        pH = (4.0, 10.0)[r % 2]
        ref = NB[pH]
        for name, x, n_sites in (("n_N", n_N[r, skip:], 16), ("n_N2", n_N2[r, skip:], 2)):
            mean, sem = x.mean(), binning_sem(x)
            tol = 3.0 * sem + 0.10 * max(ref[name], 0.1) + 0.02 * n_sites  # error bar + 10 % (+ the resolution of the notebook's print)
            assert abs(mean - ref[name]) <= tol, f"pH {pH}: <{name}> = {mean:.3f} +- {sem:.3f}, notebook {ref[name]}"
        ek, ec, eb, enb, nb_free, rg = (frames[:, k, r] for k in range(6))
        # a hidden H+ is gone (M.py:256-261): the notebook's T = 2/3 E_kin / len(system.part) gives 0.996 with 174 particles
        # at pH 4 (all 18 H+ absorbed) and 1.007 with 190 at pH 10
        n_present = 192 - 18 + nb_free
        t_inst = (2.0 / 3.0) * ek / n_present
        assert np.all((t_inst > 0.75) & (t_inst < 1.25)), f"pH {pH}: T_inst frames {t_inst}"
        assert 0.94 <= t_inst.mean() <= 1.06, f"pH {pH}: T_inst {t_inst.mean():.3f}"   # NB:234-248: 0.90 .. 1.01 over four frames
        sem_k = ek.std(ddof=1) / np.sqrt(len(ek))
        assert abs(ek.mean() - ref["e_kin"]) <= 4 * sem_k + 0.05 * ref["e_kin"], (pH, ek.mean(), ref["e_kin"])
        # 51 bonds + 50 angles, kT/2 each = 50.5; the notebook's frames (54.1, 55.8 + the two at time 0.05: 55.6, 56.3)
        assert 0.9 * 50.5 <= eb.mean() <= 1.1 * 50.5, f"pH {pH}: E_bonded {eb.mean():.2f}"
        assert abs(ref["e_bonded"] - eb.mean()) <= 3.0 * max(eb.std(ddof=1), 5.0), (pH, eb.mean(), eb.std(ddof=1))
        assert np.sign(np.median(ec)) == np.sign(ref["e_coul"]), f"pH {pH}: E_coul {np.round(ec, 1)} vs notebook {ref['e_coul']}"
        assert 0.0 <= enb.mean() < 2.0                                     # notebook: 0.004 .. 0.36 kT
    if coulomb_size:
        # Size of the Debye-Hueckel energy, pooled over the streams of each pH.  The notebook holds ONE frame per pH; it must be
        # a plausible draw (within 2.5 standard deviations) of our frames under the same conditions: at pH 4 the frames in
        # which the chain has the notebook's extension (Rg 7.98 -> 7.3..8.7); at pH 10 (neutral chain: the energy is the ions'
        # alone) the last three quarters of the run.
        for pH, which in ((4.0, 0), (10.0, 1)):
            ec = frames[:, 1, which::2].ravel()
            rg = frames[:, 5, which::2].ravel()
            if pH == 4.0:
                sel = (rg > 7.3) & (rg < 8.7)
                assert sel.sum() >= 8, f"the chain never reached the notebook's extension: Rg frames {np.round(rg, 1)}"
            else:
                sel = np.tile(np.arange(frames.shape[0]) >= frames.shape[0] // 4, (frames.shape[2] // 2, 1)).T.ravel()
            mean, sd = ec[sel].mean(), max(ec[sel].std(ddof=1), 2.0)
            assert abs(NB[pH]["e_coul"] - mean) <= 2.5 * sd, f"pH {pH}: E_coul {mean:.2f} +- {sd:.2f} ({sel.sum()} frames) vs notebook {NB[pH]['e_coul']}"
            assert abs(mean) < 12.0


def test_oracle_reproduces_the_notebook_numbers(oracle_mod):
    """CPU: the restatement itself against ESPResSo's saved output.  Reduced: 300 samples after 2e4 steps (9 time units;
    the chain is still compact then, so the size of E_coul is left to the GPU run)."""
    oracle_mod.set_threads(2)
    series, frames = run_notebook_protocol(oracle_mod.OracleEngine, n_chunks=5, samples_per_chunk=60, equil_steps=20000)
    check_against_notebook(series, frames, coulomb_size=False)


@pytest.mark.gpu
def test_cuda_engine_reproduces_the_notebook_numbers():
    """GPU: cell 2's 1e5 equilibration steps, then twice the notebook's 1333 samples (83 time units) through the drop-in
    driver, two independent streams per pH."""
    series, frames = run_notebook_protocol(None, n_chunks=62, samples_per_chunk=43, equil_steps=100000, n_streams=2)
    assert series["num_As"].shape == (4, 2666)
    check_against_notebook(series, frames, coulomb_size=True)


@pytest.mark.gpu
def test_com_velocity_autocorrelation_matches_the_notebook_fit():
    """NB cell 28 (raw JSON ~ line 918 ff.): ESPResSo's COM velocity autocorrelation of this N = 52 chain, fitted with
    a exp(-b tau), gave a = 0.0608, b = 0.980.  The COM of a Langevin chain with one friction constant is an Ornstein-Uhlenbeck
    particle whatever the interactions: a = 3 kT / N = 0.0577, b = gamma = 1.  The correlator of Observables.com_vel_cor (fed by
    the on-device frame recorder inside the fused sampling loop) must give both within 6 %, and the notebook's fit within 10 %."""
    from modules.EnsembleDynamics import EnsembleDynamics
    from modules.Observables import Observables
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    from run_diffusion import fit_vacf, green_kubo
    ens = EnsembleDynamics(copy.deepcopy(NOTEBOOK_PARAMETERS), replicas=dict(pH=[4.0, 10.0, 4.0, 10.0]))
    reg = Observables()
    ids = list(range(52))
    reg.add_observable("com_vel_cor", reg.com_vel_cor, ids, 1000)          # NB cell 7
    reg.update_accumulators(ens.system)
    ens.equilibrate_pH()
    ens.num_samples = 2666
    ens.n_samp_iter = 1328
    ens.perform_sampling()
    assert ens.system._engine.last_call_ms()[1] == 1                         # the accumulators did not break the fused launch
    c = reg.observables["com_vel_cor"]
    tau, acf = c.lag_times(), c.result()[..., 0].mean(axis=1)               # mean over the four replicas
    assert c.sample_sizes()[0] > 7000
    a, b = fit_vacf(tau, acf, n=85)
    assert abs(a - 3.0 / 52) < 0.06 * 3.0 / 52 and abs(b - 1.0) < 0.06, (a, b)
    assert abs(a - 0.0608) < 0.10 * 0.0608 and abs(b - 0.980) < 0.10 * 0.980, (a, b)
    d_gk, _ = green_kubo(tau, acf)                                           # NB cell 29; Rouse: D = kT / (N gamma) = 0.0192
    assert abs(d_gk - 1.0 / 52) < 0.2 / 52, d_gk
