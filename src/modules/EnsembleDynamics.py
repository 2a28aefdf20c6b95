"""EnsembleDynamics — warm-up, pH equilibration and the sampling loop.

Drop-in for the reference's src/modules/EnsembleDynamics.py: `EnsembleDynamics(input_dict)`,
`.equilibrate_pH(pH)`, `.perform_sampling(npH)` and the result arrays `num_As`, `num_As2`,
`num_B`, `rad_gyr`, `end_2end`, `times`, `qdist`.  With `replicas=...` (see Model) every result
array gains a leading replica axis and `equilibrate_pH` accepts one pH per replica.

The loop body of perform_sampling (E.py:88-106: Bernoulli-scheduled MD bursts, two blocks of
constant-pH trial moves, observables) runs as ONE kernel launch for all samples and replicas
(`system.sample` -> cgpe_sample_loop) instead of nine engine calls per sample from Python.
"""
import time

import numpy as np

try:
    from .Model import Model
except ImportError:
    from Model import Model


class EnsembleDynamics(Model):
    def __init__(self, input_dict, replicas=None, replica_offset=0, device=-1):
        super().__init__(input_dict, replicas=replicas, replica_offset=replica_offset, device=device)
        if self.USE_WCA:
            self._warm_up_system()

    # -- E.py:24-65 ----------------------------------------------------------------------------------
    def _warm_up_system(self, max_rounds=60):
        say = print if self.verbose else (lambda *a, **k: None)
        say("\nWarm up the system:")
        warm_steps, min_dist = 25, 0.8
        system = self.system
        system.time = 0.0
        # Not in the reference: the start configuration has bonds of 1.5 (M.py:119) against r_0 ~ 1.1, so a
        # chain of N beads must contract by ~0.2 N length units.  The capped inertial dynamics below does that
        # for N <= 100 but pinches strands of long chains (min_dist then never recovers and the loop, which is
        # unbounded in the reference, cannot end).  A contracting steepest descent (gamma k_bond = 0.1, steps of
        # at most 0.02) first takes the bond strain out gently; it is a no-op for an already relaxed chain.
        # "REFERENCE_FAITHFUL": True skips it and lifts the bound on the rounds below: the reference's loop as written.
        if not self.reference_faithful:
            k_bond = self.k_bond_reduced_units.magnitude
            stretch = abs(1.5 - self.bond_length_reduced_units.magnitude)
            n_relax = int(min(60000, max(500, 30 * self.n_mon * stretch)))
            system.integrator.set_steepest_descent(f_max=0, gamma=0.1 / k_bond, max_displacement=0.02)
            system.integrator.run(n_relax)
            system.integrator.set_vv()
        else:
            max_rounds = None
        wca_cap = 5
        system.force_cap = wca_cap
        act_min_dist = np.min(system.analysis.min_dist())
        system.thermostat.set_langevin(kT=0.0, gamma=1.0, seed=4)
        rounds = 0
        # zero-temperature, force-capped relaxation until no two particles overlap (E.py:39-48)
        while act_min_dist < min_dist:
            for _ in range(warm_steps + wca_cap):
                system.integrator.run(int(0.5 * warm_steps))
            act_min_dist = np.min(system.analysis.min_dist())
            wca_cap += 1
            system.force_cap = wca_cap
            rounds += 1
            if max_rounds is not None and rounds >= max_rounds:
                # The reference loops here without bound.  Long chains start with bonds of 1.5 against
                # r_0 ~ 1.1 (M.py:119): while they contract, a pair of strands can end up pinched below
                # the threshold at T = 0 for good.  The temperature ramp below anneals such contacts.
                import warnings
                warnings.warn(f"warm-up left min_dist at {act_min_dist:.3f} (< {min_dist}) after {rounds} rounds; "
                              "continuing with the temperature ramp", RuntimeWarning)
                break
        say("Final min distance warm up", act_min_dist)
        system.force_cap = 0
        system.integrator.run(warm_steps)
        # temperature ramp 3.0 -> 1.0 (E.py:54-59)
        temp = 3.0
        while temp > 1.0:
            system.thermostat.set_langevin(kT=temp, gamma=1.0)
            system.integrator.run(warm_steps)
            temp -= 0.01
        say("Final temperature warm up", temp)
        system.thermostat.set_langevin(kT=self.KT.to("sim_energy").magnitude, gamma=1.0, seed=1434)
        system.integrator.run(warm_steps)
        say("Warm up finished")

    # -- E.py:67-73 ----------------------------------------------------------------------------------
    def equilibrate_pH(self, pH=None):
        if pH is None:
            pH = self.replica_pH if self.system.n_replicas > 1 else float(self.replica_pH[0])
        self.reaction1.constant_pH = pH
        self.reaction2.constant_pH = pH
        self.reaction1.reaction(reaction_steps=self.n_nh + 1)
        self.reaction2.reaction(reaction_steps=self.n_nh2 + 1)
        if self.USE_WCA:
            self.system.integrator.run(steps=1000)

    # -- E.py:75-119 ---------------------------------------------------------------------------------
    def perform_sampling(self, npH=None):
        system = self.system
        if len(system.auto_update_accumulators) > 0 and (not self.USE_WCA or system._frames_plan() is None):
            # accumulators the engine's frame recorder cannot feed (RDF, particle positions, ...) are updated by integrator.run:
            # take the reference's loop call by call.  The ones Observables.py registers for the chain (bond lengths, torsions,
            # COM position / velocity correlators) are recorded on the device and the loop below stays one launch.
            return self._perform_sampling_stepwise()
        system.time = 0.0
        start = time.time()
        p = self.probability_reaction
        obs, qdist = system.sample(
            n_samples=self.num_samples, md_steps_per_burst=self.number_integration_steps,
            p_burst1=p, p_burst2=p * (2.0 / (self.n_mon - 2)),                       # E.py:89,91
            mc_blocks=((self.reaction1, 5 * self.n_nh + 1), (self.reaction2, 5 * self.n_nh2 + 1)),   # E.py:94-95
            use_md=self.USE_WCA, q_snapshot_every=self.n_samp_iter, schedule_seed=10)
        squeeze = (lambda a: a[0]) if system.n_replicas == 1 else (lambda a: a)
        self.num_As = squeeze(obs[..., 0])           # E.py:97
        self.num_As2 = squeeze(obs[..., 1])          # E.py:98
        self.num_B = squeeze(obs[..., 2])            # E.py:99
        self.rad_gyr = squeeze(obs[..., 3])          # E.py:100
        self.end_2end = squeeze(obs[..., 4])         # E.py:101
        self.times = squeeze(obs[..., 5])            # E.py:102
        # the reference allocates (n_samp_iter, n_mon) and fills rows i // n_samp_iter (E.py:82,103-106)
        rows = np.zeros((system.n_replicas, self.n_samp_iter, self.n_mon))
        if qdist is not None:
            rows[:, :qdist.shape[1], :] = qdist
        self.qdist = squeeze(rows)
        self.elapsed_seconds = time.time() - start
        if self.verbose:
            print(f" Completed. Time elapsed: {self.elapsed_seconds:.2f} seconds.")

    def _perform_sampling_stepwise(self):
        """E.py:88-106 as written there: one engine call per line, the burst decisions drawn from numpy's
        global stream (seeded 10 in Properties, P.py:376) and therefore shared by all replicas."""
        system, R = self.system, self.system.n_replicas
        system.time = 0.0
        start = time.time()
        p = self.probability_reaction
        out = {k: np.zeros((R, self.num_samples)) for k in ("num_As", "num_As2", "num_B", "rad_gyr", "end_2end", "times")}
        qdist = np.zeros((R, self.n_samp_iter, self.n_mon))
        for i in range(self.num_samples):
            if self.USE_WCA:
                if np.random.random() < p:
                    system.integrator.run(self.number_integration_steps)
                if np.random.random() < p * (2.0 / (self.n_mon - 2)):
                    system.integrator.run(self.number_integration_steps)
            self.reaction1.reaction(reaction_steps=5 * self.n_nh + 1)
            self.reaction2.reaction(reaction_steps=5 * self.n_nh2 + 1)
            out["num_As"][:, i] = system.number_of_particles(type=self.type_particles["N"])
            out["num_As2"][:, i] = system.number_of_particles(type=self.type_particles["N2"])
            out["num_B"][:, i] = (system.number_of_particles(type=self.type_particles["B"]) if self.ion_model == "explicit"
                                  else out["num_As"][:, i] + out["num_As2"][:, i])
            out["rad_gyr"][:, i] = np.reshape(system.analysis.calc_rg(chain_start=0, number_of_chains=1, chain_length=self.n_mon)[0], -1)
            out["end_2end"][:, i] = np.reshape(system.analysis.calc_re(chain_start=0, number_of_chains=1, chain_length=self.n_mon)[0], -1)
            out["times"][:, i] = system.time
            if i % self.n_samp_iter == 0 and i // self.n_samp_iter < self.n_samp_iter:
                qdist[:, i // self.n_samp_iter, :] = np.reshape(system.part.by_ids(np.arange(self.n_mon)).q, (R, self.n_mon))
        squeeze = (lambda a: a[0]) if R == 1 else (lambda a: a)
        for k, v in out.items():
            setattr(self, k, squeeze(v))
        self.qdist = squeeze(qdist)
        self.elapsed_seconds = time.time() - start

    # -- notebook cell 4 (NB:276-326): the pH sweep as one batched call ------------------------------
    def titration_curve(self):
        """Mean degree of deprotonation per replica from the last perform_sampling:
        (alpha of the NH sites, alpha of the NH2 sites)."""
        return (np.mean(self.num_As, axis=-1) / self.n_nh, np.mean(self.num_As2, axis=-1) / self.n_nh2)
