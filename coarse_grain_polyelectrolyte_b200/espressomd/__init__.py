"""The slice of the ESPResSo Python API that the reference's src/modules calls, backed by libcgpe.so.

The reference (Asureda/coarse_grain_polyelectrolyte) hands every floating-point operation of its
hot path to ESPResSo through the calls listed in SURVEY.md §8b.  This package offers exactly
those names with the same keyword arguments, so `Model` / `EnsembleDynamics` read like the
reference's, while the work runs in the hand-written sm_100a kernels behind include/cgpe.h:

    espressomd.System(box_l=...)                          P.py:372      -> cgpe_create (lazily)
    system.part.add / by_id / by_ids / len                M.py:127-198  -> cgpe_set_state / get_state
    system.bonded_inter.add, interactions.*               M.py:34-54    -> cgpe_config bonded terms
    system.non_bonded_inter[a, b].wca.set_params          M.py:74-76    -> cgpe_set_wca
    electrostatics.DH, system.actors.add                  M.py:97-109   -> cgpe_set_replica_params
    system.integrator.run / set_vv / set_steepest_descent M.py:78-80    -> cgpe_md_run / cgpe_steepest_descent
    system.thermostat.set_langevin                        E.py:35,57,62 -> cgpe_set_thermostat_seed, kT, gamma
    reaction_methods.ConstantpHEnsemble                   M.py:211-261  -> cgpe_mc_run
    system.analysis.{min_dist,energy,calc_rg,calc_re}     E.py:34-101   -> cgpe_min_dist / energy / observables
    system.number_of_particles(type=)                     E.py:97-99    -> cgpe_get_state (types)

One extension: `System(..., n_replicas=R)` holds R independent copies of the system (the pH x
ionic-strength replica axis); per-replica quantities then accept arrays of length R and results
gain a leading axis of length R.  With the default R = 1 every call has ESPResSo's shapes.

There is no CPU path: the first call that needs arithmetic creates the engine and fails loudly
if libcgpe.so or a CUDA device is missing.
"""
import numpy as np

from .. import _abi, topology

FEATURES = ("WCA", "ELECTROSTATICS", "LENNARD_JONES", "EXTERNAL_FORCES", "MASS")
_current_system = None


def assert_features(features):
    """espressomd.assert_features(["WCA", "ELECTROSTATICS"]) (M.py:7, E.py:6)."""
    missing = [f for f in features if f not in FEATURES]
    if missing:
        raise RuntimeError(f"features not available in this engine: {missing}")


def features():
    return list(FEATURES)


class _CellSystem:
    """cell_system.skin (P.py:374) is an ESPResSo tuning knob; the engine keeps its own Verlet
    skin (cgpe_config.skin).  The value is stored and otherwise ignored."""

    def __init__(self):
        self.skin = 0.4
        self.n_square = False

    def set_n_square(self, use_verlet_lists=True):   # M.py:112
        self.n_square = True

    def set_regular_decomposition(self, **_):
        self.n_square = False


class _WcaHandle:
    def __init__(self, system, key):
        self._system, self._key = system, key

    def set_params(self, epsilon, sigma):
        if epsilon < 0 or sigma < 0:
            raise ValueError("WCA epsilon and sigma must be non-negative")
        a, b = self._key
        s = self._system
        n = max(a, b) + 1
        if n > s._wca_eps.shape[0]:
            grow = np.zeros((n, n))
            grow[: s._wca_eps.shape[0], : s._wca_eps.shape[0]] = s._wca_eps
            s._wca_eps = grow
            grow = np.zeros((n, n))
            grow[: s._wca_sig.shape[0], : s._wca_sig.shape[0]] = s._wca_sig
            s._wca_sig = grow
        s._wca_eps[a, b] = s._wca_eps[b, a] = epsilon
        s._wca_sig[a, b] = s._wca_sig[b, a] = sigma
        s._use_wca = True
        s._wca_changed()

    def get_params(self):
        a, b = self._key
        s = self._system
        if max(a, b) >= s._wca_eps.shape[0]:
            return {"epsilon": 0.0, "sigma": 0.0}
        return {"epsilon": float(s._wca_eps[a, b]), "sigma": float(s._wca_sig[a, b])}


class _PairHandle:
    def __init__(self, system, key):
        self.wca = _WcaHandle(system, key)


class _NonBondedInter:
    def __init__(self, system):
        self._system = system

    def __getitem__(self, key):
        a, b = (int(k) for k in key)
        if min(a, b) < 0 or max(a, b) >= _abi.MAX_TYPES:
            raise IndexError(f"particle types must lie in [0, {_abi.MAX_TYPES})")
        return _PairHandle(self._system, (a, b))


class _BondedInter:
    def __init__(self, system):
        self._system, self._bonds = system, []

    def add(self, bond):
        from . import interactions
        if not isinstance(bond, interactions.BondedInteraction):
            raise TypeError("bonded_inter.add needs a bonded interaction object")
        bond._id = len(self._bonds)
        self._bonds.append(bond)
        self._system._structure_changed()
        return bond._id

    def __len__(self):
        return len(self._bonds)

    def __getitem__(self, i):
        return self._bonds[i]


class _Actors:
    def __init__(self, system):
        self._system, self._actors = system, []

    def add(self, actor):
        from . import electrostatics
        if not isinstance(actor, electrostatics.DH):
            raise TypeError("only electrostatics.DH actors are supported")
        if any(isinstance(a, electrostatics.DH) for a in self._actors):
            raise RuntimeError("an electrostatics actor is already active")
        self._actors.append(actor)
        actor._system = self._system
        self._system._dh = actor
        self._system._structure_changed()

    def remove(self, actor):
        self._actors.remove(actor)
        if self._system._dh is actor:
            self._system._dh = None
        self._system._structure_changed()

    def clear(self):
        for a in list(self._actors):
            self.remove(a)

    def __len__(self):
        return len(self._actors)

    def __iter__(self):
        return iter(self._actors)


class _ParticleHandle:
    def __init__(self, system, pid):
        self._system, self.id = system, int(pid)

    def add_bond(self, bond):
        """part.by_id(i).add_bond((bond_object, *partner_ids)) (M.py:160-170)."""
        bond_obj, partners = bond[0], tuple(int(p) for p in bond[1:])
        self._system._add_bond(self.id, bond_obj, partners)

    def _get(self, what):
        return self._system.part.by_ids([self.id])._get(what)[..., 0, :] if what in ("pos", "v", "f") else \
            self._system.part.by_ids([self.id])._get(what)[..., 0]

    pos = property(lambda self: self._get("pos"))
    v = property(lambda self: self._get("v"))
    q = property(lambda self: self._get("q"))
    type = property(lambda self: self._get("type"))


class _ParticleSlice:
    def __init__(self, system, ids):
        self._system, self.id = system, np.asarray(ids, dtype=np.int64)

    def _get(self, what):
        s = self._system
        xyzq, vel, typ = s._state()
        sel = self.id
        out = {"pos": xyzq[:, sel, :3], "q": xyzq[:, sel, 3], "v": vel[:, sel, :3], "type": typ[:, sel].astype(np.int64)}[what]
        out = out.astype(np.float64) if what != "type" else out
        return out[0] if s.n_replicas == 1 else out

    pos = property(lambda self: self._get("pos"))
    v = property(lambda self: self._get("v"))
    q = property(lambda self: self._get("q"))
    type = property(lambda self: self._get("type"))

    def __len__(self):
        return len(self.id)


class _ParticleList:
    def __init__(self, system):
        self._system = system

    def __len__(self):
        s = self._system
        if s._hidden_type is None:
            return s._n_part
        return int((s._state()[2][0] < s._hidden_type).sum())     # hidden ion slots are not particles

    def add(self, id=None, pos=None, type=0, q=0.0, v=None, fix=None, **unsupported):
        """Scalar form (M.py:130-155) and vectorised form (M.py:175-198).  fix=[True]*3 pins the
        particle(s) in space (all three axes or none; the engine fixes whole particle types)."""
        if unsupported:
            raise NotImplementedError(f"particle properties not supported: {sorted(unsupported)}")
        if pos is None:
            raise ValueError("pos is required")
        s = self._system
        pos = np.asarray(pos, dtype=np.float64)
        single = pos.ndim == 1
        pos = np.atleast_2d(pos)
        n = pos.shape[0]
        if pos.shape[1] != 3:
            raise ValueError("pos must have 3 components")
        types = np.broadcast_to(np.asarray(type, dtype=np.int64), (n,))
        qs = np.broadcast_to(np.asarray(q, dtype=np.float64), (n,))
        vs = np.zeros((n, 3)) if v is None else np.broadcast_to(np.asarray(v, dtype=np.float64), (n, 3))
        if id is None:
            ids = np.arange(s._n_part, s._n_part + n)
        else:
            ids = np.atleast_1d(np.asarray(id, dtype=np.int64))
        if len(ids) != n or not np.array_equal(ids, np.arange(s._n_part, s._n_part + n)):
            raise NotImplementedError("particle ids must be consecutive, starting at len(system.part)")
        if np.any(types < 0) or np.any(types >= _abi.MAX_TYPES):
            raise ValueError("particle type out of range")
        fixed = False
        if fix is not None:
            fx = np.asarray(fix, dtype=bool).reshape(-1)
            if fx.size != 3 or (fx.any() and not fx.all()):
                raise NotImplementedError("fix: all three axes or none")
            fixed = bool(fx.all())
        s._append_particles(pos, vs, types, qs)
        s._fixed = np.concatenate([s._fixed, np.full(n, fixed)])
        if single:
            return _ParticleHandle(s, ids[0])
        return _ParticleSlice(s, ids)

    def by_id(self, pid):
        if not 0 <= int(pid) < self._system._n_part:
            raise IndexError(f"particle {pid} does not exist")
        return _ParticleHandle(self._system, pid)

    def by_ids(self, ids):
        ids = np.asarray(ids, dtype=np.int64)
        if ids.size and (ids.min() < 0 or ids.max() >= self._system._n_part):
            raise IndexError("particle id out of range")
        return _ParticleSlice(self._system, ids)

    def all(self):
        return _ParticleSlice(self._system, np.arange(self._system._n_part))

    def select(self, type=None):
        typ = self._system._state()[2][0]
        ids = np.arange(self._system._n_part) if type is None else np.nonzero(typ == type)[0]
        return _ParticleSlice(self._system, ids)


class _Integrator:
    def __init__(self, system):
        self._system = system
        self._mode = "vv"
        self._sd = {}

    def set_vv(self):                                            # M.py:80
        self._mode = "vv"

    def set_steepest_descent(self, f_max, gamma, max_displacement):   # M.py:78
        if f_max < 0 or gamma < 0 or max_displacement < 0:
            raise ValueError("steepest descent parameters must be non-negative")
        self._mode = "sd"
        self._sd = dict(f_max=float(f_max), gamma=float(gamma), max_displacement=float(max_displacement))

    def run(self, steps=1, **_):
        """integrator.run(n) (E.py:41,52,58,64,73,90,92; M.py:79)."""
        steps = int(steps)
        if steps < 0:
            raise ValueError("steps must be non-negative")
        eng = self._system._ensure_engine()
        if self._mode == "sd":
            eng.steepest_descent(steps, **self._sd)
            self._system._state_cache = None
            return
        self._system.thermostat._require()
        acc = self._system.auto_update_accumulators
        if len(acc) > 0 and self._system._frames_ready(eng, lambda every: steps // every + 2):
            # every registered accumulator can be fed from the engine's frame recorder: ONE launch for all the steps
            eng.md_run(steps)
            self._system._state_cache = None
            self._system._frames_replay(eng)
            for a in acc:                                # keep the host-side phase in step (a later call may fall back)
                a._since_update = (a._since_update + steps) % a.delta_N
            return
        if len(acc) == 0 and self._system._frames_state is not None:
            self._system._frames_off(eng)
        while steps > 0:   # stop at every step count at which a registered accumulator is due (O.py:13-15)
            n = steps if len(acc) == 0 else min(steps, acc.steps_until_next())
            eng.md_run(n)
            self._system._state_cache = None
            acc.advance(n)
            steps -= n


class _Thermostat:
    def __init__(self, system):
        self._system = system
        self.kT = None
        self.gamma = None
        self.seed = None
        self._pending_seed = None

    def set_langevin(self, kT, gamma, seed=None):                # E.py:35,57,61-63
        kT_a, g_a = np.asarray(kT, dtype=np.float64), np.asarray(gamma, dtype=np.float64)
        if np.any(kT_a < 0) or np.any(g_a < 0):
            raise ValueError("kT and gamma must be non-negative")
        if seed is None and self.seed is None:
            raise ValueError("A seed has to be given as keyword argument on first activation of the thermostat")
        self.kT, self.gamma = kT, gamma
        if seed is not None:
            self.seed = int(seed)
            self._pending_seed = int(seed)
        s = self._system
        if s._engine is not None:
            self._push(s._engine)

    def _push(self, eng):
        eng.set_replica_params(kT=self.kT, gamma=self.gamma)
        if self._pending_seed is not None:
            eng.set_thermostat_seed(self._pending_seed)
            self._pending_seed = None

    def turn_off(self):
        self.set_langevin(kT=0.0, gamma=0.0, seed=self.seed or 0)

    def _require(self):
        if self.kT is None:
            # ESPResSo integrates without thermostat when none is set: NVE
            self.set_langevin(kT=0.0, gamma=0.0, seed=0)


class _Analysis:
    def __init__(self, system):
        self._system = system

    def _squeeze(self, a):
        return a[0] if self._system.n_replicas == 1 else a

    def min_dist(self):                                          # E.py:34,45
        return self._squeeze(self._system._ensure_engine().min_dist())

    def energy(self):                                            # E.py:42
        e = self._system._ensure_engine().energy()
        return {k: self._squeeze(e[:, i]) for i, k in enumerate(_abi.ENERGY_COLUMNS)}

    def _chain_args(self, chain_start, number_of_chains, chain_length):
        s = self._system
        if chain_start != 0 or number_of_chains != 1 or chain_length != s._n_beads_per_chain():
            raise NotImplementedError("calc_rg/calc_re are evaluated for chain 0 over its full length")

    def calc_rg(self, chain_start, number_of_chains, chain_length):   # E.py:100
        """[mean Rg, its std over chains, mean Rg^2, its std] like ESPResSo; one chain -> std 0."""
        self._chain_args(chain_start, number_of_chains, chain_length)
        rg, _, _ = self._system._ensure_engine().observables()
        rg = self._squeeze(rg)
        z = np.zeros_like(rg)
        return np.array([rg, z, rg * rg, z])

    def calc_re(self, chain_start, number_of_chains, chain_length):   # E.py:101
        self._chain_args(chain_start, number_of_chains, chain_length)
        _, ree, _ = self._system._ensure_engine().observables()
        ree = self._squeeze(ree)
        z = np.zeros_like(ree)
        return np.array([ree, z, ree * ree, z])


class System:
    """espressomd.System(box_l=[L, L, L]) (P.py:372) with an optional replica axis."""

    # Class that drives the C-ABI handle; None = the product's `Engine` over libcgpe.so (CUDA; raises without a
    # device).  A test may point it at another implementation of the same ABI to run the driver code elsewhere.
    engine_factory = None

    def __init__(self, box_l, n_replicas=1, replica_offset=0, device=-1, skin=0.0):
        global _current_system
        box_l = np.atleast_1d(np.asarray(box_l, dtype=np.float64))
        if box_l.size == 1:
            box_l = np.repeat(box_l, 3)
        if box_l.size != 3 or np.any(box_l <= 0):
            raise ValueError("box_l must be three positive lengths")
        if not np.allclose(box_l, box_l[0]):
            raise NotImplementedError("only cubic boxes are supported")
        if int(n_replicas) < 1:
            raise ValueError("n_replicas must be positive")
        self.box_l = box_l
        self.n_replicas = int(n_replicas)
        self.replica_offset = int(replica_offset)
        self._device = device
        self._engine_skin = float(skin)
        self.time_step = None
        self.periodicity = [True, True, True]
        self.cell_system = _CellSystem()
        self.part = _ParticleList(self)
        self.bonded_inter = _BondedInter(self)
        self.non_bonded_inter = _NonBondedInter(self)
        self.actors = _Actors(self)
        self.integrator = _Integrator(self)
        self.thermostat = _Thermostat(self)
        self.analysis = _Analysis(self)
        self.auto_update_accumulators = accumulators.AutoUpdateAccumulators()
        self._force_cap = 0.0
        self._time = 0.0
        # host-side description until (and between) engine builds
        self._pos = np.zeros((0, 3)); self._vel = np.zeros((0, 3))
        self._type = np.zeros(0, np.int64); self._q = np.zeros(0)
        self._n_part = 0
        self._bond_records = []                 # (particle, bond object, partners)
        self._wca_eps = np.zeros((1, 1)); self._wca_sig = np.zeros((1, 1))
        self._use_wca = False
        self._dh = None
        self._reactions = []                    # reaction_methods._Registered
        self._engine = None
        self._engine_particles = 0
        self._fixed = np.zeros(0, dtype=bool)   # per particle: fix=[True]*3
        self._hidden_type = None               # type index of hidden (non-interacting) ion slots once they exist
        self._state_cache = None
        self._frames_state = None              # on-device frame recorder armed for the registered accumulators
        _current_system = self

    # -- ESPResSo attributes ---------------------------------------------------------------------
    @property
    def time(self):                                               # E.py:30,77,102
        if self._engine is None:
            return self._time
        t = self._engine.get_time()
        return float(t[0]) if self.n_replicas == 1 else t

    @time.setter
    def time(self, value):
        self._time = float(value)
        if self._engine is not None:
            self._engine.set_time(self._time)

    @property
    def force_cap(self):                                          # E.py:32,48,51
        return self._force_cap

    @force_cap.setter
    def force_cap(self, value):
        if value < 0:
            raise ValueError("force_cap must be non-negative")
        self._force_cap = float(value)
        if self._engine is not None:
            self._engine.set_force_cap(self._force_cap)

    def number_of_particles(self, type):                          # E.py:97-99
        typ = self._state()[2]
        n = (typ == int(type)).sum(axis=1)
        return int(n[0]) if self.n_replicas == 1 else n

    # -- host-side bookkeeping -------------------------------------------------------------------
    def _append_particles(self, pos, vel, types, qs):
        self._pull_state()
        self._pos = np.concatenate([self._pos, pos]); self._vel = np.concatenate([self._vel, vel])
        self._type = np.concatenate([self._type, types]); self._q = np.concatenate([self._q, qs])
        self._n_part = len(self._type)
        self._structure_changed()

    def _add_bond(self, pid, bond_obj, partners):
        if bond_obj not in self.bonded_inter._bonds:
            raise ValueError("the bonded interaction was not added to system.bonded_inter")
        if len(partners) != bond_obj.n_partners:
            raise ValueError(f"{type(bond_obj).__name__} needs {bond_obj.n_partners} partners")
        self._bond_records.append((int(pid), bond_obj, partners))
        self._structure_changed()

    def _structure_changed(self):
        if self._engine is not None:
            self._pull_state()
            self._time = float(self._engine.get_time()[0])
            self._engine.close()
            self._engine = None

    def _wca_changed(self):
        if self._engine is None:
            return
        if self._wca_eps.shape[0] > self._engine.cfg.n_types or not self._engine.cfg.use_wca:
            self._structure_changed()
        else:
            self._engine.set_wca(self._wca_eps, self._wca_sig)

    def _pull_state(self):
        """Bring positions / velocities / types back from the device before the host copy is
        edited.  Only possible without loss when the replicas still share one state."""
        if self._engine is None:
            return
        xyzq, vel, typ = self._engine.get_state()
        if self.n_replicas > 1 and not all(np.array_equal(a[0], a[r]) for a in (xyzq, vel, typ) for r in range(1, self.n_replicas)):
            raise NotImplementedError("the system cannot be restructured once the replicas have diverged")
        self._pos = xyzq[0, :, :3].astype(np.float64); self._q = xyzq[0, :, 3].astype(np.float64)
        self._vel = vel[0, :, :3].astype(np.float64); self._type = typ[0].astype(np.int64)

    def _n_beads_per_chain(self):
        return self._topology()[1]

    def _topology(self):
        """Recognise the reference's chain topology (M.py:159-170) in the bond records: linear
        chains of equal length on consecutive ids starting at 0, bond (i-1,i) on every bead but
        the first, optional angle (i-1,i,i+1) and dihedral (i-1,i,i+1,i+2) on the inner beads.
        Returns (n_chains, n_beads, bond, angle, dihedral)."""
        from . import interactions
        pair = {}; ang = {}; dih = {}
        bond_t = ang_t = dih_t = None
        for pid, b, partners in self._bond_records:
            if isinstance(b, (interactions.HarmonicBond, interactions.FeneBond)):
                if partners != (pid - 1,):
                    raise NotImplementedError("pair bonds must connect a bead to its predecessor (id-1)")
                if bond_t not in (None, b):
                    raise NotImplementedError("one pair-bond type per system")
                bond_t = b; pair[pid] = True
            elif isinstance(b, interactions.AngleHarmonic):
                if partners != (pid - 1, pid + 1):
                    raise NotImplementedError("angle bonds must be (id-1, id+1) around the centre bead")
                if ang_t not in (None, b):
                    raise NotImplementedError("one angle type per system")
                ang_t = b; ang[pid] = True
            elif isinstance(b, interactions.Dihedral):
                if partners != (pid - 1, pid + 1, pid + 2):
                    raise NotImplementedError("dihedrals must be (id-1, id+1, id+2) on the second bead")
                if dih_t not in (None, b):
                    raise NotImplementedError("one dihedral type per system")
                dih_t = b; dih[pid] = True
        if not pair:
            # no bonds: every particle is its own "chain" of one bead is not useful; treat all
            # particles as one unbonded chain so that Rg/Ree stay defined
            return 1, max(self._n_part, 1), None, None, None
        bonded = sorted(pair)
        # chain starts are the beads without a bond to their predecessor
        n_chain_beads = bonded[-1] + 1
        starts = [i for i in range(n_chain_beads) if i not in pair]
        lengths = np.diff(starts + [n_chain_beads])
        if starts[0] != 0 or len(set(lengths.tolist())) != 1:
            raise NotImplementedError("chains must be contiguous from id 0 and of equal length")
        n = int(lengths[0])
        for c0 in starts:
            for l in range(n):
                i = c0 + l
                if ang_t is not None and (1 <= l <= n - 2) != (i in ang):
                    raise NotImplementedError("angle bonds must sit on every inner bead")
                if dih_t is not None and (1 <= l <= n - 3) != (i in dih):
                    raise NotImplementedError("dihedrals must sit on beads 1..n-3 of every chain")
        return len(starts), n, bond_t, ang_t, dih_t

    def _build_config(self):
        from . import interactions
        if self.time_step is None or self.time_step <= 0:
            raise RuntimeError("system.time_step must be set before integrating")
        n_chains, n_beads, bond_t, ang_t, dih_t = self._topology()
        n_chain = n_chains * n_beads
        if self._n_part < n_chain:
            raise RuntimeError("bond records reference particles that were not added")
        real = self._type if self._hidden_type is None else self._type[self._type != self._hidden_type]
        T = int(max(self._wca_eps.shape[0], real.max(initial=0) + 1,
                    max([max(r.type_prot, r.type_deprot, r.type_ion) + 1 for r in self._reactions], default=1)))
        if self._hidden_type is not None and T != self._hidden_type:
            raise NotImplementedError("particle types cannot be added once hidden ion slots exist")
        # free particles (explicit ions, M.py:173-198) live in ion slots behind the chain; one spare
        # hidden slot per titratable site lets every site release its H+ (ESPResSo creates particles
        # on demand; here the pool is sized up front)
        # fixed particles (fix=[True]*3): the engine pins whole types
        fixed = np.concatenate([self._fixed, np.zeros(self._n_part - len(self._fixed), dtype=bool)])
        fixed_types = sorted({int(t) for t in self._type[fixed]})
        if any(not f and int(t) in fixed_types for f, t in zip(fixed, self._type)):
            raise NotImplementedError("a particle type is either fixed (fix=[True]*3 on all its particles) or free")
        fixed_mask = sum(1 << t for t in fixed_types)
        # free particles that are all fixed are spectators (a charged nanoparticle): the reactions stay
        # the implicit-ion ones and need no H+ pool
        spectators = self._n_part > n_chain and bool(np.all(fixed[n_chain:]))
        if self._n_part > n_chain and not spectators:
            reactive = {t for r in self._reactions for t in (r.type_prot, r.type_deprot)}
            n_sites = int(sum(int(t) in reactive for t in self._type[:n_chain]))
            n_hidden = 0 if self._hidden_type is None else int((self._type == self._hidden_type).sum())
            n_spare = max(0, n_sites - n_hidden)
            if n_spare > 0:
                self._hidden_type = T
                # parked at scattered positions (they do not interact; a revived slot gets a fresh position)
                park = np.random.RandomState(12345).random_sample((n_spare, 3)) * float(self.box_l[0])
                self._pos = np.concatenate([self._pos, park]); self._vel = np.concatenate([self._vel, np.zeros((n_spare, 3))])
                self._type = np.concatenate([self._type, np.full(n_spare, T, np.int64)]); self._q = np.concatenate([self._q, np.zeros(n_spare)])
                self._n_part = len(self._type)
        n_ion_slots = self._n_part - n_chain

        kw = dict(
            n_replicas=self.n_replicas, replica_offset=self.replica_offset, n_chains=n_chains, n_beads=n_beads,
            n_ion_slots=n_ion_slots, n_types=T, device=self._device, skin=self._engine_skin,
            box_l=float(self.box_l[0]), time_step=float(self.time_step),
            bond_kind=_abi.BOND_NONE, use_angle=0, use_dihedral=0, use_wca=int(self._use_wca), use_dh=0,
            wca_eps=self._wca_eps, wca_sig=self._wca_sig, fixed_type_mask=fixed_mask,
        )
        if isinstance(bond_t, interactions.HarmonicBond):
            kw.update(bond_kind=_abi.BOND_HARMONIC, bond_k=bond_t.k, bond_r0=bond_t.r_0)
        elif isinstance(bond_t, interactions.FeneBond):
            kw.update(bond_kind=_abi.BOND_FENE, bond_k=bond_t.k, bond_r0=bond_t.r_0, fene_drmax=bond_t.d_r_max)
        if ang_t is not None:
            kw.update(use_angle=1, angle_k=ang_t.bend, angle_phi0=ang_t.phi0)
        if dih_t is not None:
            kw.update(use_dihedral=1, dihedral_k=dih_t.bend, dihedral_mult=dih_t.mult, dihedral_phase=dih_t.phase,
                      dihedral_sin_min=getattr(dih_t, "sin_min", 0.0))
        if self._dh is not None:
            kw.update(use_dh=1, dh_prefactor=self._dh.prefactor,
                      kappa=float(np.atleast_1d(self._dh.kappa)[0]), r_cut=float(np.max(self._dh.r_cut)))
        kw["reactions"] = [r.as_dict() for r in self._reactions]
        if spectators:
            for rx in kw["reactions"]:
                rx["type_ion"] = -1
        return _abi.make_config(**kw)

    def _ensure_engine(self):
        if self._engine is None:
            Engine = type(self).engine_factory
            if Engine is None:
                from ..engine import Engine
            if self._n_part == 0:
                raise RuntimeError("the system holds no particles")
            cfg = self._build_config()
            eng = Engine(cfg)
            xyzq = np.concatenate([self._pos, self._q[:, None]], axis=1).astype(np.float32)
            eng.set_state(xyzq, self._vel.astype(np.float32), self._type.astype(np.uint8))
            if self._dh is not None:
                eng.set_replica_params(kappa=self._dh.kappa, r_cut=self._dh.r_cut)
            if self._reactions:
                eng.set_replica_params(pH=self._reactions[0].owner.constant_pH)
            eng.set_force_cap(self._force_cap)
            eng.set_time(self._time)
            self._engine = eng
            if self.thermostat.kT is not None:
                self.thermostat._pending_seed = self.thermostat.seed
                self.thermostat._push(eng)
            self._state_cache = None
        return self._engine

    # -- accumulators fed from the engine's on-device frame recorder (cgpe_frames_*) ---------------------
    def _frames_plan(self):
        """None, or how the registered accumulators map onto recorded frames: (every, flags, [(accumulator, source, stride)]).
        Served: TimeSeries / MeanVarianceCalculator / Correlator of ComPosition, ComVelocity, ParticleDistances, BondDihedrals
        over the beads of chain 0 in order (what Observables.py:54-76 registers), all in phase."""
        import math
        from . import observables as obs_mod
        accs = list(self.auto_update_accumulators)
        if not accs or self._n_part == 0:
            return None
        try:
            _, n_beads, *_ = self._topology()
        except Exception:      # noqa: BLE001  (not a chain system)
            return None
        chain = np.arange(n_beads)
        every = 0
        for a in accs:
            every = math.gcd(every, a.delta_N)
        flags, items = 0, []
        for a in accs:
            o = getattr(a, "obs", None) or getattr(a, "obs1", None)
            if o is None or getattr(a, "obs2", None) is not None:
                return None
            if o.ids is None or not np.array_equal(o.ids, chain) or (o._bound is not None and o._bound is not self):
                return None
            src = {obs_mod.ComPosition: "com", obs_mod.ComVelocity: "vcom", obs_mod.ParticleDistances: "bonds",
                   obs_mod.BondDihedrals: "dih"}.get(type(o))
            if src is None:
                return None
            flags |= _abi.FRAME_BONDS if src == "bonds" else _abi.FRAME_DIHEDRALS if src == "dih" else 0
            items.append((a, src, a.delta_N // every))
        return every, flags, items

    def _frames_ready(self, eng, capacity_of):
        """Arm the frame recorder for the registered accumulators (capacity_of(every) frames per replica); False if they need
        the step-by-step host path."""
        plan = self._frames_plan()
        if plan is None or not hasattr(eng, "frames_begin"):
            self._frames_off(eng)
            return False
        every, flags, items = plan
        capacity = int(capacity_of(every))
        key = (every, flags, tuple(id(a) for a, _, _ in items))
        st = self._frames_state
        if st is None or st["key"] != key or st["engine"] is not eng or st["capacity"] < capacity:
            grow = st is not None and st["engine"] is eng and st["key"] == key
            if grow:
                self._frames_replay(eng)                 # nothing recorded so far may be lost; phase and counters carry over
            else:
                self._frames_off(eng)
                if any(a._since_update != 0 for a, _, _ in items):
                    return False                         # accumulators in mid-interval: the recorder starts at phase 0
            try:
                eng.frames_begin(every, flags, int(capacity))
            except _abi.CgpeError:
                self._frames_state = None
                return False
            keep = self._frames_state if grow else dict(backlog=None, count=0)
            self._frames_state = dict(key=key, engine=eng, capacity=int(capacity), backlog=keep["backlog"], count=keep["count"])
        return True

    def _frames_off(self, eng):
        if self._frames_state is not None:
            try:
                self._frames_state["engine"].frames_end()
            except Exception:      # noqa: BLE001
                pass
            self._frames_state = None

    def _frames_replay(self, eng):
        """Fetch what the recorder holds and feed it to the accumulators in order.  Replicas of a fused sampling loop record
        different numbers of frames (each has its own burst schedule); the replica-batched accumulators take the frames all
        replicas have, the surplus waits for the next call."""
        st = self._frames_state
        if st is None:
            return
        plan = self._frames_plan()
        if plan is None or (plan[0], plan[1], tuple(id(a) for a, _, _ in plan[2])) != st["key"]:
            self._frames_off(eng)                        # the accumulators were cleared or replaced: nothing to feed
            return
        _, _, items = plan
        fr, bo, di = eng.frames_fetch()
        R = self.n_replicas
        new = [dict(com=fr[r][:, 0:3], vcom=fr[r][:, 3:6], bonds=None if bo is None else bo[r], dih=None if di is None else di[r])
               for r in range(R)]
        if st["backlog"] is not None:
            new = [{k: (v if st["backlog"][r][k] is None else np.concatenate([st["backlog"][r][k], v]))
                    for k, v in new[r].items()} for r in range(R)]
        n = min(len(d["com"]) for d in new)
        for k in range(n):
            st["count"] += 1
            for acc, src, stride in items:
                if st["count"] % stride == 0:
                    val = np.stack([new[r][src][k] for r in range(R)])
                    acc.update_with(val[0] if R == 1 else val)
        st["backlog"] = [{k: (None if v is None else v[n:]) for k, v in d.items()} for d in new]

    def _state(self):
        """(xyzq [R][P][4], vel [R][P][4], type [R][P]) of the current configuration."""
        if self._engine is None:
            R = self.n_replicas
            xyzq = np.concatenate([self._pos, self._q[:, None]], axis=1).astype(np.float32)
            vel = np.concatenate([self._vel, np.zeros((self._n_part, 1))], axis=1).astype(np.float32)
            return (np.broadcast_to(xyzq, (R,) + xyzq.shape), np.broadcast_to(vel, (R,) + vel.shape),
                    np.broadcast_to(self._type.astype(np.uint8), (R, self._n_part)))
        if self._state_cache is None:
            self._state_cache = self._engine.get_state()
        return self._state_cache

    # -- extension: the fused sampling loop --------------------------------------------------------
    def sample(self, n_samples, md_steps_per_burst, p_burst1, p_burst2, mc_blocks, use_md=True,
               q_snapshot_every=0, schedule_seed=10):
        """Run the loop body of EnsembleDynamics.perform_sampling (E.py:88-106) for `n_samples`
        samples in ONE kernel launch.  mc_blocks: sequence of (ConstantpHEnsemble, n_trials).
        Returns (obs [R][n_samples][6], qdist [R][rows][n_beads] or None); columns of obs:
        n_deprot(reaction 0), n_deprot(reaction 1), n_ion, Rg, Ree, time."""
        eng = self._ensure_engine()
        if use_md:
            self.thermostat._require()
        pHs = {tuple(np.broadcast_to(np.asarray(rx.constant_pH, float), (self.n_replicas,))) for rx, _ in mc_blocks}
        if len(pHs) > 1:
            raise NotImplementedError("all reaction objects of a fused sampling loop must share constant_pH")
        blocks = []
        for rx, n in mc_blocks:
            rx._push_pH(eng)
            blocks.append((rx._single_reaction().index, int(n)))
        rows = (int(n_samples) + q_snapshot_every - 1) // q_snapshot_every if q_snapshot_every > 0 else 0
        sp = eng.sample_params(n_samples, md_steps_per_burst=md_steps_per_burst, use_md=use_md, mc_blocks=blocks,
                               q_snapshot_every=q_snapshot_every, q_snapshot_rows=rows,
                               schedule_seed=schedule_seed, p_burst1=p_burst1, p_burst2=p_burst2)
        if len(self.auto_update_accumulators) > 0:
            # registered accumulators (Observables.py) are fed from frames the kernel records every delta_N steps: the loop
            # stays ONE launch.  Capacity: expected frames of the schedule + 40 %, but never more than every burst taken.
            def cap(every):
                per_burst = md_steps_per_burst // every + 1
                return int(min(2 * n_samples * per_burst, 1.4 * n_samples * (p_burst1 + p_burst2) * per_burst + 64)) + 2
            if not (use_md and self._frames_ready(eng, cap)):
                raise RuntimeError("the registered accumulators cannot be fed by the engine's frame recorder; use the step-by-step "
                                   "loop (EnsembleDynamics._perform_sampling_stepwise)")
        elif self._frames_state is not None:
            self._frames_off(eng)
        out = eng.sample_loop(sp)
        self._state_cache = None
        if self._frames_state is not None:
            self._frames_replay(eng)
        return out


from . import accumulators, electrostatics, interactions, observables, polymer, reaction_methods  # noqa: E402,F401
