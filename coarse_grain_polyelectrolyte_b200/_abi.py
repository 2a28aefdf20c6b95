"""ctypes mirror of include/cgpe.h and a thin numpy-facing wrapper over a handle.

`Binding(lib, prefix)` attaches argument types to the C-ABI entry points of a loaded shared
library; `EngineBase` drives one handle through such a binding with numpy host buffers.  The
product binds `libcgpe.so` (see engine.py); the tests bind the CPU oracle with the same class,
which is what keeps the parity tests symmetrical.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 6
MAX_TYPES = 16
MAX_REACTIONS = 4
MAX_MC_BLOCKS = 4
N_OBS = 6
N_ENERGY = 5
FRAME_COLS = 8
FRAME_BONDS, FRAME_DIHEDRALS = 1, 2

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_CAPACITY, ERR_STATE, ERR_NUMERIC = 0, -1, -2, -3, -4, -5, -6
BOND_NONE, BOND_HARMONIC, BOND_FENE = 0, 1, 2

OBS_COLUMNS = ("n_deprot_0", "n_deprot_1", "n_ion", "rg", "ree", "time")
ENERGY_COLUMNS = ("total", "kinetic", "coulomb", "bonded", "non_bonded")


class Reaction(C.Structure):
    _fields_ = [
        ("type_prot", C.c_int32), ("type_deprot", C.c_int32), ("type_ion", C.c_int32), ("seed", C.c_int32),
        ("q_prot", C.c_double), ("q_deprot", C.c_double), ("q_ion", C.c_double),
        ("pK", C.c_double), ("kT", C.c_double), ("exclusion_range", C.c_double),
    ]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_replicas", C.c_int32), ("replica_offset", C.c_int32),
        ("n_chains", C.c_int32), ("n_beads", C.c_int32), ("n_ion_slots", C.c_int32),
        ("n_types", C.c_int32), ("bond_kind", C.c_int32), ("use_angle", C.c_int32),
        ("use_dihedral", C.c_int32), ("dihedral_mult", C.c_int32), ("use_wca", C.c_int32),
        ("use_dh", C.c_int32), ("n_reactions", C.c_int32), ("cap_wca", C.c_int32),
        ("cap_dh", C.c_int32), ("device", C.c_int32), ("fixed_type_mask", C.c_int32),
        ("bond_k", C.c_double), ("bond_r0", C.c_double), ("fene_drmax", C.c_double),
        ("angle_k", C.c_double), ("angle_phi0", C.c_double),
        ("dihedral_k", C.c_double), ("dihedral_phase", C.c_double),
        ("dh_prefactor", C.c_double), ("box_l", C.c_double), ("time_step", C.c_double),
        ("skin", C.c_double),
        ("pH", C.c_double), ("kappa", C.c_double), ("r_cut", C.c_double), ("kT", C.c_double),
        ("gamma", C.c_double), ("dihedral_sin_min", C.c_double),
        ("wca_eps", C.c_double * (MAX_TYPES * MAX_TYPES)),
        ("wca_sig", C.c_double * (MAX_TYPES * MAX_TYPES)),
        ("reactions", Reaction * MAX_REACTIONS),
    ]


class SampleParams(C.Structure):
    _fields_ = [
        ("n_samples", C.c_int32), ("md_steps_per_burst", C.c_int32), ("use_md", C.c_int32),
        ("n_mc_blocks", C.c_int32),
        ("mc_reaction", C.c_int32 * MAX_MC_BLOCKS), ("mc_trials", C.c_int32 * MAX_MC_BLOCKS),
        ("q_snapshot_every", C.c_int32), ("q_snapshot_rows", C.c_int32),
        ("schedule_seed", C.c_int32), ("reserved0", C.c_int32),
        ("p_burst1", C.c_double), ("p_burst2", C.c_double),
    ]


TRIAL_DTYPE = np.dtype(
    [("particle", "<i4"), ("ion_slot", "<i4"), ("direction", "i1"), ("accepted", "i1"),
     ("excluded", "i1"), ("reserved", "i1"), ("delta_e", "<f4"), ("bf", "<f4"), ("u", "<f4")],
    align=True,
)
assert TRIAL_DTYPE.itemsize == 24


class CgpeError(RuntimeError):
    """A C-ABI call returned a negative cgpe_status."""

    def __init__(self, status, message, call):
        super().__init__(f"{call} failed with status {status}: {message}")
        self.status = status


_p = C.c_void_p
_dp, _fp, _u8p, _i32p, _i64p = (C.POINTER(t) for t in (C.c_double, C.c_float, C.c_uint8, C.c_int32, C.c_int64))

# name -> (restype, argtypes); exactly the symbols include/cgpe.h declares
SIGNATURES = {
    "abi_version": (C.c_int, []),
    "create": (C.c_int, [C.POINTER(Config), C.POINTER(_p)]),
    "destroy": (None, [_p]),
    "last_error": (C.c_char_p, [_p]),
    "device_info": (C.c_int, [_p, _i32p, _i32p, _i32p, C.c_char_p, C.c_int32]),
    "set_stream": (C.c_int, [_p, _p]),
    "synchronize": (C.c_int, [_p]),
    "set_state": (C.c_int, [_p, _p, _p, _p]),
    "get_state": (C.c_int, [_p, _p, _p, _p]),
    "set_replica_params": (C.c_int, [_p, _p, _p, _p, _p, _p]),
    "set_wca": (C.c_int, [_p, _p, _p]),
    "set_thermostat_seed": (C.c_int, [_p, C.c_int32]),
    "set_force_cap": (C.c_int, [_p, C.c_double]),
    "set_time": (C.c_int, [_p, C.c_double]),
    "get_time": (C.c_int, [_p, _p]),
    "md_run": (C.c_int, [_p, C.c_int32]),
    "steepest_descent": (C.c_int, [_p, C.c_int32, C.c_double, C.c_double, C.c_double]),
    "mc_run": (C.c_int, [_p, C.c_int32, C.c_int32, _p]),
    "sample_loop": (C.c_int, [_p, C.POINTER(SampleParams), _p, _p]),
    "sample_loop_async": (C.c_int, [_p, C.POINTER(SampleParams)]),
    "fetch_samples": (C.c_int, [_p, _p, _p]),
    "energy": (C.c_int, [_p, _p]),
    "forces": (C.c_int, [_p, _p]),
    "md_forces": (C.c_int, [_p, _p]),
    "observables": (C.c_int, [_p, _p, _p, _p]),
    "min_dist": (C.c_int, [_p, _p]),
    "get_counters": (C.c_int, [_p, _p, _p, _p, _p]),
    "list_stats": (C.c_int, [_p, _p]),
    "set_engine": (C.c_int, [_p, C.c_int32]),
    "engine_info": (C.c_int, [_p, _i32p, _p, _p]),
    "frames_begin": (C.c_int, [_p, C.c_int32, C.c_int32, C.c_int32]),
    "frames_fetch": (C.c_int, [_p, _p, _p, _p, _p]),
    "frames_end": (C.c_int, [_p]),
    "last_call_ms": (C.c_int, [_p, _fp, _i32p]),
}


class Binding:
    """Typed access to `<prefix><name>` for every entry of SIGNATURES."""

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, prefix + name)  # AttributeError if a declared symbol is missing
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_config(**kw):
    """Build a Config from keyword arguments; wca_eps/wca_sig accept [T][T] arrays and
    reactions a list of dicts."""
    cfg = Config()
    cfg.abi_version = ABI_VERSION
    cfg.device = -1
    cfg.n_chains = 1
    cfg.dihedral_mult = 3
    cfg.kT, cfg.gamma, cfg.pH = 1.0, 1.0, 7.0
    n_types = kw.get("n_types", 1)
    for k, v in kw.items():
        if k in ("wca_eps", "wca_sig"):
            full = np.zeros((MAX_TYPES, MAX_TYPES))
            v = np.asarray(v, dtype=np.float64)
            full[: v.shape[0], : v.shape[1]] = v
            getattr(cfg, k)[:] = full.ravel().tolist()
        elif k == "reactions":
            cfg.n_reactions = len(v)
            for m, rx in enumerate(v):
                r = cfg.reactions[m]
                r.kT = 1.0
                r.type_ion = -1
                for kk, vv in rx.items():
                    setattr(r, kk, vv)
        else:
            if not hasattr(cfg, k):
                raise TypeError(f"unknown config field {k!r}")
            setattr(cfg, k, v)
    cfg.n_types = n_types
    return cfg


class EngineBase:
    """One handle of R replicas driven through a Binding with numpy host buffers."""

    def __init__(self, binding, cfg):
        self._b = binding
        self._h = C.c_void_p()
        self.cfg = cfg
        rc = binding.create(C.byref(cfg), C.byref(self._h))
        if rc != OK:
            msg = binding.last_error(None)
            raise CgpeError(rc, msg.decode() if msg else "", binding.prefix + "create")
        self.R = cfg.n_replicas
        self.P = cfg.n_chains * cfg.n_beads + cfg.n_ion_slots
        self.n_beads = cfg.n_beads
        self.n_reactions = cfg.n_reactions

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc, call):
        if rc != OK:
            msg = self._b.last_error(self._h)
            raise CgpeError(rc, msg.decode() if msg else "", self._b.prefix + call)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._b.destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _per_replica(self, v):
        if v is None:
            return None
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.R,)))
        return a

    # -- state ---------------------------------------------------------------------------------
    def set_state(self, xyzq, vel=None, type=None):
        xyzq = np.ascontiguousarray(np.broadcast_to(np.asarray(xyzq, np.float32), (self.R, self.P, 4)))
        type = np.ascontiguousarray(np.broadcast_to(np.asarray(type, np.uint8), (self.R, self.P)))
        if vel is not None:
            vel = np.asarray(vel, np.float32)
            if vel.shape[-1] == 3:
                vel = np.concatenate([vel, np.zeros(vel.shape[:-1] + (1,), np.float32)], axis=-1)
            vel = np.ascontiguousarray(np.broadcast_to(vel, (self.R, self.P, 4)))
        self._check(self._b.set_state(self._h, _ptr(xyzq), _ptr(vel), _ptr(type)), "set_state")

    def get_state(self):
        xyzq = np.empty((self.R, self.P, 4), np.float32)
        vel = np.empty((self.R, self.P, 4), np.float32)
        type = np.empty((self.R, self.P), np.uint8)
        self._check(self._b.get_state(self._h, _ptr(xyzq), _ptr(vel), _ptr(type)), "get_state")
        return xyzq, vel, type

    def set_replica_params(self, pH=None, kappa=None, r_cut=None, kT=None, gamma=None):
        arrs = [self._per_replica(v) for v in (pH, kappa, r_cut, kT, gamma)]
        self._check(self._b.set_replica_params(self._h, *[_ptr(a) for a in arrs]), "set_replica_params")

    def set_wca(self, eps, sig):
        fe, fs = np.zeros((MAX_TYPES, MAX_TYPES)), np.zeros((MAX_TYPES, MAX_TYPES))
        eps, sig = np.asarray(eps, np.float64), np.asarray(sig, np.float64)
        fe[: eps.shape[0], : eps.shape[1]] = eps
        fs[: sig.shape[0], : sig.shape[1]] = sig
        self._check(self._b.set_wca(self._h, _ptr(fe), _ptr(fs)), "set_wca")

    def set_thermostat_seed(self, seed):
        self._check(self._b.set_thermostat_seed(self._h, int(seed)), "set_thermostat_seed")

    def set_force_cap(self, cap):
        self._check(self._b.set_force_cap(self._h, float(cap)), "set_force_cap")

    def set_time(self, t):
        self._check(self._b.set_time(self._h, float(t)), "set_time")

    def get_time(self):
        t = np.empty(self.R, np.float64)
        self._check(self._b.get_time(self._h, _ptr(t)), "get_time")
        return t

    def set_stream(self, cuda_stream):
        self._check(self._b.set_stream(self._h, C.c_void_p(cuda_stream or 0)), "set_stream")

    def synchronize(self):
        self._check(self._b.synchronize(self._h), "synchronize")

    def device_info(self):
        sm, maj, mnr = C.c_int32(), C.c_int32(), C.c_int32()
        name = C.create_string_buffer(256)
        self._check(self._b.device_info(self._h, C.byref(sm), C.byref(maj), C.byref(mnr), name, 256), "device_info")
        return {"sm_count": sm.value, "cc": (maj.value, mnr.value), "name": name.value.decode()}

    # -- hot path ------------------------------------------------------------------------------
    def md_run(self, n_steps):
        self._check(self._b.md_run(self._h, int(n_steps)), "md_run")

    def steepest_descent(self, n_steps, f_max=0.0, gamma=0.1, max_displacement=0.1):
        self._check(self._b.steepest_descent(self._h, int(n_steps), float(f_max), float(gamma),
                                             float(max_displacement)), "steepest_descent")

    def mc_run(self, reaction_id, n_trials, trace=False):
        rec = np.zeros((self.R, int(n_trials)), TRIAL_DTYPE) if trace else None
        self._check(self._b.mc_run(self._h, int(reaction_id), int(n_trials), _ptr(rec)), "mc_run")
        return rec

    @staticmethod
    def sample_params(n_samples, md_steps_per_burst=500, use_md=True, mc_blocks=(), q_snapshot_every=0,
                      q_snapshot_rows=0, schedule_seed=10, p_burst1=0.0, p_burst2=0.0):
        sp = SampleParams()
        sp.n_samples, sp.md_steps_per_burst, sp.use_md = int(n_samples), int(md_steps_per_burst), int(bool(use_md))
        sp.n_mc_blocks = len(mc_blocks)
        for b, (rx, nt) in enumerate(mc_blocks):
            sp.mc_reaction[b], sp.mc_trials[b] = int(rx), int(nt)
        sp.q_snapshot_every, sp.q_snapshot_rows = int(q_snapshot_every), int(q_snapshot_rows)
        sp.schedule_seed = int(schedule_seed)
        sp.p_burst1, sp.p_burst2 = float(p_burst1), float(p_burst2)
        return sp

    def sample_loop(self, sp):
        obs = np.empty((self.R, sp.n_samples, N_OBS), np.float64)
        qd = np.zeros((self.R, sp.q_snapshot_rows, self.n_beads), np.float32) if sp.q_snapshot_every > 0 and sp.q_snapshot_rows > 0 else None
        self._check(self._b.sample_loop(self._h, C.byref(sp), _ptr(obs), _ptr(qd)), "sample_loop")
        return obs, qd

    def sample_loop_async(self, sp):
        self._last_sp = (sp.n_samples, sp.q_snapshot_rows if sp.q_snapshot_every > 0 else 0)
        self._check(self._b.sample_loop_async(self._h, C.byref(sp)), "sample_loop_async")

    def fetch_samples(self):
        ns, rows = self._last_sp
        obs = np.empty((self.R, ns, N_OBS), np.float64)
        qd = np.zeros((self.R, rows, self.n_beads), np.float32) if rows > 0 else None
        self._check(self._b.fetch_samples(self._h, _ptr(obs), _ptr(qd)), "fetch_samples")
        return obs, qd

    # -- analysis ------------------------------------------------------------------------------
    def energy(self):
        out = np.empty((self.R, N_ENERGY), np.float64)
        self._check(self._b.energy(self._h, _ptr(out)), "energy")
        return out

    def forces(self):
        out = np.empty((self.R, self.P, 4), np.float32)
        self._check(self._b.forces(self._h, _ptr(out)), "forces")
        return out[..., :3]

    def md_forces(self):
        out = np.empty((self.R, self.P, 4), np.float32)
        self._check(self._b.md_forces(self._h, _ptr(out)), "md_forces")
        return out[..., :3]

    def observables(self):
        rg, ree = np.empty(self.R, np.float64), np.empty(self.R, np.float64)
        counts = np.empty((self.R, 3), np.int32)
        self._check(self._b.observables(self._h, _ptr(rg), _ptr(ree), _ptr(counts)), "observables")
        return rg, ree, counts

    def min_dist(self):
        out = np.empty(self.R, np.float64)
        self._check(self._b.min_dist(self._h, _ptr(out)), "min_dist")
        return out

    def counters(self):
        nr = max(self.n_reactions, 1)
        md = np.zeros(self.R, np.int64)
        tr, ac = np.zeros((self.R, nr), np.int64), np.zeros((self.R, nr), np.int64)
        rb = np.zeros(self.R, np.int64)
        self._check(self._b.get_counters(self._h, _ptr(md), _ptr(tr), _ptr(ac), _ptr(rb)), "get_counters")
        return {"md_steps": md, "trials": tr, "accepted": ac, "list_rebuilds": rb}

    def list_stats(self):
        """[R][4]: directed WCA list entries, in-cutoff ones, directed DH entries, charged in-cutoff ones."""
        out = np.empty((self.R, 4), np.float64)
        self._check(self._b.list_stats(self._h, _ptr(out)), "list_stats")
        return out

    ENGINES = {"auto": -1, "solo": 0, "duo": 1}

    def set_engine(self, engine):
        """'solo' (one CTA per replica), 'duo' (cluster of two CTAs per replica) or 'auto'."""
        self._check(self._b.set_engine(self._h, int(self.ENGINES.get(engine, engine))), "set_engine")

    def engine_info(self):
        eng = C.c_int32()
        cb = np.zeros(self.R, np.int64)
        ns = np.zeros((self.R, 2, 2), np.int64)
        self._check(self._b.engine_info(self._h, C.byref(eng), _ptr(cb), _ptr(ns)), "engine_info")
        return {"engine": ("solo", "duo", "wide")[eng.value] if 0 <= eng.value <= 2 else "oracle", "compact_builds": cb,
                "cta_ns": ns}

    # -- on-device accumulation (Observables.py:54-76) ---------------------------------------------
    def frames_begin(self, every, flags=0, capacity=1024):
        self._check(self._b.frames_begin(self._h, int(every), int(flags), int(capacity)), "frames_begin")
        self._frames = (int(flags), int(capacity))

    def frames_fetch(self):
        """(frames [R][n][8], bonds [R][n][N-1] or None, dihedrals [R][n][N-3] or None) per replica as lists of arrays:
        the frames recorded since the last fetch (COM x,y,z, COM velocity x,y,z, Rg, Ree)."""
        flags, cap = self._frames
        fr = np.empty((self.R, cap, FRAME_COLS), np.float64)
        bo = np.empty((self.R, cap, self.n_beads - 1), np.float32) if flags & FRAME_BONDS and self.n_beads > 1 else None
        di = np.empty((self.R, cap, self.n_beads - 3), np.float32) if flags & FRAME_DIHEDRALS and self.n_beads > 3 else None
        n = np.zeros(self.R, np.int32)
        self._check(self._b.frames_fetch(self._h, _ptr(fr), _ptr(bo), _ptr(di), _ptr(n)), "frames_fetch")
        cut = lambda a: None if a is None else [a[r, :n[r]] for r in range(self.R)]   # noqa: E731
        return cut(fr), cut(bo), cut(di)

    def frames_end(self):
        self._check(self._b.frames_end(self._h), "frames_end")

    def last_call_ms(self):
        ms, n = C.c_float(), C.c_int32()
        self._check(self._b.last_call_ms(self._h, C.byref(ms), C.byref(n)), "last_call_ms")
        return ms.value, n.value
