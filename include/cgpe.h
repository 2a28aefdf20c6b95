/*
 * cgpe.h — C-ABI of the B200-native constant-pH Monte Carlo + Langevin MD engine.
 *
 * The reference (Asureda/coarse_grain_polyelectrolyte) has no FFI of its own: every
 * floating-point operation of its hot path runs inside the ESPResSo engine, reached through
 * `espressomd` Python calls made from src/modules.  This header declares the entry points a
 * binding for that call surface needs; each one cites the reference call site it replaces.
 * Citations are relative to the reference checkout:
 *   P.py = src/modules/Properties.py   M.py = src/modules/Model.py
 *   E.py = src/modules/EnsembleDynamics.py
 *
 * Conventions
 *   - plain C, no C++/torch types; caller owns every host buffer, the library owns device memory;
 *   - every function returns CGPE_OK (0) or a negative cgpe_status; nothing throws across the
 *     boundary; cgpe_last_error() returns a human-readable message for the last failure;
 *   - a handle is bound to one CUDA device and is not thread-safe: one handle per GPU/process;
 *   - a handle holds R independent replicas of the same system (same topology, box, force
 *     field) that differ in pH, ionic strength (kappa, r_cut), temperature and RNG stream;
 *   - calls that copy results out synchronise the handle's stream, all others are asynchronous;
 *   - arrays are C-contiguous; [R] is the replica axis, [P] the particle axis
 *     (P = n_chains*n_beads + n_ion_slots).
 *
 * Reduced units are the reference's (P.py:260-265): energy kT, length l_B/2, charge e, time ps.
 */
#ifndef CGPE_H
#define CGPE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CGPE_API __attribute__((visibility("default")))
#else
#define CGPE_API
#endif

#define CGPE_ABI_VERSION 6
#define CGPE_MAX_TYPES 16     /* interacting particle types (the reference uses 8, M.py:58-76) */
#define CGPE_MAX_REACTIONS 4  /* reaction objects (the reference uses 2, M.py:211-233)         */
#define CGPE_MAX_MC_BLOCKS 4  /* MC blocks per sample (the reference uses 2, E.py:94-95)       */
#define CGPE_N_OBS 6          /* observable columns written per sample                         */
#define CGPE_N_ENERGY 5       /* total, kinetic, coulomb, bonded, non_bonded                   */
#define CGPE_FRAME_COLS 8     /* columns of a recorded frame: COM x,y,z, COM velocity x,y,z, Rg, Ree */
#define CGPE_FRAME_BONDS 1    /* cgpe_frames_begin flags: also record the N-1 bond lengths ...   */
#define CGPE_FRAME_DIHEDRALS 2 /* ... and the N-3 torsion angles of chain 0                      */

typedef enum cgpe_status {
  CGPE_OK = 0,
  CGPE_ERR_INVALID = -1,     /* bad argument / inconsistent configuration                      */
  CGPE_ERR_CUDA = -2,        /* CUDA runtime failure (message holds cudaGetErrorString)        */
  CGPE_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve (size, mode)             */
  CGPE_ERR_CAPACITY = -4,    /* a neighbour list or ion pool overflowed its capacity           */
  CGPE_ERR_STATE = -5,       /* call order violated (e.g. run before set_state)                */
  CGPE_ERR_NUMERIC = -6      /* non-finite coordinate / broken FENE bond detected on device    */
} cgpe_status;

enum { CGPE_BOND_NONE = 0, CGPE_BOND_HARMONIC = 1, CGPE_BOND_FENE = 2 };

/* One constant-pH reaction object: HA <-> A + B.
 * Replaces espressomd.reaction_methods.ConstantpHEnsemble(kT, exclusion_range, seed, constant_pH)
 * + .add_reaction(gamma, reactant_types, product_types, default_charges)   (M.py:211-255). */
typedef struct cgpe_reaction {
  int32_t type_prot;      /* reactant_types[0], "NH"  (M.py:237)                               */
  int32_t type_deprot;    /* product_types[0],  "N"   (M.py:238)                               */
  int32_t type_ion;       /* product_types[1],  "B"   (M.py:238); < 0 = implicit ions          */
  int32_t seed;           /* seed=77 (M.py:214)                                                */
  double q_prot;          /* default_charges (M.py:239-243)                                    */
  double q_deprot;
  double q_ion;
  double pK;              /* gamma = 10^-pK (M.py:236)                                         */
  double kT;              /* kT of the acceptance rule (M.py:212)                              */
  double exclusion_range; /* M.py:213; only meaningful with explicit ions                      */
} cgpe_reaction;

/* Static description of the system shared by all replicas of a handle. */
typedef struct cgpe_config {
  int32_t abi_version;    /* must be CGPE_ABI_VERSION                                          */
  int32_t n_replicas;     /* R held by this handle                                             */
  int32_t replica_offset; /* global index of replica 0: RNG streams are keyed on the global
                             index so results do not depend on how replicas shard over GPUs   */
  int32_t n_chains;       /* "Number of polymers" (P.py:384)                                   */
  int32_t n_beads;        /* "Number of monomers" per chain (P.py:390)                         */
  int32_t n_ion_slots;    /* particles per replica behind the chains.  With reactions that name an ion
                             type (type_ion >= 0): the explicit-ion pool (free H+, salt, hidden spares).
                             With type_ion < 0 in every reaction: implicit-ion reactions, the slots hold
                             spectator particles (e.g. one fixed charged nanoparticle); 0 = none      */
  int32_t n_types;        /* T interacting types; type index T is the non-interacting type
                             (M.py:256-261)                                                    */
  int32_t bond_kind;      /* HarmonicBond / FeneBond (M.py:34-40)                              */
  int32_t use_angle;      /* AngleHarmonic on (i-1,i,i+1) (M.py:42-46,161-165)                 */
  int32_t use_dihedral;   /* Dihedral on (i-1,i,i+1,i+2)  (M.py:48-54,166-170)                 */
  int32_t dihedral_mult;  /* mult=3 (M.py:51)                                                  */
  int32_t use_wca;        /* M.py:57                                                           */
  int32_t use_dh;         /* electrostatics.DH (M.py:97-107)                                   */
  int32_t n_reactions;
  int32_t cap_wca;        /* WCA neighbour-list row capacity, 0 = library default              */
  int32_t cap_dh;         /* Debye-Hueckel row capacity of the large-replica engine (P > 2048);
                             the shared-memory engine keeps a fixed-size bit matrix instead    */
  int32_t device;         /* CUDA ordinal, -1 = current device                                 */
  int32_t fixed_type_mask; /* bit t set: particles of type t never move (ESPResSo's `fix=[True]*3`): zero
                             force and velocity, no thermostat noise.  They still act on every other
                             particle: a charged nanoparticle among the chains (BASELINE config 4)     */
  double bond_k, bond_r0, fene_drmax;   /* M.py:35-39                                          */
  double angle_k, angle_phi0;           /* M.py:43-45                                          */
  double dihedral_k, dihedral_phase;    /* M.py:49-53                                          */
  double dh_prefactor;                  /* l_B kT / e^2 in reduced units (M.py:98-104)         */
  double box_l;                         /* cubic periodic box (P.py:372,375)                   */
  double time_step;                     /* P.py:373                                            */
  double skin;                          /* Verlet skin of this library's own lists (not the
                                           reference's cell_system.skin, which only tunes
                                           ESPResSo's internals); 0 = library default          */
  /* defaults broadcast to every replica; override with cgpe_set_replica_params                */
  double pH, kappa, r_cut, kT, gamma;
  /* The torsion gradient grows like 1/sin(theta) when a bond angle theta of the quadruple
     straightens (ESPResSo's expression shares this and only drops the term for |b1 x b2| <= 1e-4).
     With soft angle terms straight angles are thermally reachable and the force diverges.  The
     gradient is therefore evaluated with sin(theta) >= dihedral_sin_min.  0 = library default
     (0.02, i.e. within 1.1 degrees of collinear); negative = ESPResSo's guard only.               */
  double dihedral_sin_min;
  double wca_eps[CGPE_MAX_TYPES * CGPE_MAX_TYPES]; /* row-major [t1][t2] (M.py:74-76)          */
  double wca_sig[CGPE_MAX_TYPES * CGPE_MAX_TYPES];
  cgpe_reaction reactions[CGPE_MAX_REACTIONS];
} cgpe_config;

/* Schedule of EnsembleDynamics.perform_sampling (E.py:88-106). */
typedef struct cgpe_sample_params {
  int32_t n_samples;          /* num_samples (P.py:360-362)                                    */
  int32_t md_steps_per_burst; /* number_integration_steps = 500 (P.py:369)                     */
  int32_t use_md;             /* the USE_WCA gate of E.py:89,91                                */
  int32_t n_mc_blocks;
  int32_t mc_reaction[CGPE_MAX_MC_BLOCKS]; /* reaction1, reaction2 (E.py:94-95)                */
  int32_t mc_trials[CGPE_MAX_MC_BLOCKS];   /* 5*n_nh+1, 5*n_nh2+1                              */
  int32_t q_snapshot_every;   /* n_samp_iter (E.py:103); 0 = no charge snapshots               */
  int32_t q_snapshot_rows;    /* rows available in the qdist output                            */
  int32_t schedule_seed;      /* np.random.seed(10) (P.py:376)                                 */
  int32_t reserved0;
  double p_burst1;            /* probability_reaction (E.py:89)                                */
  double p_burst2;            /* probability_reaction * 2/(n_mon-2) (E.py:91)                  */
} cgpe_sample_params;

/* One trial move as the device executed it (parity / audit log). */
typedef struct cgpe_trial_record {
  int32_t particle;   /* id of the site whose type/charge was swapped, -1 if no reactant       */
  int32_t ion_slot;   /* explicit ions: particle id of the inserted / hidden ion, else -1      */
  int8_t direction;   /* +1 forward (HA -> A + B), -1 backward                                 */
  int8_t accepted;
  int8_t excluded;    /* exclusion range touched -> E_new = +max                                */
  int8_t reserved;
  float delta_e;      /* E_pot,new - E_pot,old                                                 */
  float bf;           /* acceptance probability before min(1, .)                               */
  float u;            /* the uniform it was compared with                                      */
} cgpe_trial_record;

typedef struct cgpe_handle cgpe_handle;

/* ---- lifetime ------------------------------------------------------------------------------ */
CGPE_API int cgpe_abi_version(void);
/* espressomd.System(box_l=) + bonded_inter.add + non_bonded_inter[..].wca.set_params +
 * electrostatics.DH + ConstantpHEnsemble  (P.py:372-375, M.py:32-109,200-261) */
CGPE_API int cgpe_create(const cgpe_config *cfg, cgpe_handle **out);
CGPE_API void cgpe_destroy(cgpe_handle *h);
/* message of the last failing call on h (or of the last failing cgpe_create if h == NULL) */
CGPE_API const char *cgpe_last_error(const cgpe_handle *h);
/* SM count, compute capability and name of the device h lives on */
CGPE_API int cgpe_device_info(const cgpe_handle *h, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor,
                     char *name, int32_t name_len);
/* run on a caller-provided cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream);
 * NULL restores the handle's own stream */
CGPE_API int cgpe_set_stream(cgpe_handle *h, void *cuda_stream);
CGPE_API int cgpe_synchronize(cgpe_handle *h);

/* ---- state --------------------------------------------------------------------------------- */
/* system.part.add(id,pos,type,q) (M.py:130-155,175-198).  xyzq [R][P][4] = x,y,z,q (positions
 * unfolded), vel [R][P][4] = vx,vy,vz,unused (NULL = zero), type [R][P].  Resets the reaction
 * bookkeeping lists to id order and invalidates cached forces. */
CGPE_API int cgpe_set_state(cgpe_handle *h, const float *xyzq, const float *vel, const uint8_t *type);
/* system.part.by_ids(ids).q (E.py:105) and positions/velocities read-back; any pointer may be NULL */
CGPE_API int cgpe_get_state(cgpe_handle *h, float *xyzq, float *vel, uint8_t *type);
/* reactionX.constant_pH = pH (E.py:68-69), DH(kappa, r_cut) (M.py:105-106), thermostat kT/gamma
 * (E.py:35,57,61-63), per replica [R]; NULL leaves a parameter unchanged */
CGPE_API int cgpe_set_replica_params(cgpe_handle *h, const double *pH, const double *kappa,
                            const double *r_cut, const double *kT, const double *gamma);
/* non_bonded_inter[t1,t2].wca.set_params(epsilon=, sigma=) for the whole table (M.py:74-76) */
CGPE_API int cgpe_set_wca(cgpe_handle *h, const double *eps, const double *sig);
/* thermostat.set_langevin(seed=) (E.py:35,62): negative seed keeps the current seed/counter */
CGPE_API int cgpe_set_thermostat_seed(cgpe_handle *h, int32_t seed);
/* system.force_cap (E.py:32,48,51); 0 disables */
CGPE_API int cgpe_set_force_cap(cgpe_handle *h, double cap);
/* system.time (E.py:30,77,102), per replica [R] */
CGPE_API int cgpe_set_time(cgpe_handle *h, double t);
CGPE_API int cgpe_get_time(cgpe_handle *h, double *t);

/* ---- hot path ------------------------------------------------------------------------------ */
/* system.integrator.run(n) after set_vv() with the Langevin thermostat
 * (E.py:41,52,58,64,73,90,92): velocity Verlet, friction -gamma v, uniform noise,
 * harmonic/FENE + angle + dihedral + WCA + Debye-Hueckel forces */
CGPE_API int cgpe_md_run(cgpe_handle *h, int32_t n_steps);
/* integrator.set_steepest_descent(f_max, gamma, max_displacement); integrator.run(n) (M.py:78-79) */
CGPE_API int cgpe_steepest_descent(cgpe_handle *h, int32_t n_steps, double f_max, double gamma,
                          double max_displacement);
/* reactionX.reaction(reaction_steps=n) (E.py:70-71,94-95).  trace may be NULL, else
 * [R][n_trials] records */
CGPE_API int cgpe_mc_run(cgpe_handle *h, int32_t reaction_id, int32_t n_trials, cgpe_trial_record *trace);
/* the whole loop body of perform_sampling for n_samples samples (E.py:88-106), fused on device.
 * obs [R][n_samples][CGPE_N_OBS] = n_deprot(reaction 0), n_deprot(reaction 1), n_ion, Rg, Ree,
 * time (E.py:97-102); qdist [R][q_snapshot_rows][n_beads] charge snapshots (E.py:103-106) or NULL */
CGPE_API int cgpe_sample_loop(cgpe_handle *h, const cgpe_sample_params *sp, double *obs, float *qdist);
/* same, leaving the results in device memory (read them later with cgpe_fetch_samples) */
CGPE_API int cgpe_sample_loop_async(cgpe_handle *h, const cgpe_sample_params *sp);
CGPE_API int cgpe_fetch_samples(cgpe_handle *h, double *obs, float *qdist);

/* ---- analysis ------------------------------------------------------------------------------ */
/* system.analysis.energy() (E.py:42): out [R][5] = total, kinetic, coulomb, bonded, non_bonded */
CGPE_API int cgpe_energy(cgpe_handle *h, double *out);
/* interaction forces without thermostat, all-pairs evaluation: out [R][P][4] (w unused) */
CGPE_API int cgpe_forces(cgpe_handle *h, float *out);
/* the same forces as the MD kernel computes them (Verlet lists, owner-computes bonded terms) */
CGPE_API int cgpe_md_forces(cgpe_handle *h, float *out);
/* analysis.calc_rg / calc_re of chain 0 (E.py:100-101) and number_of_particles(type)
 * (E.py:97-99): rg [R], ree [R], counts [R][3] = n_deprot(reaction 0), n_deprot(reaction 1), n_ion */
CGPE_API int cgpe_observables(cgpe_handle *h, double *rg, double *ree, int32_t *counts);
/* analysis.min_dist() (E.py:34,45): out [R] */
CGPE_API int cgpe_min_dist(cgpe_handle *h, double *out);
/* bookkeeping totals per replica: md_steps [R], trials [R][n_reactions], accepted [R][n_reactions],
 * list_rebuilds [R]; any pointer may be NULL */
CGPE_API int cgpe_get_counters(cgpe_handle *h, int64_t *md_steps, int64_t *trials, int64_t *accepted,
                      int64_t *list_rebuilds);
/* Verlet-list census of the current configuration, out [R][4] = directed WCA list entries,
 * those inside their cutoff, directed Debye-Hueckel list entries, those between two charged
 * particles inside r_cut.  (Unique pairs = directed / 2.)  Feeds the roofline arithmetic. */
CGPE_API int cgpe_list_stats(cgpe_handle *h, double *out);
/* Which kernels serve integrator.run / reaction.reaction / the fused sampling loop (E.py:88-95) of this handle:
 *   0 = one persistent CTA per replica (shared-memory engine), 1 = one thread-block cluster of two CTAs per
 *   replica (implicit-ion replicas of 8..2048 chain beads without spectator particles), -1 = the library
 *   chooses (default: the cluster engine wherever it applies).  Replicas beyond 2048 particles always run on
 *   the global-memory engine.  State between calls is engine-agnostic; the Markov chain is the same. */
CGPE_API int cgpe_set_engine(cgpe_handle *h, int32_t engine);
/* engine: the one in force (0, 1 as above, 2 = global-memory engine); compact_builds [R] (may be NULL): how
 * often each replica wrote the compact Debye-Hueckel partner lists its MD force loop walks -- lets a test
 * assert that it measured the code path the benchmark times.  cta_ns [R][2][2] (may be NULL): start and end,
 * in ns of the GPU's global timer, of the CTA(s) that ran each replica in the last fused sampling launch
 * (second CTA zero for the one-CTA engines): the load-balance picture of a launch. */
CGPE_API int cgpe_engine_info(cgpe_handle *h, int32_t *engine, int64_t *compact_builds, int64_t *cta_ns);
/* On-device accumulation for the observables the reference registers with system.auto_update_accumulators
 * (Observables.py:54-76: ParticleDistances / BondDihedrals TimeSeries, ComPosition / ComVelocity correlators, delta_N = 100):
 * after cgpe_frames_begin every `every`-th MD step of a replica -- inside cgpe_md_run and inside the fused cgpe_sample_loop,
 * without leaving the launch -- appends one frame of chain 0 to device buffers: COM position, COM velocity, Rg, Ree
 * (CGPE_FRAME_COLS doubles) and, on request (flags), the N-1 bond lengths and the N-3 torsion angles (float).  `capacity` =
 * frames per replica the buffers hold; frames beyond it are dropped and reported by the next synchronising call
 * (CGPE_ERR_CAPACITY).  cgpe_frames_fetch copies out what was recorded since the last fetch -- frames [R][capacity][8],
 * bonds [R][capacity][N-1], dihedrals [R][capacity][N-3], n_frames [R]; any pointer may be NULL -- and empties the buffers;
 * the step phase of the recorder carries over, also across a cgpe_frames_begin with the same `every` (a larger capacity).
 * Shared-memory engine (<= 2048 particles per replica). */
CGPE_API int cgpe_frames_begin(cgpe_handle *h, int32_t every, int32_t flags, int32_t capacity);
CGPE_API int cgpe_frames_fetch(cgpe_handle *h, double *frames, float *bonds, float *dihedrals, int32_t *n_frames);
CGPE_API int cgpe_frames_end(cgpe_handle *h);
/* device time of the last hot-path call in milliseconds (CUDA events on the handle's stream) and
 * the number of kernels it launched */
CGPE_API int cgpe_last_call_ms(cgpe_handle *h, float *ms, int32_t *n_launches);

#ifdef __cplusplus
}
#endif
#endif /* CGPE_H */
