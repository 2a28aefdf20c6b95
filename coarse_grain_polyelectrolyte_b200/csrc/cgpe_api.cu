// cgpe_api.cu — the C-ABI of libcgpe.so (include/cgpe.h): handle management, validation, host
// <-> device copies and kernel launches.  No torch types, no C++ types across the boundary.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "duo_engine.cuh"
#include "wide_engine.cuh"

using namespace cgpe;

struct cgpe_handle {
  cgpe_config cfg;
  int *d_choice = nullptr;     // [1] engine the plan picked for the last fused sampling launch issued on both shared-memory engines
  bool choice_live = false;    // ... and whether that launch was such a one
  int *d_order = nullptr;      // [R] CTA -> replica of the one-CTA engine's fused sampling launch (duo_plan_waves)
  bool solo_cap_low = false;   // the one-CTA engine's shared memory leaves too few WCA row entries for this replica size (choose_capacities)
  DevParams p;
  DevState g;
  ChainLaunch launch;
  WideState w;          // global-memory engine (replicas beyond the shared-memory engine)
  bool wide;
  std::vector<void *> wide_allocs;
  std::vector<int> chg_host;   // [R][n_chg] chargeable particle ids (host copy)
  WideGraphCache wgraph;       // captured block of MD steps of the global-memory engine
  void *dhc_alloc = nullptr;   // DevState::dhc
  size_t dhc_bytes = 0;
  size_t smem_optin = 0;       // cudaDevAttrMaxSharedMemoryPerBlockOptin, read once (cudaGetDeviceProperties is slow,
                               // and slower still when eight processes ask at once)
  bool two_radii = false;      // see upload_wca
  DuoShape duo;                // cluster engine (two CTAs per replica): shape, or duo.ok == 0
  int engine_pref = -1;        // cgpe_set_engine: -1 = choose, 0 = one CTA per replica, 1 = cluster engine
  int n_sm = 0;
  size_t dhc_need_solo = 0;
  void *frame_alloc[3] = {nullptr, nullptr, nullptr};   // DevState::frames, frame_bonds, frame_dih
  void *mc_e_alloc = nullptr;  // DevState::mc_e
  size_t mc_e_bytes = 0;
  cudaStream_t own_stream, stream;
  cudaEvent_t ev0, ev1;
  int device;
  bool have_state;
  bool timing_pending;
  int last_launches;
  // scratch / result buffers on device
  double *d_obs; float *d_qdist; size_t obs_cap, qdist_cap; int last_ns, last_qrows;
  double *d_tmp_d; int *d_tmp_i; float4 *d_tmp_f4;     // [R*8] doubles, [R*3] ints, [R][P] float4
  cgpe_trial_record *d_trace; size_t trace_cap;
  std::vector<void *> allocs;
  char err[512];
};

static char g_err[512] = "";

static int fail(cgpe_handle *h, int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(h ? h->err : g_err, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) return fail(h, CGPE_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

namespace {

struct DeviceGuard {
  int prev;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (dev != prev) cudaSetDevice(dev); }
  ~DeviceGuard() { int cur; cudaGetDevice(&cur); if (cur != prev) cudaSetDevice(prev); }
};

template <typename T>
cudaError_t dev_alloc(cgpe_handle *h, T **ptr, size_t n) {
  cudaError_t e = cudaMalloc((void **)ptr, sizeof(T) * (n > 0 ? n : 1));
  if (e == cudaSuccess) {
    h->allocs.push_back((void *)*ptr);
    e = cudaMemset(*ptr, 0, sizeof(T) * (n > 0 ? n : 1));
  }
  return e;
}

int validate(const cgpe_config *c) {
  if (!c) return fail(nullptr, CGPE_ERR_INVALID, "config is NULL");
  if (c->abi_version != CGPE_ABI_VERSION)
    return fail(nullptr, CGPE_ERR_INVALID, "abi_version %d != %d", c->abi_version, CGPE_ABI_VERSION);
  if (c->n_replicas < 1 || c->n_chains < 1 || c->n_beads < 1 || c->n_ion_slots < 0)
    return fail(nullptr, CGPE_ERR_INVALID, "n_replicas, n_chains, n_beads must be >= 1");
  if (c->n_types < 1 || c->n_types > CGPE_MAX_TYPES) return fail(nullptr, CGPE_ERR_INVALID, "n_types out of range");
  if (c->fixed_type_mask < 0 || (c->fixed_type_mask >> c->n_types) != 0)
    return fail(nullptr, CGPE_ERR_INVALID, "fixed_type_mask names a type >= n_types");
  if (c->n_reactions < 0 || c->n_reactions > CGPE_MAX_REACTIONS)
    return fail(nullptr, CGPE_ERR_INVALID, "n_reactions out of range");
  if (!(c->box_l > 0.0) || !(c->time_step > 0.0))
    return fail(nullptr, CGPE_ERR_INVALID, "box_l and time_step must be positive");
  if (c->use_dh && (!(c->r_cut > 0.0) || c->r_cut > 0.5 * c->box_l))
    return fail(nullptr, CGPE_ERR_INVALID, "r_cut must lie in (0, box_l/2]");
  if (c->bond_kind < CGPE_BOND_NONE || c->bond_kind > CGPE_BOND_FENE)
    return fail(nullptr, CGPE_ERR_INVALID, "unknown bond_kind");
  if (c->use_dihedral && (c->dihedral_mult < 0 || c->dihedral_mult > 12))
    return fail(nullptr, CGPE_ERR_INVALID, "dihedral_mult out of range");
  for (int m = 0; m < c->n_reactions; ++m) {
    const cgpe_reaction &r = c->reactions[m];
    if (r.type_prot < 0 || r.type_prot >= c->n_types || r.type_deprot < 0 || r.type_deprot >= c->n_types ||
        r.type_prot == r.type_deprot)
      return fail(nullptr, CGPE_ERR_INVALID, "reaction %d: bad type indices", m);
    if (!(r.kT > 0.0)) return fail(nullptr, CGPE_ERR_INVALID, "reaction %d: kT must be positive", m);
    for (int m2 = 0; m2 < m; ++m2) {
      const cgpe_reaction &q = c->reactions[m2];
      if (q.type_prot == r.type_prot || q.type_prot == r.type_deprot || q.type_deprot == r.type_prot ||
          q.type_deprot == r.type_deprot)
        return fail(nullptr, CGPE_ERR_INVALID, "reactions %d and %d share a particle type", m2, m);
    }
  }
  return CGPE_OK;
}

// (re)build the WCA table {24 eps, sigma^2, rc^2, eps} incl. the non-interacting row/column, and
// the list radius that depends on it
int upload_wca(cgpe_handle *h) {
  const int T = h->cfg.n_types, S = T + 1;
  std::vector<float4> tab((size_t)S * S, make_float4(0.f, 0.f, 0.f, 0.f));
  double rc_max = 0.0, rc_free = 0.0;   // over all type pairs / over pairs of types that move
  const unsigned fixed = (unsigned)h->cfg.fixed_type_mask;
  for (int a = 0; a < T; ++a)
    for (int b = 0; b < T; ++b) {
      const double eps = h->cfg.wca_eps[a * CGPE_MAX_TYPES + b], sig = h->cfg.wca_sig[a * CGPE_MAX_TYPES + b];
      if (eps < 0.0 || sig < 0.0) return fail(h, CGPE_ERR_INVALID, "WCA epsilon and sigma must be >= 0");
      if (eps == 0.0 || sig == 0.0) continue;
      const double rc2 = 1.2599210498948731648 * sig * sig;
      tab[(size_t)a * S + b] = make_float4((float)(24.0 * eps), (float)(sig * sig), (float)rc2, (float)eps);
      rc_max = std::fmax(rc_max, std::sqrt(rc2));
      if (!((fixed >> a) & 1u) && !((fixed >> b) & 1u)) rc_free = std::fmax(rc_free, std::sqrt(rc2));
    }
  CU(cudaMemcpyAsync(h->g.wca_tab, tab.data(), sizeof(float4) * tab.size(), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  const double rl = rc_max + h->p.skin;
  if (h->cfg.use_wca && rl > 0.5 * h->cfg.box_l) return fail(h, CGPE_ERR_INVALID, "WCA cutoff + skin exceeds box_l/2");
  h->p.rlist_wf2 = (float)(rl * rl);
  // one big fixed sphere must not lengthen every row: when the fixed particles all sit in slots and the
  // shared-memory engine serves the handle (set_state decides both), the rows take free partners within
  // the free-pair radius and fixed ones within the full radius (chain_engine.cu::build_rows)
  const double rl_free = h->two_radii ? rc_free + h->p.skin : rl;
  h->p.rlist_w2 = (float)(rl_free * rl_free);
  h->p.rc_w_max2 = (float)(rc_max * rc_max * (1.0 + 1e-6));
  return CGPE_OK;
}

// List shapes.  The Debye-Hueckel list is a bit matrix over the chargeable particles (fixed size);
// the WCA row capacity takes what is left of the 227 KB shared-memory budget, up to 64 entries,
// unless the config gives one.
constexpr long kMinAutoCapW = 12;

int choose_capacities(cgpe_handle *h) {
  h->solo_cap_low = false;
  DevParams &p = h->p;
  const int n_chg = p.n_chg;
  p.dw = ((n_chg + 31) / 32) | 1;   // odd row stride: thread-per-row reads hit distinct banks
  {
    // sub-threads per Debye-Hueckel row: as many as the block has threads for, at most 4
    // (counted on one thread per particle: with two particles per thread a thread takes two sub-rows)
    const int threads = ((std::min(p.P, 1024) + 31) / 32) * 32;
    int grp = n_chg > 0 ? threads / n_chg : 1;
    p.dh_group = grp < 1 ? 1 : grp > 4 ? 4 : grp;
  }
  if (h->cfg.cap_wca > 0) p.cap_w = h->cfg.cap_wca;
  else {
    p.cap_w = 0;
    const size_t fixed = chain_smem_bytes(p), budget = h->smem_optin;
    long cap = budget > fixed + 64 ? (long)((budget - fixed - 64) / (2 * (size_t)p.P)) : 0;
    if (cap > 64) cap = 64;
    // Rows of the example chains hold 6-7 entries on average, twice that at most.  Near 2000 particles the fixed arrays
    // (Debye-Hueckel bit matrix ~ P^2/72 bytes, slots, positions) leave room for a handful only: rows overflow, pairs are
    // dropped, the replica blows up.  Then the cluster engine (each CTA holds rows for half the particles) or the
    // global-memory engine serves the replica; cgpe_set_engine(solo) or an explicit cap_wca still force this one.
    h->solo_cap_low = cap < kMinAutoCapW;
    if (cap < 4) cap = 4;   // too little: chain_launch_shape / the smem check will refuse
    p.cap_w = (int)cap;
  }
  if (p.cap_w > (p.P > 1 ? p.P - 1 : 1)) p.cap_w = p.P > 1 ? p.P - 1 : 1;
  chain_smem_bytes(p);   // final layout -> p.off
  h->launch = chain_launch_shape(p);
  return CGPE_OK;
}

bool use_duo(const cgpe_handle *h);

int check_device_errors(cgpe_handle *h) {
  if (h->p.frame_every > 0) {
    std::vector<int> nf(h->p.R);
    CU(cudaMemcpyAsync(nf.data(), h->g.n_frames, sizeof(int) * h->p.R, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int r = 0; r < h->p.R; ++r)
      if (nf[r] > h->p.frame_cap)
        return fail(h, CGPE_ERR_CAPACITY, "frame recorder overflow: replica %d recorded %d frames, capacity %d (cgpe_frames_begin); fetch more often",
                    r, nf[r], h->p.frame_cap);
  }
  std::vector<int> err(h->p.R);
  CU(cudaMemcpyAsync(err.data(), h->g.err, sizeof(int) * h->p.R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  int any = 0, first = -1;
  for (int r = 0; r < h->p.R; ++r)
    if (err[r]) { any |= err[r]; if (first < 0) first = r; }
  if (!any) return CGPE_OK;
  CU(cudaMemsetAsync(h->g.err, 0, sizeof(int) * h->p.R, h->stream));
  if (any & kErrWcaCap)   // first: dropped pairs are what makes a replica blow up afterwards
    return fail(h, CGPE_ERR_CAPACITY, "WCA neighbour list overflow (cap_wca=%d, first replica %d): raise cap_wca or lower skin",
                h->wide ? h->w.cap_w : (use_duo(h) ? h->duo.cap_w : h->p.cap_w), first);
  if (any & kErrNonFinite) return fail(h, CGPE_ERR_NUMERIC, "non-finite coordinate or velocity (first replica %d)", first);
  if (any & kErrFene) return fail(h, CGPE_ERR_NUMERIC, "FENE bond stretched beyond d_r_max (first replica %d)", first);
  if (any & kErrDiverged)
    return fail(h, CGPE_ERR_NUMERIC, "replica %d diverged (a velocity exceeds 1000 sqrt(kT/m)): the integration is unstable, reduce the time step", first);
  if (any & kErrIonPool)
    return fail(h, CGPE_ERR_CAPACITY, "explicit-ion pool exhausted (n_ion_slots=%d, first replica %d): add spare hidden slots", h->p.n_ions, first);
  if (any & kErrDhCap)
    return fail(h, CGPE_ERR_CAPACITY, "Debye-Hueckel neighbour list overflow (cap_dh=%d, first replica %d): raise cap_dh", h->w.cap_d, first);
  return fail(h, CGPE_ERR_CAPACITY, "WCA neighbour list overflow (cap_wca=%d, first replica %d): raise cap_wca or lower skin",
              h->wide ? h->w.cap_w : h->p.cap_w, first);
}

int require_engine(cgpe_handle *h) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  if (h->launch.items == 0 && !h->wide)
    return fail(h, CGPE_ERR_UNSUPPORTED, "P=%d particles per replica: no engine available", h->p.P);
  return CGPE_OK;
}

// The cluster engine serves implicit-ion replicas without spectator particles.  Its step carries two cluster barriers
// and the mirrored position stores (~1.2 us), which it earns back only where the step is long and the launch leaves SMs
// idle (measured, tools/prof_engines.py -> profiles/r2_engines.txt, final library: 128 x N=1000 at 1 mM 1.04x, 64 x N=1000
// 1.04x, 20 mM 1.01x; but 0.65x at N=100, 0.77x at N=300, 0.85x with 256 replicas).  So the library considers it for long
// Debye-Hueckel rows (list radius >= 0.6 of the ideal-chain extent), chains of >= 512 beads and no more replicas than SMs;
// for a fused sampling launch the plan then decides on the device (cgpe_sample_loop, k_duo_plan).  cgpe_set_engine or
// CGPE_ENGINE=solo|duo overrides.
// kernels of one cluster-engine call: the engine kernel, the plan, and the pair count when the plan has a choice (duo_engine.cu::launch_cluster)
int duo_launches(const cgpe_handle *h) { return 2 + ((h->p.use_dh && h->p.R > h->duo.n_sm / 2 && h->p.R <= h->duo.n_sm) ? 1 : 0); }

bool use_duo(const cgpe_handle *h) {
  if (h->wide || !h->duo.ok || h->launch.items == 0) return false;
  if (h->p.frame_every > 0) return false;   // the frame recorder lives in the one-CTA engine
  if (h->engine_pref >= 0) return h->engine_pref == 1;
  if (h->solo_cap_low) return true;   // the one-CTA engine would overflow its WCA rows
  const DevParams &p = h->p;
  const double b = h->cfg.bond_r0 > 0.0 ? h->cfg.bond_r0 : 1.0;
  const bool long_rows = p.use_dh && h->cfg.r_cut + p.skin >= 0.6 * b * std::sqrt((double)p.N);
  return p.P >= 512 && p.R <= h->n_sm && long_rows;
}

// shape of the cluster engine for the current state (after choose_capacities)
int plan_duo(cgpe_handle *h) {
  DevParams &p = h->p;
  h->duo.ok = 0;
  if (p.n_ions != 0 || p.P != p.NC || p.fixed_mask != 0u || p.P < 8 || p.P > 2048 || h->launch.items == 0) return CGPE_OK;
  const int half = (p.P + 1) / 2;
  int n_loc = 0;
  for (int r = 0; r < p.R; ++r) {
    int a = 0;
    for (int i = 0; i < p.n_chg; ++i) a += h->chg_host[(size_t)r * p.n_chg + i] < half ? 1 : 0;
    n_loc = std::max(n_loc, std::max(a, p.n_chg - a));
  }
  // two CTAs of up to 512 threads share an SM: each may take half of its shared memory (minus 1 KB the system keeps per CTA)
  const int nt = std::max(32, (half + 31) / 32 * 32);
  const size_t budget = nt <= 512 ? std::min(h->smem_optin, (size_t)(233472 / 2 - 1024)) : h->smem_optin;
  int *order = h->duo.order;
  duo_layout(p, n_loc, budget, h->cfg.cap_wca, &h->duo);
  h->duo.order = order;
  h->duo.n_sm = h->n_sm;
  if (h->duo.ok && !h->duo.order) {
    CU(cudaMalloc((void **)&h->duo.order, sizeof(int) * p.R));
    h->allocs.push_back((void *)h->duo.order);
  }
  return CGPE_OK;
}

void free_wide(cgpe_handle *h) {
  if (h->wgraph.exec) { cudaStreamSynchronize(h->stream); wide_graph_release(&h->wgraph); }
  for (void *ptr : h->wide_allocs) cudaFree(ptr);
  h->wide_allocs.clear();
  h->wide = false;
}

// How the Debye-Hueckel rows of the global-memory engine are walked: by 8 lanes when they are short.
// Expected entries = chargeable particles of a uniform box in the list sphere + those of a coil
// (Flory exponent 0.6) of the own chain.
void update_wide_policy(cgpe_handle *h) {
  const DevParams &p = h->p;
  const double rl = h->cfg.r_cut + p.skin, L = h->cfg.box_l;
  const double uniform = p.n_chg * 4.18879 * rl * rl * rl / (L * L * L);
  const double b = h->cfg.bond_r0 > 0.0 ? h->cfg.bond_r0 : 1.0;
  const double frac = p.NC > 0 ? std::fmax(0.0, (double)(p.n_chg - p.n_ions)) / p.NC : 0.0;
  const double coil = frac * std::fmin((double)p.N, std::pow(rl / b, 1.0 / 0.6));
  const int short_rows = uniform + coil < 48.0 ? 1 : 0;
  if (short_rows != h->w.dh_rows_short && h->wgraph.exec) { cudaStreamSynchronize(h->stream); wide_graph_release(&h->wgraph); }
  h->w.dh_rows_short = short_rows;
}

// arrays of the global-memory engine, sized for the current shapes
int ensure_wide(cgpe_handle *h, const std::vector<int> &chg_packed) {
  free_wide(h);
  const DevParams &p = h->p;
  WideState &w = h->w;
  std::memset(&w, 0, sizeof(w));
  w.cap_w = h->cfg.cap_wca > 0 ? h->cfg.cap_wca : 64;
  if (h->cfg.cap_dh > 0) w.cap_d = h->cfg.cap_dh;
  else {
    // HBM is plentiful (180 GB) and rows cost only what they hold: size them for compact starts
    int cap = (int)(256 + 24.0 * (h->cfg.r_cut + p.skin));
    w.cap_d = (cap + 7) / 8 * 8;
  }
  if (w.cap_d > (p.n_chg > 1 ? p.n_chg - 1 : 1)) w.cap_d = p.n_chg > 1 ? p.n_chg - 1 : 1;
  const size_t RP = (size_t)p.R * p.P, RC = (size_t)p.R * (p.n_chg > 0 ? p.n_chg : 1);
  auto alloc = [&](void **ptr, size_t bytes) {
    cudaError_t e = cudaMalloc(ptr, bytes > 0 ? bytes : 1);
    if (e == cudaSuccess) { h->wide_allocs.push_back(*ptr); e = cudaMemset(*ptr, 0, bytes > 0 ? bytes : 1); }
    return e;
  };
  CU(alloc((void **)&w.ref, sizeof(float4) * RP));
  CU(alloc((void **)&w.wl, sizeof(uint32_t) * RP * w.cap_w));
  CU(alloc((void **)&w.wn, sizeof(int) * RP));
  CU(alloc((void **)&w.dl, sizeof(uint32_t) * RC * w.cap_d));
  if (p.use_dh && p.n_chg >= 64 && p.n_chg < 65536 && sizeof(float4) * (size_t)p.n_chg <= 96 * 1024 && !std::getenv("CGPE_NO_DH_TILE"))
  {   // tile mode (wide_engine.cu): rows by chargeable index, all and charged-only
    CU(alloc((void **)&w.dlc, sizeof(uint16_t) * RC * w.cap_d));
    CU(alloc((void **)&w.dlq, sizeof(uint16_t) * RC * w.cap_d));
    CU(alloc((void **)&w.dnq, sizeof(int) * RC));
    CU(alloc((void **)&w.nowrap, sizeof(int) * p.R));
  }
  CU(alloc((void **)&w.dn, sizeof(int) * RC));
  CU(alloc((void **)&w.fdh, sizeof(float4) * RC));
  CU(alloc((void **)&w.cidx, sizeof(int) * RP));
  CU(alloc((void **)&w.flag, sizeof(int) * (p.R + 1)));
  CU(alloc((void **)&w.nsteps, sizeof(int) * p.R));
  CU(alloc((void **)&w.sched_max, sizeof(int)));
  CU(alloc((void **)&w.step, 2 * sizeof(int)));
  CU(alloc((void **)&w.phi, sizeof(float) * RC));
  CU(alloc((void **)&w.wcache, sizeof(float) * RC));
  CU(alloc((void **)&w.rxof, RC));
  CU(alloc((void **)&w.wflag, RC));
  CU(alloc((void **)&w.stale, RP));
  w.mc_grid_ok = 1;
  w.cells_multi_min = 32768;
  if (const char *e = std::getenv("CGPE_CELLS_MULTI_MIN")) w.cells_multi_min = std::atoi(e);   // diagnostics / tests
  w.mc_stage_limit = 220 * 1024;
  if (const char *e = std::getenv("CGPE_MC_STAGE_LIMIT")) w.mc_stage_limit = (size_t)std::atoll(e);   // diagnostics: force the
  if (const char *e = std::getenv("CGPE_MC_GRID")) w.mc_grid_ok = std::atoi(e);                        // other trial-loop variants
  // cell grids of the list build: at most ~2 cells per member, edge never below the list radius
  // (the kernels pick the cells per edge from the list radius in force at each build)
  auto grid = [&](CellGrid &cg, int n) -> cudaError_t {
    cg.n = n;
    int nc = (int)std::floor(std::cbrt(2.0 * (n > 0 ? n : 1)));
    if (const char *force = std::getenv("CGPE_GRID_NC_MAX")) nc = std::max(1, std::min(nc, std::atoi(force)));   // diagnostics only
    cg.nc_max = nc >= 3 ? nc : 1;
    cg.m_max = cg.nc_max * cg.nc_max * cg.nc_max;
    const size_t Rn = (size_t)p.R * (n > 0 ? n : 1);
    cudaError_t e;
    if ((e = alloc((void **)&cg.start, sizeof(int) * p.R * (cg.m_max + 1))) != cudaSuccess) return e;
    if ((e = alloc((void **)&cg.cursor, sizeof(int) * p.R * cg.m_max)) != cudaSuccess) return e;
    if ((e = alloc((void **)&cg.cell_of, sizeof(int) * Rn)) != cudaSuccess) return e;
    if ((e = alloc((void **)&cg.tmp, sizeof(int) * Rn)) != cudaSuccess) return e;
    if ((e = alloc((void **)&cg.order, sizeof(int) * Rn)) != cudaSuccess) return e;
    if ((e = alloc((void **)&cg.spos, sizeof(float4) * Rn)) != cudaSuccess) return e;
    if ((e = alloc((void **)&cg.nf, sizeof(int) * p.R)) != cudaSuccess) return e;
    return alloc((void **)&cg.nc, sizeof(int) * p.R);
  };
  CU(grid(w.cw, p.P));
  CU(grid(w.cd, p.n_chg));
  std::vector<int> cidx(RP, -1);
  for (int r = 0; r < p.R; ++r)
    for (int c = 0; c < p.n_chg; ++c) cidx[(size_t)r * p.P + chg_packed[(size_t)r * p.n_chg + c]] = c;
  CU(cudaMemcpy(w.cidx, cidx.data(), sizeof(int) * RP, cudaMemcpyHostToDevice));
  h->wide = true;
  update_wide_policy(h);
  return CGPE_OK;
}

void begin_timing(cgpe_handle *h) { cudaEventRecord(h->ev0, h->stream); h->last_launches = 0; }
void end_timing(cgpe_handle *h, int launches) {
  cudaEventRecord(h->ev1, h->stream);
  h->timing_pending = true;
  h->last_launches = launches;
}

uint32_t bernoulli_threshold(double p) {
  if (!(p > 0.0)) return 0u;
  if (p >= 1.0) return 16777216u;
  return (uint32_t)std::ceil(p * 16777216.0);
}

}  // namespace

extern "C" {

int cgpe_abi_version(void) { return CGPE_ABI_VERSION; }

const char *cgpe_last_error(const cgpe_handle *h) { return h ? h->err : g_err; }

// Slots behind the chains hold explicit ions that the reactions create and destroy (M.py:173-198) --
// unless every reaction names no ion type (type_ion < 0): then the reactions are the implicit-ion
// ones and the slots hold spectator particles (a charged nanoparticle, BASELINE config 4).
static bool explicit_ions(const cgpe_config *c) {
  if (c->n_ion_slots <= 0) return false;
  for (int m = 0; m < c->n_reactions; ++m)
    if (c->reactions[m].type_ion >= 0) return true;
  return c->n_reactions == 0;
}

int cgpe_create(const cgpe_config *cfg, cgpe_handle **out) {
  if (!out) return fail(nullptr, CGPE_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int rc = validate(cfg);
  if (rc) return rc;
  if (explicit_ions(cfg)) {
    double rc_ion = 0.0;   // largest WCA cutoff of the ion type
    const int tb = cfg->n_reactions > 0 ? cfg->reactions[0].type_ion : 0;
    for (int m = 0; m < cfg->n_reactions; ++m) {
      const cgpe_reaction &r = cfg->reactions[m];
      if (r.type_ion != tb || r.q_ion != cfg->reactions[0].q_ion || tb < 0 || tb >= cfg->n_types)
        return fail(nullptr, CGPE_ERR_INVALID, "explicit ions: all reactions must release the same ion type");
      if (tb == r.type_prot || tb == r.type_deprot) return fail(nullptr, CGPE_ERR_INVALID, "explicit ions: the ion type must differ from the site types");
    }
    if (cfg->use_wca && cfg->n_reactions > 0)
      for (int t = 0; t < cfg->n_types; ++t)
        if (cfg->wca_eps[tb * CGPE_MAX_TYPES + t] > 0.0)
          rc_ion = std::fmax(rc_ion, 1.122462048309373 * cfg->wca_sig[tb * CGPE_MAX_TYPES + t]);
    for (int m = 0; m < cfg->n_reactions; ++m)
      if (cfg->reactions[m].exclusion_range < rc_ion)
        return fail(nullptr, CGPE_ERR_UNSUPPORTED,
                    "reaction %d: exclusion_range %g is below the ion's largest WCA cutoff %g (an inserted ion would "
                    "carry WCA energy, which this build does not evaluate)", m, cfg->reactions[m].exclusion_range, rc_ion);
  }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, CGPE_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path",
                cudaGetErrorString(ce));
  int dev = cfg->device;
  if (dev < 0) cudaGetDevice(&dev);
  if (dev >= ndev) return fail(nullptr, CGPE_ERR_INVALID, "device %d out of range (%d devices)", dev, ndev);

  cgpe_handle *h = new (std::nothrow) cgpe_handle();
  if (!h) return fail(nullptr, CGPE_ERR_CUDA, "out of host memory");
  h->cfg = *cfg;
  h->device = dev;
  h->err[0] = 0;
  DeviceGuard guard(dev);

  {
    int optin = 0;
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || optin <= 0) {
      delete h;
      return fail(nullptr, CGPE_ERR_CUDA, "cannot read the shared-memory limit of device %d", dev);
    }
    h->smem_optin = (size_t)optin;
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    h->n_sm = n_sm > 0 ? n_sm : 148;
    std::memset(&h->duo, 0, sizeof(h->duo));
  }
  DevParams &p = h->p;
  std::memset(&p, 0, sizeof(p));
  p.R = cfg->n_replicas; p.N = cfg->n_beads; p.NC = cfg->n_chains * cfg->n_beads;
  p.P = p.NC + cfg->n_ion_slots; p.T = cfg->n_types; p.n_rx = cfg->n_reactions; p.n_ions = explicit_ions(cfg) ? cfg->n_ion_slots : 0;
  p.bond_kind = cfg->bond_kind; p.use_angle = cfg->use_angle; p.use_dih = cfg->use_dihedral;
  p.dih_mult = cfg->dihedral_mult; p.use_wca = cfg->use_wca; p.use_dh = cfg->use_dh;
  p.replica_offset = cfg->replica_offset;
  p.bond_k = (float)cfg->bond_k; p.bond_r0 = (float)cfg->bond_r0;
  p.fene_drmax2 = (float)(cfg->fene_drmax * cfg->fene_drmax);
  p.angle_k = (float)cfg->angle_k; p.angle_phi0 = (float)cfg->angle_phi0;
  p.dih_k = (float)cfg->dihedral_k;
  p.dih_cos_phase = (float)std::cos(cfg->dihedral_phase); p.dih_sin_phase = (float)std::sin(cfg->dihedral_phase);
  {
    const double smin = cfg->dihedral_sin_min == 0.0 ? 0.02 : (cfg->dihedral_sin_min < 0.0 ? 0.0 : cfg->dihedral_sin_min);
    p.dih_sin_min2 = (float)(smin * smin);
  }
  p.dh_pref = (float)cfg->dh_prefactor;
  p.box_l = (float)cfg->box_l; p.inv_box_l = (float)(1.0 / cfg->box_l);
  p.dt = (float)cfg->time_step; p.half_dt = (float)(0.5 * cfg->time_step); p.dt_d = cfg->time_step;
  // default Verlet skin (both engines: list builds are cheap next to the steps they serve)
  const double skin = cfg->skin > 0.0 ? cfg->skin : 0.8;
  p.skin = (float)skin; p.half_skin2 = (float)(0.25 * skin * skin);
  p.force_cap = 0.f;
  p.fixed_mask = (uint32_t)cfg->fixed_type_mask;
  p.thermostat_seed = 0u;
  if (const char *dp = std::getenv("CGPE_DENSE_POLICY")) p.dense_policy = std::atoi(dp);   // diagnostics only
  if (const char *en = std::getenv("CGPE_ENGINE")) h->engine_pref = std::strcmp(en, "solo") == 0 ? 0 : std::strcmp(en, "duo") == 0 ? 1 : -1;
  for (int m = 0; m < p.n_rx; ++m) {
    const cgpe_reaction &r = cfg->reactions[m];
    DevReaction &d = p.rx[m];
    d.type_prot = r.type_prot; d.type_deprot = r.type_deprot; d.type_ion = r.type_ion;
    d.seed = (uint32_t)r.seed;
    d.q_prot = (float)r.q_prot; d.q_deprot = (float)r.q_deprot; d.q_ion = (float)r.q_ion;
    d.inv_kT = (float)(1.0 / r.kT); d.pK = r.pK;
    d.excl2 = (float)(r.exclusion_range * r.exclusion_range); d.sqrt_kT = (float)std::sqrt(r.kT);
    if (p.n_ions > 0) p.grid_w_min = std::fmax(p.grid_w_min, (float)(r.exclusion_range * 1.001));
  }
  // chargeable particles are fixed at set_state; size the lists for "every third bead" + ions
  p.n_chg = 0;
  p.cap_w = cfg->cap_wca;   // 0: chosen at set_state (choose_capacities)
  if (p.cap_w > 255 || p.cap_w < 0) { delete h; return fail(nullptr, CGPE_ERR_INVALID, "list capacities must be <= 255"); }

  auto bail = [&](cudaError_t e, const char *what) {
    fail(nullptr, CGPE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    cgpe_destroy(h);
    return CGPE_ERR_CUDA;
  };
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
  h->stream = h->own_stream;
  if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
  const size_t RP = (size_t)p.R * p.P, nrx = p.n_rx > 0 ? p.n_rx : 1;
  DevState &g = h->g;
#define ALLOC(ptr, n) if ((e = dev_alloc(h, &(ptr), (n))) != cudaSuccess) return bail(e, "cudaMalloc " #ptr)
  ALLOC(g.pos, RP); ALLOC(g.vel, RP); ALLOC(g.frc, RP); ALLOC(g.type, RP);
  ALLOC(g.wca_tab, (size_t)(p.T + 1) * (p.T + 1));
  ALLOC(g.chg, RP);
  ALLOC(g.lst_prot, (size_t)p.R * nrx * p.P); ALLOC(g.lst_deprot, (size_t)p.R * nrx * p.P);
  ALLOC(g.cnt, (size_t)p.R * nrx * 2);
  ALLOC(g.ion_active, (size_t)p.R * (p.n_ions > 0 ? p.n_ions : 1)); ALLOC(g.ion_free, (size_t)p.R * (p.n_ions > 0 ? p.n_ions : 1));
  ALLOC(g.ion_cnt, (size_t)p.R * 2);
  ALLOC(g.kappa, p.R); ALLOC(g.rcut, p.R); ALLOC(g.kT, p.R); ALLOC(g.gamma, p.R); ALLOC(g.pH, p.R);
  ALLOC(g.time, p.R); ALLOC(g.md_counter, p.R); ALLOC(g.md_steps, p.R); ALLOC(g.samples_done, p.R);
  ALLOC(g.rebuilds, p.R); ALLOC(g.compact_builds, p.R); ALLOC(g.cta_ns, (size_t)p.R * 4);
  ALLOC(g.n_frames, p.R); ALLOC(g.frame_phase, p.R); ALLOC(g.trials, (size_t)p.R * nrx); ALLOC(g.accepted, (size_t)p.R * nrx);
  ALLOC(g.f_valid, p.R); ALLOC(g.err, p.R);
  ALLOC(g.plan, (size_t)p.R * 8); ALLOC(h->d_order, p.R); ALLOC(h->d_choice, 1);
  g.order = nullptr;
  g.engine_choice = nullptr;
  ALLOC(h->d_tmp_d, (size_t)p.R * 8); ALLOC(h->d_tmp_i, (size_t)p.R * 3); ALLOC(h->d_tmp_f4, RP);
#undef ALLOC
  std::vector<double> d(p.R);
  std::vector<float> f(p.R);
  auto up_f = [&](float *dst, double v) { for (auto &x : f) x = (float)v; return cudaMemcpy(dst, f.data(), sizeof(float) * p.R, cudaMemcpyHostToDevice); };
  if ((e = up_f(g.kappa, cfg->kappa)) != cudaSuccess) return bail(e, "upload kappa");
  if ((e = up_f(g.rcut, cfg->use_dh ? cfg->r_cut : 1.0)) != cudaSuccess) return bail(e, "upload rcut");
  if ((e = up_f(g.kT, cfg->kT)) != cudaSuccess) return bail(e, "upload kT");
  if ((e = up_f(g.gamma, cfg->gamma)) != cudaSuccess) return bail(e, "upload gamma");
  for (auto &x : d) x = cfg->pH;
  if ((e = cudaMemcpy(g.pH, d.data(), sizeof(double) * p.R, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload pH");
  if ((rc = upload_wca(h)) != CGPE_OK) { std::snprintf(g_err, 512, "%s", h->err); cgpe_destroy(h); return rc; }
  *out = h;
  return CGPE_OK;
}

void cgpe_destroy(cgpe_handle *h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  wide_graph_release(&h->wgraph);
  for (void *ptr : h->allocs) cudaFree(ptr);
  for (void *ptr : h->wide_allocs) cudaFree(ptr);
  if (h->dhc_alloc) cudaFree(h->dhc_alloc);
  for (void *ptr : h->frame_alloc) if (ptr) cudaFree(ptr);
  if (h->mc_e_alloc) cudaFree(h->mc_e_alloc);
  if (h->d_obs) cudaFree(h->d_obs);
  if (h->d_qdist) cudaFree(h->d_qdist);
  if (h->d_trace) cudaFree(h->d_trace);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

int cgpe_device_info(const cgpe_handle *hc, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor, char *name,
                     int32_t name_len) {
  cgpe_handle *h = const_cast<cgpe_handle *>(hc);
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, h->device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (name && name_len > 0) std::snprintf(name, (size_t)name_len, "%s", prop.name);
  return CGPE_OK;
}

int cgpe_set_stream(cgpe_handle *h, void *cuda_stream) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  DeviceGuard guard(h->device);
  CU(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return CGPE_OK;
}

int cgpe_synchronize(cgpe_handle *h) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  DeviceGuard guard(h->device);
  CU(cudaStreamSynchronize(h->stream));
  return h->have_state ? check_device_errors(h) : CGPE_OK;
}

int cgpe_set_state(cgpe_handle *h, const float *xyzq, const float *vel, const uint8_t *type) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!xyzq || !type) return fail(h, CGPE_ERR_INVALID, "xyzq and type are required");
  DeviceGuard guard(h->device);
  DevParams &p = h->p;
  const size_t RP = (size_t)p.R * p.P;
  for (size_t k = 0; k < RP; ++k) {
    if (type[k] > p.T) return fail(h, CGPE_ERR_INVALID, "type index %d out of range at particle %zu", (int)type[k], k);
    if (!std::isfinite(xyzq[4 * k]) || !std::isfinite(xyzq[4 * k + 1]) || !std::isfinite(xyzq[4 * k + 2]))
      return fail(h, CGPE_ERR_INVALID, "non-finite coordinate at particle %zu", k);
  }
  // the trial moves swap a site between its two default charges (M.py:239-243): a titratable
  // particle must carry the default charge of its current type
  for (int m = 0; m < p.n_rx; ++m) {
    const cgpe_reaction &rx = h->cfg.reactions[m];
    for (size_t k = 0; k < RP; ++k) {
      const float want = type[k] == rx.type_prot ? (float)rx.q_prot : type[k] == rx.type_deprot ? (float)rx.q_deprot : xyzq[4 * k + 3];
      if (xyzq[4 * k + 3] != want)
        return fail(h, CGPE_ERR_INVALID, "particle %zu of titratable type %d carries charge %g, default_charges says %g",
                    k % p.P, (int)type[k], xyzq[4 * k + 3], want);
    }
  }
  // particles that can carry charge: charged now, or of a type a reaction can charge
  unsigned reactive = 0;
  for (int m = 0; m < p.n_rx; ++m) reactive |= (1u << h->cfg.reactions[m].type_prot) | (1u << h->cfg.reactions[m].type_deprot);
  std::vector<int> chg(RP, 0);
  int n_chg = -1;
  for (int r = 0; r < p.R; ++r) {
    int n = 0;
    for (int i = 0; i < p.P; ++i) {
      const size_t k = (size_t)r * p.P + i;
      if (xyzq[4 * k + 3] != 0.f || ((reactive >> type[k]) & 1u) || i >= p.NC) chg[(size_t)r * p.P + n++] = i;
    }
    if (n_chg >= 0 && n != n_chg) return fail(h, CGPE_ERR_INVALID, "replicas differ in the number of chargeable particles");
    n_chg = n;
  }
  std::vector<int> chg_packed((size_t)p.R * (n_chg > 0 ? n_chg : 1));
  for (int r = 0; r < p.R; ++r)
    for (int i = 0; i < n_chg; ++i) chg_packed[(size_t)r * n_chg + i] = chg[(size_t)r * p.P + i];
  p.n_chg = n_chg;
  h->chg_host = chg_packed;
  {
    int rc_cap = choose_capacities(h);
    if (rc_cap) return rc_cap;
  }
  if (h->launch.items != 0 && h->launch.smem > h->smem_optin) h->launch.items = 0;   // -> global-memory engine
  if (h->launch.items != 0 && h->solo_cap_low && p.n_ions > 0 && h->engine_pref != 0) h->launch.items = 0;   // explicit ions, too few WCA row entries: ditto
  h->g.dhc = nullptr; p.dhc_cap = 0; p.dhc_stride = 0;
  h->duo.ok = 0;
  if (h->launch.items == 0) {
    if (p.frame_every > 0) return fail(h, CGPE_ERR_UNSUPPORTED, "the frame recorder is served by the shared-memory engine (<= 2048 particles per replica)");
    if (p.P > p.NC && p.n_ions == 0)
      return fail(h, CGPE_ERR_UNSUPPORTED, "spectator particles (slots without an ion reaction) are served by the shared-memory engine only (<= 2048 particles per replica)");
    int rc_w = ensure_wide(h, chg_packed);
    if (rc_w) return rc_w;
  } else {
    free_wide(h);
    if (p.use_dh && n_chg > 0) {   // compact partner lists of the shared-memory engine (chain_engine.cu::build_dh_compact)
      const int G = p.dh_group > 0 ? p.dh_group : 1;
      p.dhc_stride = (G * n_chg + 31) / 32 * 32;
      p.dhc_cap = (((32 + G - 1) / G) * ((n_chg + 31) / 32) + 7) / 8 * 8;   // whole 16-byte groups of eight entries
      size_t need = sizeof(uint16_t) * (size_t)p.R * p.dhc_cap * p.dhc_stride;
      {
        int rc_duo = plan_duo(h);
        if (rc_duo) return rc_duo;
        if (h->duo.ok) need = std::max(need, sizeof(uint16_t) * (size_t)p.R * 2 * h->duo.dhc_cap * h->duo.dhc_stride);
      }
      if (need > h->dhc_bytes) {   // kept across set_state calls: cudaFree/cudaMalloc synchronise the device
        if (h->dhc_alloc) { cudaFree(h->dhc_alloc); h->dhc_alloc = nullptr; h->dhc_bytes = 0; }
        CU(cudaMalloc(&h->dhc_alloc, need));
        h->dhc_bytes = need;
      }
      h->g.dhc = static_cast<uint16_t *>(h->dhc_alloc);
    } else {
      int rc_duo = plan_duo(h);
      if (rc_duo) return rc_duo;
    }
  }
  // pair-kernel matrix of the trial moves (chain_engine.cu mc_prepare / mc_block): implicit-ion replicas of the shared-memory
  // engines, as long as it stays a small part of HBM
  h->g.mc_e = nullptr; h->g.mc_e_stride = 0;
  if (h->launch.items != 0 && p.n_ions == 0 && p.use_dh && p.n_rx > 0 && n_chg > 0 && !std::getenv("CGPE_NO_MC_MATRIX")) {
    const int stride = (n_chg + 127) / 128 * 128;
    const size_t need = sizeof(float) * (size_t)p.R * n_chg * stride;
    if (need <= ((size_t)8 << 30)) {
      if (need > h->mc_e_bytes) {
        if (h->mc_e_alloc) { cudaFree(h->mc_e_alloc); h->mc_e_alloc = nullptr; h->mc_e_bytes = 0; }
        CU(cudaMalloc(&h->mc_e_alloc, need));
        h->mc_e_bytes = need;
      }
      h->g.mc_e = static_cast<float *>(h->mc_e_alloc);
      h->g.mc_e_stride = stride;
    }
  }
  {   // list radii: see upload_wca
    bool fixed_in_chain = false;
    for (int r = 0; r < p.R && !fixed_in_chain; ++r)
      for (int i = 0; i < p.NC; ++i)
        if ((p.fixed_mask >> type[(size_t)r * p.P + i]) & 1u) { fixed_in_chain = true; break; }
    const bool two = p.fixed_mask != 0u && !fixed_in_chain && h->launch.items != 0;
    if (two != h->two_radii) {
      h->two_radii = two;
      int rc_r = upload_wca(h);
      if (rc_r) return rc_r;
    }
  }
  // bookkeeping lists in particle-id order
  const size_t nrx = p.n_rx > 0 ? p.n_rx : 1;
  std::vector<int> lp((size_t)p.R * nrx * p.P, 0), ld((size_t)p.R * nrx * p.P, 0), cnt((size_t)p.R * nrx * 2, 0);
  for (int r = 0; r < p.R; ++r)
    for (int m = 0; m < p.n_rx; ++m) {
      int np = 0, nd = 0;
      const size_t o = ((size_t)r * nrx + m) * p.P;
      for (int i = 0; i < p.P; ++i) {
        const int t = type[(size_t)r * p.P + i];
        if (t == h->cfg.reactions[m].type_prot) lp[o + np++] = i;
        else if (t == h->cfg.reactions[m].type_deprot) ld[o + nd++] = i;
      }
      cnt[((size_t)r * nrx + m) * 2] = np; cnt[((size_t)r * nrx + m) * 2 + 1] = nd;
    }
  std::vector<float> zero;
  if (!vel) zero.assign(RP * 4, 0.f);
  std::vector<float> xq_fix, vel_fix;
  if (p.n_ions > 0) {
    // hidden slots (non-interacting type) carry neither charge nor velocity; pools in id order
    xq_fix.assign(xyzq, xyzq + RP * 4);
    vel_fix.assign(vel ? vel : zero.data(), (vel ? vel : zero.data()) + RP * 4);
    std::vector<int> act((size_t)p.R * p.n_ions, 0), fre((size_t)p.R * p.n_ions, 0), icnt((size_t)p.R * 2, 0);
    const int tb = p.n_rx > 0 ? h->cfg.reactions[0].type_ion : -1;
    for (int r = 0; r < p.R; ++r)
      for (int i = p.NC; i < p.P; ++i) {
        const size_t k = (size_t)r * p.P + i;
        if (type[k] >= p.T) {
          xq_fix[4 * k + 3] = 0.f; vel_fix[4 * k] = vel_fix[4 * k + 1] = vel_fix[4 * k + 2] = 0.f;
          fre[(size_t)r * p.n_ions + icnt[r * 2 + 1]++] = i;
        } else if (type[k] == tb) {
          act[(size_t)r * p.n_ions + icnt[r * 2]++] = i;
        }
      }
    xyzq = xq_fix.data(); vel = vel_fix.data();
    CU(cudaMemcpyAsync(h->g.ion_active, act.data(), sizeof(int) * act.size(), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->g.ion_free, fre.data(), sizeof(int) * fre.size(), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->g.ion_cnt, icnt.data(), sizeof(int) * icnt.size(), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  std::vector<float> vel_still;
  if (p.fixed_mask != 0u && vel) {   // fixed particles (fix=[True]*3) hold no velocity
    vel_still.assign(vel, vel + RP * 4);
    for (size_t k = 0; k < RP; ++k)
      if ((p.fixed_mask >> type[k]) & 1u) vel_still[4 * k] = vel_still[4 * k + 1] = vel_still[4 * k + 2] = 0.f;
    vel = vel_still.data();
  }
  CU(cudaMemcpyAsync(h->g.pos, xyzq, sizeof(float4) * RP, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->g.vel, vel ? vel : zero.data(), sizeof(float4) * RP, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->g.type, type, RP, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->g.chg, chg_packed.data(), sizeof(int) * chg_packed.size(), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->g.lst_prot, lp.data(), sizeof(int) * lp.size(), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->g.lst_deprot, ld.data(), sizeof(int) * ld.size(), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->g.cnt, cnt.data(), sizeof(int) * cnt.size(), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemsetAsync(h->g.f_valid, 0, sizeof(int) * p.R, h->stream));
  CU(cudaMemsetAsync(h->g.err, 0, sizeof(int) * p.R, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->have_state = true;
  return CGPE_OK;
}

int cgpe_get_state(cgpe_handle *h, float *xyzq, float *vel, uint8_t *type) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  DeviceGuard guard(h->device);
  const size_t RP = (size_t)h->p.R * h->p.P;
  if (xyzq) CU(cudaMemcpyAsync(xyzq, h->g.pos, sizeof(float4) * RP, cudaMemcpyDeviceToHost, h->stream));
  if (vel) CU(cudaMemcpyAsync(vel, h->g.vel, sizeof(float4) * RP, cudaMemcpyDeviceToHost, h->stream));
  if (type) CU(cudaMemcpyAsync(type, h->g.type, RP, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return check_device_errors(h);
}

int cgpe_set_replica_params(cgpe_handle *h, const double *pH, const double *kappa, const double *r_cut,
                            const double *kT, const double *gamma) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  DeviceGuard guard(h->device);
  const int R = h->p.R;
  std::vector<float> f(R);
  auto up = [&](float *dst, const double *src) {
    for (int r = 0; r < R; ++r) f[r] = (float)src[r];
    cudaError_t e = cudaMemcpyAsync(dst, f.data(), sizeof(float) * R, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    return e;
  };
  if (r_cut) {
    double rc_max = 0.0;
    for (int r = 0; r < R; ++r) {
      if (!(r_cut[r] > 0.0) || r_cut[r] + h->p.skin > 0.5 * h->cfg.box_l)
        return fail(h, CGPE_ERR_INVALID, "r_cut + skin must lie in (0, box_l/2] (replica %d: %g)", r, r_cut[r]);
      rc_max = std::fmax(rc_max, r_cut[r]);
    }
    const bool grew = rc_max > h->cfg.r_cut;
    h->cfg.r_cut = rc_max;
    if (h->have_state) {
      int rc_cap = choose_capacities(h);
      if (rc_cap) return rc_cap;
      if (h->wide && grew && h->cfg.cap_dh <= 0) {   // longer rows are needed
        rc_cap = ensure_wide(h, h->chg_host);
        if (rc_cap) return rc_cap;
      }
      if (h->wide) update_wide_policy(h);
    }
    CU(up(h->g.rcut, r_cut));
  }
  if (kappa) CU(up(h->g.kappa, kappa));
  if (kT) {
    for (int r = 0; r < R; ++r) if (kT[r] < 0.0) return fail(h, CGPE_ERR_INVALID, "kT must be >= 0");
    CU(up(h->g.kT, kT));
  }
  if (gamma) {
    for (int r = 0; r < R; ++r) if (gamma[r] < 0.0) return fail(h, CGPE_ERR_INVALID, "gamma must be >= 0");
    CU(up(h->g.gamma, gamma));
  }
  if (pH) {
    CU(cudaMemcpyAsync(h->g.pH, pH, sizeof(double) * R, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  if (kappa || r_cut || kT || gamma) CU(cudaMemsetAsync(h->g.f_valid, 0, sizeof(int) * R, h->stream));
  return CGPE_OK;
}

int cgpe_set_wca(cgpe_handle *h, const double *eps, const double *sig) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!eps || !sig) return fail(h, CGPE_ERR_INVALID, "eps and sig are required");
  DeviceGuard guard(h->device);
  std::memcpy(h->cfg.wca_eps, eps, sizeof(h->cfg.wca_eps));
  std::memcpy(h->cfg.wca_sig, sig, sizeof(h->cfg.wca_sig));
  int rc = upload_wca(h);
  if (rc) return rc;
  CU(cudaMemsetAsync(h->g.f_valid, 0, sizeof(int) * h->p.R, h->stream));
  return CGPE_OK;
}

int cgpe_set_thermostat_seed(cgpe_handle *h, int32_t seed) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  DeviceGuard guard(h->device);
  if (seed >= 0) {
    h->p.thermostat_seed = (uint32_t)seed;
    CU(cudaMemsetAsync(h->g.md_counter, 0, sizeof(long long) * h->p.R, h->stream));
  }
  CU(cudaMemsetAsync(h->g.f_valid, 0, sizeof(int) * h->p.R, h->stream));
  return CGPE_OK;
}

int cgpe_set_force_cap(cgpe_handle *h, double cap) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (cap < 0.0) return fail(h, CGPE_ERR_INVALID, "force cap must be >= 0");
  DeviceGuard guard(h->device);
  h->p.force_cap = (float)cap;
  CU(cudaMemsetAsync(h->g.f_valid, 0, sizeof(int) * h->p.R, h->stream));
  return CGPE_OK;
}

int cgpe_set_time(cgpe_handle *h, double t) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  DeviceGuard guard(h->device);
  std::vector<double> v(h->p.R, t);
  CU(cudaMemcpyAsync(h->g.time, v.data(), sizeof(double) * h->p.R, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return CGPE_OK;
}

int cgpe_get_time(cgpe_handle *h, double *t) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!t) return fail(h, CGPE_ERR_INVALID, "t is NULL");
  DeviceGuard guard(h->device);
  CU(cudaMemcpyAsync(t, h->g.time, sizeof(double) * h->p.R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return CGPE_OK;
}

int cgpe_md_run(cgpe_handle *h, int32_t n_steps) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (n_steps < 0) return fail(h, CGPE_ERR_INVALID, "n_steps must be >= 0");
  int rc = require_engine(h);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  begin_timing(h);
  int launches = 1;
  if (h->wide) { launches = 0; CU(wide_md_run(h->p, h->g, h->w, n_steps, h->stream, &launches, &h->wgraph)); }
  else if (use_duo(h)) { launches = duo_launches(h); CU(duo_md_run(h->p, h->g, h->duo, n_steps, h->stream)); }
  else CU(launch_md_run(h->p, h->g, h->launch, n_steps, h->stream));
  end_timing(h, launches);
  return CGPE_OK;
}

int cgpe_steepest_descent(cgpe_handle *h, int32_t n_steps, double f_max, double gamma, double max_displacement) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (n_steps < 0 || f_max < 0.0 || gamma < 0.0 || max_displacement < 0.0)
    return fail(h, CGPE_ERR_INVALID, "steepest descent parameters must be >= 0");
  int rc = require_engine(h);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  begin_timing(h);
  int launches = 1;
  if (h->wide) { launches = 0; CU(wide_steepest_descent(h->p, h->g, h->w, n_steps, (float)f_max, (float)gamma, (float)max_displacement, h->stream, &launches)); }
  else CU(launch_steepest_descent(h->p, h->g, h->launch, n_steps, (float)f_max, (float)gamma, (float)max_displacement, h->stream));
  end_timing(h, launches);
  return CGPE_OK;
}

int cgpe_mc_run(cgpe_handle *h, int32_t reaction_id, int32_t n_trials, cgpe_trial_record *trace) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (reaction_id < 0 || reaction_id >= h->p.n_rx) return fail(h, CGPE_ERR_INVALID, "reaction_id out of range");
  if (n_trials < 0) return fail(h, CGPE_ERR_INVALID, "n_trials must be >= 0");
  int rc = require_engine(h);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  cgpe_trial_record *d_trace = nullptr;
  const size_t n_rec = (size_t)h->p.R * n_trials;
  if (trace && n_rec > 0) {
    if (n_rec > h->trace_cap) {
      if (h->d_trace) cudaFree(h->d_trace);
      h->d_trace = nullptr; h->trace_cap = 0;
      CU(cudaMalloc((void **)&h->d_trace, sizeof(cgpe_trial_record) * n_rec));
      h->trace_cap = n_rec;
    }
    d_trace = h->d_trace;
  }
  begin_timing(h);
  int launches = 1;
  if (h->wide) { launches = 0; CU(wide_mc_run(h->p, h->g, h->w, reaction_id, n_trials, d_trace, h->stream, &launches)); }
  else if (use_duo(h)) { launches = duo_launches(h); CU(duo_mc_run(h->p, h->g, h->duo, reaction_id, n_trials, d_trace, h->stream)); }
  else CU(launch_mc_run(h->p, h->g, h->launch, reaction_id, n_trials, d_trace, h->stream));
  end_timing(h, launches);
  if (d_trace) {
    CU(cudaMemcpyAsync(trace, d_trace, sizeof(cgpe_trial_record) * n_rec, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return check_device_errors(h);
  }
  return CGPE_OK;
}

int cgpe_sample_loop_async(cgpe_handle *h, const cgpe_sample_params *sp) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!sp || sp->n_samples < 0 || sp->n_mc_blocks < 0 || sp->n_mc_blocks > CGPE_MAX_MC_BLOCKS ||
      sp->md_steps_per_burst < 0)
    return fail(h, CGPE_ERR_INVALID, "bad sample params");
  for (int b = 0; b < sp->n_mc_blocks; ++b)
    if (sp->mc_reaction[b] < 0 || sp->mc_reaction[b] >= h->p.n_rx || sp->mc_trials[b] < 0)
      return fail(h, CGPE_ERR_INVALID, "bad MC block %d", b);
  int rc = require_engine(h);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  const size_t n_obs = (size_t)h->p.R * (sp->n_samples > 0 ? sp->n_samples : 1) * CGPE_N_OBS;
  const int rows = sp->q_snapshot_every > 0 ? sp->q_snapshot_rows : 0;
  const size_t n_q = (size_t)h->p.R * (rows > 0 ? rows : 0) * h->p.N;
  if (n_obs > h->obs_cap) {
    if (h->d_obs) cudaFree(h->d_obs);
    h->d_obs = nullptr; h->obs_cap = 0;
    CU(cudaMalloc((void **)&h->d_obs, sizeof(double) * n_obs));
    h->obs_cap = n_obs;
  }
  if (n_q > h->qdist_cap) {
    if (h->d_qdist) cudaFree(h->d_qdist);
    h->d_qdist = nullptr; h->qdist_cap = 0;
    CU(cudaMalloc((void **)&h->d_qdist, sizeof(float) * n_q));
    h->qdist_cap = n_q;
  }
  if (n_q > 0) CU(cudaMemsetAsync(h->d_qdist, 0, sizeof(float) * n_q, h->stream));
  h->last_ns = sp->n_samples; h->last_qrows = rows;
  begin_timing(h);
  int launches = 1;
  h->choice_live = false;
  if (h->wide) {
    launches = 0;
    CU(wide_sample_loop(h->p, h->g, h->w, *sp, bernoulli_threshold(sp->p_burst1), bernoulli_threshold(sp->p_burst2),
                        h->d_obs, n_q > 0 ? h->d_qdist : nullptr, h->d_tmp_d, h->d_tmp_i, h->stream, &launches, &h->wgraph));
  } else if (use_duo(h)) {
    launches = duo_launches(h);   // the placement plan (k_duo_count_pairs, k_duo_plan) + the loop itself
    DevState g = h->g;
    // left to itself (no cgpe_set_engine, no CGPE_ENGINE) the library issues the launch on both shared-memory engines and
    // the plan decides on the device which one runs it (the other kernel returns at once): no host round trip
    const bool both = h->engine_pref < 0 && !h->solo_cap_low && h->p.R <= h->n_sm && h->launch.items != 0;
    g.engine_choice = both ? h->d_choice : nullptr;
    CU(duo_sample_loop(h->p, g, h->duo, *sp, bernoulli_threshold(sp->p_burst1), bernoulli_threshold(sp->p_burst2),
                       h->d_obs, n_q > 0 ? h->d_qdist : nullptr, h->stream));
    if (both) {
      CU(launch_sample_loop(h->p, g, h->launch, *sp, bernoulli_threshold(sp->p_burst1), bernoulli_threshold(sp->p_burst2),
                            h->d_obs, n_q > 0 ? h->d_qdist : nullptr, h->stream));
      launches += 1;
    }
    h->choice_live = both;
  } else {
    DevState g = h->g;
    g.order = nullptr;
    // more replicas than SMs (and few enough for the planner's shared memory): the launch runs in waves, most expensive first
    if (h->p.R > h->n_sm && h->p.R <= 4096 && h->p.n_ions == 0 && !std::getenv("CGPE_NO_WAVE_PLAN")) {
      CU(duo_plan_waves(h->p, h->g, h->d_order, *sp, bernoulli_threshold(sp->p_burst1), bernoulli_threshold(sp->p_burst2), h->stream));
      g.order = h->d_order;
      launches += h->p.use_dh ? 2 : 1;
    }
    CU(launch_sample_loop(h->p, g, h->launch, *sp, bernoulli_threshold(sp->p_burst1), bernoulli_threshold(sp->p_burst2),
                          h->d_obs, n_q > 0 ? h->d_qdist : nullptr, h->stream));
  }
  end_timing(h, launches);
  return CGPE_OK;
}

int cgpe_fetch_samples(cgpe_handle *h, double *obs, float *qdist) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!h->d_obs) return fail(h, CGPE_ERR_STATE, "no sample loop has run");
  DeviceGuard guard(h->device);
  if (obs && h->last_ns > 0)
    CU(cudaMemcpyAsync(obs, h->d_obs, sizeof(double) * (size_t)h->p.R * h->last_ns * CGPE_N_OBS, cudaMemcpyDeviceToHost, h->stream));
  if (qdist && h->last_qrows > 0)
    CU(cudaMemcpyAsync(qdist, h->d_qdist, sizeof(float) * (size_t)h->p.R * h->last_qrows * h->p.N, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return check_device_errors(h);
}

int cgpe_sample_loop(cgpe_handle *h, const cgpe_sample_params *sp, double *obs, float *qdist) {
  int rc = cgpe_sample_loop_async(h, sp);
  if (rc) return rc;
  return cgpe_fetch_samples(h, obs, qdist);
}

int cgpe_energy(cgpe_handle *h, double *out) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!out) return fail(h, CGPE_ERR_INVALID, "out is NULL");
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  DeviceGuard guard(h->device);
  CU(launch_energy(h->p, h->g, h->d_tmp_d, h->stream));
  CU(cudaMemcpyAsync(out, h->d_tmp_d, sizeof(double) * h->p.R * CGPE_N_ENERGY, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return check_device_errors(h);
}

int cgpe_forces(cgpe_handle *h, float *out) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!out) return fail(h, CGPE_ERR_INVALID, "out is NULL");
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  DeviceGuard guard(h->device);
  CU(launch_forces(h->p, h->g, h->d_tmp_f4, h->stream));
  CU(cudaMemcpyAsync(out, h->d_tmp_f4, sizeof(float4) * (size_t)h->p.R * h->p.P, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return check_device_errors(h);
}

int cgpe_md_forces(cgpe_handle *h, float *out) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!out) return fail(h, CGPE_ERR_INVALID, "out is NULL");
  int rc = require_engine(h);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  if (h->wide) CU(wide_md_forces(h->p, h->g, h->w, h->d_tmp_f4, h->stream));
  else if (use_duo(h)) CU(duo_md_forces(h->p, h->g, h->duo, h->d_tmp_f4, h->stream));
  else CU(launch_md_forces(h->p, h->g, h->launch, h->d_tmp_f4, h->stream));
  CU(cudaMemcpyAsync(out, h->d_tmp_f4, sizeof(float4) * (size_t)h->p.R * h->p.P, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return check_device_errors(h);
}

int cgpe_observables(cgpe_handle *h, double *rg, double *ree, int32_t *counts) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  DeviceGuard guard(h->device);
  const int R = h->p.R;
  CU(launch_observables(h->p, h->g, h->d_tmp_d, h->d_tmp_d + R, h->d_tmp_i, h->stream));
  if (rg) CU(cudaMemcpyAsync(rg, h->d_tmp_d, sizeof(double) * R, cudaMemcpyDeviceToHost, h->stream));
  if (ree) CU(cudaMemcpyAsync(ree, h->d_tmp_d + R, sizeof(double) * R, cudaMemcpyDeviceToHost, h->stream));
  if (counts) CU(cudaMemcpyAsync(counts, h->d_tmp_i, sizeof(int) * R * 3, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return check_device_errors(h);
}

int cgpe_min_dist(cgpe_handle *h, double *out) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!out) return fail(h, CGPE_ERR_INVALID, "out is NULL");
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  DeviceGuard guard(h->device);
  CU(launch_min_dist(h->p, h->g, h->d_tmp_d, h->stream));
  CU(cudaMemcpyAsync(out, h->d_tmp_d, sizeof(double) * h->p.R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return check_device_errors(h);
}

int cgpe_get_counters(cgpe_handle *h, int64_t *md_steps, int64_t *trials, int64_t *accepted, int64_t *list_rebuilds) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  DeviceGuard guard(h->device);
  const int R = h->p.R, nrx = h->p.n_rx;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
  if (md_steps) CU(cudaMemcpyAsync(md_steps, h->g.md_steps, sizeof(int64_t) * R, cudaMemcpyDeviceToHost, h->stream));
  if (trials && nrx > 0) CU(cudaMemcpyAsync(trials, h->g.trials, sizeof(int64_t) * R * nrx, cudaMemcpyDeviceToHost, h->stream));
  if (accepted && nrx > 0) CU(cudaMemcpyAsync(accepted, h->g.accepted, sizeof(int64_t) * R * nrx, cudaMemcpyDeviceToHost, h->stream));
  if (list_rebuilds) CU(cudaMemcpyAsync(list_rebuilds, h->g.rebuilds, sizeof(int64_t) * R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return CGPE_OK;
}

int cgpe_list_stats(cgpe_handle *h, double *out) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!out) return fail(h, CGPE_ERR_INVALID, "out is NULL");
  int rc = require_engine(h);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  if (h->wide) CU(wide_list_stats(h->p, h->g, h->w, h->d_tmp_d, h->stream));
  else CU(launch_list_stats(h->p, h->g, h->launch, h->d_tmp_d, h->stream));
  CU(cudaMemcpyAsync(out, h->d_tmp_d, sizeof(double) * h->p.R * 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return check_device_errors(h);
}

int cgpe_set_engine(cgpe_handle *h, int32_t engine) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (engine < -1 || engine > 1) return fail(h, CGPE_ERR_INVALID, "engine must be -1 (choose), 0 (one CTA per replica) or 1 (cluster of two CTAs)");
  if (engine == 1 && h->have_state && (!h->duo.ok || h->wide))
    return fail(h, CGPE_ERR_UNSUPPORTED, "the cluster engine serves implicit-ion replicas of 8..2048 chain beads without spectator particles");
  h->engine_pref = engine;
  return CGPE_OK;
}

int cgpe_engine_info(cgpe_handle *h, int32_t *engine, int64_t *compact_builds, int64_t *cta_ns) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  DeviceGuard guard(h->device);
  if (engine) {
    *engine = h->wide ? 2 : use_duo(h) ? 1 : 0;
    if (*engine == 1 && h->choice_live) {   // the last fused sampling launch chose on the device
      int choice = 1;
      CU(cudaMemcpyAsync(&choice, h->d_choice, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      *engine = choice;
    }
  }
  if (compact_builds) CU(cudaMemcpyAsync(compact_builds, h->g.compact_builds, sizeof(int64_t) * h->p.R, cudaMemcpyDeviceToHost, h->stream));
  if (cta_ns) CU(cudaMemcpyAsync(cta_ns, h->g.cta_ns, sizeof(int64_t) * h->p.R * 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return CGPE_OK;
}

int cgpe_frames_end(cgpe_handle *h) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  DeviceGuard guard(h->device);
  CU(cudaStreamSynchronize(h->stream));
  h->p.frame_every = 0; h->p.frame_flags = 0; h->p.frame_cap = 0;
  for (void *&ptr : h->frame_alloc) { if (ptr) cudaFree(ptr); ptr = nullptr; }
  h->g.frames = nullptr; h->g.frame_bonds = nullptr; h->g.frame_dih = nullptr;
  return CGPE_OK;
}

int cgpe_frames_begin(cgpe_handle *h, int32_t every, int32_t flags, int32_t capacity) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (every < 1 || capacity < 1 || (flags & ~(CGPE_FRAME_BONDS | CGPE_FRAME_DIHEDRALS)))
    return fail(h, CGPE_ERR_INVALID, "frames: every and capacity must be positive, flags a combination of CGPE_FRAME_BONDS, CGPE_FRAME_DIHEDRALS");
  if (h->have_state && (h->wide || h->launch.items == 0))
    return fail(h, CGPE_ERR_UNSUPPORTED, "the frame recorder is served by the shared-memory engine (<= 2048 particles per replica)");
  const bool keep_phase = h->p.frame_every == every;   // re-arming with the same interval (a larger capacity): the step phase carries over
  int rc = cgpe_frames_end(h);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  const DevParams &p = h->p;
  const size_t n = (size_t)p.R * capacity;
  CU(cudaMalloc(&h->frame_alloc[0], sizeof(double) * n * CGPE_FRAME_COLS));
  h->g.frames = static_cast<double *>(h->frame_alloc[0]);
  if ((flags & CGPE_FRAME_BONDS) && p.N > 1) {
    CU(cudaMalloc(&h->frame_alloc[1], sizeof(float) * n * (p.N - 1)));
    h->g.frame_bonds = static_cast<float *>(h->frame_alloc[1]);
  }
  if ((flags & CGPE_FRAME_DIHEDRALS) && p.N > 3) {
    CU(cudaMalloc(&h->frame_alloc[2], sizeof(float) * n * (p.N - 3)));
    h->g.frame_dih = static_cast<float *>(h->frame_alloc[2]);
  }
  CU(cudaMemsetAsync(h->g.n_frames, 0, sizeof(int) * p.R, h->stream));
  if (!keep_phase) CU(cudaMemsetAsync(h->g.frame_phase, 0, sizeof(int) * p.R, h->stream));
  h->p.frame_every = every; h->p.frame_flags = flags; h->p.frame_cap = capacity;
  return CGPE_OK;
}

int cgpe_frames_fetch(cgpe_handle *h, double *frames, float *bonds, float *dihedrals, int32_t *n_frames) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  if (h->p.frame_every <= 0) return fail(h, CGPE_ERR_STATE, "cgpe_frames_begin has not been called");
  DeviceGuard guard(h->device);
  const DevParams &p = h->p;
  const size_t n = (size_t)p.R * p.frame_cap;
  if (frames) CU(cudaMemcpyAsync(frames, h->g.frames, sizeof(double) * n * CGPE_FRAME_COLS, cudaMemcpyDeviceToHost, h->stream));
  if (bonds && h->g.frame_bonds) CU(cudaMemcpyAsync(bonds, h->g.frame_bonds, sizeof(float) * n * (p.N - 1), cudaMemcpyDeviceToHost, h->stream));
  if (dihedrals && h->g.frame_dih) CU(cudaMemcpyAsync(dihedrals, h->g.frame_dih, sizeof(float) * n * (p.N - 3), cudaMemcpyDeviceToHost, h->stream));
  std::vector<int> nf(p.R);
  CU(cudaMemcpyAsync(nf.data(), h->g.n_frames, sizeof(int) * p.R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  int rc = check_device_errors(h);
  if (rc) return rc;
  if (n_frames) for (int r = 0; r < p.R; ++r) n_frames[r] = nf[r];
  CU(cudaMemsetAsync(h->g.n_frames, 0, sizeof(int) * p.R, h->stream));
  return CGPE_OK;
}

int cgpe_last_call_ms(cgpe_handle *h, float *ms, int32_t *n_launches) {
  if (!h) return fail(nullptr, CGPE_ERR_INVALID, "handle is NULL");
  DeviceGuard guard(h->device);
  float t = 0.f;
  if (h->timing_pending) {
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(&t, h->ev0, h->ev1));
  }
  if (ms) *ms = t;
  if (n_launches) *n_launches = h->last_launches;
  return CGPE_OK;
}

}  // extern "C"
