// duo_engine.cu — the cluster engine: one replica = one thread-block cluster of TWO CTAs (sm_100a).
//
// Why: a launch of the one-CTA-per-replica engine (chain_engine.cu) lasts as long as its most charged
// replica while the SMs of the neutral replicas idle (128 replicas on 148 SMs, mean/max cost 0.68: an SM
// was busy 59 % of the launch).  Here every replica is split over two CTAs of half the threads, two
// such CTAs fit on one SM, and the replica -> cluster assignment puts half of a heavy replica next to
// half of a light one (replica_of_cluster): the SMs carry equal sums, and the serial sections of one
// CTA (trial moves on one warp, barriers) are covered by the other CTA's warps.
//
//   * CTA h of the pair owns the contiguous half [lo, hi) of the particles: velocity, force and list
//     reference position in the registers of the owning thread, WCA rows and the Debye-Hueckel rows /
//     compact partner lists of its own particles only;
//   * BOTH CTAs hold all positions (float4 x,y,z,q) and types: after the drift every thread stores its
//     particle's new position into its own copy and, through distributed shared memory, into the
//     peer's.  One cluster barrier per MD step makes them visible (arrive.release / wait.acquire); a
//     second one, split into an early arrive after the force phase and a late wait before the next
//     store, keeps a fast CTA from overwriting positions the peer still reads;
//   * chain-local terms keep the owner-computes scheme; the three owners whose hand-over slots cross
//     the boundary are evaluated on both sides (ghost owners on spare threads), so no force crosses
//     the cluster and the step needs no further barrier;
//   * the trial moves run on warp 0 of CTA 0 (the leader) with its other warps as helpers; CTA 1
//     contributes the per-site caches of its half (mc_prepare) and sleeps in a cluster barrier
//     meanwhile -- its SM share goes to the co-resident CTA of another replica.  The leader keeps the
//     whole Debye-Hueckel bit matrix (CTA 1 mirrors its rows) and mirrors every accepted flip
//     (type, charge, charge mask) into the peer.
//
// Same Markov chain as chain_engine.cu and the oracle: same Philox streams, same list semantics, same
// bookkeeping order.  Implicit-ion replicas without spectator particles only (the headline path);
// everything else stays on chain_engine.cu / wide_engine.cu.  State in HBM between launches is
// engine-agnostic, so the engines can be mixed call by call.
//
// Reference call sites replaced (relative to /root/reference/src/modules):
//   integrator.run            EnsembleDynamics.py:41,52,58,64,73,90,92
//   reaction.reaction         EnsembleDynamics.py:70-71,94-95
//   calc_rg / calc_re / number_of_particles   EnsembleDynamics.py:97-101
#include "duo_engine.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace cgpe {

static inline size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t duo_layout(const DevParams &p, int n_loc, size_t smem_budget, int cap_w_cfg, DuoShape *d) {
  DuoShape &s = *d;
  s.ok = 0;
  s.half = (p.P + 1) / 2;
  s.nt = std::max(32, (s.half + 31) / 32 * 32);
  if (s.nt > 1024) return 0;
  s.n_loc = n_loc > 0 ? n_loc : 1;
  int grp = s.nt / s.n_loc;
  s.G = grp < 1 ? 1 : grp > 4 ? 4 : grp;
  const size_t n_chg = p.n_chg > 0 ? p.n_chg : 1, n_rx = p.n_rx > 0 ? p.n_rx : 1;
  s.dw = (int)((n_chg + 31) / 32) | 1;
  const int TT = (p.T + 1) * (p.T + 1);
  size_t o = 0;
  auto take = [&](size_t bytes, size_t align) { o = up(o, align); size_t at = o; o += bytes; return (int)at; };
  s.pos = take(sizeof(float4) * (p.P + 1), 16);   // +1: far-away, uncharged dummy that pads the compact partner lists
  s.tab = take(sizeof(float4) * TT, 16);
  s.red = take(sizeof(double) * 32 * 4, 16);   // also holds the block spheres of a list build (float4)
  s.slot = take(sizeof(float) * 9 * s.nt, 4);
  s.fd = take(sizeof(float) * 3 * s.G * s.n_loc, 4);
  s.dhn = take(sizeof(uint16_t) * s.G * s.n_loc, 2);
  s.cnt = take(sizeof(int) * kMaxRx * 2, 4);
  s.flag = take(sizeof(int) * 4, 4);
  s.db = take(sizeof(uint32_t) * (size_t)s.dw * n_chg, 4);
  s.qm = take(sizeof(uint32_t) * (size_t)s.dw, 4);
  s.chg = take(sizeof(uint16_t) * n_chg, 2);
  s.cidx = take(sizeof(uint16_t) * p.P, 2);
  s.lp = take(sizeof(uint16_t) * n_rx * n_chg, 2);
  s.ld = take(sizeof(uint16_t) * n_rx * n_chg, 2);
  s.wn = take((size_t)s.nt, 1);
  s.type = take((size_t)p.P, 1);
  s.phi = take(sizeof(float) * n_chg, 4);
  s.wc = take(sizeof(float) * n_chg, 4);
  s.rxof = take(n_chg, 1);
  s.wrange = take(2 * n_chg, 1);
  s.wflag = take(n_chg, 1);
  s.ratio = take(sizeof(float) * n_rx * (n_chg + 1), 4);
  s.ord = take(sizeof(float) * 8, 4);
  s.qloc = take(sizeof(uint16_t) * (s.n_loc + 2), 2);   // owned chargeable particles that carry charge (local index) + their number
  // the WCA rows take what is left, up to 64 entries per particle (two CTAs per SM when nt <= 512)
  const size_t fixed = up(o, 4);
  long cap = cap_w_cfg > 0 ? cap_w_cfg : 0;
  if (cap == 0) {
    cap = smem_budget > fixed + 64 ? (long)((smem_budget - fixed - 64) / (2 * (size_t)s.nt)) : 0;
    if (cap > 64) cap = 64;
  }
  if (cap > (p.P > 1 ? p.P - 1 : 1)) cap = p.P > 1 ? p.P - 1 : 1;
  if (cap < 4 && cap < p.P - 1) return 0;
  s.cap_w = (int)cap;
  s.wl = take(sizeof(uint16_t) * (size_t)s.cap_w * s.nt, 4);
  s.total = (int)up(o, 16);
  if ((size_t)s.total > smem_budget) return 0;
  s.dhc_stride = (s.G * s.n_loc + 31) / 32 * 32;
  s.dhc_cap = (((32 + s.G - 1) / s.G) * (int)((n_chg + 31) / 32) + 7) / 8 * 8;
  s.ok = 1;
  return (size_t)s.total;
}

namespace {

struct Sm {
  float4 *pos, *tab;
  double *red;
  float *slot, *fd;
  uint16_t *dhn;
  int *cnt, *flag;
  uint16_t *wl;
  uint32_t *dbits, *qmask;
  uint16_t *chg, *cidx, *lp, *ld;
  unsigned char *wn, *type;
  float *phi, *wcache;
  unsigned char *rxof, *wrange, *wflag;
  float *ratio, *ord;
  uint16_t *qloc;
};

struct Thr {
  float x[3], v[3], f[3], ref[3];
  float q;
  int i;    // owned particle, -1 = none
  int l;    // index within its chain
};

struct Ctx {
  const DevParams &p;
  const DevState &g;
  const DuoShape &d;
  Sm s;               // this CTA's shared memory
  Sm q;               // the peer's, through the cluster's distributed shared memory (generic pointers)
  int r;              // replica (local index)
  int rank;           // CTA rank in the pair; 0 = leader
  int tid, nt, lane, warp;
  int lo, hi;         // owned particles
  int clo, chi;       // owned chargeable indices
  float kappa, rc2, neg_kappa_log2e, rlist_d2, rlist_max, gamma, noise_pref;
  bool wrap, compact;
  int err;
  long long rebuilds;
};

// cluster barrier, split form.  Every thread alternates arrive / wait strictly.
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cl_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ void cl_sync() { cl_arrive(); cl_wait(); }

__device__ __forceinline__ uint32_t sel_mask(int G, int g) {
  return G == 1 ? 0xFFFFFFFFu : G == 2 ? (0x55555555u << g) : G == 3 ? (0x49249249u << g) : (0x11111111u << g);
}

template <bool WRAP>
__device__ __forceinline__ float dist2(const Ctx &c, float4 a, float4 b, V3 *d) {
  if (WRAP) {
    d->x = min_image(a.x - b.x, c.p.box_l, c.p.inv_box_l);
    d->y = min_image(a.y - b.y, c.p.box_l, c.p.inv_box_l);
    d->z = min_image(a.z - b.z, c.p.box_l, c.p.inv_box_l);
  } else {
    d->x = a.x - b.x; d->y = a.y - b.y; d->z = a.z - b.z;
  }
  return fmaf(d->x, d->x, fmaf(d->y, d->y, d->z * d->z));
}

extern __shared__ __align__(16) unsigned char duo_smem[];

__device__ __forceinline__ void bind(const DuoShape &d, unsigned char *base, Sm *s) {
  s->pos = reinterpret_cast<float4 *>(base + d.pos);
  s->tab = reinterpret_cast<float4 *>(base + d.tab);
  s->red = reinterpret_cast<double *>(base + d.red);
  s->slot = reinterpret_cast<float *>(base + d.slot);
  s->fd = reinterpret_cast<float *>(base + d.fd);
  s->dhn = reinterpret_cast<uint16_t *>(base + d.dhn);
  s->cnt = reinterpret_cast<int *>(base + d.cnt);
  s->flag = reinterpret_cast<int *>(base + d.flag);
  s->wl = reinterpret_cast<uint16_t *>(base + d.wl);
  s->dbits = reinterpret_cast<uint32_t *>(base + d.db);
  s->qmask = reinterpret_cast<uint32_t *>(base + d.qm);
  s->chg = reinterpret_cast<uint16_t *>(base + d.chg);
  s->cidx = reinterpret_cast<uint16_t *>(base + d.cidx);
  s->lp = reinterpret_cast<uint16_t *>(base + d.lp);
  s->ld = reinterpret_cast<uint16_t *>(base + d.ld);
  s->wn = base + d.wn;
  s->type = base + d.type;
  s->phi = reinterpret_cast<float *>(base + d.phi);
  s->wcache = reinterpret_cast<float *>(base + d.wc);
  s->rxof = base + d.rxof;
  s->wrange = base + d.wrange;
  s->wflag = base + d.wflag;
  s->ratio = reinterpret_cast<float *>(base + d.ratio);
  s->ord = reinterpret_cast<float *>(base + d.ord);
  s->qloc = reinterpret_cast<uint16_t *>(base + d.qloc);
}

// ---- which replica does a cluster take? ------------------------------------------------------------
// With two CTAs resident per SM the hardware places CTA b and CTA b + n_sm on the same SM (measured on B200:
// smid = (b + 142) mod 148 for cluster launches, profiles/r2_ubench_issue.txt), i.e. cluster c shares its two SMs
// with cluster c + n_sm/2; with R <= n_sm replicas, n_sm - R clusters have their SMs to themselves.  A launch lasts as
// long as its slowest cluster, so the plan is a small makespan problem with KNOWN job lengths:
//   * MD steps of the launch: the burst schedule is a Philox stream of (seed, replica, sample index) - the planner
//     draws the same numbers the kernel will draw (fused sampling launches; other launches: equal steps);
//   * cost per MD step: a chain-local part + the charged Debye-Hueckel pairs inside the list radius, counted from the
//     state the launch starts from (k_duo_count_pairs: all pairs of charged sites, ~10 us for 128 replicas) - not timed:
//     a CTA's time depends on whom it shares its SM with;
//   * trial moves: the acceptance measured so far (an accepted flip costs 2-3x a rejected trial).
// Constants in microseconds from tools/prof_phases.py / prof_mc.py at N = 1000; only their ratios matter.
// Assignment: n_alone consecutive replicas of the cost order stay alone, the rest pair k-th heaviest with k-th
// lightest; the start of the alone block is the one with the smallest predicted makespan (a replica alone runs at
// ~0.97 of its one-CTA time; two that share take ~1.2x while both run, after which the longer one finishes alone:
// profiles/r2_timeline_solo_vs_duo.txt).  The launch cannot end before its most expensive replica does.  With more
// clusters than the device holds at once: heaviest first.  A heuristic only: replicas are independent, results do
// not depend on it.  Its own small launch in front of every engine launch: the counts it reads change while the
// engine kernel runs.
__global__ void __launch_bounds__(256) k_duo_count_pairs(const __grid_constant__ DevParams p, const __grid_constant__ DevState g) {
  extern __shared__ float4 qpos[];   // the charged chargeable particles of the replica
  __shared__ int n_q;
  __shared__ unsigned warp_cnt[8];
  const int r = blockIdx.x;
  if (threadIdx.x == 0) n_q = 0;
  __syncthreads();
  for (int ci = threadIdx.x; ci < p.n_chg; ci += blockDim.x) {
    const int i = g.chg[(size_t)r * p.n_chg + ci];
    const float4 x = g.pos[(size_t)r * p.P + i];
    if (x.w != 0.f && g.type[(size_t)r * p.P + i] < p.T) qpos[atomicAdd(&n_q, 1)] = x;
  }
  __syncthreads();
  const int n = n_q;
  const float rl = g.rcut[r] + p.skin, rl2 = rl * rl;
  unsigned cnt = 0;
  for (int a = threadIdx.x; a < n; a += blockDim.x) {
    const float4 pa = qpos[a];
    for (int b = a + 1; b < n; ++b) {
      const float4 pb = qpos[b];
      const float dx = min_image(pa.x - pb.x, p.box_l, p.inv_box_l), dy = min_image(pa.y - pb.y, p.box_l, p.inv_box_l),
                  dz = min_image(pa.z - pb.z, p.box_l, p.inv_box_l);
      cnt += fmaf(dx, dx, fmaf(dy, dy, dz * dz)) < rl2 ? 1u : 0u;
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_cnt[w];
    g.plan[(size_t)r * 8 + 4] = 2.0 * (double)t;   // directed pairs, as the partner lists hold them
  }
}

struct PlanJob { int n_samples, steps_per_burst, use_md, trials_per_sample; uint32_t thr1, thr2, seed; int fixed_steps; };

__global__ void k_duo_plan(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, int n_sm, int *order,
                           const __grid_constant__ PlanJob job) {
  extern __shared__ float plan_smem[];
  float *cost = plan_smem, *cand = plan_smem + p.R;
  int *srt = reinterpret_cast<int *>(plan_smem + 2 * p.R);
  __shared__ int best_j;
  const int R = p.R, hs = n_sm / 2;
  const float kStepUs = 6.4f, kPairUs = 2.4e-4f, kTrialUs = 0.33f, kFlipUs = 0.45f, kAlone = 0.97f, kShare = 0.2f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    int w = 0;
    long long tr = 0, ac = 0;
    for (int m = 0; m < p.n_rx; ++m) {
      const int np = g.cnt[((size_t)r * p.n_rx + m) * 2], nd = g.cnt[((size_t)r * p.n_rx + m) * 2 + 1];
      w += (p.rx[m].q_prot != 0.f ? np : 0) + (p.rx[m].q_deprot != 0.f ? nd : 0);
      tr += g.trials[(size_t)r * p.n_rx + m]; ac += g.accepted[(size_t)r * p.n_rx + m];
    }
    double *st = g.plan + (size_t)r * 8;
    if ((double)tr > st[2]) st[5] = ((double)ac - st[3]) / ((double)tr - st[2]);
    else if (st[2] == 0.0) st[5] = 0.5;
    st[2] = (double)tr; st[3] = (double)ac;
    const float pairs = p.use_dh ? (float)st[4] : 0.f;   // k_duo_count_pairs
    long long steps = job.fixed_steps;
    if (job.n_samples > 20000 && job.use_md) {
      // a very long launch: the schedules of the replicas differ by less than a per cent, the expectation will do
      steps = (long long)((double)job.n_samples * job.steps_per_burst * ((double)job.thr1 + (double)job.thr2) / 16777216.0);
    } else if (job.n_samples > 0 && job.use_md) {
      steps = 0;
      const unsigned long long s0 = (unsigned long long)g.samples_done[r];
      for (int s = 0; s < job.n_samples; ++s) {
        const unsigned long long sc = s0 + (unsigned long long)s;
        const uint4 o = philox4x32_10((uint32_t)sc, (uint32_t)(sc >> 32), 0u, kTagSchedule, job.seed, (uint32_t)(p.replica_offset + r));
        steps += ((o.x >> 8) < job.thr1 ? job.steps_per_burst : 0) + ((o.y >> 8) < job.thr2 ? job.steps_per_burst : 0);
      }
    }
    cost[r] = (float)steps * (kStepUs + kPairUs * pairs)
              + (float)job.n_samples * (float)job.trials_per_sample * (kTrialUs + kFlipUs * (float)st[5])
              + 1.0e-3f * (float)w;   // launches without steps or trials: order by charged sites as before
    st[6] = cost[r]; st[7] = (double)steps;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float w = cost[r];
    int k = 0;   // sorted rank, 0 = heaviest
    for (int j = 0; j < R; ++j) {
      const float wj = cost[j];
      k += (wj > w || (wj == w && j < r)) ? 1 : 0;
    }
    srt[k] = r;
  }
  __syncthreads();
  if (R <= hs || R > 2 * hs) {
    for (int k = threadIdx.x; k < R; k += blockDim.x) order[R <= hs ? srt[k] : k] = srt[k];   // all alone / heaviest first
    if (threadIdx.x == 0 && g.engine_choice) *g.engine_choice = 1;
    return;
  }
  const int n_pair = R - hs, n_alone = 2 * hs - R;
  // candidate j: the alone block is srt[j .. j + n_alone); rest(m) = srt[m < j ? m : m + n_alone]
  const int n_cand = n_alone > 0 ? R - n_alone + 1 : 1;   // no cluster alone: one way to pair
  for (int j = threadIdx.x; j < n_cand; j += blockDim.x) {
    float worst = n_alone > 0 ? kAlone * cost[srt[j]] : 0.f;
    for (int i = 0; i < n_pair; ++i) {
      const int a = i, b = 2 * n_pair - 1 - i;
      // while both run they advance at ~1/(1 + kShare) of their lone pace; the longer one finishes alone
      worst = fmaxf(worst, cost[srt[a < j ? a : a + n_alone]] + kShare * cost[srt[b < j ? b : b + n_alone]]);
    }
    cand[j] = worst;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int bj = 0;
    for (int j = 1; j < n_cand; ++j) if (cand[j] < cand[bj]) bj = j;
    best_j = bj;
    // Which engine for this launch?  One CTA per replica: every replica on its own SM, the launch lasts as long as the most
    // expensive one.  Clusters: cand[bj].  A short launch has a few outliers (long burst schedules) that gain from two SMs of
    // their own; over a long one all highly charged replicas cost about the same, only n_alone of them can be alone and the
    // others lose more by sharing than the alone ones gain (measured: full 4571-sample protocol 21.1 s against 24.1 s).
    if (g.engine_choice) *g.engine_choice = cand[bj] < cost[srt[0]] ? 1 : 0;
  }
  __syncthreads();
  const int j0 = best_j;
  for (int k = threadIdx.x; k < R; k += blockDim.x) {
    int cluster;
    if (k >= j0 && k < j0 + n_alone) cluster = n_pair + (k - j0);          // alone on its two SMs
    else {
      const int m = k < j0 ? k : k - n_alone;                              // index among the sharing replicas, heaviest first
      cluster = m < n_pair ? m : hs + (2 * n_pair - 1 - m);                // m-th heaviest next to the m-th lightest
    }
    order[cluster] = srt[k];
  }
}

__device__ void init_ctx(Ctx &c) {
  const DevParams &p = c.p;
  c.tid = threadIdx.x; c.nt = blockDim.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
  c.lo = c.rank == 0 ? 0 : c.d.half;
  c.hi = c.rank == 0 ? min(c.d.half, p.P) : p.P;
  c.kappa = c.g.kappa[c.r];
  const float rc = c.g.rcut[c.r];
  c.rc2 = rc * rc;
  c.neg_kappa_log2e = -c.kappa * kLog2e;
  c.rlist_d2 = (rc + p.skin) * (rc + p.skin);
  c.rlist_max = fmaxf(p.use_dh ? rc + p.skin : 0.f, p.use_wca ? sqrtf(p.rlist_wf2) : 0.f);
  c.wrap = true;
  c.compact = false;
  c.gamma = c.g.gamma[c.r];
  c.noise_pref = sqrtf(24.0f * c.gamma * c.g.kT[c.r] / p.dt);
  c.err = 0;
  c.rebuilds = 0;
}

// ---- load / store ------------------------------------------------------------------------------
// Local work only (the peer may not have started yet); the caller puts a cluster barrier behind it.
__device__ void load_state(Ctx &c, Thr &t, bool f_valid) {
  const DevParams &p = c.p;
  const size_t base = (size_t)c.r * p.P;
  for (int i = c.tid; i < p.P; i += c.nt) {
    c.s.pos[i] = c.g.pos[base + i];
    c.s.type[i] = c.g.type[base + i];
    c.s.cidx[i] = 0xFFFFu;
  }
  t.i = (c.lo + c.tid < c.hi) ? c.lo + c.tid : -1;
  t.l = -1;
  t.q = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) { t.x[a] = 0.f; t.v[a] = 0.f; t.f[a] = 0.f; t.ref[a] = 0.f; }
  if (t.i >= 0) {
    const float4 pq = c.g.pos[base + t.i], vv = c.g.vel[base + t.i];
    t.x[0] = pq.x; t.x[1] = pq.y; t.x[2] = pq.z; t.q = pq.w;
    t.v[0] = vv.x; t.v[1] = vv.y; t.v[2] = vv.z;
    if (f_valid) { const float4 ff = c.g.frc[base + t.i]; t.f[0] = ff.x; t.f[1] = ff.y; t.f[2] = ff.z; }
    t.l = t.i < p.NC ? (p.NC == p.N ? t.i : t.i % p.N) : -1;
  }
  if (c.tid < c.nt) c.s.wn[c.tid] = 0;
  for (int i = c.tid; i < 9 * c.nt; i += c.nt) c.s.slot[i] = 0.0f;
  for (int i = c.tid; i < (p.T + 1) * (p.T + 1); i += c.nt) c.s.tab[i] = c.g.wca_tab[i];
  for (int i = c.tid; i < p.n_chg; i += c.nt) {
    c.s.chg[i] = (uint16_t)c.g.chg[(size_t)c.r * p.n_chg + i];
    c.s.wrange[2 * i] = 0; c.s.wrange[2 * i + 1] = 0;
  }
  for (int i = c.tid; i < 3 * c.d.G * c.d.n_loc; i += c.nt) c.s.fd[i] = 0.0f;
  for (int m = 0; m < p.n_rx; ++m) {
    const int *gp = c.g.lst_prot + ((size_t)c.r * p.n_rx + m) * p.P;
    const int *gd = c.g.lst_deprot + ((size_t)c.r * p.n_rx + m) * p.P;
    const int np = c.g.cnt[((size_t)c.r * p.n_rx + m) * 2], nd = c.g.cnt[((size_t)c.r * p.n_rx + m) * 2 + 1];
    for (int i = c.tid; i < np; i += c.nt) c.s.lp[(size_t)m * p.n_chg + i] = (uint16_t)gp[i];
    for (int i = c.tid; i < nd; i += c.nt) c.s.ld[(size_t)m * p.n_chg + i] = (uint16_t)gd[i];
    if (c.tid == 0) { c.s.cnt[m * 2] = np; c.s.cnt[m * 2 + 1] = nd; }
  }
  if (c.tid == 0) {
    c.s.pos[p.P] = make_float4(1e10f, 1e10f, 1e10f, 0.0f);   // list padding (build_dh_compact)
    c.s.flag[0] = 0; c.s.flag[1] = 0;
  }
  __syncthreads();
  for (int i = c.tid; i < p.n_chg; i += c.nt) c.s.cidx[c.s.chg[i]] = (uint16_t)i;
  // live mask of the chargeable particles that carry charge right now (the leader's MC warp keeps both copies current)
  for (int w = c.tid; w < c.d.dw; w += c.nt) {
    uint32_t m = 0;
    for (int b = 0; b < 32; ++b) {
      const int cj = w * 32 + b;
      if (cj < p.n_chg && c.s.pos[c.s.chg[cj]].w != 0.0f) m |= 1u << b;
    }
    c.s.qmask[w] = m;
  }
  // owned chargeable indices: chg is ascending in particle id
  {
    int a = 0, b = p.n_chg;
    while (a < b) { const int mid = (a + b) >> 1; if ((int)c.s.chg[mid] < c.d.half) a = mid + 1; else b = mid; }
    c.clo = c.rank == 0 ? 0 : a;
    c.chi = c.rank == 0 ? a : p.n_chg;
  }
  __syncthreads();
}

__device__ void store_state(Ctx &c, Thr &t, bool store_forces) {
  const DevParams &p = c.p;
  const size_t base = (size_t)c.r * p.P;
  __syncthreads();
  if (t.i >= 0) {
    const int i = t.i;
    c.g.pos[base + i] = make_float4(t.x[0], t.x[1], t.x[2], c.s.pos[i].w);
    c.g.vel[base + i] = make_float4(t.v[0], t.v[1], t.v[2], 0.f);
    if (store_forces) c.g.frc[base + i] = make_float4(t.f[0], t.f[1], t.f[2], 0.f);
    c.g.type[base + i] = c.s.type[i];
    if (!isfinite(t.x[0] + t.x[1] + t.x[2] + t.v[0] + t.v[1] + t.v[2])) c.err |= kErrNonFinite;
    if (t.v[0] * t.v[0] + t.v[1] * t.v[1] + t.v[2] * t.v[2] > 1.0e6f * fmaxf(c.g.kT[c.r], 1.0f)) c.err |= kErrDiverged;
  }
  if (c.rank == 0) {
    for (int m = 0; m < p.n_rx; ++m) {
      int *gp = c.g.lst_prot + ((size_t)c.r * p.n_rx + m) * p.P;
      int *gd = c.g.lst_deprot + ((size_t)c.r * p.n_rx + m) * p.P;
      const int np = c.s.cnt[m * 2], nd = c.s.cnt[m * 2 + 1];
      for (int i = c.tid; i < np; i += c.nt) gp[i] = c.s.lp[(size_t)m * p.n_chg + i];
      for (int i = c.tid; i < nd; i += c.nt) gd[i] = c.s.ld[(size_t)m * p.n_chg + i];
      if (c.tid == 0) { c.g.cnt[((size_t)c.r * p.n_rx + m) * 2] = np; c.g.cnt[((size_t)c.r * p.n_rx + m) * 2 + 1] = nd; }
    }
    if (c.tid == 0 && c.rebuilds) c.g.rebuilds[c.r] += c.rebuilds;
  }
  if (c.err) atomicOr(&c.g.err[c.r], c.err);
}

// ---- Verlet lists ------------------------------------------------------------------------------
// Same lists as chain_engine.cu: WCA rows (uint16 ids, column per particle) of the OWNED particles
// without the chain neighbours i+-1, i+-2; Debye-Hueckel bit rows of the OWNED chargeable particles,
// CTA 1 mirroring its rows into the leader's matrix (the trial moves need every row).
// Requires the local position copy current; every distance test runs on it.
template <bool WRAP>
__device__ void build_rows(Ctx &c, Thr &t) {
  const DevParams &p = c.p;
  if (t.i >= 0) {
    t.ref[0] = t.x[0]; t.ref[1] = t.x[1]; t.ref[2] = t.x[2];
    if (p.use_wca) {
      const int i = t.i, li = i - c.lo;
      const float4 pi = c.s.pos[i];
      int ex_lo = i, ex_hi = i;
      if (t.l >= 0) { ex_lo = i - min(2, t.l); ex_hi = i + min(2, p.N - 1 - t.l); }
      int n = 0;
      const float4 *sph = reinterpret_cast<const float4 *>(c.s.red);   // block spheres (build_lists)
      const float rl_w = sqrtf(p.rlist_w2);
      for (int b = 0, nb = (p.P + 31) >> 5; b < nb; ++b) {   // blocks of 32 consecutive particles: sphere first
        const float4 sb = sph[b];
        V3 dd;
        const float reach = sb.w + rl_w;
        if (dist2<WRAP>(c, pi, sb, &dd) > reach * reach) continue;
        const int j1 = min(p.P, b * 32 + 32);
        for (int j = b * 32; j < j1; ++j) {
          const float r2 = dist2<WRAP>(c, pi, c.s.pos[j], &dd);
          if (r2 < p.rlist_w2 && (j < ex_lo || j > ex_hi)) {
            if (n < c.d.cap_w) c.s.wl[(size_t)n * c.nt + li] = (uint16_t)j;
            ++n;
          }
        }
      }
      if (n > c.d.cap_w) { c.err |= kErrWcaCap; n = c.d.cap_w; }
      c.s.wn[li] = (unsigned char)n;
    }
  }
  if (p.use_dh) {
    const int nw = (p.n_chg + 31) >> 5, n_mine = c.chi - c.clo;
    // work item = (word, row), rows fastest: the lanes of a warp test one partner against 32 rows
    for (int item = c.tid; item < n_mine * nw; item += c.nt) {
      const int w = item / n_mine, ci = c.clo + (item - w * n_mine);
      const float4 pi = c.s.pos[c.s.chg[ci]];
      uint32_t m = 0;
      const int nb = min(32, p.n_chg - w * 32);
      for (int b = 0; b < nb; ++b) {
        V3 dd;
        const float r2 = dist2<WRAP>(c, pi, c.s.pos[c.s.chg[w * 32 + b]], &dd);
        if (r2 < c.rlist_d2) m |= 1u << b;
      }
      if ((ci >> 5) == w) m &= ~(1u << (ci & 31));   // not its own partner
      c.s.dbits[(size_t)ci * c.d.dw + w] = m;
      if (c.rank != 0) c.q.dbits[(size_t)ci * c.d.dw + w] = m;
    }
    __syncthreads();
    for (int ci = c.clo + c.tid; ci < c.chi; ci += c.nt) {
      int w_lo = nw, w_hi = 0;
      for (int w = 0; w < nw; ++w)
        if (c.s.dbits[(size_t)ci * c.d.dw + w]) { w_lo = min(w_lo, w); w_hi = w + 1; }
      c.s.wrange[2 * ci] = (unsigned char)min(w_lo, w_hi);
      c.s.wrange[2 * ci + 1] = (unsigned char)w_hi;
    }
  }
}

// Both CTAs run this in step.  Contains block barriers.
__device__ void build_lists(Ctx &c, Thr &t) {
  const DevParams &p = c.p;
  // bounding box of the replica from the local copy of all positions: both CTAs find the same answer
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int j = c.tid; j < p.P; j += c.nt) {
    const float4 x = c.s.pos[j];
    lo[0] = fminf(lo[0], x.x); hi[0] = fmaxf(hi[0], x.x);
    lo[1] = fminf(lo[1], x.y); hi[1] = fmaxf(hi[1], x.y);
    lo[2] = fminf(lo[2], x.z); hi[2] = fmaxf(hi[2], x.z);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  float *scr = reinterpret_cast<float *>(c.s.red);   // 32 warps x 6 floats fit in the scratch
  __syncthreads();
  if (c.lane == 0)
#pragma unroll
    for (int a = 0; a < 3; ++a) { scr[c.warp * 6 + a] = lo[a]; scr[c.warp * 6 + 3 + a] = hi[a]; }
  __syncthreads();
  float ext = 0.f;
  {
    const int nwarps = (c.nt + 31) >> 5;
    for (int a = 0; a < 3; ++a) {
      float l_ = 3.0e38f, h_ = -3.0e38f;
      for (int w = 0; w < nwarps; ++w) { l_ = fminf(l_, scr[w * 6 + a]); h_ = fmaxf(h_, scr[w * 6 + 3 + a]); }
      ext = fmaxf(ext, h_ - l_);
    }
  }
  c.wrap = !(ext + c.rlist_max + p.skin < p.box_l);   // NaN-safe: wraps
  __syncthreads();   // every thread has read the bounding-box scratch: it now takes the block spheres
  block_spheres(c.s.pos, p.P, reinterpret_cast<float4 *>(c.s.red), c.tid, c.nt);
  __syncthreads();
  if (c.wrap) build_rows<true>(c, t); else build_rows<false>(c, t);
  __syncthreads();
  c.rebuilds += 1;
  c.compact = false;
}

// Partners of every (owned site, sub-thread) that are listed AND charged, written out once per list
// build / burst start as byte offsets into the position array: [k / 8][sub-thread][k % 8] in global
// memory (L2-resident), read back by the thread that wrote them, eight per 16-byte load.
__device__ void build_dh_compact(Ctx &c) {
  const DevParams &p = c.p;
  c.compact = false;
  if (!p.use_dh || c.g.dhc == nullptr) return;
  // Work items exist for the owned particles that carry charge only, packed (chain_engine.cu::build_dh_compact): warp 0
  // lists them, the force slots of the others are zeroed once, nobody writes them until the next build.  Called by all
  // threads of the CTA (one block barrier inside).
  const int n_mine = c.chi - c.clo, n_loc = c.d.n_loc;
  if (c.warp == 0) {
    int base = 0;
    for (int k0 = 0; k0 < n_mine; k0 += 32) {
      const int cl = k0 + c.lane;
      const bool q = cl < n_mine && c.s.pos[c.s.chg[c.clo + cl]].w != 0.0f;
      const uint32_t m = __ballot_sync(0xffffffffu, q);
      if (q) c.s.qloc[base + __popc(m & ((1u << c.lane) - 1u))] = (uint16_t)cl;
      base += __popc(m);
    }
    if (c.lane == 0) c.s.qloc[n_loc] = (uint16_t)base;
  }
  for (int cl = c.tid; cl < n_mine; cl += c.nt)
    if (c.s.pos[c.s.chg[c.clo + cl]].w == 0.0f)
      for (int k = 0; k < 3 * c.d.G; ++k) c.s.fd[(size_t)k * n_loc + cl] = 0.0f;
  __syncthreads();
  const int nq = c.s.qloc[n_loc];
  const int G = c.d.G;
  uint16_t *lst0 = c.g.dhc + ((size_t)c.r * 2 + c.rank) * c.d.dhc_cap * c.d.dhc_stride;
  for (int idx = c.tid; idx < G * nq; idx += c.nt) {
    const int g = idx / nq, ci = c.clo + (int)c.s.qloc[idx - g * nq];
    const uint32_t sel = sel_mask(G, g);
    uint16_t *lst = lst0 + (size_t)idx * 8;
    int n = 0;
    const int w_end = c.s.wrange[2 * ci + 1];
    for (int w = c.s.wrange[2 * ci]; w < w_end; ++w) {
      uint32_t m = c.s.dbits[(size_t)ci * c.d.dw + w] & c.s.qmask[w] & sel;
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        lst[((size_t)(n >> 3) * c.d.dhc_stride) * 8 + (n & 7)] = (uint16_t)((int)c.s.chg[w * 32 + b] * 16);
        ++n;
      }
    }
    c.s.dhn[idx] = (uint16_t)n;
    for (; n & 3; ++n) lst[((size_t)(n >> 3) * c.d.dhc_stride) * 8 + (n & 7)] = (uint16_t)(p.P * 16);   // whole half groups
  }
  c.compact = true;   // every list is read only by the thread that wrote it: no barrier
  if (c.rank == 0 && c.tid == 0) c.g.compact_builds[c.r] += 1;
}

// ---- forces ------------------------------------------------------------------------------------
// Everything that is local to the chain, evaluated by the owner of bead i (chain_engine.cu): bond (i-1,i),
// angle (i-1,i,i+1), dihedral (i-1,i,i+1,i+2), WCA of (i-1,i) and (i-1,i+1).  Returns the owner's share and
// puts the partners' shares into the slots of the particles this CTA owns.
__device__ __forceinline__ V3 owner_terms(Ctx &c, int i, int l) {
  const DevParams &p = c.p;
  const int S = p.T + 1, nt = c.nt;
  float *sa = c.s.slot, *sb = c.s.slot + 3 * nt, *sc = c.s.slot + 6 * nt;
  V3 f = {0.f, 0.f, 0.f};
  const V3 pi = xyz(c.s.pos[i]);
  const int ti = c.s.type[i];
  const V3 pm1 = xyz(c.s.pos[i - 1]);
  const int tm1 = c.s.type[i - 1];
  V3 to_prev = {0.f, 0.f, 0.f}, to_next = {0.f, 0.f, 0.f};
  const V3 d01 = pi - pm1;
  if (p.bond_kind != CGPE_BOND_NONE) {  // bond (i-1,i): M.py:159-160
    V3 fb;
    bond_term(p, pi, pm1, &fb, &c.err);
    f = f + fb;
    to_prev = to_prev - fb;
  }
  if (p.use_wca) {                      // WCA of the bonded pair (no exclusions in the reference)
    const float4 tab = c.s.tab[ti * S + tm1];
    const float r2 = dot(d01, d01);
    if (r2 < tab.z) {
      const V3 fw = wca_force_over_r(tab, r2) * d01;
      f = f + fw;
      to_prev = to_prev - fw;
    }
  }
  const bool has_angle = p.use_angle && l <= p.N - 2;  // M.py:161-165
  const bool has_dih = p.use_dih && l <= p.N - 3;      // M.py:166-170
  if (l <= p.N - 2 && (p.use_angle || p.use_dih || p.use_wca)) {
    const V3 pp1 = xyz(c.s.pos[i + 1]);
    if (p.use_wca) {                    // WCA of the next-nearest pair (i-1,i+1)
      const float4 tab = c.s.tab[tm1 * S + c.s.type[i + 1]];
      const V3 dd = pp1 - pm1;
      const float r2 = dot(dd, dd);
      if (r2 < tab.z) {
        const V3 fw = wca_force_over_r(tab, r2) * dd;
        to_next = to_next + fw;
        to_prev = to_prev - fw;
      }
    }
    if (has_angle) {
      V3 fl, fm, fr;
      angle_term(p, pm1, pi, pp1, &fl, &fm, &fr);
      to_prev = to_prev + fl; f = f + fm; to_next = to_next + fr;
    }
    if (has_dih) {
      const V3 pp2 = xyz(c.s.pos[i + 2]);
      V3 f1, f2, f3, f4;
      dihedral_term(p, pm1, pi, pp1, pp2, &f1, &f2, &f3, &f4);
      to_prev = to_prev + f1; f = f + f2; to_next = to_next + f3;
      if (i + 2 >= c.lo && i + 2 < c.hi) { const int k = i + 2 - c.lo; sc[k] = f4.x; sc[nt + k] = f4.y; sc[2 * nt + k] = f4.z; }
    }
    if (i + 1 >= c.lo && i + 1 < c.hi) { const int k = i + 1 - c.lo; sb[k] = to_next.x; sb[nt + k] = to_next.y; sb[2 * nt + k] = to_next.z; }
  }
  if (i - 1 >= c.lo && i - 1 < c.hi) { const int k = i - 1 - c.lo; sa[k] = to_prev.x; sa[nt + k] = to_prev.y; sa[2 * nt + k] = to_prev.z; }
  return f;
}

// Phase 1 reads the local position copy (must be current and visible), phase 2 gathers the slots and adds
// the thermostat: -gamma v + sqrt(24 gamma kT/dt)(U-1/2), U from Philox(counter, particle) (E.py:61-63).
// arrive_between: signal the cluster after the last read of the positions (see md_steps).
template <bool WRAP>
__device__ void compute_forces_impl(Ctx &c, Thr &t, bool thermostat, unsigned long long counter, bool arrive_between) {
  const DevParams &p = c.p;
  const int S = p.T + 1, nt = c.nt;
  V3 f = {0.f, 0.f, 0.f};
  if (t.i >= 0) {
    const int i = t.i;
    if (t.l >= 1) f = owner_terms(c, i, t.l);
    if (p.use_wca) {
      const float4 pi4 = c.s.pos[i];
      const float4 *tab_i = c.s.tab + (int)c.s.type[i] * S;
      const int n = c.s.wn[i - c.lo];
      for (int e = 0; e < n; ++e) {
        const int j = c.s.wl[(size_t)e * nt + (i - c.lo)];
        const float4 pj4 = c.s.pos[j];
        const float4 tab = tab_i[c.s.type[j]];
        V3 dd;
        const float r2 = dist2<WRAP>(c, pi4, pj4, &dd);
        if (r2 < tab.z) f = f + wca_force_over_r(tab, r2) * dd;
      }
    }
  }
  // ghost owners: the owners outside [lo, hi) whose hand-over slots reach inside it (lo-2, lo-1, hi), on the
  // last threads of the CTA (spare ones whenever the half does not fill the block)
  {
    const int k = nt - 1 - c.tid;
    if (k < 3) {
      const int gi = k == 0 ? c.hi : c.lo - 1 - (k - 1);
      if (gi >= 1 && gi < p.NC && (gi < c.lo || gi >= c.hi)) {
        const int gl = p.NC == p.N ? gi : gi % p.N;
        if (gl >= 1) (void)owner_terms(c, gi, gl);
      }
    }
  }
  if (p.use_dh) {
    // Debye-Hueckel: G sub-threads per owned row, each with its own compact partner list
    const int G = c.d.G, n_mine = c.chi - c.clo, n_loc = c.d.n_loc;
    const int nq = c.compact ? (int)c.s.qloc[n_loc] : n_mine;   // compact lists: items of the charged particles only, packed
    for (int idx = c.tid; idx < G * nq; idx += nt) {
      const int g = idx / nq, cl = c.compact ? (int)c.s.qloc[idx - g * nq] : idx - g * nq, ci = c.clo + cl;
      const float4 pi4 = c.s.pos[c.s.chg[ci]];
      V3 fd = {0.f, 0.f, 0.f};
      if (pi4.w != 0.0f && c.compact) {
        const int n = c.s.dhn[idx];
        const uint4 *lst = reinterpret_cast<const uint4 *>(c.g.dhc + ((size_t)c.r * 2 + c.rank) * c.d.dhc_cap * c.d.dhc_stride) + idx;
        uint32_t pos_sa = (uint32_t)__cvta_generic_to_shared(c.s.pos);
        float nk = c.neg_kappa_log2e, rc2 = c.rc2;
        asm volatile("" : "+r"(pos_sa), "+f"(nk), "+f"(rc2));
        uint4 cur = n > 0 ? __ldcg(lst) : make_uint4(0u, 0u, 0u, 0u);
        for (int k0 = 0; k0 < n; k0 += 8) {
          lst += c.d.dhc_stride;
          const uint4 nxt = k0 + 8 < n ? __ldcg(lst) : make_uint4(0u, 0u, 0u, 0u);
          const uint32_t pk[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half == 1 && k0 + 4 >= n) break;
#pragma unroll
            for (int e = 4 * half; e < 4 * half + 4; ++e) {
              const uint32_t off = (e & 1) ? (pk[e >> 1] >> 16) : (pk[e >> 1] & 0xFFFFu);
              const float4 pj4 = lds_f4(pos_sa + off);
              V3 dd;
              const float r2 = dist2<WRAP>(c, pi4, pj4, &dd);
              bool ok = r2 < rc2;
              if (WRAP) ok = ok && (k0 + e < n);
              float fr = dh_force_unit_q(r2, nk, pj4.w);
              fr = ok ? fr : 0.0f;
              fd.x = fmaf(fr, dd.x, fd.x); fd.y = fmaf(fr, dd.y, fd.y); fd.z = fmaf(fr, dd.z, fd.z);
            }
          }
          cur = nxt;
        }
        const float qi = p.dh_pref * pi4.w;
        fd.x *= qi; fd.y *= qi; fd.z *= qi;
      } else if (pi4.w != 0.0f) {
        const float qi = p.dh_pref * pi4.w;
        const uint32_t sel = sel_mask(G, g);
        const int w_end = c.s.wrange[2 * ci + 1];
        for (int w = c.s.wrange[2 * ci]; w < w_end; ++w) {
          uint32_t m = c.s.dbits[(size_t)ci * c.d.dw + w] & c.s.qmask[w] & sel;   // listed AND charged AND mine
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const float4 pj4 = c.s.pos[c.s.chg[w * 32 + b]];
            V3 dd;
            const float r2 = dist2<WRAP>(c, pi4, pj4, &dd);
            if (r2 < c.rc2) {
              float fr;
              dh_energy_unit(r2, c.kappa, c.neg_kappa_log2e, &fr);
              fr *= qi * pj4.w;
              fd.x = fmaf(fr, dd.x, fd.x); fd.y = fmaf(fr, dd.y, fd.y); fd.z = fmaf(fr, dd.z, fd.z);
            }
          }
        }
      }
      float *fdg = c.s.fd + (size_t)g * 3 * n_loc;
      fdg[cl] = fd.x; fdg[n_loc + cl] = fd.y; fdg[2 * n_loc + cl] = fd.z;
    }
  }
  // This thread has read the positions for the last time in this step.  Every value it loaded has been consumed by the
  // stores above, so the arrive needs no release fence (measured: the fence costs 1.2 % of the step).
  if (arrive_between) cl_arrive_relaxed();
  __syncthreads();
  if (t.i >= 0) {
    const int i = t.i, li = i - c.lo;
    float fx = f.x, fy = f.y, fz = f.z;
    if (t.l >= 0) {
      const float *sa = c.s.slot, *sb = c.s.slot + 3 * nt, *sc = c.s.slot + 6 * nt;
      fx += sa[li] + sb[li] + sc[li];
      fy += sa[nt + li] + sb[nt + li] + sc[nt + li];
      fz += sa[2 * nt + li] + sb[2 * nt + li] + sc[2 * nt + li];
    }
    if (p.use_dh) {
      const int ci = c.s.cidx[i];
      if (ci != 0xFFFF && c.s.pos[i].w != 0.0f) {   // a particle without charge feels no Debye-Hueckel force: its slots hold zeros
        const int cl = ci - c.clo, n_loc = c.d.n_loc;
        for (int g = 0; g < c.d.G; ++g) {
          const float *fdg = c.s.fd + (size_t)g * 3 * n_loc;
          fx += fdg[cl]; fy += fdg[n_loc + cl]; fz += fdg[2 * n_loc + cl];
        }
      }
    }
    if (thermostat && c.s.type[i] < p.T) {
      const uint4 o = philox4x32_10((uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)i, kTagLangevin,
                                    p.thermostat_seed, (uint32_t)(p.replica_offset + c.r));
      fx += fmaf(-c.gamma, t.v[0], c.noise_pref * centred_u23(o.x));
      fy += fmaf(-c.gamma, t.v[1], c.noise_pref * centred_u23(o.y));
      fz += fmaf(-c.gamma, t.v[2], c.noise_pref * centred_u23(o.z));
    }
    if (p.force_cap > 0.0f) {  // system.force_cap (E.py:32,48,51)
      const float n2 = fx * fx + fy * fy + fz * fz;
      if (n2 > p.force_cap * p.force_cap) {
        const float sc_ = p.force_cap * rsqrtf(n2);
        fx *= sc_; fy *= sc_; fz *= sc_;
      }
    }
    t.f[0] = fx; t.f[1] = fy; t.f[2] = fz;
  }
}

__device__ __forceinline__ void compute_forces(Ctx &c, Thr &t, bool thermostat, unsigned long long counter, bool arrive_between) {
  if (c.wrap) compute_forces_impl<true>(c, t, thermostat, counter, arrive_between);
  else compute_forces_impl<false>(c, t, thermostat, counter, arrive_between);
}

// ---- velocity Verlet + Langevin (R2) ------------------------------------------------------------
// Cluster protocol of one step (every thread alternates arrive / wait):
//   drift in registers
//   wait  B   the peer has finished reading the positions of the previous step
//   store the new position into both copies (+ the rebuild vote into both flag words)
//   arrive + wait  A   all positions of this step are visible in both CTAs
//   [rebuild lists]  forces phase 1 (reads positions)   arrive B   __syncthreads   phase 2, half kick
// Entered and left with no arrival pending.
__device__ void md_steps(Ctx &c, Thr &t, int n_steps, unsigned long long &counter, bool &f_valid, bool &lists_valid) {
  const DevParams &p = c.p;
  if (n_steps <= 0) return;
  if (t.i >= 0) t.q = c.s.pos[t.i].w;   // trial moves may have changed the charge
  if (!lists_valid) { build_lists(c, t); lists_valid = true; }
  build_dh_compact(c);
  if (!f_valid) { compute_forces(c, t, true, counter, false); f_valid = true; }
  cl_arrive();   // B of "step -1": everything this thread read or mirrored before the burst is done
  for (int step = 0; step < n_steps; ++step) {
    const int par = step & 1;
    bool moved = false;
    if (t.i >= 0) {
      float d2 = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        t.v[a] = fmaf(p.half_dt, t.f[a], t.v[a]);
        t.x[a] = fmaf(p.dt, t.v[a], t.x[a]);
        const float dd = t.x[a] - t.ref[a];
        d2 = fmaf(dd, dd, d2);
      }
      moved = !(d2 <= p.half_skin2);  // also true for NaN
    }
    cl_wait();
    if (t.i >= 0) {
      const float4 np = make_float4(t.x[0], t.x[1], t.x[2], t.q);
      c.s.pos[t.i] = np;
      c.q.pos[t.i] = np;
      if (moved) { c.s.flag[par] = 1; c.q.flag[par] = 1; }
    }
    if (c.tid == 0) c.s.flag[par ^ 1] = 0;   // next step's vote: last read two barriers ago, next written after our arrive B
    cl_arrive();
    cl_wait();
    if (c.s.flag[par]) { build_lists(c, t); build_dh_compact(c); }
    counter += 1;
    compute_forces(c, t, true, counter, true);
#pragma unroll
    for (int a = 0; a < 3; ++a) t.v[a] = fmaf(p.half_dt, t.f[a], t.v[a]);
  }
  cl_wait();
  if (c.tid == 0) { c.s.flag[0] = 0; c.s.flag[1] = 0; }
  c.compact = false;   // trial moves change the charges
  __syncthreads();
}

// ---- constant-pH trial moves (R1) ----------------------------------------------------------------
// As in chain_engine.cu: positions rest during a reaction block, so every chargeable particle gets
//   phi[ci]    = sum_j q_j exp(-kappa r_ij)/r_ij over the listed, charged partners inside r_cut
//   wcache[ci] = sum_j WCA(t_deprot, t_j, r_ij) - WCA(t_prot, t_j, r_ij) over chain neighbours + row
// from the CTA that owns it (CTA 1 writes its half into the leader), and a trial on the leader's warp 0 is a
// handful of dependent loads.

// WCA type-swap energy of chargeable particle ci.  The row of its particle may live in the peer.
__device__ float wca_swap_energy(const Ctx &c, int ci, bool *titratable_partner = nullptr) {
  const DevParams &p = c.p;
  const int m = c.s.rxof[ci];
  if (titratable_partner) *titratable_partner = false;
  if (m == 255 || !p.use_wca) return 0.f;
  const int S = p.T + 1, i = c.s.chg[ci];
  const int tp = p.rx[m].type_prot, td = p.rx[m].type_deprot;
  const float4 pi4 = c.s.pos[i];
  const int l = i < p.NC ? (p.NC == p.N ? i : i % p.N) : -1;
  const bool mine = i >= c.lo && i < c.hi;
  const int li = mine ? i - c.lo : i - (c.rank == 0 ? c.d.half : 0);
  const unsigned char *wn = mine ? c.s.wn : c.q.wn;
  const uint16_t *wl = mine ? c.s.wl : c.q.wl;
  const int n_w = wn[li];
  float e = 0.f;
  for (int k = 0; k < n_w + 4; ++k) {
    int j = -1;
    if (k < 4) {
      const int off = k < 2 ? k - 2 : k - 1;   // -2,-1,+1,+2: the chain neighbours are not in the rows
      if (l >= 0 && l + off >= 0 && l + off < p.N) j = i + off;
    } else {
      j = wl[(size_t)(k - 4) * c.nt + li];
    }
    if (j < 0) continue;
    const int tj = c.s.type[j];
    if (tj >= p.T) continue;
    V3 dd;
    const float r2 = dist2<true>(c, pi4, c.s.pos[j], &dd);
    if (titratable_partner && r2 < p.rc_w_max2 && c.s.cidx[j] != 0xFFFF && c.s.rxof[c.s.cidx[j]] != 255)
      *titratable_partner = true;
    e += wca_energy(c.s.tab[td * S + tj], r2) - wca_energy(c.s.tab[tp * S + tj], r2);
  }
  return e;
}

// Both CTAs; needs positions, types, rows and qmask current; the caller puts a cluster barrier behind it.
__device__ void mc_prepare(Ctx &c) {
  const DevParams &p = c.p;
  const int G = c.d.G, n_mine = c.chi - c.clo, n_loc = c.d.n_loc;
  if (p.use_dh) {
    for (int idx = c.tid; idx < G * n_mine; idx += c.nt) {
      const int g = idx / n_mine, cl = idx - g * n_mine, ci = c.clo + cl;
      const uint32_t sel = sel_mask(G, g);
      const float4 pi4 = c.s.pos[c.s.chg[ci]];
      float sum = 0.f;
      const int w_end = c.s.wrange[2 * ci + 1];
      for (int w = c.s.wrange[2 * ci]; w < w_end; ++w) {
        uint32_t m = c.s.dbits[(size_t)ci * c.d.dw + w] & c.s.qmask[w] & sel;
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const float4 pj4 = c.s.pos[c.s.chg[w * 32 + b]];
          V3 dd;
          const float r2 = c.wrap ? dist2<true>(c, pi4, pj4, &dd) : dist2<false>(c, pi4, pj4, &dd);
          if (r2 < c.rc2) {
            float fr;
            sum = fmaf(pj4.w, dh_energy_unit(r2, c.kappa, c.neg_kappa_log2e, &fr), sum);
          }
        }
      }
      c.s.fd[(size_t)g * 3 * n_loc + cl] = sum;   // the x plane of the force slots doubles as scratch
    }
  }
  // reaction of every chargeable particle (both CTAs, all of them: wca_swap_energy looks at partners' entries)
  for (int ci = c.tid; ci < p.n_chg; ci += c.nt) {
    const int t = c.s.type[c.s.chg[ci]];
    int m = 255;
    for (int k = 0; k < p.n_rx; ++k)
      if (t == p.rx[k].type_prot || t == p.rx[k].type_deprot) m = k;
    c.s.rxof[ci] = (unsigned char)m;
  }
  if (c.rank == 0)
    for (int m = 0; m < p.n_rx; ++m) {   // N/(N_sites - N + 1): the N_react/(N_prod+1) factor without a division per trial
      const int n_sites = c.s.cnt[m * 2] + c.s.cnt[m * 2 + 1];
      for (int n = c.tid; n <= n_sites; n += c.nt)
        c.s.ratio[(size_t)m * (p.n_chg + 1) + n] = (float)n / (float)(n_sites - n + 1);
    }
  __syncthreads();
  for (int ci = c.clo + c.tid; ci < c.chi; ci += c.nt) {
    const int cl = ci - c.clo;
    float ph = 0.f;
    if (p.use_dh)
      for (int g = 0; g < G; ++g) ph += c.s.fd[(size_t)g * 3 * n_loc + cl];
    bool flag;
    const float wc = wca_swap_energy(c, ci, &flag);
    c.s.phi[ci] = ph; c.s.wcache[ci] = wc; c.s.wflag[ci] = flag ? 1 : 0;
    if (c.rank != 0) { c.q.phi[ci] = ph; c.q.wcache[ci] = wc; c.q.wflag[ci] = flag ? 1 : 0; }
  }
  __syncthreads();
}

__device__ __forceinline__ void mc_bar(int n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }

__device__ __forceinline__ void mc_phi_word(Ctx &c, int ci, int w, float dq, float4 pi4) {
  const uint32_t mw = c.s.dbits[(size_t)ci * c.d.dw + w];
  if ((mw >> c.lane) & 1u) {
    const int ck = w * 32 + c.lane;
    V3 dd;
    const float4 pk4 = c.s.pos[c.s.chg[ck]];
    const float r2 = c.wrap ? dist2<true>(c, pi4, pk4, &dd) : dist2<false>(c, pi4, pk4, &dd);
    if (r2 < c.rc2) {
      float fr;
      c.s.phi[ck] = fmaf(dq, dh_energy_unit(r2, c.kappa, c.neg_kappa_log2e, &fr), c.s.phi[ck]);
    }
  }
}

// Leader CTA only (all of its threads).  Ends with a block barrier.
__device__ void mc_block(Ctx &c, int m, int n_trials, unsigned long long trial_ctr, long long &accepted,
                         cgpe_trial_record *trace) {
  const DevParams &p = c.p;
  const bool helpers = p.use_dh && c.nt > 32;   // same for every thread
  float *ord = c.s.ord;   // {x, y, z, -, dq, ci, op}
  if (c.warp != 0 && n_trials > 0 && helpers) {
    const int nw = (p.n_chg + 31) >> 5, n_help = (c.nt >> 5) - 1;
    for (;;) {
      mc_bar(c.nt);   // an order is posted
      if (__float_as_int(ord[6]) == 0) break;
      const int ci = __float_as_int(ord[5]);
      const float dq = ord[4];
      const float4 pi4 = make_float4(ord[0], ord[1], ord[2], 0.f);
      for (int w = c.warp - 1; w < nw; w += n_help) mc_phi_word(c, ci, w, dq, pi4);
      mc_bar(c.nt);   // the potentials are current
    }
  }
  if (c.warp == 0 && n_trials > 0) {
    const DevReaction rx = p.rx[m];
    uint16_t *lp = c.s.lp + (size_t)m * p.n_chg, *ld = c.s.ld + (size_t)m * p.n_chg;
    int n_p = c.s.cnt[m * 2], n_d = c.s.cnt[m * 2 + 1];
    const float arg_fwd = (float)(2.302585092994045684 * (c.g.pH[c.r] - rx.pK)) * kLog2e;
    const float neg_beta_log2e = -rx.inv_kT * kLog2e;
    const float dq_fwd = rx.q_deprot - rx.q_prot, coul_fwd = p.dh_pref * dq_fwd;
    const uint32_t k0 = rx.seed, k1 = (uint32_t)(p.replica_offset + c.r);
    const int nw = (p.n_chg + 31) >> 5;
    const float *ratio = c.s.ratio + (size_t)m * (p.n_chg + 1);
    long long acc = 0;
    for (int base = 0; base < n_trials; base += 32) {
      const unsigned long long cnt = trial_ctr + (unsigned long long)(base + c.lane);
      const uint4 o = philox4x32_10((uint32_t)cnt, (uint32_t)(cnt >> 32), 0u, kTagMc + (uint32_t)m, k0, k1);
      const int nb = min(32, n_trials - base);
      for (int tt = 0; tt < nb; ++tt) {
        const uint32_t a0 = __shfl_sync(0xffffffffu, o.x, tt), a1 = __shfl_sync(0xffffffffu, o.y, tt),
                       a2 = __shfl_sync(0xffffffffu, o.z, tt);
        const bool fwd = (a0 >> 31) == 0;
        const int n_src = fwd ? n_p : n_d, n_dst = fwd ? n_d : n_p;
        uint16_t *src = fwd ? lp : ld, *dst = fwd ? ld : lp;
        const float u = u24(a2);
        if (n_src <= 0) {   // no reactant: counted as tried, rejected
          if (trace && c.lane == 0) {
            cgpe_trial_record rec = {-1, -1, (int8_t)(fwd ? 1 : -1), 0, 0, 0, 0.f, 0.f, u};
            trace[base + tt] = rec;
          }
          continue;
        }
        const int k = (int)__umulhi(a1, (uint32_t)n_src);
        const int i = src[k];
        const int ci = c.s.cidx[i];
        const int t_new = fwd ? rx.type_deprot : rx.type_prot;
        const float q_new = fwd ? rx.q_deprot : rx.q_prot, dq = fwd ? dq_fwd : -dq_fwd;
        // E_new - E_old of the type/charge swap of site i from the per-site caches
        float de = c.s.wcache[ci];
        if (p.use_dh) de = fmaf(coul_fwd, c.s.phi[ci], de);
        if (!fwd) de = -de;
        // bf = N_react/(N_prod+1) exp(-dE/kT + nu ln10 (pH-pK))   
        const float bf = ratio[n_src] * exp2f(fmaf(neg_beta_log2e, de, fwd ? arg_fwd : -arg_fwd));
        const bool accept = u < bf;
        if (trace && c.lane == 0) {
          cgpe_trial_record rec = {i, -1, (int8_t)(fwd ? 1 : -1), (int8_t)accept, 0, 0, de, bf, u};
          trace[base + tt] = rec;
        }
        if (accept) {
          const float4 pi4 = c.s.pos[i];
          // the flipped site's row: lane w fetches word w (before the bookkeeping stores)
          const uint32_t row = (p.use_dh && c.lane < nw) ? c.s.dbits[(size_t)ci * c.d.dw + c.lane] : 0u;
          const bool refresh = p.use_wca && c.s.wflag[ci] != 0;
          if (c.lane == 0) {
            uint32_t qm = c.s.qmask[ci >> 5];
            if (q_new != 0.f) qm |= 1u << (ci & 31); else qm &= ~(1u << (ci & 31));
            c.s.type[i] = (unsigned char)t_new; c.s.pos[i].w = q_new; c.s.qmask[ci >> 5] = qm;
            c.q.type[i] = (unsigned char)t_new; c.q.pos[i].w = q_new; c.q.qmask[ci >> 5] = qm;   // the peer's copy
            src[k] = src[n_src - 1];
            dst[n_dst] = (uint16_t)i;
          }
          if (fwd) { n_p -= 1; n_d += 1; } else { n_d -= 1; n_p += 1; }
          acc += 1;
          if (p.use_dh && dq != 0.f) {
            // phi of every listed partner inside r_cut changes by dq e(r): lane b handles bit b of each non-empty
            // word, the helper warps one word each when the row is long
            uint32_t live = __ballot_sync(0xffffffffu, row != 0u);
            if (helpers && __popc(live) > (c.nt > 256 ? 2 : 1)) {
              if (c.lane == 0) {
                ord[0] = pi4.x; ord[1] = pi4.y; ord[2] = pi4.z; ord[4] = dq;
                ord[5] = __int_as_float(ci); ord[6] = __int_as_float(1);
              }
              mc_bar(c.nt);
              mc_bar(c.nt);
            } else {
              while (live) {
                const int w = __ffs(live) - 1;
                live &= live - 1;
                mc_phi_word(c, ci, w, dq, pi4);
              }
            }
          }
          __syncwarp();
          if (refresh) {
            // partners of i that are titratable themselves see a new neighbour type: refresh their cache
            const int l = i < p.NC ? (p.NC == p.N ? i : i % p.N) : -1;
            const bool mine = i >= c.lo && i < c.hi;
            const int li = mine ? i - c.lo : i - c.d.half;
            const int n_w = mine ? c.s.wn[li] : c.q.wn[li];
            const uint16_t *wl = mine ? c.s.wl : c.q.wl;
            for (int e = c.lane; e < n_w + 4; e += 32) {
              int j = -1;
              if (e < 4) {
                const int off = e < 2 ? e - 2 : e - 1;
                if (l >= 0 && l + off >= 0 && l + off < p.N) j = i + off;
              } else {
                j = wl[(size_t)(e - 4) * c.nt + li];
              }
              if (j >= 0) {
                const int cj = c.s.cidx[j];
                if (cj != 0xFFFF && c.s.rxof[cj] != 255) c.s.wcache[cj] = wca_swap_energy(c, cj);
              }
            }
            __syncwarp();
          }
        }
      }
    }
    if (helpers) {   // release the helpers
      if (c.lane == 0) ord[6] = __int_as_float(0);
      mc_bar(c.nt);
    }
    if (c.lane == 0) { c.s.cnt[m * 2] = n_p; c.s.cnt[m * 2 + 1] = n_d; }
    accepted += acc;  // only meaningful on warp 0; thread 0 writes it back
  }
  __syncthreads();
}

// ---- observables (R5, R6): leader CTA, from its copy of all positions ------------------------------
__device__ void chain_observables(Ctx &c, double *rg, double *ree) {
  const DevParams &p = c.p;
  const float4 p0 = c.s.pos[0];
  double sx = 0, sy = 0, sz = 0, s2 = 0;
  for (int i = c.tid; i < p.N; i += c.nt) {
    const float4 x = c.s.pos[i];
    const double dx = (double)x.x - p0.x, dy = (double)x.y - p0.y, dz = (double)x.z - p0.z;
    sx += dx; sy += dy; sz += dz; s2 += dx * dx + dy * dy + dz * dz;
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); s2 = warp_sum(s2);
  __syncthreads();
  if (c.lane == 0) { c.s.red[c.warp * 4] = sx; c.s.red[c.warp * 4 + 1] = sy; c.s.red[c.warp * 4 + 2] = sz; c.s.red[c.warp * 4 + 3] = s2; }
  __syncthreads();
  const int nw = (c.nt + 31) / 32;
  double tx = 0, ty = 0, tz = 0, t2 = 0;
  for (int w = 0; w < nw; ++w) { tx += c.s.red[w * 4]; ty += c.s.red[w * 4 + 1]; tz += c.s.red[w * 4 + 2]; t2 += c.s.red[w * 4 + 3]; }
  const double inv = 1.0 / p.N;
  const double v = t2 * inv - (tx * tx + ty * ty + tz * tz) * inv * inv;
  *rg = sqrt(v > 0.0 ? v : 0.0);
  const float4 pe = c.s.pos[p.N - 1];
  const double ex = (double)pe.x - p0.x, ey = (double)pe.y - p0.y, ez = (double)pe.z - p0.z;
  *ree = sqrt(ex * ex + ey * ey + ez * ez);
}

// common prologue: replica choice, shared-memory binding (own + peer), context
__device__ __forceinline__ void prologue(Ctx &c) {
  cg::cluster_group cluster = cg::this_cluster();
  c.rank = (int)cluster.block_rank();
  bind(c.d, duo_smem, &c.s);
  bind(c.d, static_cast<unsigned char *>(cluster.map_shared_rank(static_cast<void *>(duo_smem), c.rank ^ 1)), &c.q);
  c.r = c.d.order[blockIdx.x >> 1];
  init_ctx(c);
}

// ---- kernels -----------------------------------------------------------------------------------
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_duo_md_run(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ DuoShape d, int n_steps) {
  Ctx c{p, g, d};
  prologue(c);
  Thr t;
  bool f_valid = g.f_valid[c.r] != 0, lists_valid = false;
  unsigned long long counter = (unsigned long long)g.md_counter[c.r];
  load_state(c, t, f_valid);
  cl_sync();
  md_steps(c, t, n_steps, counter, f_valid, lists_valid);
  cl_sync();
  store_state(c, t, true);
  if (c.rank == 0 && c.tid == 0) {
    g.md_counter[c.r] = (long long)counter;
    g.md_steps[c.r] += n_steps;
    g.time[c.r] += (double)n_steps * p.dt_d;
    g.f_valid[c.r] = f_valid ? 1 : 0;
  }
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_duo_mc_run(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ DuoShape d, int m,
             int n_trials, cgpe_trial_record *trace) {
  Ctx c{p, g, d};
  prologue(c);
  Thr t;
  load_state(c, t, true);
  cl_sync();
  build_lists(c, t);
  cl_sync();
  mc_prepare(c);
  cl_sync();
  const unsigned long long trial_ctr = (unsigned long long)g.trials[(size_t)c.r * p.n_rx + m];
  long long accepted = 0;
  if (c.rank == 0) mc_block(c, m, n_trials, trial_ctr, accepted, trace ? trace + (size_t)c.r * n_trials : nullptr);
  cl_sync();
  store_state(c, t, true);
  if (c.rank == 0 && c.tid == 0) {
    g.trials[(size_t)c.r * p.n_rx + m] = (long long)(trial_ctr + (unsigned long long)n_trials);
    g.accepted[(size_t)c.r * p.n_rx + m] += accepted;
    if (n_trials > 0) g.f_valid[c.r] = 0;
  }
}

// forces exactly as the MD loop computes them (compact partner lists included), no thermostat
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_duo_md_forces(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ DuoShape d, float4 *out) {
  Ctx c{p, g, d};
  prologue(c);
  Thr t;
  load_state(c, t, false);
  cl_sync();
  build_lists(c, t);
  build_dh_compact(c);
  compute_forces(c, t, false, 0ull, false);
  if (t.i >= 0) out[(size_t)c.r * p.P + t.i] = make_float4(t.f[0], t.f[1], t.f[2], 0.f);
  if (c.err) atomicOr(&c.g.err[c.r], c.err);
  cl_sync();   // CTA 1 mirrored its rows into the leader: nobody leaves before those stores have landed
}

// The sampling loop of perform_sampling (E.py:88-106), one launch for all samples.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_duo_sample_loop(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ DuoShape d,
                  const __grid_constant__ cgpe_sample_params sp, uint32_t thr1, uint32_t thr2, double *obs, float *qdist) {
  if (g.engine_choice && *g.engine_choice != 1) return;   // the plan of this launch gave it to the one-CTA engine (whole clusters leave)
  Ctx c{p, g, d};
  prologue(c);
  Thr t;
  if (c.tid == 0) g.cta_ns[c.r * 4 + c.rank * 2] = (long long)global_timer_ns();
  bool f_valid = g.f_valid[c.r] != 0, lists_valid = false;
  unsigned long long counter = (unsigned long long)g.md_counter[c.r];
  unsigned long long sample_ctr = (unsigned long long)g.samples_done[c.r];
  unsigned long long trial_ctr[kMaxRx];
  long long accepted[kMaxRx];
  for (int m = 0; m < kMaxRx; ++m) {
    trial_ctr[m] = m < p.n_rx ? (unsigned long long)g.trials[(size_t)c.r * p.n_rx + m] : 0ull;
    accepted[m] = 0;
  }
  load_state(c, t, f_valid);
  cl_sync();
  const double t0 = g.time[c.r];
  long long steps = 0;
  const uint32_t k0 = (uint32_t)sp.schedule_seed, k1 = (uint32_t)(p.replica_offset + c.r);
  for (int s = 0; s < sp.n_samples; ++s) {
    // every thread of both CTAs draws the same schedule numbers
    const uint4 o = philox4x32_10((uint32_t)sample_ctr, (uint32_t)(sample_ctr >> 32), 0u, kTagSchedule, k0, k1);
    sample_ctr += 1;
    int n = 0;
    if (sp.use_md)
      n = ((o.x >> 8) < thr1 ? sp.md_steps_per_burst : 0) +   // E.py:89-90
          ((o.y >> 8) < thr2 ? sp.md_steps_per_burst : 0);    // E.py:91-92
    if (n > 0) {
      md_steps(c, t, n, counter, f_valid, lists_valid);
      steps += n;
    } else if (!lists_valid) {
      build_lists(c, t); lists_valid = true;
      cl_sync();   // rows mirrored into the leader
    }
    if (sp.n_mc_blocks > 0) {
      mc_prepare(c);      // per-site caches for this sample's trial moves
      cl_sync();
      for (int b = 0; b < sp.n_mc_blocks; ++b) {   // E.py:94-95
        const int m = sp.mc_reaction[b];
        if (c.rank == 0) mc_block(c, m, sp.mc_trials[b], trial_ctr[m], accepted[m], nullptr);
        trial_ctr[m] += (unsigned long long)sp.mc_trials[b];
        if (sp.mc_trials[b] > 0) f_valid = false;
      }
    }
    if (c.rank == 0) {
      double rg, ree;
      chain_observables(c, &rg, &ree);   // E.py:100-101
      if (c.tid == 0) {
        double *o6 = obs + ((size_t)c.r * sp.n_samples + s) * CGPE_N_OBS;
        int n_ion = 0;
        for (int m = 0; m < p.n_rx; ++m) n_ion += c.s.cnt[m * 2 + 1];
        o6[0] = p.n_rx > 0 ? c.s.cnt[1] : 0;   // E.py:97
        o6[1] = p.n_rx > 1 ? c.s.cnt[3] : 0;   // E.py:98
        o6[2] = n_ion;                         // E.py:99 (implicit ions: N_B = N_N + N_N2)
        o6[3] = rg; o6[4] = ree;
        o6[5] = t0 + (double)steps * p.dt_d;   // E.py:102
      }
      if (qdist && sp.q_snapshot_every > 0 && s % sp.q_snapshot_every == 0) {   // E.py:103-106
        const int row = s / sp.q_snapshot_every;
        if (row < sp.q_snapshot_rows)
          for (int i = c.tid; i < p.N; i += c.nt)
            qdist[((size_t)c.r * sp.q_snapshot_rows + row) * p.N + i] = c.s.pos[i].w;
      }
    }
    cl_sync();   // flips mirrored into the peer; the leader is done with its copy of the positions
  }
  store_state(c, t, true);
  if (c.tid == 0) g.cta_ns[c.r * 4 + c.rank * 2 + 1] = (long long)global_timer_ns();
  if (c.rank == 0 && c.tid == 0) {
    g.md_counter[c.r] = (long long)counter;
    g.md_steps[c.r] += steps;
    g.samples_done[c.r] = (long long)sample_ctr;
    g.time[c.r] = t0 + (double)steps * p.dt_d;
    g.f_valid[c.r] = f_valid ? 1 : 0;
    for (int m = 0; m < p.n_rx; ++m) {
      g.trials[(size_t)c.r * p.n_rx + m] = (long long)trial_ctr[m];
      g.accepted[(size_t)c.r * p.n_rx + m] += accepted[m];
    }
  }
}

template <typename K, typename... Args>
cudaError_t launch_cluster(K kernel, const DevParams &p, const DevState &g, const DuoShape &d, cudaStream_t st, const PlanJob &job,
                           Args... args) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, d.total);
  if (e != cudaSuccess) return e;
  if (p.use_dh && p.R > d.n_sm / 2 && p.R <= d.n_sm) k_duo_count_pairs<<<p.R, 256, sizeof(float4) * (p.n_chg > 0 ? p.n_chg : 1), st>>>(p, g);
  k_duo_plan<<<1, 256, 3 * sizeof(float) * p.R, st>>>(p, g, d.n_sm, d.order, job);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (std::getenv("CGPE_PLAN_DEBUG")) {   // diagnostics: the plan of this launch on stderr
    std::vector<int> ord(p.R);
    std::vector<double> pl((size_t)p.R * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(ord.data(), d.order, sizeof(int) * p.R, cudaMemcpyDeviceToHost);
    cudaMemcpy(pl.data(), g.plan, sizeof(double) * 8 * p.R, cudaMemcpyDeviceToHost);
    std::fprintf(stderr, "[cgpe plan] samples %d fixed_steps %d\n", job.n_samples, job.fixed_steps);
    for (int c = 0; c < p.R; ++c) {
      const int r = ord[c];
      std::fprintf(stderr, "[cgpe plan] cluster %3d replica %3d cost %.0f steps %.0f pairs %.0f acc %.2f\n", c, r, pl[r * 8 + 6], pl[r * 8 + 7],
                   pl[r * 8 + 4], pl[r * 8 + 5]);
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * p.R);
  cfg.blockDim = dim3(d.nt);
  cfg.dynamicSmemBytes = (size_t)d.total;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace

// Order of the replicas for a fused sampling launch of the one-CTA engine that runs in waves (more replicas than SMs):
// predicted cost as in k_duo_plan, most expensive first, so that the last wave holds the cheap ones (longest processing
// time first).  order: device array of R ints.  Two small launches.
cudaError_t duo_plan_waves(const DevParams &p, const DevState &g, int *order, const cgpe_sample_params &sp, uint32_t thr1,
                           uint32_t thr2, cudaStream_t st) {
  int trials = 0;
  for (int b = 0; b < sp.n_mc_blocks; ++b) trials += sp.mc_trials[b];
  const PlanJob job{sp.n_samples, sp.md_steps_per_burst, sp.use_md, trials, thr1, thr2, (uint32_t)sp.schedule_seed, 0};
  if (p.use_dh) k_duo_count_pairs<<<p.R, 256, sizeof(float4) * (p.n_chg > 0 ? p.n_chg : 1), st>>>(p, g);
  k_duo_plan<<<1, 256, 3 * sizeof(float) * p.R, st>>>(p, g, 0, order, job);   // n_sm = 0: every launch is "more clusters than the device holds"
  return cudaGetLastError();
}

cudaError_t duo_md_run(const DevParams &p, const DevState &g, const DuoShape &d, int n_steps, cudaStream_t st) {
  return d.nt <= 512 ? launch_cluster(k_duo_md_run<512, 2>, p, g, d, st, PlanJob{0, 0, 0, 0, 0u, 0u, 0u, n_steps}, p, g, d, n_steps)
                     : launch_cluster(k_duo_md_run<1024, 1>, p, g, d, st, PlanJob{0, 0, 0, 0, 0u, 0u, 0u, n_steps}, p, g, d, n_steps);
}

cudaError_t duo_mc_run(const DevParams &p, const DevState &g, const DuoShape &d, int m, int n_trials,
                       cgpe_trial_record *trace, cudaStream_t st) {
  return d.nt <= 512 ? launch_cluster(k_duo_mc_run<512, 2>, p, g, d, st, PlanJob{1, 0, 0, n_trials, 0u, 0u, 0u, 0}, p, g, d, m, n_trials, trace)
                     : launch_cluster(k_duo_mc_run<1024, 1>, p, g, d, st, PlanJob{1, 0, 0, n_trials, 0u, 0u, 0u, 0}, p, g, d, m, n_trials, trace);
}

cudaError_t duo_md_forces(const DevParams &p, const DevState &g, const DuoShape &d, float4 *out, cudaStream_t st) {
  return d.nt <= 512 ? launch_cluster(k_duo_md_forces<512, 2>, p, g, d, st, PlanJob{0, 0, 0, 0, 0u, 0u, 0u, 1}, p, g, d, out)
                     : launch_cluster(k_duo_md_forces<1024, 1>, p, g, d, st, PlanJob{0, 0, 0, 0, 0u, 0u, 0u, 1}, p, g, d, out);
}

cudaError_t duo_sample_loop(const DevParams &p, const DevState &g, const DuoShape &d, const cgpe_sample_params &sp,
                            uint32_t thr1, uint32_t thr2, double *obs, float *qdist, cudaStream_t st) {
  int trials = 0;
  for (int b = 0; b < sp.n_mc_blocks; ++b) trials += sp.mc_trials[b];
  const PlanJob job{sp.n_samples, sp.md_steps_per_burst, sp.use_md, trials, thr1, thr2, (uint32_t)sp.schedule_seed, 0};
  return d.nt <= 512 ? launch_cluster(k_duo_sample_loop<512, 2>, p, g, d, st, job, p, g, d, sp, thr1, thr2, obs, qdist)
                     : launch_cluster(k_duo_sample_loop<1024, 1>, p, g, d, st, job, p, g, d, sp, thr1, thr2, obs, qdist);
}

}  // namespace cgpe
