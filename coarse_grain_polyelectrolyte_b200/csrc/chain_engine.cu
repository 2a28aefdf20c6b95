// chain_engine.cu — the hot path: one persistent CTA per replica, whole replica resident in
// shared memory (sm_100a, up to 227 KB per CTA).
//
//   * positions+charge float4 in shared memory, velocity / force / list reference position in
//     registers of the owning thread (thread t owns particles t, t+blockDim, ...);
//   * two Verlet lists in shared memory (uint16 ids): a short one for WCA (all particles, radius
//     max WCA cutoff + skin) and one for Debye-Hueckel over the particles that can carry charge
//     (titratable sites [+ ions], radius r_cut + skin); rebuilt when any particle has moved
//     more than skin/2, detected with one __syncthreads_or per step;
//   * bonded terms are computed once by an owner thread (bond (i-1,i), angle (i-1,i,i+1),
//     dihedral (i-1,i,i+1,i+2) belong to bead i) and handed to the partner beads through three
//     conflict-free shared-memory slots, so a step costs two block barriers;
//   * the constant-pH trial moves run on warp 0: Philox draws for 32 trials at a time, Delta-E as
//     a warp-shuffle reduction over the site's list rows, decision and O(1) list bookkeeping;
//   * the sampling loop of EnsembleDynamics.perform_sampling (E.py:88-106) runs entirely inside
//     one launch: MD bursts, MC blocks, observables, no host round trip per sample.
//
// Reference call sites replaced (relative to /root/reference/src/modules):
//   integrator.run            EnsembleDynamics.py:41,52,58,64,73,90,92
//   reaction.reaction         EnsembleDynamics.py:70-71,94-95
//   calc_rg / calc_re / number_of_particles   EnsembleDynamics.py:97-101
#include "chain_engine.cuh"

#include <algorithm>
#include <cstdlib>

namespace cgpe {

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up (same function on host, for the size, and on device, for the pointers)
// ------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Host: lay the arrays out for the shapes in p and record the offsets in p.off.
size_t chain_smem_bytes(DevParams &p) {
  size_t o = 0;
  const int TT = (p.T + 1) * (p.T + 1);
  const size_t n_chg = p.n_chg > 0 ? p.n_chg : 1, n_rx = p.n_rx > 0 ? p.n_rx : 1, grp = p.dh_group > 0 ? p.dh_group : 1;
  auto take = [&](size_t bytes, size_t align) { o = align_up(o, align); size_t at = o; o += bytes; return (int)at; };
  SmemOff &f = p.off;
  f.pos = take(sizeof(float4) * p.P, 16);
  f.tab = take(sizeof(float4) * TT, 16);
  f.red = take(sizeof(double) * 32 * 4, 16);   // also holds the block spheres of a list build (float4)
  f.slot = take(sizeof(float) * 9 * p.NC, 4);
  f.fd = take(sizeof(float) * 3 * grp * n_chg, 4);
  f.dhn = take(sizeof(uint16_t) * grp * n_chg, 2);
  f.cnt = take(sizeof(int) * kMaxRx * 2, 4);
  f.flag = take(sizeof(int) * 4, 4);
  f.wl = take(sizeof(uint16_t) * (size_t)p.cap_w * p.P, 4);
  f.db = take(sizeof(uint32_t) * (size_t)(p.dw > 0 ? p.dw : 1) * n_chg, 4);
  f.qm = take(sizeof(uint32_t) * (size_t)(p.dw > 0 ? p.dw : 1), 4);
  f.chg = take(sizeof(uint16_t) * n_chg, 2);
  f.cidx = take(sizeof(uint16_t) * p.P, 2);
  f.lp = take(sizeof(uint16_t) * n_rx * n_chg, 2);
  f.ld = take(sizeof(uint16_t) * n_rx * n_chg, 2);
  f.wn = take((size_t)p.P, 1);
  f.type = take((size_t)p.P, 1);
  f.phi = take(sizeof(float) * ((n_chg + 127) / 128 * 128), 16);   // padded to the row stride of DevState::mc_e (float4 updates)
  f.wc = take(sizeof(float) * n_chg, 4);
  f.rxof = take(n_chg, 1);
  f.cpos = take(sizeof(float4) * (n_chg + 1), 16);   // +1: far-away, uncharged dummy that pads the compact partner lists
  f.wrange = take(2 * n_chg, 1);
  f.wflag = take(n_chg, 1);
  f.ratio = take(sizeof(float) * n_rx * (n_chg + 1), 4);
  f.qlist = take(sizeof(uint16_t) * n_chg, 2);
  f.ionA = take(sizeof(uint16_t) * (p.n_ions > 0 ? p.n_ions : 1), 2);
  f.ionF = take(sizeof(uint16_t) * (p.n_ions > 0 ? p.n_ions : 1), 2);
  f.ionV = take(sizeof(float4) * (p.n_ions > 0 ? p.n_ions : 1), 16);
  f.total = (int)align_up(o, 16);
  return (size_t)f.total;
}

namespace {

// per-thread registers: ITEMS particles per thread
template <int ITEMS>
struct Thr {
  float x[ITEMS][3], v[ITEMS][3], f[ITEMS][3], ref[ITEMS][3];
  int l[ITEMS];   // index within its chain, -1 for particles that are not chain beads
  int ci[ITEMS];  // index among the chargeable particles, -1 if the particle can never carry charge
};

struct Ctx {
  const DevParams &p;
  const DevState &g;
  ChainSmem s;
  int r;              // replica (local)
  int tid, nt, lane, warp;
  float kappa, rc2, neg_kappa_log2e, rlist_d2, rlist_max, gamma, noise_pref;
  bool wrap;          // false: every listed pair is closer directly than through any periodic image
  bool compact;       // DevState::dhc holds the partners of the current rows and charges (see build_dh_compact)
  bool dense;         // Debye-Hueckel list radius ~ replica extent: loop over all charged particles instead of the bit rows
  int dh_g0, dh_ci0;  // this thread's first (sub-thread, particle) work item of the packed compact lists (build_dh_compact)
  int err;            // thread-local error bits, merged at exit
  long long rebuilds; // uniform
  int frame_since, frame_n;   // frame recorder: MD steps since the last frame, frames recorded so far (uniform)
};

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

// ---- load / store ------------------------------------------------------------------------------
template <int ITEMS>
__device__ void load_state(Ctx &c, Thr<ITEMS> &t, bool f_valid) {
  const DevParams &p = c.p;
  const size_t base = (size_t)c.r * p.P;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    if (i < p.P) {
      const float4 pq = c.g.pos[base + i], vv = c.g.vel[base + i];
      c.s.pos[i] = pq;
      t.x[k][0] = pq.x; t.x[k][1] = pq.y; t.x[k][2] = pq.z;
      t.v[k][0] = vv.x; t.v[k][1] = vv.y; t.v[k][2] = vv.z;
      float4 ff = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f_valid) ff = c.g.frc[base + i];
      t.f[k][0] = ff.x; t.f[k][1] = ff.y; t.f[k][2] = ff.z;
      c.s.type[i] = c.g.type[base + i];
      c.s.cidx[i] = 0xFFFFu;
      c.s.wn[i] = 0;
    }
    t.l[k] = (i < p.NC) ? (p.NC == p.N ? i : i % p.N) : -1;
  }
  for (int i = c.tid; i < 9 * p.NC; i += c.nt) c.s.slot[i] = 0.0f;
  for (int i = c.tid; i < (p.T + 1) * (p.T + 1); i += c.nt) c.s.tab[i] = c.g.wca_tab[i];
  for (int i = c.tid; i < p.n_chg; i += c.nt) {
    c.s.chg[i] = (uint16_t)c.g.chg[(size_t)c.r * p.n_chg + i];
    for (int gsub = 0; gsub < 3 * p.dh_group; ++gsub) c.s.fd[gsub * p.n_chg + i] = 0.0f;
  }
  for (int m = 0; m < p.n_rx; ++m) {
    const int *gp = c.g.lst_prot + ((size_t)c.r * p.n_rx + m) * p.P;
    const int *gd = c.g.lst_deprot + ((size_t)c.r * p.n_rx + m) * p.P;
    const int np = c.g.cnt[((size_t)c.r * p.n_rx + m) * 2], nd = c.g.cnt[((size_t)c.r * p.n_rx + m) * 2 + 1];
    for (int i = c.tid; i < np; i += c.nt) c.s.lp[(size_t)m * p.n_chg + i] = (uint16_t)gp[i];
    for (int i = c.tid; i < nd; i += c.nt) c.s.ld[(size_t)m * p.n_chg + i] = (uint16_t)gd[i];
    if (c.tid == 0) { c.s.cnt[m * 2] = np; c.s.cnt[m * 2 + 1] = nd; }
  }
  if (p.n_ions > 0) {   // explicit ions: pools of free H+ and of hidden slots; flag[1], flag[2] = their sizes
    const int na = c.g.ion_cnt[c.r * 2], nf = c.g.ion_cnt[c.r * 2 + 1];
    for (int i = c.tid; i < na; i += c.nt) c.s.ion_active[i] = (uint16_t)c.g.ion_active[(size_t)c.r * p.n_ions + i];
    for (int i = c.tid; i < nf; i += c.nt) c.s.ion_free[i] = (uint16_t)c.g.ion_free[(size_t)c.r * p.n_ions + i];
    for (int i = c.tid; i < p.n_ions; i += c.nt) c.s.ion_vel[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c.tid == 0) { c.s.flag[1] = na; c.s.flag[2] = nf; c.s.flag[3] = 0; }
  }
  __syncthreads();
  for (int i = c.tid; i < p.n_chg; i += c.nt) {
    c.s.cidx[c.s.chg[i]] = (uint16_t)i;
    c.s.cpos[i] = c.s.pos[c.s.chg[i]];
    c.s.wrange[2 * i] = 0; c.s.wrange[2 * i + 1] = 0;
  }
  if (c.tid == 0) c.s.cpos[p.n_chg] = make_float4(1e10f, 1e10f, 1e10f, 0.0f);   // list padding (build_dh_compact)
  // live mask of the chargeable particles that carry charge right now (the MC warp keeps it current)
  for (int w = c.tid; w < p.dw; w += c.nt) {
    uint32_t m = 0;
    for (int b = 0; b < 32; ++b) {
      const int cj = w * 32 + b;
      if (cj < p.n_chg && c.s.pos[c.s.chg[cj]].w != 0.0f) m |= 1u << b;
    }
    c.s.qmask[w] = m;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    t.ci[k] = (i < p.P && c.s.cidx[i] != 0xFFFF) ? (int)c.s.cidx[i] : -1;
  }
}

// Explicit ions: a trial move may have placed an ion somewhere else or hidden it; its owner thread
// takes position and (if pending) velocity over from shared memory.  Call behind a barrier.
template <int ITEMS>
__device__ void sync_ions(Ctx &c, Thr<ITEMS> &t) {
  const DevParams &p = c.p;
  if (p.n_ions == 0) return;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    if (i >= p.NC && i < p.P) {
      const float4 x = c.s.pos[i];
      t.x[k][0] = x.x; t.x[k][1] = x.y; t.x[k][2] = x.z;
      const float4 v = c.s.ion_vel[i - p.NC];
      if (v.w != 0.f) {
        t.v[k][0] = v.x; t.v[k][1] = v.y; t.v[k][2] = v.z;
        c.s.ion_vel[i - p.NC].w = 0.f;
      }
    }
  }
}

template <int ITEMS>
__device__ void store_state(Ctx &c, Thr<ITEMS> &t, bool store_forces) {
  const DevParams &p = c.p;
  const size_t base = (size_t)c.r * p.P;
  __syncthreads();
  sync_ions(c, t);
  int bad = 0;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    if (i < p.P) {
      const float q = c.s.pos[i].w;
      c.g.pos[base + i] = make_float4(t.x[k][0], t.x[k][1], t.x[k][2], q);
      c.g.vel[base + i] = make_float4(t.v[k][0], t.v[k][1], t.v[k][2], 0.f);
      if (store_forces) c.g.frc[base + i] = make_float4(t.f[k][0], t.f[k][1], t.f[k][2], 0.f);
      c.g.type[base + i] = c.s.type[i];
      if (!isfinite(t.x[k][0] + t.x[k][1] + t.x[k][2] + t.v[k][0] + t.v[k][1] + t.v[k][2])) bad = 1;
      if (t.v[k][0] * t.v[k][0] + t.v[k][1] * t.v[k][1] + t.v[k][2] * t.v[k][2] > 1.0e6f * fmaxf(c.g.kT[c.r], 1.0f)) bad |= 2;
    }
  }
  if (bad & 1) c.err |= kErrNonFinite;
  if (bad & 2) c.err |= kErrDiverged;
  for (int m = 0; m < p.n_rx; ++m) {
    int *gp = c.g.lst_prot + ((size_t)c.r * p.n_rx + m) * p.P;
    int *gd = c.g.lst_deprot + ((size_t)c.r * p.n_rx + m) * p.P;
    const int np = c.s.cnt[m * 2], nd = c.s.cnt[m * 2 + 1];
    for (int i = c.tid; i < np; i += c.nt) gp[i] = c.s.lp[(size_t)m * p.n_chg + i];
    for (int i = c.tid; i < nd; i += c.nt) gd[i] = c.s.ld[(size_t)m * p.n_chg + i];
    if (c.tid == 0) { c.g.cnt[((size_t)c.r * p.n_rx + m) * 2] = np; c.g.cnt[((size_t)c.r * p.n_rx + m) * 2 + 1] = nd; }
  }
  if (p.n_ions > 0) {
    const int na = c.s.flag[1], nf = c.s.flag[2];
    for (int i = c.tid; i < na; i += c.nt) c.g.ion_active[(size_t)c.r * p.n_ions + i] = c.s.ion_active[i];
    for (int i = c.tid; i < nf; i += c.nt) c.g.ion_free[(size_t)c.r * p.n_ions + i] = c.s.ion_free[i];
    if (c.tid == 0) { c.g.ion_cnt[c.r * 2] = na; c.g.ion_cnt[c.r * 2 + 1] = nf; }
  }
  if (c.err) atomicOr(&c.g.err[c.r], c.err);
  if (c.tid == 0 && c.rebuilds) c.g.rebuilds[c.r] += c.rebuilds;
  if (c.tid == 0 && p.frame_every > 0) { c.g.frame_phase[c.r] = c.frame_since; c.g.n_frames[c.r] = c.frame_n; }
}

// ---- Verlet lists ------------------------------------------------------------------------------
// WCA rows hold every partner inside (max WCA cutoff + skin) EXCEPT the same-chain neighbours
// i+-1, i+-2, which the bonded owner evaluates directly (compute_forces).  The Debye-Hueckel list is
// one bit per (chargeable, chargeable) pair inside r_cut + skin: fixed size, cannot overflow.
// Requires s.pos current and visible.  Contains block barriers (bounding-box reduction that
// decides whether minimum-image arithmetic can be skipped until the next rebuild, and a trailing
// one that publishes the rows: the Debye-Hueckel rows are read by other threads).
template <int ITEMS, bool WRAP>
__device__ void build_rows(Ctx &c, Thr<ITEMS> &t) {
  const DevParams &p = c.p;
  const float4 *sph = reinterpret_cast<const float4 *>(c.s.red);   // block spheres (build_lists)
  const int nb = (p.P + 31) >> 5;
  const float rl_w = sqrtf(p.rlist_w2);
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    if (i < p.P) {
      t.ref[k][0] = t.x[k][0]; t.ref[k][1] = t.x[k][1]; t.ref[k][2] = t.x[k][2];
      if (p.use_wca && (c.s.type[i] >= p.T || ((p.fixed_mask >> c.s.type[i]) & 1u))) {
        c.s.wn[i] = 0;   // hidden slot: not a particle (it gets rows when a trial move revives it); fixed particle: its force is discarded
      } else if (p.use_wca) {
        const float4 pi = c.s.pos[i];
        // same-chain partners handled by the bonded owner (also excludes j == i)
        int ex_lo = i, ex_hi = i;
        if (t.l[k] >= 0) { ex_lo = i - min(2, t.l[k]); ex_hi = i + min(2, p.N - 1 - t.l[k]); }
        int n = 0;
        for (int b = 0; b < nb; ++b) {   // blocks of 32 consecutive particles: sphere first (block_spheres)
          const float4 sb = sph[b];
          V3 d;
          const float reach = sb.w + rl_w;
          if (dist2<WRAP>(c, pi, sb, &d) > reach * reach) continue;
          const int j1 = min(p.P, b * 32 + 32);
          for (int j = b * 32; j < j1; ++j) {
            const float r2 = dist2<WRAP>(c, pi, c.s.pos[j], &d);
            if (r2 < p.rlist_w2 && (j < ex_lo || j > ex_hi)) {
              if (n < p.cap_w) c.s.wl[(size_t)n * p.P + i] = (uint16_t)j;
              ++n;
            }
          }
        }
        if (p.rlist_wf2 > p.rlist_w2)   // fixed particles sit in slots and may be large: their own list radius
          for (int j = p.NC; j < p.P; ++j) {
            if (!((p.fixed_mask >> c.s.type[j]) & 1u) || j == i) continue;
            V3 d;
            const float r2 = dist2<WRAP>(c, pi, c.s.pos[j], &d);
            if (r2 >= p.rlist_w2 && r2 < p.rlist_wf2) {
              if (n < p.cap_w) c.s.wl[(size_t)n * p.P + i] = (uint16_t)j;
              ++n;
            }
          }
        if (n > p.cap_w) { c.err |= kErrWcaCap; n = p.cap_w; }
        c.s.wn[i] = (unsigned char)n;
      }
    }
  }
  if (p.use_dh) {
    const int nw = (p.n_chg + 31) >> 5;
    for (int ci = c.tid; ci < p.n_chg; ci += c.nt) {
      const float4 pi = c.s.cpos[ci];
      int w_lo = nw, w_hi = 0;
      for (int w = 0; w < nw; ++w) {
        uint32_t m = 0;
        const int hi = min(32, p.n_chg - w * 32);
        for (int b = 0; b < hi; ++b) {
          V3 d;
          const float r2 = dist2<WRAP>(c, pi, c.s.cpos[w * 32 + b], &d);
          if (r2 < c.rlist_d2) m |= 1u << b;
        }
        if ((ci >> 5) == w) m &= ~(1u << (ci & 31));   // not its own partner
        c.s.dbits[(size_t)ci * p.dw + w] = m;
        if (m) { w_lo = min(w_lo, w); w_hi = w + 1; }
      }
      c.s.wrange[2 * ci] = (unsigned char)min(w_lo, w_hi);   // backbone neighbours cluster in a few words
      c.s.wrange[2 * ci + 1] = (unsigned char)w_hi;
    }
  }
}

template <int ITEMS>
__device__ void build_lists(Ctx &c, Thr<ITEMS> &t) {
  const DevParams &p = c.p;
  sync_ions(c, t);   // positions a trial move gave to explicit ions (no-op otherwise)
  if (p.P > p.NC) {
    // free particles (explicit ions) wander: fold them back into the box so FP32 keeps its resolution
    bool any = false;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int i = c.tid + k * c.nt;
      if (i >= p.NC && i < p.P) {
#pragma unroll
        for (int a = 0; a < 3; ++a) t.x[k][a] -= p.box_l * floorf(t.x[k][a] * p.inv_box_l);
        float4 *sp = &c.s.pos[i];
        sp->x = t.x[k][0]; sp->y = t.x[k][1]; sp->z = t.x[k][2];
        if (t.ci[k] >= 0) { float4 *cp = &c.s.cpos[t.ci[k]]; cp->x = t.x[k][0]; cp->y = t.x[k][1]; cp->z = t.x[k][2]; }
        any = true;
      }
    }
    (void)any;
    __syncthreads();
  }
  // bounding box of the replica: if extent + list radius + skin < L no listed pair can be closer
  // through a periodic image than directly, and it stays so until the next rebuild
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    if (i < p.P)
#pragma unroll
      for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], t.x[k][a]); hi[a] = fmaxf(hi[a], t.x[k][a]); }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  float *scr = reinterpret_cast<float *>(c.s.red);   // 32 warps x 6 floats fit in the scratch
  __syncthreads();
  if (c.tid == 0 && p.n_ions > 0) c.s.flag[3] = 0;   // rows are about to be current again
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
  c.dense = p.use_dh && (p.dense_policy == 2 || (p.dense_policy == 0 && c.rlist_d2 >= ext * ext));   // list radius >= extent: the rows prune nothing (measured: no gain below that)
  __syncthreads();   // every thread has read the bounding-box scratch: it now takes the block spheres
  block_spheres(c.s.pos, p.P, reinterpret_cast<float4 *>(c.s.red), c.tid, c.nt);
  __syncthreads();
  if (c.wrap) build_rows<ITEMS, true>(c, t); else build_rows<ITEMS, false>(c, t);
  __syncthreads();   // Debye-Hueckel rows are read by other threads (sub-thread split, MC warp)
  c.rebuilds += 1;
  c.compact = false;
}

// Compact list of the chargeable particles that carry charge right now (warp 0; callers put a
// barrier behind it).  Order = chargeable index, so results do not depend on who built it.
__device__ void build_qlist(Ctx &c) {
  const DevParams &p = c.p;
  if (c.warp == 0) {
    const int nw = (p.n_chg + 31) >> 5;
    int base = 0;
    for (int w = 0; w < nw; ++w) {
      const uint32_t m = c.s.qmask[w];
      if ((m >> c.lane) & 1u) c.s.qlist[base + __popc(m & ((1u << c.lane) - 1u))] = (uint16_t)(w * 32 + c.lane);
      base += __popc(m);
    }
    if (c.lane == 0) c.s.flag[0] = base;
  }
}

// The bit rows AND the charge mask say which partners a sub-thread visits, but walking them costs a
// find-first-set chain per partner and, worse, the lanes of a warp fall out of step at every word
// (each word lasts as long as its fullest lane).  Between two list builds of an MD burst neither
// rows nor charges change, so every sub-thread (site ci, class g) writes its partners out once,
// as a plain list it alone reads back: [k/8][sub-thread][k%8] in global memory (L2-resident, one
// coalesced 16-byte load per eight partners, the next one in flight while these are evaluated).
// The force loop is then a straight run of pairs, lanes in step until their lists end.
__device__ void build_dh_compact(Ctx &c) {
  const DevParams &p = c.p;
  c.compact = false;
  if (!p.use_dh || c.dense || c.g.dhc == nullptr || p.n_chg < 97) return;   // short rows (<= 3 words): the walk is cheap, few warps hide the loads
  // Work items exist for the particles that carry charge only (qlist, build_qlist): item idx = (sub-thread g, k-th charged
  // particle), packed, so that the warps of the force loop are full whatever the degree of ionisation (at pH 2.5 and
  // 1 mM a quarter of the sites is neutral: their lanes used to idle next to busy ones).  The force slots of the
  // others are zeroed here once, nobody writes them until the next build.
  const int G = p.dh_group, nq = c.s.flag[0];
  for (int ci = c.tid; ci < p.n_chg; ci += c.nt)
    if (c.s.cpos[ci].w == 0.0f)
      for (int k = 0; k < 3 * G; ++k) c.s.fd[(size_t)k * p.n_chg + ci] = 0.0f;
  uint16_t *lst0 = c.g.dhc + (size_t)c.r * p.dhc_cap * p.dhc_stride;
  for (int idx = c.tid; idx < G * nq; idx += c.nt) {
    const int g = idx / nq, ci = c.s.qlist[idx - g * nq];
    if (idx == c.tid) { c.dh_g0 = g; c.dh_ci0 = ci; }
    const uint32_t sel = G == 1 ? 0xFFFFFFFFu : G == 2 ? (0x55555555u << g) : G == 3 ? (0x49249249u << g) : (0x11111111u << g);
    uint16_t *lst = lst0 + (size_t)idx * 8;   // [k / 8][item][k % 8]: eight partners per 16-byte load
    int n = 0;
    const int w_end = c.s.wrange[2 * ci + 1];
    for (int w = c.s.wrange[2 * ci]; w < w_end; ++w) {
      uint32_t m = c.s.dbits[(size_t)ci * p.dw + w] & c.s.qmask[w] & sel;
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        lst[((size_t)(n >> 3) * p.dhc_stride) * 8 + (n & 7)] = (uint16_t)((w * 32 + b) * 16);
        ++n;
      }
    }
    c.s.dhn[idx] = (uint16_t)n;
    for (; n & 3; ++n) lst[((size_t)(n >> 3) * p.dhc_stride) * 8 + (n & 7)] = (uint16_t)(p.n_chg * 16);   // whole half groups
  }
  c.compact = true;   // every list is read only by the thread that wrote it: no barrier
  if (c.tid == 0) c.g.compact_builds[c.r] += 1;
}

// ---- forces ------------------------------------------------------------------------------------
// Phase 1 reads s.pos (must be visible), phase 2 gathers the slots.  thermostat: add
// -gamma v + sqrt(24 gamma kT/dt)(U-1/2) with U from Philox(counter, particle) (E.py:61-63).
// Owner-computes scheme for everything that is local to the chain: bead i evaluates bond (i-1,i),
// angle (i-1,i,i+1), dihedral (i-1,i,i+1,i+2) and the WCA pairs (i-1,i) and (i-1,i+1) once, keeps
// its own share and hands the partners' shares over through slots A (i-1), B (i+1), C (i+2).
template <int ITEMS, bool WRAP>
__device__ void compute_forces_impl(Ctx &c, Thr<ITEMS> &t, bool thermostat, unsigned long long counter) {
  const DevParams &p = c.p;
  const int NC = p.NC, S = p.T + 1;
  float *sa = c.s.slot, *sb = c.s.slot + 3 * NC, *sc = c.s.slot + 6 * NC;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    V3 f = {0.f, 0.f, 0.f};
    if (i < p.P) {
      const float4 pi4 = c.s.pos[i];
      const V3 pi = xyz(pi4);
      const int ti = c.s.type[i];
      const float4 *tab_i = c.s.tab + ti * S;
      const int l = t.l[k];
      if (l >= 1) {
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
          const float4 tab = tab_i[tm1];
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
            const V3 d = pp1 - pm1;
            const float r2 = dot(d, d);
            if (r2 < tab.z) {
              const V3 fw = wca_force_over_r(tab, r2) * d;
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
            sc[i + 2] = f4.x; sc[NC + i + 2] = f4.y; sc[2 * NC + i + 2] = f4.z;
          }
          sb[i + 1] = to_next.x; sb[NC + i + 1] = to_next.y; sb[2 * NC + i + 1] = to_next.z;
        }
        sa[i - 1] = to_prev.x; sa[NC + i - 1] = to_prev.y; sa[2 * NC + i - 1] = to_prev.z;
      }
      if (p.use_wca && ti < p.T) {
        const int n = c.s.wn[i];
        for (int e = 0; e < n; ++e) {
          const int j = c.s.wl[(size_t)e * p.P + i];
          const float4 pj4 = c.s.pos[j];
          const float4 tab = tab_i[c.s.type[j]];
          V3 d;
          const float r2 = dist2<WRAP>(c, pi4, pj4, &d);
          if (r2 < tab.z) f = f + wca_force_over_r(tab, r2) * d;
        }
      }
    }
    t.f[k][0] = f.x; t.f[k][1] = f.y; t.f[k][2] = f.z;
  }
  if (p.use_dh) {
    // Debye-Hueckel rows: G sub-threads per row, sub-thread g takes the bits b = g (mod G) of every
    // word, so the backbone neighbours ci+-1, ci+-2, ... of a site land on different sub-threads;
    // each writes its partial force into its own slot plane (summed in phase 2)
    const int G = p.dh_group;
    const int nq = c.compact ? c.s.flag[0] : p.n_chg;   // compact lists: work items of the charged particles only, packed
    for (int idx = c.tid; idx < G * nq; idx += c.nt) {
      int g, ci;
      if (c.compact) {
        if (idx == c.tid) { g = c.dh_g0; ci = c.dh_ci0; }
        else { g = idx / nq; ci = c.s.qlist[idx - g * nq]; }
      } else {
        g = idx / p.n_chg; ci = idx - g * p.n_chg;
      }
      const uint32_t sel = G == 1 ? 0xFFFFFFFFu : G == 2 ? (0x55555555u << g) : G == 3 ? (0x49249249u << g) : (0x11111111u << g);
      const float4 pi4 = c.s.cpos[ci];
      V3 f = {0.f, 0.f, 0.f};
      if (pi4.w != 0.0f && c.dense) {
        // dense mode: every charged particle is a candidate; sub-thread g takes every G-th one
        const float qi = p.dh_pref * pi4.w;
        const int nq = c.s.flag[0];
        for (int k = g; k < nq; k += G) {
          const int cj = c.s.qlist[k];
          const float4 pj4 = c.s.cpos[cj];
          V3 d;
          const float r2 = dist2<WRAP>(c, pi4, pj4, &d);
          if (r2 < c.rc2 && cj != ci) {
            float fr;
            dh_energy_unit(r2, c.kappa, c.neg_kappa_log2e, &fr);
            fr *= qi * pj4.w;
            f.x = fmaf(fr, d.x, f.x); f.y = fmaf(fr, d.y, f.y); f.z = fmaf(fr, d.z, f.z);
          }
        }
      } else if (pi4.w != 0.0f && c.compact) {
        // entries are byte offsets into cpos, groups of eight padded with the far-away uncharged dummy:
        // no per-entry bounds test, one per half group (WRAP keeps it: a folded dummy may land anywhere), no branch on the
        // cutoff (91 % of the listed pairs are inside it; the rest select a zero), the shared address of
        // cpos and the two loop constants pinned in registers instead of re-derived per pair
        const int n = c.s.dhn[idx];
        const uint4 *lst = reinterpret_cast<const uint4 *>(c.g.dhc + (size_t)c.r * p.dhc_cap * p.dhc_stride) + idx;
        uint32_t cpos_sa = (uint32_t)__cvta_generic_to_shared(c.s.cpos);
        float nk = c.neg_kappa_log2e, rc2 = c.rc2;
        asm volatile("" : "+r"(cpos_sa), "+f"(nk), "+f"(rc2));
        uint4 cur = n > 0 ? __ldcg(lst) : make_uint4(0u, 0u, 0u, 0u);
        for (int k0 = 0; k0 < n; k0 += 8) {
          lst += p.dhc_stride;
          const uint4 nxt = k0 + 8 < n ? __ldcg(lst) : make_uint4(0u, 0u, 0u, 0u);
          const uint32_t pk[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half == 1 && k0 + 4 >= n) break;   // short rows (high salt): the padding stops at the half group
#pragma unroll
            for (int e = 4 * half; e < 4 * half + 4; ++e) {
              const uint32_t off = (e & 1) ? (pk[e >> 1] >> 16) : (pk[e >> 1] & 0xFFFFu);
              const float4 pj4 = lds_f4(cpos_sa + off);
              V3 d;
              const float r2 = dist2<WRAP>(c, pi4, pj4, &d);
              bool ok = r2 < rc2;
              if (WRAP) ok = ok && (k0 + e < n);
              float fr = dh_force_unit_q(r2, nk, pj4.w);
              fr = ok ? fr : 0.0f;
              f.x = fmaf(fr, d.x, f.x); f.y = fmaf(fr, d.y, f.y); f.z = fmaf(fr, d.z, f.z);
            }
          }
          cur = nxt;
        }
        const float qi = p.dh_pref * pi4.w;
        f.x *= qi; f.y *= qi; f.z *= qi;
      } else if (pi4.w != 0.0f) {
        const float qi = p.dh_pref * pi4.w;
        const int w_end = c.s.wrange[2 * ci + 1];
        for (int w = c.s.wrange[2 * ci]; w < w_end; ++w) {
          uint32_t m = c.s.dbits[(size_t)ci * p.dw + w] & c.s.qmask[w] & sel;   // listed AND charged AND mine
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const float4 pj4 = c.s.cpos[w * 32 + b];
            V3 d;
            const float r2 = dist2<WRAP>(c, pi4, pj4, &d);
            if (r2 < c.rc2) {
              float fr;
              dh_energy_unit(r2, c.kappa, c.neg_kappa_log2e, &fr);
              fr *= qi * pj4.w;
              f.x = fmaf(fr, d.x, f.x); f.y = fmaf(fr, d.y, f.y); f.z = fmaf(fr, d.z, f.z);
            }
          }
        }
      }
      float *fdg = c.s.fd + (size_t)g * 3 * p.n_chg;
      fdg[ci] = f.x; fdg[p.n_chg + ci] = f.y; fdg[2 * p.n_chg + ci] = f.z;
    }
  }
  __syncthreads();
  const uint32_t k0 = p.thermostat_seed, k1 = (uint32_t)(p.replica_offset + c.r);
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    if (i < p.P) {
      float fx = t.f[k][0], fy = t.f[k][1], fz = t.f[k][2];
      if (t.l[k] >= 0) {
        fx += sa[i] + sb[i] + sc[i];
        fy += sa[NC + i] + sb[NC + i] + sc[NC + i];
        fz += sa[2 * NC + i] + sb[2 * NC + i] + sc[2 * NC + i];
      }
      if (p.use_dh) {
        const int ci = c.s.cidx[i];
        if (ci != 0xFFFF && c.s.cpos[ci].w != 0.0f)   // a particle without charge feels no Debye-Hueckel force: its slots hold zeros
          for (int g = 0; g < p.dh_group; ++g) {
            const float *fdg = c.s.fd + (size_t)g * 3 * p.n_chg;
            fx += fdg[ci]; fy += fdg[p.n_chg + ci]; fz += fdg[2 * p.n_chg + ci];
          }
      }
      if (thermostat && c.s.type[i] < p.T) {
        const uint4 o = philox4x32_10((uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)i, kTagLangevin, k0, k1);
        fx += fmaf(-c.gamma, t.v[k][0], c.noise_pref * centred_u23(o.x));
        fy += fmaf(-c.gamma, t.v[k][1], c.noise_pref * centred_u23(o.y));
        fz += fmaf(-c.gamma, t.v[k][2], c.noise_pref * centred_u23(o.z));
      }
      if (p.force_cap > 0.0f) {  // system.force_cap (E.py:32,48,51)
        const float n2 = fx * fx + fy * fy + fz * fz;
        if (n2 > p.force_cap * p.force_cap) {
          const float sc_ = p.force_cap * rsqrtf(n2);
          fx *= sc_; fy *= sc_; fz *= sc_;
        }
      }
      if (p.fixed_mask != 0u && ((p.fixed_mask >> c.s.type[i]) & 1u)) fx = fy = fz = 0.0f;   // fix=[True]*3
      t.f[k][0] = fx; t.f[k][1] = fy; t.f[k][2] = fz;
    }
  }
}

template <int ITEMS>
__device__ __forceinline__ void compute_forces(Ctx &c, Thr<ITEMS> &t, bool thermostat, unsigned long long counter) {
  if (c.wrap) compute_forces_impl<ITEMS, true>(c, t, thermostat, counter);
  else compute_forces_impl<ITEMS, false>(c, t, thermostat, counter);
}

// ---- frame recorder (cgpe_frames_begin): the observables of Observables.py:54-76 without leaving the launch ----
// COM position / velocity, Rg, Ree of chain 0 [+ bond lengths, torsion angles] appended to the replica's frame buffers.
// All threads; positions in shared memory current; contains block barriers.  Not inlined, and handed plain values instead of
// the kernel's context: it runs once in a hundred steps and must not cost the step loop registers.
template <int ITEMS>
struct FrameArgs { float x[ITEMS][3], v[ITEMS][3]; };

template <int ITEMS>
__device__ __noinline__ void record_frame(const DevParams &p, const DevState &g, const float4 *pos, double *red, int r, int k,
                                          FrameArgs<ITEMS> t) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const float4 p0 = pos[0];
  double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
#pragma unroll
  for (int q = 0; q < ITEMS; ++q) {
    const int i = tid + q * nt;
    if (i < p.N) {
      const double dx = (double)t.x[q][0] - p0.x, dy = (double)t.x[q][1] - p0.y, dz = (double)t.x[q][2] - p0.z;
      a[0] += dx; a[1] += dy; a[2] += dz; a[3] += dx * dx + dy * dy + dz * dz;
      b[0] += t.v[q][0]; b[1] += t.v[q][1]; b[2] += t.v[q][2];
    }
  }
  double out[8];
  for (int pass = 0; pass < 2; ++pass) {
    double *v = pass == 0 ? a : b;
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = warp_sum(v[q]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
      for (int q = 0; q < 4; ++q) red[warp * 4 + q] = v[q];
    __syncthreads();
    double s[4] = {0, 0, 0, 0};
    for (int w = 0; w < (nt + 31) / 32; ++w)
#pragma unroll
      for (int q = 0; q < 4; ++q) s[q] += red[w * 4 + q];
    const double inv = 1.0 / p.N;
    if (pass == 0) {
      out[0] = p0.x + s[0] * inv; out[1] = p0.y + s[1] * inv; out[2] = p0.z + s[2] * inv;
      const double var = s[3] * inv - (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) * inv * inv;
      out[6] = sqrt(var > 0.0 ? var : 0.0);
    } else {
      out[3] = s[0] * inv; out[4] = s[1] * inv; out[5] = s[2] * inv;
    }
  }
  if (k >= p.frame_cap) return;   // dropped; n_frames keeps counting and the host reports it
  const size_t slot = (size_t)r * p.frame_cap + k;
  if (tid == 0) {
    const float4 pe = pos[p.N - 1];
    const double ex = (double)pe.x - p0.x, ey = (double)pe.y - p0.y, ez = (double)pe.z - p0.z;
    out[7] = sqrt(ex * ex + ey * ey + ez * ez);
    double *f = g.frames + slot * CGPE_FRAME_COLS;
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = out[q];
  }
  if ((p.frame_flags & CGPE_FRAME_BONDS) && g.frame_bonds)
    for (int i = tid + 1; i < p.N; i += nt) {
      const V3 d = xyz(pos[i]) - xyz(pos[i - 1]);
      g.frame_bonds[slot * (p.N - 1) + i - 1] = sqrtf(dot(d, d));
    }
  if ((p.frame_flags & CGPE_FRAME_DIHEDRALS) && g.frame_dih && p.N >= 4)
    for (int i = tid; i < p.N - 3; i += nt) {   // torsion of (i, i+1, i+2, i+3), the convention of observables.BondDihedrals
      const V3 b1 = xyz(pos[i + 1]) - xyz(pos[i]), b2 = xyz(pos[i + 2]) - xyz(pos[i + 1]), b3 = xyz(pos[i + 3]) - xyz(pos[i + 2]);
      const V3 n1 = cross(b1, b2), n2 = cross(b2, b3);
      const V3 m = cross(n1, rsqrtf(dot(b2, b2)) * b2);
      g.frame_dih[slot * (p.N - 3) + i] = atan2f(dot(m, n2), dot(n1, n2));
    }
}

// ---- velocity Verlet + Langevin (R2) ------------------------------------------------------------
template <int ITEMS, bool FRAMES>
__device__ void md_steps(Ctx &c, Thr<ITEMS> &t, int n_steps, unsigned long long &counter, bool &f_valid,
                         bool &lists_valid) {
  const DevParams &p = c.p;
  if (n_steps <= 0) return;
  sync_ions(c, t);   // explicit ions moved or hidden by trial moves since the last burst
  if (!lists_valid) { build_lists(c, t); lists_valid = true; }
  if (p.use_dh) { build_qlist(c); __syncthreads(); }   // charges may have changed since the last burst
  build_dh_compact(c);
  if (!f_valid) { compute_forces(c, t, true, counter); f_valid = true; }
  for (int step = 0; step < n_steps; ++step) {
    int moved = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int i = c.tid + k * c.nt;
      if (i < p.P) {
        float d2 = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          t.v[k][a] = fmaf(p.half_dt, t.f[k][a], t.v[k][a]);
          t.x[k][a] = fmaf(p.dt, t.v[k][a], t.x[k][a]);
          const float d = t.x[k][a] - t.ref[k][a];
          d2 = fmaf(d, d, d2);
        }
        float4 *sp = &c.s.pos[i];
        sp->x = t.x[k][0]; sp->y = t.x[k][1]; sp->z = t.x[k][2];
        if (t.ci[k] >= 0) { float4 *cp = &c.s.cpos[t.ci[k]]; cp->x = t.x[k][0]; cp->y = t.x[k][1]; cp->z = t.x[k][2]; }
        moved |= !(d2 <= p.half_skin2);  // also true for NaN
      }
    }
    if (__syncthreads_or(moved)) { build_lists(c, t); build_dh_compact(c); }
    counter += 1;
    compute_forces(c, t, true, counter);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
#pragma unroll
      for (int a = 0; a < 3; ++a) t.v[k][a] = fmaf(p.half_dt, t.f[k][a], t.v[k][a]);
    if (FRAMES && ++c.frame_since >= p.frame_every) {   // uniform: every thread counts the same steps (compiled in only when armed)
      c.frame_since = 0;
      FrameArgs<ITEMS> fa;
#pragma unroll
      for (int k = 0; k < ITEMS; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a) { fa.x[k][a] = t.x[k][a]; fa.v[k][a] = t.v[k][a]; }
      record_frame<ITEMS>(p, c.g, c.s.pos, c.s.red, c.r, c.frame_n, fa);
      c.frame_n += 1;
    }
  }
  c.compact = false;   // trial moves change the charges
}

// ---- constant-pH trial moves (R1) ----------------------------------------------------------------
// Positions do not move inside a reaction block, so everything a trial needs is cached per
// chargeable particle by ALL threads (mc_prepare):
//   phi[ci]    = sum_j q_j exp(-kappa r_ij)/r_ij over the listed, charged partners inside r_cut
//   wcache[ci] = sum_j WCA(t_deprot, t_j, r_ij) - WCA(t_prot, t_j, r_ij) over chain neighbours + row
// A trial (warp 0) is then: direction and site from the Philox draw, Delta-E = +-wcache + pref dq phi,
// Metropolis decision, O(1) swap-remove bookkeeping: a handful of dependent loads.  Only an ACCEPTED
// flip does lane-parallel work: it adds dq e(r) to phi of every listed partner of the flipped site
// and refreshes wcache of partners that are themselves titratable.

// WCA type-swap energy of chargeable particle ci (serial over its partners)
__device__ float wca_swap_energy(const Ctx &c, int ci, bool *titratable_partner = nullptr) {
  const DevParams &p = c.p;
  const int m = c.s.rxof[ci];
  if (titratable_partner) *titratable_partner = false;
  if (m == 255 || !p.use_wca) return 0.f;
  const int S = p.T + 1, i = c.s.chg[ci];
  const int tp = p.rx[m].type_prot, td = p.rx[m].type_deprot;
  const float4 pi4 = c.s.pos[i];
  const int l = (p.NC == p.N) ? (i < p.NC ? i : -1) : (i < p.NC ? i % p.N : -1);
  const int n_w = c.s.wn[i];
  float e = 0.f;
  for (int k = 0; k < n_w + 4; ++k) {
    int j = -1;
    if (k < 4) {
      const int off = k < 2 ? k - 2 : k - 1;   // -2,-1,+1,+2: the chain neighbours are not in the rows
      if (l >= 0 && l + off >= 0 && l + off < p.N) j = i + off;
    } else {
      j = c.s.wl[(size_t)(k - 4) * p.P + i];
    }
    if (j < 0) continue;
    const int tj = c.s.type[j];
    if (tj >= p.T) continue;
    V3 d;
    const float r2 = dist2<true>(c, pi4, c.s.pos[j], &d);
    // a titratable partner inside the largest WCA cutoff: its own cache depends on this particle's type
    if (titratable_partner && r2 < p.rc_w_max2 && c.s.cidx[j] != 0xFFFF && c.s.rxof[c.s.cidx[j]] != 255)
      *titratable_partner = true;
    e += wca_energy(c.s.tab[td * S + tj], r2) - wca_energy(c.s.tab[tp * S + tj], r2);
  }
  return e;
}

// all threads; needs positions, types, rows and qmask current; ends with a barrier
__device__ void mc_prepare(Ctx &c) {
  const DevParams &p = c.p;
  const int G = p.dh_group;
  if (p.use_dh) {
    for (int idx = c.tid; idx < G * p.n_chg; idx += c.nt) {
      const int g = idx / p.n_chg, ci = idx - g * p.n_chg;
      const uint32_t sel = G == 1 ? 0xFFFFFFFFu : G == 2 ? (0x55555555u << g) : G == 3 ? (0x49249249u << g) : (0x11111111u << g);
      const float4 pi4 = c.s.cpos[ci];
      float sum = 0.f;
      const int w_end = c.s.wrange[2 * ci + 1];
      for (int w = c.s.wrange[2 * ci]; w < w_end; ++w) {
        uint32_t m = c.s.dbits[(size_t)ci * p.dw + w] & c.s.qmask[w] & sel;
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const float4 pj4 = c.s.cpos[w * 32 + b];
          V3 d;
          const float r2 = c.wrap ? dist2<true>(c, pi4, pj4, &d) : dist2<false>(c, pi4, pj4, &d);
          if (r2 < c.rc2) {
            float fr;
            sum = fmaf(pj4.w, dh_energy_unit(r2, c.kappa, c.neg_kappa_log2e, &fr), sum);
          }
        }
      }
      c.s.fd[(size_t)g * 3 * p.n_chg + ci] = sum;   // the x plane of the force slots doubles as scratch
    }
  }
  for (int ci = c.tid; ci < p.n_chg; ci += c.nt) {
    const int t = c.s.type[c.s.chg[ci]];
    int m = 255;
    for (int k = 0; k < p.n_rx; ++k)
      if (t == p.rx[k].type_prot || t == p.rx[k].type_deprot) m = k;
    c.s.rxof[ci] = (unsigned char)m;
  }
  // N/(N_sites - N + 1) for N = 0..N_sites: the N_react/(N_prod+1) factor without a division per trial
  for (int m = 0; m < p.n_rx; ++m) {
    const int n_sites = c.s.cnt[m * 2] + c.s.cnt[m * 2 + 1];
    for (int n = c.tid; n <= n_sites; n += c.nt)
      c.s.ratio[(size_t)m * (p.n_chg + 1) + n] = (float)n / (float)(n_sites - n + 1);
  }
  if (p.use_dh && c.g.mc_e) {
    // Positions rest during the reaction blocks of this sample: the pair kernel e(r_ij) = exp(-kappa r)/r of every listed pair
    // inside r_cut is evaluated ONCE here, by all threads, into an L2-resident matrix; an accepted flip then updates the cached
    // potentials of its partners with one row read (mc_block) instead of a distance + rsqrt + ex2 chain per partner behind two
    // block-wide barriers.  Same arithmetic as mc_phi_word, so the Markov chain is unchanged bit for bit.
    const int stride = c.g.mc_e_stride;
    float *e0 = c.g.mc_e + (size_t)c.r * p.n_chg * stride;
    for (int idx = c.tid; idx < p.n_chg * stride; idx += c.nt) {
      const int ci = idx / stride, ck = idx - ci * stride;
      float e = 0.f;
      if (ck < p.n_chg && ((c.s.dbits[(size_t)ci * p.dw + (ck >> 5)] >> (ck & 31)) & 1u)) {
        V3 d;
        const float4 pi4 = c.s.cpos[ci], pk4 = c.s.cpos[ck];
        const float r2 = c.wrap ? dist2<true>(c, pi4, pk4, &d) : dist2<false>(c, pi4, pk4, &d);
        if (r2 < c.rc2) {
          float fr;
          e = dh_energy_unit(r2, c.kappa, c.neg_kappa_log2e, &fr);
        }
      }
      __stcg(e0 + idx, e);
    }
  }
  __syncthreads();
  for (int ci = c.tid; ci < p.n_chg; ci += c.nt) {
    float ph = 0.f;
    if (p.use_dh)
      for (int g = 0; g < G; ++g) ph += c.s.fd[(size_t)g * 3 * p.n_chg + ci];
    c.s.phi[ci] = ph;
    bool flag;
    c.s.wcache[ci] = wca_swap_energy(c, ci, &flag);
    c.s.wflag[ci] = flag ? 1 : 0;
  }
  __syncthreads();
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// With explicit ions (p.n_ions > 0) a forward move also inserts one H+ ("B") at a uniform position
// with a Maxwell-Boltzmann velocity and a backward move takes a random free H+ away:
// the ion's Debye-Hueckel energy with every particle and its distance to the nearest one
// (exclusion range, M.py:202-206) come from one lane-parallel pass over all particles.  The ion
// carries no WCA energy because cgpe_create insists on exclusion_range >= its largest WCA cutoff.
//
// An accepted flip changes the cached potential of every listed partner of the site: up to one
// 32-lane pass per word of its row, a chain of dependent shared-memory accesses and MUFU calls each.
// Warp 0 posts the flip as an order (site, dq, position) and the other warps, parked at a named
// barrier while the trial loop runs, take one word each.
__device__ __forceinline__ void mc_bar(int n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }

__device__ __forceinline__ void mc_phi_word(Ctx &c, int ci, int w, float dq, float4 pi4) {
  const uint32_t mw = c.s.dbits[(size_t)ci * c.p.dw + w];
  if ((mw >> c.lane) & 1u) {
    const int ck = w * 32 + c.lane;
    V3 d;
    const float4 pk4 = c.s.cpos[ck];
    const float r2 = c.wrap ? dist2<true>(c, pi4, pk4, &d) : dist2<false>(c, pi4, pk4, &d);
    if (r2 < c.rc2) {
      float fr;
      c.s.phi[ck] = fmaf(dq, dh_energy_unit(r2, c.kappa, c.neg_kappa_log2e, &fr), c.s.phi[ck]);
    }
  }
}

template <bool IONS>
__device__ void mc_block(Ctx &c, int m, int n_trials, unsigned long long &trial_ctr, long long &accepted,
                         cgpe_trial_record *trace) {
  const DevParams &p = c.p;
  const bool use_matrix = !IONS && p.use_dh && c.g.mc_e != nullptr;   // implicit ions: partner updates from the pair-kernel matrix
  const bool helpers = p.use_dh && c.nt > 32 && !use_matrix;   // same for every thread
  float *ord = reinterpret_cast<float *>(c.s.red);   // {x, y, z, -, dq, ci, op}: the reduction scratch is idle here
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
    constexpr bool ions = IONS;   // compile-time: the implicit-ion kernels carry none of the ion code
    uint16_t *lp = c.s.lp + (size_t)m * p.n_chg, *ld = c.s.ld + (size_t)m * p.n_chg;
    int n_p = c.s.cnt[m * 2], n_d = c.s.cnt[m * 2 + 1];
    int n_act = ions ? c.s.flag[1] : 0, n_free = ions ? c.s.flag[2] : 0;
    const float arg_fwd = (float)(2.302585092994045684 * (c.g.pH[c.r] - rx.pK)) * kLog2e;
    const float neg_beta_log2e = -rx.inv_kT * kLog2e;
    const float dq_fwd = rx.q_deprot - rx.q_prot, coul_fwd = p.dh_pref * dq_fwd;
    const uint32_t k0 = rx.seed, k1 = (uint32_t)(p.replica_offset + c.r);
    const int nw = (p.n_chg + 31) >> 5;
    const float *ratio = c.s.ratio + (size_t)m * (p.n_chg + 1);
    long long acc = 0;
    bool ion_moved = false;
    for (int base = 0; base < n_trials; base += 32) {
      const unsigned long long cnt = trial_ctr + (unsigned long long)(base + c.lane);
      const uint4 o = philox4x32_10((uint32_t)cnt, (uint32_t)(cnt >> 32), 0u, kTagMc + (uint32_t)m, k0, k1);
      uint4 o1 = make_uint4(0, 0, 0, 0), o2 = o1;
      if (ions) {   // block 1: insertion position, block 2: insertion velocity
        o1 = philox4x32_10((uint32_t)cnt, (uint32_t)(cnt >> 32), 1u, kTagMc + (uint32_t)m, k0, k1);
        o2 = philox4x32_10((uint32_t)cnt, (uint32_t)(cnt >> 32), 2u, kTagMc + (uint32_t)m, k0, k1);
      }
      const int nb = min(32, n_trials - base);
      for (int tt = 0; tt < nb; ++tt) {
        const uint32_t a0 = __shfl_sync(0xffffffffu, o.x, tt), a1 = __shfl_sync(0xffffffffu, o.y, tt),
                       a2 = __shfl_sync(0xffffffffu, o.z, tt);
        const bool fwd = (a0 >> 31) == 0;
        const int n_src = fwd ? n_p : n_d, n_dst = fwd ? n_d : n_p;
        uint16_t *src = fwd ? lp : ld, *dst = fwd ? ld : lp;
        const float u = u24(a2);
        // every reactant must exist: the site, and with explicit ions a free H+ to take (backward)
        // or a spare slot to put one (forward)
        bool possible = n_src > 0;
        if (ions && !fwd && n_act == 0) possible = false;
        if (ions && fwd && n_free == 0) { possible = false; c.err |= kErrIonPool; }
        if (!possible) {
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
        int b = -1, kb = 0;
        bool excluded = false;
        float4 pb4 = make_float4(0.f, 0.f, 0.f, rx.q_ion);
        if (ions) {
          if (fwd) {
            b = c.s.ion_free[n_free - 1];
            pb4.x = u24(__shfl_sync(0xffffffffu, o1.x, tt)) * p.box_l;
            pb4.y = u24(__shfl_sync(0xffffffffu, o1.y, tt)) * p.box_l;
            pb4.z = u24(__shfl_sync(0xffffffffu, o1.z, tt)) * p.box_l;
          } else {
            kb = (int)__umulhi(__shfl_sync(0xffffffffu, o.w, tt), (uint32_t)n_act);
            b = c.s.ion_active[kb];
            const float4 x = c.s.pos[b];
            pb4.x = x.x; pb4.y = x.y; pb4.z = x.z;
          }
          float esum = 0.f, dmin2 = 3.0e38f;
          for (int j = c.lane; j < p.P; j += 32) {
            if (j == b) continue;
            const int tj = j == i ? t_new : (int)c.s.type[j];   // the site is seen in its new state
            if (tj >= p.T) continue;
            const float4 pj4 = c.s.pos[j];
            const float qj = j == i ? q_new : pj4.w;
            V3 d;
            const float r2 = dist2<true>(c, pb4, pj4, &d);
            dmin2 = fminf(dmin2, r2);
            if (p.use_dh && qj != 0.f && r2 < c.rc2) {
              float fr;
              esum = fmaf(qj, dh_energy_unit(r2, c.kappa, c.neg_kappa_log2e, &fr), esum);
            }
          }
          esum = warp_sum(esum);
          dmin2 = warp_min(dmin2);
          const float e_ion = p.dh_pref * rx.q_ion * esum;
          de += fwd ? e_ion : -e_ion;
          excluded = dmin2 < rx.excl2;
        }
        // bf = N_react/(N_prod+1) exp(-dE/kT + nu ln10 (pH-pK))   (SURVEY 8a-R1); E_new = +max if excluded
        const float bf = excluded ? 0.f : ratio[n_src] * exp2f(fmaf(neg_beta_log2e, de, fwd ? arg_fwd : -arg_fwd));
        const bool accept = u < bf;
        if (trace && c.lane == 0) {
          cgpe_trial_record rec = {i, b, (int8_t)(fwd ? 1 : -1), (int8_t)accept, (int8_t)excluded, 0, de, bf, u};
          trace[base + tt] = rec;
        }
        if (accept) {
          const float4 pi4 = c.s.cpos[ci];
          // the flipped site's row: lane w fetches word w (before the bookkeeping stores)
          const uint32_t row = (p.use_dh && c.lane < nw) ? c.s.dbits[(size_t)ci * p.dw + c.lane] : 0u;
          const bool refresh = p.use_wca && c.s.wflag[ci] != 0;
          if (c.lane == 0) {
            c.s.type[i] = (unsigned char)t_new;
            c.s.pos[i].w = q_new;
            c.s.cpos[ci].w = q_new;
            if (q_new != 0.f) c.s.qmask[ci >> 5] |= 1u << (ci & 31);
            else c.s.qmask[ci >> 5] &= ~(1u << (ci & 31));
            src[k] = src[n_src - 1];
            dst[n_dst] = (uint16_t)i;
          }
          if (fwd) { n_p -= 1; n_d += 1; } else { n_d -= 1; n_p += 1; }
          acc += 1;
          if (use_matrix && dq != 0.f) {
            // phi of every partner changes by dq e(r): one coalesced read of the flipped site's matrix row (four partners per
            // lane and load, all loads of a round in flight together), zeros for partners that are not listed or out of range
            const int stride = c.g.mc_e_stride;
            const float4 *erow = reinterpret_cast<const float4 *>(c.g.mc_e + ((size_t)c.r * p.n_chg + ci) * stride);
            float4 *phi4 = reinterpret_cast<float4 *>(c.s.phi);
            for (int g0 = 0; g0 * 128 < p.n_chg; g0 += 4) {
              float4 e[4];
#pragma unroll
              for (int g = 0; g < 4; ++g)
                e[g] = (g0 + g) * 128 < p.n_chg ? __ldcg(erow + (g0 + g) * 32 + c.lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if ((g0 + g) * 128 < p.n_chg) {
                  float4 ph = phi4[(g0 + g) * 32 + c.lane];
                  ph.x = fmaf(dq, e[g].x, ph.x); ph.y = fmaf(dq, e[g].y, ph.y); ph.z = fmaf(dq, e[g].z, ph.z); ph.w = fmaf(dq, e[g].w, ph.w);
                  phi4[(g0 + g) * 32 + c.lane] = ph;
                }
            }
          } else if (p.use_dh && dq != 0.f) {
            // phi of every listed partner inside r_cut changes by dq e(r): lane b handles bit b of
            // each non-empty word (the helpers one word each, see above)
            // (a row of one or two words -- high salt -- costs this warp less than two barriers of 32 warps;
            // barriers of a small CTA are cheap: measured at N = 100, 8 CTAs per SM)
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
          if (ions) {
            const int cb = c.s.cidx[b];
            const float dqb = fwd ? rx.q_ion : -rx.q_ion;
            if (c.lane == 0) {
              if (fwd) {   // the hidden slot becomes a free H+ at the drawn position
                c.s.type[b] = (unsigned char)rx.type_ion;
                c.s.pos[b] = pb4;
                c.s.cpos[cb] = pb4;
                c.s.qmask[cb >> 5] |= 1u << (cb & 31);
              } else {     // the H+ turns into the non-interacting type and is gone (M.py:256-261)
                c.s.type[b] = (unsigned char)p.T;
                c.s.pos[b].w = 0.f;
                c.s.cpos[cb].w = 0.f;
                c.s.qmask[cb >> 5] &= ~(1u << (cb & 31));
                c.s.ion_vel[b - p.NC] = make_float4(0.f, 0.f, 0.f, 1.f);
                c.s.ion_active[kb] = c.s.ion_active[n_act - 1];
                c.s.ion_free[n_free] = (uint16_t)b;
              }
            }
            if (fwd) {
              // velocity of the new ion: all lanes fetch the draw, lane 0 stores it
              const uint32_t w0 = __shfl_sync(0xffffffffu, o2.x, tt), w1 = __shfl_sync(0xffffffffu, o2.y, tt),
                             w2 = __shfl_sync(0xffffffffu, o2.z, tt), w3 = __shfl_sync(0xffffffffu, o2.w, tt);
              if (c.lane == 0) {
                const float ua = ((float)(w0 >> 8) + 0.5f) * (1.0f / 16777216.0f), ub = ((float)(w1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
                const float uc = ((float)(w2 >> 8) + 0.5f) * (1.0f / 16777216.0f), ud = ((float)(w3 >> 8) + 0.5f) * (1.0f / 16777216.0f);
                const float ra = rx.sqrt_kT * sqrtf(-2.0f * logf(ua)), rb = rx.sqrt_kT * sqrtf(-2.0f * logf(uc));
                float sa, ca, sb_, cb_;
                sincosf(6.283185307179586f * ub, &sa, &ca);
                sincosf(6.283185307179586f * ud, &sb_, &cb_);
                c.s.ion_vel[b - p.NC] = make_float4(ra * ca, ra * sa, rb * cb_, 1.f);
                c.s.ion_active[n_act] = (uint16_t)b;
              }
              n_free -= 1; n_act += 1;
            } else {
              n_act -= 1; n_free += 1;
            }
            // the ion's charge appears in / leaves the potential of every chargeable particle in range
            if (p.use_dh)
              for (int ck = c.lane; ck < p.n_chg; ck += 32) {
                if (ck == cb) continue;
                V3 d;
                const float r2 = dist2<true>(c, pb4, c.s.cpos[ck], &d);
                if (r2 < c.rc2) {
                  float fr;
                  c.s.phi[ck] = fmaf(dqb, dh_energy_unit(r2, c.kappa, c.neg_kappa_log2e, &fr), c.s.phi[ck]);
                }
              }
            ion_moved = true;
            __syncwarp();
          }
          if (refresh) {
            // partners of i that are titratable themselves see a new neighbour type: refresh their cache
            const int l = (p.NC == p.N) ? (i < p.NC ? i : -1) : (i < p.NC ? i % p.N : -1);
            const int n_w = c.s.wn[i];
            for (int e = c.lane; e < n_w + 4; e += 32) {
              int j = -1;
              if (e < 4) {
                const int off = e < 2 ? e - 2 : e - 1;
                if (l >= 0 && l + off >= 0 && l + off < p.N) j = i + off;
              } else {
                j = c.s.wl[(size_t)(e - 4) * p.P + i];
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
    if (c.lane == 0) {
      c.s.cnt[m * 2] = n_p; c.s.cnt[m * 2 + 1] = n_d;
      if (ions) { c.s.flag[1] = n_act; c.s.flag[2] = n_free; if (ion_moved) c.s.flag[3] = 1; }
    }
    accepted += acc;  // only meaningful on warp 0; thread 0 writes it back
  }
  trial_ctr += (unsigned long long)n_trials;
  __syncthreads();
}

// ---- observables (R5, R6) ----------------------------------------------------------------------
// Rg, Ree of chain 0 from unfolded coordinates; double accumulation (N*12 flop per sample).
template <int ITEMS>
__device__ void chain_observables(Ctx &c, Thr<ITEMS> &t, double *rg, double *ree) {
  const DevParams &p = c.p;
  const float4 p0 = c.s.pos[0];
  double sx = 0, sy = 0, sz = 0, s2 = 0;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    if (i < p.N) {
      const double dx = (double)t.x[k][0] - p0.x, dy = (double)t.x[k][1] - p0.y, dz = (double)t.x[k][2] - p0.z;
      sx += dx; sy += dy; sz += dz; s2 += dx * dx + dy * dy + dz * dz;
    }
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); s2 = warp_sum(s2);
  __syncthreads();
  if (c.lane == 0) { c.s.red[c.warp * 4] = sx; c.s.red[c.warp * 4 + 1] = sy; c.s.red[c.warp * 4 + 2] = sz; c.s.red[c.warp * 4 + 3] = s2; }
  __syncthreads();
  const int nw = (c.nt + 31) / 32;
  double tx = 0, ty = 0, tz = 0, t2 = 0;
  for (int w = 0; w < nw; ++w) { tx += c.s.red[w * 4]; ty += c.s.red[w * 4 + 1]; tz += c.s.red[w * 4 + 2]; t2 += c.s.red[w * 4 + 3]; }
  const double inv = 1.0 / p.N;
  // mean |r - r_cm|^2 = mean |d|^2 - |mean d|^2 with d = r - r_0 (shifted to the chain, so no cancellation)
  const double v = t2 * inv - (tx * tx + ty * ty + tz * tz) * inv * inv;
  *rg = sqrt(v > 0.0 ? v : 0.0);
  const float4 pe = c.s.pos[p.N - 1];
  const double ex = (double)pe.x - p0.x, ey = (double)pe.y - p0.y, ez = (double)pe.z - p0.z;
  *ree = sqrt(ex * ex + ey * ey + ez * ez);
}

__device__ void init_ctx(Ctx &c) {
  const DevParams &p = c.p;
  c.tid = threadIdx.x; c.nt = blockDim.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
  c.kappa = c.g.kappa[c.r];
  const float rc = c.g.rcut[c.r];
  c.rc2 = rc * rc;
  c.neg_kappa_log2e = -c.kappa * kLog2e;
  c.rlist_d2 = (rc + p.skin) * (rc + p.skin);
  c.rlist_max = fmaxf(p.use_dh ? rc + p.skin : 0.f, p.use_wca ? sqrtf(p.rlist_wf2) : 0.f);
  c.wrap = true;
  c.dense = false;
  c.compact = false;
  c.dh_g0 = 0;
  c.dh_ci0 = 0;
  c.gamma = c.g.gamma[c.r];
  c.noise_pref = sqrtf(24.0f * c.gamma * c.g.kT[c.r] / p.dt);
  c.err = 0;
  c.rebuilds = 0;
  c.frame_since = p.frame_every > 0 ? c.g.frame_phase[c.r] : 0;
  c.frame_n = p.frame_every > 0 ? c.g.n_frames[c.r] : 0;
}

extern __shared__ __align__(16) unsigned char smem_raw[];

__device__ __forceinline__ void bind_smem(const DevParams &p, ChainSmem *s) {
  const SmemOff &f = p.off;
  s->pos = reinterpret_cast<float4 *>(smem_raw + f.pos);
  s->tab = reinterpret_cast<float4 *>(smem_raw + f.tab);
  s->red = reinterpret_cast<double *>(smem_raw + f.red);
  s->slot = reinterpret_cast<float *>(smem_raw + f.slot);
  s->fd = reinterpret_cast<float *>(smem_raw + f.fd);
  s->dhn = reinterpret_cast<uint16_t *>(smem_raw + f.dhn);
  s->cnt = reinterpret_cast<int *>(smem_raw + f.cnt);
  s->flag = reinterpret_cast<int *>(smem_raw + f.flag);
  s->wl = reinterpret_cast<uint16_t *>(smem_raw + f.wl);
  s->dbits = reinterpret_cast<uint32_t *>(smem_raw + f.db);
  s->qmask = reinterpret_cast<uint32_t *>(smem_raw + f.qm);
  s->chg = reinterpret_cast<uint16_t *>(smem_raw + f.chg);
  s->cidx = reinterpret_cast<uint16_t *>(smem_raw + f.cidx);
  s->lp = reinterpret_cast<uint16_t *>(smem_raw + f.lp);
  s->ld = reinterpret_cast<uint16_t *>(smem_raw + f.ld);
  s->wn = smem_raw + f.wn;
  s->type = smem_raw + f.type;
  s->phi = reinterpret_cast<float *>(smem_raw + f.phi);
  s->wcache = reinterpret_cast<float *>(smem_raw + f.wc);
  s->rxof = smem_raw + f.rxof;
  s->cpos = reinterpret_cast<float4 *>(smem_raw + f.cpos);
  s->wrange = smem_raw + f.wrange;
  s->wflag = smem_raw + f.wflag;
  s->ratio = reinterpret_cast<float *>(smem_raw + f.ratio);
  s->qlist = reinterpret_cast<uint16_t *>(smem_raw + f.qlist);
  s->ion_active = reinterpret_cast<uint16_t *>(smem_raw + f.ionA);
  s->ion_free = reinterpret_cast<uint16_t *>(smem_raw + f.ionF);
  s->ion_vel = reinterpret_cast<float4 *>(smem_raw + f.ionV);
}

// ---- kernels -----------------------------------------------------------------------------------
template <int ITEMS, bool FRAMES>
__global__ void __launch_bounds__(1024, 1)
k_md_run(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, int n_steps) {
  Ctx c{p, g};
  c.r = blockIdx.x;
  bind_smem(p, &c.s);
  init_ctx(c);
  Thr<ITEMS> t;
  bool f_valid = g.f_valid[c.r] != 0, lists_valid = false;
  unsigned long long counter = (unsigned long long)g.md_counter[c.r];
  load_state(c, t, f_valid);
  md_steps<ITEMS, FRAMES>(c, t, n_steps, counter, f_valid, lists_valid);
  store_state(c, t, true);
  if (c.tid == 0) {
    g.md_counter[c.r] = (long long)counter;
    g.md_steps[c.r] += n_steps;
    g.time[c.r] += (double)n_steps * p.dt_d;
    g.f_valid[c.r] = f_valid ? 1 : 0;
  }
}

// steepest descent (M.py:78-79): x += clamp(gamma f, +-max_disp), forces without thermostat
template <int ITEMS>
__global__ void __launch_bounds__(1024, 1)
k_steepest_descent(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, int n_steps,
                   float f_max, float sd_gamma, float max_disp) {
  Ctx c{p, g};
  c.r = blockIdx.x;
  bind_smem(p, &c.s);
  init_ctx(c);
  Thr<ITEMS> t;
  load_state(c, t, false);
  build_lists(c, t);
  if (p.use_dh) { build_qlist(c); __syncthreads(); }
  compute_forces(c, t, false, 0ull);
  for (int step = 0; step < n_steps; ++step) {
    int moved = 0, big = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int i = c.tid + k * c.nt;
      if (i < p.P) {
        float d2 = 0.f, n2 = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float dp = fminf(fmaxf(sd_gamma * t.f[k][a], -max_disp), max_disp);
          t.x[k][a] += dp;
          n2 = fmaf(t.f[k][a], t.f[k][a], n2);
          const float d = t.x[k][a] - t.ref[k][a];
          d2 = fmaf(d, d, d2);
        }
        float4 *sp = &c.s.pos[i];
        sp->x = t.x[k][0]; sp->y = t.x[k][1]; sp->z = t.x[k][2];
        if (t.ci[k] >= 0) { float4 *cp = &c.s.cpos[t.ci[k]]; cp->x = t.x[k][0]; cp->y = t.x[k][1]; cp->z = t.x[k][2]; }
        moved |= !(d2 <= p.half_skin2);
        big |= !(n2 < f_max * f_max);
      }
    }
    // __syncthreads_or is a vote (zero / non-zero), so the two questions need two barriers
    const int any_moved = __syncthreads_or(moved);
    if (!__syncthreads_or(big)) break;  // every |f| < f_max: converged (never with f_max = 0)
    if (any_moved) build_lists(c, t);
    compute_forces(c, t, false, 0ull);
  }
  store_state(c, t, false);
  if (c.tid == 0) g.f_valid[c.r] = 0;
}

template <int ITEMS, bool IONS>
__global__ void __launch_bounds__(1024, 1)
k_mc_run(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, int m, int n_trials,
         cgpe_trial_record *trace) {
  Ctx c{p, g};
  c.r = blockIdx.x;
  bind_smem(p, &c.s);
  init_ctx(c);
  Thr<ITEMS> t;
  load_state(c, t, true);
  build_lists(c, t);
  mc_prepare(c);
  unsigned long long trial_ctr = (unsigned long long)g.trials[(size_t)c.r * p.n_rx + m];
  long long accepted = 0;
  mc_block<IONS>(c, m, n_trials, trial_ctr, accepted, trace ? trace + (size_t)c.r * n_trials : nullptr);
  store_state(c, t, true);
  if (c.tid == 0) {
    g.trials[(size_t)c.r * p.n_rx + m] = (long long)trial_ctr;
    g.accepted[(size_t)c.r * p.n_rx + m] += accepted;
    if (n_trials > 0) g.f_valid[c.r] = 0;
  }
}

// forces exactly as the MD loop computes them, no thermostat (parity / debugging)
template <int ITEMS>
__global__ void __launch_bounds__(1024, 1)
k_md_forces(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, float4 *out) {
  Ctx c{p, g};
  c.r = blockIdx.x;
  bind_smem(p, &c.s);
  init_ctx(c);
  Thr<ITEMS> t;
  load_state(c, t, false);
  build_lists(c, t);
  if (p.use_dh) { build_qlist(c); __syncthreads(); }
  build_dh_compact(c);   // the partner lists the MD loop walks (md_steps)
  compute_forces(c, t, false, 0ull);
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    if (i < p.P) out[(size_t)c.r * p.P + i] = make_float4(t.f[k][0], t.f[k][1], t.f[k][2], 0.f);
  }
  if (c.err) atomicOr(&c.g.err[c.r], c.err);
}

// census of the Verlet lists for the current configuration (roofline arithmetic)
template <int ITEMS>
__global__ void __launch_bounds__(1024, 1)
k_list_stats(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, double *out) {
  Ctx c{p, g};
  c.r = blockIdx.x;
  bind_smem(p, &c.s);
  init_ctx(c);
  Thr<ITEMS> t;
  load_state(c, t, false);
  build_lists(c, t);
  double nw = 0, nw_in = 0, nd = 0, nd_in = 0;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = c.tid + k * c.nt;
    if (i < p.P && p.use_wca) {
      const int ti = c.s.type[i];
      const float4 pi4 = c.s.pos[i];
      for (int e = -4; e < (int)c.s.wn[i]; ++e) {
        int j = -1;
        if (e < 0) {   // the chain neighbours the bonded owner evaluates (directed count, like the rows)
          const int off = e < -2 ? e + 2 : e + 3;   // -2,-1,+1,+2
          if (t.l[k] >= 0 && t.l[k] + off >= 0 && t.l[k] + off < p.N) j = i + off;
        } else {
          j = c.s.wl[(size_t)e * p.P + i];
        }
        if (j < 0) continue;
        V3 d;
        const float r2 = dist2<true>(c, pi4, c.s.pos[j], &d);
        nw += 1;
        if (ti < p.T && r2 < c.s.tab[ti * (p.T + 1) + c.s.type[j]].z) nw_in += 1;
      }
    }
  }
  if (p.use_dh)
    for (int ci = c.tid; ci < p.n_chg; ci += c.nt) {
      const float4 pi4 = c.s.pos[c.s.chg[ci]];
      for (int w = 0; w < (p.n_chg + 31) >> 5; ++w) {
        uint32_t m = c.s.dbits[(size_t)ci * p.dw + w];
        nd += __popc(m);
        m &= c.s.qmask[w];
        while (m && pi4.w != 0.f) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          V3 d;
          if (dist2<true>(c, pi4, c.s.pos[c.s.chg[w * 32 + b]], &d) < c.rc2) nd_in += 1;
        }
      }
    }
  nw = warp_sum(nw); nw_in = warp_sum(nw_in); nd = warp_sum(nd); nd_in = warp_sum(nd_in);
  __syncthreads();
  if (c.lane == 0) { c.s.red[c.warp * 4] = nw; c.s.red[c.warp * 4 + 1] = nw_in; c.s.red[c.warp * 4 + 2] = nd; c.s.red[c.warp * 4 + 3] = nd_in; }
  __syncthreads();
  if (c.tid == 0) {
    double a[4] = {0, 0, 0, 0};
    for (int w = 0; w < (c.nt + 31) / 32; ++w) for (int q = 0; q < 4; ++q) a[q] += c.s.red[w * 4 + q];
    for (int q = 0; q < 4; ++q) out[c.r * 4 + q] = a[q];
  }
  if (c.err) atomicOr(&c.g.err[c.r], c.err);
}

// The sampling loop of perform_sampling (E.py:88-106), one launch for all samples.
template <int ITEMS, bool IONS, int MAXT, bool FRAMES>
__global__ void __launch_bounds__(MAXT, 1)
k_sample_loop(const __grid_constant__ DevParams p, const __grid_constant__ DevState g,
              const __grid_constant__ cgpe_sample_params sp, uint32_t thr1, uint32_t thr2, double *obs,
              float *qdist) {
  if (g.engine_choice && *g.engine_choice != 0) return;   // the plan of this launch gave it to the cluster engine
  Ctx c{p, g};
  c.r = g.order ? g.order[blockIdx.x] : (int)blockIdx.x;   // more replicas than SMs: the most expensive start first (duo_plan_waves)
  bind_smem(p, &c.s);
  init_ctx(c);
  Thr<ITEMS> t;
  if (c.tid == 0) { g.cta_ns[c.r * 4] = (long long)global_timer_ns(); g.cta_ns[c.r * 4 + 2] = 0; g.cta_ns[c.r * 4 + 3] = 0; }
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
  if (!sp.use_md) { build_lists(c, t); lists_valid = true; __syncthreads(); }
  const double t0 = g.time[c.r];
  long long steps = 0;
  const uint32_t k0 = (uint32_t)sp.schedule_seed, k1 = (uint32_t)(p.replica_offset + c.r);
  for (int s = 0; s < sp.n_samples; ++s) {
    // every thread draws the same schedule numbers: cheaper than a broadcast + barrier
    const uint4 o = philox4x32_10((uint32_t)sample_ctr, (uint32_t)(sample_ctr >> 32), 0u, kTagSchedule, k0, k1);
    sample_ctr += 1;
    if (sp.use_md) {
      const int n = ((o.x >> 8) < thr1 ? sp.md_steps_per_burst : 0) +   // E.py:89-90
                    ((o.y >> 8) < thr2 ? sp.md_steps_per_burst : 0);    // E.py:91-92
      if (n > 0) {
        if (!lists_valid) { build_lists(c, t); lists_valid = true; }
        md_steps<ITEMS, FRAMES>(c, t, n, counter, f_valid, lists_valid);
        steps += n;
        __syncthreads();  // list rows and positions visible to the MC warp
      } else if (!lists_valid) {
        build_lists(c, t); lists_valid = true;
        __syncthreads();
      }
    }
    if (sp.n_mc_blocks > 0) mc_prepare(c);      // per-site caches for this sample's trial moves
    for (int b = 0; b < sp.n_mc_blocks; ++b) {   // E.py:94-95
      const int m = sp.mc_reaction[b];
      mc_block<IONS>(c, m, sp.mc_trials[b], trial_ctr[m], accepted[m], nullptr);
      if (sp.mc_trials[b] > 0) f_valid = false;
    }
    if (IONS && c.s.flag[3]) lists_valid = false;   // ions were inserted / taken away: rows are stale
    double rg, ree;
    chain_observables(c, t, &rg, &ree);   // E.py:100-101
    if (c.tid == 0) {
      double *o6 = obs + ((size_t)c.r * sp.n_samples + s) * CGPE_N_OBS;
      int n_ion = 0;
      for (int m = 0; m < p.n_rx; ++m) n_ion += c.s.cnt[m * 2 + 1];
      o6[0] = p.n_rx > 0 ? c.s.cnt[1] : 0;   // E.py:97
      o6[1] = p.n_rx > 1 ? c.s.cnt[3] : 0;   // E.py:98
      o6[2] = IONS ? c.s.flag[1] : n_ion;   // E.py:99 (implicit ions: N_B = N_N + N_N2)
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
  store_state(c, t, true);
  if (c.tid == 0) {
    g.md_counter[c.r] = (long long)counter;
    g.md_steps[c.r] += steps;
    g.samples_done[c.r] = (long long)sample_ctr;
    g.cta_ns[c.r * 4 + 1] = (long long)global_timer_ns();
    g.time[c.r] = t0 + (double)steps * p.dt_d;
    g.f_valid[c.r] = f_valid ? 1 : 0;
    for (int m = 0; m < p.n_rx; ++m) {
      g.trials[(size_t)c.r * p.n_rx + m] = (long long)trial_ctr[m];
      g.accepted[(size_t)c.r * p.n_rx + m] += accepted[m];
    }
  }
}

template <typename K>
cudaError_t opt_in_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int chain_items(int P) {
  int items = (P + 1023) / 1024;
  if (const char *e = std::getenv("CGPE_ITEMS")) items = std::max(items, std::atoi(e));   // diagnostics: particles per thread
  return items;
}

ChainLaunch chain_launch_shape(const DevParams &p) {
  ChainLaunch L{};
  L.items = chain_items(p.P);
  if (L.items > 2) { L.items = 0; return L; }   // beyond the shared-memory engine
  const int per = (p.P + L.items - 1) / L.items;
  L.threads = ((per + 31) / 32) * 32;
  if (L.threads < 32) L.threads = 32;
  L.smem = (size_t)p.off.total;
  return L;
}

#define CGPE_DISPATCH(items, expr1, expr2) ((items) == 1 ? (expr1) : (expr2))

cudaError_t launch_md_run(const DevParams &p, const DevState &g, const ChainLaunch &L, int n_steps, cudaStream_t st) {
  auto k = p.frame_every > 0 ? (L.items == 1 ? k_md_run<1, true> : k_md_run<2, true>) : (L.items == 1 ? k_md_run<1, false> : k_md_run<2, false>);
  cudaError_t e = opt_in_smem(k, L.smem);
  if (e != cudaSuccess) return e;
  k<<<p.R, L.threads, L.smem, st>>>(p, g, n_steps);
  return cudaGetLastError();
}

cudaError_t launch_steepest_descent(const DevParams &p, const DevState &g, const ChainLaunch &L, int n_steps,
                                    float f_max, float gamma, float max_disp, cudaStream_t st) {
  auto k = L.items == 1 ? k_steepest_descent<1> : k_steepest_descent<2>;
  cudaError_t e = opt_in_smem(k, L.smem);
  if (e != cudaSuccess) return e;
  k<<<p.R, L.threads, L.smem, st>>>(p, g, n_steps, f_max, gamma, max_disp);
  return cudaGetLastError();
}

cudaError_t launch_mc_run(const DevParams &p, const DevState &g, const ChainLaunch &L, int m, int n_trials,
                          cgpe_trial_record *trace, cudaStream_t st) {
  auto k = p.n_ions > 0 ? (L.items == 1 ? k_mc_run<1, true> : k_mc_run<2, true>) : (L.items == 1 ? k_mc_run<1, false> : k_mc_run<2, false>);
  cudaError_t e = opt_in_smem(k, L.smem);
  if (e != cudaSuccess) return e;
  k<<<p.R, L.threads, L.smem, st>>>(p, g, m, n_trials, trace);
  return cudaGetLastError();
}

cudaError_t launch_md_forces(const DevParams &p, const DevState &g, const ChainLaunch &L, float4 *out, cudaStream_t st) {
  auto k = L.items == 1 ? k_md_forces<1> : k_md_forces<2>;
  cudaError_t e = opt_in_smem(k, L.smem);
  if (e != cudaSuccess) return e;
  k<<<p.R, L.threads, L.smem, st>>>(p, g, out);
  return cudaGetLastError();
}

cudaError_t launch_list_stats(const DevParams &p, const DevState &g, const ChainLaunch &L, double *out, cudaStream_t st) {
  auto k = L.items == 1 ? k_list_stats<1> : k_list_stats<2>;
  cudaError_t e = opt_in_smem(k, L.smem);
  if (e != cudaSuccess) return e;
  k<<<p.R, L.threads, L.smem, st>>>(p, g, out);
  return cudaGetLastError();
}

cudaError_t launch_sample_loop(const DevParams &p, const DevState &g, const ChainLaunch &L,
                               const cgpe_sample_params &sp, uint32_t thr1, uint32_t thr2, double *obs,
                               float *qdist, cudaStream_t st) {
  auto k = p.n_ions > 0 ? (L.items == 1 ? k_sample_loop<1, true, 1024, false> : k_sample_loop<2, true, 1024, false>)
                        : (L.items == 1 ? k_sample_loop<1, false, 1024, false> : L.threads <= 512 ? k_sample_loop<2, false, 512, false> : k_sample_loop<2, false, 1024, false>);
  if (p.frame_every > 0)   // the frame recorder armed (cgpe_frames_begin): its own instantiations, so that the common path carries none of it
    k = p.n_ions > 0 ? (L.items == 1 ? k_sample_loop<1, true, 1024, true> : k_sample_loop<2, true, 1024, true>)
                     : (L.items == 1 ? k_sample_loop<1, false, 1024, true> : L.threads <= 512 ? k_sample_loop<2, false, 512, true> : k_sample_loop<2, false, 1024, true>);
  cudaError_t e = opt_in_smem(k, L.smem);
  if (e != cudaSuccess) return e;
  k<<<p.R, L.threads, L.smem, st>>>(p, g, sp, thr1, thr2, obs, qdist);
  return cudaGetLastError();
}

}  // namespace cgpe
