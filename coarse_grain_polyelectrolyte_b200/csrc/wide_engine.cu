// wide_engine.cu — the same hot path for replicas that do not fit one CTA's shared memory
// (P > 2048 particles, e.g. the N = 10^4 chains of the metric): many CTAs per replica, state and
// Verlet lists in global memory (L2-resident: 160 KB of positions per N = 10^4 replica), one
// thread per particle.
//
//   MD step   = wk_drift (half kick + drift + skin/2 vote)  ->  wk_rebuild (only replicas that
//               voted)  ->  wk_forces (all terms + thermostat + second half kick)
//   MC block  = wk_mc_prepare (per-site potentials, all threads) -> wk_mc_block (one warp per
//               replica, same trial logic and Philox streams as the shared-memory engine)
//   sampling  = the loop of perform_sampling driven from the host, one small read-back per sample
//
// Physics, Philox streams and bookkeeping are those of chain_engine.cu, so both engines follow
// the same Markov chain and are checked against the same oracle; only the summation order of the
// forces differs.  Reference call sites: see chain_engine.cu.
#include "wide_engine.cuh"

#include <cstring>

namespace cgpe {
namespace {

constexpr int kWT = 256;   // threads per CTA
constexpr int kStepFromDevice = -1000;   // kernels of a captured step read the step index from WideState::step
constexpr int kGraphChunk = 25;          // MD steps per captured graph

__device__ __forceinline__ float wdist2(const DevParams &p, float4 a, float4 b, V3 *d) {
  d->x = min_image(a.x - b.x, p.box_l, p.inv_box_l);
  d->y = min_image(a.y - b.y, p.box_l, p.inv_box_l);
  d->z = min_image(a.z - b.z, p.box_l, p.inv_box_l);
  return fmaf(d->x, d->x, fmaf(d->y, d->y, d->z * d->z));
}

__device__ __forceinline__ int chain_index(const DevParams &p, int i) {
  return i < p.NC ? (p.NC == p.N ? i : i % p.N) : -1;
}

// ---- MD ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWT)
wk_drift(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w,
         int step, float sd_gamma, float sd_max_disp) {
  const int r = blockIdx.y, i = blockIdx.x * kWT + threadIdx.x;
  if (step == kStepFromDevice) step = w.step[0];
  if (step >= w.nsteps[r] || i >= p.P) return;
  const size_t k = (size_t)r * p.P + i;
  float4 x = g.pos[k];
  const float4 f = g.frc[k], ref = w.ref[k];
  if (sd_gamma > 0.f) {   // steepest descent: x += clamp(gamma f, +-max_disp)   (M.py:78-79)
    x.x += fminf(fmaxf(sd_gamma * f.x, -sd_max_disp), sd_max_disp);
    x.y += fminf(fmaxf(sd_gamma * f.y, -sd_max_disp), sd_max_disp);
    x.z += fminf(fmaxf(sd_gamma * f.z, -sd_max_disp), sd_max_disp);
  } else {
    float4 v = g.vel[k];
    v.x = fmaf(p.half_dt, f.x, v.x); v.y = fmaf(p.half_dt, f.y, v.y); v.z = fmaf(p.half_dt, f.z, v.z);
    x.x = fmaf(p.dt, v.x, x.x); x.y = fmaf(p.dt, v.y, x.y); x.z = fmaf(p.dt, v.z, x.z);
    g.vel[k] = v;
  }
  if (i >= p.NC) {   // free particles (explicit ions) wander: keep them inside the box
    x.x -= p.box_l * floorf(x.x * p.inv_box_l); x.y -= p.box_l * floorf(x.y * p.inv_box_l); x.z -= p.box_l * floorf(x.z * p.inv_box_l);
  }
  g.pos[k] = x;
  V3 dd;
  if (!(wdist2(p, x, ref, &dd) <= p.half_skin2)) w.flag[r] = 1;   // benign race: all writers store 1 (minimum image: ions are folded)
}

__global__ void __launch_bounds__(kWT)
wk_rebuild(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w) {
  const int r = blockIdx.y, i = blockIdx.x * kWT + threadIdx.x;
  if (r == 0 && i == 0) w.step[1] = w.step[0];   // the force kernels of this step read step[1]
  if (!w.flag[r]) return;
  __shared__ float4 tile[kWT];
  const size_t base = (size_t)r * p.P;
  const bool exists = i < p.P;
  const bool own = exists && g.type[base + i] < p.T;   // hidden slots build no rows
  const float4 pi = exists ? g.pos[base + i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const int l = own ? chain_index(p, i) : -1, ci = own ? w.cidx[base + i] : -1;
  int ex_lo = i, ex_hi = i;
  if (l >= 0) { ex_lo = i - min(2, l); ex_hi = i + min(2, p.N - 1 - l); }
  const float rc = g.rcut[r], rlist_d2 = (rc + p.skin) * (rc + p.skin);
  int nw = 0, nd = 0;
  uint32_t *wl = w.wl + (size_t)r * w.cap_w * p.P;
  uint32_t *dl = w.dl + ((size_t)r * p.n_chg + (ci >= 0 ? ci : 0)) * w.cap_d;   // this site's row
  for (int j0 = 0; j0 < p.P; j0 += kWT) {
    __syncthreads();
    if (j0 + threadIdx.x < p.P) tile[threadIdx.x] = g.pos[base + j0 + threadIdx.x];
    __syncthreads();
    if (!own) continue;
    const int jn = min(kWT, p.P - j0);
    for (int jj = 0; jj < jn; ++jj) {
      const int j = j0 + jj;
      V3 d;
      const float r2 = wdist2(p, pi, tile[jj], &d);
      if (p.use_wca && r2 < p.rlist_w2 && (j < ex_lo || j > ex_hi)) {
        if (nw < w.cap_w) wl[(size_t)nw * p.P + i] = (uint32_t)j;
        ++nw;
      }
      if (p.use_dh && ci >= 0 && r2 < rlist_d2 && j != i) {
        const int cj = w.cidx[base + j];
        if (cj >= 0) {
          if (nd < w.cap_d) dl[nd] = (uint32_t)cj;
          ++nd;
        }
      }
    }
  }
  if (!exists) return;
  if (nw > w.cap_w) { atomicOr(&g.err[r], kErrWcaCap); nw = w.cap_w; }
  if (nd > w.cap_d) { atomicOr(&g.err[r], kErrDhCap); nd = w.cap_d; }
  w.wn[base + i] = nw;
  if (ci >= 0) w.dn[(size_t)r * p.n_chg + ci] = nd;
  w.ref[base + i] = pi;
  if (i == 0) g.rebuilds[r] += 1;
}

// Debye-Hueckel force (mode 0) or potential sum_j q_j e(r) (mode 1) of every chargeable particle:
// one warp per site, lanes stride over its contiguous row, warp-shuffle reduction.
__global__ void __launch_bounds__(kWT)
wk_dh_rows(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w,
           int step, int mode) {
  const int r = blockIdx.y, lane = threadIdx.x & 31, ci = blockIdx.x * (kWT / 32) + (threadIdx.x >> 5);
  if (step == kStepFromDevice) step = w.step[1];
  if (mode == 0 && step >= w.nsteps[r]) return;
  if (ci >= p.n_chg) return;
  const size_t base = (size_t)r * p.P, cb = (size_t)r * p.n_chg;
  const int *chg = g.chg + cb;
  const float4 pi4 = g.pos[base + chg[ci]];
  float fx = 0.f, fy = 0.f, fz = 0.f, ph = 0.f;
  if (mode == 1 || pi4.w != 0.f) {
    const float kappa = g.kappa[r], rc = g.rcut[r], rc2 = rc * rc, nkl = -kappa * kLog2e;
    const uint32_t *row = w.dl + (cb + ci) * w.cap_d;
    const int n = w.dn[cb + ci];
    for (int e = lane; e < n; e += 32) {
      const float4 pj4 = g.pos[base + chg[row[e]]];
      if (pj4.w == 0.f) continue;
      V3 d;
      const float r2 = wdist2(p, pi4, pj4, &d);
      if (r2 < rc2) {
        float fr;
        const float en = dh_energy_unit(r2, kappa, nkl, &fr);
        ph = fmaf(pj4.w, en, ph);
        fr *= pj4.w;
        fx = fmaf(fr, d.x, fx); fy = fmaf(fr, d.y, fy); fz = fmaf(fr, d.z, fz);
      }
    }
  }
  if (mode == 1) {
    ph = warp_sum(ph);
    if (lane == 0) w.phi[cb + ci] = ph;
  } else {
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    const float s = p.dh_pref * pi4.w;
    if (lane == 0) w.fdh[cb + ci] = make_float4(s * fx, s * fy, s * fz, 0.f);
  }
}

// mode: 0 = forces only (initial evaluation / md_forces / steepest descent), 1 = forces + second half kick
__global__ void __launch_bounds__(kWT)
wk_forces(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w,
          int step, int mode, int thermostat, float4 *out) {
  const int r = blockIdx.y, i = blockIdx.x * kWT + threadIdx.x;
  if (step == kStepFromDevice) {
    step = w.step[1];
    if (r == 0 && i == 0) w.step[0] = step + 1;   // nothing in this launch reads step[0]; the next drift does
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) w.flag[r] = 0;   // every block of wk_rebuild has seen it
  if (step >= w.nsteps[r]) return;
  if (mode == 0 && step < 0 && g.f_valid[r] && !out) return;   // cached forces are still good
  if (i >= p.P) return;
  const size_t base = (size_t)r * p.P;
  const float4 *pos = g.pos + base;
  const uint8_t *type = g.type + base;
  const float4 pi4 = pos[i];
  const V3 pi = xyz(pi4);
  const int ti = type[i], S = p.T + 1, l = chain_index(p, i), N = p.N;
  int err = 0;
  V3 f = {0.f, 0.f, 0.f};
  if (l >= 0) {
    V3 f1, f2, f3, f4;
    // bonded terms this bead takes part in (M.py:159-170)
    if (p.bond_kind != CGPE_BOND_NONE) {
      if (l >= 1) { bond_term(p, pi, xyz(pos[i - 1]), &f1, &err); f = f + f1; }
      if (l <= N - 2) { bond_term(p, pi, xyz(pos[i + 1]), &f1, &err); f = f + f1; }
    }
    if (p.use_angle) {
      if (l >= 2) { angle_term(p, xyz(pos[i - 2]), xyz(pos[i - 1]), pi, &f1, &f2, &f3); f = f + f3; }
      if (l >= 1 && l <= N - 2) { angle_term(p, xyz(pos[i - 1]), pi, xyz(pos[i + 1]), &f1, &f2, &f3); f = f + f2; }
      if (l <= N - 3) { angle_term(p, pi, xyz(pos[i + 1]), xyz(pos[i + 2]), &f1, &f2, &f3); f = f + f1; }
    }
    if (p.use_dih)
      for (int role = 0; role < 4; ++role) {
        const int o = l + 1 - role;   // owner bead of the quadruple in which this bead is p1..p4
        if (o >= 1 && o <= N - 3) {
          const int b = i + 1 - role;
          dihedral_term(p, xyz(pos[b - 1]), xyz(pos[b]), xyz(pos[b + 1]), xyz(pos[b + 2]), &f1, &f2, &f3, &f4);
          f = f + (role == 0 ? f1 : role == 1 ? f2 : role == 2 ? f3 : f4);
        }
      }
    // WCA with the chain neighbours i+-1, i+-2 (not in the rows)
    if (p.use_wca && ti < p.T)
      for (int off = -2; off <= 2; ++off) {
        if (off == 0 || l + off < 0 || l + off >= N) continue;
        const int j = i + off, tj = type[j];
        const float4 tab = g.wca_tab[ti * S + tj];
        const V3 d = pi - xyz(pos[j]);
        const float r2 = dot(d, d);
        if (tj < p.T && r2 < tab.z) f = f + wca_force_over_r(tab, r2) * d;
      }
  }
  if (p.use_wca && ti < p.T) {
    const uint32_t *wl = w.wl + (size_t)r * w.cap_w * p.P;
    const int n = w.wn[base + i];
    for (int e = 0; e < n; ++e) {
      const int j = wl[(size_t)e * p.P + i];
      const float4 tab = g.wca_tab[ti * S + type[j]];
      V3 d;
      const float r2 = wdist2(p, pi4, pos[j], &d);
      if (r2 < tab.z) f = f + wca_force_over_r(tab, r2) * d;
    }
  }
  const int ci = w.cidx[base + i];
  if (p.use_dh && ci >= 0 && pi4.w != 0.f) {   // summed by wk_dh_forces, one warp per site
    const float4 fd = w.fdh[(size_t)r * p.n_chg + ci];
    f.x += fd.x; f.y += fd.y; f.z += fd.z;
  }
  float4 v = g.vel[base + i];
  if (thermostat && ti < p.T) {   // E.py:61-63
    const float gamma = g.gamma[r], pref = sqrtf(24.0f * gamma * g.kT[r] / p.dt);
    const unsigned long long cnt = (unsigned long long)g.md_counter[r] + (unsigned long long)(step + 1);
    const uint4 o = philox4x32_10((uint32_t)cnt, (uint32_t)(cnt >> 32), (uint32_t)i, kTagLangevin, p.thermostat_seed,
                                  (uint32_t)(p.replica_offset + r));
    f.x += fmaf(-gamma, v.x, pref * centred_u23(o.x));
    f.y += fmaf(-gamma, v.y, pref * centred_u23(o.y));
    f.z += fmaf(-gamma, v.z, pref * centred_u23(o.z));
  }
  if (p.force_cap > 0.f) {
    const float n2 = dot(f, f);
    if (n2 > p.force_cap * p.force_cap) f = (p.force_cap * rsqrtf(n2)) * f;
  }
  if (out) { out[base + i] = make_float4(f.x, f.y, f.z, 0.f); }
  else {
    g.frc[base + i] = make_float4(f.x, f.y, f.z, 0.f);
    if (mode == 1) {
      v.x = fmaf(p.half_dt, f.x, v.x); v.y = fmaf(p.half_dt, f.y, v.y); v.z = fmaf(p.half_dt, f.z, v.z);
      g.vel[base + i] = v;
      if (!isfinite(v.x + v.y + v.z + pi4.x + pi4.y + pi4.z)) err |= kErrNonFinite;
      else if (dot(V3{v.x, v.y, v.z}, V3{v.x, v.y, v.z}) > 1.0e6f * fmaxf(g.kT[r], 1.0f)) err |= kErrDiverged;
    }
  }
  if (err) atomicOr(&g.err[r], err);
}

// per-replica bookkeeping after a burst of n steps (n from nsteps[r] when n_uniform < 0)
__global__ void wk_finish_md(const __grid_constant__ DevParams p, const __grid_constant__ DevState g,
                             const __grid_constant__ WideState w, int n_uniform, int is_md) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  const int n = n_uniform >= 0 ? n_uniform : w.nsteps[r];
  if (is_md) {
    if (n > 0) { g.md_counter[r] += n; g.md_steps[r] += n; g.time[r] += (double)n * p.dt_d; g.f_valid[r] = 1; }
  } else {
    g.f_valid[r] = 0;
  }
}

__global__ void wk_set_int(int *a, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// ---- MC ----------------------------------------------------------------------------------------
__device__ float wide_wca_swap(const DevParams &p, const DevState &g, const WideState &w, int r, int ci, bool *tit) {
  const int m = w.rxof[(size_t)r * p.n_chg + ci];
  if (tit) *tit = false;
  if (m == 255 || !p.use_wca) return 0.f;
  const size_t base = (size_t)r * p.P;
  const int S = p.T + 1, i = g.chg[(size_t)r * p.n_chg + ci];
  const int tp = p.rx[m].type_prot, td = p.rx[m].type_deprot, l = chain_index(p, i);
  const float4 pi4 = g.pos[base + i];
  const uint32_t *wl = w.wl + (size_t)r * w.cap_w * p.P;
  const int n_w = w.wn[base + i];
  float e = 0.f;
  for (int k = 0; k < n_w + 4; ++k) {
    int j = -1;
    if (k < 4) {
      const int off = k < 2 ? k - 2 : k - 1;
      if (l >= 0 && l + off >= 0 && l + off < p.N) j = i + off;
    } else {
      j = wl[(size_t)(k - 4) * p.P + i];
    }
    if (j < 0) continue;
    const int tj = g.type[base + j];
    if (tj >= p.T) continue;
    V3 d;
    const float r2 = wdist2(p, pi4, g.pos[base + j], &d);
    if (tit && r2 < p.rc_w_max2) {
      const int cj = w.cidx[base + j];
      if (cj >= 0 && w.rxof[(size_t)r * p.n_chg + cj] != 255) *tit = true;
    }
    e += wca_energy(g.wca_tab[td * S + tj], r2) - wca_energy(g.wca_tab[tp * S + tj], r2);
  }
  return e;
}

__global__ void __launch_bounds__(kWT)
wk_mc_classify(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w) {
  const int r = blockIdx.y, ci = blockIdx.x * kWT + threadIdx.x;
  if (ci >= p.n_chg) return;
  const int t = g.type[(size_t)r * p.P + g.chg[(size_t)r * p.n_chg + ci]];
  int m = 255;
  for (int k = 0; k < p.n_rx; ++k)
    if (t == p.rx[k].type_prot || t == p.rx[k].type_deprot) m = k;
  w.rxof[(size_t)r * p.n_chg + ci] = (uint8_t)m;
}

__global__ void __launch_bounds__(kWT)
wk_mc_prepare(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w) {
  const int r = blockIdx.y, ci = blockIdx.x * kWT + threadIdx.x;
  if (ci >= p.n_chg) return;
  const size_t cb = (size_t)r * p.n_chg;
  if (!p.use_dh) w.phi[cb + ci] = 0.f;   // otherwise filled by wk_dh_rows(mode 1)
  bool tit;
  w.wcache[cb + ci] = wide_wca_swap(p, g, w, r, ci, &tit);
  w.wflag[cb + ci] = tit ? 1 : 0;
}

__device__ __forceinline__ float wwarp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// one warp per replica: the trial loop of chain_engine.cu::mc_block on global-memory caches,
// including the explicit-ion insertion / deletion (p.n_ions > 0)
__global__ void __launch_bounds__(32)
wk_mc_block(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w,
            int m, int n_trials, cgpe_trial_record *trace_all) {
  const int r = blockIdx.x, lane = threadIdx.x;
  const DevReaction rx = p.rx[m];
  const bool ions = p.n_ions > 0;
  const size_t base = (size_t)r * p.P, cb = (size_t)r * p.n_chg;
  int *lp = g.lst_prot + ((size_t)r * p.n_rx + m) * p.P, *ld = g.lst_deprot + ((size_t)r * p.n_rx + m) * p.P;
  int *ion_active = g.ion_active + (size_t)r * (ions ? p.n_ions : 1), *ion_free = g.ion_free + (size_t)r * (ions ? p.n_ions : 1);
  int n_p = g.cnt[((size_t)r * p.n_rx + m) * 2], n_d = g.cnt[((size_t)r * p.n_rx + m) * 2 + 1];
  int n_act = ions ? g.ion_cnt[r * 2] : 0, n_free = ions ? g.ion_cnt[r * 2 + 1] : 0;
  const int n_sites = n_p + n_d;
  const float arg_fwd = (float)(2.302585092994045684 * (g.pH[r] - rx.pK)) * kLog2e;
  const float neg_beta_log2e = -rx.inv_kT * kLog2e;
  const float dq_fwd = rx.q_deprot - rx.q_prot, coul_fwd = p.dh_pref * dq_fwd;
  const float kappa = g.kappa[r], rc = g.rcut[r], rc2 = rc * rc, nkl = -kappa * kLog2e;
  const uint32_t k0 = rx.seed, k1 = (uint32_t)(p.replica_offset + r);
  const unsigned long long trial_ctr = (unsigned long long)g.trials[(size_t)r * p.n_rx + m];
  cgpe_trial_record *trace = trace_all ? trace_all + (size_t)r * n_trials : nullptr;
  const int *chg = g.chg + cb;
  long long acc = 0;
  bool ion_moved = false;
  int err = 0;
  for (int b0 = 0; b0 < n_trials; b0 += 32) {
    const unsigned long long cnt = trial_ctr + (unsigned long long)(b0 + lane);
    const uint4 o = philox4x32_10((uint32_t)cnt, (uint32_t)(cnt >> 32), 0u, kTagMc + (uint32_t)m, k0, k1);
    uint4 o1 = make_uint4(0, 0, 0, 0), o2 = o1;
    if (ions) {
      o1 = philox4x32_10((uint32_t)cnt, (uint32_t)(cnt >> 32), 1u, kTagMc + (uint32_t)m, k0, k1);
      o2 = philox4x32_10((uint32_t)cnt, (uint32_t)(cnt >> 32), 2u, kTagMc + (uint32_t)m, k0, k1);
    }
    const int nb = min(32, n_trials - b0);
    for (int tt = 0; tt < nb; ++tt) {
      const uint32_t a0 = __shfl_sync(0xffffffffu, o.x, tt), a1 = __shfl_sync(0xffffffffu, o.y, tt),
                     a2 = __shfl_sync(0xffffffffu, o.z, tt);
      const bool fwd = (a0 >> 31) == 0;
      const int n_src = fwd ? n_p : n_d, n_dst = fwd ? n_d : n_p;
      int *src = fwd ? lp : ld, *dst = fwd ? ld : lp;
      const float u = u24(a2);
      bool possible = n_src > 0;
      if (ions && !fwd && n_act == 0) possible = false;
      if (ions && fwd && n_free == 0) { possible = false; err |= kErrIonPool; }
      if (!possible) {
        if (trace && lane == 0) { cgpe_trial_record rec = {-1, -1, (int8_t)(fwd ? 1 : -1), 0, 0, 0, 0.f, 0.f, u}; trace[b0 + tt] = rec; }
        continue;
      }
      const int k = (int)__umulhi(a1, (uint32_t)n_src);
      const int i = src[k], ci = w.cidx[base + i];
      const int t_new = fwd ? rx.type_deprot : rx.type_prot;
      const float q_new = fwd ? rx.q_deprot : rx.q_prot, dq = fwd ? dq_fwd : -dq_fwd;
      float de = w.wcache[cb + ci];
      if (p.use_dh) de = fmaf(coul_fwd, w.phi[cb + ci], de);
      if (!fwd) de = -de;
      int b = -1, kb = 0;
      bool excluded = false;
      float4 pb4 = make_float4(0.f, 0.f, 0.f, rx.q_ion);
      if (ions) {   // see chain_engine.cu::mc_block
        if (fwd) {
          b = ion_free[n_free - 1];
          pb4.x = u24(__shfl_sync(0xffffffffu, o1.x, tt)) * p.box_l;
          pb4.y = u24(__shfl_sync(0xffffffffu, o1.y, tt)) * p.box_l;
          pb4.z = u24(__shfl_sync(0xffffffffu, o1.z, tt)) * p.box_l;
        } else {
          kb = (int)__umulhi(__shfl_sync(0xffffffffu, o.w, tt), (uint32_t)n_act);
          b = ion_active[kb];
          const float4 x = g.pos[base + b];
          pb4.x = x.x; pb4.y = x.y; pb4.z = x.z;
        }
        float esum = 0.f, dmin2 = 3.0e38f;
        for (int j = lane; j < p.P; j += 32) {
          if (j == b) continue;
          const int tj = j == i ? t_new : (int)g.type[base + j];
          if (tj >= p.T) continue;
          const float4 pj4 = g.pos[base + j];
          const float qj = j == i ? q_new : pj4.w;
          V3 d;
          const float r2 = wdist2(p, pb4, pj4, &d);
          dmin2 = fminf(dmin2, r2);
          if (p.use_dh && qj != 0.f && r2 < rc2) {
            float fr;
            esum = fmaf(qj, dh_energy_unit(r2, kappa, nkl, &fr), esum);
          }
        }
        esum = warp_sum(esum);
        dmin2 = wwarp_min(dmin2);
        const float e_ion = p.dh_pref * rx.q_ion * esum;
        de += fwd ? e_ion : -e_ion;
        excluded = dmin2 < rx.excl2;
      }
      const float bf = excluded ? 0.f : ((float)n_src / (float)(n_sites - n_src + 1)) * exp2f(fmaf(neg_beta_log2e, de, fwd ? arg_fwd : -arg_fwd));
      const bool accept = u < bf;
      if (trace && lane == 0) { cgpe_trial_record rec = {i, b, (int8_t)(fwd ? 1 : -1), (int8_t)accept, (int8_t)excluded, 0, de, bf, u}; trace[b0 + tt] = rec; }
      if (accept) {
        const float4 pi4 = g.pos[base + i];
        const bool refresh = p.use_wca && w.wflag[cb + ci] != 0;
        if (lane == 0) {
          g.type[base + i] = (uint8_t)t_new;
          g.pos[base + i].w = q_new;
          src[k] = src[n_src - 1];
          dst[n_dst] = i;
        }
        if (fwd) { n_p -= 1; n_d += 1; } else { n_d -= 1; n_p += 1; }
        acc += 1;
        if (p.use_dh && dq != 0.f) {
          const int n = w.dn[cb + ci];
          const uint32_t *row = w.dl + (cb + ci) * w.cap_d;
          for (int e = lane; e < n; e += 32) {
            const int ck = row[e];
            V3 d;
            const float r2 = wdist2(p, pi4, g.pos[base + chg[ck]], &d);
            if (r2 < rc2) {
              float fr;
              w.phi[cb + ck] = fmaf(dq, dh_energy_unit(r2, kappa, nkl, &fr), w.phi[cb + ck]);
            }
          }
        }
        __syncwarp();
        if (ions) {
          const int cbi = w.cidx[base + b];
          const float dqb = fwd ? rx.q_ion : -rx.q_ion;
          const uint32_t w0 = __shfl_sync(0xffffffffu, o2.x, tt), w1 = __shfl_sync(0xffffffffu, o2.y, tt),
                         w2 = __shfl_sync(0xffffffffu, o2.z, tt), w3 = __shfl_sync(0xffffffffu, o2.w, tt);
          if (lane == 0) {
            if (fwd) {
              g.type[base + b] = (uint8_t)rx.type_ion;
              g.pos[base + b] = pb4;
              const float ua = ((float)(w0 >> 8) + 0.5f) * (1.0f / 16777216.0f), ub = ((float)(w1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
              const float uc = ((float)(w2 >> 8) + 0.5f) * (1.0f / 16777216.0f), ud = ((float)(w3 >> 8) + 0.5f) * (1.0f / 16777216.0f);
              const float ra = rx.sqrt_kT * sqrtf(-2.0f * logf(ua)), rb = rx.sqrt_kT * sqrtf(-2.0f * logf(uc));
              float sa, ca, sb_, cb_;
              sincosf(6.283185307179586f * ub, &sa, &ca);
              sincosf(6.283185307179586f * ud, &sb_, &cb_);
              g.vel[base + b] = make_float4(ra * ca, ra * sa, rb * cb_, 0.f);
              ion_active[n_act] = b;
            } else {
              g.type[base + b] = (uint8_t)p.T;
              g.pos[base + b].w = 0.f;
              g.vel[base + b] = make_float4(0.f, 0.f, 0.f, 0.f);
              ion_active[kb] = ion_active[n_act - 1];
              ion_free[n_free] = b;
            }
          }
          if (fwd) { n_free -= 1; n_act += 1; } else { n_act -= 1; n_free += 1; }
          if (p.use_dh)
            for (int ck = lane; ck < p.n_chg; ck += 32) {
              if (ck == cbi) continue;
              V3 d;
              const float r2 = wdist2(p, pb4, g.pos[base + chg[ck]], &d);
              if (r2 < rc2) {
                float fr;
                w.phi[cb + ck] = fmaf(dqb, dh_energy_unit(r2, kappa, nkl, &fr), w.phi[cb + ck]);
              }
            }
          ion_moved = true;
          __syncwarp();
        }
        if (refresh) {
          const int l = chain_index(p, i), n_w = w.wn[base + i];
          const uint32_t *wl = w.wl + (size_t)r * w.cap_w * p.P;
          for (int e = lane; e < n_w + 4; e += 32) {
            int j = -1;
            if (e < 4) {
              const int off = e < 2 ? e - 2 : e - 1;
              if (l >= 0 && l + off >= 0 && l + off < p.N) j = i + off;
            } else {
              j = wl[(size_t)(e - 4) * p.P + i];
            }
            if (j >= 0) {
              const int cj = w.cidx[base + j];
              if (cj >= 0 && w.rxof[cb + cj] != 255) w.wcache[cb + cj] = wide_wca_swap(p, g, w, r, cj, nullptr);
            }
          }
          __syncwarp();
        }
      }
    }
  }
  if (lane == 0) {
    g.cnt[((size_t)r * p.n_rx + m) * 2] = n_p; g.cnt[((size_t)r * p.n_rx + m) * 2 + 1] = n_d;
    g.trials[(size_t)r * p.n_rx + m] = (long long)(trial_ctr + (unsigned long long)n_trials);
    g.accepted[(size_t)r * p.n_rx + m] += acc;
    if (n_trials > 0) g.f_valid[r] = 0;
    if (ions) {
      g.ion_cnt[r * 2] = n_act; g.ion_cnt[r * 2 + 1] = n_free;
      if (ion_moved) w.flag[r] = 1;   // rows are stale: rebuild before the next force evaluation
    }
    if (err) atomicOr(&g.err[r], err);
  }
}

// ---- sampling loop helpers ---------------------------------------------------------------------
__global__ void wk_schedule(const __grid_constant__ DevParams p, const __grid_constant__ DevState g,
                            const __grid_constant__ WideState w, const __grid_constant__ cgpe_sample_params sp,
                            uint32_t thr1, uint32_t thr2) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  const unsigned long long sc = (unsigned long long)g.samples_done[r];
  const uint4 o = philox4x32_10((uint32_t)sc, (uint32_t)(sc >> 32), 0u, kTagSchedule, (uint32_t)sp.schedule_seed,
                                (uint32_t)(p.replica_offset + r));
  g.samples_done[r] = (long long)(sc + 1);
  int n = 0;
  if (sp.use_md) n = ((o.x >> 8) < thr1 ? sp.md_steps_per_burst : 0) + ((o.y >> 8) < thr2 ? sp.md_steps_per_burst : 0);
  w.nsteps[r] = n;
  atomicMax(w.sched_max, n);
}

__global__ void __launch_bounds__(kWT)
wk_record(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ cgpe_sample_params sp,
          int s, const double *rg, const double *ree, const int *counts, double *obs, float *qdist) {
  const int r = blockIdx.y, i = blockIdx.x * kWT + threadIdx.x;
  if (i == 0) {
    double *o6 = obs + ((size_t)r * sp.n_samples + s) * CGPE_N_OBS;
    o6[0] = counts[r * 3]; o6[1] = counts[r * 3 + 1]; o6[2] = counts[r * 3 + 2];
    o6[3] = rg[r]; o6[4] = ree[r]; o6[5] = g.time[r];
  }
  if (qdist && sp.q_snapshot_every > 0 && s % sp.q_snapshot_every == 0 && i < p.N) {
    const int row = s / sp.q_snapshot_every;
    if (row < sp.q_snapshot_rows) qdist[((size_t)r * sp.q_snapshot_rows + row) * p.N + i] = g.pos[(size_t)r * p.P + i].w;
  }
}

__global__ void __launch_bounds__(kWT)
wk_list_stats(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w,
              double *out) {
  const int r = blockIdx.y, i = blockIdx.x * kWT + threadIdx.x;
  if (i >= p.P) return;
  const size_t base = (size_t)r * p.P;
  const float4 pi4 = g.pos[base + i];
  const int ti = g.type[base + i], l = chain_index(p, i), S = p.T + 1;
  double a[4] = {0, 0, 0, 0};
  if (p.use_wca) {
    const uint32_t *wl = w.wl + (size_t)r * w.cap_w * p.P;
    for (int e = -4; e < w.wn[base + i]; ++e) {
      int j = -1;
      if (e < 0) { const int off = e < -2 ? e + 2 : e + 3; if (l >= 0 && l + off >= 0 && l + off < p.N) j = i + off; }
      else j = wl[(size_t)e * p.P + i];
      if (j < 0) continue;
      V3 d;
      const float r2 = wdist2(p, pi4, g.pos[base + j], &d);
      a[0] += 1;
      if (ti < p.T && r2 < g.wca_tab[ti * S + g.type[base + j]].z) a[1] += 1;
    }
  }
  const int ci = w.cidx[base + i];
  if (p.use_dh && ci >= 0) {
    const float rc = g.rcut[r];
    const uint32_t *row = w.dl + ((size_t)r * p.n_chg + ci) * w.cap_d;
    for (int e = 0; e < w.dn[(size_t)r * p.n_chg + ci]; ++e) {
      const float4 pj4 = g.pos[base + g.chg[(size_t)r * p.n_chg + row[e]]];
      V3 d;
      a[2] += 1;
      if (pi4.w != 0.f && pj4.w != 0.f && wdist2(p, pi4, pj4, &d) < rc * rc) a[3] += 1;
    }
  }
  for (int q = 0; q < 4; ++q) if (a[q] != 0.0) atomicAdd(&out[r * 4 + q], a[q]);
}

inline dim3 grid_particles(const DevParams &p) { return dim3((p.P + kWT - 1) / kWT, p.R); }
inline dim3 grid_sites_warp(const DevParams &p) { const int per = kWT / 32; return dim3((p.n_chg + per - 1) / per > 0 ? (p.n_chg + per - 1) / per : 1, p.R); }
inline dim3 grid_chargeable(const DevParams &p) { return dim3((p.n_chg + kWT - 1) / kWT > 0 ? (p.n_chg + kWT - 1) / kWT : 1, p.R); }

#define WCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t set_ints(int *a, int n, int v, cudaStream_t st) {
  wk_set_int<<<(n + 255) / 256, 256, 0, st>>>(a, n, v);
  return cudaGetLastError();
}

// lists for the current positions, then forces where the cache is stale
cudaError_t wide_prepare(const DevParams &p, const DevState &g, const WideState &w, int thermostat, cudaStream_t st, int *launches) {
  WCK(set_ints(w.flag, p.R, 1, st));
  wk_rebuild<<<grid_particles(p), kWT, 0, st>>>(p, g, w);
  if (p.use_dh) wk_dh_rows<<<grid_sites_warp(p), kWT, 0, st>>>(p, g, w, -1, 0);
  wk_forces<<<grid_particles(p), kWT, 0, st>>>(p, g, w, -1, 0, thermostat, nullptr);
  *launches += 4;
  return cudaGetLastError();
}

void launch_step(const DevParams &p, const DevState &g, const WideState &w, cudaStream_t st) {
  wk_drift<<<grid_particles(p), kWT, 0, st>>>(p, g, w, kStepFromDevice, 0.f, 0.f);
  wk_rebuild<<<grid_particles(p), kWT, 0, st>>>(p, g, w);
  if (p.use_dh) wk_dh_rows<<<grid_sites_warp(p), kWT, 0, st>>>(p, g, w, kStepFromDevice, 0);
  wk_forces<<<grid_particles(p), kWT, 0, st>>>(p, g, w, kStepFromDevice, 1, 1, nullptr);
}

// n_max velocity-Verlet steps for all replicas (replica r takes part in the first nsteps[r]).  The
// step is launch-bound (4 short kernels), so blocks of kGraphChunk steps are captured once into a
// CUDA graph and replayed; the kernels read the running step index from device memory.
cudaError_t wide_steps(const DevParams &p, const DevState &g, const WideState &w, int n_max, cudaStream_t st, int *launches,
                       WideGraphCache *gc) {
  WCK(cudaMemsetAsync(w.step, 0, 2 * sizeof(int), st));
  int done = 0;
  if (gc && n_max >= 2 * kGraphChunk) {
    if (gc->exec && (gc->chunk != kGraphChunk || std::memcmp(&gc->p_key, &p, sizeof(DevParams)) != 0)) {
      WCK(cudaStreamSynchronize(st));
      wide_graph_release(gc);
    }
    if (!gc->exec) {
      cudaGraph_t graph = nullptr;
      WCK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      for (int s = 0; s < kGraphChunk; ++s) launch_step(p, g, w, st);
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (e != cudaSuccess) return e;
      e = cudaGraphInstantiate(&gc->exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) { gc->exec = nullptr; return e; }
      gc->p_key = p;
      gc->chunk = kGraphChunk;
    }
    for (; done + kGraphChunk <= n_max; done += kGraphChunk) WCK(cudaGraphLaunch(gc->exec, st));
  }
  for (; done < n_max; ++done) launch_step(p, g, w, st);
  *launches += (p.use_dh ? 4 : 3) * n_max;
  return cudaGetLastError();
}

}  // namespace

// ------------------------------------------------------------------------------------------------
void wide_graph_release(WideGraphCache *c) {
  if (c && c->exec) { cudaGraphExecDestroy(c->exec); c->exec = nullptr; }
}

cudaError_t wide_md_run(const DevParams &p, const DevState &g, const WideState &w, int n_steps, cudaStream_t st, int *launches,
                        WideGraphCache *gc) {
  if (n_steps <= 0) return cudaSuccess;
  WCK(set_ints(w.nsteps, p.R, n_steps, st));
  WCK(wide_prepare(p, g, w, 1, st, launches));
  WCK(wide_steps(p, g, w, n_steps, st, launches, gc));
  wk_finish_md<<<(p.R + 127) / 128, 128, 0, st>>>(p, g, w, n_steps, 1);
  *launches += 2;
  return cudaGetLastError();
}

cudaError_t wide_steepest_descent(const DevParams &p, const DevState &g, const WideState &w, int n_steps, float f_max,
                                  float gamma, float max_disp, cudaStream_t st, int *launches) {
  (void)f_max;   // convergence test on f_max is not evaluated here: all n_steps are taken (f_max = 0 in the reference)
  WCK(set_ints(w.nsteps, p.R, n_steps > 0 ? n_steps : 1, st));
  WCK(cudaMemsetAsync(g.f_valid, 0, sizeof(int) * p.R, st));
  WCK(wide_prepare(p, g, w, 0, st, launches));
  for (int s = 0; s < n_steps; ++s) {
    wk_drift<<<grid_particles(p), kWT, 0, st>>>(p, g, w, s, gamma, max_disp);
    wk_rebuild<<<grid_particles(p), kWT, 0, st>>>(p, g, w);
    if (p.use_dh) wk_dh_rows<<<grid_sites_warp(p), kWT, 0, st>>>(p, g, w, s, 0);
    wk_forces<<<grid_particles(p), kWT, 0, st>>>(p, g, w, s, 0, 0, nullptr);
  }
  *launches += 4 * n_steps + 1;
  wk_finish_md<<<(p.R + 127) / 128, 128, 0, st>>>(p, g, w, 0, 0);
  return cudaGetLastError();
}

cudaError_t wide_md_forces(const DevParams &p, const DevState &g, const WideState &w, float4 *out, cudaStream_t st) {
  WCK(set_ints(w.nsteps, p.R, 1, st));
  WCK(set_ints(w.flag, p.R, 1, st));
  wk_rebuild<<<grid_particles(p), kWT, 0, st>>>(p, g, w);
  if (p.use_dh) wk_dh_rows<<<grid_sites_warp(p), kWT, 0, st>>>(p, g, w, 0, 0);
  wk_forces<<<grid_particles(p), kWT, 0, st>>>(p, g, w, 0, 0, 0, out);
  return cudaGetLastError();
}

static cudaError_t wide_mc_blocks(const DevParams &p, const DevState &g, const WideState &w, int n_blocks, const int *rx,
                                  const int *trials, cgpe_trial_record *trace, cudaStream_t st, int *launches) {
  wk_mc_classify<<<grid_chargeable(p), kWT, 0, st>>>(p, g, w);
  if (p.use_dh) wk_dh_rows<<<grid_sites_warp(p), kWT, 0, st>>>(p, g, w, 0, 1);
  wk_mc_prepare<<<grid_chargeable(p), kWT, 0, st>>>(p, g, w);
  for (int b = 0; b < n_blocks; ++b)
    if (trials[b] > 0) wk_mc_block<<<p.R, 32, 0, st>>>(p, g, w, rx[b], trials[b], trace);
  *launches += 2 + n_blocks;
  return cudaGetLastError();
}

cudaError_t wide_mc_run(const DevParams &p, const DevState &g, const WideState &w, int m, int n_trials,
                        cgpe_trial_record *trace, cudaStream_t st, int *launches) {
  WCK(set_ints(w.nsteps, p.R, 0, st));
  WCK(set_ints(w.flag, p.R, 1, st));
  wk_rebuild<<<grid_particles(p), kWT, 0, st>>>(p, g, w);
  WCK(set_ints(w.flag, p.R, 0, st));
  *launches += 4;
  return wide_mc_blocks(p, g, w, 1, &m, &n_trials, trace, st, launches);
}

cudaError_t wide_list_stats(const DevParams &p, const DevState &g, const WideState &w, double *out, cudaStream_t st) {
  WCK(set_ints(w.flag, p.R, 1, st));
  wk_rebuild<<<grid_particles(p), kWT, 0, st>>>(p, g, w);
  WCK(set_ints(w.flag, p.R, 0, st));
  WCK(cudaMemsetAsync(out, 0, sizeof(double) * 4 * p.R, st));
  wk_list_stats<<<grid_particles(p), kWT, 0, st>>>(p, g, w, out);
  return cudaGetLastError();
}

// perform_sampling (E.py:88-106), host-driven: one 4-byte read-back per sample (longest burst)
cudaError_t wide_sample_loop(const DevParams &p, const DevState &g, const WideState &w, const cgpe_sample_params &sp,
                             uint32_t thr1, uint32_t thr2, double *obs, float *qdist, double *tmp_d, int *tmp_i,
                             cudaStream_t st, int *launches, WideGraphCache *gc) {
  WCK(set_ints(w.nsteps, p.R, 0, st));
  WCK(set_ints(w.flag, p.R, 1, st));
  wk_rebuild<<<grid_particles(p), kWT, 0, st>>>(p, g, w);
  WCK(set_ints(w.flag, p.R, 0, st));
  *launches += 4;
  for (int s = 0; s < sp.n_samples; ++s) {
    WCK(cudaMemsetAsync(w.sched_max, 0, sizeof(int), st));
    wk_schedule<<<(p.R + 127) / 128, 128, 0, st>>>(p, g, w, sp, thr1, thr2);
    int n_max = 0;
    WCK(cudaMemcpyAsync(&n_max, w.sched_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    WCK(cudaStreamSynchronize(st));
    *launches += 1;
    if (n_max > 0) {
      wk_rebuild<<<grid_particles(p), kWT, 0, st>>>(p, g, w);   // only replicas whose ions moved in the last MC block
      if (p.use_dh) wk_dh_rows<<<grid_sites_warp(p), kWT, 0, st>>>(p, g, w, -1, 0);
      wk_forces<<<grid_particles(p), kWT, 0, st>>>(p, g, w, -1, 0, 1, nullptr);   // stale forces only
      WCK(wide_steps(p, g, w, n_max, st, launches, gc));
      wk_finish_md<<<(p.R + 127) / 128, 128, 0, st>>>(p, g, w, -1, 1);
      *launches += 2;
    }
    if (sp.n_mc_blocks > 0) WCK(wide_mc_blocks(p, g, w, sp.n_mc_blocks, sp.mc_reaction, sp.mc_trials, nullptr, st, launches));
    WCK(launch_observables(p, g, tmp_d, tmp_d + p.R, tmp_i, st));
    wk_record<<<grid_particles(p), kWT, 0, st>>>(p, g, sp, s, tmp_d, tmp_d + p.R, tmp_i, obs, qdist);
    *launches += 2;
  }
  return cudaGetLastError();
}

}  // namespace cgpe
