// wide_engine.cu — the same hot path for replicas that do not fit one CTA's shared memory
// (P > 2048 particles, e.g. the N = 10^4 chains of the metric): many CTAs per replica, state and
// Verlet lists in global memory (L2-resident: 160 KB of positions per N = 10^4 replica), one
// thread per particle.
//
//   MD step   = wk_drift (half kick + drift + skin/2 vote)  ->  wk_cells + wk_rows (cell-grid list
//               build, only replicas that voted)  ->  wk_forces (all terms + thermostat + second half kick)
//   MC block  = wk_mc_prepare (per-site potentials, all threads) -> wk_mc_block (one CTA per
//               replica, same trial logic and Philox streams as the shared-memory engine)
//   sampling  = the loop of perform_sampling driven from the host, one small read-back per sample
//
// Physics, Philox streams and bookkeeping are those of chain_engine.cu, so both engines follow
// the same Markov chain and are checked against the same oracle; only the summation order of the
// forces differs.  Reference call sites: see chain_engine.cu.
#include "wide_engine.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace cgpe {
namespace {

constexpr int kWT = 256;   // threads per CTA
constexpr int kTabSmem = 128; // WCA table entries ((T+1)^2) the force kernel stages in shared memory
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
  if (!(wdist2(p, x, ref, &dd) <= p.half_skin2)) w.flag[p.R] = 1;   // see rebuild_due; benign race: all writers store 1 (minimum image: ions are folded)
}

// ---- list build: cell grids, then rows -----------------------------------------------------------
// flag[r]: replica r needs new rows (forced, or its ions were moved by a trial).  flag[R]: some
// particle of some replica has left its skin/2 sphere in this step; then every replica taking part
// in the burst rebuilds.  A build is a chain of short dependent phases, so R replicas building
// together cost one such chain instead of R of them scattered over the burst, and since the
// largest displacement among thousands of particles grows at nearly the same pace everywhere the
// replicas lose little by sharing the vote.
__device__ __forceinline__ bool rebuild_due(const DevParams &p, const WideState &w, int r, int step) {
  if (step == kStepFromDevice) step = w.step[0];
  return w.flag[r] || (w.flag[p.R] && step < w.nsteps[r]);
}

constexpr int kCT = 1024;   // threads of the one CTA that sorts a replica's particle set into cells

// Geometry of a grid: the box is cut into nf^3 cells with edge >= the list radius.  A chain fills a
// small part of a dilute box, so the nf^3 cells are mostly empty; they are folded onto a table of
// A^3 buckets (A^3 ~ 2 members, bucket coordinate = cell coordinate mod A).  Members of aliased
// cells share a bucket and are told apart by the distance test, the 27 cells around a particle
// fall into 27 different buckets (A >= 3 and nf mod A not in {1, 2}, so that the periodic wrap
// does not alias neighbours either).
struct GridDims { int nf, A; };

__device__ __forceinline__ GridDims grid_dims(const DevParams &p, float rlist, int nc_max) {
  GridDims d;
  d.nf = (int)fminf(p.box_l / (rlist * 1.0001f), 1.0e6f);
  d.A = min(d.nf, nc_max);
  if (d.A < 3) { d.nf = d.A = 1; return d; }
  if (d.nf > d.A && d.nf % d.A >= 1 && d.nf % d.A <= 2) d.nf -= d.nf % d.A;
  return d;
}
__device__ __forceinline__ int cell_coord(const DevParams &p, float x, int nf) {
  const float xf = x - p.box_l * floorf(x * p.inv_box_l);
  return min(max((int)(xf * p.inv_box_l * (float)nf), 0), nf - 1);
}
__device__ __forceinline__ int bucket_id(const DevParams &p, float4 x, GridDims d) {
  return ((cell_coord(p, x.x, d.nf) % d.A) * d.A + cell_coord(p, x.y, d.nf) % d.A) * d.A + cell_coord(p, x.z, d.nf) % d.A;
}

// Sort the particles (blockIdx.x = 0: all, for the WCA rows) or the chargeable particles
// (blockIdx.x = 1, for the Debye-Hueckel rows) of replica blockIdx.y into buckets: count,
// exclusive scan, scatter, then order every bucket by member index so that the rows - and with
// them the summation order of the forces - do not depend on the order in which the atomics landed.
__global__ void __launch_bounds__(kCT)
wk_cells(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w, int step) {
  const int r = blockIdx.y, kind = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (!rebuild_due(p, w, r, step) || (kind == 0 ? !(p.use_wca || p.n_ions > 0) : !p.use_dh)) return;
  const CellGrid &cg = kind == 0 ? w.cw : w.cd;
  const GridDims gd = grid_dims(p, kind == 0 ? fmaxf(sqrtf(p.rlist_w2), p.grid_w_min) : g.rcut[r] + p.skin, cg.nc_max);
  const int M = gd.A * gd.A * gd.A, n = cg.n;
  const size_t base = (size_t)r * p.P;
  const int *chg = g.chg + (size_t)r * p.n_chg;
  int *start = cg.start + (size_t)r * (cg.m_max + 1), *cursor = cg.cursor + (size_t)r * cg.m_max;
  int *cell_of = cg.cell_of + (size_t)r * n, *tmp = cg.tmp + (size_t)r * n, *order = cg.order + (size_t)r * n;
  float4 *spos = cg.spos + (size_t)r * n;
  __shared__ int warp_tot[2][32];

  for (int c = tid; c < M; c += kCT) cursor[c] = 0;
  __syncthreads();
#pragma unroll 2
  for (int e = tid; e < n; e += kCT) {
    const int i = kind == 0 ? e : chg[e];
    int c = -1;
    if (g.type[base + i] < p.T) {   // hidden slots stay out of the grid
      c = bucket_id(p, g.pos[base + i], gd);
      atomicAdd(&cursor[c], 1);
    }
    cell_of[e] = c;
  }
  __syncthreads();
  int carry = 0;   // exclusive scan of the counts, one barrier per 1024 buckets
  for (int c0 = 0, buf = 0; c0 < M; c0 += kCT, buf ^= 1) {
    const int c = c0 + tid, v = c < M ? cursor[c] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[buf][warp] = incl;
    __syncthreads();
    int t = warp_tot[buf][lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
    const int before = __shfl_sync(0xffffffffu, t, warp > 0 ? warp - 1 : 0);
    const int excl = carry + (warp > 0 ? before : 0) + incl - v;
    if (c < M) { start[c] = excl; cursor[c] = excl; }
    carry += __shfl_sync(0xffffffffu, t, 31);
  }
  const int members = carry;
  if (tid == 0) { start[M] = members; cg.nc[r] = gd.A; cg.nf[r] = gd.nf; }
  __syncthreads();
#pragma unroll 2
  for (int e = tid; e < n; e += kCT) {
    const int c = cell_of[e];
    if (c >= 0) tmp[atomicAdd(&cursor[c], 1)] = e;
  }
  __syncthreads();
#pragma unroll 2
  for (int s = tid; s < members; s += kCT) {
    const int e = tmp[s], c = cell_of[e], lo = start[c], hi = start[c + 1];
    int rank = 0;
    for (int t = lo; t < hi; ++t) rank += tmp[t] < e ? 1 : 0;
    const int i = kind == 0 ? e : chg[e];   // chg ascends, so the order by member index is the order by particle
    order[lo + rank] = i;
    spos[lo + rank] = g.pos[base + i];
  }
}

// The same sort for particle sets too large for one CTA (a multi-chain box of 10^5 .. 10^6 beads with its ions is ONE
// replica: wk_cells would walk millions of members with 1024 threads): count and scatter over the whole grid, the scan
// on one CTA per set (four buckets per thread and round), the in-bucket ordering over the whole grid again.  The
// cursor array is all zero between builds (allocation, then wk_cells_order), which saves a clearing pass.
// (WideState::cells_multi_min: particles per replica from which the multi-CTA build is used)
constexpr int kCB = 256;

struct CellSet { const CellGrid *cg; GridDims gd; int M; };
__device__ __forceinline__ bool cells_set(const DevParams &p, const DevState &g, const WideState &w, int r, int kind, int step,
                                          const CellGrid *&cg, GridDims &gd) {
  if (!rebuild_due(p, w, r, step) || (kind == 0 ? !(p.use_wca || p.n_ions > 0) : !p.use_dh)) return false;
  cg = kind == 0 ? &w.cw : &w.cd;
  gd = grid_dims(p, kind == 0 ? fmaxf(sqrtf(p.rlist_w2), p.grid_w_min) : g.rcut[r] + p.skin, cg->nc_max);
  return true;
}

__global__ void __launch_bounds__(kCB)
wk_cells_count(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w, int step) {
  const int kind = blockIdx.y, r = blockIdx.z;
  const CellGrid *cg; GridDims gd;
  if (!cells_set(p, g, w, r, kind, step, cg, gd)) return;
  const int n = cg->n;
  const size_t base = (size_t)r * p.P;
  const int *chg = g.chg + (size_t)r * p.n_chg;
  int *cursor = cg->cursor + (size_t)r * cg->m_max, *cell_of = cg->cell_of + (size_t)r * n;
  for (int e = blockIdx.x * kCB + threadIdx.x; e < n; e += gridDim.x * kCB) {
    const int i = kind == 0 ? e : chg[e];
    int c = -1;
    if (g.type[base + i] < p.T) {   // hidden slots stay out of the grid
      c = bucket_id(p, g.pos[base + i], gd);
      atomicAdd(&cursor[c], 1);
    }
    cell_of[e] = c;
  }
}

__global__ void __launch_bounds__(kCT)
wk_cells_scan(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w, int step) {
  const int kind = blockIdx.x, r = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const CellGrid *cg; GridDims gd;
  if (!cells_set(p, g, w, r, kind, step, cg, gd)) return;
  const int M = gd.A * gd.A * gd.A;
  int *start = cg->start + (size_t)r * (cg->m_max + 1), *cursor = cg->cursor + (size_t)r * cg->m_max;
  __shared__ int warp_tot[2][32];
  int carry = 0;
  for (int c0 = 0, buf = 0; c0 < M; c0 += 4 * kCT, buf ^= 1) {
    const int c = c0 + 4 * tid;
    int v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = c + q < M ? cursor[c + q] : 0;
    const int mine = v[0] + v[1] + v[2] + v[3];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[buf][warp] = incl;
    __syncthreads();
    int t = warp_tot[buf][lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
    const int before = __shfl_sync(0xffffffffu, t, warp > 0 ? warp - 1 : 0);
    int excl = carry + (warp > 0 ? before : 0) + incl - mine;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (c + q < M) { start[c + q] = excl; cursor[c + q] = excl; excl += v[q]; }
    carry += __shfl_sync(0xffffffffu, t, 31);
  }
  if (tid == 0) { start[M] = carry; cg->nc[r] = gd.A; cg->nf[r] = gd.nf; }
}

__global__ void __launch_bounds__(kCB)
wk_cells_scatter(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w, int step) {
  const int kind = blockIdx.y, r = blockIdx.z;
  const CellGrid *cg; GridDims gd;
  if (!cells_set(p, g, w, r, kind, step, cg, gd)) return;
  const int n = cg->n;
  int *cursor = cg->cursor + (size_t)r * cg->m_max, *tmp = cg->tmp + (size_t)r * n;
  const int *cell_of = cg->cell_of + (size_t)r * n;
  for (int e = blockIdx.x * kCB + threadIdx.x; e < n; e += gridDim.x * kCB) {
    const int c = cell_of[e];
    if (c >= 0) tmp[atomicAdd(&cursor[c], 1)] = e;
  }
}

__global__ void __launch_bounds__(kCB)
wk_cells_order(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w, int step) {
  const int kind = blockIdx.y, r = blockIdx.z;
  const CellGrid *cg; GridDims gd;
  if (!cells_set(p, g, w, r, kind, step, cg, gd)) return;
  const int n = cg->n, M = gd.A * gd.A * gd.A;
  const size_t base = (size_t)r * p.P;
  const int *chg = g.chg + (size_t)r * p.n_chg;
  const int *start = cg->start + (size_t)r * (cg->m_max + 1), *cell_of = cg->cell_of + (size_t)r * n, *tmp = cg->tmp + (size_t)r * n;
  int *order = cg->order + (size_t)r * n, *cursor = cg->cursor + (size_t)r * cg->m_max;
  float4 *spos = cg->spos + (size_t)r * n;
  const int members = start[M];
  for (int s = blockIdx.x * kCB + threadIdx.x; s < members; s += gridDim.x * kCB) {
    const int e = tmp[s], c = cell_of[e], lo = start[c], hi = start[c + 1];
    int rank = 0;
    for (int t = lo; t < hi; ++t) rank += tmp[t] < e ? 1 : 0;
    const int i = kind == 0 ? e : chg[e];
    order[lo + rank] = i;
    spos[lo + rank] = g.pos[base + i];
  }
  for (int c = blockIdx.x * kCB + threadIdx.x; c < M; c += gridDim.x * kCB) cursor[c] = 0;   // nobody reads the cursors any more
}

// Runs [lo, hi) of grid slots that hold the buckets of the 27 cells around x: the buckets of one
// (x, y) column are contiguous in z, so a column gives one run where the three z are consecutive.
template <class F>
__device__ __forceinline__ void for_cell_runs(const DevParams &p, const int *start, GridDims d, float4 x, F &&visit) {
  if (d.A == 1) { visit(0, start[1]); return; }
  const int cx = cell_coord(p, x.x, d.nf), cy = cell_coord(p, x.y, d.nf), cz = cell_coord(p, x.z, d.nf);
  const int z0 = (cz > 0 ? cz - 1 : d.nf - 1) % d.A, z1 = cz % d.A, z2 = (cz < d.nf - 1 ? cz + 1 : 0) % d.A;
  const bool one_run = z0 + 1 == z1 && z1 + 1 == z2;
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy) {
      const int X = ((cx + dx + d.nf) % d.nf) % d.A, Y = ((cy + dy + d.nf) % d.nf) % d.A, col = (X * d.A + Y) * d.A;
      if (one_run) visit(start[col + z0], start[col + z2 + 1]);
      else { visit(start[col + z0], start[col + z0 + 1]); visit(start[col + z1], start[col + z1 + 1]); visit(start[col + z2], start[col + z2 + 1]); }
    }
}

// One warp visits the members of the 27 buckets around x: lanes 0..26 take one bucket each (a
// single bucket: all lanes stride over it); round t hands visit(slot, active) the t-th member of
// every bucket, so the loads of a round are independent.  visit runs converged (ballots allowed).
template <class F>
__device__ __forceinline__ void warp_for_neighbours(const DevParams &p, const int *start, GridDims gd, float4 x, int lane, F &&visit) {
  int s = 0, end = 0, stride = 1;
  if (gd.A == 1) { s = lane; end = start[1]; stride = 32; }
  else if (lane < 27) {
    const int cx = cell_coord(p, x.x, gd.nf) + lane / 9 - 1, cy = cell_coord(p, x.y, gd.nf) + (lane / 3) % 3 - 1,
              cz = cell_coord(p, x.z, gd.nf) + lane % 3 - 1;
    const int bkt = ((((cx + gd.nf) % gd.nf) % gd.A) * gd.A + ((cy + gd.nf) % gd.nf) % gd.A) * gd.A + ((cz + gd.nf) % gd.nf) % gd.A;
    s = start[bkt];
    end = start[bkt + 1];
  }
  for (; __any_sync(0xffffffffu, s < end); s += stride) visit(s, s < end);
}

constexpr int kRowSiteBlocks = 64;   // blocks per replica that build Debye-Hueckel rows (warps stride over the sites)

// Verlet rows of the replicas that rebuild.  Blocks [0, nbw): one thread per particle collects its
// WCA row from the 27 cells around it; the other blocks: one warp per chargeable site collects
// the Debye-Hueckel row, lanes striding over each run of slots (ballot compaction keeps the row
// in slot order).
__global__ void __launch_bounds__(kWT)
wk_rows(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w, int nbw,
        int step) {
  const int r = blockIdx.y;
  if (r == 0 && blockIdx.x == 0 && threadIdx.x == 0) w.step[1] = w.step[0];   // the force kernels of this step read step[1]
  if (!rebuild_due(p, w, r, step)) return;
  const size_t base = (size_t)r * p.P;
  if ((int)blockIdx.x < nbw) {
    const int i = blockIdx.x * kWT + threadIdx.x;
    if (i >= p.P) return;
    const float4 pi = g.pos[base + i];
    int nw = 0;
    if (p.use_wca && g.type[base + i] < p.T) {   // hidden slots build no rows
      const int l = chain_index(p, i);
      int ex_lo = i, ex_hi = i;
      if (l >= 0) { ex_lo = i - min(2, l); ex_hi = i + min(2, p.N - 1 - l); }
      const CellGrid &cg = w.cw;
      const GridDims gd = {cg.nf[r], cg.nc[r]};
      const int *order = cg.order + (size_t)r * cg.n;
      const float4 *spos = cg.spos + (size_t)r * cg.n;
      uint32_t *wl = w.wl + (size_t)r * w.cap_w * p.P;
      for_cell_runs(p, cg.start + (size_t)r * (cg.m_max + 1), gd, pi, [&](int lo, int hi) {
        for (int s = lo; s < hi; ++s) {
          V3 d;
          if (wdist2(p, pi, spos[s], &d) < p.rlist_w2) {
            const int j = order[s];
            if (j < ex_lo || j > ex_hi) {
              if (nw < w.cap_w) wl[(size_t)nw * p.P + i] = (uint32_t)j;
              ++nw;
            }
          }
        }
      });
      if (nw > w.cap_w) { atomicOr(&g.err[r], kErrWcaCap); nw = w.cap_w; }
    }
    w.wn[base + i] = nw;
    w.ref[base + i] = pi;
    if (i == 0) g.rebuilds[r] += 1;
  } else {
    const int lane = threadIdx.x & 31, per = kWT / 32;
    const size_t cb = (size_t)r * p.n_chg;
    const CellGrid &cg = w.cd;
    const GridDims gd = {cg.nf[r], cg.nc[r]};
    const int *order = cg.order + (size_t)r * cg.n, *start = cg.start + (size_t)r * (cg.m_max + 1);
    const float4 *spos = cg.spos + (size_t)r * cg.n;
    const float rl = g.rcut[r] + p.skin, rlist_d2 = rl * rl;
    for (int ci = ((int)blockIdx.x - nbw) * per + (threadIdx.x >> 5); ci < p.n_chg; ci += ((int)gridDim.x - nbw) * per) {
      const int i = g.chg[cb + ci];
      int nd = 0;
      if (g.type[base + i] < p.T) {
        const float4 pi = g.pos[base + i];
        uint32_t *dl = w.dl + (cb + ci) * w.cap_d;
        warp_for_neighbours(p, start, gd, pi, lane, [&](int s, bool active) {
          int j = -1;
          if (active) {
            V3 d;
            if (wdist2(p, pi, spos[s], &d) < rlist_d2) j = order[s];
            if (j == i) j = -1;
          }
          const uint32_t hit = __ballot_sync(0xffffffffu, j >= 0);
          const int slot = nd + __popc(hit & ((1u << lane) - 1u));
          if (j >= 0 && slot < w.cap_d) dl[slot] = (uint32_t)j;
          nd += __popc(hit);
        });
        if (nd > w.cap_d) { if (lane == 0) atomicOr(&g.err[r], kErrDhCap); nd = w.cap_d; }
      }
      if (lane == 0) w.dn[cb + ci] = nd;
    }
  }
}

// Debye-Hueckel force (mode 0) or potential sum_j q_j e(r) (mode 1) of every chargeable particle:
// LPS lanes per site stride over its contiguous row (LPS = 32 for the long rows of weak screening,
// 8 for the short rows of a salty box), shuffle reduction within the group.
template <int LPS>
__global__ void __launch_bounds__(kWT)
wk_dh_rows(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w,
           int step, int mode) {
  const int r = blockIdx.y, sub = threadIdx.x % LPS, ci = blockIdx.x * (kWT / LPS) + threadIdx.x / LPS;
  if (step == kStepFromDevice) step = w.step[1];
  if (mode == 0 && step >= w.nsteps[r]) return;
  const bool site = ci < p.n_chg;   // no early exit: the shuffles below name the full warp
  const size_t base = (size_t)r * p.P, cb = (size_t)r * p.n_chg;
  const float4 pi4 = site ? g.pos[base + g.chg[cb + ci]] : make_float4(0.f, 0.f, 0.f, 0.f);
  float fx = 0.f, fy = 0.f, fz = 0.f, ph = 0.f;
  if (site && (mode == 1 || pi4.w != 0.f)) {
    const float kappa = g.kappa[r], rc = g.rcut[r], rc2 = rc * rc, nkl = -kappa * kLog2e;
    const uint32_t *row = w.dl + (cb + ci) * w.cap_d;
    const int n = w.dn[cb + ci];
    for (int e = sub; e < n; e += LPS) {
      const float4 pj4 = g.pos[base + row[e]];
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
#pragma unroll
  for (int o = LPS / 2; o > 0; o >>= 1) {
    if (mode == 1) ph += __shfl_xor_sync(0xffffffffu, ph, o);
    else {
      fx += __shfl_xor_sync(0xffffffffu, fx, o); fy += __shfl_xor_sync(0xffffffffu, fy, o); fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
  }
  if (!site || sub != 0) return;
  if (mode == 1) w.phi[cb + ci] = ph;
  else {
    const float sc = p.dh_pref * pi4.w;
    w.fdh[cb + ci] = make_float4(sc * fx, sc * fy, sc * fz, 0.f);
  }
}

// ---- tile mode: replicas whose chargeable particles fit one CTA's shared memory --------------------
// A chain of N = 10^4 beads has 3334 chargeable particles: 53 KB as float4.  Every CTA stages ALL of them once
// (position + charge, one gather) and then serves kTileSites sites, one warp per site, from shared memory:
//   wk_rows_tile    Verlet rows by brute force over the staged set (ascending chargeable index; 15x faster than the
//                   cell walk when the list radius is a good part of the chain's extent), written three ways: particle
//                   ids (dl: trial moves, statistics), chargeable indices (dlc) and the CHARGED partners only (dlq)
//   wk_dh_compact   dlq again from dlc after trial moves changed the charges (burst start, start of a reaction block)
//   wk_dh_tile      force / potential sums over dlq
// wk_dh_rows pays a dependent 16-byte L2 gather per listed partner, charged or not; averaged over a titration most
// listed partners carry no charge, and at low salt (hundreds of partners per site) that walk bounded the MD step.
constexpr int kTileThreads = 256, kTileSites = 256;

__host__ __device__ inline size_t tile_smem_bytes(int n_chg) { return sizeof(float4) * (size_t)n_chg + (size_t)((n_chg + 15) / 16 * 16); }

__global__ void __launch_bounds__(kTileThreads)
wk_rows_tile(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w, int step) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  float4 *tcpos = reinterpret_cast<float4 *>(tile_smem);
  uint8_t *vis = tile_smem + sizeof(float4) * (size_t)p.n_chg;
  const int r = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (!rebuild_due(p, w, r, step)) return;
  const size_t base = (size_t)r * p.P, cb = (size_t)r * p.n_chg;
  const int *chg = g.chg + cb;
  for (int ck = tid; ck < p.n_chg; ck += kTileThreads) {
    const int j = chg[ck];
    tcpos[ck] = g.pos[base + j];
    vis[ck] = g.type[base + j] < p.T ? 1 : 0;   // hidden slots build no rows and are nobody's partner
  }
  __syncthreads();
  const float rl = g.rcut[r] + p.skin, rlist_d2 = rl * rl;
  if (blockIdx.x == 0) {
    // extent of the chargeable set: if extent + list radius + skin < L, no listed pair can be closer through a periodic
    // image than directly until the next build (as in chain_engine.cu::build_lists), and wk_dh_tile skips the image search
    __shared__ float bb[6][kTileThreads / 32];
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int ck = tid; ck < p.n_chg; ck += kTileThreads)
      if (vis[ck]) {
        const float4 x = tcpos[ck];
        lo[0] = fminf(lo[0], x.x); hi[0] = fmaxf(hi[0], x.x); lo[1] = fminf(lo[1], x.y); hi[1] = fmaxf(hi[1], x.y);
        lo[2] = fminf(lo[2], x.z); hi[2] = fmaxf(hi[2], x.z);
      }
#pragma unroll
    for (int a = 0; a < 3; ++a)
      for (int o = 16; o > 0; o >>= 1) {
        lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
        hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
      }
    if (lane == 0)
#pragma unroll
      for (int a = 0; a < 3; ++a) { bb[a][warp] = lo[a]; bb[3 + a][warp] = hi[a]; }
    __syncthreads();
    if (tid == 0) {
      float ext = 0.f;
      for (int a = 0; a < 3; ++a) {
        float l_ = 3.0e38f, h_ = -3.0e38f;
        for (int wv = 0; wv < kTileThreads / 32; ++wv) { l_ = fminf(l_, bb[a][wv]); h_ = fmaxf(h_, bb[3 + a][wv]); }
        ext = fmaxf(ext, h_ - l_);
      }
      w.nowrap[r] = (ext + rl + p.skin < p.box_l) ? 1 : 0;   // NaN-safe: wraps
    }
  }
  const int c1 = min((int)(blockIdx.x + 1) * kTileSites, p.n_chg);
  for (int ci = blockIdx.x * kTileSites + warp; ci < c1; ci += kTileThreads / 32) {
    int nd = 0, nq = 0;
    if (vis[ci]) {
      const float4 pi = tcpos[ci];
      uint32_t *dl = w.dl + (cb + ci) * w.cap_d;
      uint16_t *dlc = w.dlc + (cb + ci) * w.cap_d, *dlq = w.dlq + (cb + ci) * w.cap_d;
      for (int c0 = 0; c0 < p.n_chg; c0 += 32) {
        const int ck = c0 + lane;
        bool hit = false, charged = false;
        if (ck < p.n_chg && ck != ci && vis[ck]) {
          const float4 pj = tcpos[ck];
          V3 d;
          hit = wdist2(p, pi, pj, &d) < rlist_d2;
          charged = hit && pj.w != 0.f;
        }
        const uint32_t mh = __ballot_sync(0xffffffffu, hit), mq = __ballot_sync(0xffffffffu, charged);
        const uint32_t below = (1u << lane) - 1u;
        const int slot = nd + __popc(mh & below), slotq = nq + __popc(mq & below);
        if (hit && slot < w.cap_d) { dl[slot] = (uint32_t)chg[ck]; dlc[slot] = (uint16_t)ck; }
        if (charged && slotq < w.cap_d) dlq[slotq] = (uint16_t)ck;
        nd += __popc(mh);
        nq += __popc(mq);
      }
      if (nd > w.cap_d) { if (lane == 0) atomicOr(&g.err[r], kErrDhCap); nd = w.cap_d; nq = min(nq, w.cap_d); }
    }
    if (lane == 0) { w.dn[cb + ci] = nd; w.dnq[cb + ci] = nq; }
  }
}

__global__ void __launch_bounds__(kTileThreads)
wk_dh_compact(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w) {
  const int r = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t base = (size_t)r * p.P, cb = (size_t)r * p.n_chg;
  const int *chg = g.chg + cb;
  const int c1 = min((int)(blockIdx.x + 1) * kTileSites, p.n_chg);
  for (int ci = blockIdx.x * kTileSites + warp; ci < c1; ci += kTileThreads / 32) {
    const uint16_t *dlc = w.dlc + (cb + ci) * w.cap_d;
    uint16_t *dlq = w.dlq + (cb + ci) * w.cap_d;
    const int n = w.dn[cb + ci];
    int nq = 0;
    for (int e0 = 0; e0 < n; e0 += 32) {
      const int e = e0 + lane;
      int ck = -1;
      bool charged = false;
      if (e < n) { ck = dlc[e]; charged = g.pos[base + chg[ck]].w != 0.f; }
      const uint32_t mq = __ballot_sync(0xffffffffu, charged);
      if (charged) dlq[nq + __popc(mq & ((1u << lane) - 1u))] = (uint16_t)ck;
      nq += __popc(mq);
    }
    if (lane == 0) w.dnq[cb + ci] = nq;
  }
}

// Lane striding and arithmetic are those of wk_dh_rows<32> (over the charged partners only).
__global__ void __launch_bounds__(kTileThreads)
wk_dh_tile(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w,
           int step, int mode) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  float4 *tcpos = reinterpret_cast<float4 *>(tile_smem);
  const int r = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (step == kStepFromDevice) step = w.step[1];
  if (mode == 0 && step >= w.nsteps[r]) return;
  const size_t base = (size_t)r * p.P, cb = (size_t)r * p.n_chg;
  for (int ck = tid; ck < p.n_chg; ck += kTileThreads) tcpos[ck] = g.pos[base + g.chg[cb + ck]];
  __syncthreads();
  const float kappa = g.kappa[r], rc = g.rcut[r], rc2 = rc * rc, nkl = -kappa * kLog2e;
  const bool nowrap = w.nowrap[r] != 0;   // uniform: set by the row build in force
  const int c1 = min((int)(blockIdx.x + 1) * kTileSites, p.n_chg);
  for (int ci = blockIdx.x * kTileSites + warp; ci < c1; ci += kTileThreads / 32) {
    const float4 pi4 = tcpos[ci];
    float fx = 0.f, fy = 0.f, fz = 0.f, ph = 0.f;
    if (mode == 1 || pi4.w != 0.f) {
      const uint16_t *row = w.dlq + (cb + ci) * w.cap_d;
      const int n = w.dnq[cb + ci];
      for (int e0 = lane; e0 < n; e0 += 128) {
        int ck[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) ck[q] = e0 + 32 * q < n ? (int)row[e0 + 32 * q] : -1;   // four entries in flight
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (ck[q] < 0) continue;
          const float4 pj4 = tcpos[ck[q]];
          V3 d;
          float r2;
          if (nowrap) { d.x = pi4.x - pj4.x; d.y = pi4.y - pj4.y; d.z = pi4.z - pj4.z; r2 = fmaf(d.x, d.x, fmaf(d.y, d.y, d.z * d.z)); }
          else r2 = wdist2(p, pi4, pj4, &d);
          if (r2 < rc2) {
            float fr;
            const float en = dh_energy_unit(r2, kappa, nkl, &fr);
            ph = fmaf(pj4.w, en, ph);
            fr *= pj4.w;
            fx = fmaf(fr, d.x, fx); fy = fmaf(fr, d.y, fy); fz = fmaf(fr, d.z, fz);
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (mode == 1) ph += __shfl_xor_sync(0xffffffffu, ph, o);
      else {
        fx += __shfl_xor_sync(0xffffffffu, fx, o); fy += __shfl_xor_sync(0xffffffffu, fy, o); fz += __shfl_xor_sync(0xffffffffu, fz, o);
      }
    }
    if (lane != 0) continue;
    if (mode == 1) w.phi[cb + ci] = ph;
    else {
      const float sc = p.dh_pref * pi4.w;
      w.fdh[cb + ci] = make_float4(sc * fx, sc * fy, sc * fz, 0.f);
    }
  }
}

inline bool short_dh_rows(const DevParams &p, const WideState &w);
inline bool tile_mode(const DevParams &p, const WideState &w) { return p.use_dh && w.dlc != nullptr && !short_dh_rows(p, w); }
inline dim3 grid_tiles(const DevParams &p) { return dim3((p.n_chg + kTileSites - 1) / kTileSites, p.R); }

// rows are short when a uniform box would put fewer than ~2 dozen chargeable particles in the list sphere
inline bool short_dh_rows(const DevParams &p, const WideState &w) { return w.dh_rows_short != 0; }
void launch_dh_rows(const DevParams &p, const DevState &g, const WideState &w, int step, int mode, cudaStream_t st) {
  if (tile_mode(p, w)) {
    const int bytes = (int)tile_smem_bytes(p.n_chg);
    cudaFuncSetAttribute(wk_dh_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    wk_dh_tile<<<grid_tiles(p), kTileThreads, bytes, st>>>(p, g, w, step, mode);
  } else if (short_dh_rows(p, w)) {
    const int per = kWT / 8;
    wk_dh_rows<8><<<dim3((p.n_chg + per - 1) / per > 0 ? (p.n_chg + per - 1) / per : 1, p.R), kWT, 0, st>>>(p, g, w, step, mode);
  } else {
    const int per = kWT / 32;
    wk_dh_rows<32><<<dim3((p.n_chg + per - 1) / per > 0 ? (p.n_chg + per - 1) / per : 1, p.R), kWT, 0, st>>>(p, g, w, step, mode);
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
  if (blockIdx.x == 0 && threadIdx.x == 0) { w.flag[r] = 0; if (r == 0) w.flag[p.R] = 0; }   // every block of wk_cells / wk_rows has seen them
  if (step >= w.nsteps[r]) return;
  if (mode == 0 && step < 0 && g.f_valid[r] && !out) return;   // cached forces are still good
  const size_t base = (size_t)r * p.P;
  const float4 *pos = g.pos + base;
  const uint8_t *type = g.type + base;
  const int S = p.T + 1, N = p.N, t0 = blockIdx.x * kWT, tid = threadIdx.x;
  int err = 0;
  // Chain-local terms, each computed once by its owner bead o: bond (o, o+1), angle (o-1, o, o+1),
  // torsion (o-1, o, o+1, o+2) (M.py:159-170) and the WCA pairs (o, o+1), (o, o+2) that the rows
  // leave out.  The four beads of an owner receive their shares through four slot planes (self,
  // from o-1, from o+1, from o-2: one writer per entry, no atomics); owners t0-2 .. t0+kWT serve the
  // tile, so the three beyond it are computed twice, once per neighbouring tile.
  __shared__ float4 tpos[kWT + 6];
  __shared__ uint8_t ttype[kWT + 6];
  __shared__ float slot[4][3][kWT];
  __shared__ float4 stab[kTabSmem];   // WCA table: one dependent global load less per row entry
  const bool tab_staged = S * S <= kTabSmem;
  if (tab_staged) for (int k = tid; k < S * S; k += kWT) stab[k] = g.wca_tab[k];
  const float4 *wca_tab = tab_staged ? stab : g.wca_tab;
  if (t0 >= p.NC) __syncthreads();   // (the chain tiles synchronise below)
  if (t0 < p.NC) {
    for (int k = tid; k < kWT + 6; k += kWT) {
      const int j = t0 - 3 + k;
      const bool in = j >= 0 && j < p.P;
      tpos[k] = in ? pos[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      ttype[k] = in ? type[j] : (uint8_t)p.T;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { slot[q][0][tid] = 0.f; slot[q][1][tid] = 0.f; slot[q][2][tid] = 0.f; }
    __syncthreads();
    for (int oi = tid; oi < kWT + 3; oi += kWT) {
      const int o = t0 - 2 + oi;
      if (o < 0 || o >= p.NC) continue;
      const int l = chain_index(p, o), k = oi + 1;   // tpos[k] is bead o
      const V3 xm = xyz(tpos[k - 1]), x0 = xyz(tpos[k]), x1 = xyz(tpos[k + 1]), x2 = xyz(tpos[k + 2]);
      V3 fs = {0.f, 0.f, 0.f}, fa = fs, fb = fs, fc = fs, f1, f2, f3, f4;   // self, to o+1, to o-1, to o+2
      if (p.bond_kind != CGPE_BOND_NONE && l <= N - 2) { bond_term(p, x0, x1, &f1, &err); fs = fs + f1; fa = fa - f1; }
      if (p.use_angle && l >= 1 && l <= N - 2) { angle_term(p, xm, x0, x1, &f1, &f2, &f3); fb = fb + f1; fs = fs + f2; fa = fa + f3; }
      if (p.use_dih && l >= 1 && l <= N - 3) {
        dihedral_term(p, xm, x0, x1, x2, &f1, &f2, &f3, &f4);
        fb = fb + f1; fs = fs + f2; fa = fa + f3; fc = fc + f4;
      }
      const int to = ttype[k];
      if (p.use_wca && to < p.T) {
        if (l <= N - 2 && ttype[k + 1] < p.T) {
          const float4 tab = wca_tab[to * S + ttype[k + 1]];
          const V3 d = x0 - x1;
          const float r2 = dot(d, d);
          if (r2 < tab.z) { const V3 fw = wca_force_over_r(tab, r2) * d; fs = fs + fw; fa = fa - fw; }
        }
        if (l <= N - 3 && ttype[k + 2] < p.T) {
          const float4 tab = wca_tab[to * S + ttype[k + 2]];
          const V3 d = x0 - x2;
          const float r2 = dot(d, d);
          if (r2 < tab.z) { const V3 fw = wca_force_over_r(tab, r2) * d; fs = fs + fw; fc = fc - fw; }
        }
      }
      const int e = oi - 2;   // tile index of bead o
      if (e >= 0 && e < kWT) { slot[0][0][e] = fs.x; slot[0][1][e] = fs.y; slot[0][2][e] = fs.z; }
      if (e + 1 >= 0 && e + 1 < kWT) { slot[1][0][e + 1] = fa.x; slot[1][1][e + 1] = fa.y; slot[1][2][e + 1] = fa.z; }
      if (e - 1 >= 0 && e - 1 < kWT) { slot[2][0][e - 1] = fb.x; slot[2][1][e - 1] = fb.y; slot[2][2][e - 1] = fb.z; }
      if (e + 2 >= 0 && e + 2 < kWT) { slot[3][0][e + 2] = fc.x; slot[3][1][e + 2] = fc.y; slot[3][2][e + 2] = fc.z; }
    }
    __syncthreads();
  }
  if (err) atomicOr(&g.err[r], err);
  err = 0;
  if (i >= p.P) return;
  const float4 pi4 = pos[i];
  const int ti = type[i];
  V3 f = {0.f, 0.f, 0.f};
  if (i < p.NC) {
    f.x = (slot[0][0][tid] + slot[1][0][tid]) + (slot[2][0][tid] + slot[3][0][tid]);
    f.y = (slot[0][1][tid] + slot[1][1][tid]) + (slot[2][1][tid] + slot[3][1][tid]);
    f.z = (slot[0][2][tid] + slot[1][2][tid]) + (slot[2][2][tid] + slot[3][2][tid]);
  }
  if (p.use_wca && ti < p.T) {
    const uint32_t *wl = w.wl + (size_t)r * w.cap_w * p.P;
    const int n = w.wn[base + i];
    for (int e = 0; e < n; ++e) {
      const int j = wl[(size_t)e * p.P + i];
      const float4 tab = wca_tab[ti * S + type[j]];
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
  if (p.fixed_mask != 0u && ((p.fixed_mask >> ti) & 1u)) f = V3{0.f, 0.f, 0.f};   // fix=[True]*3
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

// Work a trial hands to the whole CTA (explicit ions): the sums over all particles.
struct McOrder { int op, b, cbi, ci, t_new; float q_new, dqb; float4 pb; };
enum { kMcExit = 0, kMcIonEnergy = 1, kMcIonPhi = 2, kMcSitePhi = 3 };
constexpr int kMcSiteRowMin = 96;     // accepted flips with a longer Debye-Hueckel row are shared with the helper warps
constexpr int kMcThreads = 512;       // CTA of a block of trials with explicit ions
constexpr int kMcChunk = 256;          // trials per launch when the ion sums come from the cell grids (IONS = 2)
constexpr int kMcLoadThreads = 512;   // without ions: threads that fill the mirrors, then help with the rows of accepted flips

__device__ __forceinline__ void mc_bar(int n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }

// What the trial loop reads.  STAGED: mirrors in shared memory, filled when the block starts and
// written through to the global arrays, so a trial's chain of dependent reads (list entry -> site
// -> cached potentials -> row partners) stays on the SM; otherwise the global arrays themselves.
struct McView {
  float4 *cpos;        // [n_chg] chargeable particles: position + charge           (staged only)
  uint8_t *ctype;      // [n_chg] their types                                        (staged only)
  float4 *npos;        // [P - n_chg] the other particles: never hidden, never moved  (staged, ions)
  float *phi, *wcache; // [n_chg]
  uint8_t *wflag;      // [n_chg]
  int *cidx;           // [P]
  int *lp, *ld;        // bookkeeping lists of the reaction
  int *ion_active, *ion_free;
  int n_other;
};

// bytes of the mirrors (host and device agree through this one function)
__host__ __device__ inline size_t mc_stage_bytes(int P, int n_chg, int n_ions, bool with_ions) {
  const size_t sites = (size_t)(n_chg - n_ions > 0 ? n_chg - n_ions : 0);
  size_t b = (size_t)n_chg * (16 + 4 + 4) + (size_t)P * 4 + 2 * 4 * sites + 2 * (((size_t)n_chg + 15) / 16 * 16);
  if (with_ions) b += (size_t)(P - n_chg) * 16 + 2 * 4 * (size_t)n_ions;
  return b + 64;
}

// interaction of the ion at o.pb with every visible particle (site o.ci taken in its trial state):
// this thread's share of  sum_j q_j e(r)  and of the smallest squared distance
template <bool STAGED>
__device__ __forceinline__ void mc_ion_energy_share(const DevParams &p, const DevState &g, const McView &v, size_t base, size_t cb,
                                                    const McOrder &o, float kappa, float nkl, float rc2, int tid, int n_threads,
                                                    float *esum, float *dmin2) {
  float es = 0.f, dm = 3.0e38f;
  const int *chg = g.chg + cb;
  // batches of four particles per thread: the (global-memory) loads of a batch are issued together
  for (int c0 = tid; c0 < p.n_chg; c0 += 4 * n_threads) {
    int tj[4];
    float4 pj[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ck = c0 + q * n_threads;
      tj[q] = p.T;
      if (ck < p.n_chg && ck != o.cbi) {
        const int j = STAGED ? 0 : chg[ck];
        tj[q] = ck == o.ci ? o.t_new : (int)(STAGED ? v.ctype[ck] : g.type[base + j]);
        pj[q] = STAGED ? v.cpos[ck] : g.pos[base + j];
        if (ck == o.ci) pj[q].w = o.q_new;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (tj[q] >= p.T) continue;
      V3 d;
      const float r2 = wdist2(p, o.pb, pj[q], &d);
      dm = fminf(dm, r2);
      if (p.use_dh && pj[q].w != 0.f && r2 < rc2) {
        float fr;
        es = fmaf(pj[q].w, dh_energy_unit(r2, kappa, nkl, &fr), es);
      }
    }
  }
  if (STAGED) {
    for (int k = tid; k < v.n_other; k += n_threads) {
      V3 d;
      dm = fminf(dm, wdist2(p, o.pb, v.npos[k], &d));
    }
  } else {
#pragma unroll 4
    for (int j = tid; j < p.P; j += n_threads) {
      if (v.cidx[j] >= 0 || g.type[base + j] >= p.T) continue;
      V3 d;
      dm = fminf(dm, wdist2(p, o.pb, g.pos[base + j], &d));
    }
  }
  *esum = warp_sum(es);
  *dmin2 = wwarp_min(dm);
}

// the cached site potentials after the ion at o.pb appeared (dqb > 0) or vanished
template <bool STAGED>
__device__ __forceinline__ void mc_ion_phi_share(const DevParams &p, const DevState &g, const WideState &w, const McView &v,
                                                 size_t base, size_t cb, const McOrder &o, float kappa, float nkl, float rc2,
                                                 int tid, int n_threads) {
  const int *chg = g.chg + cb;
  const int n_site = p.n_chg - p.n_ions;   // chg ascends and the ion slots follow the chains: only sites own a potential
  for (int c0 = tid; c0 < n_site; c0 += 4 * n_threads) {
    float4 pj[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ck = c0 + q * n_threads;
      if (ck < n_site && ck != o.cbi) pj[q] = STAGED ? v.cpos[ck] : g.pos[base + chg[ck]];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ck = c0 + q * n_threads;
      if (ck >= n_site || ck == o.cbi) continue;
      V3 d;
      const float r2 = wdist2(p, o.pb, pj[q], &d);
      if (r2 < rc2) {
        float fr;
        const float ph = fmaf(o.dqb, dh_energy_unit(r2, kappa, nkl, &fr), v.phi[ck]);
        v.phi[ck] = ph;
        if (STAGED) w.phi[cb + ck] = ph;
      }
    }
  }
}

// the cached potentials of the listed partners of site o.ci after its charge changed by o.dqb (accepted flip): this
// thread's share of the row, four entries in flight (row entry -> chargeable index -> position are dependent loads)
template <bool STAGED>
__device__ __forceinline__ void mc_site_phi_share(const DevParams &p, const DevState &g, const WideState &w, const McView &v,
                                                  size_t base, size_t cb, int ci, float dq, float4 pi4, float kappa, float nkl,
                                                  float rc2, int tid, int n_threads) {
  const int n = w.dn[cb + ci];
  const uint32_t *row = w.dl + (cb + ci) * w.cap_d;
  for (int e0 = tid; e0 < n; e0 += 4 * n_threads) {
    int jj[4], ck[4];
    float4 pj[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) jj[q] = e0 + n_threads * q < n ? (int)row[e0 + n_threads * q] : -1;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (jj[q] >= 0) {
        ck[q] = v.cidx[jj[q]];
        pj[q] = STAGED ? v.cpos[ck[q]] : g.pos[base + jj[q]];
      }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (jj[q] < 0) continue;
      V3 d;
      const float r2 = wdist2(p, pi4, pj[q], &d);
      if (r2 < rc2) {
        float fr;
        const float ph = fmaf(dq, dh_energy_unit(r2, kappa, nkl, &fr), v.phi[ck[q]]);
        v.phi[ck[q]] = ph;
        if (STAGED) w.phi[cb + ck[q]] = ph;
      }
    }
  }
}

// One CTA per replica: warp 0 runs the trial loop of chain_engine.cu::mc_block.  With explicit ions
// (IONS, blockDim = kMcThreads) an insertion / deletion needs sums over all particles; warp 0
// posts them as orders and the other warps, parked at a named barrier, take their share.  Without
// ions the other warps fill the mirrors and then take their share of the partner row of every accepted flip (long rows:
// N = 10^4 at low salt lists hundreds of partners per site, one warp alone spends microseconds on them).
//
// IONS = 2 (replicas too large for the mirrors, e.g. a multi-chain box of 10^5 beads): the sums of
// an ion move are local - charged particles within r_cut, any particle within the exclusion range
// - and are taken from the cell grids by warp 0 alone.  The host hands such a block over in chunks
// of kMcChunk trials with the grids rebuilt before each, so the grids know every particle where it
// is except the ions inserted during the chunk: those are marked `stale` (skipped in the buckets)
// and kept in a journal that every ion move also scans.
template <int IONS, bool STAGED>
__global__ void __launch_bounds__(kMcThreads)
wk_mc_block(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, const __grid_constant__ WideState w,
            int m, int n_trials, cgpe_trial_record *trace_all, int trace_stride, int trace_offset) {
  const int r = blockIdx.x, lane = threadIdx.x & 31, tid = threadIdx.x, n_threads = blockDim.x, n_warps = n_threads >> 5;
  const DevReaction rx = p.rx[m];
  constexpr bool ions = IONS != 0;
  const size_t base = (size_t)r * p.P, cb = (size_t)r * p.n_chg;
  const float kappa = g.kappa[r], rc = g.rcut[r], rc2 = rc * rc, nkl = -kappa * kLog2e;
  const int *chg = g.chg + cb;
  int *g_lp = g.lst_prot + ((size_t)r * p.n_rx + m) * p.P, *g_ld = g.lst_deprot + ((size_t)r * p.n_rx + m) * p.P;
  int *g_active = g.ion_active + (size_t)r * (ions ? p.n_ions : 1), *g_free = g.ion_free + (size_t)r * (ions ? p.n_ions : 1);
  int n_p = g.cnt[((size_t)r * p.n_rx + m) * 2], n_d = g.cnt[((size_t)r * p.n_rx + m) * 2 + 1];
  int n_act = ions ? g.ion_cnt[r * 2] : 0, n_free = ions ? g.ion_cnt[r * 2 + 1] : 0;
  __shared__ McOrder ord;
  __shared__ float red_sum[32], red_min[32];
  __shared__ int n_other_s;
  extern __shared__ __align__(16) unsigned char mc_smem[];
  McView v;
  if (STAGED) {
    const int sites = max(p.n_chg - p.n_ions, 0), pad = (p.n_chg + 15) / 16 * 16;
    unsigned char *q = mc_smem;
    v.cpos = reinterpret_cast<float4 *>(q); q += (size_t)p.n_chg * 16;
    v.npos = reinterpret_cast<float4 *>(q); q += ions ? (size_t)(p.P - p.n_chg) * 16 : 0;
    v.phi = reinterpret_cast<float *>(q); q += (size_t)p.n_chg * 4;
    v.wcache = reinterpret_cast<float *>(q); q += (size_t)p.n_chg * 4;
    v.cidx = reinterpret_cast<int *>(q); q += (size_t)p.P * 4;
    v.lp = reinterpret_cast<int *>(q); q += (size_t)sites * 4;
    v.ld = reinterpret_cast<int *>(q); q += (size_t)sites * 4;
    v.ion_active = reinterpret_cast<int *>(q); q += ions ? (size_t)p.n_ions * 4 : 0;
    v.ion_free = reinterpret_cast<int *>(q); q += ions ? (size_t)p.n_ions * 4 : 0;
    v.ctype = q; q += pad;
    v.wflag = q;
    if (tid == 0) n_other_s = 0;
    __syncthreads();
    for (int ck = tid; ck < p.n_chg; ck += n_threads) {
      const int j = chg[ck];
      v.cpos[ck] = g.pos[base + j]; v.ctype[ck] = g.type[base + j];
      v.phi[ck] = w.phi[cb + ck]; v.wcache[ck] = w.wcache[cb + ck]; v.wflag[ck] = w.wflag[cb + ck];
    }
    for (int j = tid; j < p.P; j += n_threads) {
      const int c = w.cidx[base + j];
      v.cidx[j] = c;
      if (ions && c < 0 && g.type[base + j] < p.T) v.npos[atomicAdd(&n_other_s, 1)] = g.pos[base + j];
    }
    for (int k = tid; k < n_p; k += n_threads) v.lp[k] = g_lp[k];
    for (int k = tid; k < n_d; k += n_threads) v.ld[k] = g_ld[k];
    if (ions) {
      for (int k = tid; k < n_act; k += n_threads) v.ion_active[k] = g_active[k];
      for (int k = tid; k < n_free; k += n_threads) v.ion_free[k] = g_free[k];
    }
    __syncthreads();
    v.n_other = n_other_s;
  } else {
    v.cpos = nullptr; v.ctype = nullptr; v.npos = nullptr; v.n_other = 0;
    v.phi = w.phi + cb; v.wcache = w.wcache + cb; v.wflag = w.wflag + cb; v.cidx = w.cidx + base;
    v.lp = g_lp; v.ld = g_ld; v.ion_active = g_active; v.ion_free = g_free;
  }
  const bool site_helpers = IONS == 0 && p.use_dh && n_threads > 32;   // uniform
  if constexpr (IONS != 1) if (tid >= 32) {
    if (!site_helpers) return;
    for (;;) {
      mc_bar(n_threads);   // an order is posted
      if (ord.op == kMcExit) return;
      mc_site_phi_share<STAGED>(p, g, w, v, base, cb, ord.ci, ord.dqb, ord.pb, kappa, nkl, rc2, tid, n_threads);
      mc_bar(n_threads);   // the potentials are current
    }
  }
  if constexpr (IONS == 1) if (tid >= 32) {   // helpers
    for (;;) {
      mc_bar(n_threads);   // an order is posted
      const McOrder o = ord;
      if (o.op == kMcExit) return;
      if (o.op == kMcIonEnergy) {
        float es, dm;
        mc_ion_energy_share<STAGED>(p, g, v, base, cb, o, kappa, nkl, rc2, tid, n_threads, &es, &dm);
        if (lane == 0) { red_sum[tid >> 5] = es; red_min[tid >> 5] = dm; }
      } else {
        mc_ion_phi_share<STAGED>(p, g, w, v, base, cb, o, kappa, nkl, rc2, tid, n_threads);
      }
      mc_bar(n_threads);   // shares are in
    }
  }
  const int n_sites = n_p + n_d;
  const float arg_fwd = (float)(2.302585092994045684 * (g.pH[r] - rx.pK)) * kLog2e;
  const float neg_beta_log2e = -rx.inv_kT * kLog2e;
  const float dq_fwd = rx.q_deprot - rx.q_prot, coul_fwd = p.dh_pref * dq_fwd;
  const uint32_t k0 = rx.seed, k1 = (uint32_t)(p.replica_offset + r);
  const unsigned long long trial_ctr = (unsigned long long)g.trials[(size_t)r * p.n_rx + m];
  cgpe_trial_record *trace = trace_all ? trace_all + (size_t)r * trace_stride + trace_offset : nullptr;
  __shared__ int journal[kMcChunk];
  int n_journal = 0;
  uint8_t *stale = w.stale + base;
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
      int *src = fwd ? v.lp : v.ld, *dst = fwd ? v.ld : v.lp;
      int *g_src = fwd ? g_lp : g_ld, *g_dst = fwd ? g_ld : g_lp;
      const float u = u24(a2);
      bool possible = n_src > 0;
      if (ions && !fwd && n_act == 0) possible = false;
      if (ions && fwd && n_free == 0) { possible = false; err |= kErrIonPool; }
      if (!possible) {
        if (trace && lane == 0) { cgpe_trial_record rec = {-1, -1, (int8_t)(fwd ? 1 : -1), 0, 0, 0, 0.f, 0.f, u}; trace[b0 + tt] = rec; }
        continue;
      }
      const int k = (int)__umulhi(a1, (uint32_t)n_src);
      const int i = src[k], ci = v.cidx[i];
      const int t_new = fwd ? rx.type_deprot : rx.type_prot;
      const float q_new = fwd ? rx.q_deprot : rx.q_prot, dq = fwd ? dq_fwd : -dq_fwd;
      float de = v.wcache[ci];
      if (p.use_dh) de = fmaf(coul_fwd, v.phi[ci], de);
      if (!fwd) de = -de;
      int b = -1, kb = 0, cbi = -1;
      bool excluded = false;
      float4 pb4 = make_float4(0.f, 0.f, 0.f, rx.q_ion);
      if (ions) {   // see chain_engine.cu::mc_block
        if (fwd) {
          b = v.ion_free[n_free - 1];
          cbi = v.cidx[b];
          pb4.x = u24(__shfl_sync(0xffffffffu, o1.x, tt)) * p.box_l;
          pb4.y = u24(__shfl_sync(0xffffffffu, o1.y, tt)) * p.box_l;
          pb4.z = u24(__shfl_sync(0xffffffffu, o1.z, tt)) * p.box_l;
        } else {
          kb = (int)__umulhi(__shfl_sync(0xffffffffu, o.w, tt), (uint32_t)n_act);
          b = v.ion_active[kb];
          cbi = v.cidx[b];
          const float4 x = STAGED ? v.cpos[cbi] : g.pos[base + b];
          pb4.x = x.x; pb4.y = x.y; pb4.z = x.z;
        }
        float esum, dmin2;
        if (IONS == 2) {
          float es = 0.f, dm = 3.0e38f;
          // the particle in slot j as the trial sees it: type (hidden slots do not interact), position, charge
          auto look = [&](int j, bool charged_only) {
            if (j == b) return;
            const int tj = j == i ? t_new : (int)g.type[base + j];
            if (tj >= p.T) return;
            const float4 pj4 = g.pos[base + j];
            const float qj = j == i ? q_new : pj4.w;
            V3 d;
            const float r2 = wdist2(p, pb4, pj4, &d);
            if (!charged_only) dm = fminf(dm, r2);
            else if (p.use_dh && qj != 0.f && r2 < rc2) {
              float fr;
              es = fmaf(qj, dh_energy_unit(r2, kappa, nkl, &fr), es);
            }
          };
          if (p.use_dh) {
            const CellGrid &cg = w.cd;
            const int *order = cg.order + (size_t)r * cg.n;
            warp_for_neighbours(p, cg.start + (size_t)r * (cg.m_max + 1), GridDims{cg.nf[r], cg.nc[r]}, pb4, lane,
                                [&](int s, bool active) { if (active) { const int j = order[s]; if (!stale[j]) look(j, true); } });
          }
          {
            const CellGrid &cg = w.cw;
            const int *order = cg.order + (size_t)r * cg.n;
            warp_for_neighbours(p, cg.start + (size_t)r * (cg.m_max + 1), GridDims{cg.nf[r], cg.nc[r]}, pb4, lane,
                                [&](int s, bool active) { if (active) { const int j = order[s]; if (!stale[j]) look(j, false); } });
          }
          for (int k = lane; k < n_journal; k += 32) { look(journal[k], true); look(journal[k], false); }
          esum = warp_sum(es);
          dmin2 = wwarp_min(dm);
        } else {
        if (lane == 0) { ord.op = kMcIonEnergy; ord.b = b; ord.cbi = cbi; ord.ci = ci; ord.t_new = t_new; ord.q_new = q_new; ord.pb = pb4; }
        mc_bar(n_threads);
        mc_ion_energy_share<STAGED>(p, g, v, base, cb, ord, kappa, nkl, rc2, tid, n_threads, &esum, &dmin2);
        if (lane == 0) { red_sum[0] = esum; red_min[0] = dmin2; }
        mc_bar(n_threads);
        esum = warp_sum(lane < n_warps ? red_sum[lane] : 0.f);
        dmin2 = wwarp_min(lane < n_warps ? red_min[lane] : 3.0e38f);
        }
        const float e_ion = p.dh_pref * rx.q_ion * esum;
        de += fwd ? e_ion : -e_ion;
        excluded = dmin2 < rx.excl2;
      }
      const float bf = excluded ? 0.f : ((float)n_src / (float)(n_sites - n_src + 1)) * exp2f(fmaf(neg_beta_log2e, de, fwd ? arg_fwd : -arg_fwd));
      const bool accept = u < bf;
      if (trace && lane == 0) { cgpe_trial_record rec = {i, b, (int8_t)(fwd ? 1 : -1), (int8_t)accept, (int8_t)excluded, 0, de, bf, u}; trace[b0 + tt] = rec; }
      if (accept) {
        const float4 pi4 = STAGED ? v.cpos[ci] : g.pos[base + i];
        const bool refresh = p.use_wca && v.wflag[ci] != 0;
        __syncwarp();
        if (lane == 0) {
          g.type[base + i] = (uint8_t)t_new;
          g.pos[base + i].w = q_new;
          const int last = src[n_src - 1];
          src[k] = last;
          dst[n_dst] = i;
          if (STAGED) { v.ctype[ci] = (uint8_t)t_new; v.cpos[ci].w = q_new; g_src[k] = last; g_dst[n_dst] = i; }
        }
        if (fwd) { n_p -= 1; n_d += 1; } else { n_d -= 1; n_p += 1; }
        acc += 1;
        if (p.use_dh && dq != 0.f && site_helpers && w.dn[cb + ci] > kMcSiteRowMin) {
          if (lane == 0) { ord.op = kMcSitePhi; ord.ci = ci; ord.dqb = dq; ord.pb = pi4; }
          mc_bar(n_threads);
          mc_site_phi_share<STAGED>(p, g, w, v, base, cb, ci, dq, pi4, kappa, nkl, rc2, tid, n_threads);
          mc_bar(n_threads);
        } else if (p.use_dh && dq != 0.f) {
          const int n = w.dn[cb + ci];
          const uint32_t *row = w.dl + (cb + ci) * w.cap_d;
          // four row entries in flight per lane: the loads of a batch do not wait for one another
          for (int e0 = lane; e0 < n; e0 += 128) {
            int jj[4], ck[4];
            float4 pj[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) jj[q] = e0 + 32 * q < n ? (int)row[e0 + 32 * q] : -1;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (jj[q] >= 0) {
                ck[q] = v.cidx[jj[q]];
                pj[q] = STAGED ? v.cpos[ck[q]] : g.pos[base + jj[q]];
              }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (jj[q] < 0) continue;
              V3 d;
              const float r2 = wdist2(p, pi4, pj[q], &d);
              if (r2 < rc2) {
                float fr;
                const float ph = fmaf(dq, dh_energy_unit(r2, kappa, nkl, &fr), v.phi[ck[q]]);
                v.phi[ck[q]] = ph;
                if (STAGED) w.phi[cb + ck[q]] = ph;
              }
            }
          }
        }
        __syncwarp();
        if (ions) {
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
              v.ion_active[n_act] = b;
              if (STAGED) { v.ctype[cbi] = (uint8_t)rx.type_ion; v.cpos[cbi] = pb4; g_active[n_act] = b; }
            } else {
              g.type[base + b] = (uint8_t)p.T;
              g.pos[base + b].w = 0.f;
              g.vel[base + b] = make_float4(0.f, 0.f, 0.f, 0.f);
              const int last = v.ion_active[n_act - 1];
              v.ion_active[kb] = last;
              v.ion_free[n_free] = b;
              if (STAGED) { v.ctype[cbi] = (uint8_t)p.T; v.cpos[cbi].w = 0.f; g_active[kb] = last; g_free[n_free] = b; }
            }
          }
          if (fwd) { n_free -= 1; n_act += 1; } else { n_act -= 1; n_free += 1; }
          if (IONS == 2) {
            __syncwarp();
            if (fwd && !stale[b]) {   // the grids hold slot b elsewhere (or not at all): journal it
              if (lane == 0) { stale[b] = 1; journal[n_journal] = b; }
              n_journal += 1;
            }
            if (p.use_dh) {   // cached potentials of the titratable sites around the ion
              const CellGrid &cg = w.cd;
              const int *order = cg.order + (size_t)r * cg.n;
              warp_for_neighbours(p, cg.start + (size_t)r * (cg.m_max + 1), GridDims{cg.nf[r], cg.nc[r]}, pb4, lane,
                                  [&](int s, bool active) {
                if (!active) return;
                const int j = order[s], ck = v.cidx[j];
                if (j == b || w.rxof[cb + ck] == 255) return;
                V3 d;
                const float r2 = wdist2(p, pb4, g.pos[base + j], &d);
                if (r2 < rc2) {
                  float fr;
                  v.phi[ck] = fmaf(dqb, dh_energy_unit(r2, kappa, nkl, &fr), v.phi[ck]);
                }
              });
            }
            __syncwarp();
          } else if (p.use_dh) {
            __syncwarp();
            if (lane == 0) { ord.op = kMcIonPhi; ord.cbi = cbi; ord.dqb = dqb; ord.pb = pb4; }
            mc_bar(n_threads);
            mc_ion_phi_share<STAGED>(p, g, w, v, base, cb, ord, kappa, nkl, rc2, tid, n_threads);
            mc_bar(n_threads);
          }
          ion_moved = true;
          __syncwarp();
        }
        if (refresh) {   // rare: a titratable neighbour within the WCA cutoff; reads the (written-through) global state
          const int l = chain_index(p, i), n_w = w.wn[base + i];
          const uint32_t *wl = w.wl + (size_t)r * w.cap_w * p.P;
          __threadfence_block();
          for (int e = lane; e < n_w + 4; e += 32) {
            int j = -1;
            if (e < 4) {
              const int off = e < 2 ? e - 2 : e - 1;
              if (l >= 0 && l + off >= 0 && l + off < p.N) j = i + off;
            } else {
              j = wl[(size_t)(e - 4) * p.P + i];
            }
            if (j >= 0) {
              const int cj = v.cidx[j];
              if (cj >= 0 && w.rxof[cb + cj] != 255) {
                const float wc = wide_wca_swap(p, g, w, r, cj, nullptr);
                v.wcache[cj] = wc;
                if (STAGED) w.wcache[cb + cj] = wc;
              }
            }
          }
          __syncwarp();
        }
      }
    }
  }
  if (IONS == 1 || site_helpers) {   // release the helpers
    if (lane == 0) ord.op = kMcExit;
    mc_bar(n_threads);
  }
  if (IONS == 2) {
    __syncwarp();
    for (int k = lane; k < n_journal; k += 32) stale[journal[k]] = 0;
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
      const float4 pj4 = g.pos[base + row[e]];
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

// cell grids of the replicas whose flag is set
void launch_cells(const DevParams &p, const DevState &g, const WideState &w, cudaStream_t st, int step) {
  if (!(p.use_wca || p.use_dh || p.n_ions > 0)) return;
  const int kinds = p.use_dh ? 2 : 1;
  if (p.P < w.cells_multi_min) { wk_cells<<<dim3(kinds, p.R), kCT, 0, st>>>(p, g, w, step); return; }
  const int nb = std::max(1, std::min((p.P + 4 * kCB - 1) / (4 * kCB), 148 * 8));
  wk_cells_count<<<dim3(nb, kinds, p.R), kCB, 0, st>>>(p, g, w, step);
  wk_cells_scan<<<dim3(kinds, p.R), kCT, 0, st>>>(p, g, w, step);
  wk_cells_scatter<<<dim3(nb, kinds, p.R), kCB, 0, st>>>(p, g, w, step);
  wk_cells_order<<<dim3(nb, kinds, p.R), kCB, 0, st>>>(p, g, w, step);
}

// kernels a call of launch_rebuild / launch_dh_rows / launch_step puts on the stream (the library reports its launch count)
inline int rebuild_launches(const DevParams &p, const WideState &w) {
  const int cells = !(p.use_wca || p.use_dh || p.n_ions > 0) ? 0 : p.P < w.cells_multi_min ? 1 : 4;
  return cells + 1 + (tile_mode(p, w) ? 1 : 0);
}
inline int step_launches(const DevParams &p, const WideState &w) { return 2 + rebuild_launches(p, w) + (p.use_dh ? 1 : 0); }

// cell grids + rows of the replicas whose flag is set
void launch_rebuild(const DevParams &p, const DevState &g, const WideState &w, cudaStream_t st, int step = -1) {
  const bool tile = tile_mode(p, w);
  const int nbw = (p.P + kWT - 1) / kWT, per = kWT / 32, nbd = p.use_dh && !tile ? std::min((p.n_chg + per - 1) / per, std::max(kRowSiteBlocks, 148 * 16 / std::max(p.R, 1))) : 0;
  launch_cells(p, g, w, st, step);
  wk_rows<<<dim3(nbw + nbd, p.R), kWT, 0, st>>>(p, g, w, nbw, step);
  if (tile) {
    const int bytes = (int)tile_smem_bytes(p.n_chg);
    cudaFuncSetAttribute(wk_rows_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    wk_rows_tile<<<grid_tiles(p), kTileThreads, bytes, st>>>(p, g, w, step);
  }
}

// tile mode: the charged-partner rows follow the charges (call where trial moves may have changed them since the last build)
void launch_dh_compact(const DevParams &p, const DevState &g, const WideState &w, cudaStream_t st) {
  if (tile_mode(p, w)) wk_dh_compact<<<grid_tiles(p), kTileThreads, 0, st>>>(p, g, w);
}

// lists for the current positions, then forces where the cache is stale
cudaError_t wide_prepare(const DevParams &p, const DevState &g, const WideState &w, int thermostat, cudaStream_t st, int *launches) {
  WCK(set_ints(w.flag, p.R, 1, st));
  launch_rebuild(p, g, w, st);
  if (p.use_dh) launch_dh_rows(p, g, w, -1, 0, st);
  wk_forces<<<grid_particles(p), kWT, 0, st>>>(p, g, w, -1, 0, thermostat, nullptr);
  *launches += 2 + rebuild_launches(p, w) + (p.use_dh ? 1 : 0);
  return cudaGetLastError();
}

void launch_step(const DevParams &p, const DevState &g, const WideState &w, cudaStream_t st) {
  wk_drift<<<grid_particles(p), kWT, 0, st>>>(p, g, w, kStepFromDevice, 0.f, 0.f);
  launch_rebuild(p, g, w, st, kStepFromDevice);
  if (p.use_dh) launch_dh_rows(p, g, w, kStepFromDevice, 0, st);
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
  *launches += step_launches(p, w) * n_max;
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
    launch_rebuild(p, g, w, st, s);
    if (p.use_dh) launch_dh_rows(p, g, w, s, 0, st);
    wk_forces<<<grid_particles(p), kWT, 0, st>>>(p, g, w, s, 0, 0, nullptr);
  }
  *launches += step_launches(p, w) * n_steps + 1;
  wk_finish_md<<<(p.R + 127) / 128, 128, 0, st>>>(p, g, w, 0, 0);
  return cudaGetLastError();
}

cudaError_t wide_md_forces(const DevParams &p, const DevState &g, const WideState &w, float4 *out, cudaStream_t st) {
  WCK(set_ints(w.nsteps, p.R, 1, st));
  WCK(set_ints(w.flag, p.R, 1, st));
  launch_rebuild(p, g, w, st);
  if (p.use_dh) launch_dh_rows(p, g, w, 0, 0, st);
  wk_forces<<<grid_particles(p), kWT, 0, st>>>(p, g, w, 0, 0, 0, out);
  return cudaGetLastError();
}

static cudaError_t wide_mc_blocks(const DevParams &p, const DevState &g, const WideState &w, int n_blocks, const int *rx,
                                  const int *trials, cgpe_trial_record *trace, cudaStream_t st, int *launches) {
  wk_mc_classify<<<grid_chargeable(p), kWT, 0, st>>>(p, g, w);
  launch_dh_compact(p, g, w, st);
  if (p.use_dh) launch_dh_rows(p, g, w, 0, 1, st);
  wk_mc_prepare<<<grid_chargeable(p), kWT, 0, st>>>(p, g, w);
  for (int b = 0; b < n_blocks; ++b)
    if (trials[b] > 0) {
      const bool ions = p.n_ions > 0;
      const size_t bytes = mc_stage_bytes(p.P, p.n_chg, p.n_ions, ions);
      const bool staged = bytes <= w.mc_stage_limit;
      auto launch = [&](auto kernel, int threads, int n, int offset) -> cudaError_t {
        if (staged) {
          cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
          if (e != cudaSuccess) return e;
        }
        kernel<<<p.R, threads, staged ? bytes : 0, st>>>(p, g, w, rx[b], n, trace, trials[b], offset);
        return cudaGetLastError();
      };
      if (ions && staged) WCK(launch(wk_mc_block<1, true>, kMcThreads, trials[b], 0));
      else if (ions && w.mc_grid_ok) {
        for (int t0 = 0; t0 < trials[b]; t0 += kMcChunk) {   // grids rebuilt before every chunk (see wk_mc_block)
          WCK(set_ints(w.flag, p.R, 1, st));
          launch_cells(p, g, w, st, -1);
          WCK(launch(wk_mc_block<2, false>, 32, std::min(kMcChunk, trials[b] - t0), t0));
          *launches += 2 + rebuild_launches(p, w) - 1 - (tile_mode(p, w) ? 1 : 0);
        }
      }
      else if (ions) WCK(launch(wk_mc_block<1, false>, kMcThreads, trials[b], 0));
      else if (staged) WCK(launch(wk_mc_block<0, true>, kMcLoadThreads, trials[b], 0));
      else WCK(launch(wk_mc_block<0, false>, 32, trials[b], 0));
    }
  *launches += 2 + n_blocks + (p.use_dh ? 1 : 0) + (tile_mode(p, w) ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t wide_mc_run(const DevParams &p, const DevState &g, const WideState &w, int m, int n_trials,
                        cgpe_trial_record *trace, cudaStream_t st, int *launches) {
  WCK(set_ints(w.nsteps, p.R, 0, st));
  WCK(set_ints(w.flag, p.R, 1, st));
  launch_rebuild(p, g, w, st);
  WCK(set_ints(w.flag, p.R, 0, st));
  *launches += 3 + rebuild_launches(p, w);
  return wide_mc_blocks(p, g, w, 1, &m, &n_trials, trace, st, launches);
}

cudaError_t wide_list_stats(const DevParams &p, const DevState &g, const WideState &w, double *out, cudaStream_t st) {
  WCK(set_ints(w.flag, p.R, 1, st));
  launch_rebuild(p, g, w, st);
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
  launch_rebuild(p, g, w, st);
  WCK(set_ints(w.flag, p.R, 0, st));
  *launches += 3 + rebuild_launches(p, w);
  for (int s = 0; s < sp.n_samples; ++s) {
    WCK(cudaMemsetAsync(w.sched_max, 0, sizeof(int), st));
    wk_schedule<<<(p.R + 127) / 128, 128, 0, st>>>(p, g, w, sp, thr1, thr2);
    int n_max = 0;
    WCK(cudaMemcpyAsync(&n_max, w.sched_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    WCK(cudaStreamSynchronize(st));
    *launches += 1;
    if (n_max > 0) {
      launch_rebuild(p, g, w, st);   // only replicas whose ions moved in the last MC block
      launch_dh_compact(p, g, w, st);   // the last reaction blocks changed the charges
      if (p.use_dh) launch_dh_rows(p, g, w, -1, 0, st);
      wk_forces<<<grid_particles(p), kWT, 0, st>>>(p, g, w, -1, 0, 1, nullptr);   // stale forces only
      WCK(wide_steps(p, g, w, n_max, st, launches, gc));
      wk_finish_md<<<(p.R + 127) / 128, 128, 0, st>>>(p, g, w, -1, 1);
      *launches += 2 + rebuild_launches(p, w) + (p.use_dh ? 1 : 0) + (tile_mode(p, w) ? 1 : 0);
    }
    if (sp.n_mc_blocks > 0) WCK(wide_mc_blocks(p, g, w, sp.n_mc_blocks, sp.mc_reaction, sp.mc_trials, nullptr, st, launches));
    WCK(launch_observables(p, g, tmp_d, tmp_d + p.R, tmp_i, st));
    wk_record<<<grid_particles(p), kWT, 0, st>>>(p, g, sp, s, tmp_d, tmp_d + p.R, tmp_i, obs, qdist);
    *launches += 2;
  }
  return cudaGetLastError();
}

}  // namespace cgpe
