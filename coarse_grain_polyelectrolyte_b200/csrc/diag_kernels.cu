// diag_kernels.cu — analysis kernels that do not depend on the Verlet lists of the MD engine:
// energies, forces, minimum distance and chain observables evaluated over ALL pairs from global
// memory, one CTA per replica.  They replace system.analysis.energy() / min_dist() / calc_rg /
// calc_re / number_of_particles (EnsembleDynamics.py:34,42,45,97-101 of the reference) and double
// as an independent check of the list-based force path.  Pair terms are FP32, sums are FP64.
#include "chain_engine.cuh"

namespace cgpe {
namespace {

constexpr int kDiagThreads = 256;

__device__ __forceinline__ float pair_dist2(const DevParams &p, float4 a, float4 b, V3 *d) {
  d->x = min_image(a.x - b.x, p.box_l, p.inv_box_l);
  d->y = min_image(a.y - b.y, p.box_l, p.inv_box_l);
  d->z = min_image(a.z - b.z, p.box_l, p.inv_box_l);
  return fmaf(d->x, d->x, fmaf(d->y, d->y, d->z * d->z));
}

__device__ double block_sum(double v, double *scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) t += scratch[w];
  return t;
}

// out[r][5] = total, kinetic, coulomb, bonded, non_bonded
__global__ void __launch_bounds__(kDiagThreads)
k_energy(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, double *out) {
  __shared__ double scratch[32];
  const int r = blockIdx.x;
  const float4 *pos = g.pos + (size_t)r * p.P;
  const float4 *vel = g.vel + (size_t)r * p.P;
  const uint8_t *type = g.type + (size_t)r * p.P;
  const float kappa = g.kappa[r], rc = g.rcut[r], rc2 = rc * rc, nkl = -kappa * kLog2e;
  double e_kin = 0, e_coul = 0, e_bond = 0, e_wca = 0;
  int err = 0;
  for (int i = threadIdx.x; i < p.P; i += blockDim.x) {
    const float4 pi = pos[i];
    const int ti = type[i];
    if (ti < p.T) {
      const float4 v = vel[i];
      e_kin += 0.5 * ((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z);
    }
    if (i < p.NC) {
      const int l = i % p.N;
      if (l >= 1) {
        V3 f1, f2, f3, f4;
        if (p.bond_kind != CGPE_BOND_NONE) e_bond += bond_term(p, xyz(pi), xyz(pos[i - 1]), &f1, &err);
        if (p.use_angle && l <= p.N - 2) e_bond += angle_term(p, xyz(pos[i - 1]), xyz(pi), xyz(pos[i + 1]), &f1, &f2, &f3);
        if (p.use_dih && l <= p.N - 3)
          e_bond += dihedral_term(p, xyz(pos[i - 1]), xyz(pi), xyz(pos[i + 1]), xyz(pos[i + 2]), &f1, &f2, &f3, &f4);
      }
    }
    if (ti >= p.T) continue;
    for (int j = i + 1; j < p.P; ++j) {
      const int tj = type[j];
      if (tj >= p.T) continue;
      const float4 pj = pos[j];
      V3 d;
      const float r2 = pair_dist2(p, pi, pj, &d);
      if (p.use_wca) e_wca += wca_energy(g.wca_tab[ti * (p.T + 1) + tj], r2);
      if (p.use_dh && pi.w != 0.f && pj.w != 0.f && r2 < rc2) {
        float fr;
        e_coul += (double)(p.dh_pref * pi.w * pj.w * dh_energy_unit(r2, kappa, nkl, &fr));
      }
    }
  }
  e_kin = block_sum(e_kin, scratch);
  e_coul = block_sum(e_coul, scratch);
  e_bond = block_sum(e_bond, scratch);
  e_wca = block_sum(e_wca, scratch);
  if (threadIdx.x == 0) {
    double *o = out + (size_t)r * CGPE_N_ENERGY;
    o[0] = e_kin + e_coul + e_bond + e_wca; o[1] = e_kin; o[2] = e_coul; o[3] = e_bond; o[4] = e_wca;
  }
  if (err) atomicOr(&g.err[r], err);
}

// interaction forces, every bead sums the terms it takes part in (no hand-over, no lists)
__global__ void __launch_bounds__(kDiagThreads)
k_forces(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, float4 *out) {
  const int r = blockIdx.x;
  const float4 *pos = g.pos + (size_t)r * p.P;
  const uint8_t *type = g.type + (size_t)r * p.P;
  const float kappa = g.kappa[r], rc = g.rcut[r], rc2 = rc * rc, nkl = -kappa * kLog2e;
  int err = 0;
  for (int i = threadIdx.x; i < p.P; i += blockDim.x) {
    const float4 pi = pos[i];
    const int ti = type[i];
    double fx = 0, fy = 0, fz = 0;
    auto add = [&](V3 f) { fx += f.x; fy += f.y; fz += f.z; };
    if (i < p.NC) {
      const int l = i % p.N, N = p.N;
      V3 f1, f2, f3, f4;
      if (p.bond_kind != CGPE_BOND_NONE) {
        if (l >= 1) { bond_term(p, xyz(pi), xyz(pos[i - 1]), &f1, &err); add(f1); }
        if (l <= N - 2) { bond_term(p, xyz(pi), xyz(pos[i + 1]), &f1, &err); add(f1); }
      }
      if (p.use_angle) {   // angles centred on l-1, l, l+1 (centre must lie in 1..N-2)
        if (l - 1 >= 1 && l - 1 <= N - 2) { angle_term(p, xyz(pos[i - 2]), xyz(pos[i - 1]), xyz(pi), &f1, &f2, &f3); add(f3); }
        if (l >= 1 && l <= N - 2) { angle_term(p, xyz(pos[i - 1]), xyz(pi), xyz(pos[i + 1]), &f1, &f2, &f3); add(f2); }
        if (l + 1 >= 1 && l + 1 <= N - 2) { angle_term(p, xyz(pi), xyz(pos[i + 1]), xyz(pos[i + 2]), &f1, &f2, &f3); add(f1); }
      }
      if (p.use_dih) {     // dihedral owned by bead o covers (o-1,o,o+1,o+2), o in 1..N-3
        for (int role = 0; role < 4; ++role) {
          const int o = l + 1 - role;   // this bead is p1 (role 0) .. p4 (role 3)
          if (o >= 1 && o <= N - 3) {
            const int b = i + 1 - role;
            dihedral_term(p, xyz(pos[b - 1]), xyz(pos[b]), xyz(pos[b + 1]), xyz(pos[b + 2]), &f1, &f2, &f3, &f4);
            add(role == 0 ? f1 : role == 1 ? f2 : role == 2 ? f3 : f4);
          }
        }
      }
    }
    if (ti < p.T) {
      for (int j = 0; j < p.P; ++j) {
        const int tj = type[j];
        if (j == i || tj >= p.T) continue;
        const float4 pj = pos[j];
        V3 d;
        const float r2 = pair_dist2(p, pi, pj, &d);
        float fr = 0.f;
        if (p.use_wca) {
          const float4 tab = g.wca_tab[ti * (p.T + 1) + tj];
          if (r2 < tab.z) fr += wca_force_over_r(tab, r2);
        }
        if (p.use_dh && pi.w != 0.f && pj.w != 0.f && r2 < rc2) {
          float f_;
          dh_energy_unit(r2, kappa, nkl, &f_);
          fr += p.dh_pref * pi.w * pj.w * f_;
        }
        if (fr != 0.f) { fx += (double)(fr * d.x); fy += (double)(fr * d.y); fz += (double)(fr * d.z); }
      }
    }
    if ((p.fixed_mask >> ti) & 1u) fx = fy = fz = 0;   // fix=[True]*3
    out[(size_t)r * p.P + i] = make_float4((float)fx, (float)fy, (float)fz, 0.f);
  }
  if (err) atomicOr(&g.err[r], err);
}

__global__ void __launch_bounds__(kDiagThreads)
k_min_dist(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, double *out) {
  __shared__ float scratch[32];
  const int r = blockIdx.x;
  const float4 *pos = g.pos + (size_t)r * p.P;
  const uint8_t *type = g.type + (size_t)r * p.P;
  float best = 3.0e38f;
  for (int i = threadIdx.x; i < p.P; i += blockDim.x) {
    if (type[i] >= p.T) continue;   // hidden slots are not particles
    const float4 pi = pos[i];
    for (int j = i + 1; j < p.P; ++j) {
      V3 d;
      if (type[j] < p.T) best = fminf(best, pair_dist2(p, pi, pos[j], &d));
    }
  }
  for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) best = fminf(best, scratch[w]);
    out[r] = sqrt((double)best);
  }
}

__global__ void __launch_bounds__(kDiagThreads)
k_observables(const __grid_constant__ DevParams p, const __grid_constant__ DevState g, double *rg, double *ree,
              int *counts) {
  __shared__ double scratch[32];
  const int r = blockIdx.x;
  const float4 *pos = g.pos + (size_t)r * p.P;
  const float4 p0 = pos[0];
  double sx = 0, sy = 0, sz = 0, s2 = 0;
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
    const float4 q = pos[i];
    const double dx = (double)q.x - p0.x, dy = (double)q.y - p0.y, dz = (double)q.z - p0.z;
    sx += dx; sy += dy; sz += dz; s2 += dx * dx + dy * dy + dz * dz;
  }
  sx = block_sum(sx, scratch); sy = block_sum(sy, scratch); sz = block_sum(sz, scratch); s2 = block_sum(s2, scratch);
  if (threadIdx.x == 0) {
    const double inv = 1.0 / p.N;
    const double v = s2 * inv - (sx * sx + sy * sy + sz * sz) * inv * inv;
    if (rg) rg[r] = sqrt(v > 0.0 ? v : 0.0);
    const float4 pe = pos[p.N - 1];
    const double ex = (double)pe.x - p0.x, ey = (double)pe.y - p0.y, ez = (double)pe.z - p0.z;
    if (ree) ree[r] = sqrt(ex * ex + ey * ey + ez * ez);
    if (counts) {
      int n_ion = 0;
      counts[r * 3] = counts[r * 3 + 1] = 0;
      for (int m = 0; m < p.n_rx; ++m) {
        const int nd = g.cnt[((size_t)r * p.n_rx + m) * 2 + 1];
        if (m < 2) counts[r * 3 + m] = nd;
        n_ion += nd;
      }
      counts[r * 3 + 2] = p.n_ions > 0 ? g.ion_cnt[r * 2] : n_ion;   // implicit ions: N_B = N_N + N_N2
    }
  }
}

}  // namespace

cudaError_t launch_energy(const DevParams &p, const DevState &g, double *out, cudaStream_t st) {
  k_energy<<<p.R, kDiagThreads, 0, st>>>(p, g, out);
  return cudaGetLastError();
}
cudaError_t launch_forces(const DevParams &p, const DevState &g, float4 *out, cudaStream_t st) {
  k_forces<<<p.R, kDiagThreads, 0, st>>>(p, g, out);
  return cudaGetLastError();
}
cudaError_t launch_min_dist(const DevParams &p, const DevState &g, double *out, cudaStream_t st) {
  k_min_dist<<<p.R, kDiagThreads, 0, st>>>(p, g, out);
  return cudaGetLastError();
}
cudaError_t launch_observables(const DevParams &p, const DevState &g, double *rg, double *ree, int *counts,
                               cudaStream_t st) {
  k_observables<<<p.R, kDiagThreads, 0, st>>>(p, g, rg, ree, counts);
  return cudaGetLastError();
}

}  // namespace cgpe
