// cgpe_device.cuh — device-side building blocks shared by the kernels of libcgpe.so (sm_100a).
//
// Everything here is FP32 arithmetic on float4 (x,y,z,q) particles.  The functional forms follow
// the reference's force field as configured in src/modules/Model.py (cited per function) and are
// checked against the FP64 CPU oracle by tests/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cgpe.h"

namespace cgpe {

constexpr int kMaxRx = CGPE_MAX_REACTIONS;
constexpr float kLog2e = 1.4426950408889634f;

// error bits raised by kernels into DevState::err[r]
enum : int { kErrWcaCap = 1, kErrDhCap = 2, kErrNonFinite = 4, kErrFene = 8, kErrDiverged = 16, kErrIonPool = 32 };

struct DevReaction {
  int type_prot, type_deprot, type_ion;
  uint32_t seed;
  float q_prot, q_deprot, q_ion;
  float inv_kT;
  float excl2;       // exclusion_range^2 (explicit ions)
  float sqrt_kT;     // Maxwell-Boltzmann width of an inserted ion's velocity
  double pK;
};

// Byte offsets of the shared-memory arrays of the replica engine (chain_engine.cu).  They live in
// the kernel-parameter constant bank, so an address is one add away and no pointer stays in a register.
struct SmemOff {
  int pos, tab, red, slot, fd, dhn, cnt, flag, wl, db, qm, chg, cidx, lp, ld, wn, type, phi, wc, rxof, cpos, wrange, wflag, ratio, qlist, ionA, ionF, ionV, total;
};

// Uniform (replica-independent) parameters, passed to kernels by value.
struct DevParams {
  int R, P, NC, N, T;           // replicas, particles, chain beads (n_chains*N), beads per chain, types
  int n_chg;                    // chargeable particles per replica (sites [+ ions])
  int n_ions;                   // explicit-ion slots per replica (0 = implicit ions)
  int n_rx;
  int bond_kind, use_angle, use_dih, dih_mult, use_wca, use_dh;
  int cap_w;                    // WCA list row capacity
  int dw;                       // words per row of the Debye-Hueckel bit list (odd: conflict-free)
  int dh_group;                 // sub-threads sharing one Debye-Hueckel row (1..4)
  int dense_policy;             // 0 = choose per rebuild, 1 = never dense, 2 = always dense (diagnostics)
  int replica_offset;
  uint32_t thermostat_seed;
  float bond_k, bond_r0, fene_drmax2;
  float angle_k, angle_phi0;
  float dih_k, dih_cos_phase, dih_sin_phase, dih_sin_min2;
  float dh_pref;
  float box_l, inv_box_l;
  float dt, half_dt;
  double dt_d;                  // time step in double for the clock (system.time)
  float skin, half_skin2;       // Verlet skin, (skin/2)^2
  float rlist_w2;               // (max WCA cutoff + skin)^2; with fixed particles in slots only (rlist_wf2 > rlist_w2):
                                // over the pairs of free types
  float rlist_wf2;              // (max WCA cutoff + skin)^2 over all type pairs: list radius towards a fixed particle
  float rc_w_max2;              // (max WCA cutoff)^2
  int dhc_cap, dhc_stride;      // compact Debye-Hueckel partner lists: entries per sub-thread, sub-threads per replica (padded)
  float grid_w_min;             // smallest cell edge of the all-particle grid (explicit ions: the largest exclusion range)
  float force_cap;              // 0 = off
  uint32_t fixed_mask;          // bit t: particles of type t never move (cgpe_config::fixed_type_mask)
  int frame_every;              // > 0: every frame_every MD steps a frame of chain-0 observables is recorded (cgpe_frames_begin)
  int frame_flags;              // CGPE_FRAME_BONDS | CGPE_FRAME_DIHEDRALS
  int frame_cap;                // frames per replica the buffers hold
  SmemOff off;
  DevReaction rx[kMaxRx];
};

// Device arrays of one handle.
struct DevState {
  float4 *pos;        // [R][P] x,y,z,q (unfolded)
  float4 *vel;        // [R][P]
  float4 *frc;        // [R][P] total force of the last force evaluation (valid iff f_valid[r])
  uint8_t *type;      // [R][P]
  float4 *wca_tab;    // [(T+1)^2] {24 eps, sigma^2, rc^2, eps}
  int *chg;           // [R][n_chg] particle ids that can carry charge
  int *lst_prot;      // [R][n_rx][P] bookkeeping lists (particle ids), swap-remove order
  int *lst_deprot;    // [R][n_rx][P]
  int *cnt;           // [R][n_rx][2] n_prot, n_deprot
  float *mc_e;        // [R][n_chg][mc_e_stride] exp(-kappa r)/r of every listed pair inside r_cut, 0 elsewhere: written by mc_prepare for
                      // the reaction blocks of a sample (positions rest meanwhile), read row-wise by an accepted flip.  NULL: not in use
  int mc_e_stride;    // row stride in floats (multiple of 128)
  uint16_t *dhc;      // [R][dhc_cap / 8][dhc_stride][8] listed AND charged partners of every (site, sub-thread), shared-memory engine
  int *ion_active;    // [R][n_ions] particle ids of the free H+ (type_ion), swap-remove order
  int *ion_free;      // [R][n_ions] hidden slots, used as a stack
  int *ion_cnt;       // [R][2] n_active, n_free
  float *kappa, *rcut, *kT, *gamma;  // [R]
  double *pH;         // [R]
  double *time;       // [R]
  long long *md_counter, *md_steps, *samples_done, *rebuilds;   // [R]
  double *frames;              // [R][frame_cap][CGPE_FRAME_COLS] COM position, COM velocity, Rg, Ree of chain 0
  float *frame_bonds;          // [R][frame_cap][N-1] bond lengths (CGPE_FRAME_BONDS)
  float *frame_dih;            // [R][frame_cap][N-3] torsion angles in (-pi, pi] (CGPE_FRAME_DIHEDRALS)
  int *n_frames;               // [R] frames recorded since the last fetch (may exceed frame_cap: the excess was dropped)
  int *frame_phase;            // [R] MD steps since the last frame
  long long *cta_ns;           // [R][2][2] start / end (GPU global timer, ns) of the CTA(s) of each replica in the last fused sampling launch
  long long *compact_builds;   // [R] times the compact Debye-Hueckel partner lists were written (cgpe_engine_info)
  long long *trials, *accepted;                                  // [R][n_rx]
  double *plan;                // [R][8] cluster-engine planner: [2],[3] trials / accepted at the last plan, [4] charged pairs inside the list radius, [5] acceptance, [6],[7] cost and MD steps predicted for the launch
  const int *order;            // [R] or NULL: replica taken by each CTA of the one-CTA engine's fused sampling launch (most expensive first when the launch runs in waves)
  int *engine_choice;          // NULL, or: the fused sampling launch was issued on both shared-memory engines and the plan (k_duo_plan) wrote which one runs - 0 one CTA per replica, 1 cluster; the other kernel returns at once
  int *f_valid;       // [R]
  int *err;           // [R]
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Stream layout (identical in oracle/cgpe_oracle.c):
//   key = (seed, global replica index), counter = (lo, hi, sub, tag)
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kTagLangevin = 0x4C414E47u;
constexpr uint32_t kTagMc = 0x4D430000u;
constexpr uint32_t kTagSchedule = 0x53434844u;

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// uniform in [0,1) on a 2^-24 grid (exact in float)
__device__ __forceinline__ float u24(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// zero-mean uniform in (-1/2,1/2) on a 2^-23 grid (exact in float)
__device__ __forceinline__ float centred_u23(uint32_t x) {
  return fmaf((float)((int)(x >> 9) - 4194304), 1.0f / 8388608.0f, 1.0f / 16777216.0f);
}

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
// minimum image of one component; round-to-nearest through the 1.5*2^23 trick keeps it on the
// FP32 FMA pipe (FRND would issue at a quarter of the rate)
__device__ __forceinline__ float min_image(float d, float L, float invL) {
  const float n = fmaf(d, invL, 12582912.0f) - 12582912.0f;
  return fmaf(-L, n, d);
}

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ V3 xyz(float4 p) { return {p.x, p.y, p.z}; }

// ------------------------------------------------------------------------------------------------
// pair terms
// ------------------------------------------------------------------------------------------------
// MUFU.RCP, ~1 ulp; the IEEE-rounded __frcp_rn costs a Newton step and a fix-up path on top
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// WCA (M.py:74-76).  tab = {24 eps, sigma^2, rc^2, eps}.  Returns F/r; r2 must be < tab.z.
__device__ __forceinline__ float wca_force_over_r(float4 tab, float r2) {
  const float inv_r2 = fast_rcp(r2);
  const float s2 = tab.y * inv_r2, s6 = s2 * s2 * s2;
  return tab.x * s6 * fmaf(2.0f, s6, -1.0f) * inv_r2;
}
// energy 4 eps [s^12 - s^6 + 1/4]
__device__ __forceinline__ float wca_energy(float4 tab, float r2) {
  if (r2 >= tab.z) return 0.0f;
  const float s2 = tab.y * fast_rcp(r2), s6 = s2 * s2 * s2;
  return 4.0f * tab.w * (fmaf(s6, s6, -s6) + 0.25f);
}
// Debye-Hueckel (M.py:97-107): e = exp(-kappa r)/r ; F/r = e (1 + kappa r)/r^2 ; both without the
// pref*qi*qj factor.  neg_kappa_log2e = -kappa*log2(e).
__device__ __forceinline__ float dh_energy_unit(float r2, float kappa, float neg_kappa_log2e, float *f_over_r) {
  const float rinv = rsqrtf(r2), r = r2 * rinv;
  const float e = fast_ex2(neg_kappa_log2e * r) * rinv;
  *f_over_r = e * fmaf(kappa, r, 1.0f) * rinv * rinv;
  return e;
}

// F/r of the same term times q_j, written for the inner loop of the MD force phase: the only constant
// is u's factor (1 + kappa r = 1 - u ln 2 with u = -kappa log2(e) r), the two MUFU results meet late.
__device__ __forceinline__ float dh_force_unit_q(float r2, float neg_kappa_log2e, float qj) {
  const float rinv = rsqrtf(r2), r = r2 * rinv, u = neg_kappa_log2e * r;
  const float w = fast_ex2(u) * ((rinv * qj) * (rinv * rinv));
  return fmaf(w * u, -0.693147180559945f, w);
}

// 16-byte shared-memory load from a 32-bit shared-space address
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// ------------------------------------------------------------------------------------------------
// bonded terms (unfolded coordinates).  Each returns the energy; forces by reference.
// ------------------------------------------------------------------------------------------------
// HarmonicBond / FeneBond (M.py:34-40): force on a of the bond (a,b)
__device__ __forceinline__ float bond_term(const DevParams &p, V3 a, V3 b, V3 *fa, int *err) {
  const V3 d = a - b;
  const float r2 = dot(d, d), rinv = rsqrtf(fmaxf(r2, 1e-30f)), r = r2 * rinv, dr = r - p.bond_r0;
  float e, fmag;
  if (p.bond_kind == CGPE_BOND_HARMONIC) {
    e = 0.5f * p.bond_k * dr * dr;
    fmag = -p.bond_k * dr;
  } else {
    const float arg = 1.0f - dr * dr / p.fene_drmax2;
    if (arg <= 0.0f) { *err |= kErrFene; }
    const float a_ = fmaxf(arg, 1e-12f);
    e = -0.5f * p.bond_k * p.fene_drmax2 * logf(a_);
    fmag = -p.bond_k * dr / a_;
  }
  *fa = (fmag * rinv) * d;
  return e;
}

// AngleHarmonic 1/2 k (phi-phi0)^2, phi at the middle bead (M.py:42-46,163-165)
__device__ __forceinline__ float angle_term(const DevParams &p, V3 pl, V3 pm, V3 pr, V3 *fl, V3 *fm, V3 *fr) {
  const V3 a = pl - pm, b = pr - pm;
  const float ia = rsqrtf(dot(a, a)), ib = rsqrtf(dot(b, b));
  const V3 ah = ia * a, bh = ib * b, n = cross(ah, bh);
  const float c = dot(ah, bh), s = sqrtf(dot(n, n));
  const float dphi = atan2f(s, c) - p.angle_phi0;
  const float pref = p.angle_k * dphi / fmaxf(s, 1e-9f);
  *fl = (pref * ia) * (bh - c * ah);
  *fr = (pref * ib) * (ah - c * bh);
  *fm = (-1.0f) * (*fl + *fr);
  return 0.5f * p.angle_k * dphi * dphi;
}

// Dihedral k [1 - cos(mult phi - phase)] on (p1,p2,p3,p4) (M.py:48-54,168-170).  cos/sin(mult phi)
// come from cos phi, sin phi by the multiple-angle recurrence: no trigonometric call.
__device__ __forceinline__ float dihedral_term(const DevParams &p, V3 p1, V3 p2, V3 p3, V3 p4,
                                               V3 *f1, V3 *f2, V3 *f3, V3 *f4) {
  const V3 b1 = p2 - p1, b2 = p3 - p2, b3 = p4 - p3;
  const V3 m = cross(b1, b2), n = cross(b2, b3);
  const float m2 = dot(m, m), n2 = dot(n, n), l2sq = dot(b2, b2);
  if (m2 <= 1e-8f || n2 <= 1e-8f) {   // |b1 x b2| <= 1e-4: torsion undefined (ESPResSo's guard)
    *f1 = *f2 = *f3 = *f4 = V3{0.f, 0.f, 0.f};
    return 0.0f;
  }
  // gradient evaluated with sin(bond angle) >= dihedral_sin_min (see cgpe.h)
  const float m2g = fmaxf(m2, p.dih_sin_min2 * dot(b1, b1) * l2sq), n2g = fmaxf(n2, p.dih_sin_min2 * dot(b3, b3) * l2sq);
  const float l2 = sqrtf(l2sq), inv_mn = rsqrtf(m2 * n2);
  const float c1 = dot(m, n) * inv_mn, s1 = l2 * dot(b1, n) * inv_mn;   // cos phi, sin phi
  float cm = 1.0f, sm = 0.0f;                                            // cos/sin(mult phi)
  for (int k = 0; k < p.dih_mult; ++k) {
    const float t = cm * c1 - sm * s1;
    sm = fmaf(sm, c1, cm * s1);
    cm = t;
  }
  const float cos_arg = fmaf(cm, p.dih_cos_phase, sm * p.dih_sin_phase);
  const float sin_arg = fmaf(sm, p.dih_cos_phase, -cm * p.dih_sin_phase);
  const float dedphi = p.dih_k * (float)p.dih_mult * sin_arg;
  const V3 g1 = (-l2 * fast_rcp(m2g)) * m, g4 = (l2 * fast_rcp(n2g)) * n;   // dphi/dp1, dphi/dp4
  const float inv_l2sq = fast_rcp(l2sq), s12 = dot(b1, b2) * inv_l2sq, s32 = dot(b3, b2) * inv_l2sq;
  const V3 g2 = (-1.0f - s12) * g1 + s32 * g4;
  const V3 g3 = s12 * g1 + (-1.0f - s32) * g4;
  *f1 = (-dedphi) * g1; *f2 = (-dedphi) * g2; *f3 = (-dedphi) * g3; *f4 = (-dedphi) * g4;
  return p.dih_k * (1.0f - cos_arg);
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Bounding spheres of the blocks of 32 consecutive particles (sph[b] = centre, .w = radius with a safety margin for FP32
// rounding).  Chain neighbours are spatial neighbours, so a list build tests a block's sphere before its 32 members and skips
// the block when even the nearest point of the sphere is out of range - with minimum-image arithmetic where the caller uses
// it: |x_i - x_m| >= |x_i - c_b| - rho_b holds image by image.  All threads of the block call it (nt a multiple of 32); the
// caller puts barriers around it.  Rows keep their order (partners ascending), so results do not change.
__device__ __forceinline__ void block_spheres(const float4 *pos, int P, float4 *sph, int tid, int nt) {
  const int lane = tid & 31, nb = (P + 31) >> 5;
  for (int b = tid >> 5; b < nb; b += nt >> 5) {
    const int j = b * 32 + lane;
    const bool in = j < P;
    const float4 x = in ? pos[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float inv = 1.0f / (float)min(32, P - b * 32);
    const float cx = warp_sum(x.x) * inv, cy = warp_sum(x.y) * inv, cz = warp_sum(x.z) * inv;
    const float ex = x.x - cx, ey = x.y - cy, ez = x.z - cz;
    float d2 = in ? fmaf(ex, ex, fmaf(ey, ey, ez * ez)) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, o));
    if (lane == 0) sph[b] = make_float4(cx, cy, cz, sqrtf(d2) * 1.0001f + 1.0e-3f);
  }
}

}  // namespace cgpe
