// chain_engine.cuh — interface of the shared-memory replica engine (chain_engine.cu).
#pragma once
#include "cgpe_device.cuh"

namespace cgpe {

// pointers into the dynamic shared memory of one CTA (= one replica)
struct ChainSmem {
  float4 *pos;        // [P] x,y,z,q
  float4 *tab;        // [(T+1)^2] WCA table
  double *red;        // [32][4] block-reduction scratch
  float *slot;        // [3 slots][3][NC] bonded hand-over slots
  float *fd;          // [dh_group][3][n_chg] partial Debye-Hueckel forces per chargeable particle
  uint16_t *dhn;      // [dh_group * n_chg] entries of each sub-thread's compact partner list (DevState::dhc)
  int *cnt;           // [n_rx][2] n_prot, n_deprot
  int *flag;          // [4]
  uint16_t *wl;       // [cap_w][P] WCA Verlet list (column per particle)
  uint32_t *dbits;    // [n_chg][dw] Debye-Hueckel Verlet list, one bit per chargeable partner
  uint32_t *qmask;    // [dw] chargeable particles that carry charge right now
  uint16_t *chg;      // [n_chg] chargeable particle ids
  uint16_t *cidx;     // [P] particle id -> chargeable index (0xFFFF = none)
  uint16_t *lp, *ld;  // [n_rx][n_chg] protonated / deprotonated bookkeeping lists
  unsigned char *wn;        // WCA list lengths
  unsigned char *type;      // [P]
  float *phi;               // [n_chg] Debye-Hueckel potential sum_j q_j exp(-kappa r)/r at each chargeable particle
  float *wcache;            // [n_chg] WCA energy change of the type swap prot -> deprot
  unsigned char *rxof;      // [n_chg] reaction the particle belongs to (255 = none)
  float4 *cpos;             // [n_chg] x,y,z,q of the chargeable particles, compact (no id indirection)
  unsigned char *wrange;    // [n_chg][2] first / one-past-last non-empty word of the Debye-Hueckel row
  unsigned char *wflag;     // [n_chg] 1 if a WCA partner of the particle is titratable itself
  uint16_t *ion_active, *ion_free;   // [n_ions] explicit-ion pools (particle ids)
  float4 *ion_vel;          // [n_ions] velocity hand-over for ions touched by a trial move (w = 1: pending)
  uint16_t *qlist;          // [n_chg] chargeable indices of the particles that carry charge right now (dense mode)
  float *ratio;             // [n_rx][n_chg+1] N/(N_sites - N + 1): the factorial factor of the acceptance
};

struct ChainLaunch {
  int items;      // particles per thread (0 = system too large for this engine)
  int threads;
  size_t smem;
};

size_t chain_smem_bytes(DevParams &p);   // lays the arrays out, fills p.off, returns the size
int chain_items(int P);
ChainLaunch chain_launch_shape(const DevParams &p);

cudaError_t launch_md_run(const DevParams &p, const DevState &g, const ChainLaunch &L, int n_steps, cudaStream_t st);
cudaError_t launch_steepest_descent(const DevParams &p, const DevState &g, const ChainLaunch &L, int n_steps,
                                    float f_max, float gamma, float max_disp, cudaStream_t st);
cudaError_t launch_mc_run(const DevParams &p, const DevState &g, const ChainLaunch &L, int m, int n_trials,
                          cgpe_trial_record *trace, cudaStream_t st);
cudaError_t launch_md_forces(const DevParams &p, const DevState &g, const ChainLaunch &L, float4 *out, cudaStream_t st);
cudaError_t launch_list_stats(const DevParams &p, const DevState &g, const ChainLaunch &L, double *out, cudaStream_t st);
cudaError_t launch_sample_loop(const DevParams &p, const DevState &g, const ChainLaunch &L,
                               const cgpe_sample_params &sp, uint32_t thr1, uint32_t thr2, double *obs,
                               float *qdist, cudaStream_t st);

// all-pairs analysis kernels (diag_kernels.cu)
cudaError_t launch_energy(const DevParams &p, const DevState &g, double *out /*[R][5]*/, cudaStream_t st);
cudaError_t launch_forces(const DevParams &p, const DevState &g, float4 *out /*[R][P]*/, cudaStream_t st);
cudaError_t launch_min_dist(const DevParams &p, const DevState &g, double *out /*[R]*/, cudaStream_t st);
cudaError_t launch_observables(const DevParams &p, const DevState &g, double *rg, double *ree, int *counts,
                               cudaStream_t st);

}  // namespace cgpe
