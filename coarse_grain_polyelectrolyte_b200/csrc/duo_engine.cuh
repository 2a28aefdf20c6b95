// duo_engine.cuh — interface of the cluster engine (duo_engine.cu): one replica = one thread-block
// cluster of two CTAs on two SMs, each CTA owning half of the particles.
#pragma once
#include "cgpe_device.cuh"

namespace cgpe {

// Launch shape and shared-memory layout of one CTA of the pair (identical in both CTAs, so that a
// shared-memory offset names the same array in the peer: mapa / map_shared_rank need nothing else).
struct DuoShape {
  int ok;          // 0: this handle cannot use the cluster engine
  int half;        // CTA 0 owns particles [0, half), CTA 1 owns [half, P)
  int nt;          // threads per CTA (one particle per thread)
  int n_loc;       // largest number of chargeable particles one CTA owns
  int G;           // sub-threads per Debye-Hueckel row
  int cap_w;       // WCA row capacity
  int dw;          // words per row of the Debye-Hueckel bit matrix
  int dhc_cap, dhc_stride;   // compact partner lists: entries per sub-thread, sub-threads per CTA (padded)
  int n_sm;        // SMs of the device (replica -> cluster placement heuristic)
  int *order;      // [R] device: replica taken by each cluster (k_duo_plan, in front of every launch)
  // byte offsets
  int pos, tab, red, slot, fd, dhn, cnt, flag, wl, db, qm, chg, cidx, lp, ld, wn, type, phi, wc, rxof, wrange, wflag, ratio, ord, qloc;
  int total;
};

// Fills the layout for the shapes in p (p.n_chg, p.P, ... must be final); n_loc = chargeable
// particles of the fuller half.  Returns the bytes of dynamic shared memory per CTA.
size_t duo_layout(const DevParams &p, int n_loc, size_t smem_budget, int cap_w_cfg, DuoShape *d);

cudaError_t duo_plan_waves(const DevParams &p, const DevState &g, int *order, const cgpe_sample_params &sp, uint32_t thr1,
                           uint32_t thr2, cudaStream_t st);
cudaError_t duo_md_run(const DevParams &p, const DevState &g, const DuoShape &d, int n_steps, cudaStream_t st);
cudaError_t duo_mc_run(const DevParams &p, const DevState &g, const DuoShape &d, int m, int n_trials,
                       cgpe_trial_record *trace, cudaStream_t st);
cudaError_t duo_md_forces(const DevParams &p, const DevState &g, const DuoShape &d, float4 *out, cudaStream_t st);
cudaError_t duo_sample_loop(const DevParams &p, const DevState &g, const DuoShape &d, const cgpe_sample_params &sp,
                            uint32_t thr1, uint32_t thr2, double *obs, float *qdist, cudaStream_t st);

}  // namespace cgpe
