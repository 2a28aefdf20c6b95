// wide_engine.cuh — interface of the global-memory engine for large replicas (wide_engine.cu).
#pragma once
#include "chain_engine.cuh"

namespace cgpe {

// Cell grid of one particle set per replica: members sorted by cell (counting sort, ties by member
// index, so the order is reproducible), rebuilt together with the Verlet rows.
struct CellGrid {
  int *start;         // [R][m_max + 1] first slot of each bucket; start[nc^3] = number of members
  int *cursor;        // [R][m_max] counting / scatter cursor
  int *cell_of;       // [R][n] cell of each member (-1: hidden slot, not in the grid)
  int *tmp;           // [R][n] members by slot before the in-cell ordering
  int *order;         // [R][n] member index by slot
  float4 *spos;       // [R][n] member position (+ charge) by slot
  int *nc;            // [R] buckets per edge of the table in use (1: a single bucket, every member is a candidate)
  int *nf;            // [R] cells per box edge (folded onto the table modulo nc)
  int n, m_max, nc_max;
};

// device arrays the wide engine adds to DevState
struct WideState {
  float4 *ref;        // [R][P] positions at the last list build
  uint32_t *wl;       // [R][cap_w][P] WCA Verlet rows (particle ids), column per particle
  int *wn;            // [R][P]
  uint32_t *dl;       // [R][n_chg][cap_d] Debye-Hueckel rows (particle ids of chargeable particles), one contiguous row per site
  uint16_t *dlc;      // [R][n_chg][cap_d] the same rows by chargeable index (nullptr unless all chargeable particles fit one CTA's shared memory)
  uint16_t *dlq;      // [R][n_chg][cap_d] of those the partners that carry charge right now (wk_dh_compact)
  int *dnq;           // [R][n_chg]
  int *nowrap;        // [R] tile mode: 1 if no listed pair of chargeable particles can be closer through a periodic image (set at every row build)
  float4 *fdh;        // [R][n_chg] Debye-Hueckel force on each chargeable particle (wk_dh_forces)
  int *dn;            // [R][n_chg]
  int *cidx;          // [R][P] particle -> chargeable index, -1 = never charged
  int *flag;          // [R + 1] rows of replica r are stale; [R]: a particle left its skin/2 sphere (shared vote)
  int *nsteps;        // [R] MD steps of the current burst
  int *sched_max;     // [1] longest burst of the current sample
  int *step;          // [2] step index read by the kernels of a captured step (graph replay)
  float *phi;         // [R][n_chg] per-site Debye-Hueckel potential (see chain_engine.cu, mc_prepare)
  float *wcache;      // [R][n_chg] WCA type-swap energy
  uint8_t *rxof;      // [R][n_chg]
  uint8_t *wflag;     // [R][n_chg]
  CellGrid cw, cd;    // cells over all particles (edge >= WCA list radius) / over chargeable ones (edge >= DH list radius)
  uint8_t *stale;     // [R][P] ion slots inserted since the grids were built (trial moves through the grids)
  int cap_w, cap_d;
  int mc_grid_ok;     // ion moves of replicas too large for the mirrors take their sums from the cell grids
  size_t mc_stage_limit;   // largest set of trial-loop mirrors kept in shared memory (bytes)
  int cells_multi_min; // particles per replica from which the cell grids are built by many CTAs (wk_cells_count ...)
  int dh_rows_short;  // rows are expected to hold a few dozen entries at most: 8 lanes per site instead of a warp
};

// One captured block of MD steps (CUDA graph), re-used while the kernel arguments stay the same.
struct WideGraphCache {
  cudaGraphExec_t exec = nullptr;
  DevParams p_key;     // arguments baked into the captured kernels
  int chunk = 0;
};
void wide_graph_release(WideGraphCache *c);

cudaError_t wide_md_run(const DevParams &p, const DevState &g, const WideState &w, int n_steps, cudaStream_t st, int *launches,
                        WideGraphCache *gc);
cudaError_t wide_steepest_descent(const DevParams &p, const DevState &g, const WideState &w, int n_steps, float f_max,
                                  float gamma, float max_disp, cudaStream_t st, int *launches);
cudaError_t wide_md_forces(const DevParams &p, const DevState &g, const WideState &w, float4 *out, cudaStream_t st);
cudaError_t wide_mc_run(const DevParams &p, const DevState &g, const WideState &w, int m, int n_trials,
                        cgpe_trial_record *trace, cudaStream_t st, int *launches);
cudaError_t wide_list_stats(const DevParams &p, const DevState &g, const WideState &w, double *out, cudaStream_t st);
cudaError_t wide_sample_loop(const DevParams &p, const DevState &g, const WideState &w, const cgpe_sample_params &sp,
                             uint32_t thr1, uint32_t thr2, double *obs, float *qdist, double *tmp_d, int *tmp_i,
                             cudaStream_t st, int *launches, WideGraphCache *gc);

}  // namespace cgpe
