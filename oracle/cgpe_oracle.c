/*
 * cgpe_oracle.c — CPU restatement (FP64, scalar C) of the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (coarse_grain_polyelectrolyte_b200/, src/)
 * may import, link or execute this file; it is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in ESPResSo (module
 * `espressomd`, version not pinned by the reference; API generation 4.2.x), which is not part of
 * /root/reference, is not installed and cannot be fetched.  The reference holds no golden vector
 * for any energy, force, acceptance or observable.  The functional forms below are the published
 * ESPResSo 4.2 ones restated from its documentation; they are anchored on analytic two-body
 * values, finite differences, the ideal titration curve (src/modules/utils.py:15-16) and the
 * reference's own call sites (cited per function, relative to /root/reference):
 *   P.py = src/modules/Properties.py  M.py = src/modules/Model.py  E.py = src/modules/EnsembleDynamics.py
 *
 * The exported orc_* functions mirror include/cgpe.h one to one so the same ctypes binding can
 * drive either library from the tests.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cgpe.h"
#include <pthread.h>

#define ORC_LN10 2.302585092994045684

/* ------------------------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator both sides draw from.      */
/* ------------------------------------------------------------------------------------------- */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
  philox4x32_10(ctr, key, out);
}

/* stream tags (word 3 of the counter) */
#define TAG_LANGEVIN 0x4C414E47u /* 'LANG' */
#define TAG_MC 0x4D430000u       /* 'MC' + reaction id */
#define TAG_SCHEDULE 0x53434844u /* 'SCHD' */

/* uniform in [0,1) on a 2^-24 grid: exact in both float and double */
static double u24(uint32_t x) { return (double)(x >> 8) * (1.0 / 16777216.0); }
/* zero-mean uniform in (-1/2, 1/2) on a 2^-23 grid, exact in float */
static double centred_u23(uint32_t x) {
  return ((double)(int32_t)(x >> 9) - 4194304.0) * (1.0 / 8388608.0) + (1.0 / 16777216.0);
}
static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* ------------------------------------------------------------------------------------------- */
typedef struct rx_lists {
  int32_t *prot;   /* particle ids currently of type_prot, in bookkeeping order */
  int32_t *deprot; /* particle ids currently of type_deprot */
  int32_t n_prot, n_deprot;
} rx_lists;

struct orc_handle {
  cgpe_config cfg;
  int R, P, T, NC; /* replicas, particles, types, chain beads total */
  double *x, *v, *f, *q; /* [R][P][3], [R][P] */
  uint8_t *type;         /* [R][P] */
  double *pH, *kappa, *rcut, *kT, *gamma; /* [R] */
  double *time;          /* [R] */
  int64_t *md_counter;   /* [R] Langevin noise counter */
  int64_t *md_steps;     /* [R] */
  int64_t *trials, *accepted; /* [R][n_reactions] */
  int64_t *samples_done; /* [R] */
  int *f_valid;          /* [R] */
  rx_lists *lists;       /* [R][n_reactions] */
  int32_t *ion_active, *ion_free, *n_active, *n_free; /* explicit ions: [R][n_ion_slots] pools, [R] counts */
  int32_t thermostat_seed;
  double force_cap;
  int have_state;
  double *last_obs; float *last_qdist; int last_ns, last_qrows;
  /* optional Verlet lists (performance mode for the CPU baseline; results are identical up to
     summation order).  Full rows: wl row i holds every j within rc_wca_max+skin, dl row i (only
     chargeable i) every chargeable j within r_cut+skin. */
  int use_lists; double skin;
  int **wl, **wn, **dl, **dn; int *cap_wl, *cap_dl;   /* per replica: [P][cap], [P]; rows grow on demand */
  double *ref;                                  /* [R][P][3] positions at the last build */
  int *lists_valid;                             /* [R] */
  uint8_t *chargeable;                          /* [R][P] */
  int64_t *rebuilds;                            /* [R] */
  char err[512];
};
typedef struct orc_handle orc_handle;

static char g_err[512];

static int fail(orc_handle *h, int code, const char *msg) {
  snprintf(h ? h->err : g_err, 512, "%s", msg);
  return code;
}

int orc_abi_version(void) { return CGPE_ABI_VERSION; }
/* threads used by the loops over replicas (replicas are independent); 1 = scalar port */
static int g_threads = 1;
void orc_set_threads(int32_t n) { g_threads = n > 0 ? n : 1; }

typedef void (*replica_fn)(void *ctx, int r);
typedef struct par_job { replica_fn fn; void *ctx; int R; int next; pthread_mutex_t mu; } par_job;
static void *par_worker(void *arg) {
  par_job *j = (par_job *)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    int r = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (r >= j->R) return NULL;
    j->fn(j->ctx, r);
  }
}
static void for_each_replica(int R, replica_fn fn, void *ctx) {
  int nt = g_threads < R ? g_threads : R;
  if (nt <= 1) { for (int r = 0; r < R; ++r) fn(ctx, r); return; }
  par_job j = {fn, ctx, R, 0, PTHREAD_MUTEX_INITIALIZER};
  pthread_t th[256];
  if (nt > 256) nt = 256;
  for (int t = 0; t < nt; ++t) pthread_create(&th[t], NULL, par_worker, &j);
  for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}
const char *orc_last_error(const orc_handle *h) { return h ? h->err : g_err; }

/* ------------------------------------------------------------------------------------------- */
/* Pair potentials                                                                              */
/* ------------------------------------------------------------------------------------------- */
static inline double min_image(double d, double L) { return d - L * nearbyint(d / L); }

/* WCA (M.py:74-76): U = 4 eps [(s/r)^12 - (s/r)^6 + 1/4] for r < 2^(1/6) s.
 * Returns the energy, *fr = -dU/dr / r  (force on i = fr * (ri - rj)). */
static inline double wca(double eps, double sig, double r2, double *fr) {
  *fr = 0.0;
  if (eps == 0.0 || sig == 0.0) return 0.0;
  double rc2 = 1.2599210498948731648 * sig * sig; /* 2^(1/3) sig^2 */
  if (r2 >= rc2) return 0.0;
  double s2 = sig * sig / r2, s6 = s2 * s2 * s2;
  *fr = 48.0 * eps * s6 * (s6 - 0.5) / r2;
  return 4.0 * eps * (s6 * s6 - s6 + 0.25);
}

/* Debye-Hueckel (M.py:97-107): U = pref qi qj exp(-kappa r)/r for r < r_cut, unshifted. */
static inline double dh(double pref_qq, double kappa, double rcut, double r2, double *fr) {
  *fr = 0.0;
  if (pref_qq == 0.0 || r2 >= rcut * rcut) return 0.0;
  double r = sqrt(r2);
  double e = pref_qq * exp(-kappa * r) / r;
  *fr = e * (1.0 + kappa * r) / r2;
  return e;
}

static inline const double *X(const orc_handle *h, int r, int i) {
  return h->x + ((size_t)r * h->P + i) * 3;
}

/* ------------------------------------------------------------------------------------------- */
/* Bonded terms on chain c (beads c*N .. c*N+N-1), unfolded coordinates                         */
/* ------------------------------------------------------------------------------------------- */
static double bond_energy_force(const cgpe_config *c, const double *a, const double *b, double *fa) {
  /* harmonic 1/2 k (r-r0)^2 ; FENE -1/2 k drmax^2 ln(1-((r-r0)/drmax)^2)  (M.py:34-40) */
  double d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
  double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  double dr = r - c->bond_r0, e, fmag; /* fmag = -dU/dr */
  if (c->bond_kind == CGPE_BOND_HARMONIC) {
    e = 0.5 * c->bond_k * dr * dr;
    fmag = -c->bond_k * dr;
  } else {
    double m2 = c->fene_drmax * c->fene_drmax;
    double arg = 1.0 - dr * dr / m2;
    if (arg <= 0.0) { e = INFINITY; fmag = NAN; }
    else { e = -0.5 * c->bond_k * m2 * log(arg); fmag = -c->bond_k * dr / arg; }
  }
  for (int k = 0; k < 3; ++k) fa[k] = (r > 0.0) ? fmag * d[k] / r : 0.0;
  return e;
}

static void cross(const double *a, const double *b, double *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* AngleHarmonic 1/2 bend (phi-phi0)^2 with phi at the middle bead (M.py:42-46,163-165). */
static double angle_energy_force(const cgpe_config *c, const double *pl, const double *pm,
                                 const double *pr, double *fl, double *fm, double *fr_) {
  double a[3], b[3], n[3];
  for (int k = 0; k < 3; ++k) { a[k] = pl[k] - pm[k]; b[k] = pr[k] - pm[k]; }
  double la = sqrt(dot(a, a)), lb = sqrt(dot(b, b));
  cross(a, b, n);
  double cosv = dot(a, b) / (la * lb), sinv = sqrt(dot(n, n)) / (la * lb);
  double phi = atan2(sinv, cosv), dphi = phi - c->angle_phi0;
  double e = 0.5 * c->angle_k * dphi * dphi;
  double s = sinv < 1e-9 ? 1e-9 : sinv;
  double pref = c->angle_k * dphi / s; /* F_l = pref * d(cos)/d r_l */
  for (int k = 0; k < 3; ++k) {
    double ah = a[k] / la, bh = b[k] / lb;
    fl[k] = pref * (bh - cosv * ah) / la;
    fr_[k] = pref * (ah - cosv * bh) / lb;
    fm[k] = -fl[k] - fr_[k];
  }
  return e;
}

/* Dihedral bend [1 - cos(mult phi - phase)] on (p1,p2,p3,p4) (M.py:48-54,168-170);
 * phi is the IUPAC torsion (trans = pi), i.e. the angle between the normals b1xb2 and b2xb3. */
static double dihedral_energy_force(const cgpe_config *c, const double *p1, const double *p2,
                                    const double *p3, const double *p4, double *f1, double *f2,
                                    double *f3, double *f4) {
  double b1[3], b2[3], b3[3], m[3], n[3];
  for (int k = 0; k < 3; ++k) { b1[k] = p2[k] - p1[k]; b2[k] = p3[k] - p2[k]; b3[k] = p4[k] - p3[k]; }
  cross(b1, b2, m);
  cross(b2, b3, n);
  double m2 = dot(m, m), n2 = dot(n, n), l2sq = dot(b2, b2), l2 = sqrt(l2sq);
  for (int k = 0; k < 3; ++k) f1[k] = f2[k] = f3[k] = f4[k] = 0.0;
  if (m2 <= 1e-8 || n2 <= 1e-8) return 0.0; /* |b1 x b2| <= 1e-4: torsion undefined, no force (ESPResSo's guard) */
  /* gradient evaluated with sin(theta) >= dihedral_sin_min (see cgpe.h) */
  const double smin = c->dihedral_sin_min == 0.0 ? 0.02 : (c->dihedral_sin_min < 0.0 ? 0.0 : c->dihedral_sin_min);
  const double m2g = fmax(m2, smin * smin * dot(b1, b1) * l2sq), n2g = fmax(n2, smin * smin * dot(b3, b3) * l2sq);
  double phi = atan2(l2 * dot(b1, n), dot(m, n));
  double arg = c->dihedral_mult * phi - c->dihedral_phase;
  double e = c->dihedral_k * (1.0 - cos(arg));
  double dedphi = c->dihedral_k * c->dihedral_mult * sin(arg);
  /* dphi/dp1 = -|b2|/|m|^2 m ; dphi/dp4 = +|b2|/|n|^2 n */
  double g1[3], g4[3];
  for (int k = 0; k < 3; ++k) { g1[k] = -l2 / m2g * m[k]; g4[k] = l2 / n2g * n[k]; }
  double s12 = dot(b1, b2) / l2sq, s32 = dot(b3, b2) / l2sq;
  for (int k = 0; k < 3; ++k) {
    double g2 = -g1[k] - s12 * g1[k] + s32 * g4[k];
    double g3 = -g4[k] + s12 * g1[k] - s32 * g4[k];
    f1[k] = -dedphi * g1[k];
    f2[k] = -dedphi * g2;
    f3[k] = -dedphi * g3;
    f4[k] = -dedphi * g4[k];
  }
  return e;
}

/* ------------------------------------------------------------------------------------------- */
/* Full interaction energy / forces of replica r, all pairs (R3, R4)                            */
/* f may be NULL.  e[0..2] = coulomb, bonded, non_bonded                                        */
/* ------------------------------------------------------------------------------------------- */
static void nonbonded_listed(orc_handle *h, int r, double *f, double e[3]);

static void interactions(const orc_handle *h, int r, double *f, double e[3]) {
  const cgpe_config *c = &h->cfg;
  const int P = h->P, T = h->T;
  const uint8_t *ty = h->type + (size_t)r * P;
  const double *q = h->q + (size_t)r * P;
  const double L = c->box_l;
  e[0] = e[1] = e[2] = 0.0;
  if (f) memset(f, 0, sizeof(double) * 3 * P);
  /* bonded */
  for (int ch = 0; ch < c->n_chains; ++ch) {
    const int N = c->n_beads, o = ch * N;
    for (int i = 1; i < N; ++i) {
      if (c->bond_kind != CGPE_BOND_NONE) { /* bond (i-1,i): M.py:159-160 */
        double fa[3];
        e[1] += bond_energy_force(c, X(h, r, o + i), X(h, r, o + i - 1), fa);
        if (f) for (int k = 0; k < 3; ++k) { f[(o + i) * 3 + k] += fa[k]; f[(o + i - 1) * 3 + k] -= fa[k]; }
      }
      if (c->use_angle && i < N - 1) { /* M.py:161-165 */
        double fl[3], fm[3], fr_[3];
        e[1] += angle_energy_force(c, X(h, r, o + i - 1), X(h, r, o + i), X(h, r, o + i + 1), fl, fm, fr_);
        if (f) for (int k = 0; k < 3; ++k) {
          f[(o + i - 1) * 3 + k] += fl[k]; f[(o + i) * 3 + k] += fm[k]; f[(o + i + 1) * 3 + k] += fr_[k];
        }
      }
      if (c->use_dihedral && i < N - 2) { /* M.py:166-170 */
        double f1[3], f2[3], f3[3], f4[3];
        e[1] += dihedral_energy_force(c, X(h, r, o + i - 1), X(h, r, o + i), X(h, r, o + i + 1),
                                      X(h, r, o + i + 2), f1, f2, f3, f4);
        if (f) for (int k = 0; k < 3; ++k) {
          f[(o + i - 1) * 3 + k] += f1[k]; f[(o + i) * 3 + k] += f2[k];
          f[(o + i + 1) * 3 + k] += f3[k]; f[(o + i + 2) * 3 + k] += f4[k];
        }
      }
    }
  }
  /* non-bonded: every pair, bonded neighbours included (the reference sets no exclusions) */
  if (!c->use_wca && !c->use_dh) return;
  if (h->use_lists) { nonbonded_listed((orc_handle *)h, r, f, e); return; }
  for (int i = 0; i < P; ++i) {
    if (ty[i] >= T) continue; /* non-interacting type */
    const double *xi = X(h, r, i);
    for (int j = i + 1; j < P; ++j) {
      if (ty[j] >= T) continue;
      const double *xj = X(h, r, j);
      double d[3] = {min_image(xi[0] - xj[0], L), min_image(xi[1] - xj[1], L), min_image(xi[2] - xj[2], L)};
      double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], fr = 0.0, fr2;
      if (c->use_wca) {
        e[2] += wca(c->wca_eps[ty[i] * CGPE_MAX_TYPES + ty[j]], c->wca_sig[ty[i] * CGPE_MAX_TYPES + ty[j]], r2, &fr2);
        fr += fr2;
      }
      if (c->use_dh) {
        e[0] += dh(c->dh_prefactor * q[i] * q[j], h->kappa[r], h->rcut[r], r2, &fr2);
        fr += fr2;
      }
      if (f && fr != 0.0) for (int k = 0; k < 3; ++k) { f[i * 3 + k] += fr * d[k]; f[j * 3 + k] -= fr * d[k]; }
    }
  }
}

/* ------------------------------------------------------------------------------------------- */
/* Verlet-list mode: the same sums as interactions(), restricted to listed pairs                 */
/* ------------------------------------------------------------------------------------------- */
static double wca_rc_max(const orc_handle *h) {
  double m = 0.0;
  for (int a = 0; a < h->T; ++a) for (int b = 0; b < h->T; ++b) {
    double e = h->cfg.wca_eps[a * CGPE_MAX_TYPES + b], s = h->cfg.wca_sig[a * CGPE_MAX_TYPES + b];
    if (e != 0.0 && s != 0.0 && 1.122462048309373 * s > m) m = 1.122462048309373 * s;
  }
  return m;
}

static void build_lists(orc_handle *h, int r) {
  const int P = h->P;
  const double L = h->cfg.box_l;
  const double rw = wca_rc_max(h) + h->skin, rw2 = rw * rw;
  const double rd = h->rcut[r] + h->skin, rd2 = rd * rd;
  const uint8_t *chg = h->chargeable + (size_t)r * P;
  for (;;) {   /* restart with longer rows when one overflows */
    int *wl = h->wl[r], *wn = h->wn[r], *dl = h->dl[r], *dn = h->dn[r];
    const int cw = h->cap_wl[r], cd = h->cap_dl[r];
    int overflow_w = 0, overflow_d = 0;
    memset(wn, 0, sizeof(int) * P);
    memset(dn, 0, sizeof(int) * P);
    for (int i = 0; i < P && !overflow_w && !overflow_d; ++i) {
      const double *xi = X(h, r, i);
      for (int j = i + 1; j < P; ++j) {
        const double *xj = X(h, r, j);
        double d0 = min_image(xi[0] - xj[0], L), d1 = min_image(xi[1] - xj[1], L), d2 = min_image(xi[2] - xj[2], L);
        double r2 = d0 * d0 + d1 * d1 + d2 * d2;
        if (h->cfg.use_wca && r2 < rw2) {
          if (wn[i] >= cw || wn[j] >= cw) { overflow_w = 1; break; }
          wl[(size_t)i * cw + wn[i]++] = j;
          wl[(size_t)j * cw + wn[j]++] = i;
        }
        if (h->cfg.use_dh && chg[i] && chg[j] && r2 < rd2) {
          if (dn[i] >= cd || dn[j] >= cd) { overflow_d = 1; break; }
          dl[(size_t)i * cd + dn[i]++] = j;
          dl[(size_t)j * cd + dn[j]++] = i;
        }
      }
    }
    if (!overflow_w && !overflow_d) break;
    if (overflow_w) { h->cap_wl[r] *= 2; free(h->wl[r]); h->wl[r] = malloc(sizeof(int) * (size_t)P * h->cap_wl[r]); }
    if (overflow_d) { h->cap_dl[r] *= 2; free(h->dl[r]); h->dl[r] = malloc(sizeof(int) * (size_t)P * h->cap_dl[r]); }
  }
  memcpy(h->ref + (size_t)r * P * 3, h->x + (size_t)r * P * 3, sizeof(double) * 3 * P);
  h->lists_valid[r] = 1;
  h->rebuilds[r] += 1;
}

static void ensure_lists(orc_handle *h, int r) {
  if (!h->lists_valid[r]) { build_lists(h, r); return; }
  const double *x = h->x + (size_t)r * h->P * 3, *ref = h->ref + (size_t)r * h->P * 3;
  const double lim = 0.25 * h->skin * h->skin;
  for (int i = 0; i < h->P; ++i) {
    double a = x[i * 3] - ref[i * 3], b = x[i * 3 + 1] - ref[i * 3 + 1], c = x[i * 3 + 2] - ref[i * 3 + 2];
    if (!(a * a + b * b + c * c <= lim)) { build_lists(h, r); return; }
  }
}

/* non-bonded part over listed pairs (each unique pair once: j > i) */
static void nonbonded_listed(orc_handle *h, int r, double *f, double e[3]) {
  const cgpe_config *c = &h->cfg;
  const int P = h->P, T = h->T;
  const uint8_t *ty = h->type + (size_t)r * P;
  const double *q = h->q + (size_t)r * P;
  const double L = c->box_l;
  ensure_lists(h, r);
  const int *wl = h->wl[r], *wn = h->wn[r], *dl = h->dl[r], *dn = h->dn[r];
  const int cw = h->cap_wl[r], cd = h->cap_dl[r];
  for (int i = 0; i < P; ++i) {
    if (ty[i] >= T) continue;
    const double *xi = X(h, r, i);
    if (c->use_wca)
      for (int k = 0; k < wn[i]; ++k) {
        const int j = wl[(size_t)i * cw + k];
        if (j < i || ty[j] >= T) continue;
        const double *xj = X(h, r, j);
        double d[3] = {min_image(xi[0] - xj[0], L), min_image(xi[1] - xj[1], L), min_image(xi[2] - xj[2], L)};
        double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], fr;
        e[2] += wca(c->wca_eps[ty[i] * CGPE_MAX_TYPES + ty[j]], c->wca_sig[ty[i] * CGPE_MAX_TYPES + ty[j]], r2, &fr);
        if (f && fr != 0.0) for (int a = 0; a < 3; ++a) { f[i * 3 + a] += fr * d[a]; f[j * 3 + a] -= fr * d[a]; }
      }
    if (c->use_dh && q[i] != 0.0)
      for (int k = 0; k < dn[i]; ++k) {
        const int j = dl[(size_t)i * cd + k];
        if (j < i || ty[j] >= T || q[j] == 0.0) continue;
        const double *xj = X(h, r, j);
        double d[3] = {min_image(xi[0] - xj[0], L), min_image(xi[1] - xj[1], L), min_image(xi[2] - xj[2], L)};
        double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], fr;
        e[0] += dh(c->dh_prefactor * q[i] * q[j], h->kappa[r], h->rcut[r], r2, &fr);
        if (f && fr != 0.0) for (int a = 0; a < 3; ++a) { f[i * 3 + a] += fr * d[a]; f[j * 3 + a] -= fr * d[a]; }
      }
  }
}

/* ------------------------------------------------------------------------------------------- */
/* lifetime / state                                                                             */
/* ------------------------------------------------------------------------------------------- */
static int check_config(const cgpe_config *c, char *err) {
  if (!c) { snprintf(err, 512, "config is NULL"); return CGPE_ERR_INVALID; }
  if (c->abi_version != CGPE_ABI_VERSION) { snprintf(err, 512, "abi_version %d != %d", c->abi_version, CGPE_ABI_VERSION); return CGPE_ERR_INVALID; }
  if (c->n_replicas < 1 || c->n_chains < 1 || c->n_beads < 1 || c->n_ion_slots < 0) { snprintf(err, 512, "n_replicas, n_chains, n_beads must be >= 1"); return CGPE_ERR_INVALID; }
  if (c->n_types < 1 || c->n_types > CGPE_MAX_TYPES) { snprintf(err, 512, "n_types out of range"); return CGPE_ERR_INVALID; }
  if (c->n_reactions < 0 || c->n_reactions > CGPE_MAX_REACTIONS) { snprintf(err, 512, "n_reactions out of range"); return CGPE_ERR_INVALID; }
  if (!(c->box_l > 0.0) || !(c->time_step > 0.0)) { snprintf(err, 512, "box_l and time_step must be positive"); return CGPE_ERR_INVALID; }
  if (c->use_dh && (!(c->r_cut > 0.0) || c->r_cut > 0.5 * c->box_l)) { snprintf(err, 512, "r_cut must lie in (0, box_l/2]"); return CGPE_ERR_INVALID; }
  return CGPE_OK;
}

/* Slots behind the chains hold explicit ions that the reactions create and destroy (M.py:173-198) --
   unless every reaction names no ion type (type_ion < 0): then the reactions are the implicit-ion
   ones and the slots hold spectator particles (a charged nanoparticle, BASELINE config 4). */
static int explicit_ions(const cgpe_config *c) {
  if (c->n_ion_slots <= 0) return 0;
  for (int m = 0; m < c->n_reactions; ++m)
    if (c->reactions[m].type_ion >= 0) return 1;
  return c->n_reactions == 0;
}

int orc_create(const cgpe_config *cfg, orc_handle **out) {
  int rc = check_config(cfg, g_err);
  if (rc) return rc;
  if (explicit_ions(cfg)) {
    for (int m = 0; m < cfg->n_reactions; ++m)
      if (cfg->reactions[m].type_ion != cfg->reactions[0].type_ion || cfg->reactions[m].q_ion != cfg->reactions[0].q_ion ||
          cfg->reactions[m].type_ion < 0 || cfg->reactions[m].type_ion >= cfg->n_types) {
        snprintf(g_err, 512, "explicit ions: all reactions must release the same ion type"); return CGPE_ERR_INVALID;
      }
  }
  orc_handle *h = (orc_handle *)calloc(1, sizeof(orc_handle));
  h->cfg = *cfg;
  h->R = cfg->n_replicas; h->T = cfg->n_types; h->NC = cfg->n_chains * cfg->n_beads;
  h->P = h->NC + cfg->n_ion_slots;
  size_t RP = (size_t)h->R * h->P;
  h->x = calloc(RP * 3, sizeof(double)); h->v = calloc(RP * 3, sizeof(double));
  h->f = calloc(RP * 3, sizeof(double)); h->q = calloc(RP, sizeof(double));
  h->type = calloc(RP, 1);
  h->pH = malloc(sizeof(double) * h->R); h->kappa = malloc(sizeof(double) * h->R);
  h->rcut = malloc(sizeof(double) * h->R); h->kT = malloc(sizeof(double) * h->R);
  h->gamma = malloc(sizeof(double) * h->R);
  h->time = calloc(h->R, sizeof(double));
  h->md_counter = calloc(h->R, sizeof(int64_t)); h->md_steps = calloc(h->R, sizeof(int64_t));
  h->samples_done = calloc(h->R, sizeof(int64_t));
  h->f_valid = calloc(h->R, sizeof(int));
  int nr = cfg->n_reactions > 0 ? cfg->n_reactions : 1;
  h->trials = calloc((size_t)h->R * nr, sizeof(int64_t));
  h->accepted = calloc((size_t)h->R * nr, sizeof(int64_t));
  h->lists = calloc((size_t)h->R * nr, sizeof(rx_lists));
  for (size_t k = 0; k < (size_t)h->R * nr; ++k) {
    h->lists[k].prot = malloc(sizeof(int32_t) * h->P);
    h->lists[k].deprot = malloc(sizeof(int32_t) * h->P);
  }
  for (int r = 0; r < h->R; ++r) {
    h->pH[r] = cfg->pH; h->kappa[r] = cfg->kappa; h->rcut[r] = cfg->r_cut;
    h->kT[r] = cfg->kT; h->gamma[r] = cfg->gamma;
  }
  h->thermostat_seed = 0;
  h->use_lists = 0; h->skin = 0.8;
  {
    const int ni = cfg->n_ion_slots > 0 ? cfg->n_ion_slots : 1;
    h->ion_active = calloc((size_t)h->R * ni, sizeof(int32_t)); h->ion_free = calloc((size_t)h->R * ni, sizeof(int32_t));
    h->n_active = calloc(h->R, sizeof(int32_t)); h->n_free = calloc(h->R, sizeof(int32_t));
  }
  h->lists_valid = calloc(h->R, sizeof(int));
  h->rebuilds = calloc(h->R, sizeof(int64_t));
  h->chargeable = calloc(RP, 1);
  *out = h;
  return CGPE_OK;
}

void orc_destroy(orc_handle *h) {
  if (!h) return;
  int nr = h->cfg.n_reactions > 0 ? h->cfg.n_reactions : 1;
  for (size_t k = 0; k < (size_t)h->R * nr; ++k) { free(h->lists[k].prot); free(h->lists[k].deprot); }
  free(h->lists); free(h->x); free(h->v); free(h->f); free(h->q); free(h->type);
  free(h->pH); free(h->kappa); free(h->rcut); free(h->kT); free(h->gamma); free(h->time);
  free(h->md_counter); free(h->md_steps); free(h->samples_done); free(h->f_valid);
  free(h->trials); free(h->accepted); free(h->last_obs); free(h->last_qdist);
  if (h->wl) for (int r = 0; r < h->R; ++r) { free(h->wl[r]); free(h->wn[r]); free(h->dl[r]); free(h->dn[r]); }
  free(h->wl); free(h->wn); free(h->dl); free(h->dn); free(h->cap_wl); free(h->cap_dl);
  free(h->ref); free(h->lists_valid); free(h->rebuilds); free(h->chargeable);
  free(h->ion_active); free(h->ion_free); free(h->n_active); free(h->n_free);
  free(h);
}

int orc_device_info(const orc_handle *h, int32_t *sm, int32_t *maj, int32_t *min, char *name, int32_t len) {
  (void)h;
  if (sm) *sm = 0;
  if (maj) *maj = 0;
  if (min) *min = 0;
  if (name && len > 0) snprintf(name, len, "cpu-oracle-fp64");
  return CGPE_OK;
}
int orc_set_stream(orc_handle *h, void *s) { (void)h; (void)s; return CGPE_OK; }
int orc_synchronize(orc_handle *h) { (void)h; return CGPE_OK; }

static void rebuild_lists(orc_handle *h) {
  /* bookkeeping lists in particle-id order (set_state contract) */
  if (explicit_ions(&h->cfg) && h->cfg.n_reactions > 0)
    for (int r = 0; r < h->R; ++r) {
      const int ni = h->cfg.n_ion_slots, tb = h->cfg.reactions[0].type_ion;
      h->n_active[r] = h->n_free[r] = 0;
      for (int i = h->NC; i < h->P; ++i) {
        const uint8_t t = h->type[(size_t)r * h->P + i];
        if (t == tb) h->ion_active[(size_t)r * ni + h->n_active[r]++] = i;
        else if (t >= h->T) h->ion_free[(size_t)r * ni + h->n_free[r]++] = i;
      }
    }
  for (int r = 0; r < h->R; ++r)
    for (int m = 0; m < h->cfg.n_reactions; ++m) {
      rx_lists *l = &h->lists[(size_t)r * h->cfg.n_reactions + m];
      const cgpe_reaction *rx = &h->cfg.reactions[m];
      l->n_prot = l->n_deprot = 0;
      for (int i = 0; i < h->P; ++i) {
        uint8_t t = h->type[(size_t)r * h->P + i];
        if (t == rx->type_prot) l->prot[l->n_prot++] = i;
        else if (t == rx->type_deprot) l->deprot[l->n_deprot++] = i;
      }
    }
}

int orc_set_state(orc_handle *h, const float *xyzq, const float *vel, const uint8_t *type) {
  if (!xyzq || !type) return fail(h, CGPE_ERR_INVALID, "xyzq and type are required");
  size_t RP = (size_t)h->R * h->P;
  for (size_t k = 0; k < RP; ++k) {
    for (int c = 0; c < 3; ++c) {
      h->x[k * 3 + c] = xyzq[k * 4 + c];
      h->v[k * 3 + c] = vel ? vel[k * 4 + c] : 0.0;
    }
    h->q[k] = xyzq[k * 4 + 3];
    if (type[k] > h->T) return fail(h, CGPE_ERR_INVALID, "type index out of range");
    h->type[k] = type[k];
    if (type[k] >= h->T) { h->v[k * 3] = h->v[k * 3 + 1] = h->v[k * 3 + 2] = 0.0; h->q[k] = 0.0; }
    else if (((unsigned)h->cfg.fixed_type_mask >> type[k]) & 1u) h->v[k * 3] = h->v[k * 3 + 1] = h->v[k * 3 + 2] = 0.0;
  }
  rebuild_lists(h);
  {
    unsigned reactive = 0;
    for (int m = 0; m < h->cfg.n_reactions; ++m) reactive |= (1u << h->cfg.reactions[m].type_prot) | (1u << h->cfg.reactions[m].type_deprot);
    for (size_t k = 0; k < RP; ++k)
      h->chargeable[k] = (h->q[k] != 0.0) || ((reactive >> h->type[k]) & 1u) || ((int)(k % h->P) >= h->NC);
  }
  for (int r = 0; r < h->R; ++r) { h->f_valid[r] = 0; h->lists_valid[r] = 0; }
  h->have_state = 1;
  return CGPE_OK;
}

int orc_get_state(orc_handle *h, float *xyzq, float *vel, uint8_t *type) {
  size_t RP = (size_t)h->R * h->P;
  for (size_t k = 0; k < RP; ++k) {
    if (xyzq) { for (int c = 0; c < 3; ++c) xyzq[k * 4 + c] = (float)h->x[k * 3 + c]; xyzq[k * 4 + 3] = (float)h->q[k]; }
    if (vel) { for (int c = 0; c < 3; ++c) vel[k * 4 + c] = (float)h->v[k * 3 + c]; vel[k * 4 + 3] = 0.f; }
    if (type) type[k] = h->type[k];
  }
  return CGPE_OK;
}

/* double-precision state access for the tests (the float path above rounds) */
int orc_get_state_f64(orc_handle *h, double *x, double *v, double *q) {
  size_t RP = (size_t)h->R * h->P;
  if (x) memcpy(x, h->x, RP * 3 * sizeof(double));
  if (v) memcpy(v, h->v, RP * 3 * sizeof(double));
  if (q) memcpy(q, h->q, RP * sizeof(double));
  return CGPE_OK;
}

/* double-precision position upload for finite-difference tests */
int orc_set_positions_f64(orc_handle *h, const double *x) {
  memcpy(h->x, x, sizeof(double) * 3 * (size_t)h->R * h->P);
  for (int r = 0; r < h->R; ++r) { h->f_valid[r] = 0; h->lists_valid[r] = 0; }
  return CGPE_OK;
}

/* Verlet-list mode on/off (performance only; used for the CPU baseline timing) */
int orc_set_list_mode(orc_handle *h, int32_t on, double skin) {
  h->use_lists = on != 0;
  if (skin > 0.0) h->skin = skin;
  if (on && !h->wl) {
    h->wl = calloc(h->R, sizeof(int *)); h->wn = calloc(h->R, sizeof(int *));
    h->dl = calloc(h->R, sizeof(int *)); h->dn = calloc(h->R, sizeof(int *));
    h->cap_wl = malloc(sizeof(int) * h->R); h->cap_dl = malloc(sizeof(int) * h->R);
    for (int r = 0; r < h->R; ++r) {
      h->cap_wl[r] = 32; h->cap_dl[r] = 64;
      h->wl[r] = malloc(sizeof(int) * (size_t)h->P * h->cap_wl[r]); h->wn[r] = calloc(h->P, sizeof(int));
      h->dl[r] = malloc(sizeof(int) * (size_t)h->P * h->cap_dl[r]); h->dn[r] = calloc(h->P, sizeof(int));
    }
    h->ref = calloc((size_t)h->R * h->P * 3, sizeof(double));
  }
  for (int r = 0; r < h->R; ++r) h->lists_valid[r] = 0;
  return CGPE_OK;
}

int orc_set_replica_params(orc_handle *h, const double *pH, const double *kappa, const double *rcut,
                           const double *kT, const double *gamma) {
  for (int r = 0; r < h->R; ++r) {
    if (pH) h->pH[r] = pH[r];
    if (kappa) h->kappa[r] = kappa[r];
    if (rcut) {
      if (!(rcut[r] > 0.0) || rcut[r] > 0.5 * h->cfg.box_l) return fail(h, CGPE_ERR_INVALID, "r_cut must lie in (0, box_l/2]");
      h->rcut[r] = rcut[r];
    }
    if (kT) h->kT[r] = kT[r];
    if (gamma) h->gamma[r] = gamma[r];
    if (kappa || rcut || kT || gamma) h->f_valid[r] = 0;
    if (rcut) h->lists_valid[r] = 0;
  }
  return CGPE_OK;
}

int orc_set_wca(orc_handle *h, const double *eps, const double *sig) {
  memcpy(h->cfg.wca_eps, eps, sizeof(h->cfg.wca_eps));
  memcpy(h->cfg.wca_sig, sig, sizeof(h->cfg.wca_sig));
  for (int r = 0; r < h->R; ++r) { h->f_valid[r] = 0; h->lists_valid[r] = 0; }
  return CGPE_OK;
}

int orc_set_thermostat_seed(orc_handle *h, int32_t seed) {
  if (seed >= 0) {
    h->thermostat_seed = seed;
    for (int r = 0; r < h->R; ++r) h->md_counter[r] = 0;
  }
  for (int r = 0; r < h->R; ++r) h->f_valid[r] = 0;
  return CGPE_OK;
}

int orc_set_force_cap(orc_handle *h, double cap) {
  h->force_cap = cap;
  for (int r = 0; r < h->R; ++r) h->f_valid[r] = 0;
  return CGPE_OK;
}

int orc_set_time(orc_handle *h, double t) { for (int r = 0; r < h->R; ++r) h->time[r] = t; return CGPE_OK; }
int orc_get_time(orc_handle *h, double *t) { memcpy(t, h->time, sizeof(double) * h->R); return CGPE_OK; }

/* ------------------------------------------------------------------------------------------- */
/* R2: velocity Verlet + Langevin thermostat (E.py:61-63; integrator.run E.py:90)               */
/* ------------------------------------------------------------------------------------------- */
/* fix=[True]*3 (cgpe_config::fixed_type_mask): particles of a fixed type feel no force, hold no
   velocity, get no thermostat noise; they still act on everything else. */
static void release_fixed(const orc_handle *h, int r, double *f) {
  const unsigned mask = (unsigned)h->cfg.fixed_type_mask;
  if (!mask) return;
  for (int i = 0; i < h->P; ++i)
    if ((mask >> h->type[(size_t)r * h->P + i]) & 1u) f[i * 3] = f[i * 3 + 1] = f[i * 3 + 2] = 0.0;
}

static void total_force(orc_handle *h, int r, int with_thermostat) {
  /* f = F_interactions + ( -gamma v + sqrt(24 gamma kT / dt) (U - 1/2) ), then the force cap */
  double e[3];
  double *f = h->f + (size_t)r * h->P * 3;
  interactions(h, r, f, e);
  const double dt = h->cfg.time_step;
  if (with_thermostat) {
    const double g = h->gamma[r], pref = sqrt(24.0 * g * h->kT[r] / dt);
    const double *v = h->v + (size_t)r * h->P * 3;
    uint32_t key[2] = {(uint32_t)h->thermostat_seed, (uint32_t)(h->cfg.replica_offset + r)};
    uint64_t cnt = (uint64_t)h->md_counter[r];
    for (int i = 0; i < h->P; ++i) {
      if (h->type[(size_t)r * h->P + i] >= h->T) continue;
      uint32_t ctr[4] = {(uint32_t)cnt, (uint32_t)(cnt >> 32), (uint32_t)i, TAG_LANGEVIN}, o[4];
      philox4x32_10(ctr, key, o);
      for (int k = 0; k < 3; ++k) f[i * 3 + k] += -g * v[i * 3 + k] + pref * centred_u23(o[k]);
    }
  }
  if (h->force_cap > 0.0) {
    for (int i = 0; i < h->P; ++i) {
      double n2 = f[i * 3] * f[i * 3] + f[i * 3 + 1] * f[i * 3 + 1] + f[i * 3 + 2] * f[i * 3 + 2];
      if (n2 > h->force_cap * h->force_cap) {
        double s = h->force_cap / sqrt(n2);
        for (int k = 0; k < 3; ++k) f[i * 3 + k] *= s;
      }
    }
  }
  release_fixed(h, r, f);
}

static void md_run_replica(orc_handle *h, int r, int n_steps) {
  const double dt = h->cfg.time_step;
  double *x = h->x + (size_t)r * h->P * 3, *v = h->v + (size_t)r * h->P * 3, *f = h->f + (size_t)r * h->P * 3;
  if (n_steps <= 0) return;
  if (!h->f_valid[r]) { total_force(h, r, 1); h->f_valid[r] = 1; }
  for (int s = 0; s < n_steps; ++s) {
    for (int k = 0; k < 3 * h->P; ++k) { v[k] += 0.5 * dt * f[k]; x[k] += dt * v[k]; }
    h->md_counter[r] += 1;
    total_force(h, r, 1);
    for (int k = 0; k < 3 * h->P; ++k) v[k] += 0.5 * dt * f[k];
    h->time[r] += dt;
    h->md_steps[r] += 1;
  }
}

typedef struct md_job { orc_handle *h; int n; } md_job;
static void md_job_fn(void *c, int r) { md_job *j = (md_job *)c; md_run_replica(j->h, r, j->n); }

int orc_md_run(orc_handle *h, int32_t n_steps) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  if (n_steps < 0) return fail(h, CGPE_ERR_INVALID, "n_steps must be >= 0");
  md_job j = {h, n_steps};
  for_each_replica(h->R, md_job_fn, &j);
  return CGPE_OK;
}

/* R8: steepest descent (M.py:78-79): x += clamp(gamma f, +-max_displacement) per coordinate. */
int orc_steepest_descent(orc_handle *h, int32_t n_steps, double f_max, double gamma, double max_disp) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  for (int r = 0; r < h->R; ++r) {
    double *x = h->x + (size_t)r * h->P * 3, *f = h->f + (size_t)r * h->P * 3;
    total_force(h, r, 0);
    for (int s = 0; s < n_steps; ++s) {
      double fmax2 = 0.0;
      for (int i = 0; i < h->P; ++i) {
        double n2 = 0.0;
        for (int k = 0; k < 3; ++k) {
          double dp = gamma * f[i * 3 + k];
          if (dp > max_disp) dp = max_disp;
          if (dp < -max_disp) dp = -max_disp;
          x[i * 3 + k] += dp;
          n2 += f[i * 3 + k] * f[i * 3 + k];
        }
        if (n2 > fmax2) fmax2 = n2;
      }
      if (sqrt(fmax2) < f_max) break; /* converged: positions moved, no further force update */
      total_force(h, r, 0);
    }
    h->f_valid[r] = 0;
  }
  return CGPE_OK;
}

/* ------------------------------------------------------------------------------------------- */
/* R1: constant-pH trial moves (reaction.reaction, E.py:94-95; acceptance per SURVEY 8a-R1)     */
/* ------------------------------------------------------------------------------------------- */
static double site_delta_e(const orc_handle *h, int r, int i, int t_old, int t_new, double q_old, double q_new) {
  /* E_new - E_old for swapping type and charge of particle i: only i's pair terms change */
  const cgpe_config *c = &h->cfg;
  const uint8_t *ty = h->type + (size_t)r * h->P;
  const double *q = h->q + (size_t)r * h->P;
  const double *xi = X(h, r, i);
  double de_w = 0.0, de_c = 0.0, fr;
  if (h->use_lists) {
    orc_handle *hm = (orc_handle *)h;
    ensure_lists(hm, r);
    const int *wl = h->wl[r] + (size_t)i * h->cap_wl[r], *dl = h->dl[r] + (size_t)i * h->cap_dl[r];
    const int nw = h->wn[r][i], nd = h->dn[r][i];
    if (c->use_wca)
      for (int k = 0; k < nw; ++k) {
        const int j = wl[k];
        if (ty[j] >= h->T) continue;
        const double *xj = X(h, r, j);
        double d0 = min_image(xi[0] - xj[0], c->box_l), d1 = min_image(xi[1] - xj[1], c->box_l), d2 = min_image(xi[2] - xj[2], c->box_l);
        double r2 = d0 * d0 + d1 * d1 + d2 * d2;
        if (t_new < h->T) de_w += wca(c->wca_eps[t_new * CGPE_MAX_TYPES + ty[j]], c->wca_sig[t_new * CGPE_MAX_TYPES + ty[j]], r2, &fr);
        if (t_old < h->T) de_w -= wca(c->wca_eps[t_old * CGPE_MAX_TYPES + ty[j]], c->wca_sig[t_old * CGPE_MAX_TYPES + ty[j]], r2, &fr);
      }
    if (c->use_dh)
      for (int k = 0; k < nd; ++k) {
        const int j = dl[k];
        if (ty[j] >= h->T || q[j] == 0.0) continue;
        const double *xj = X(h, r, j);
        double d0 = min_image(xi[0] - xj[0], c->box_l), d1 = min_image(xi[1] - xj[1], c->box_l), d2 = min_image(xi[2] - xj[2], c->box_l);
        de_c += dh(c->dh_prefactor * (q_new - q_old) * q[j], h->kappa[r], h->rcut[r], d0 * d0 + d1 * d1 + d2 * d2, &fr);
      }
    return de_w + de_c;
  }
  for (int j = 0; j < h->P; ++j) {
    if (j == i || ty[j] >= h->T) continue;
    const double *xj = X(h, r, j);
    double d0 = min_image(xi[0] - xj[0], c->box_l), d1 = min_image(xi[1] - xj[1], c->box_l),
           d2 = min_image(xi[2] - xj[2], c->box_l);
    double r2 = d0 * d0 + d1 * d1 + d2 * d2;
    if (c->use_wca) {
      if (t_new < h->T) de_w += wca(c->wca_eps[t_new * CGPE_MAX_TYPES + ty[j]], c->wca_sig[t_new * CGPE_MAX_TYPES + ty[j]], r2, &fr);
      if (t_old < h->T) de_w -= wca(c->wca_eps[t_old * CGPE_MAX_TYPES + ty[j]], c->wca_sig[t_old * CGPE_MAX_TYPES + ty[j]], r2, &fr);
    }
    if (c->use_dh && q[j] != 0.0)
      de_c += dh(c->dh_prefactor * (q_new - q_old) * q[j], h->kappa[r], h->rcut[r], r2, &fr);
  }
  return de_w + de_c;
}

/* Energy of an ion of type tb / charge qb at position xb with every interacting particle except
 * `skip`, where particle `ov` is seen with type ov_t and charge ov_q (the site being swapped in the
 * same trial).  *dmin2 returns the squared distance to the nearest interacting particle
 * (exclusion-range test, M.py:202-206,213). */
static double ion_energy(const orc_handle *h, int r, const double *xb, int tb, double qb, int skip, int ov, int ov_t,
                         double ov_q, double *dmin2) {
  const cgpe_config *c = &h->cfg;
  const uint8_t *ty = h->type + (size_t)r * h->P;
  const double *q = h->q + (size_t)r * h->P;
  double e = 0.0, fr, best = INFINITY;
  for (int j = 0; j < h->P; ++j) {
    if (j == skip) continue;
    const int tj = j == ov ? ov_t : ty[j];
    const double qj = j == ov ? ov_q : q[j];
    if (tj >= h->T) continue;
    const double *xj = X(h, r, j);
    double d0 = min_image(xb[0] - xj[0], c->box_l), d1 = min_image(xb[1] - xj[1], c->box_l), d2 = min_image(xb[2] - xj[2], c->box_l);
    double r2 = d0 * d0 + d1 * d1 + d2 * d2;
    if (r2 < best) best = r2;
    if (c->use_wca) e += wca(c->wca_eps[tb * CGPE_MAX_TYPES + tj], c->wca_sig[tb * CGPE_MAX_TYPES + tj], r2, &fr);
    if (c->use_dh && qj != 0.0) e += dh(c->dh_prefactor * qb * qj, h->kappa[r], h->rcut[r], r2, &fr);
  }
  *dmin2 = best;
  return e;
}

static double unit_open(uint32_t x) { return ((double)(x >> 8) + 0.5) * (1.0 / 16777216.0); } /* (0,1) */

static void mc_run_replica_follow(orc_handle *h, int r, int m, int n_trials, cgpe_trial_record *trace,
                                  const cgpe_trial_record *follow, double tie_tol, int32_t *n_ties) {
  const cgpe_reaction *rx = &h->cfg.reactions[m];
  rx_lists *l = &h->lists[(size_t)r * h->cfg.n_reactions + m];
  uint32_t key[2] = {(uint32_t)rx->seed, (uint32_t)(h->cfg.replica_offset + r)};
  int64_t *tried = &h->trials[(size_t)r * h->cfg.n_reactions + m];
  int64_t *acc = &h->accepted[(size_t)r * h->cfg.n_reactions + m];
  uint8_t *ty = h->type + (size_t)r * h->P;
  double *q = h->q + (size_t)r * h->P;
  for (int t = 0; t < n_trials; ++t) {
    uint64_t cnt = (uint64_t)(*tried);
    uint32_t ctr[4] = {(uint32_t)cnt, (uint32_t)(cnt >> 32), 0u, TAG_MC + (uint32_t)m}, o[4];
    philox4x32_10(ctr, key, o);
    *tried += 1;
    const int forward = (o[0] >> 31) == 0;
    int32_t *src = forward ? l->prot : l->deprot, *dst = forward ? l->deprot : l->prot;
    int32_t *n_src = forward ? &l->n_prot : &l->n_deprot, *n_dst = forward ? &l->n_deprot : &l->n_prot;
    cgpe_trial_record rec = {-1, -1, (int8_t)(forward ? 1 : -1), 0, 0, 0, 0.f, 0.f, (float)u24(o[2])};
    const int explicit_ions_ = explicit_ions(&h->cfg);
    const int ni = explicit_ions_ ? h->cfg.n_ion_slots : 1;
    int32_t *ion_active = h->ion_active + (size_t)r * ni, *ion_free = h->ion_free + (size_t)r * ni;
    /* all reactant particles must exist (backward also needs a free H+ to take; forward a spare slot) */
    int possible = *n_src > 0;
    if (explicit_ions_ && !forward && h->n_active[r] == 0) possible = 0;
    if (explicit_ions_ && forward && h->n_free[r] == 0) possible = 0;
    if (possible) {
      const uint32_t k = mulhi32(o[1], (uint32_t)*n_src);
      const int i = src[k];
      const int t_old = forward ? rx->type_prot : rx->type_deprot, t_new = forward ? rx->type_deprot : rx->type_prot;
      const double q_old = q[i], q_new = forward ? rx->q_deprot : rx->q_prot;
      double de = site_delta_e(h, r, i, t_old, t_new, q_old, q_new);
      int b = -1, kb = 0, excluded = 0;
      double xb[3] = {0, 0, 0}, vb[3] = {0, 0, 0};
      if (explicit_ions_) {
        double dmin2;
        if (forward) { /* insert one B at a uniform position with a Maxwell-Boltzmann velocity */
          uint32_t c1[4] = {ctr[0], ctr[1], 1u, ctr[3]}, c2[4] = {ctr[0], ctr[1], 2u, ctr[3]}, o1[4], o2[4];
          philox4x32_10(c1, key, o1);
          philox4x32_10(c2, key, o2);
          b = ion_free[h->n_free[r] - 1];
          for (int a = 0; a < 3; ++a) xb[a] = u24(o1[a]) * h->cfg.box_l;
          const double s = sqrt(rx->kT), ra = sqrt(-2.0 * log(unit_open(o2[0]))), rb = sqrt(-2.0 * log(unit_open(o2[2])));
          vb[0] = s * ra * cos(6.283185307179586 * unit_open(o2[1]));
          vb[1] = s * ra * sin(6.283185307179586 * unit_open(o2[1]));
          vb[2] = s * rb * cos(6.283185307179586 * unit_open(o2[3]));
          de += ion_energy(h, r, xb, rx->type_ion, rx->q_ion, b, i, t_new, q_new, &dmin2);
        } else {       /* take a random free B away */
          kb = (int)mulhi32(o[3], (uint32_t)h->n_active[r]);
          b = ion_active[kb];
          de -= ion_energy(h, r, X(h, r, b), rx->type_ion, rx->q_ion, b, i, t_new, q_new, &dmin2);
        }
        excluded = dmin2 < rx->exclusion_range * rx->exclusion_range;
        rec.ion_slot = b; rec.excluded = (int8_t)excluded;
      }
      const double nu = forward ? 1.0 : -1.0;
      const double bf = excluded ? 0.0 : (double)(*n_src) / (double)(*n_dst + 1) *
                        exp(-de / rx->kT + nu * ORC_LN10 * (h->pH[r] - rx->pK));
      const double u = u24(o[2]);
      rec.particle = i; rec.delta_e = (float)de; rec.bf = (float)bf;
      int accept = u < bf;
      if (follow && follow[t].particle == i && (follow[t].accepted != 0) != accept &&
          fabs(u - bf) <= tie_tol * bf) {
        accept = follow[t].accepted != 0; /* near-tie: stay on the device's path, and count it */
        if (n_ties) *n_ties += 1;
      }
      if (accept) {
        ty[i] = (uint8_t)t_new; q[i] = q_new;
        src[k] = src[*n_src - 1]; *n_src -= 1; /* swap-remove */
        dst[*n_dst] = i; *n_dst += 1;
        if (explicit_ions_) {
          double *xs = h->x + ((size_t)r * h->P + b) * 3, *vs = h->v + ((size_t)r * h->P + b) * 3;
          if (forward) {
            ty[b] = (uint8_t)rx->type_ion; q[b] = rx->q_ion;
            for (int a = 0; a < 3; ++a) { xs[a] = xb[a]; vs[a] = vb[a]; }
            h->n_free[r] -= 1;
            ion_active[h->n_active[r]++] = b;
          } else {
            ty[b] = (uint8_t)h->T; q[b] = 0.0; /* the non-interacting type (M.py:256-261), then gone */
            vs[0] = vs[1] = vs[2] = 0.0;
            ion_active[kb] = ion_active[h->n_active[r] - 1]; h->n_active[r] -= 1;
            ion_free[h->n_free[r]++] = b;
          }
          h->lists_valid[r] = 0;
        }
        *acc += 1;
        rec.accepted = 1;
      }
    }
    if (trace) trace[t] = rec;
  }
  if (n_trials > 0) h->f_valid[r] = 0;
}

static void mc_run_replica(orc_handle *h, int r, int m, int n_trials, cgpe_trial_record *trace) {
  mc_run_replica_follow(h, r, m, n_trials, trace, NULL, 0.0, NULL);
}

typedef struct mc_job { orc_handle *h; int m, n; cgpe_trial_record *trace; } mc_job;
static void mc_job_fn(void *c, int r) {
  mc_job *j = (mc_job *)c;
  mc_run_replica(j->h, r, j->m, j->n, j->trace ? j->trace + (size_t)r * j->n : NULL);
}

int orc_mc_run(orc_handle *h, int32_t reaction_id, int32_t n_trials, cgpe_trial_record *trace) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  if (reaction_id < 0 || reaction_id >= h->cfg.n_reactions) return fail(h, CGPE_ERR_INVALID, "reaction_id out of range");
  if (n_trials < 0) return fail(h, CGPE_ERR_INVALID, "n_trials must be >= 0");
  mc_job j = {h, reaction_id, n_trials, trace};
  for_each_replica(h->R, mc_job_fn, &j);
  return CGPE_OK;
}

/* Parity helper: run the trials like orc_mc_run, but where this FP64 decision and the device's
 * FP32 decision (follow, [R][n_trials]) differ on a near-tie |u - bf| <= tie_tol*bf, take the
 * device's decision so both Markov chains stay on one path.  n_ties counts those trials; any
 * other difference is left in place for the test to find. */
int orc_mc_run_follow(orc_handle *h, int32_t reaction_id, int32_t n_trials, cgpe_trial_record *trace,
                      const cgpe_trial_record *follow, double tie_tol, int32_t *n_ties) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  if (reaction_id < 0 || reaction_id >= h->cfg.n_reactions) return fail(h, CGPE_ERR_INVALID, "reaction_id out of range");
  if (n_ties) *n_ties = 0;
  for (int r = 0; r < h->R; ++r)
    mc_run_replica_follow(h, r, reaction_id, n_trials, trace ? trace + (size_t)r * n_trials : NULL,
                          follow ? follow + (size_t)r * n_trials : NULL, tie_tol, n_ties);
  return CGPE_OK;
}

/* ------------------------------------------------------------------------------------------- */
/* R5/R6 observables (E.py:97-101)                                                              */
/* ------------------------------------------------------------------------------------------- */
static void observables_replica(const orc_handle *h, int r, double *rg, double *ree, int32_t counts[3]) {
  const int N = h->cfg.n_beads;
  double cm[3] = {0, 0, 0};
  for (int i = 0; i < N; ++i) for (int k = 0; k < 3; ++k) cm[k] += X(h, r, i)[k];
  for (int k = 0; k < 3; ++k) cm[k] /= N;
  double s = 0.0;
  for (int i = 0; i < N; ++i) for (int k = 0; k < 3; ++k) { double d = X(h, r, i)[k] - cm[k]; s += d * d; }
  *rg = sqrt(s / N);
  const double *a = X(h, r, 0), *b = X(h, r, N - 1);
  *ree = sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
  counts[0] = counts[1] = counts[2] = 0;
  for (int m = 0; m < h->cfg.n_reactions && m < 2; ++m)
    counts[m] = h->lists[(size_t)r * h->cfg.n_reactions + m].n_deprot;
  if (explicit_ions(&h->cfg)) counts[2] = h->n_active[r]; /* number_of_particles(type=B), E.py:99 */
  else /* implicit ions: every deprotonated site has released one (virtual) H+ : N_B = N_N + N_N2 */
    for (int m = 0; m < h->cfg.n_reactions; ++m)
      counts[2] += h->lists[(size_t)r * h->cfg.n_reactions + m].n_deprot;
}

int orc_observables(orc_handle *h, double *rg, double *ree, int32_t *counts) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  for (int r = 0; r < h->R; ++r) {
    double g, e; int32_t c[3];
    observables_replica(h, r, &g, &e, c);
    if (rg) rg[r] = g;
    if (ree) ree[r] = e;
    if (counts) memcpy(counts + 3 * r, c, sizeof(c));
  }
  return CGPE_OK;
}

int orc_energy(orc_handle *h, double *out) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  for (int r = 0; r < h->R; ++r) {
    double e[3], kin = 0.0;
    interactions(h, r, NULL, e);
    const double *v = h->v + (size_t)r * h->P * 3;
    for (int i = 0; i < h->P; ++i)
      if (h->type[(size_t)r * h->P + i] < h->T) kin += 0.5 * (v[i * 3] * v[i * 3] + v[i * 3 + 1] * v[i * 3 + 1] + v[i * 3 + 2] * v[i * 3 + 2]);
    out[r * 5 + 0] = kin + e[0] + e[1] + e[2];
    out[r * 5 + 1] = kin; out[r * 5 + 2] = e[0]; out[r * 5 + 3] = e[1]; out[r * 5 + 4] = e[2];
  }
  return CGPE_OK;
}

int orc_forces_f64(orc_handle *h, double *out /*[R][P][3]*/) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  double e[3];
  for (int r = 0; r < h->R; ++r) { interactions(h, r, out + (size_t)r * h->P * 3, e); release_fixed(h, r, out + (size_t)r * h->P * 3); }
  return CGPE_OK;
}

int orc_forces(orc_handle *h, float *out /*[R][P][4]*/) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  double *tmp = malloc(sizeof(double) * 3 * h->P), e[3];
  for (int r = 0; r < h->R; ++r) {
    interactions(h, r, tmp, e);
    release_fixed(h, r, tmp);
    for (int i = 0; i < h->P; ++i) {
      float *o = out + ((size_t)r * h->P + i) * 4;
      o[0] = (float)tmp[i * 3]; o[1] = (float)tmp[i * 3 + 1]; o[2] = (float)tmp[i * 3 + 2]; o[3] = 0.f;
    }
  }
  free(tmp);
  return CGPE_OK;
}
int orc_md_forces(orc_handle *h, float *out) { return orc_forces(h, out); }

int orc_min_dist(orc_handle *h, double *out) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  for (int r = 0; r < h->R; ++r) {
    double best = INFINITY;
    for (int i = 0; i < h->P; ++i) for (int j = i + 1; j < h->P; ++j) {
      if (h->type[(size_t)r * h->P + i] >= h->T || h->type[(size_t)r * h->P + j] >= h->T) continue; /* hidden slots */
      const double *a = X(h, r, i), *b = X(h, r, j);
      double d0 = min_image(a[0] - b[0], h->cfg.box_l), d1 = min_image(a[1] - b[1], h->cfg.box_l),
             d2 = min_image(a[2] - b[2], h->cfg.box_l);
      double r2 = d0 * d0 + d1 * d1 + d2 * d2;
      if (r2 < best) best = r2;
    }
    out[r] = sqrt(best);
  }
  return CGPE_OK;
}

int orc_get_counters(orc_handle *h, int64_t *md_steps, int64_t *trials, int64_t *accepted, int64_t *rebuilds) {
  if (md_steps) memcpy(md_steps, h->md_steps, sizeof(int64_t) * h->R);
  if (trials) memcpy(trials, h->trials, sizeof(int64_t) * h->R * h->cfg.n_reactions);
  if (accepted) memcpy(accepted, h->accepted, sizeof(int64_t) * h->R * h->cfg.n_reactions);
  if (rebuilds) memcpy(rebuilds, h->rebuilds, sizeof(int64_t) * h->R);
  return CGPE_OK;
}

int orc_list_stats(orc_handle *h, double *out) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  const int was = h->use_lists;
  if (!was) orc_set_list_mode(h, 1, 0.0);
  const cgpe_config *c = &h->cfg;
  for (int r = 0; r < h->R; ++r) {
    build_lists(h, r);
    double a[4] = {0, 0, 0, 0};
    const uint8_t *ty = h->type + (size_t)r * h->P;
    const double *q = h->q + (size_t)r * h->P;
    for (int i = 0; i < h->P; ++i) {
      const double *xi = X(h, r, i);
      for (int k = 0; k < h->wn[r][i]; ++k) {
        const int j = h->wl[r][(size_t)i * h->cap_wl[r] + k];
        const double *xj = X(h, r, j);
        double d0 = min_image(xi[0] - xj[0], c->box_l), d1 = min_image(xi[1] - xj[1], c->box_l), d2 = min_image(xi[2] - xj[2], c->box_l);
        double s = ty[i] < h->T && ty[j] < h->T ? c->wca_sig[ty[i] * CGPE_MAX_TYPES + ty[j]] : 0.0;
        a[0] += 1;
        if (d0 * d0 + d1 * d1 + d2 * d2 < 1.2599210498948731648 * s * s) a[1] += 1;
      }
      for (int k = 0; k < h->dn[r][i]; ++k) {
        const int j = h->dl[r][(size_t)i * h->cap_dl[r] + k];
        const double *xj = X(h, r, j);
        double d0 = min_image(xi[0] - xj[0], c->box_l), d1 = min_image(xi[1] - xj[1], c->box_l), d2 = min_image(xi[2] - xj[2], c->box_l);
        a[2] += 1;
        if (q[i] != 0.0 && q[j] != 0.0 && d0 * d0 + d1 * d1 + d2 * d2 < h->rcut[r] * h->rcut[r]) a[3] += 1;
      }
    }
    memcpy(out + 4 * r, a, sizeof(a));
  }
  if (!was) { h->use_lists = 0; }
  return CGPE_OK;
}

/* the oracle has one engine; the entry points exist so that the same binding drives both libraries */
int orc_set_engine(orc_handle *h, int32_t engine) { (void)h; (void)engine; return CGPE_OK; }
int orc_engine_info(orc_handle *h, int32_t *engine, int64_t *compact_builds, int64_t *cta_ns) {
  if (cta_ns) for (int r = 0; r < 4 * h->cfg.n_replicas; ++r) cta_ns[r] = 0;
  if (engine) *engine = -1;
  if (compact_builds) for (int r = 0; r < h->cfg.n_replicas; ++r) compact_builds[r] = 0;
  return CGPE_OK;
}
/* the frame recorder belongs to the CUDA engines; the oracle's role is taken by the host-side accumulators of the facade */
int orc_frames_begin(orc_handle *h, int32_t every, int32_t flags, int32_t capacity) { (void)every; (void)flags; (void)capacity; snprintf(h->err, 512, "the oracle records no frames"); return CGPE_ERR_UNSUPPORTED; }
int orc_frames_fetch(orc_handle *h, double *frames, float *bonds, float *dihedrals, int32_t *n_frames) { (void)frames; (void)bonds; (void)dihedrals; (void)n_frames; snprintf(h->err, 512, "the oracle records no frames"); return CGPE_ERR_UNSUPPORTED; }
int orc_frames_end(orc_handle *h) { (void)h; return CGPE_OK; }
int orc_last_call_ms(orc_handle *h, float *ms, int32_t *n) { (void)h; if (ms) *ms = 0.f; if (n) *n = 0; return CGPE_OK; }

/* ------------------------------------------------------------------------------------------- */
/* R7: the sampling loop body (E.py:88-106)                                                     */
/* ------------------------------------------------------------------------------------------- */
static uint32_t bernoulli_threshold(double p) {
  /* u24 < p  <=>  (x >> 8) < ceil(p 2^24), evaluated in integers on both sides */
  if (p <= 0.0) return 0u;
  if (p >= 1.0) return 16777216u;
  return (uint32_t)ceil(p * 16777216.0);
}

typedef struct sample_job { orc_handle *h; const cgpe_sample_params *sp; } sample_job;
static void sample_job_fn(void *c, int r) {
  orc_handle *h = ((sample_job *)c)->h;
  const cgpe_sample_params *sp = ((sample_job *)c)->sp;
  const int ns = sp->n_samples, N = h->cfg.n_beads;
  const uint32_t thr1 = bernoulli_threshold(sp->p_burst1), thr2 = bernoulli_threshold(sp->p_burst2);
  {
    uint32_t key[2] = {(uint32_t)sp->schedule_seed, (uint32_t)(h->cfg.replica_offset + r)};
    for (int s = 0; s < ns; ++s) {
      uint64_t cnt = (uint64_t)h->samples_done[r];
      uint32_t ctr[4] = {(uint32_t)cnt, (uint32_t)(cnt >> 32), 0u, TAG_SCHEDULE}, o[4];
      philox4x32_10(ctr, key, o);
      h->samples_done[r] += 1;
      if (sp->use_md) {
        if ((o[0] >> 8) < thr1) md_run_replica(h, r, sp->md_steps_per_burst); /* E.py:89-90 */
        if ((o[1] >> 8) < thr2) md_run_replica(h, r, sp->md_steps_per_burst); /* E.py:91-92 */
      }
      for (int b = 0; b < sp->n_mc_blocks; ++b) /* E.py:94-95 */
        mc_run_replica(h, r, sp->mc_reaction[b], sp->mc_trials[b], NULL);
      double rg, ree; int32_t c[3];
      observables_replica(h, r, &rg, &ree, c); /* E.py:97-102 */
      double *o6 = h->last_obs + ((size_t)r * ns + s) * CGPE_N_OBS;
      o6[0] = c[0]; o6[1] = c[1]; o6[2] = c[2]; o6[3] = rg; o6[4] = ree; o6[5] = h->time[r];
      if (sp->q_snapshot_every > 0 && s % sp->q_snapshot_every == 0) { /* E.py:103-106 */
        int row = s / sp->q_snapshot_every;
        if (row < h->last_qrows)
          for (int i = 0; i < N; ++i)
            h->last_qdist[((size_t)r * h->last_qrows + row) * N + i] = (float)h->q[(size_t)r * h->P + i];
      }
    }
  }
}

int orc_sample_loop_async(orc_handle *h, const cgpe_sample_params *sp) {
  if (!h->have_state) return fail(h, CGPE_ERR_STATE, "set_state has not been called");
  if (!sp || sp->n_samples < 0 || sp->n_mc_blocks < 0 || sp->n_mc_blocks > CGPE_MAX_MC_BLOCKS)
    return fail(h, CGPE_ERR_INVALID, "bad sample params");
  for (int b = 0; b < sp->n_mc_blocks; ++b)
    if (sp->mc_reaction[b] < 0 || sp->mc_reaction[b] >= h->cfg.n_reactions || sp->mc_trials[b] < 0)
      return fail(h, CGPE_ERR_INVALID, "bad MC block");
  const int ns = sp->n_samples, N = h->cfg.n_beads;
  free(h->last_obs); free(h->last_qdist);
  h->last_obs = calloc((size_t)h->R * (ns > 0 ? ns : 1) * CGPE_N_OBS, sizeof(double));
  h->last_qrows = sp->q_snapshot_every > 0 ? sp->q_snapshot_rows : 0;
  h->last_qdist = calloc((size_t)h->R * (h->last_qrows > 0 ? h->last_qrows : 1) * N, sizeof(float));
  h->last_ns = ns;
  sample_job j = {h, sp};
  for_each_replica(h->R, sample_job_fn, &j);
  return CGPE_OK;
}

int orc_fetch_samples(orc_handle *h, double *obs, float *qdist) {
  if (!h->last_obs) return fail(h, CGPE_ERR_STATE, "no sample loop has run");
  if (obs) memcpy(obs, h->last_obs, sizeof(double) * (size_t)h->R * h->last_ns * CGPE_N_OBS);
  if (qdist && h->last_qrows > 0)
    memcpy(qdist, h->last_qdist, sizeof(float) * (size_t)h->R * h->last_qrows * h->cfg.n_beads);
  return CGPE_OK;
}

int orc_sample_loop(orc_handle *h, const cgpe_sample_params *sp, double *obs, float *qdist) {
  int rc = orc_sample_loop_async(h, sp);
  if (rc) return rc;
  return orc_fetch_samples(h, obs, qdist);
}
