/*
 * moc_oracle.c - CPU restatement (plain C) of the reference's MOC source
 * iteration.  TEST INFRASTRUCTURE ONLY - see moc_oracle.h for the rules and
 * the pinning status (PINNED against the reference's golden files).
 *
 * Conventions follow the reference CPUSolver (double FP_PRECISION build):
 *   psi   float, index (t*2+dir)*F + p*G + e           src/Solver.h:49-54
 *   phi,q double, index r*G + e                        src/Solver.h:34-46
 *   sigma_s[dest*G+orig], fiss_matrix[G_dest*G+g]      src/Material.cpp:728,977
 */
#include "moc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define VEC_LENGTH 4                      /* profile/Makefile:196-201 (double) */
#define FOUR_PI 12.566370614359172        /* src/constants.h:27 */
#define ONE_OVER_FOUR_PI 0.07957747154594767 /* src/constants.h:30 */
#define FLUX_EPSILON 1.0E-25              /* src/constants.h:15 */
#define VOL_EPSILON 1.0E-12               /* the reference's FLT_EPSILON, constants.h:12 */

enum { BC_VACUUM = 0, BC_REFLECTIVE = 1, BC_PERIODIC = 2, BC_INTERFACE = 3 };
enum { STAB_DIAGONAL = 0, STAB_YAMAMOTO = 1, STAB_GLOBAL = 2 };

struct moc_oracle {
  int G, A, P, solve_3d, NP, F, n_mat;
  int64_t n_trk, n_seg, n_fsr, n_fissionable;
  double* seg_len; int32_t* seg_fsr;
  int64_t* trk_off; int32_t* trk_azim; int32_t* trk_polar;
  int64_t* next_fwd; int64_t* next_bwd;
  uint8_t* flags; uint8_t* bc_fwd; uint8_t* bc_bwd;
  double* weight; double* sin_theta;
  double* vol; int32_t* fsr_mat;
  double* sigma_t; double* sigma_s; double* fiss; double* nu_sigma_f; double* sigma_f; double* chi;
  uint8_t* fissionable;
  /* state */
  double* phi; double* phi_old; double* q; double* fixed; double* stab;
  float* psi_start; float* psi_bound;
  double k_eff;
  int fixed_on, stabilize, stab_type, threads, balance;
  float* leakage; double* sigma_a;
  double stab_factor;
  double sweep_seconds;
  double* scratch;
};

static void* dup_mem(const void* src, size_t bytes) {
  void* p = malloc(bytes ? bytes : 1);
  if (src && bytes) memcpy(p, src, bytes);
  return p;
}

/* src/pairwise_sum.h:17-39 */
static double pairwise_sum(const double* v, int64_t length) {
  double sum = 0;
  if (length < VEC_LENGTH) {
    for (int64_t i = 0; i < length; i++) sum += v[i];
  } else {
    int64_t offset = length % 2;
    length = length / 2;
    sum = pairwise_sum(v, length) + pairwise_sum(v + length, length + offset);
  }
  return sum;
}

/* src/exponentials.h:156-192: (1-exp(-x))/x, 5/6-order rational */
static inline double expF1_fractional(double x) {
  const double p0 = 1.0;
  const double p1 = 2.4172687328033081 * 1E-1;
  const double p2 = 6.2804790965268531 * 1E-2;
  const double p3 = 1.0567595009016521 * 1E-2;
  const double p4 = 1.0059468082903561 * 1E-3;
  const double p5 = 1.9309063097411041 * 1E-4;
  const double d0 = 1.0;
  const double d1 = 7.4169266112320541 * 1E-1;
  const double d2 = 2.6722515319494311 * 1E-1;
  const double d3 = 6.1643725066901411 * 1E-2;
  const double d4 = 1.0590759992367811 * 1E-2;
  const double d5 = 1.0057980007137651 * 1E-3;
  const double d6 = 1.9309063097411041 * 1E-4;
  double num, den;
  den = d6 * x + d5;
  den = den * x + d4;
  den = den * x + d3;
  den = den * x + d2;
  den = den * x + d1;
  den = den * x + d0;
  den = 1. / den;
  num = p5 * x + p4;
  num = num * x + p3;
  num = num * x + p2;
  num = num * x + p1;
  num = num * x + p0;
  return num * den;
}

double moc_oracle_expF1(double x) { return expF1_fractional(x); }

moc_oracle* moc_oracle_create(
    int num_groups, int num_azim, int num_polar, int solve_3d,
    int64_t n_tracks, int64_t n_segments, int64_t n_fsrs, int n_materials,
    const double* seg_length, const int32_t* seg_fsr,
    const int64_t* trk_seg_offset, const int32_t* trk_azim, const int32_t* trk_polar,
    const int64_t* trk_next_fwd, const int64_t* trk_next_bwd,
    const uint8_t* trk_flags, const uint8_t* trk_bc_fwd, const uint8_t* trk_bc_bwd,
    const double* quad_weight, const double* quad_sin_theta,
    const double* fsr_volume, const int32_t* fsr_mat,
    const double* mat_sigma_t, const double* mat_sigma_s, const double* mat_fiss_matrix,
    const double* mat_nu_sigma_f, const double* mat_sigma_f, const double* mat_chi,
    const uint8_t* mat_fissionable) {
  moc_oracle* o = (moc_oracle*)calloc(1, sizeof(moc_oracle));
  int G = num_groups;
  o->G = G; o->A = num_azim; o->P = num_polar; o->solve_3d = solve_3d;
  o->NP = solve_3d ? 1 : num_polar / 2;
  o->F = G * o->NP;                                   /* src/Solver.cpp:432-449 */
  o->n_trk = n_tracks; o->n_seg = n_segments; o->n_fsr = n_fsrs; o->n_mat = n_materials;
  o->seg_len = dup_mem(seg_length, n_segments * 8);
  o->seg_fsr = dup_mem(seg_fsr, n_segments * 4);
  o->trk_off = dup_mem(trk_seg_offset, (n_tracks + 1) * 8);
  o->trk_azim = dup_mem(trk_azim, n_tracks * 4);
  o->trk_polar = dup_mem(trk_polar, n_tracks * 4);
  o->next_fwd = dup_mem(trk_next_fwd, n_tracks * 8);
  o->next_bwd = dup_mem(trk_next_bwd, n_tracks * 8);
  o->flags = dup_mem(trk_flags, n_tracks);
  o->bc_fwd = dup_mem(trk_bc_fwd, n_tracks);
  o->bc_bwd = dup_mem(trk_bc_bwd, n_tracks);
  o->weight = dup_mem(quad_weight, (size_t)(num_azim / 2) * num_polar * 8);
  o->sin_theta = dup_mem(quad_sin_theta, (size_t)(num_azim / 2) * num_polar * 8);
  o->vol = dup_mem(fsr_volume, n_fsrs * 8);
  o->fsr_mat = dup_mem(fsr_mat, n_fsrs * 4);
  o->sigma_t = dup_mem(mat_sigma_t, (size_t)n_materials * G * 8);
  o->sigma_s = dup_mem(mat_sigma_s, (size_t)n_materials * G * G * 8);
  o->fiss = dup_mem(mat_fiss_matrix, (size_t)n_materials * G * G * 8);
  o->nu_sigma_f = dup_mem(mat_nu_sigma_f, (size_t)n_materials * G * 8);
  o->sigma_f = dup_mem(mat_sigma_f, (size_t)n_materials * G * 8);
  o->chi = dup_mem(mat_chi, (size_t)n_materials * G * 8);
  o->fissionable = dup_mem(mat_fissionable, n_materials);
  size_t nphi = (size_t)n_fsrs * G;
  o->phi = calloc(nphi, 8); o->phi_old = calloc(nphi, 8); o->q = calloc(nphi, 8);
  o->fixed = calloc(nphi, 8); o->stab = calloc(nphi, 8);
  size_t npsi = (size_t)n_tracks * 2 * o->F;
  o->psi_start = calloc(npsi, 4); o->psi_bound = calloc(npsi, 4);
  o->scratch = calloc(n_fsrs > 0 ? n_fsrs : 1, 8);
  o->leakage = calloc(n_tracks > 0 ? n_tracks : 1, 4);
  /* Material::getSigmaA when never set: sigma_t minus the out-scatter sum (src/Material.cpp:241-249) */
  o->sigma_a = calloc((size_t)n_materials * G, 8);
  for (int m = 0; m < n_materials; m++)
    for (int g = 0; g < G; g++) {
      double a = o->sigma_t[(size_t)m * G + g];
      for (int gp = 0; gp < G; gp++) a -= o->sigma_s[((size_t)m * G + gp) * G + g];
      o->sigma_a[(size_t)m * G + g] = a;
    }
  o->k_eff = 1.0;
  o->threads = 1;
  /* src/Solver.cpp:882-892 */
  for (int64_t r = 0; r < n_fsrs; r++)
    if (o->fissionable[o->fsr_mat[r]]) o->n_fissionable++;
  return o;
}

void moc_oracle_destroy(moc_oracle* o) {
  if (!o) return;
  free(o->seg_len); free(o->seg_fsr); free(o->trk_off); free(o->trk_azim); free(o->trk_polar);
  free(o->next_fwd); free(o->next_bwd); free(o->flags); free(o->bc_fwd); free(o->bc_bwd);
  free(o->weight); free(o->sin_theta); free(o->vol); free(o->fsr_mat);
  free(o->sigma_t); free(o->sigma_s); free(o->fiss); free(o->nu_sigma_f); free(o->sigma_f);
  free(o->chi); free(o->fissionable);
  free(o->phi); free(o->phi_old); free(o->q); free(o->fixed); free(o->stab);
  free(o->psi_start); free(o->psi_bound); free(o->scratch); free(o->leakage); free(o->sigma_a);
  free(o);
}

void moc_oracle_set_num_threads(moc_oracle* o, int n) { o->threads = n > 0 ? n : 1; }

/* src/CPUSolver.cpp:463-477 */
void moc_oracle_zero_track_fluxes(moc_oracle* o) {
  size_t n = (size_t)o->n_trk * 2 * o->F;
  memset(o->psi_start, 0, n * 4);
  memset(o->psi_bound, 0, n * 4);
}

/* src/CPUSolver.cpp:1816-1824 */
void moc_oracle_flatten_fsr_fluxes(moc_oracle* o, double value) {
  for (int64_t i = 0; i < o->n_fsr * o->G; i++) o->phi[i] = value;
}

/* src/CPUSolver.cpp:1846-1853 */
void moc_oracle_store_fsr_fluxes(moc_oracle* o) {
  memcpy(o->phi_old, o->phi, (size_t)o->n_fsr * o->G * 8);
}

/* src/CPUSolver.cpp:1860-1931 */
double moc_oracle_normalize_fluxes(moc_oracle* o) {
  int G = o->G;
  double* gs = (double*)malloc(G * 8);
  for (int64_t r = 0; r < o->n_fsr; r++) {
    const double* nsf = o->nu_sigma_f + (size_t)o->fsr_mat[r] * G;
    double volume = o->vol[r];
    for (int e = 0; e < G; e++) gs[e] = nsf[e] * o->phi[r * G + e] * volume;
    o->scratch[r] = pairwise_sum(gs, G);
  }
  free(gs);
  double tot = pairwise_sum(o->scratch, o->n_fsr);
  double norm_factor = o->n_fsr / tot;
  for (int64_t i = 0; i < o->n_fsr * G; i++) o->phi[i] *= norm_factor;
  size_t n = (size_t)o->n_trk * 2 * o->F;
  for (size_t i = 0; i < n; i++) {
    o->psi_start[i] *= norm_factor;   /* float *= double, rounded to float */
    o->psi_bound[i] *= norm_factor;
  }
  return norm_factor;
}

/* src/CPUSolver.cpp:1939-2023 */
void moc_oracle_compute_fsr_sources(moc_oracle* o, int iteration) {
  int G = o->G;
#pragma omp parallel num_threads(o->threads)
  {
    double* fs = (double*)malloc(G * 8);
    double* ss = (double*)malloc(G * 8);
#pragma omp for schedule(static)
    for (int64_t r = 0; r < o->n_fsr; r++) {
      int m = o->fsr_mat[r];
      const double* sigma_s = o->sigma_s + (size_t)m * G * G;
      const double* fm = o->fiss + (size_t)m * G * G;
      int fissionable = o->fissionable[m];
      for (int Gd = 0; Gd < G; Gd++) {
        int first = Gd * G;
        for (int g = 0; g < G; g++) {
          double fiss_mat = fissionable ? fm[first + g] : 0.;
          ss[g] = sigma_s[first + g] * o->phi[r * G + g];
          fs[g] = o->phi[r * G + g] * fiss_mat;
        }
        double scatter_source = pairwise_sum(ss, G);
        double fission_source = pairwise_sum(fs, G);
        fission_source /= o->k_eff;
        double q = fission_source;
        q += scatter_source;
        if (o->fixed_on) q += o->fixed[r * G + Gd];
        q *= ONE_OVER_FOUR_PI;
        if (q < 0.0 && iteration < 30) q = FLUX_EPSILON;
        o->q[r * G + Gd] = q;
      }
    }
    free(fs); free(ss);
  }
}

/* src/CPUSolver.cpp:2030-2066 */
void moc_oracle_compute_fsr_fission_sources(moc_oracle* o) {
  int G = o->G;
  double* fs = (double*)malloc(G * 8);
  for (int64_t r = 0; r < o->n_fsr; r++) {
    int m = o->fsr_mat[r];
    const double* fm = o->fiss + (size_t)m * G * G;
    for (int g = 0; g < G; g++) {
      for (int gp = 0; gp < G; gp++) {
        double fiss_mat = o->fissionable[m] ? fm[g * G + gp] : 0.;
        fs[gp] = fiss_mat * o->phi[r * G + gp];
      }
      o->q[r * G + g] = pairwise_sum(fs, G) * ONE_OVER_FOUR_PI;
    }
  }
  free(fs);
}

/* src/CPUSolver.cpp:2072-2104 */
void moc_oracle_compute_fsr_scatter_sources(moc_oracle* o) {
  int G = o->G;
  double* ss = (double*)malloc(G * 8);
  for (int64_t r = 0; r < o->n_fsr; r++) {
    const double* sigma_s = o->sigma_s + (size_t)o->fsr_mat[r] * G * G;
    for (int g = 0; g < G; g++) {
      for (int gp = 0; gp < G; gp++) ss[gp] = sigma_s[g * G + gp] * o->phi[r * G + gp];
      o->q[r * G + g] = pairwise_sum(ss, G) * ONE_OVER_FOUR_PI;
    }
  }
  free(ss);
}

/* One track, both directions: src/TrackTraversingAlgorithms.cpp:890-1052 with
 * tallyScalarFlux (src/CPUSolver.cpp:2402-2497), accumulateScalarFluxContribution
 * (:2507-2527) and transferBoundaryFlux (:2560-2601) inlined. */
static void sweep_track(moc_oracle* o, int64_t t, double* fsr_flux) {
  const int G = o->G, NP = o->NP, F = o->F, P = o->P;
  const int azim = o->trk_azim[t];
  const int polar = o->trk_polar[t];
  const int64_t s0 = o->trk_off[t], s1 = o->trk_off[t + 1];
  const double* wrow = o->weight + (size_t)azim * P;
  /* 2D: the evaluator of azim a serves A/2-1-a too (src/Solver.cpp:763-779) */
  int a_eval = azim;
  if (a_eval >= o->A / 4) a_eval = o->A / 2 - 1 - azim;
  const double* srow = o->sin_theta + (size_t)a_eval * P;
  const double weight3d = o->solve_3d ? wrow[polar] : 1.0;

  memset(fsr_flux, 0, G * 8);

  for (int dir = 0; dir < 2; dir++) {
    float* track_flux = o->psi_bound + ((size_t)t * 2 + dir) * F;
    int64_t s = dir == 0 ? s0 : s1 - 1;
    const int64_t step = dir == 0 ? 1 : -1;
    for (int64_t n = 0; n < s1 - s0; n++, s += step) {
      const int64_t fsr = o->seg_fsr[s];
      const double length = o->seg_len[s];
      const double* sigma_t = o->sigma_t + (size_t)o->fsr_mat[fsr] * G;
      const double* q = o->q + fsr * G;
      if (o->solve_3d) {
        for (int e = 0; e < G; e++) {
          double tau = sigma_t[e] * length;
          double exponential = expF1_fractional(tau);
          double delta_psi = (tau * track_flux[e] - length * q[e]) * exponential;
          track_flux[e] -= delta_psi;
          fsr_flux[e] += delta_psi;
        }
      } else {
        for (int p = 0; p < NP; p++) {
          double inv_sin = 1.0 / srow[p];
          double wgt = wrow[p];
          for (int e = 0; e < G; e++) {
            int pe = p * G + e;
            double tau = sigma_t[e] * length;
            /* ExpEvaluator::computeExponential, src/ExpEvaluator.h:170-183 */
            double exponential = inv_sin * expF1_fractional(tau * inv_sin);
            double delta_psi = (tau * track_flux[pe] - length * q[e]) * exponential;
            track_flux[pe] -= delta_psi;
            fsr_flux[e] += delta_psi * wgt;
          }
        }
      }
      /* flush before the FSR changes; the last forward segment is carried
       * into the backward pass (TrackTraversingAlgorithms.cpp:982 vs 1030) */
      int flush;
      if (dir == 0) flush = (s < s1 - 1) && (fsr != o->seg_fsr[s + 1]);
      else flush = (s == s0) || (fsr != o->seg_fsr[s - 1]);
      if (flush) {
        for (int e = 0; e < G; e++) {
          double add = weight3d * fsr_flux[e];
#pragma omp atomic update
          o->phi[fsr * G + e] += add;
          fsr_flux[e] = 0.;
        }
      }
    }
    /* transferBoundaryFlux */
    uint8_t bc = dir == 0 ? o->bc_fwd[t] : o->bc_bwd[t];
    int64_t nxt = dir == 0 ? o->next_fwd[t] : o->next_bwd[t];
    int next_is_fwd = dir == 0 ? (o->flags[t] & 1) : ((o->flags[t] >> 1) & 1);
    if (bc == BC_REFLECTIVE || bc == BC_PERIODIC) {
      float* out = o->psi_start + ((size_t)nxt * 2 + (next_is_fwd ? 0 : 1)) * F;
      memcpy(out, track_flux, F * 4);
    }
    /* leakage tally (src/CPUSolver.cpp:2592-2600): weight of (azim, polar_index), where
     * polar_index is 0 for every 2D track (TrackTraversingAlgorithms.cpp:901) */
    if (o->balance && bc == BC_VACUUM) {
      double weight = wrow[o->solve_3d ? polar : 0];
      for (int pe = 0; pe < F; pe++) o->leakage[t] += weight * track_flux[pe];
    }
  }
}

/* src/CPUSolver.cpp:2338-2389 */
void moc_oracle_transport_sweep(moc_oracle* o) {
  double t0 = omp_get_wtime();
  memset(o->phi, 0, (size_t)o->n_fsr * o->G * 8);                    /* :2347 */
  memcpy(o->psi_bound, o->psi_start, (size_t)o->n_trk * 2 * o->F * 4); /* :2351 */
  memset(o->leakage, 0, (size_t)o->n_trk * 4);                          /* :2360 */
#pragma omp parallel num_threads(o->threads)
  {
    double* fsr_flux = (double*)malloc(o->G * 8);
#pragma omp for schedule(dynamic)
    for (int64_t t = 0; t < o->n_trk; t++) sweep_track(o, t, fsr_flux);
    free(fsr_flux);
  }
  o->sweep_seconds += omp_get_wtime() - t0;
}

double moc_oracle_sweep_seconds(moc_oracle* o, int reset) {
  double s = o->sweep_seconds;
  if (reset) o->sweep_seconds = 0.;
  return s;
}

/* src/CPUSolver.cpp:2608-2659 */
void moc_oracle_add_source_to_scalar_flux(moc_oracle* o) {
  int G = o->G;
  for (int64_t r = 0; r < o->n_fsr; r++) {
    double volume = o->vol[r];
    const double* sigma_t = o->sigma_t + (size_t)o->fsr_mat[r] * G;
    if (volume < VOL_EPSILON) volume = 1e30;
    for (int e = 0; e < G; e++) {
      o->phi[r * G + e] /= (sigma_t[e] * volume);
      o->phi[r * G + e] += FOUR_PI * o->q[r * G + e] / sigma_t[e];
      if (o->phi[r * G + e] < 0.0) o->phi[r * G + e] = FLUX_EPSILON;
    }
  }
}

/* src/CPUSolver.cpp:2258-2328 (fission-rate form, the default) */
void moc_oracle_compute_keff(moc_oracle* o) {
  int G = o->G;
  double* gr = (double*)malloc(G * 8);
  for (int64_t r = 0; r < o->n_fsr; r++) {
    const double* sigma = o->nu_sigma_f + (size_t)o->fsr_mat[r] * G;
    for (int e = 0; e < G; e++) gr[e] = sigma[e] * o->phi[r * G + e];
    o->scratch[r] = pairwise_sum(gr, G);
    o->scratch[r] *= o->vol[r];
  }
  free(gr);
  double rate = pairwise_sum(o->scratch, o->n_fsr);
  if (!o->balance) {
    o->k_eff *= rate / o->n_fsr;
    return;
  }
  /* k = fission / (absorption + leakage), src/CPUSolver.cpp:2264-2325 */
  gr = (double*)malloc(G * 8);
  for (int64_t r = 0; r < o->n_fsr; r++) {
    const double* sigma = o->sigma_a + (size_t)o->fsr_mat[r] * G;
    for (int e = 0; e < G; e++) gr[e] = sigma[e] * o->phi[r * G + e];
    o->scratch[r] = pairwise_sum(gr, G);
    o->scratch[r] *= o->vol[r];
  }
  free(gr);
  double absorption = pairwise_sum(o->scratch, o->n_fsr);
  double leak = 0.;
  for (int64_t t = 0; t < o->n_trk; t++) leak += o->leakage[t];
  o->k_eff = rate / (absorption + leak);
}

void moc_oracle_set_keff_from_neutron_balance(moc_oracle* o, int on) { o->balance = on; }

/* src/CPUSolver.cpp:2113-2252 */
double moc_oracle_compute_residual(moc_oracle* o, int res_type) {
  int G = o->G;
  int64_t norm;
  double* residuals = o->scratch;
  memset(residuals, 0, o->n_fsr * 8);
  const double* ref = o->phi_old;
  if (res_type == MOC_RES_SCALAR_FLUX) {
    norm = o->n_fsr;
    for (int64_t r = 0; r < o->n_fsr; r++)
      for (int e = 0; e < G; e++)
        if (ref[r * G + e] > 0.)
          residuals[r] += pow((o->phi[r * G + e] - ref[r * G + e]) / ref[r * G + e], 2);
  } else if (res_type == MOC_RES_FISSION_SOURCE) {
    norm = o->n_fissionable;
    for (int64_t r = 0; r < o->n_fsr; r++) {
      int m = o->fsr_mat[r];
      if (!o->fissionable[m]) continue;
      const double* nsf = o->nu_sigma_f + (size_t)m * G;
      double nw = 0., old = 0.;
      for (int e = 0; e < G; e++) {
        nw += o->phi[r * G + e] * nsf[e];
        old += ref[r * G + e] * nsf[e];
      }
      if (old > 0.) residuals[r] = pow((nw - old) / old, 2);
    }
  } else {
    norm = o->n_fsr;
    double inverse_k_eff = 1.0 / o->k_eff;
    for (int64_t r = 0; r < o->n_fsr; r++) {
      int m = o->fsr_mat[r];
      double nw = 0., old = 0.;
      if (o->fissionable[m]) {
        const double* nsf = o->nu_sigma_f + (size_t)m * G;
        for (int e = 0; e < G; e++) {
          nw += o->phi[r * G + e] * nsf[e];
          old += ref[r * G + e] * nsf[e];
        }
        nw *= inverse_k_eff;
        old *= inverse_k_eff;
      }
      const double* sigma_s = o->sigma_s + (size_t)m * G * G;
      for (int Gd = 0; Gd < G; Gd++)
        for (int g = 0; g < G; g++) {
          nw += sigma_s[Gd * G + g] * o->phi[r * G + g];
          old += sigma_s[Gd * G + g] * ref[r * G + g];
        }
      if (old > 0.) residuals[r] = pow((nw - old) / old, 2);
    }
  }
  double residual = pairwise_sum(residuals, o->n_fsr);
  if (residual < 0.0) residual = 0.0;
  if (norm <= 0) norm = 1;
  return sqrt(residual / norm);
}

/* src/CPUSolver.cpp:2665-2728 */
void moc_oracle_compute_stabilizing_flux(moc_oracle* o) {
  int G = o->G;
  if (o->stab_type == STAB_DIAGONAL) {
    for (int64_t r = 0; r < o->n_fsr; r++) {
      int m = o->fsr_mat[r];
      for (int e = 0; e < G; e++) {
        double sigma_s = o->sigma_s[(size_t)m * G * G + e * G + e];
        if (sigma_s < 0.0)
          o->stab[r * G + e] = -o->phi[r * G + e] * o->stab_factor * sigma_s /
                               o->sigma_t[(size_t)m * G + e];
      }
    }
  } else if (o->stab_type == STAB_YAMAMOTO) {
    for (int e = 0; e < G; e++) {
      double max_ratio = 0.0;
      for (int64_t r = 0; r < o->n_fsr; r++) {
        int m = o->fsr_mat[r];
        double ratio = fabs(o->sigma_s[(size_t)m * G * G + e * G + e] / o->sigma_t[(size_t)m * G + e]);
        if (ratio > max_ratio) max_ratio = ratio;
      }
      max_ratio *= o->stab_factor;
      for (int64_t r = 0; r < o->n_fsr; r++) o->stab[r * G + e] = o->phi[r * G + e] * max_ratio;
    }
  } else {
    double mult = 1.0 / o->stab_factor - 1.0;
    for (int64_t i = 0; i < o->n_fsr * G; i++) o->stab[i] = mult * o->phi[i];
  }
}

/* src/CPUSolver.cpp:2736-2805 */
void moc_oracle_stabilize_flux(moc_oracle* o) {
  int G = o->G;
  if (o->stab_type == STAB_DIAGONAL) {
    for (int64_t r = 0; r < o->n_fsr; r++) {
      int m = o->fsr_mat[r];
      for (int e = 0; e < G; e++) {
        double sigma_s = o->sigma_s[(size_t)m * G * G + e * G + e];
        if (sigma_s < 0.0) {
          o->phi[r * G + e] += o->stab[r * G + e];
          o->phi[r * G + e] /= (1.0 - o->stab_factor * sigma_s / o->sigma_t[(size_t)m * G + e]);
        }
      }
    }
  } else if (o->stab_type == STAB_YAMAMOTO) {
    for (int e = 0; e < G; e++) {
      double max_ratio = 0.0;
      for (int64_t r = 0; r < o->n_fsr; r++) {
        int m = o->fsr_mat[r];
        double ratio = fabs(o->sigma_s[(size_t)m * G * G + e * G + e] / o->sigma_t[(size_t)m * G + e]);
        if (ratio > max_ratio) max_ratio = ratio;
      }
      max_ratio *= o->stab_factor;
      for (int64_t r = 0; r < o->n_fsr; r++) {
        o->phi[r * G + e] += o->stab[r * G + e];
        o->phi[r * G + e] /= (1 + max_ratio);
      }
    }
  } else {
    for (int64_t i = 0; i < o->n_fsr * G; i++) {
      o->phi[i] += o->stab[i];
      o->phi[i] *= o->stab_factor;
    }
  }
}

/* src/Solver.cpp:1542-1689 (no CMFD) + computeInitialFluxGuess :1710-1731 */
int moc_oracle_compute_eigenvalue(moc_oracle* o, int max_iters, double tol, int res_type) {
  int num_iterations = 0;
  double previous_residual = 1.0, residual = 0.;
  o->k_eff = 1.;
  moc_oracle_zero_track_fluxes(o);            /* initializeFluxArrays allocates zeros */
  memset(o->phi_old, 0, (size_t)o->n_fsr * o->G * 8);
  moc_oracle_flatten_fsr_fluxes(o, 1.0);
  moc_oracle_normalize_fluxes(o);
  moc_oracle_store_fsr_fluxes(o);
  double k_prev = o->k_eff;
  for (int i = 0; i < max_iters; i++) {
    if (i > 0 && o->stabilize) moc_oracle_compute_stabilizing_flux(o);
    moc_oracle_compute_fsr_sources(o, i);
    moc_oracle_transport_sweep(o);
    moc_oracle_add_source_to_scalar_flux(o);
    moc_oracle_compute_keff(o);
    if (i > 0 && o->stabilize) moc_oracle_stabilize_flux(o);
    moc_oracle_normalize_fluxes(o);
    residual = moc_oracle_compute_residual(o, res_type);
    int dk = 1e5 * (o->k_eff - k_prev);
    previous_residual = residual;
    k_prev = o->k_eff;
    moc_oracle_store_fsr_fluxes(o);
    num_iterations++;
    if (residual < tol && abs(dk) < 1) break;
  }
  (void)previous_residual;
  return num_iterations;
}

/* src/Solver.cpp:1352-1420 */
int moc_oracle_compute_flux(moc_oracle* o, int max_iters, double tol, int only_fixed_source) {
  o->k_eff = 1.;
  double residual = 0.;
  if (only_fixed_source) {
    moc_oracle_zero_track_fluxes(o);
    moc_oracle_flatten_fsr_fluxes(o, 0.);
    moc_oracle_store_fsr_fluxes(o);
  }
  moc_oracle_compute_fsr_sources(o, 0);
  for (int i = 0; i < max_iters; i++) {
    moc_oracle_transport_sweep(o);
    moc_oracle_add_source_to_scalar_flux(o);
    residual = moc_oracle_compute_residual(o, MOC_RES_SCALAR_FLUX);
    moc_oracle_store_fsr_fluxes(o);
    if (i > 1 && residual < tol) return i;
  }
  return max_iters;
}

/* src/Solver.cpp:1459-1516 */
int moc_oracle_compute_source(moc_oracle* o, int max_iters, double k_eff, double tol, int res_type) {
  o->k_eff = k_eff;
  double residual = 0.;
  moc_oracle_zero_track_fluxes(o);
  moc_oracle_flatten_fsr_fluxes(o, 1.0);      /* computeInitialFluxGuess(true) */
  moc_oracle_store_fsr_fluxes(o);
  for (int i = 0; i < max_iters; i++) {
    moc_oracle_compute_fsr_sources(o, i);
    moc_oracle_transport_sweep(o);
    moc_oracle_add_source_to_scalar_flux(o);
    residual = moc_oracle_compute_residual(o, res_type);
    moc_oracle_store_fsr_fluxes(o);
    if (i > 1 && residual < tol) return i;
  }
  return max_iters;
}

double moc_oracle_get_keff(moc_oracle* o) { return o->k_eff; }
void moc_oracle_set_keff(moc_oracle* o, double k) { o->k_eff = k; }
void moc_oracle_get_fluxes(moc_oracle* o, double* out) { memcpy(out, o->phi, (size_t)o->n_fsr * o->G * 8); }
void moc_oracle_set_fluxes(moc_oracle* o, const double* in) { memcpy(o->phi, in, (size_t)o->n_fsr * o->G * 8); }
void moc_oracle_get_sources(moc_oracle* o, double* out) { memcpy(out, o->q, (size_t)o->n_fsr * o->G * 8); }
void moc_oracle_set_sources(moc_oracle* o, const double* in) { memcpy(o->q, in, (size_t)o->n_fsr * o->G * 8); }
void moc_oracle_get_start_fluxes(moc_oracle* o, float* out) { memcpy(out, o->psi_start, (size_t)o->n_trk * 2 * o->F * 4); }
void moc_oracle_set_start_fluxes(moc_oracle* o, const float* in) { memcpy(o->psi_start, in, (size_t)o->n_trk * 2 * o->F * 4); }

/* src/Solver.cpp:479-497 + src/CPUSolver.cpp:425-456 (group0 is 0-based here) */
void moc_oracle_set_fixed_source(moc_oracle* o, int64_t fsr, int group0, double value) {
  o->fixed_on = 1;
  o->fixed[fsr * o->G + group0] = value;
}

/* src/Solver.cpp stabilizeTransport */
void moc_oracle_stabilize_transport(moc_oracle* o, double factor, int type) {
  o->stabilize = 1; o->stab_factor = factor; o->stab_type = type;
}

/* src/CPUSolver.cpp:2825-2856 */
void moc_oracle_compute_fission_rates(moc_oracle* o, double* out, int nu) {
  int G = o->G;
  for (int64_t r = 0; r < o->n_fsr; r++) {
    const double* sig = (nu ? o->nu_sigma_f : o->sigma_f) + (size_t)o->fsr_mat[r] * G;
    out[r] = 0.;
    for (int e = 0; e < G; e++) out[r] += sig[e] * o->phi[r * G + e] * o->vol[r];
  }
}
