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
  int neg_allowed;            /* Solver::allowNegativeFluxes (src/Solver.cpp) */
  double* stab_m;             /* stabilising flux moments [r][3][G] (CPULSSolver.cpp:888-970) */
  double* fixed_m;            /* fixed source moments [r][3][G] (CPULSSolver.cpp:154-205), NULL until set */
  float* leakage; double* sigma_a;
  double stab_factor;
  double sweep_seconds;
  double* scratch;
  /* linear source (CPULSSolver): see moc_oracle_enable_linear_source */
  int ls, nc;                       /* nc: 3 coefficients in 2D, 6 in 3D */
  double* seg_start;                /* [n_seg][3] relative to the FSR centroid; walked in place by the sweep */
  double* seg_start0;               /* as uploaded: 3D (on-the-fly) segments are re-traced before every sweep */
  double* trk_phi; double* trk_theta;
  double* phi_m; double* q_m;       /* [r][c][e] = r*3G + c*G + e   (CPULSSolver.h:22-26) */
  double* lin_exp;                  /* [r][nc] inverse expansion matrix */
  double* src_const;                /* [r][i][e] = r*G*nc + i*G + e */
  int num_flat;
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
  free(o->phi); free(o->phi_old); free(o->q); free(o->fixed); free(o->stab); free(o->fixed_m); free(o->stab_m);
  free(o->psi_start); free(o->psi_bound); free(o->scratch); free(o->leakage); free(o->sigma_a);
  free(o->seg_start); free(o->seg_start0); free(o->trk_phi); free(o->trk_theta); free(o->phi_m); free(o->q_m);
  free(o->lin_exp); free(o->src_const);
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
  if (o->ls) memset(o->phi_m, 0, (size_t)o->n_fsr * o->G * 3 * 8);   /* src/CPULSSolver.cpp:342-354 */
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
  if (o->ls)                          /* src/CPULSSolver.cpp:360-372 */
    for (int64_t i = 0; i < o->n_fsr * G * 3; i++) o->phi_m[i] *= norm_factor;
  return norm_factor;
}

static void ls_sources(moc_oracle* o, int iteration);

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
        if (q < 0.0 && iteration < 30 && !o->neg_allowed) q = FLUX_EPSILON;     /* CPUSolver.cpp:1981 */
        o->q[r * G + Gd] = q;
      }
    }
    free(fs); free(ss);
  }
  if (o->ls) ls_sources(o, iteration);
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


/* ===================================================================================== */
/* Linear source: restatement of CPULSSolver (src/CPULSSolver.cpp) and of its pre-pass     */
/* LinearExpansionGenerator (src/TrackTraversingAlgorithms.cpp:470-831)                   */
/* ===================================================================================== */
#define MIN_DET 1E-10                     /* src/constants.h:70 */

/* src/exponentials.h:110-145: 1/x - (1-exp(-x))/x^2 */
static inline double expG_fractional(double x) {
  const double p0 = 0.5;
  const double p1 = 1.76558112351595 * 1E-1;
  const double p2 = 4.041584305811143 * 1E-2;
  const double p3 = 6.178333902037397 * 1E-3;
  const double p4 = 6.429894635552992 * 1E-4;
  const double p5 = 6.064409107557148 * 1E-5;
  const double d0 = 1.0;
  const double d1 = 6.864462055546078 * 1E-1;
  const double d2 = 2.263358514260129 * 1E-1;
  const double d3 = 4.721469893686252 * 1E-2;
  const double d4 = 6.883236664917246 * 1E-3;
  const double d5 = 7.036272419147752 * 1E-4;
  const double d6 = 6.064409107557148 * 1E-5;
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

/* src/exponentials.h:293-323 */
static inline double expG2_fractional(double x) {
  const double a1 = -8.335775885589858 * 1E-2;
  const double a2 = -3.603942303847604 * 1E-3;
  const double a3 = 3.7673183263550827 * 1E-3;
  const double a4 = 1.124183494990467 * 1E-5;
  const double a5 = 1.6837426505799449 * 1E-4;
  const double b1 = 7.454048371823628 * 1E-1;
  const double b2 = 2.3794300531408347 * 1E-1;
  const double b3 = 5.367250964303789 * 1E-2;
  const double b4 = 6.125197988351906 * 1E-3;
  const double b5 = 1.0102514456857377 * 1E-3;
  double num, den;
  num = a5 * x + a4;
  num = num * x + a3;
  num = num * x + a2;
  num = num * x + a1;
  num *= x;
  den = b5 * x + b4;
  den = den * x + b3;
  den = den * x + b2;
  den = den * x + b1;
  den = den * x + 1.;
  return num / den;
}

/* LinearExpansionGenerator::onTrack + execute (TrackTraversingAlgorithms.cpp:536-831) */
static void ls_prepass(moc_oracle* o, const double* azim_spacing, const double* azim_weight,
                       const double* polar_spacing, const double* polar_weight) {
  const int G = o->G, nc = o->nc, P = o->P;
  double* lem = (double*)calloc((size_t)o->n_fsr * nc, 8);
  double* tsc = (double*)malloc((size_t)G * nc * 8);
  memset(o->src_const, 0, (size_t)o->n_fsr * G * nc * 8);
  for (int64_t t = 0; t < o->n_trk; t++) {
    const int azim = o->trk_azim[t], polar = o->trk_polar[t];
    const double phi = o->trk_phi[t];
    const double sin_phi = sin(phi), cos_phi = cos(phi);
    double wgt = azim_spacing[azim] * azim_weight[azim];
    double sin_theta = 1, cos_theta = 0;
    if (o->solve_3d) {
      const double theta = o->trk_theta[t];
      sin_theta = sin(theta);
      cos_theta = cos(theta);
      wgt *= polar_spacing[azim * P + polar] * polar_weight[azim * P + polar];
    }
    for (int64_t s = o->trk_off[t]; s < o->trk_off[t + 1]; s++) {
      const int64_t fsr = o->seg_fsr[s];
      const double* sigma_t = o->sigma_t + (size_t)o->fsr_mat[fsr] * G;
      const double length = o->seg_len[s];
      const double volume = o->vol[fsr];
      const double x = o->seg_start[3 * s], y = o->seg_start[3 * s + 1], z = o->seg_start[3 * s + 2];
      const double xc = x + length * 0.5 * cos_phi * sin_theta;
      const double yc = y + length * 0.5 * sin_phi * sin_theta;
      const double zc = z + length * 0.5 * cos_theta;
      const double vol_impact = wgt * length / volume;
      const double src_constant = vol_impact * length / 2.0;
      for (int g = 0; g < G; g++) {
        tsc[g] = vol_impact * xc * xc;
        tsc[G + g] = vol_impact * yc * yc;
        tsc[2 * G + g] = vol_impact * xc * yc;
        if (o->solve_3d) {
          tsc[3 * G + g] = vol_impact * xc * zc;
          tsc[4 * G + g] = vol_impact * yc * zc;
          tsc[5 * G + g] = vol_impact * zc * zc;
        }
        const double tau = length * sigma_t[g];
        if (!o->solve_3d) {
          for (int p = 0; p < P / 2; p++) {
            const double st = o->sin_theta[azim * P + p];
            const double G2_src = length * expG2_fractional(tau / st) * src_constant * 2
                                  * polar_weight[azim * P + p] * st;
            tsc[g] += cos_phi * cos_phi * G2_src;
            tsc[G + g] += sin_phi * sin_phi * G2_src;
            tsc[2 * G + g] += sin_phi * cos_phi * G2_src;
          }
        } else {
          const double G2_src = expG2_fractional(tau) * length * src_constant;
          tsc[g] += cos_phi * cos_phi * G2_src * sin_theta * sin_theta;
          tsc[G + g] += sin_phi * sin_phi * G2_src * sin_theta * sin_theta;
          tsc[2 * G + g] += sin_phi * cos_phi * G2_src * sin_theta * sin_theta;
          tsc[3 * G + g] += cos_phi * cos_theta * G2_src * sin_theta;
          tsc[4 * G + g] += sin_phi * cos_theta * G2_src * sin_theta;
          tsc[5 * G + g] += cos_theta * cos_theta * G2_src;
        }
      }
      lem[fsr * nc] += wgt * length / volume * (xc * xc + pow(cos_phi * sin_theta * length, 2) / 12.0);
      lem[fsr * nc + 1] += wgt * length / volume * (yc * yc + pow(sin_phi * sin_theta * length, 2) / 12.0);
      lem[fsr * nc + 2] += wgt * length / volume * (xc * yc + sin_phi * cos_phi * pow(sin_theta * length, 2) / 12.0);
      if (o->solve_3d) {
        lem[fsr * nc + 3] += wgt * length / volume * (xc * zc + cos_phi * cos_theta * sin_theta * pow(length, 2) / 12.0);
        lem[fsr * nc + 4] += wgt * length / volume * (yc * zc + sin_phi * cos_theta * sin_theta * pow(length, 2) / 12.0);
        lem[fsr * nc + 5] += wgt * length / volume * (zc * zc + pow(cos_theta * length, 2) / 12.0);
      }
      for (int g = 0; g < G; g++)
        for (int i = 0; i < nc; i++) o->src_const[fsr * G * nc + i * G + g] += tsc[i * G + g];
    }
  }
  /* invert the symmetric expansion matrix per FSR (:570-633) */
  double* ilem = o->lin_exp;
  o->num_flat = 0;
  for (int64_t r = 0; r < o->n_fsr; r++) {
    if (o->solve_3d) {
      double det = lem[r*nc + 0] * lem[r*nc + 1] * lem[r*nc + 5] + lem[r*nc + 2] * lem[r*nc + 4] * lem[r*nc + 3] +
                   lem[r*nc + 3] * lem[r*nc + 2] * lem[r*nc + 4] - lem[r*nc + 0] * lem[r*nc + 4] * lem[r*nc + 4] -
                   lem[r*nc + 3] * lem[r*nc + 1] * lem[r*nc + 3] - lem[r*nc + 2] * lem[r*nc + 2] * lem[r*nc + 5];
      if (fabs(det) < MIN_DET || o->vol[r] < 1e-6) {
        o->num_flat++;
        for (int i = 0; i < 6; i++) ilem[r*nc + i] = 0.0;
      } else {
        ilem[r*nc + 0] = (lem[r*nc + 1] * lem[r*nc + 5] - lem[r*nc + 4] * lem[r*nc + 4]) / det;
        ilem[r*nc + 1] = (lem[r*nc + 0] * lem[r*nc + 5] - lem[r*nc + 3] * lem[r*nc + 3]) / det;
        ilem[r*nc + 2] = (lem[r*nc + 3] * lem[r*nc + 4] - lem[r*nc + 2] * lem[r*nc + 5]) / det;
        ilem[r*nc + 3] = (lem[r*nc + 2] * lem[r*nc + 4] - lem[r*nc + 3] * lem[r*nc + 1]) / det;
        ilem[r*nc + 4] = (lem[r*nc + 3] * lem[r*nc + 2] - lem[r*nc + 0] * lem[r*nc + 4]) / det;
        ilem[r*nc + 5] = (lem[r*nc + 0] * lem[r*nc + 1] - lem[r*nc + 2] * lem[r*nc + 2]) / det;
      }
    } else {
      double det = lem[r*nc] * lem[r*nc + 1] - lem[r*nc + 2] * lem[r*nc + 2];
      if (fabs(det) < MIN_DET) {
        o->num_flat++;
        ilem[r*nc] = ilem[r*nc + 1] = ilem[r*nc + 2] = 0.0;
      } else {
        ilem[r*nc + 0] = lem[r*nc + 1] / det;
        ilem[r*nc + 1] = lem[r*nc + 0] / det;
        ilem[r*nc + 2] = -lem[r*nc + 2] / det;
      }
    }
  }
  free(lem); free(tsc);
}

int moc_oracle_enable_linear_source(moc_oracle* o, const double* seg_start, const double* trk_phi,
                                    const double* trk_theta, const double* azim_spacing,
                                    const double* azim_weight, const double* polar_spacing,
                                    const double* polar_weight) {
  const int G = o->G;
  o->ls = 1;
  o->nc = o->solve_3d ? 6 : 3;
  o->seg_start = dup_mem(seg_start, (size_t)o->n_seg * 3 * 8);
  o->seg_start0 = dup_mem(seg_start, (size_t)o->n_seg * 3 * 8);
  o->trk_phi = dup_mem(trk_phi, (size_t)o->n_trk * 8);
  o->trk_theta = dup_mem(trk_theta, (size_t)o->n_trk * 8);
  o->phi_m = calloc((size_t)o->n_fsr * G * 3, 8);
  o->q_m = calloc((size_t)o->n_fsr * G * 3, 8);
  o->lin_exp = calloc((size_t)o->n_fsr * o->nc, 8);
  o->src_const = calloc((size_t)o->n_fsr * G * o->nc, 8);
  ls_prepass(o, azim_spacing, azim_weight, polar_spacing, polar_weight);
  return o->num_flat;
}

void moc_oracle_get_flux_moments(moc_oracle* o, double* out) {
  memcpy(out, o->phi_m, (size_t)o->n_fsr * o->G * 3 * 8);
}
void moc_oracle_get_linear_source_tables(moc_oracle* o, double* lin_exp, double* src_const) {
  memcpy(lin_exp, o->lin_exp, (size_t)o->n_fsr * o->nc * 8);
  memcpy(src_const, o->src_const, (size_t)o->n_fsr * o->G * o->nc * 8);
}

/* CPULSSolver::computeFSRSources, moment part (src/CPULSSolver.cpp:386-524) */
static void ls_sources(moc_oracle* o, int iteration) {
  const int G = o->G, nc = o->nc;
  double* buf = (double*)malloc(G * 8);
  for (int64_t r = 0; r < o->n_fsr; r++) {
    const int m = o->fsr_mat[r];
    const double* sigma_s = o->sigma_s + (size_t)m * G * G;
    const double* fm = o->fiss + (size_t)m * G * G;
    const double* pm = o->phi_m + r * 3 * G;
    for (int g = 0; g < G; g++) {
      double fis[3] = {0., 0., 0.}, sca[3];
      if (o->fissionable[m]) {
        for (int c = 0; c < 3; c++) {
          for (int gp = 0; gp < G; gp++) buf[gp] = fm[g * G + gp] * pm[c * G + gp];
          fis[c] = pairwise_sum(buf, G) / o->k_eff;
        }
      }
      for (int c = 0; c < 3; c++) {
        for (int gp = 0; gp < G; gp++) buf[gp] = sigma_s[g * G + gp] * pm[c * G + gp];
        sca[c] = pairwise_sum(buf, G);
      }
      double src_x = sca[0] + fis[0], src_y = sca[1] + fis[1], src_z = sca[2] + fis[2];
      if (o->fixed_m != NULL) {          /* _fixed_source_moments_on, CPULSSolver.cpp:475-479 */
        src_x += o->fixed_m[r * 3 * G + g];
        src_y += o->fixed_m[r * 3 * G + G + g];
        src_z += o->fixed_m[r * 3 * G + 2 * G + g];
      }
      double* qm = o->q_m + r * 3 * G;
      const double* M = o->lin_exp + r * nc;
      if (o->neg_allowed || o->q[r * G + g] > 10 * FLUX_EPSILON || iteration > 29) {
        if (o->solve_3d) {
          qm[g] = ONE_OVER_FOUR_PI / 2 * (M[0] * src_x + M[2] * src_y + M[3] * src_z);
          qm[G + g] = ONE_OVER_FOUR_PI / 2 * (M[2] * src_x + M[1] * src_y + M[4] * src_z);
          qm[2 * G + g] = ONE_OVER_FOUR_PI / 2 * (M[3] * src_x + M[4] * src_y + M[5] * src_z);
        } else {
          qm[g] = ONE_OVER_FOUR_PI / 2 * (M[0] * src_x + M[2] * src_y);
          qm[G + g] = ONE_OVER_FOUR_PI / 2 * (M[2] * src_x + M[1] * src_y);
        }
      } else {
        qm[g] = qm[G + g] = 0;
        if (o->solve_3d) qm[2 * G + g] = 0;
      }
    }
  }
  free(buf);
}

/* One track, both directions, linear source: TransportSweep::onTrack
 * (TrackTraversingAlgorithms.cpp:890-1052) with CPULSSolver::tallyLSScalarFlux
 * (CPULSSolver.cpp:542-739) and accumulateLinearFluxContribution (:749-780) inlined.
 * buf = 4*G doubles: flux, x, y, z contributions of the current FSR run. */
static void ls_sweep_track(moc_oracle* o, int64_t t, double* buf) {
  const int G = o->G, NP = o->NP, F = o->F, P = o->P;
  const int azim = o->trk_azim[t], polar = o->trk_polar[t];
  const int64_t s0 = o->trk_off[t], s1 = o->trk_off[t + 1];
  const double* wrow = o->weight + (size_t)azim * P;
  int a_eval = azim;
  if (a_eval >= o->A / 4) a_eval = o->A / 2 - 1 - azim;
  const double* srow = o->sin_theta + (size_t)a_eval * P;
  const double weight3d = o->solve_3d ? wrow[polar] : 1.0;
  double* fx = buf + G; double* fy = buf + 2 * G; double* fz = buf + 3 * G;
  double direction[3];
  {
    const double phi = o->trk_phi[t];
    double cos_theta = 0.0, sin_theta = 1.0;
    if (o->solve_3d) { cos_theta = cos(o->trk_theta[t]); sin_theta = sin(o->trk_theta[t]); }
    direction[0] = cos(phi) * sin_theta;
    direction[1] = sin(phi) * sin_theta;
    direction[2] = cos_theta;
  }
  memset(buf, 0, 4 * G * 8);
  for (int dir = 0; dir < 2; dir++) {
    float* track_flux = o->psi_bound + ((size_t)t * 2 + dir) * F;
    int64_t s = dir == 0 ? s0 : s1 - 1;
    const int64_t step = dir == 0 ? 1 : -1;
    for (int64_t n = 0; n < s1 - s0; n++, s += step) {
      const int64_t fsr = o->seg_fsr[s];
      const double length = o->seg_len[s];
      const double* sigma_t = o->sigma_t + (size_t)o->fsr_mat[fsr] * G;
      const double* q = o->q + fsr * G;
      const double* qm = o->q_m + fsr * 3 * G;
      double* position = o->seg_start + 3 * s;
      if (o->solve_3d) {
        double center_x2[3];
        for (int i = 0; i < 3; i++) center_x2[i] = 2 * position[i] + length * direction[i];
        for (int e = 0; e < G; e++) {
          double src_flat = q[e];
          for (int i = 0; i < 3; i++) src_flat += qm[i * G + e] * center_x2[i];
          double src_linear = qm[e] * direction[0];
          src_linear += qm[G + e] * direction[1];
          src_linear += qm[2 * G + e] * direction[2];
          const double tau = length * sigma_t[e];
          const double exp_G = expG_fractional(tau > 1e-8 ? tau : 1e-8);
          const double exp_F1 = 1. - tau * exp_G;
          const double exp_F2 = 2. * exp_G - exp_F1;
          double exp_H = exp_F1 - exp_G;
          exp_H *= length * track_flux[e] * tau;
          const double delta_psi = (tau * track_flux[e] - length * src_flat) * exp_F1
                                   - src_linear * length * length * exp_F2;
          track_flux[e] -= delta_psi;
          buf[e] += delta_psi;
          fx[e] += exp_H * direction[0] + delta_psi * position[0];
          fy[e] += exp_H * direction[1] + delta_psi * position[1];
          fz[e] += exp_H * direction[2] + delta_psi * position[2];
        }
      } else {
        double center[2];
        for (int i = 0; i < 2; i++) center[i] = 2 * position[i] + length * direction[i];
        for (int p = 0; p < NP; p++) {
          const double inv_sin_theta = 1.0 / srow[p];
          const double wgt = wrow[p];
          for (int e = 0; e < G; e++) {
            const int pe = p * G + e;
            const double tau = sigma_t[e] * length;
            /* ExpEvaluator::retrieveExponentialComponents, src/ExpEvaluator.h:349-381 */
            double tp = tau * inv_sin_theta;
            if (tp < 1e-8) tp = 1e-8;
            double exp_G = expG_fractional(tp);
            double exp_F1 = 1. - tp * exp_G;
            exp_F1 *= inv_sin_theta;
            exp_G *= inv_sin_theta;
            const double exp_F2 = 2. * exp_G - exp_F1;
            double exp_H = exp_F1 - exp_G;
            double src_flat = q[e];
            for (int i = 0; i < 2; i++) src_flat += qm[i * G + e] * center[i];
            double src_linear = direction[0] * qm[e];
            src_linear += direction[1] * qm[G + e];
            exp_H *= wgt * tau * length * track_flux[pe];
            double delta_psi = (tau * track_flux[pe] - length * src_flat) * exp_F1
                               - length * length * src_linear * exp_F2;
            track_flux[pe] -= delta_psi;
            delta_psi *= wgt;
            buf[e] += delta_psi;
            fx[e] += exp_H * direction[0] + delta_psi * position[0];
            fy[e] += exp_H * direction[1] + delta_psi * position[1];
          }
        }
      }
      /* move the starting position to the end of the segment for the opposite direction (:736-738) */
      for (int i = 0; i < 3; i++) position[i] += direction[i] * length;

      int flush;
      if (dir == 0) flush = (s < s1 - 1) && (fsr != o->seg_fsr[s + 1]);
      else flush = (s == s0) || (fsr != o->seg_fsr[s - 1]);
      if (flush) {
        double* pm = o->phi_m + fsr * 3 * G;
        for (int e = 0; e < G; e++) {
#pragma omp atomic update
          o->phi[fsr * G + e] += weight3d * buf[e];
#pragma omp atomic update
          pm[e] += weight3d * fx[e];
#pragma omp atomic update
          pm[G + e] += weight3d * fy[e];
#pragma omp atomic update
          pm[2 * G + e] += weight3d * fz[e];
        }
        memset(buf, 0, 4 * G * 8);
      }
    }
    uint8_t bc = dir == 0 ? o->bc_fwd[t] : o->bc_bwd[t];
    int64_t nxt = dir == 0 ? o->next_fwd[t] : o->next_bwd[t];
    int next_is_fwd = dir == 0 ? (o->flags[t] & 1) : ((o->flags[t] >> 1) & 1);
    if (bc == BC_REFLECTIVE || bc == BC_PERIODIC) {
      float* out = o->psi_start + ((size_t)nxt * 2 + (next_is_fwd ? 0 : 1)) * F;
      memcpy(out, track_flux, F * 4);
    }
    for (int i = 0; i < 3; i++) direction[i] *= -1;     /* reverse the direction (:1006-1008) */
  }
}

/* CPULSSolver::addSourceToScalarFlux (src/CPULSSolver.cpp:787-882) */
static void ls_closure(moc_oracle* o) {
  const int G = o->G, nc = o->nc;
  for (int64_t r = 0; r < o->n_fsr; r++) {
    double volume = o->vol[r];
    if (volume < VOL_EPSILON) volume = 1e30;
    const double* sigma_t = o->sigma_t + (size_t)o->fsr_mat[r] * G;
    double* pm = o->phi_m + r * 3 * G;
    const double* qm = o->q_m + r * 3 * G;
    const double* sc = o->src_const + r * G * nc;
    for (int e = 0; e < G; e++) {
      const double flux_const = FOUR_PI * 2;
      o->phi[r * G + e] /= volume;
      o->phi[r * G + e] += FOUR_PI * o->q[r * G + e];
      o->phi[r * G + e] /= sigma_t[e];
      pm[e] /= volume;
      pm[e] += flux_const * qm[e] * sc[e];
      pm[e] += flux_const * qm[G + e] * sc[2 * G + e];
      pm[G + e] /= volume;
      pm[G + e] += flux_const * qm[e] * sc[2 * G + e];
      pm[G + e] += flux_const * qm[G + e] * sc[G + e];
      if (o->solve_3d) {
        pm[e] += flux_const * qm[2 * G + e] * sc[3 * G + e];
        pm[G + e] += flux_const * qm[2 * G + e] * sc[4 * G + e];
        pm[2 * G + e] /= volume;
        pm[2 * G + e] += flux_const * qm[e] * sc[3 * G + e];
        pm[2 * G + e] += flux_const * qm[G + e] * sc[4 * G + e];
        pm[2 * G + e] += flux_const * qm[2 * G + e] * sc[5 * G + e];
      }
      pm[e] /= sigma_t[e];
      pm[G + e] /= sigma_t[e];
      if (o->solve_3d) pm[2 * G + e] /= sigma_t[e];
      if (o->phi[r * G + e] < 0.0 && !o->neg_allowed) {
        o->phi[r * G + e] = o->phi_old[r * G + e] > FLUX_EPSILON ? o->phi_old[r * G + e] : FLUX_EPSILON;
        pm[e] = pm[G + e] = pm[2 * G + e] = 0;
      }
    }
  }
}

/* src/CPUSolver.cpp:2338-2389 */
void moc_oracle_transport_sweep(moc_oracle* o) {
  double t0 = omp_get_wtime();
  memset(o->phi, 0, (size_t)o->n_fsr * o->G * 8);                    /* :2347 */
  memcpy(o->psi_bound, o->psi_start, (size_t)o->n_trk * 2 * o->F * 4); /* :2351 */
  memset(o->leakage, 0, (size_t)o->n_trk * 4);                          /* :2360 */
  if (o->ls) memset(o->phi_m, 0, (size_t)o->n_fsr * o->G * 3 * 8);     /* CPULSSolver::flattenFSRFluxes */
  /* explicit 2D segments keep the positions the previous sweep left (forward: += d*l, backward:
   * -= d*l, CPULSSolver.cpp:736-738); on-the-fly 3D segments are traced afresh every sweep */
  if (o->ls && o->solve_3d) memcpy(o->seg_start, o->seg_start0, (size_t)o->n_seg * 3 * 8);
#pragma omp parallel num_threads(o->threads)
  {
    double* fsr_flux = (double*)malloc(4 * o->G * 8);
#pragma omp for schedule(dynamic)
    for (int64_t t = 0; t < o->n_trk; t++) {
      if (o->ls) ls_sweep_track(o, t, fsr_flux);
      else sweep_track(o, t, fsr_flux);
    }
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
  if (o->ls) { ls_closure(o); return; }
  for (int64_t r = 0; r < o->n_fsr; r++) {
    double volume = o->vol[r];
    const double* sigma_t = o->sigma_t + (size_t)o->fsr_mat[r] * G;
    if (volume < VOL_EPSILON) volume = 1e30;
    for (int e = 0; e < G; e++) {
      o->phi[r * G + e] /= (sigma_t[e] * volume);
      o->phi[r * G + e] += FOUR_PI * o->q[r * G + e] / sigma_t[e];
      if (o->phi[r * G + e] < 0.0 && !o->neg_allowed) o->phi[r * G + e] = FLUX_EPSILON;   /* CPUSolver.cpp:2630 */
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
  if (!o->ls) return;
  /* CPULSSolver::computeStabilizingFlux, moment part (src/CPULSSolver.cpp:894-970; _stabilize_moments = true) */
  if (o->stab_m == NULL) o->stab_m = calloc((size_t)o->n_fsr * 3 * G, 8);
  for (int64_t r = 0; r < o->n_fsr; r++) {
    int m = o->fsr_mat[r];
    for (int e = 0; e < G; e++)
      for (int c = 0; c < 3; c++) {
        const size_t k = (size_t)r * 3 * G + c * G + e;
        if (o->stab_type == STAB_DIAGONAL) {
          double sigma_s = o->sigma_s[(size_t)m * G * G + e * G + e];
          if (sigma_s < 0.0) o->stab_m[k] = -o->phi_m[k] * o->stab_factor * sigma_s / o->sigma_t[(size_t)m * G + e];
        } else if (o->stab_type == STAB_YAMAMOTO) {
          o->stab_m[k] = o->phi_m[k] * 0.0;      /* max_ratio stays 0 in the reference: `ratio = max_ratio` (:943) */
        } else {
          o->stab_m[k] = o->phi_m[k] * (1.0 / o->stab_factor - 1.0);
        }
      }
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
  if (!o->ls || o->stab_m == NULL) return;
  /* CPULSSolver::stabilizeFlux, moment part (src/CPULSSolver.cpp:978-1052) */
  for (int64_t r = 0; r < o->n_fsr; r++) {
    int m = o->fsr_mat[r];
    for (int e = 0; e < G; e++)
      for (int c = 0; c < 3; c++) {
        const size_t k = (size_t)r * 3 * G + c * G + e;
        if (o->stab_type == STAB_DIAGONAL) {
          double sigma_s = o->sigma_s[(size_t)m * G * G + e * G + e];
          if (sigma_s < 0.0) {
            o->phi_m[k] += o->stab_m[k];
            o->phi_m[k] /= (1.0 - o->stab_factor * sigma_s / o->sigma_t[(size_t)m * G + e]);
          }
        } else if (o->stab_type == STAB_YAMAMOTO) {
          o->phi_m[k] += o->stab_m[k];
          o->phi_m[k] /= (1 + 0.0);
        } else {
          o->phi_m[k] += o->stab_m[k];
          o->phi_m[k] *= o->stab_factor;
        }
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
/* CPULSSolver::setFixedSourceMomentByFSR (src/CPULSSolver.cpp): volume-averaged x, y, z moments */
void moc_oracle_set_fixed_source_moments(moc_oracle* o, int64_t fsr, int group0, double sx, double sy, double sz) {
  if (o->fixed_m == NULL) o->fixed_m = calloc((size_t)o->n_fsr * 3 * o->G, 8);
  o->fixed_m[fsr * 3 * o->G + group0] = sx;
  o->fixed_m[fsr * 3 * o->G + o->G + group0] = sy;
  o->fixed_m[fsr * 3 * o->G + 2 * o->G + group0] = sz;
}
void moc_oracle_allow_negative_fluxes(moc_oracle* o, int allowed) { o->neg_allowed = allowed != 0; }

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
