/*
 * ls_prepass.cuh - the linear-source pre-pass on the device.
 *
 * Replaces LinearExpansionGenerator::onTrack / execute (src/TrackTraversingAlgorithms.cpp:470-831): per FSR
 * the symmetric moment matrix of the track-based centroid expansion (3 entries in 2D, 6 in 3D), inverted
 * (`_FSR_lin_exp_matrix`), and per (FSR, group) the source constants (`_FSR_source_constants`,
 * src/CPULSSolver.h:36-50).  One thread per segment, tallies with RED.ADD.F64; a second kernel inverts.
 * The reference's host pre-pass stays available (the plug-in inherits it from CPULSSolver); this is what a
 * track file alone needs (Python B200Solver) and what keeps setup off the host for large decks.
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

/* src/exponentials.h:293-323, the 5/5-order rational */
__device__ __forceinline__ double ls_expG2(double x) {
  const double a1 = -8.335775885589858e-2, a2 = -3.603942303847604e-3, a3 = 3.7673183263550827e-3,
               a4 = 1.124183494990467e-5, a5 = 1.6837426505799449e-4;
  const double b1 = 7.454048371823628e-1, b2 = 2.3794300531408347e-1, b3 = 5.367250964303789e-2,
               b4 = 6.125197988351906e-3, b5 = 1.0102514456857377e-3;
  double num = a5 * x + a4;
  num = num * x + a3; num = num * x + a2; num = num * x + a1; num = num * x;
  double den = b5 * x + b4;
  den = den * x + b3; den = den * x + b2; den = den * x + b1; den = den * x + 1.0;
  return num / den;
}

struct LsPrepassArgs {
  int G, P, solve_3d, nc;
  int64_t n_trk, n_seg, n_fsr;
  const double* seg_len; const int32_t* seg_fsr; const double* seg_start;   /* [n_seg][3], centroid-relative */
  const int64_t* trk_off; const int32_t* trk_azim; const int32_t* trk_polar;
  const double* trk_phi; const double* trk_theta;
  const double* azim_spacing; const double* azim_weight;                    /* [A/2] */
  const double* polar_spacing; const double* polar_weight; const double* sin_theta;   /* [A/2][P] */
  const double* volume; const int32_t* fsr_mat; const double* sigma_t;      /* [n_mat][G] */
  double* lem;          /* [n_fsr][nc] moment matrix, accumulated */
  double* src_const;    /* [n_fsr][nc][G], accumulated */
};

/* one thread per track walks its segments (the segment -> track map then needs no search) */
__global__ void ls_prepass_kernel(const LsPrepassArgs a) {
  const int G = a.G, P = a.P, nc = a.nc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.n_trk; t += (int64_t)gridDim.x * blockDim.x) {
    const int azim = a.trk_azim[t];
    const double phi = a.trk_phi[t];
    const double sin_phi = sin(phi), cos_phi = cos(phi);
    double wgt = a.azim_spacing[azim] * a.azim_weight[azim];
    double sin_t = 1.0, cos_t = 0.0;
    if (a.solve_3d) {
      const int polar = a.trk_polar[t];
      const double theta = a.trk_theta[t];
      sin_t = sin(theta); cos_t = cos(theta);
      wgt *= a.polar_spacing[azim * P + polar] * a.polar_weight[azim * P + polar];
    }
    for (int64_t s = a.trk_off[t]; s < a.trk_off[t + 1]; s++) {
      const int64_t r = a.seg_fsr[s];
      const double length = a.seg_len[s];
      const double volume = a.volume[r];
      const double xc = a.seg_start[3 * s] + length * 0.5 * cos_phi * sin_t;
      const double yc = a.seg_start[3 * s + 1] + length * 0.5 * sin_phi * sin_t;
      const double zc = a.seg_start[3 * s + 2] + length * 0.5 * cos_t;
      const double vol_impact = wgt * length / volume;
      const double src_constant = vol_impact * length / 2.0;
      const double l2 = length * length;
      double geo[6] = {xc * xc, yc * yc, xc * yc, xc * zc, yc * zc, zc * zc};
      double dirt[6] = {cos_phi * cos_phi * sin_t * sin_t, sin_phi * sin_phi * sin_t * sin_t, sin_phi * cos_phi * sin_t * sin_t,
                        cos_phi * cos_t * sin_t, sin_phi * cos_t * sin_t, cos_t * cos_t};
      for (int i = 0; i < nc; i++) atomicAdd(&a.lem[r * nc + i], vol_impact * (geo[i] + dirt[i] * l2 / 12.0));
      const double* sig = a.sigma_t + (int64_t)a.fsr_mat[r] * G;
      for (int g = 0; g < G; g++) {
        const double tau = length * sig[g];
        if (a.solve_3d) {
          const double g2 = ls_expG2(tau) * (length * src_constant);
          for (int i = 0; i < nc; i++) atomicAdd(&a.src_const[(r * nc + i) * G + g], vol_impact * geo[i] + dirt[i] * g2);
        } else {
          double acc = 0.0;      /* polar sum of 2 * w_p * sin(theta_p) * G2(tau / sin(theta_p)) */
          for (int p = 0; p < P / 2; p++) {
            const double st = a.sin_theta[azim * P + p];
            acc += length * ls_expG2(tau / st) * (src_constant * 2 * a.polar_weight[azim * P + p] * st);
          }
          const double d2[3] = {cos_phi * cos_phi, sin_phi * sin_phi, sin_phi * cos_phi};
          for (int i = 0; i < 3; i++) atomicAdd(&a.src_const[(r * nc + i) * G + g], vol_impact * geo[i] + d2[i] * acc);
        }
      }
    }
  }
}

/* inverse of the symmetric 2x2 / 3x3 moment matrix; singular FSRs fall back to a flat source (zeros) and
 * are counted (src/TrackTraversingAlgorithms.cpp:745-831, MIN_DET = 1e-10, src/constants.h:70) */
__global__ void ls_invert_kernel(const double* __restrict__ lem, const double* __restrict__ volume, double* __restrict__ ilem,
                                 int64_t n_fsr, int solve_3d, int* __restrict__ n_flat) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_fsr; r += (int64_t)gridDim.x * blockDim.x) {
    if (solve_3d) {
      const double* m = lem + r * 6;
      double* o = ilem + r * 6;
      const double det = m[0] * m[1] * m[5] + m[2] * m[4] * m[3] + m[3] * m[2] * m[4]
                       - m[0] * m[4] * m[4] - m[3] * m[1] * m[3] - m[2] * m[2] * m[5];
      if (fabs(det) < 1e-10 || volume[r] < 1e-6) {
        for (int i = 0; i < 6; i++) o[i] = 0.0;
        atomicAdd(n_flat, 1);
      } else {
        o[0] = (m[1] * m[5] - m[4] * m[4]) / det; o[1] = (m[0] * m[5] - m[3] * m[3]) / det;
        o[2] = (m[3] * m[4] - m[2] * m[5]) / det; o[3] = (m[2] * m[4] - m[3] * m[1]) / det;
        o[4] = (m[3] * m[2] - m[0] * m[4]) / det; o[5] = (m[0] * m[1] - m[2] * m[2]) / det;
      }
    } else {
      const double* m = lem + r * 3;
      double* o = ilem + r * 3;
      const double det = m[0] * m[1] - m[2] * m[2];
      if (fabs(det) < 1e-10) {
        o[0] = o[1] = o[2] = 0.0;
        atomicAdd(n_flat, 1);
      } else {
        o[0] = m[1] / det; o[1] = m[0] / det; o[2] = -m[2] / det;
      }
    }
  }
}

}  // namespace b200
