/*
 * fsr_kernels.cuh - the per-FSR steps around the sweep as fused elementwise /
 * reduction kernels.  Each kernel cites the CPUSolver method it replaces
 * (src/CPUSolver.cpp) and the reference GPUSolver kernel it supersedes
 * (src/accel/cuda/GPUSolver.cu).
 *
 * Reductions are two-pass and order-fixed (per-block partials, then one block
 * folds them in index order) so results do not depend on block scheduling.
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

constexpr double FOUR_PI = 12.566370614359172;          /* src/constants.h:27 */
constexpr double ONE_OVER_FOUR_PI = 0.07957747154594767; /* src/constants.h:30 */
constexpr double FLUX_EPSILON = 1.0e-25;                 /* src/constants.h:15 */
constexpr double VOL_EPSILON = 1.0e-12;                  /* reference FLT_EPSILON, constants.h:12 */

constexpr int RED_THREADS = 256;
constexpr int MAX_PARTIALS = 1024;

/* device scalar block (double) */
enum { SC_KEFF = 0, SC_RATE, SC_NORM, SC_RESIDUAL, SC_KPREV, SC_TOL, SC_FXSCALE, SC_FXBOUND, SC_PSIMAX, SC_COUNT_D };
/* device scalar block (int) */
enum { SI_DONE = 0, SI_ITERS, SI_EXEC, SI_NEG_SRC, SI_NEG_FLUX, SI_COUNT_I };

struct FsrArgs {
  int G;
  int64_t n_fsr;
  int64_t n_fsr_global;
  int64_t n_fissionable;
  const int32_t* __restrict__ fsr_mat;
  const double* __restrict__ vol;
  const double* __restrict__ sigma_t;     /* [mat][G] */
  const double* __restrict__ sigma_s;     /* [mat][dest*G+orig] */
  const double* __restrict__ fiss;        /* [mat][dest*G+orig] */
  const double* __restrict__ nu_sigma_f;  /* [mat][G] */
  const double* __restrict__ sigma_f;
  const double* __restrict__ sigma_a;     /* [mat][G], Material::getSigmaA */
  const double* __restrict__ chi;
  const uint8_t* __restrict__ fissionable;
  double* __restrict__ phi;
  double* __restrict__ phi_old;
  double2* __restrict__ qst;              /* {q, sigma_t} per (fsr, group) */
  double* __restrict__ fixed;             /* may be NULL */
  double* __restrict__ stab;              /* may be NULL */
  double* __restrict__ scal;              /* SC_* */
  int* __restrict__ iscal;                /* SI_* */
  double* __restrict__ partials;          /* MAX_PARTIALS */
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

/* all threads of the block must call; result valid in thread 0 */
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[RED_THREADS / 32];
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.;
  if (w == 0) {
    r = (l < RED_THREADS / 32) ? sh[l] : 0.;
    r = warp_sum(r);
  }
  return r;
}

__device__ __forceinline__ double fold_partials(const double* partials, int n) {
  /* one block, fixed order */
  double v = 0.;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
  return block_sum(v);
}

/* ---- computeFSRSources (src/CPUSolver.cpp:1939-2023; GPUSolver.cu:288-338) ----
 * one thread per (fsr, destination group); the FSR's G fluxes are re-read from
 * L1 by its G threads.  mode 0: total, 1: fission only, 2: scatter only
 * (computeFSRFissionSources :2030, computeFSRScatterSources :2072). */
__global__ void __launch_bounds__(256)
sources_kernel(const FsrArgs a, int iteration, int mode, int neg_allowed) {
  if (a.iscal[SI_DONE]) return;
  if (iteration < 0) iteration = a.iscal[SI_EXEC];     /* CUDA-graph replay: device-side counter */
  const int G = a.G;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n_fsr * G) return;
  const int64_t r = idx / G;
  const int Gd = (int)(idx - r * G);
  const int m = a.fsr_mat[r];
  const double* __restrict__ ss = a.sigma_s + ((int64_t)m * G + Gd) * G;
  const double* __restrict__ fm = a.fiss + ((int64_t)m * G + Gd) * G;
  const double* __restrict__ ph = a.phi + r * G;
  const bool fissionable = a.fissionable[m];
  double scatter = 0., fission = 0.;
  for (int g = 0; g < G; g++) {
    const double f = ph[g];
    scatter = fma(ss[g], f, scatter);
    if (fissionable) fission = fma(fm[g], f, fission);
  }
  double q;
  if (mode == 0) {
    q = fission / a.scal[SC_KEFF];
    q += scatter;
    if (a.fixed != nullptr) q += a.fixed[idx];
  } else if (mode == 1) {
    q = fission;
  } else {
    q = scatter;
  }
  q *= ONE_OVER_FOUR_PI;
  if (mode == 0 && q < 0.0) {
    atomicAdd(&a.iscal[SI_NEG_SRC], 1);
    if (iteration < 30 && !neg_allowed) q = FLUX_EPSILON;
  }
  a.qst[idx].x = q;
}

/* sigma_t half of the {q, sigma_t} table, once per material upload */
__global__ void fill_sigma_t_kernel(const FsrArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n_fsr * a.G) return;
  const int64_t r = idx / a.G;
  const int e = (int)(idx - r * a.G);
  a.qst[idx].y = a.sigma_t[(int64_t)a.fsr_mat[r] * a.G + e];
}

/* ---- addSourceToScalarFlux (src/CPUSolver.cpp:2608-2659; GPUSolver.cu:666-695)
 *      fused with the nu-fission rate partial sums that computeKeff (:2258) and
 *      normalizeFluxes (:1860) both need ---- */
__global__ void __launch_bounds__(RED_THREADS)
closure_kernel(const FsrArgs a, int neg_allowed, int with_rate) {
  if (a.iscal[SI_DONE]) return;
  const int G = a.G;
  const int64_t n = a.n_fsr * G;
  double local = 0.;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / G;
    const int e = (int)(idx - r * G);
    double volume = a.vol[r];
    if (volume < VOL_EPSILON) volume = 1e30;
    const double2 qs = a.qst[idx];
    double f = a.phi[idx];
    f /= (qs.y * volume);
    f += FOUR_PI * qs.x / qs.y;
    if (f < 0.0 && !neg_allowed) {
      f = FLUX_EPSILON;
      atomicAdd(&a.iscal[SI_NEG_FLUX], 1);
    }
    a.phi[idx] = f;
    local += a.nu_sigma_f[(int64_t)a.fsr_mat[r] * G + e] * f * a.vol[r];
  }
  if (with_rate) {
    const double s = block_sum(local);
    if (threadIdx.x == 0) a.partials[blockIdx.x] = s;
  }
}

/* nu-fission (which=0) / fission (1) / absorption-free generic rate partials:
 * sum_r V_r sum_e sigma[e] phi[r,e]   (computeKeff :2279-2297, normalizeFluxes :1871-1888) */
__global__ void __launch_bounds__(RED_THREADS)
rate_partials_kernel(const FsrArgs a) {
  if (a.iscal[SI_DONE]) return;
  const int G = a.G;
  const int64_t n = a.n_fsr * G;
  double local = 0.;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / G;
    const int e = (int)(idx - r * G);
    local += a.nu_sigma_f[(int64_t)a.fsr_mat[r] * G + e] * a.phi[idx] * a.vol[r];
  }
  const double s = block_sum(local);
  if (threadIdx.x == 0) a.partials[blockIdx.x] = s;
}

/* ---- k_eff from the neutron balance (src/CPUSolver.cpp:2264-2325, Solver::setKeffFromNeutronBalance):
 *      k = fission / (absorption + leakage); three fixed-order reductions ---- */
__global__ void __launch_bounds__(RED_THREADS)
balance_partials_kernel(const FsrArgs a, const float* __restrict__ leakage, int64_t n_trk, double* __restrict__ out3) {
  /* grid = 3 x nb blocks: blockIdx.y selects fission / absorption / leakage */
  if (a.iscal[SI_DONE]) return;
  const int G = a.G;
  double local = 0.;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (blockIdx.y < 2) {
    const double* __restrict__ sig = blockIdx.y == 0 ? a.nu_sigma_f : a.sigma_a;
    const int64_t n = a.n_fsr * G;
    for (int64_t idx = tid; idx < n; idx += nth) {
      const int64_t r = idx / G;
      local += sig[(int64_t)a.fsr_mat[r] * G + (idx - r * G)] * a.phi[idx] * a.vol[r];
    }
  } else {
    for (int64_t t = tid; t < n_trk; t += nth) local += (double)leakage[t];
  }
  const double s = block_sum(local);
  if (threadIdx.x == 0) out3[blockIdx.y * gridDim.x + blockIdx.x] = s;
}
__global__ void __launch_bounds__(RED_THREADS)
balance_finalize_kernel(const FsrArgs a, const double* __restrict__ part3, int nb) {
  if (a.iscal[SI_DONE]) return;
  __shared__ double r[3];
  for (int k = 0; k < 3; k++) {
    const double v = fold_partials(part3 + k * nb, nb);
    if (threadIdx.x == 0) r[k] = v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a.scal[SC_KPREV] = a.scal[SC_KEFF];
    a.scal[SC_KEFF] = r[0] / (r[1] + r[2]);
    a.scal[SC_RATE] = r[0];
  }
}
__global__ void zero_float_kernel(float* __restrict__ p, int64_t n, const int* __restrict__ iscal) {
  if (iscal[SI_DONE]) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

/* fold the rate partials; op 0: rate only, 1: k *= rate/N (computeKeff :2327),
 * 2: norm = N/rate (normalizeFluxes :1910), 3: both (fused iteration) */
__global__ void __launch_bounds__(RED_THREADS)
rate_finalize_kernel(const FsrArgs a, int n_partials, int op) {
  if (a.iscal[SI_DONE]) return;
  const double rate = fold_partials(a.partials, n_partials);
  if (threadIdx.x == 0) {
    a.scal[SC_RATE] = rate;
    if (op & 1) {
      a.scal[SC_KPREV] = a.scal[SC_KEFF];
      a.scal[SC_KEFF] *= rate / (double)a.n_fsr_global;
    }
    if (op & 2) {
      a.scal[SC_NORM] = (double)a.n_fsr_global / rate;
      /* every normalisation factor is applied to the track fluxes exactly once (scale_psi_kernel):
       * the bound on |psi| that the deterministic tally keeps follows it */
      a.scal[SC_PSIMAX] *= fabs(a.scal[SC_NORM]);
    }
  }
}

/* phi *= norm (normalizeFluxes :1915-1919) */
__global__ void scale_phi_kernel(const FsrArgs a) {
  if (a.iscal[SI_DONE]) return;
  const double norm = a.scal[SC_NORM];
  const int64_t n = a.n_fsr * a.G;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x)
    a.phi[idx] *= norm;
}

/* psi *= norm, float *= double rounded to float (normalizeFluxes :1922-1928) */
__global__ void scale_psi_kernel(float* __restrict__ psi, int64_t n, const double* __restrict__ scal,
                                 const int* __restrict__ iscal) {
  if (iscal[SI_DONE]) return;
  const double norm = scal[SC_NORM];
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x)
    psi[idx] = (float)((double)psi[idx] * norm);
}

/* ---- computeResidual (src/CPUSolver.cpp:2113-2252; GPUSolver.cu:1728-1951's
 *      ~25 Thrust calls) as one pass: one thread per FSR.  Options fuse
 *      normalizeFluxes' phi scaling before and storeFSRFluxes (:1846) after. ---- */
__global__ void __launch_bounds__(RED_THREADS)
residual_kernel(const FsrArgs a, int res_type, int scale_first, int store_after) {
  if (a.iscal[SI_DONE]) return;
  const int G = a.G;
  const double norm = scale_first ? a.scal[SC_NORM] : 1.0;
  const double inv_k = 1.0 / a.scal[SC_KEFF];
  double local = 0.;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_fsr;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int m = a.fsr_mat[r];
    double* __restrict__ ph = a.phi + r * G;
    double* __restrict__ po = a.phi_old + r * G;
    if (scale_first)
      for (int e = 0; e < G; e++) ph[e] *= norm;
    double res = 0.;
    if (res_type == 0) {
      for (int e = 0; e < G; e++)
        if (po[e] > 0.) {
          const double d = (ph[e] - po[e]) / po[e];
          res += d * d;
        }
    } else if (res_type == 1) {
      if (a.fissionable[m]) {
        const double* __restrict__ nsf = a.nu_sigma_f + (int64_t)m * G;
        double nw = 0., od = 0.;
        for (int e = 0; e < G; e++) { nw += ph[e] * nsf[e]; od += po[e] * nsf[e]; }
        if (od > 0.) { const double d = (nw - od) / od; res = d * d; }
      }
    } else {
      double nw = 0., od = 0.;
      if (a.fissionable[m]) {
        const double* __restrict__ nsf = a.nu_sigma_f + (int64_t)m * G;
        for (int e = 0; e < G; e++) { nw += ph[e] * nsf[e]; od += po[e] * nsf[e]; }
        nw *= inv_k; od *= inv_k;
      }
      const double* __restrict__ ss = a.sigma_s + (int64_t)m * G * G;
      for (int Gd = 0; Gd < G; Gd++)
        for (int g = 0; g < G; g++) {
          nw += ss[Gd * G + g] * ph[g];
          od += ss[Gd * G + g] * po[g];
        }
      if (od > 0.) { const double d = (nw - od) / od; res = d * d; }
    }
    local += res;
    if (store_after)
      for (int e = 0; e < G; e++) po[e] = ph[e];
  }
  const double s = block_sum(local);
  if (threadIdx.x == 0) a.partials[blockIdx.x] = s;
}

/* fold residual partials, RMS, and (fused loops) the stopping rule of
 * Solver::computeEigenvalue (src/Solver.cpp:1643,1680): residual < tol and
 * integer-truncated |delta-k| pcm < 1.  loop_kind 0: none, 1: eigenvalue,
 * 2: flux/source loops (stop when i > 1 && residual < tol, :1409,:1506). */
__global__ void __launch_bounds__(RED_THREADS)
residual_finalize_kernel(const FsrArgs a, int n_partials, int res_type, int loop_kind, int iteration,
                         double* __restrict__ hist_k, double* __restrict__ hist_res) {
  if (a.iscal[SI_DONE]) return;
  double residual = fold_partials(a.partials, n_partials);
  if (threadIdx.x == 0) {
    if (iteration < 0) iteration = a.iscal[SI_EXEC];   /* CUDA-graph replay: device-side counter */
    int64_t norm = (res_type == 1) ? a.n_fissionable : a.n_fsr_global;
    if (residual < 0.0) residual = 0.0;
    if (norm <= 0) norm = 1;
    residual = sqrt(residual / (double)norm);
    a.scal[SC_RESIDUAL] = residual;
    if (loop_kind != 0) {
      a.iscal[SI_ITERS] = iteration + 1;
      a.iscal[SI_EXEC] = iteration + 1;   /* iterations that really ran on the device */
      if (hist_k != nullptr) { hist_k[iteration] = a.scal[SC_KEFF]; hist_res[iteration] = residual; }
      if (loop_kind == 1) {
        const int dk = (int)(1e5 * (a.scal[SC_KEFF] - a.scal[SC_KPREV]));
        if (residual < a.scal[SC_TOL] && abs(dk) < 1) a.iscal[SI_DONE] = 1;
      } else {
        if (iteration > 1 && residual < a.scal[SC_TOL]) { a.iscal[SI_DONE] = 1; a.iscal[SI_ITERS] = iteration; }
      }
    }
  }
}

/* storeFSRFluxes / flattenFSRFluxes / flattenFSRFluxesChiSpectrum (:1846,:1816,:1829) */
__global__ void copy_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n,
                            const int* __restrict__ iscal) {
  if (iscal != nullptr && iscal[SI_DONE]) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void zero_phi_kernel(double* __restrict__ dst, int64_t n, const int* __restrict__ iscal) {
  if (iscal[SI_DONE]) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) dst[i] = 0.0;
}
/* copy that only runs once the device-side loop has converged (see b200_iteration_begin) */
__global__ void copy_if_done_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n,
                                    const int* __restrict__ iscal) {
  if (!iscal[SI_DONE]) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void fill_kernel(double* __restrict__ dst, double v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void fill_chi_kernel(const FsrArgs a, int material) {
  const int64_t n = a.n_fsr * a.G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) a.phi[i] = a.chi[(int64_t)material * a.G + (i % a.G)];
}

/* ---- computeStabilizingFlux / stabilizeFlux (src/CPUSolver.cpp:2665-2805;
 *      the reference GPU solver only has GLOBAL, GPUSolver.cu:189-229) ----
 * max_ratio[e] (YAMAMOTO) is a material-table quantity: computed on the host. */
__global__ void stabilizing_flux_kernel(const FsrArgs a, int type, double factor,
                                        const double* __restrict__ max_ratio) {
  if (a.iscal[SI_DONE]) return;      /* a converged device-side loop leaves the flux untouched */
  const int G = a.G;
  const int64_t n = a.n_fsr * G;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / G;
    const int e = (int)(idx - r * G);
    const int m = a.fsr_mat[r];
    if (type == 0) {
      const double ss = a.sigma_s[((int64_t)m * G + e) * G + e];
      if (ss < 0.0) a.stab[idx] = -a.phi[idx] * factor * ss / a.sigma_t[(int64_t)m * G + e];
    } else if (type == 1) {
      a.stab[idx] = a.phi[idx] * (max_ratio[e] * factor);
    } else {
      a.stab[idx] = (1.0 / factor - 1.0) * a.phi[idx];
    }
  }
}
/* The same two steps for the three flux-moment planes of the linear source
 * (CPULSSolver::computeStabilizingFlux / stabilizeFlux, src/CPULSSolver.cpp:888-1052).  YAMAMOTO: the
 * reference's search for the largest scattering ratio never updates its maximum (`ratio = max_ratio`,
 * :943, :1028), so its moment stabilisation adds nothing and divides by one: reproduced as a no-op. */
__global__ void stabilizing_moments_kernel(const FsrArgs a, const double* __restrict__ phi_m, double* __restrict__ stab_m,
                                           int type, double factor) {
  if (a.iscal[SI_DONE]) return;
  const int G = a.G;
  const int64_t n = a.n_fsr * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = i % n, r = idx / G;
    const int e = (int)(idx - r * G);
    const int m = a.fsr_mat[r];
    if (type == 0) {
      const double ss = a.sigma_s[((int64_t)m * G + e) * G + e];
      if (ss < 0.0) stab_m[i] = -phi_m[i] * factor * ss / a.sigma_t[(int64_t)m * G + e];
    } else if (type == 1) {
      stab_m[i] = phi_m[i] * 0.0;
    } else {
      stab_m[i] = phi_m[i] * (1.0 / factor - 1.0);
    }
  }
}
__global__ void stabilize_moments_kernel(const FsrArgs a, double* __restrict__ phi_m, const double* __restrict__ stab_m,
                                         int type, double factor) {
  if (a.iscal[SI_DONE]) return;
  const int G = a.G;
  const int64_t n = a.n_fsr * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = i % n, r = idx / G;
    const int e = (int)(idx - r * G);
    const int m = a.fsr_mat[r];
    if (type == 0) {
      const double ss = a.sigma_s[((int64_t)m * G + e) * G + e];
      if (ss < 0.0) phi_m[i] = (phi_m[i] + stab_m[i]) / (1.0 - factor * ss / a.sigma_t[(int64_t)m * G + e]);
    } else if (type == 1) {
      phi_m[i] = (phi_m[i] + stab_m[i]) / (1 + 0.0);
    } else {
      phi_m[i] = (phi_m[i] + stab_m[i]) * factor;
    }
  }
}
__global__ void stabilize_flux_kernel(const FsrArgs a, int type, double factor,
                                      const double* __restrict__ max_ratio) {
  if (a.iscal[SI_DONE]) return;
  const int G = a.G;
  const int64_t n = a.n_fsr * G;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / G;
    const int e = (int)(idx - r * G);
    const int m = a.fsr_mat[r];
    if (type == 0) {
      const double ss = a.sigma_s[((int64_t)m * G + e) * G + e];
      if (ss < 0.0) {
        double f = a.phi[idx] + a.stab[idx];
        f /= (1.0 - factor * ss / a.sigma_t[(int64_t)m * G + e]);
        a.phi[idx] = f;
      }
    } else if (type == 1) {
      double f = a.phi[idx] + a.stab[idx];
      f /= (1 + max_ratio[e] * factor);
      a.phi[idx] = f;
    } else {
      a.phi[idx] = (a.phi[idx] + a.stab[idx]) * factor;
    }
  }
}

/* computeFSRFissionRates (src/CPUSolver.cpp:2825-2856; GPUSolver.cu:711-762) */
__global__ void fission_rates_kernel(const FsrArgs a, double* __restrict__ out, int nu) {
  const int G = a.G;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_fsr;
       r += (int64_t)gridDim.x * blockDim.x) {
    const double* __restrict__ sg = (nu ? a.nu_sigma_f : a.sigma_f) + (int64_t)a.fsr_mat[r] * G;
    double v = 0.;
    for (int e = 0; e < G; e++) v += sg[e] * a.phi[r * G + e] * a.vol[r];
    out[r] = v;
  }
}

/* ---- deterministic (fixed-point) tally support --------------------------------------
 * |sum w dpsi| <= 4 pi Sigma_t V M for every (FSR, group), with M the largest angular flux
 * that can occur: max(bound on |psi_in|, |q| / Sigma_t) (psi along a track is a convex combination
 * of its start value and the local q / Sigma_t).  A power-of-two scale with 2^5 head room
 * under 2^62 therefore cannot overflow; resolution is ~1e-17 of the bound. */
__global__ void __launch_bounds__(RED_THREADS)
fx_bound_kernel(const FsrArgs a, unsigned long long* __restrict__ bits) {
  /* bits[1]: max |q|/sigma_t, bits[2]: max sigma_t*V  (non-negative doubles order like integers).
   * Both are replicated quantities: every rank of a multi-GPU run derives the same scale. */
  double m_q = 0., m_sv = 0.;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t n = a.n_fsr * a.G;
  for (int64_t i = tid; i < n; i += nth) {
    const double2 qs = a.qst[i];
    m_q = fmax(m_q, fabs(qs.x) / qs.y);
    m_sv = fmax(m_sv, qs.y * a.vol[i / a.G]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m_q = fmax(m_q, __shfl_xor_sync(0xffffffffu, m_q, o));
    m_sv = fmax(m_sv, __shfl_xor_sync(0xffffffffu, m_sv, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&bits[1], (unsigned long long)__double_as_longlong(m_q));
    atomicMax(&bits[2], (unsigned long long)__double_as_longlong(m_sv));
  }
}
/* The bound on |psi| is kept as a recursion instead of a scan of this GPU's track fluxes (which
 * differ from rank to rank in a multi-GPU run and would give every rank its own scale): no flux
 * leaving this sweep exceeds max(bound on the incoming ones, max |q|/sigma_t); normalisations
 * rescale it (rate_finalize_kernel), zeroTrackFluxes resets it. */
__global__ void fx_scale_kernel(const FsrArgs a, unsigned long long* __restrict__ bits) {
  if (a.iscal[SI_DONE]) return;
  const double m = fmax(a.scal[SC_PSIMAX], __longlong_as_double((long long)bits[1]));
  a.scal[SC_PSIMAX] = m;
  double bound = FOUR_PI * __longlong_as_double((long long)bits[2]) * m;
  if (!(bound > 0.)) bound = 1.0;
  int ex;
  frexp(bound, &ex);                       /* bound < 2^ex */
  a.scal[SC_FXBOUND] = bound;
  a.scal[SC_FXSCALE] = ldexp(1.0, 62 - 5 - ex);
  bits[0] = bits[1] = bits[2] = 0ull;
}
/* GP: row pitch of the fixed-point tally (padded rows, sweep.cuh) */
__global__ void fx_to_double_kernel(const FsrArgs a, unsigned long long* __restrict__ fx, int GP) {
  if (a.iscal[SI_DONE]) return;
  const double inv = 1.0 / a.scal[SC_FXSCALE];
  const int64_t n = a.n_fsr * a.G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / a.G;
    const int64_t j = r * GP + (i - r * a.G);
    a.phi[i] = (double)(long long)fx[j] * inv;
    fx[j] = 0ull;
  }
}

__global__ void extract_q_kernel(const double2* __restrict__ qst, double* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) out[i] = qst[i].x;
}
__global__ void insert_q_kernel(double2* __restrict__ qst, const double* __restrict__ in, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) qst[i].x = in[i];
}

}  // namespace b200

/* ===================================================================================== */
/* Linear source (CPULSSolver): moment sources, closure and scaling                      */
/* ===================================================================================== */
namespace b200 {

struct LsArgs {
  int nc;                                   /* 3 in 2D, 6 in 3D */
  int solve_3d;
  const double* __restrict__ lin_exp;       /* [n_fsr][nc]  _FSR_lin_exp_matrix (CPULSSolver.h) */
  const double* __restrict__ src_const;     /* [n_fsr][nc][G]  _FSR_source_constants */
  double* __restrict__ phi_m;               /* [c*N_FSR*G + r*G+e] flux moments, one plane per component:
                                             * the sweep's moment REDs are then as coalesced as the phi RED */
  double4* __restrict__ qxyz;               /* {q_x, q_y, q_z, 0} per (r, e) */
  const double* __restrict__ fixed_m;       /* fixed source moments, planes like phi_m; NULL: none (CPULSSolver.cpp:475-479) */
};

/* CPULSSolver::computeFSRSources, moment part (src/CPULSSolver.cpp:386-524): per (FSR, group)
 * scatter + fission/k of the three flux moments, then the 2x2 / 3x3 linear expansion matrix. */
__global__ void __launch_bounds__(256)
sources_ls_kernel(const FsrArgs a, const LsArgs l, int iteration, int neg_allowed) {
  if (a.iscal[SI_DONE]) return;
  if (iteration < 0) iteration = a.iscal[SI_EXEC];
  const int G = a.G;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n_fsr * G) return;
  const int64_t r = idx / G;
  const int g = (int)(idx - r * G);
  const int m = a.fsr_mat[r];
  const double* __restrict__ ss = a.sigma_s + ((int64_t)m * G + g) * G;
  const double* __restrict__ fm = a.fiss + ((int64_t)m * G + g) * G;
  const bool fissionable = a.fissionable[m];
  double sx = 0., sy = 0., sz = 0., fx = 0., fy = 0., fz = 0.;
  for (int gp = 0; gp < G; gp++) {
    const int64_t np = a.n_fsr * G;
    const double* __restrict__ pm = l.phi_m + (r * G + gp);
    const double mx = pm[0], my = pm[np], mz = pm[2 * np];
    sx = fma(ss[gp], mx, sx); sy = fma(ss[gp], my, sy); sz = fma(ss[gp], mz, sz);
    if (fissionable) { fx = fma(fm[gp], mx, fx); fy = fma(fm[gp], my, fy); fz = fma(fm[gp], mz, fz); }
  }
  const double k = a.scal[SC_KEFF];
  double src_x = sx + fx / k, src_y = sy + fy / k, src_z = sz + fz / k;
  if (l.fixed_m != nullptr) {
    const int64_t np = a.n_fsr * G;
    src_x += l.fixed_m[idx]; src_y += l.fixed_m[np + idx]; src_z += l.fixed_m[2 * np + idx];
  }
  double4 q = make_double4(0., 0., 0., 0.);
  const double* __restrict__ M = l.lin_exp + r * l.nc;
  const double c = ONE_OVER_FOUR_PI / 2;
  if (neg_allowed || a.qst[idx].x > 10 * FLUX_EPSILON || iteration > 29) {
    if (l.solve_3d) {
      q.x = c * (M[0] * src_x + M[2] * src_y + M[3] * src_z);
      q.y = c * (M[2] * src_x + M[1] * src_y + M[4] * src_z);
      q.z = c * (M[3] * src_x + M[4] * src_y + M[5] * src_z);
    } else {
      q.x = c * (M[0] * src_x + M[2] * src_y);
      q.y = c * (M[2] * src_x + M[1] * src_y);
    }
  }
  l.qxyz[idx] = q;
}

/* CPULSSolver::addSourceToScalarFlux (src/CPULSSolver.cpp:787-882) + nu-fission partial sums */
__global__ void __launch_bounds__(RED_THREADS)
closure_ls_kernel(const FsrArgs a, const LsArgs l, int neg_allowed, int with_rate) {
  if (a.iscal[SI_DONE]) return;
  const int G = a.G;
  const int64_t n = a.n_fsr * G;
  double local = 0.;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / G;
    const int e = (int)(idx - r * G);
    double volume = a.vol[r];
    if (volume < VOL_EPSILON) volume = 1e30;
    const double2 qs = a.qst[idx];
    const double4 qm = l.qxyz[idx];
    const double* __restrict__ sc = l.src_const + r * G * l.nc;
    const double flux_const = FOUR_PI * 2;
    double f = a.phi[idx];
    f /= volume;
    f += FOUR_PI * qs.x;
    f /= qs.y;
    const int64_t np = a.n_fsr * G;
    double* __restrict__ pm = l.phi_m + idx;
    double mx = pm[0] / volume, my = pm[np] / volume, mz = pm[2 * np] / volume;
    mx += flux_const * qm.x * sc[e];
    mx += flux_const * qm.y * sc[2 * G + e];
    my += flux_const * qm.x * sc[2 * G + e];
    my += flux_const * qm.y * sc[G + e];
    if (l.solve_3d) {
      mx += flux_const * qm.z * sc[3 * G + e];
      my += flux_const * qm.z * sc[4 * G + e];
      mz += flux_const * qm.x * sc[3 * G + e];
      mz += flux_const * qm.y * sc[4 * G + e];
      mz += flux_const * qm.z * sc[5 * G + e];
    }
    mx /= qs.y; my /= qs.y;
    if (l.solve_3d) mz /= qs.y;
    if (f < 0.0 && !neg_allowed) {
      atomicAdd(&a.iscal[SI_NEG_FLUX], 1);
      f = fmax(a.phi_old[idx], FLUX_EPSILON);
      mx = my = mz = 0.;
    }
    a.phi[idx] = f;
    pm[0] = mx; pm[np] = my; pm[2 * np] = mz;
    local += a.nu_sigma_f[(int64_t)a.fsr_mat[r] * G + e] * f * a.vol[r];
  }
  if (with_rate) {
    const double s = block_sum(local);
    if (threadIdx.x == 0) a.partials[blockIdx.x] = s;
  }
}

/* CPULSSolver::normalizeFluxes, moment part (src/CPULSSolver.cpp:360-372) */
__global__ void scale_moments_kernel(const FsrArgs a, double* __restrict__ phi_m) {
  if (a.iscal[SI_DONE]) return;
  const double norm = a.scal[SC_NORM];
  const int64_t n = a.n_fsr * a.G * 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    phi_m[i] *= norm;
}

/* device [(r*G+e)*3 + c]  <->  reference layout [r*3G + c*G + e] (src/CPULSSolver.h:22-23) */
__global__ void moments_to_ref_kernel(const double* __restrict__ dev, double* __restrict__ ref, int64_t n_fsr, int G) {
  const int64_t n = n_fsr * G * 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (3 * G);
    const int rem = (int)(i - r * 3 * G);
    const int c = rem / G, e = rem - c * G;
    ref[i] = dev[(int64_t)c * n_fsr * G + r * G + e];
  }
}
__global__ void moments_from_ref_kernel(double* __restrict__ dev, const double* __restrict__ ref, int64_t n_fsr, int G) {
  const int64_t n = n_fsr * G * 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (3 * G);
    const int rem = (int)(i - r * 3 * G);
    const int c = rem / G, e = rem - c * G;
    dev[(int64_t)c * n_fsr * G + r * G + e] = ref[i];
  }
}

}  // namespace b200
