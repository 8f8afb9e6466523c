/*
 * cmfd.cuh - CMFD acceleration on the device: everything Cmfd::computeKeff
 * (src/Cmfd.cpp:1192-1295) does between two transport sweeps, on the arrays the sweep already
 * holds in HBM, so that neither the scalar flux nor the surface currents cross PCIe:
 *
 *   cmfd_split_kernel       Cmfd::splitVertexCurrents / splitEdgeCurrents   (Cmfd.cpp:2126-2331)
 *   cmfd_collapse_kernel    Cmfd::collapseXS                                 (Cmfd.cpp:720-1007)
 *   cmfd_diffusion_kernel   getDiffusionCoefficient + computeLarsensEDCFactor (Cmfd.cpp:1026, 1625)
 *   cmfd_matrix_kernel      constructMatrices + getSurfaceDiffusionCoefficient (Cmfd.cpp:1353-1500, 1048-1177)
 *   cmfd_eigen_kernel       eigenvalueSolve + linearSolve + rescaleFlux      (linalg.cpp:25-163, 179-395; Cmfd.cpp:1304)
 *   cmfd_update_kernel      updateMOCFlux + getUpdateRatio + getFluxRatio    (Cmfd.cpp:1509-1580, 3178-3290)
 *
 * The eigenvalue solve is ONE persistent kernel: the power iteration, the red/black SOR sweeps, the
 * matrix-vector products, the residual reductions and every convergence decision run on the device;
 * the only synchronisation is the barrier between the two colours (a CTA barrier with the flux in
 * shared memory for small meshes, a cooperative grid barrier otherwise).  The A and M matrices are
 * never assembled as CSR: A is a 7-point stencil per group plus a dense ncg x ncg block per cell,
 * M is block diagonal.  The operation order inside a row follows the column order of the
 * reference's CSR rows, and every reduction has a fixed order, so the solve is bitwise
 * reproducible (replicated on the GPUs of a group it yields the same bits on each).
 */
#pragma once
#include <cooperative_groups.h>
#include <cstdint>

#include "fsr_kernels.cuh"

namespace b200 {
namespace cg = cooperative_groups;

constexpr int CMFD_NS = 26;                 /* NUM_SURFACES, src/constants.h:119 */
constexpr int CMFD_NF = 6;                  /* NUM_FACES */
constexpr int CMFD_NFE = 18;                /* faces + edges */
constexpr int CMFD_BLOCK_THREADS = 512;     /* one-CTA solve: 128 registers per thread */
constexpr int CMFD_GRID_THREADS = 256;      /* cooperative solve */
#ifndef CMFD_GRID_MIN_BLOCKS
#define CMFD_GRID_MIN_BLOCKS 2              /* resident CTAs per SM the cooperative kernel is compiled for (register cap) */
#endif
constexpr double CMFD_EPS = 1.0e-12;        /* FLT_EPSILON of src/constants.h:12 (NOT the C one) */
constexpr double CMFD_FLUX_EPS = 1.0e-25;   /* FLUX_EPSILON, src/constants.h:15 */
constexpr double CMFD_ZERO_SIGMA_T = 1.0e-6;

enum { CS_KEFF = 0, CS_THRESH, CS_RES_1, CS_RES_END, CS_LIN_RES_1, CS_LIN_RES_END, CS_PF, CS_COUNT };
/* CI_PF_MAX / CI_PF_MIN are 64-bit words: keep their indices even */
enum { CI_FAIL = 0, CI_OLD_VALID, CI_POWER_ITERS, CI_LIN_ITERS_1, CI_LIN_ITERS_END, CI_LIN_TOTAL, CI_BAD_TALLY,
       CI_SOLVES, CI_PF_MAX, CI_PF_MAX_HI, CI_PF_MIN, CI_PF_MIN_HI, CI_COUNT };
static_assert(CI_PF_MAX % 2 == 0 && CI_PF_MIN % 2 == 0, "64-bit words of the CMFD integer block must be aligned");

struct CmfdArgs {
  int nx, ny, nz, ncg, G;
  int64_t n_cells, n_fsr;
  int bc[6];                               /* boundaryType per face: 0 VACUUM, 1 REFLECTIVE, 2 PERIODIC */
  int linear, flux_limiting, centroid, axial_interp, n_unbounded;
  double sor, relax, linalg_tol;
  int n_azim_2, n_polar_2;
  const double* __restrict__ wx; const double* __restrict__ wy; const double* __restrict__ wz;
  const int* __restrict__ group_idx;       /* ncg + 1 */
  const int* __restrict__ moc_to_cmfd;     /* G */
  const int64_t* __restrict__ cell_fsr_off; const int32_t* __restrict__ cell_fsrs;
  const int32_t* __restrict__ fsr_cell;
  const int32_t* __restrict__ nbr;         /* n_cells * 6, -1 outside */
  const double* __restrict__ azim_w; const double* __restrict__ sin_theta; const double* __restrict__ polar_w;
  /* MOC side */
  const int32_t* __restrict__ fsr_mat; const double* __restrict__ vol;
  const double* __restrict__ sigma_t; const double* __restrict__ sigma_s;
  const double* __restrict__ nu_sigma_f; const double* __restrict__ chi;
  double* phi; double* phi_m;
  double* cur;
  double* scal; int* iscal;
  /* split tables */
  const int64_t* __restrict__ sv_off; const int32_t* __restrict__ sv_src;
  const int64_t* __restrict__ se_off; const int32_t* __restrict__ se_src;
  /* CMFD state */
  double *rxn, *volc, *dift, *xs_t, *xs_nf, *xs_chi, *xs_s, *old_flux, *new_flux, *dcoef, *old_corr;
  double *diag, *off, *ain, *mm, *B, *SO, *SN, *partials;
  double *dq, *qd;                          /* diag / omega and omega / diag of every row (linalg.cpp:288, 318) */
  const int32_t* __restrict__ slot_cell;    /* [2][n_slots]: cell of every red / black slot, -1 for a hole */
  double* cs; int* ci;
  int x_in_smem;
  /* prolongation */
  const int64_t* __restrict__ st_off; const int32_t* __restrict__ st_cell; const double* __restrict__ st_w;
  const double* __restrict__ st_own; const int32_t* __restrict__ st_n;
  const double* __restrict__ ax_interp;
};

/* ------------------------------------------------------------------------------------------ */
/* Cmfd::splitVertexCurrents / splitEdgeCurrents as a gather: destination d = (cell, surface)  */
/* sums the currents of the vertices (edges) that split onto it; the sources are listed by the */
/* host from the reference's own rules (cmfd_impl.cuh: vertex_targets / edge_targets).         */
/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256)
cmfd_split_kernel(double* __restrict__ cur, const int64_t* __restrict__ off, const int32_t* __restrict__ src,
                  int64_t n_dest, int dest_per_cell, int ncg, double parts, const int* __restrict__ iscal) {
  if (iscal[SI_DONE]) return;
  const int64_t n = n_dest * ncg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = i / ncg;
    const int g = (int)(i - d * ncg);
    const int64_t a = off[d], b = off[d + 1];
    if (a == b) continue;
    const int64_t cell = d / dest_per_cell;
    const int surf = (int)(d - cell * dest_per_cell);
    double acc = cur[(cell * CMFD_NS + surf) * ncg + g];
    for (int64_t j = a; j < b; j++) acc += cur[(int64_t)src[j] * ncg + g] / parts;
    cur[(cell * CMFD_NS + surf) * ncg + g] = acc;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Cmfd::collapseXS: one thread per CMFD cell, the reference's loop order                      */
/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(128)
cmfd_collapse_kernel(CmfdArgs a) {
  if (a.iscal[SI_DONE]) return;
  const int G = a.G, ncg = a.ncg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_cells; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f0 = a.cell_fsr_off[i], f1 = a.cell_fsr_off[i + 1];
    double* chi_t = a.xs_chi + i * ncg;
    double* scat_t = a.xs_s + i * ncg * ncg;
    for (int e = 0; e < ncg; e++) chi_t[e] = 0.;
    double production = 0.;
    for (int64_t j = f0; j < f1; j++) {
      const int64_t r = a.cell_fsrs[j];
      const int m = a.fsr_mat[r];
      const double volume = a.vol[r];
      double np = 0.;
      for (int h = 0; h < G; h++) np += a.nu_sigma_f[m * G + h] * a.phi[r * G + h] * volume;
      for (int e = 0; e < ncg; e++) {
        double chi = 0.;
        for (int h = a.group_idx[e]; h < a.group_idx[e + 1]; h++) chi += a.chi[m * G + h];
        chi_t[e] += chi * np;
      }
      production += np;
    }
    for (int e = 0; e < ncg; e++) chi_t[e] = fabs(production) > 0. ? chi_t[e] / production : 0.;

    double vol_tally = 0.;
    for (int e = 0; e < ncg; e++) {
      double nu_fission = 0., total = 0., reaction = 0., diffusion = 0.;
      double* sc = scat_t + e * ncg;            /* origin e -> destination g */
      for (int g = 0; g < ncg; g++) sc[g] = 0.;
      for (int h = a.group_idx[e]; h < a.group_idx[e + 1]; h++) {
        vol_tally = 0.;
        double rxn_group = 0., trans_group = 0.;
        for (int64_t j = f0; j < f1; j++) {
          const int64_t r = a.cell_fsrs[j];
          const int m = a.fsr_mat[r];
          const double volume = a.vol[r];
          const double flux = a.phi[r * G + h];
          const double tot = a.sigma_t[m * G + h];
          const double nuf = a.nu_sigma_f[m * G + h];
          total += tot * flux * volume;
          nu_fission += nuf * flux * volume;
          reaction += flux * volume;
          vol_tally += volume;
          rxn_group += flux * volume;
          trans_group += tot * flux * volume;
          const double* srow = a.sigma_s + (int64_t)m * G * G;
          for (int g = 0; g < G; g++) sc[a.moc_to_cmfd[g]] += srow[g * G + h] * flux * volume;
        }
        if (fabs(trans_group) > fabs(rxn_group) * CMFD_EPS) {
          const double avg_sigma_t = trans_group / rxn_group;
          diffusion += rxn_group / (3.0 * avg_sigma_t);
        }
      }
      if (reaction <= 0.) {
        atomicAdd(&a.ci[CI_BAD_TALLY], 1);
        reaction = CMFD_ZERO_SIGMA_T;
        diffusion = CMFD_ZERO_SIGMA_T;
        total = CMFD_ZERO_SIGMA_T;
        if (nu_fission != 0.) nu_fission = CMFD_ZERO_SIGMA_T;
        for (int g = 0; g < ncg; g++) sc[g] = 0.;
      }
      a.rxn[i * ncg + e] = reaction;
      a.dift[i * ncg + e] = diffusion;
      a.xs_t[i * ncg + e] = total / reaction;
      a.xs_nf[i * ncg + e] = nu_fission / reaction;
      for (int g = 0; g < ncg; g++) sc[g] = sc[g] / reaction;
    }
    a.volc[i] = vol_tally;
    for (int e = 0; e < ncg; e++) a.old_flux[i * ncg + e] = a.rxn[i * ncg + e] / vol_tally;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* diffusion coefficient of every (cell, group) along x, y, z with Larsen's correction          */
/* ------------------------------------------------------------------------------------------ */
__device__ inline double cmfd_width(const CmfdArgs& a, int64_t cell, int axis) {
  const int ix = (int)(cell % a.nx), iy = (int)((cell % ((int64_t)a.nx * a.ny)) / a.nx), iz = (int)(cell / ((int64_t)a.nx * a.ny));
  return axis == 0 ? a.wx[ix] : (axis == 1 ? a.wy[iy] : a.wz[iz]);
}

__global__ void __launch_bounds__(256)
cmfd_diffusion_kernel(CmfdArgs a) {
  if (a.iscal[SI_DONE]) return;
  const int64_t n = a.n_cells * a.ncg * 3;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / 3;
    const int axis = (int)(t - row * 3);
    const int64_t cell = row / a.ncg;
    double d = a.dift[row] / a.rxn[row];
    if (!a.linear) {
      const double delta = cmfd_width(a, cell, axis);
      double rho = 0.;
      for (int az = 0; az < a.n_azim_2; az++) {
        const double wa = a.azim_w[az];
        for (int p = 0; p < a.n_polar_2; p++) {
          const double st = a.sin_theta[az * a.n_polar_2 + p];
          const double mu = sqrt(1.0 - st * st);
          const double expon = exp(-delta / (3 * d * mu));
          const double alpha = (1 + expon) / (1 - expon) - 2 * (3 * d * mu) / delta;
          rho += 2.0 * mu * a.polar_w[az * a.n_polar_2 + p] * wa * alpha;
        }
      }
      d *= 1.0 + delta * rho / (2 * d);
    }
    a.dcoef[t] = d;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Cmfd::constructMatrices: thread per (cell, group)                                           */
/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256)
cmfd_matrix_kernel(CmfdArgs a, int moc_iteration) {
  if (a.iscal[SI_DONE]) return;
  if (moc_iteration < 0) moc_iteration = a.iscal[SI_EXEC];
  const int ncg = a.ncg;
  const bool old_valid = a.ci[CI_OLD_VALID] != 0;
  const int64_t n = a.n_cells * ncg;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = row / ncg;
    const int e = (int)(row - i * ncg);
    const int ix = (int)(i % a.nx), iy = (int)((i % ((int64_t)a.nx * a.ny)) / a.nx), iz = (int)(i / ((int64_t)a.nx * a.ny));
    const double volume = a.volc[i];
    double diag = a.xs_t[row] * volume;
    for (int g = 0; g < ncg; g++) {
      double value = -a.xs_s[(i * ncg + g) * ncg + e] * volume;
      if (!(fabs(value) > CMFD_EPS)) value = 0.;
      if (g == e) { diag += value; value = 0.; }
      a.ain[row * ncg + g] = value;
    }
    const double flux = a.old_flux[row];
    for (int s = 0; s < CMFD_NF; s++) {
      const int axis = s % 3;
      const double delta_if = axis == 0 ? a.wy[iy] * a.wz[iz] : (axis == 1 ? a.wx[ix] * a.wz[iz] : a.wx[ix] * a.wy[iy]);
      const double delta = axis == 0 ? a.wx[ix] : (axis == 1 ? a.wy[iy] : a.wz[iz]);
      const double dif = a.dcoef[row * 3 + axis];
      const int64_t next = a.nbr[i * CMFD_NF + s];
      double ds = 0., dc = 0.;
      if (next < 0) {
        if (a.bc[s] == 0) {                        /* VACUUM */
          const double current_out = a.cur[(i * CMFD_NS + s) * ncg + e] / delta_if;
          ds = 2 * dif / delta / (1 + 4 * dif / delta);
          dc = (ds * flux - current_out) / flux;
        }
      } else {
        const int s_next = (s + CMFD_NF / 2) % CMFD_NF;
        const double current_out = a.cur[(i * CMFD_NS + s) * ncg + e];
        const double current_in = a.cur[(next * CMFD_NS + s_next) * ncg + e];
        const double delta_next = cmfd_width(a, next, axis);
        const double dif_next = a.dcoef[(next * ncg + e) * 3 + axis];
        const double flux_next = a.old_flux[next * ncg + e];
        ds = 2.0 * dif * dif_next / (delta_next * dif + delta * dif_next);
        const double current = (current_out - current_in) / delta_if;
        dc = -(ds * (flux_next - flux) + current) / (flux_next + flux);
        if (a.flux_limiting && moc_iteration > 0) {
          const double ratio = dc / ds;
          if (fabs(ratio) > 1.0) {
            ds = current > 0.0 ? fabs(current / (2.0 * flux)) : fabs(current / (2.0 * flux_next));
            dc = -(ds * (flux_next - flux) + current) / (flux_next + flux);
            ds = fmax(ds, fabs(dc));
          }
        }
      }
      if (old_valid) dc = a.relax * dc + (1.0 - a.relax) * a.old_corr[(i * CMFD_NF + s) * ncg + e];
      if (moc_iteration == 0) dc = 0.0;
      a.old_corr[(i * CMFD_NF + s) * ncg + e] = dc;
      diag += (ds - dc) * delta_if;
      a.off[row * CMFD_NF + s] = next >= 0 ? -(ds + dc) * delta_if : 0.;
    }
    a.diag[row] = diag;
    a.dq[row] = diag / a.sor;
    a.qd[row] = a.sor / diag;
    for (int g = 0; g < ncg; g++) {
      const double value = a.xs_chi[row] * a.xs_nf[i * ncg + g] * volume;
      a.mm[row * ncg + g] = fabs(value) > CMFD_EPS ? value : 0.;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* eigenvalueSolve + linearSolve + rescaleFlux in one persistent kernel                        */
/* MODE 0: one CTA, __syncthreads, flux in shared memory when it fits                          */
/* MODE 1: cooperative grid, grid barrier, flux in L2 (read with ld.global.cg)                 */
/* ------------------------------------------------------------------------------------------ */
template <int MODE>
__device__ __forceinline__ void cmfd_barrier() {
  if constexpr (MODE == 0) __syncthreads();
  else cg::this_grid().sync();
}

template <int MODE>
__device__ __forceinline__ double cmfd_ldx(const double* X, int64_t i) {
  if constexpr (MODE == 0) return X[i];
  else return __ldcg(X + i);
}

/* first half of a reduction: the CTA's sum goes to partials[slot][blockIdx] */
template <int MODE>
__device__ __forceinline__ void cmfd_reduce_put(double v, double* sh, double* partials, int slot) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.;
    const int nw = (blockDim.x + 31) >> 5;
    for (int j = 0; j < nw; j++) t += sh[j];
    if constexpr (MODE == 0) sh[32] = t;
    else partials[(int64_t)slot * gridDim.x + blockIdx.x] = t;
  }
}

/* second half, after a barrier */
template <int MODE>
__device__ __forceinline__ double cmfd_reduce_get(double* sh, const double* partials, int slot) {
  if constexpr (MODE == 0) {
    const double t = sh[32];
    __syncthreads();
    return t;
  } else {
    if (threadIdx.x < 32) {
      double t = 0.;
      for (int j = threadIdx.x; j < (int)gridDim.x; j += 32) t += __ldcg(partials + (int64_t)slot * gridDim.x + j);
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (threadIdx.x == 0) sh[33] = t;
    }
    __syncthreads();
    const double t = sh[33];
    __syncthreads();
    return t;
  }
}

template <int MODE>
__device__ __forceinline__ double cmfd_reduce(double v, double* sh, double* partials, int& slot) {
  cmfd_reduce_put<MODE>(v, sh, partials, slot);
  cmfd_barrier<MODE>();
  const double t = cmfd_reduce_get<MODE>(sh, partials, slot);
  slot ^= 1;
  return t;
}

/* cell of slot `idx` of colour `colour` (linalg.cpp:276-279), or -1 */
__device__ __forceinline__ int64_t cmfd_slot_cell(const CmfdArgs& a, int64_t idx, int colour, int hx) {
  const int64_t rowi = idx / hx;
  const int k = (int)(idx - rowi * hx);
  const int iy = (int)(rowi % a.ny), iz = (int)(rowi / a.ny);
  const int ix = 2 * k + ((iy + iz + colour) & 1);
  if (ix >= a.nx) return -1;
  return rowi * a.nx + ix;
}

/* (M X) of one cell into out[], returns the group sum */
template <int MODE>
__device__ __forceinline__ double cmfd_cell_source(const CmfdArgs& a, const double* X, int64_t cell, double* out) {
  const int ncg = a.ncg;
  double tot = 0.;
  for (int e = 0; e < ncg; e++) {
    const int64_t row = cell * ncg + e;
    double s = 0.;
    for (int g = 0; g < ncg; g++) s += a.mm[row * ncg + g] * X[cell * ncg + g];
    out[row] = s;
    tot += s;
  }
  return tot;
}

/* One red/black SOR update of a cell (linalg.cpp:281-320) followed by its new source (M X) and, when `need`, its
 * term of the residual.  The row is walked in the column order of the reference's CSR matrix: the lower
 * neighbours (z-, y-, x-), the groups of the cell itself (Gauss-Seidel inside the cell, the source at the
 * diagonal's position), the upper neighbours.  NCG > 0: group count known at compile time - the six neighbour
 * fluxes of every group are requested before any arithmetic, so a phase costs one L2 round trip. */
template <int MODE, int NCG>
__device__ __forceinline__ void cmfd_sor_cell(const CmfdArgs& a, double* X, int64_t cell, double omega, bool need,
                                              const double* old_src, double& part) {
  const int32_t* nbp = a.nbr + cell * CMFD_NF;
  if constexpr (NCG > 0) {
    int nbi[CMFD_NF];
#pragma unroll
    for (int s = 0; s < CMFD_NF; s++) nbi[s] = __ldg(nbp + s);
    /* everything this thread wrote itself (B, SO) and every flux is requested before the first store: the
     * compiler cannot move a plain load above a store through another double*, and a load issued right before
     * its use costs a full L2 round trip */
    double xn[CMFD_NF][NCG], xo[NCG], bv[NCG];
    double sold = 0.;
#pragma unroll
    for (int s = 0; s < CMFD_NF; s++)
#pragma unroll
      for (int g = 0; g < NCG; g++) xn[s][g] = nbi[s] >= 0 ? cmfd_ldx<MODE>(X, (int64_t)nbi[s] * NCG + g) : 0.;
#pragma unroll
    for (int g = 0; g < NCG; g++) { xo[g] = cmfd_ldx<MODE>(X, cell * NCG + g); bv[g] = a.B[cell * NCG + g]; }
    if (need) {
#pragma unroll
      for (int e = 0; e < NCG; e++) sold += old_src[cell * NCG + e];
    }
#pragma unroll
    for (int g = 0; g < NCG; g++) {
      const int64_t row = cell * NCG + g;
      const double* of = a.off + row * CMFD_NF;
      double v = (1.0 - omega) * xo[g] * __ldg(a.dq + row);
      v -= __ldg(of + 2) * xn[2][g];
      v -= __ldg(of + 1) * xn[1][g];
      v -= __ldg(of + 0) * xn[0][g];
#pragma unroll
      for (int g2 = 0; g2 < NCG; g2++) {
        if (g2 == g) v += bv[g];
        else v -= __ldg(a.ain + row * NCG + g2) * xo[g2];
      }
      v -= __ldg(of + 3) * xn[3][g];
      v -= __ldg(of + 4) * xn[4][g];
      v -= __ldg(of + 5) * xn[5][g];
      xo[g] = v * __ldg(a.qd + row);
    }
    double snew = 0., sn[NCG];
#pragma unroll
    for (int e = 0; e < NCG; e++) {
      const int64_t row = cell * NCG + e;
      double sv = 0.;
#pragma unroll
      for (int g = 0; g < NCG; g++) sv += __ldg(a.mm + row * NCG + g) * xo[g];
      sn[e] = sv;
      snew += sv;
    }
#pragma unroll
    for (int g = 0; g < NCG; g++) { X[cell * NCG + g] = xo[g]; a.SN[cell * NCG + g] = sn[g]; }
    if (need && fabs(sold) > CMFD_FLUX_EPS) { const double q = (snew - sold) / sold; part += q * q; }
  } else {
    const int ncg = a.ncg;
    int nbi[CMFD_NF];
    for (int s = 0; s < CMFD_NF; s++) nbi[s] = __ldg(nbp + s);
    double sold = 0.;
    if (need)
      for (int e = 0; e < ncg; e++) sold += old_src[cell * ncg + e];
    for (int g = 0; g < ncg; g++) {
      const int64_t row = cell * ncg + g;
      const double* of = a.off + row * CMFD_NF;
      double v = (1.0 - omega) * cmfd_ldx<MODE>(X, row) * __ldg(a.dq + row);
      if (nbi[2] >= 0) v -= __ldg(of + 2) * cmfd_ldx<MODE>(X, (int64_t)nbi[2] * ncg + g);
      if (nbi[1] >= 0) v -= __ldg(of + 1) * cmfd_ldx<MODE>(X, (int64_t)nbi[1] * ncg + g);
      if (nbi[0] >= 0) v -= __ldg(of + 0) * cmfd_ldx<MODE>(X, (int64_t)nbi[0] * ncg + g);
      for (int g2 = 0; g2 < ncg; g2++) {
        if (g2 == g) v += a.B[row];
        else v -= __ldg(a.ain + row * ncg + g2) * cmfd_ldx<MODE>(X, cell * ncg + g2);
      }
      if (nbi[3] >= 0) v -= __ldg(of + 3) * cmfd_ldx<MODE>(X, (int64_t)nbi[3] * ncg + g);
      if (nbi[4] >= 0) v -= __ldg(of + 4) * cmfd_ldx<MODE>(X, (int64_t)nbi[4] * ncg + g);
      if (nbi[5] >= 0) v -= __ldg(of + 5) * cmfd_ldx<MODE>(X, (int64_t)nbi[5] * ncg + g);
      X[row] = v * __ldg(a.qd + row);
    }
    double snew = 0.;
    for (int e = 0; e < ncg; e++) {
      const int64_t row = cell * ncg + e;
      double sv = 0.;
      for (int g = 0; g < ncg; g++) sv += __ldg(a.mm + row * ncg + g) * cmfd_ldx<MODE>(X, cell * ncg + g);
      a.SN[row] = sv;
      snew += sv;
    }
    if (need && fabs(sold) > CMFD_FLUX_EPS) { const double q = (snew - sold) / sold; part += q * q; }
  }
}

template <int MODE, int NCG>
__global__ void __launch_bounds__(MODE == 0 ? CMFD_BLOCK_THREADS : CMFD_GRID_THREADS, MODE == 0 ? 1 : CMFD_GRID_MIN_BLOCKS)
cmfd_eigen_kernel(CmfdArgs a, double source_thresh) {
  extern __shared__ double cmfd_smem[];
  __shared__ double sh[34];
  if (a.iscal[SI_DONE]) return;
  const int ncg = NCG > 0 ? NCG : a.ncg;
  const int64_t T = (int64_t)gridDim.x * blockDim.x, gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int hx = (a.nx + 1) / 2;
  const int64_t n_slots = (int64_t)a.nz * a.ny * hx;
  const int64_t nr = a.n_cells * ncg;
  double* X = (MODE == 0 && a.x_in_smem) ? cmfd_smem : a.new_flux;
  int slot = 0;
  if (gt == 0) a.ci[CI_OLD_VALID] = 1;          /* the matrices of this solve are built (Cmfd.cpp:1498) */
  if (source_thresh < 0.) source_thresh = a.cs[CS_THRESH];

  /* the collapsed flux is the starting guess (Cmfd.cpp:1232) */
  for (int64_t r = gt; r < nr; r += T) X[r] = a.old_flux[r];
  cmfd_barrier<MODE>();

  /* initial source, normalised to num_rows (linalg.cpp:65-80) */
  double local = 0.;
  for (int c2 = 0; c2 < 2; c2++)
    for (int64_t idx = gt; idx < n_slots; idx += T) {
      const int64_t cell = cmfd_slot_cell(a, idx, c2, hx);
      if (cell >= 0) local += cmfd_cell_source<MODE>(a, X, cell, a.B);
    }
  double sum = cmfd_reduce<MODE>(local, sh, a.partials, slot);
  double k = a.cs[CS_KEFF];
  {
    const double fb = (double)nr / sum, fx = (double)nr * k / sum;
    for (int c2 = 0; c2 < 2; c2++)
      for (int64_t idx = gt; idx < n_slots; idx += T) {
        const int64_t cell = cmfd_slot_cell(a, idx, c2, hx);
        if (cell < 0) continue;
        for (int e = 0; e < ncg; e++) { a.B[cell * ncg + e] *= fb; X[cell * ncg + e] *= fx; }
      }
  }
  cmfd_barrier<MODE>();

  const double lin_tol = fmax(a.linalg_tol, fmax(a.linalg_tol, source_thresh) * 1e-1);
  const double inv_cells = 1.0 / (double)a.n_cells;
  double initial_residual = 0., residual = 0.;
  int iter = 0, lin_total = 0, lin_iters = 0, lin_iters_1 = 0;
  double lin_res_first = 0., lin_res_1 = 0., res_1 = 0.;
  bool ok = true, converged = false;
  const double omega = a.sor;

  for (iter = 0; iter < 25000; iter++) {
    /* ---- linearSolve(A, M, X, B) ---- */
    for (int c2 = 0; c2 < 2; c2++)
      for (int64_t idx = gt; idx < n_slots; idx += T) {
        const int64_t cell = cmfd_slot_cell(a, idx, c2, hx);
        if (cell >= 0) cmfd_cell_source<MODE>(a, X, cell, a.SO);
      }
    double lres = 0., linit = 0., min_res = 1e6;
    int liter = 0;
    while (liter < 10000) {
      const bool need = liter == 0 || liter + 1 > 25;
      /* the source this sweep is compared with: the one before the solve until the 25th sweep, the previous
       * sweep's afterwards (linalg.cpp:383-386 copies new to old from then on; here the previous source is simply
       * read before it is overwritten) */
      const double* old_src = liter >= 25 ? a.SN : a.SO;
      double part = 0.;
      for (int colour = 0; colour < 2; colour++) {
        for (int64_t idx = gt; idx < n_slots; idx += T) {
          const int64_t cell = __ldg(a.slot_cell + colour * n_slots + idx);
          if (cell < 0) continue;
          cmfd_sor_cell<MODE, NCG>(a, X, cell, omega, need, old_src, part);
        }
        if (colour == 1 && need) cmfd_reduce_put<MODE>(part, sh, a.partials, slot);
        cmfd_barrier<MODE>();
      }
      double r = 0.;
      if (need) {
        const double t = cmfd_reduce_get<MODE>(sh, a.partials, slot);
        slot ^= 1;
        r = sqrt(fmax(t, 0.) * inv_cells);
      }
      if (liter == 0) { lres = r; linit = r; lin_res_first = r; }
      liter++;
      if (liter > 25) {
        lres = r;
        if (lres < min_res) min_res = lres;
        if ((lres > 1e3 * min_res && min_res > 1e-10) || !(lres == lres)) { ok = false; break; }
        if (lres / linit < 0.1 || lres < lin_tol) break;
      }
    }
    lin_total += liter;
    lin_iters = liter;
    if (liter >= 10000) ok = false;
    if (!ok) break;

    /* ---- new source, k, residual (linalg.cpp:104-125) ---- */
    local = 0.;
    for (int c2 = 0; c2 < 2; c2++)
      for (int64_t idx = gt; idx < n_slots; idx += T) {
        const int64_t cell = cmfd_slot_cell(a, idx, c2, hx);
        if (cell < 0) continue;
        for (int e = 0; e < ncg; e++) local += a.SN[cell * ncg + e];
      }
    sum = cmfd_reduce<MODE>(local, sh, a.partials, slot);
    k = sum / (double)nr;
    const double inv_k = 1.0 / k;
    local = 0.;
    for (int c2 = 0; c2 < 2; c2++)
      for (int64_t idx = gt; idx < n_slots; idx += T) {
        const int64_t cell = cmfd_slot_cell(a, idx, c2, hx);
        if (cell < 0) continue;
        double snew = 0., sold = 0.;
        for (int e = 0; e < ncg; e++) {
          const double v = a.SN[cell * ncg + e] * inv_k;
          snew += v;
          sold += a.B[cell * ncg + e];
          a.B[cell * ncg + e] = v;
        }
        if (fabs(sold) > CMFD_FLUX_EPS) { const double q = (snew - sold) / sold; local += q * q; }
      }
    residual = sqrt(fmax(cmfd_reduce<MODE>(local, sh, a.partials, slot), 0.) * inv_cells);
    if (iter == 0) {
      initial_residual = residual;
      if (initial_residual < 1e-14) initial_residual = 1e-10;
      res_1 = residual; lin_iters_1 = lin_iters; lin_res_1 = lin_res_first;
    }
    if ((residual / initial_residual < 0.03 || residual < a.linalg_tol) && iter > 25) { converged = true; break; }
  }
  if (!converged) ok = false;

  if (ok) {
    /* Cmfd::rescaleFlux (Cmfd.cpp:1304-1345): both fluxes to a unit total fission source */
    double ln = 0., lo = 0.;
    for (int c2 = 0; c2 < 2; c2++)
      for (int64_t idx = gt; idx < n_slots; idx += T) {
        const int64_t cell = cmfd_slot_cell(a, idx, c2, hx);
        if (cell < 0) continue;
        for (int e = 0; e < ncg; e++) {
          const int64_t row = cell * ncg + e;
          double sn = 0., so = 0.;
          for (int g = 0; g < ncg; g++) {
            sn += a.mm[row * ncg + g] * X[cell * ncg + g];
            so += a.mm[row * ncg + g] * a.old_flux[cell * ncg + g];
          }
          ln += sn; lo += so;
        }
      }
    const double sum_new = cmfd_reduce<MODE>(ln, sh, a.partials, slot);
    const double sum_old = cmfd_reduce<MODE>(lo, sh, a.partials, slot);
    const double fn = 1.0 / sum_new, fo = 1.0 / sum_old;
    for (int c2 = 0; c2 < 2; c2++)
      for (int64_t idx = gt; idx < n_slots; idx += T) {
        const int64_t cell = cmfd_slot_cell(a, idx, c2, hx);
        if (cell < 0) continue;
        for (int e = 0; e < ncg; e++) {
          a.new_flux[cell * ncg + e] = X[cell * ncg + e] * fn;
          a.old_flux[cell * ncg + e] *= fo;
        }
      }
  }
  if (gt == 0) {
    a.ci[CI_FAIL] = ok ? 0 : 1;
    a.ci[CI_POWER_ITERS] = iter;
    a.ci[CI_LIN_ITERS_1] = lin_iters_1;
    a.ci[CI_LIN_ITERS_END] = lin_iters;
    a.ci[CI_LIN_TOTAL] = lin_total;
    a.ci[CI_SOLVES] += 1;
    a.cs[CS_RES_1] = res_1; a.cs[CS_RES_END] = residual;
    a.cs[CS_LIN_RES_1] = lin_res_1; a.cs[CS_LIN_RES_END] = lin_res_first;
    if (ok) a.cs[CS_KEFF] = k;
    /* Solver.cpp:1628: _k_eff = _cmfd->computeKeff(); a failed solve hands back the last good CMFD k */
    a.scal[SC_KPREV] = a.scal[SC_KEFF];
    a.scal[SC_KEFF] = a.cs[CS_KEFF];
    /* prolongation-factor extremes, filled by cmfd_update_kernel */
    *reinterpret_cast<long long*>(&a.ci[CI_PF_MAX]) = 0;
    *reinterpret_cast<long long*>(&a.ci[CI_PF_MIN]) = 0;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* The same solve on ONE thread-block cluster: the flux lives in the shared memory of the CTAs  */
/* (CTA r owns the slots [r*S, (r+1)*S) of both colours), a neighbour's flux is a distributed-   */
/* shared-memory load (ld.shared::cluster), the two colours are separated by the hardware        */
/* cluster barrier, reductions go through per-CTA partials read over DSMEM in rank order.  With  */
/* ALL_SMEM the three source vectors sit in shared memory too and a colour phase touches HBM/L2  */
/* only for the (L1-resident, read-only) matrix coefficients.                                    */
/* ------------------------------------------------------------------------------------------ */
constexpr int CMFD_CLUSTER_THREADS = 512;

struct CmfdClusterArgs {
  int slots_per_cta;                       /* S */
  int all_smem;                            /* B, SO, SN in shared memory as well */
  const int32_t* __restrict__ nb_loc;      /* n_cells * 6: rank << 26 | (local slot * 2 + colour), -1 outside */
};

__device__ __forceinline__ int cmfd_slot_cell32(const CmfdArgs& a, int idx, int colour, int hx) {
  const int rowi = idx / hx, k = idx - rowi * hx;
  const int iy = rowi % a.ny, iz = rowi / a.ny;
  const int ix = 2 * k + ((iy + iz + colour) & 1);
  return ix >= a.nx ? -1 : rowi * a.nx + ix;
}

template <int NCG>
__global__ void __launch_bounds__(CMFD_CLUSTER_THREADS)
cmfd_eigen_cluster_kernel(CmfdArgs a, CmfdClusterArgs ca, double source_thresh) {
  extern __shared__ double cmfd_smem[];
  __shared__ double sh[40];                /* [0,32) warp sums, 32/33 result, 34..37 this CTA's partials (ping-pong x 2) */
  if (a.iscal[SI_DONE]) return;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), C = (int)cluster.num_blocks();
  const int S = ca.slots_per_cta, hx = (a.nx + 1) / 2;
  const int n_slots = a.nz * a.ny * hx;
  const int base = rank * S;
  const int mine = max(0, min(S, n_slots - base));
  const int64_t nr = a.n_cells * NCG;
  double* Xs = cmfd_smem;                                     /* [S*2][NCG] */
  double* Bv = ca.all_smem ? Xs + (size_t)S * 2 * NCG : a.B;
  double* SOv = ca.all_smem ? Bv + (size_t)S * 2 * NCG : a.SO;
  double* SNv = ca.all_smem ? SOv + (size_t)S * 2 * NCG : a.SN;
  const bool sm = ca.all_smem != 0;
  int slot = 0;
  if (rank == 0 && threadIdx.x == 0) a.ci[CI_OLD_VALID] = 1;
  if (source_thresh < 0.) source_thresh = a.cs[CS_THRESH];

  /* cluster-wide sum with a fixed order: CTA partial -> own shared memory -> every CTA adds the C partials */
  auto reduce_put = [&](double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.;
      for (int j = 0; j < (int)(blockDim.x >> 5); j++) t += sh[j];
      sh[34 + slot] = t;
    }
  };
  auto reduce_get = [&]() {
    if (threadIdx.x == 0) {
      double t = 0.;
      for (int r = 0; r < C; r++) t += cluster.map_shared_rank(sh, r)[34 + slot];
      sh[32] = t;
    }
    __syncthreads();
    const double t = sh[32];
    __syncthreads();
    slot ^= 1;
    return t;
  };
  auto reduce = [&](double v) { reduce_put(v); cluster.sync(); return reduce_get(); };

#define CMFD_MY_CELLS(...)                                                         \
  for (int ls = threadIdx.x; ls < mine; ls += (int)blockDim.x)                \
    for (int colour = 0; colour < 2; colour++) {                                   \
      const int cell = cmfd_slot_cell32(a, base + ls, colour, hx);                 \
      if (cell < 0) continue;                                                      \
      const int li = ls * 2 + colour;                                              \
      [[maybe_unused]] const int64_t iv = sm ? (int64_t)li : (int64_t)cell;        \
      __VA_ARGS__                                                                  \
    }

  /* starting guess and initial source (Cmfd.cpp:1232, linalg.cpp:65-80) */
  double local = 0.;
  CMFD_MY_CELLS({
    double x[NCG];
#pragma unroll
    for (int g = 0; g < NCG; g++) { x[g] = a.old_flux[(int64_t)cell * NCG + g]; }
#pragma unroll
    for (int e = 0; e < NCG; e++) {
      double sv = 0.;
#pragma unroll
      for (int g = 0; g < NCG; g++) sv += __ldg(a.mm + ((int64_t)cell * NCG + e) * NCG + g) * x[g];
      Bv[iv * NCG + e] = sv;
      local += sv;
    }
#pragma unroll
    for (int g = 0; g < NCG; g++) Xs[li * NCG + g] = x[g];
  })
  double sum = reduce(local);
  double k = a.cs[CS_KEFF];
  {
    const double fb = (double)nr / sum, fx = (double)nr * k / sum;
    CMFD_MY_CELLS({
#pragma unroll
      for (int e = 0; e < NCG; e++) { Bv[iv * NCG + e] *= fb; Xs[li * NCG + e] *= fx; }
    })
  }
  cluster.sync();

  const double lin_tol = fmax(a.linalg_tol, fmax(a.linalg_tol, source_thresh) * 1e-1);
  const double inv_cells = 1.0 / (double)a.n_cells;
  const double omega = a.sor;
  double initial_residual = 0., residual = 0.;
  int iter = 0, lin_total = 0, lin_iters = 0, lin_iters_1 = 0;
  double lin_res_first = 0., lin_res_1 = 0., res_1 = 0.;
  bool ok = true, converged = false;

  for (iter = 0; iter < 25000; iter++) {
    CMFD_MY_CELLS({
#pragma unroll
      for (int e = 0; e < NCG; e++) {
        double sv = 0.;
#pragma unroll
        for (int g = 0; g < NCG; g++) sv += __ldg(a.mm + ((int64_t)cell * NCG + e) * NCG + g) * Xs[li * NCG + g];
        SOv[iv * NCG + e] = sv;
      }
    })
    double lres = 0., linit = 0., min_res = 1e6;
    int liter = 0;
    while (liter < 10000) {
      const bool need = liter == 0 || liter + 1 > 25;
      const double* old_src = liter >= 25 ? SNv : SOv;
      double part = 0.;
      for (int colour = 0; colour < 2; colour++) {
        for (int ls = threadIdx.x; ls < mine; ls += (int)blockDim.x) {
          const int cell = __ldg(a.slot_cell + (int64_t)colour * n_slots + base + ls);
          if (cell < 0) continue;
          const int li = ls * 2 + colour;
          const int64_t iv = sm ? (int64_t)li : (int64_t)cell;
          int nb[CMFD_NF];
#pragma unroll
          for (int s = 0; s < CMFD_NF; s++) nb[s] = __ldg(ca.nb_loc + (int64_t)cell * CMFD_NF + s);
          double xn[CMFD_NF][NCG], xo[NCG], bv[NCG];
          double sold = 0.;
#pragma unroll
          for (int s = 0; s < CMFD_NF; s++) {
            if (nb[s] >= 0) {
              const double* rx = cluster.map_shared_rank(Xs, nb[s] >> 26) + (size_t)(nb[s] & 0x3ffffff) * NCG;
#pragma unroll
              for (int g = 0; g < NCG; g++) xn[s][g] = rx[g];
            } else {
#pragma unroll
              for (int g = 0; g < NCG; g++) xn[s][g] = 0.;
            }
          }
#pragma unroll
          for (int g = 0; g < NCG; g++) { xo[g] = Xs[li * NCG + g]; bv[g] = Bv[iv * NCG + g]; }
          if (need) {
#pragma unroll
            for (int e = 0; e < NCG; e++) sold += old_src[iv * NCG + e];
          }
#pragma unroll
          for (int g = 0; g < NCG; g++) {
            const int64_t row = (int64_t)cell * NCG + g;
            const double* of = a.off + row * CMFD_NF;
            double v = (1.0 - omega) * xo[g] * __ldg(a.dq + row);
            v -= __ldg(of + 2) * xn[2][g];
            v -= __ldg(of + 1) * xn[1][g];
            v -= __ldg(of + 0) * xn[0][g];
#pragma unroll
            for (int g2 = 0; g2 < NCG; g2++) {
              if (g2 == g) v += bv[g];
              else v -= __ldg(a.ain + row * NCG + g2) * xo[g2];
            }
            v -= __ldg(of + 3) * xn[3][g];
            v -= __ldg(of + 4) * xn[4][g];
            v -= __ldg(of + 5) * xn[5][g];
            xo[g] = v * __ldg(a.qd + row);
          }
          double snew = 0., sn[NCG];
#pragma unroll
          for (int e = 0; e < NCG; e++) {
            double sv = 0.;
#pragma unroll
            for (int g = 0; g < NCG; g++) sv += __ldg(a.mm + ((int64_t)cell * NCG + e) * NCG + g) * xo[g];
            sn[e] = sv;
            snew += sv;
          }
#pragma unroll
          for (int g = 0; g < NCG; g++) { Xs[li * NCG + g] = xo[g]; SNv[iv * NCG + g] = sn[g]; }
          if (need && fabs(sold) > CMFD_FLUX_EPS) { const double q = (snew - sold) / sold; part += q * q; }
        }
        if (colour == 1 && need) reduce_put(part);
        cluster.sync();
      }
      double r = 0.;
      if (need) r = sqrt(fmax(reduce_get(), 0.) * inv_cells);
      if (liter == 0) { lres = r; linit = r; lin_res_first = r; }
      liter++;
      if (liter > 25) {
        lres = r;
        if (lres < min_res) min_res = lres;
        if ((lres > 1e3 * min_res && min_res > 1e-10) || !(lres == lres)) { ok = false; break; }
        if (lres / linit < 0.1 || lres < lin_tol) break;
      }
    }
    lin_total += liter;
    lin_iters = liter;
    if (liter >= 10000) ok = false;
    if (!ok) break;

    local = 0.;
    CMFD_MY_CELLS({
#pragma unroll
      for (int e = 0; e < NCG; e++) local += SNv[iv * NCG + e];
    })
    sum = reduce(local);
    k = sum / (double)nr;
    const double inv_k = 1.0 / k;
    local = 0.;
    CMFD_MY_CELLS({
      double snew = 0., sold = 0.;
#pragma unroll
      for (int e = 0; e < NCG; e++) {
        const double v = SNv[iv * NCG + e] * inv_k;
        snew += v;
        sold += Bv[iv * NCG + e];
        Bv[iv * NCG + e] = v;
      }
      if (fabs(sold) > CMFD_FLUX_EPS) { const double q = (snew - sold) / sold; local += q * q; }
    })
    residual = sqrt(fmax(reduce(local), 0.) * inv_cells);
    if (iter == 0) {
      initial_residual = residual;
      if (initial_residual < 1e-14) initial_residual = 1e-10;
      res_1 = residual; lin_iters_1 = lin_iters; lin_res_1 = lin_res_first;
    }
    if ((residual / initial_residual < 0.03 || residual < a.linalg_tol) && iter > 25) { converged = true; break; }
  }
  if (!converged) ok = false;

  if (ok) {
    double ln = 0., lo = 0.;
    CMFD_MY_CELLS({
#pragma unroll
      for (int e = 0; e < NCG; e++) {
        double sn = 0., so = 0.;
#pragma unroll
        for (int g = 0; g < NCG; g++) {
          const double m = __ldg(a.mm + ((int64_t)cell * NCG + e) * NCG + g);
          sn += m * Xs[li * NCG + g];
          so += m * a.old_flux[(int64_t)cell * NCG + g];
        }
        ln += sn; lo += so;
      }
    })
    const double sum_new = reduce(ln);
    const double sum_old = reduce(lo);
    const double fn = 1.0 / sum_new, fo = 1.0 / sum_old;
    CMFD_MY_CELLS({
#pragma unroll
      for (int e = 0; e < NCG; e++) {
        a.new_flux[(int64_t)cell * NCG + e] = Xs[li * NCG + e] * fn;
        a.old_flux[(int64_t)cell * NCG + e] *= fo;
      }
    })
  }
#undef CMFD_MY_CELLS
  if (rank == 0 && threadIdx.x == 0) {
    a.ci[CI_FAIL] = ok ? 0 : 1;
    a.ci[CI_POWER_ITERS] = iter;
    a.ci[CI_LIN_ITERS_1] = lin_iters_1;
    a.ci[CI_LIN_ITERS_END] = lin_iters;
    a.ci[CI_LIN_TOTAL] = lin_total;
    a.ci[CI_SOLVES] += 1;
    a.cs[CS_RES_1] = res_1; a.cs[CS_RES_END] = residual;
    a.cs[CS_LIN_RES_1] = lin_res_1; a.cs[CS_LIN_RES_END] = lin_res_first;
    if (ok) a.cs[CS_KEFF] = k;
    a.scal[SC_KPREV] = a.scal[SC_KEFF];
    a.scal[SC_KEFF] = a.cs[CS_KEFF];
    *reinterpret_cast<long long*>(&a.ci[CI_PF_MAX]) = 0;
    *reinterpret_cast<long long*>(&a.ci[CI_PF_MIN]) = 0;
  }
  cluster.sync();       /* no CTA leaves while another may still read its shared memory */
}

/* ------------------------------------------------------------------------------------------ */
/* Cmfd::updateMOCFlux: thread per FSR                                                         */
/* ------------------------------------------------------------------------------------------ */
__device__ inline double cmfd_flux_ratio(const CmfdArgs& a, int64_t cell, int e, int64_t fsr) {
  const int ncg = a.ncg;
  if (a.axial_interp && a.nz >= 3) {
    const double* c = a.ax_interp + fsr * 3;
    const int64_t nxy = (int64_t)a.nx * a.ny;
    const int iz = (int)(cell / nxy);
    int64_t mid = cell;
    if (iz == 0) mid += nxy;
    else if (iz == a.nz - 1) mid -= nxy;
    const int64_t prev = mid - nxy, next = mid + nxy;
    const double old_flux = c[0] * a.old_flux[prev * ncg + e] + c[1] * a.old_flux[mid * ncg + e] + c[2] * a.old_flux[next * ncg + e];
    const double new_flux = c[0] * a.new_flux[prev * ncg + e] + c[1] * a.new_flux[mid * ncg + e] + c[2] * a.new_flux[next * ncg + e];
    double ratio = 1.0;
    if (fabs(old_flux) > CMFD_FLUX_EPS) ratio = new_flux / old_flux;
    if (ratio < 0) {
      if (fabs(a.old_flux[cell * ncg + e]) > CMFD_FLUX_EPS) ratio = a.new_flux[cell * ncg + e] / a.old_flux[cell * ncg + e];
      else ratio = 0.0;
    }
    return ratio;
  }
  if (fabs(a.old_flux[cell * ncg + e]) > CMFD_FLUX_EPS) return a.new_flux[cell * ncg + e] / a.old_flux[cell * ncg + e];
  return 0.0;
}

__device__ __forceinline__ long long cmfd_order_bits(double v) {      /* monotone double -> int64 */
  const long long b = __double_as_longlong(v);
  return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}

__global__ void __launch_bounds__(256)
cmfd_update_kernel(CmfdArgs a, int moc_iteration) {
  if (a.iscal[SI_DONE]) return;
  if (a.ci[CI_FAIL]) return;
  if (moc_iteration < 0) moc_iteration = a.iscal[SI_EXEC];
  const int ncg = a.ncg, G = a.G;
  const int64_t plane = a.n_fsr * G;
  double lmax = 0., lmin = 0.;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_fsr; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = a.fsr_cell[r];
    if (cell < 0) continue;
    for (int e = 0; e < ncg; e++) {
      double ratio = 0.;
      if (a.centroid) {
        const int64_t s0 = a.st_off[r], s1 = a.st_off[r + 1];
        for (int64_t j = s0; j < s1; j++) ratio += a.st_w[j] * cmfd_flux_ratio(a, a.st_cell[j], e, r);
        if (a.st_n[r] == 1) ratio += cmfd_flux_ratio(a, cell, e, r);
        else {
          ratio += a.st_own[r] * cmfd_flux_ratio(a, cell, e, r);
          ratio /= (double)(a.st_n[r] - 1);
        }
      } else {
        ratio = cmfd_flux_ratio(a, cell, e, r);
      }
      if (moc_iteration > a.n_unbounded) {
        if (ratio > 20.0) ratio = 20.0;
        else if (ratio < 0.05) ratio = 0.05;
      }
      const double lg = log(ratio);
      if (lg > lmax) lmax = lg;
      if (lg < lmin) lmin = lg;
      for (int h = a.group_idx[e]; h < a.group_idx[e + 1]; h++) {
        a.phi[r * G + h] *= ratio;
        if (a.linear) {
          a.phi_m[r * G + h] *= ratio;
          a.phi_m[plane + r * G + h] *= ratio;
          a.phi_m[2 * plane + r * G + h] *= ratio;
        }
      }
    }
  }
  /* MAX P.F. of the iteration report (Cmfd.cpp:1545-1573): the ratio with the largest |log| */
  for (int o = 16; o > 0; o >>= 1) {
    lmax = fmax(lmax, __shfl_down_sync(0xffffffffu, lmax, o));
    lmin = fmin(lmin, __shfl_down_sync(0xffffffffu, lmin, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (lmax > 0.) atomicMax(reinterpret_cast<long long*>(&a.ci[CI_PF_MAX]), cmfd_order_bits(lmax));
    if (lmin < 0.) atomicMax(reinterpret_cast<long long*>(&a.ci[CI_PF_MIN]), cmfd_order_bits(-lmin));
  }
}

/* after computeResidual: Solver.cpp:1671-1675, _cmfd->setSourceConvergenceThreshold(0.01 * residual) */
__global__ void cmfd_threshold_kernel(double* cs, const double* scal, const int* iscal) {
  if (iscal[SI_DONE]) return;
  double r = scal[SC_RESIDUAL];
  if (r <= 0) r = 1e-6;
  cs[CS_THRESH] = 0.01 * r;
}

}  // namespace b200
