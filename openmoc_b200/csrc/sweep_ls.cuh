/*
 * sweep_ls.cuh - linear-source transport sweep (CPULSSolver) as one sm_100a kernel.
 *
 * Replaces CPULSSolver::tallyLSScalarFlux / accumulateLinearFluxContribution
 * (src/CPULSSolver.cpp:542-780) inside TransportSweep::onTrack
 * (src/TrackTraversingAlgorithms.cpp:890-1052).  Same work decomposition as the flat
 * kernel (sweep.cuh): item = (track, direction), LPI threads per item, one energy group
 * (GPL of them) and all NP polar angles per thread.
 *
 * Per segment the source is  q(s) = q_flat + q_xyz . (x0 + s*Omega)  with x0 the segment's
 * starting point relative to the FSR centroid; the exponentials F1, F2, H all derive from
 * one rational G(tau) = 1/tau - (1-exp(-tau))/tau^2 (expG_fractional,
 * src/exponentials.h:110-145; ExpEvaluator::retrieveExponentialComponents,
 * src/ExpEvaluator.h:349-381).  Four tallies per (FSR, group): phi and its x, y, z moments.
 *
 * The reference walks the forward direction, moves each segment's starting point to its end
 * (CPULSSolver.cpp:736-738) and walks back with the direction reversed; here the backward
 * item computes that end point itself, so the two directions stay independent.
 */
#pragma once
#include "sweep.cuh"

namespace b200 {

struct SweepLSArgs {
  SweepArgs f;                              /* everything the flat sweep needs */
  const double4* __restrict__ seg_pos;      /* {x, y, z, 0} per segment, padded like seg */
  const double* __restrict__ trk_dir;       /* [n_trk][3] unit vector of the forward direction */
  const double4* __restrict__ qxyz;         /* {q_x, q_y, q_z, 0} per (FSR, group) */
  double* __restrict__ phi_m;               /* moment tallies [c*N_FSR*G + fsr*G+e] */
  double cg[12];                            /* expG coefficients p0..p5, d1..d6 */
};

/* 1/x - (1-exp(-x))/x^2, 5/6-order rational of src/exponentials.h:110-145 */
__device__ __forceinline__ double expG(double x, const double (&c)[12]) {
  double den = fma(c[11], x, c[10]);
  den = fma(den, x, c[9]);
  den = fma(den, x, c[8]);
  den = fma(den, x, c[7]);
  den = fma(den, x, c[6]);
  den = fma(den, x, 1.0);
  double num = fma(c[5], x, c[4]);
  num = fma(num, x, c[3]);
  num = fma(num, x, c[2]);
  num = fma(num, x, c[1]);
  num = fma(num, x, c[0]);
  return fast_div(num, den);
}

__device__ __forceinline__ double4 ld_pos(const double4* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
/* {q, sigma_t} and the source moments of one (FSR, group); q_z is not read in 2D */
template <bool IS3D>
__device__ __forceinline__ void ld_src(const double2* __restrict__ qst, const double4* __restrict__ qxyz,
                                       uint32_t idx, double2& qs, double4& qm) {
  qs = __ldg(&qst[idx]);
  const double2* qmp = reinterpret_cast<const double2*>(&qxyz[idx]);
  const double2 q01 = __ldg(qmp);
  double2 q23 = make_double2(0.0, 0.0);
  if (IS3D) q23 = __ldg(qmp + 1);
  qm = make_double4(q01.x, q01.y, q23.x, q23.y);
}

template <int NP, int GPL, bool IS3D>
/* resident CTAs the kernel is compiled for (register cap).  Measured on a B200 (2D C5G7 64 azim / 0.02 cm; 3D 70-group
 * lattice): one group per thread in 2D gains 5 % at three CTAs (80 registers, 32 B of spills: 2.66e11 -> 2.80e11 /s) and
 * loses 15 % at four; the three-groups-per-thread 3D variants lose 43 % at three CTAs (208 B of spills) */
#ifndef B200_LS_BLOCKS
#define B200_LS_BLOCKS 3
#endif
__global__ void __launch_bounds__(224, (GPL == 1 && !IS3D) ? B200_LS_BLOCKS : 0)   /* 0: no occupancy constraint, as before */
sweep_ls_kernel(const SweepLSArgs la) {
  const SweepArgs& a = la.f;
  if (a.done != nullptr && *a.done) return;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t item = gtid / a.lpi;
  const int sub = (int)(gtid - item * a.lpi);
  if (item >= a.n_items) return;

  const int G = a.G;
  const int64_t t = a.order[item >> 1];
  const int dir = (int)(item & 1);
  const int64_t s0 = a.trk_off[t], s1 = a.trk_off[t + 1];
  const int n = (int)(s1 - s0);
  const int cls = a.trk_class[t];

  uint32_t e[GPL];
  bool valid[GPL];
#pragma unroll
  for (int j = 0; j < GPL; j++) {
    int ej = sub + j * a.lpi;
    valid[j] = ej < G;
    e[j] = (uint32_t)(valid[j] ? ej : G - 1);
  }
  double w[NP], inv_sin[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    w[p] = a.cls_w[cls * NP + p];
    inv_sin[p] = a.cls_inv_sin[cls * NP + p];
  }
  /* direction of travel; the reverse item flips it (TrackTraversingAlgorithms.cpp:1006-1008) */
  const double sgn = dir ? -1.0 : 1.0;
  const double dx = sgn * la.trk_dir[t * 3], dy = sgn * la.trk_dir[t * 3 + 1], dz = sgn * la.trk_dir[t * 3 + 2];

  /* tally replica of this CTA: phi copies are rep_stride apart, the three moment planes of one
   * copy follow each other, so a copy of the moments is 3 * rep_stride long */
  const uint32_t rep = (uint32_t)(blockIdx.x & a.rep_mask);
  double* __restrict__ const phi = a.phi;
  double* __restrict__ const phi_m = la.phi_m;
  const uint32_t rep_phi = rep * (uint32_t)a.rep_stride, rep_m = 3u * rep_phi;

  const int F = G * NP;
  const int64_t slot_in = (t * 2 + dir) * (int64_t)F;
  float psi[NP][GPL];
#pragma unroll
  for (int p = 0; p < NP; p++)
#pragma unroll
    for (int j = 0; j < GPL; j++) psi[p][j] = a.psi_in[slot_in + p * G + e[j]];
  if (a.carry[t * 2 + dir]) {
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
      for (int j = 0; j < GPL; j++)
        if (valid[j]) a.psi_out[slot_in + p * G + e[j]] = psi[p][j];
  }

  double acc[GPL], accx[GPL], accy[GPL], accz[GPL];
#pragma unroll
  for (int j = 0; j < GPL; j++) acc[j] = accx[j] = accy[j] = accz[j] = 0.0;

  const int step = dir ? -1 : 1;
  int64_t s = dir ? s1 - 1 : s0;
  /* 3D tracks apply the quadrature weight when the tally is flushed
   * (TrackTraversingAlgorithms.cpp:904-910), 2D tracks per polar angle */
  const double wflush = IS3D ? w[0] : 1.0;

  /* Software pipeline as in the flat kernel: record and starting point two segments ahead,
   * the {q, sigma_t} / {q_x, q_y, q_z} gathers one segment ahead when a thread owns a single
   * group (with three groups the 36 extra registers cost more than the latency they hide).
   * Both streams are padded by SEG_PAD records, so the look-ahead needs no bounds test. */
  constexpr bool PFG = (GPL == 1);
  int4 r0 = ld_rec(a.seg + s), r1 = ld_rec(a.seg + s + step);
  double4 p0 = ld_pos(la.seg_pos + s), p1 = ld_pos(la.seg_pos + s + step);
  double2 qsN[GPL];
  double4 qmN[GPL];
  if (PFG) {
#pragma unroll
    for (int j = 0; j < GPL; j++) ld_src<IS3D>(a.qst, la.qxyz, (uint32_t)r0.z + e[j], qsN[j], qmN[j]);
  }

  for (int i = 0; i < n; i++, s += step) {
    const int4 r2 = ld_rec(a.seg + s + 2 * step);
    const double4 p2 = ld_pos(la.seg_pos + s + 2 * step);
    const uint32_t base = (uint32_t)r0.z, bnext = (uint32_t)r1.z;
    double2 qsC[GPL];
    double4 qmC[GPL];
#pragma unroll
    for (int j = 0; j < GPL; j++) {
      if (PFG) {
        qsC[j] = qsN[j]; qmC[j] = qmN[j];
        ld_src<IS3D>(a.qst, la.qxyz, bnext + e[j], qsN[j], qmN[j]);
      } else {
        ld_src<IS3D>(a.qst, la.qxyz, base + e[j], qsC[j], qmC[j]);
      }
    }
    const double len = __hiloint2double(r0.y, r0.x);
    const double4 pos4 = p0;
    /* starting point of this traversal: the stored point forward, the segment's end backward */
    const double px = dir ? pos4.x - dx * len : pos4.x;   /* pos + dir_fwd*len, dir_fwd = -d */
    const double py = dir ? pos4.y - dy * len : pos4.y;
    const double pz = dir ? pos4.z - dz * len : pos4.z;
    /* segment mid-point times two (CPULSSolver.cpp:560-562, 642-644) */
    const double cx = 2.0 * px + len * dx, cy = 2.0 * py + len * dy, cz = 2.0 * pz + len * dz;

#pragma unroll
    for (int j = 0; j < GPL; j++) {
      const double2 qs = qsC[j];
      const double4 qm = qmC[j];
      const double tau = qs.y * len;
      double src_flat = qs.x + qm.x * cx + qm.y * cy;
      double src_lin = qm.x * dx + qm.y * dy;
      if (IS3D) { src_flat += qm.z * cz; src_lin += qm.z * dz; }
      const double lsf = len * src_flat, l2sl = len * len * src_lin, tl = tau * len;
      /* sum over the polar angles of the (weighted) delta psi and of the H term: the three
       * moment tallies take them once per segment instead of once per polar angle */
      double sumd = 0.0, sumh = 0.0;
#pragma unroll
      for (int p = 0; p < NP; p++) {
        double f1, f2, h;
        if (IS3D) {
          const double g = expG(fmax(1e-8, tau), la.cg);      /* CPULSSolver.cpp:583-587 */
          f1 = 1.0 - tau * g;
          f2 = 2.0 * g - f1;
          h = f1 - g;
        } else {
          /* ExpEvaluator::retrieveExponentialComponents, src/ExpEvaluator.h:361-380 */
          const double tp = fmax(1e-8, tau * inv_sin[p]);
          double g = expG(tp, la.cg);
          f1 = (1.0 - tp * g) * inv_sin[p];
          g *= inv_sin[p];
          f2 = 2.0 * g - f1;
          h = f1 - g;
        }
        const double psid = (double)psi[p][j];
        const double dpsi = fma(-l2sl, f2, (tau * psid - lsf) * f1);
        psi[p][j] = (float)(psid - dpsi);
        if (IS3D) {
          sumh = fma(h * tl, psid, sumh);
          sumd += dpsi;
        } else {
          sumh = fma(h * (w[p] * tl), psid, sumh);
          sumd = fma(w[p], dpsi, sumd);
        }
      }
      acc[j] += sumd;
      accx[j] += fma(sumh, dx, sumd * px);
      accy[j] += fma(sumh, dy, sumd * py);
      if (IS3D) accz[j] += fma(sumh, dz, sumd * pz);
    }

    if (a.seg_cmfd != nullptr) {      /* CMFD surface currents, src/Cmfd.h:572-670 */
      const int2 c = a.seg_cmfd[s];
      const int surf = dir ? c.y : c.x;
      if (surf >= 0) {
#pragma unroll
        for (int j = 0; j < GPL; j++) {
          double cur = 0.0;
#pragma unroll
          for (int p = 0; p < NP; p++) cur = fma(w[p], (double)psi[p][j], cur);
          if (valid[j]) atomicAdd(&a.currents[(size_t)surf * a.ncg + a.cmfd_group[e[j]]], cur);
        }
      }
    }

    if (bnext != base || i == n - 1) {
#pragma unroll
      for (int j = 0; j < GPL; j++) {
        if (valid[j]) {
          const uint32_t idx = base + e[j];
          const uint32_t im = idx + rep_m, ns = (uint32_t)a.rep_stride;   /* 32-bit index math (checked at finalize) */
          atomicAdd(&phi[idx + rep_phi], wflush * acc[j]);
          atomicAdd(&phi_m[im], wflush * accx[j]);
          atomicAdd(&phi_m[im + ns], wflush * accy[j]);
          if (IS3D) atomicAdd(&phi_m[im + 2u * ns], wflush * accz[j]);
        }
        acc[j] = accx[j] = accy[j] = accz[j] = 0.0;
      }
    }
    r0 = r1; r1 = r2;
    p0 = p1; p1 = p2;
  }

  const int64_t out = a.out_slot[t * 2 + dir];
  if (out >= 0) {
    const int peer = (int)(out >> PEER_SHIFT);
    float* __restrict__ dst = peer ? a.peer_out.p[peer - 1] : a.psi_out;
    const int64_t base = (out & PEER_SLOT_MASK) * (int64_t)F;
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
      for (int j = 0; j < GPL; j++)
        if (valid[j]) dst[base + p * G + e[j]] = psi[p][j];
  }
}

/* padded {x, y, z, 0} stream from the uploaded [n_seg][3] starting points */
__global__ void build_segpos_kernel(double4* __restrict__ out, const double* __restrict__ xyz, int64_t n_seg) {
  const int64_t total = n_seg + 2 * SEG_PAD;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i - SEG_PAD;
    double4 r = make_double4(0., 0., 0., 0.);
    if (s >= 0 && s < n_seg) { r.x = xyz[3 * s]; r.y = xyz[3 * s + 1]; r.z = xyz[3 * s + 2]; }
    out[i] = r;
  }
}

}  // namespace b200
