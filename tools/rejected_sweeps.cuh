/*
 * rejected_sweeps.cuh - sweep-kernel variants that were built, measured on the B200 and rejected
 * (profiles/r01_sweep_ablation.md #5): a 4-deep register ring and a cp.async-staged variant.
 * Kept out of the product header for the record; to rebuild one, include this file after
 * openmoc_b200/csrc/sweep.cuh inside namespace-level scope and launch it in place of sweep_kernel.
 */
#pragma once
#include "../openmoc_b200/csrc/sweep.cuh"
#ifndef B200_RING_MINB
#define B200_RING_MINB 1
#endif
namespace b200 {
/* ------------------------------------------------------------------------- */
/* Ring variant (default for GPL <= 2): a 4-deep register ring of segment        */
/* records and {q, sigma_t} pairs, the loop unrolled by 4 so that the ring never   */
/* rotates through register moves.  A record is loaded 3 steps ahead and first     */
/* touched one step later (to form the gather address); the gather is issued 2     */
/* steps ahead.  Wherever ptxas places a load inside a step, a full step           */
/* (~1000 cycles at the measured issue rate) separates it from its first use,      */
/* which covers L2 latency; DRAM latency is covered by the L2 prefetch.            */
/* The tally flush is a predicated RED instead of a branch: one basic block/step.  */
/* ------------------------------------------------------------------------- */

template <typename T, int NP, int GPL>
__global__ void __launch_bounds__(224, B200_RING_MINB)
sweep_kernel_ring(const SweepArgs a) {
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
  T w[NP], inv_sin[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    w[p] = (T)a.cls_w[cls * NP + p];
    inv_sin[p] = (T)a.cls_inv_sin[cls * NP + p];
  }
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
  double acc[GPL];
#pragma unroll
  for (int j = 0; j < GPL; j++) acc[j] = 0.0;

  const int step = dir ? -1 : 1;
  const SegRec* __restrict__ ps = a.seg + (dir ? s1 - 1 : s0);   /* record of step 0 */
  const double2* __restrict__ const qst = a.qst;
  double* __restrict__ const phi = a.phi;
  constexpr int PF_DIST = 24;

  /* ring slots */
  int4 R0, R1, R2, R3;
  double2 Q0[GPL], Q1[GPL], Q2[GPL], Q3[GPL];
  R0 = __ldg(reinterpret_cast<const int4*>(ps));
  R1 = __ldg(reinterpret_cast<const int4*>(ps + step));
  R2 = __ldg(reinterpret_cast<const int4*>(ps + 2 * step));
#pragma unroll
  for (int j = 0; j < GPL; j++) {
    Q0[j] = __ldg(&qst[(uint32_t)R0.z + e[j]]);
    Q1[j] = __ldg(&qst[(uint32_t)R1.z + e[j]]);
  }
  ps += 3 * step;                                               /* next record to load */

/* one segment: CUR/NXT/N2/N3 are ring slots holding steps k, k+1, k+2, k+3 */
#define B200_SWEEP_STEP(RC, QC, RN, RN2, QN2, RN3, LAST)                                       \
  {                                                                                            \
    if ((reinterpret_cast<uintptr_t>(ps) & 0x70) == 0)                                         \
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + PF_DIST * step));                     \
    RN3 = __ldg(reinterpret_cast<const int4*>(ps));                                            \
    ps += step;                                                                                \
    _Pragma("unroll") for (int j = 0; j < GPL; j++) QN2[j] = __ldg(&qst[(uint32_t)RN2.z + e[j]]); \
    const T len = (T)__hiloint2double(RC.y, RC.x);                                             \
    const uint32_t bc = (uint32_t)RC.z;                                                        \
    const bool flush = ((uint32_t)RN.z != bc) || (LAST);                                       \
    _Pragma("unroll") for (int j = 0; j < GPL; j++) {                                          \
      const T tau = (T)QC[j].y * len;                                                          \
      const T lq = len * (T)QC[j].x;                                                           \
      T x[NP], f1[NP];                                                                         \
      _Pragma("unroll") for (int p = 0; p < NP; p++) x[p] = tau * inv_sin[p];                  \
      expF1_batch<T, NP>(x, f1, a.cf);                                                               \
      T sum = (T)0;                                                                            \
      _Pragma("unroll") for (int p = 0; p < NP; p++) {                                         \
        const T ex = inv_sin[p] * f1[p];                                                       \
        const T dpsi = (tau * (T)psi[p][j] - lq) * ex;                                         \
        psi[p][j] = (float)((T)psi[p][j] - dpsi);                                              \
        sum = fma(w[p], dpsi, sum);                                                            \
      }                                                                                        \
      acc[j] += (double)sum;                                                                   \
      red_add_if(&phi[bc + e[j]], acc[j], flush && valid[j]);                                  \
      acc[j] = flush ? 0.0 : acc[j];                                                           \
    }                                                                                          \
  }

  int i = 0;
  for (; i + 4 < n; i += 4) {          /* strictly less: the last segment goes through the tail */
    B200_SWEEP_STEP(R0, Q0, R1, R2, Q2, R3, false)
    B200_SWEEP_STEP(R1, Q1, R2, R3, Q3, R0, false)
    B200_SWEEP_STEP(R2, Q2, R3, R0, Q0, R1, false)
    B200_SWEEP_STEP(R3, Q3, R0, R1, Q1, R2, false)
  }
  /* tail: 1..4 segments left (n > 0), the very last one always flushes */
  const int rem = n - i;
  if (rem > 0) { B200_SWEEP_STEP(R0, Q0, R1, R2, Q2, R3, rem == 1) }
  if (rem > 1) { B200_SWEEP_STEP(R1, Q1, R2, R3, Q3, R0, rem == 2) }
  if (rem > 2) { B200_SWEEP_STEP(R2, Q2, R3, R0, Q0, R1, rem == 3) }
  if (rem > 3) { B200_SWEEP_STEP(R3, Q3, R0, R1, Q1, R2, true) }
#undef B200_SWEEP_STEP

  const int64_t out = a.out_slot[t * 2 + dir];
  if (out >= 0) {
    const int64_t base = out * (int64_t)F;
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
      for (int j = 0; j < GPL; j++)
        if (valid[j]) a.psi_out[base + p * G + e[j]] = psi[p][j];
  } else if (a.leakage != nullptr) {
    /* vacuum end: leakage tally of transferBoundaryFlux (src/CPUSolver.cpp:2592-2600); the
     * reference weighs every flux of a 2D track with the weight of polar index 0 */
    double lk = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
      for (int j = 0; j < GPL; j++)
        if (valid[j]) lk += (double)psi[p][j];
    atomicAdd(&a.leakage[t], (float)((double)a.cls_w[cls * NP] * lk));
  }
}

/* ------------------------------------------------------------------------- */
/* Staged variant: the segment records and the {q, sigma_t} gathers are copied  */
/* global -> shared with cp.async (LDGSTS) into per-thread rings, DQ segments    */
/* ahead for the gathers and 2*DQ ahead for the records, so neither DRAM nor L2   */
/* latency is ever waited for and no register holds in-flight data (ptxas sinks   */
/* plain look-ahead loads next to their use, which exposed ~20 % long-scoreboard  */
/* stalls in the register-pipelined kernel above: profiles/).                     */
/* Rings are laid out [slot][thread] in 16-byte words: conflict-free LDS.128.     */
/* ------------------------------------------------------------------------- */
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <typename T, int NP, int GPL, int DQ>
__global__ void __launch_bounds__(224)
sweep_kernel_staged(const SweepArgs a) {
  constexpr int D = 2 * DQ;                 /* record look-ahead (power of two) */
  extern __shared__ int4 smem[];            /* [D][nthr] records, then [DQ][GPL][nthr] {q, sigma_t} */
  if (a.done != nullptr && *a.done) return;
  const int nthr = blockDim.x;
  const int tid = threadIdx.x;
  const int64_t gtid = (int64_t)blockIdx.x * nthr + tid;
  const int64_t item = gtid / a.lpi;
  const int sub = (int)(gtid - item * a.lpi);
  if (item >= a.n_items) return;            /* no block-wide barrier is used below */

  int4* const rec_ring = smem + tid;                       /* slot stride nthr */
  int4* const qs_ring = smem + D * nthr + tid;             /* slot stride GPL*nthr, group stride nthr */
  const uint32_t rec_s = (uint32_t)__cvta_generic_to_shared(rec_ring);
  const uint32_t qs_s = (uint32_t)__cvta_generic_to_shared(qs_ring);
  const uint32_t slot_b = (uint32_t)nthr * 16u;

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
  T w[NP], inv_sin[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    w[p] = (T)a.cls_w[cls * NP + p];
    inv_sin[p] = (T)a.cls_inv_sin[cls * NP + p];
  }
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
  double acc[GPL];
#pragma unroll
  for (int j = 0; j < GPL; j++) acc[j] = 0.0;

  const int step = dir ? -1 : 1;
  const SegRec* __restrict__ const first = a.seg + (dir ? s1 - 1 : s0);
  const double2* __restrict__ const qst = a.qst;
  double* __restrict__ const phi = a.phi;

  /* prologue: records 0..D-DQ-1 synchronously, then DQ groups {rec[D-DQ+j], qs[j]} */
#pragma unroll
  for (int k = 0; k < D - DQ; k++) cp_async16(rec_s + k * slot_b, first + k * step);
  cp_async_commit();
  cp_async_wait<0>();
#pragma unroll
  for (int j = 0; j < DQ; j++) {
    cp_async16(rec_s + (D - DQ + j) * slot_b, first + (D - DQ + j) * step);
    const uint32_t bj = (uint32_t)rec_ring[j * nthr].z;
#pragma unroll
    for (int g = 0; g < GPL; g++) cp_async16(qs_s + (j * GPL + g) * slot_b, qst + (bj + e[g]));
    cp_async_commit();
  }
  int4 rc = rec_ring[0];
  double L0 = __hiloint2double(rc.y, rc.x);
  uint32_t b0 = (uint32_t)rc.z;

  for (int i = 0; i < n; i++) {
    cp_async_wait<DQ - 1>();                /* group of step i-DQ: qs[i], rec[i+DQ] have landed */
    const int sr = i & (D - 1), sq = i & (DQ - 1);
    double2 qs[GPL];
#pragma unroll
    for (int g = 0; g < GPL; g++) {
      const int4 v = qs_ring[(sq * GPL + g) * nthr];
      qs[g] = make_double2(__hiloint2double(v.y, v.x), __hiloint2double(v.w, v.z));
    }
    const int4 rn = rec_ring[((i + 1) & (D - 1)) * nthr];            /* next record (flush test) */
    const uint32_t bq = (uint32_t)rec_ring[((i + DQ) & (D - 1)) * nthr].z;
    /* refill the two slots just consumed: rec[i+D] -> slot of rec[i], qs[i+DQ] -> slot of qs[i] */
    cp_async16(rec_s + sr * slot_b, first + (int64_t)(i + D) * step);
#pragma unroll
    for (int g = 0; g < GPL; g++) cp_async16(qs_s + (sq * GPL + g) * slot_b, qst + (bq + e[g]));
    cp_async_commit();
    const uint32_t b1 = (uint32_t)rn.z;

    const T len = (T)L0;
#pragma unroll
    for (int j = 0; j < GPL; j++) {
      const T tau = (T)qs[j].y * len;
      const T lq = len * (T)qs[j].x;
      T x[NP], f1[NP];
#pragma unroll
      for (int p = 0; p < NP; p++) x[p] = tau * inv_sin[p];
      expF1_batch<T, NP>(x, f1, a.cf);
      T sum = (T)0;
#pragma unroll
      for (int p = 0; p < NP; p++) {
        const T ex = inv_sin[p] * f1[p];
        const T dpsi = (tau * (T)psi[p][j] - lq) * ex;
        psi[p][j] = (float)((T)psi[p][j] - dpsi);
        sum = fma(w[p], dpsi, sum);
      }
      acc[j] += (double)sum;
    }
    if (b1 != b0 || i == n - 1) {
#pragma unroll
      for (int j = 0; j < GPL; j++) {
        if (valid[j]) atomicAdd(&phi[b0 + e[j]], acc[j]);
        acc[j] = 0.0;
      }
    }
    L0 = __hiloint2double(rn.y, rn.x);
    b0 = b1;
  }
  cp_async_wait<0>();

  const int64_t out = a.out_slot[t * 2 + dir];
  if (out >= 0) {
    const int64_t base = out * (int64_t)F;
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
      for (int j = 0; j < GPL; j++)
        if (valid[j]) a.psi_out[base + p * G + e[j]] = psi[p][j];
  } else if (a.leakage != nullptr) {
    /* vacuum end: leakage tally of transferBoundaryFlux (src/CPUSolver.cpp:2592-2600); the
     * reference weighs every flux of a 2D track with the weight of polar index 0 */
    double lk = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
      for (int j = 0; j < GPL; j++)
        if (valid[j]) lk += (double)psi[p][j];
    atomicAdd(&a.leakage[t], (float)((double)a.cls_w[cls * NP] * lk));
  }
}


/* Round 2: the ring variant brought up to date (tally replicas, peer hand-offs) and tried on the 3D C5G7
 * deck (1.06e9 segments, one polar angle per track), where 46 % of the stall samples of sweep_kernel sit on
 * the long-scoreboard wait for the {q, sigma_t} gather: 47.9 ms (sweep_kernel) against 56.9 / 72.5 / 134 ms
 * with 4 / 5 / 6 resident CTAs per SM (72 / 56 / 40 registers, 24 / 80 / 168 bytes of spills).  Rejected. */
/* ------------------------------------------------------------------------------------
 * Deep-pipeline variant for the latency-bound sweeps (3D tracks: one polar angle per track, so a
 * segment is ~110 instructions of one warp and the {q, sigma_t} gather issued one segment ahead by
 * sweep_kernel is not back in time: 46 % of the stall samples of the 3D C5G7 sweep sit on that
 * long-scoreboard wait, profiles/r02_sweep3d.md).  Here the records travel three segments ahead and
 * the gathers two, through a 4-slot register ring whose rotation is unrolled away; MINB trades
 * registers for resident warps.  Plain double tally only (no fixed-point tally, no CMFD currents).
 * ------------------------------------------------------------------------------------ */
template <typename T, int NP, int GPL, int MINB>
__global__ void __launch_bounds__(224, MINB)
sweep_kernel_deep(const SweepArgs a) {
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
    valid[j] = a.exact || ej < G;
    e[j] = (uint32_t)(valid[j] ? ej : G - 1);
  }
  T w[NP], inv_sin[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    w[p] = (T)a.cls_w[cls * NP + p];
    inv_sin[p] = (T)a.cls_inv_sin[cls * NP + p];
  }
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
  double acc[GPL];
#pragma unroll
  for (int j = 0; j < GPL; j++) acc[j] = 0.0;

  const int step = dir ? -1 : 1;
  const SegRec* __restrict__ ps = a.seg + (dir ? s1 - 1 : s0);   /* record of step 0 */
  const double2* __restrict__ const qst = a.qst;
  double* __restrict__ const phi = a.phi;
  const uint32_t rep_idx = (uint32_t)(blockIdx.x & a.rep_mask) * (uint32_t)a.rep_stride;
  uint32_t et[GPL];
#pragma unroll
  for (int j = 0; j < GPL; j++) et[j] = e[j] + rep_idx;

  /* ring slots: records of steps k..k+3, gathers of steps k..k+2 */
  int4 R0, R1, R2, R3;
  double2 Q0[GPL], Q1[GPL], Q2[GPL], Q3[GPL];
  R0 = ld_rec(ps);
  R1 = ld_rec(ps + step);
  R2 = ld_rec(ps + 2 * step);
#pragma unroll
  for (int j = 0; j < GPL; j++) {
    Q0[j] = ld_qs(&qst[(uint32_t)R0.z + e[j]]);
    Q1[j] = ld_qs(&qst[(uint32_t)R1.z + e[j]]);
  }
  ps += 3 * step;                                               /* next record to load */

#define B200_DEEP_STEP(RC, QC, RN, RN2, QN2, RN3, LAST)                                        \
  {                                                                                            \
    RN3 = ld_rec(ps);                                                                          \
    ps += step;                                                                                \
    _Pragma("unroll") for (int j = 0; j < GPL; j++) QN2[j] = ld_qs(&qst[(uint32_t)RN2.z + e[j]]); \
    const T len = (T)__hiloint2double(RC.y, RC.x);                                             \
    const uint32_t bc = (uint32_t)RC.z;                                                        \
    const bool flush = ((uint32_t)RN.z != bc) || (LAST);                                       \
    _Pragma("unroll") for (int j = 0; j < GPL; j++) {                                          \
      const T tau = (T)QC[j].y * len;                                                          \
      const T lq = len * (T)QC[j].x;                                                           \
      T x[NP], f1[NP];                                                                         \
      _Pragma("unroll") for (int p = 0; p < NP; p++) x[p] = tau * inv_sin[p];                  \
      expF1_batch<T, NP>(x, f1, a.cf);                                                         \
      _Pragma("unroll") for (int p = 0; p < NP; p++) {                                         \
        const T ex = inv_sin[p] * f1[p];                                                       \
        const T dpsi = (tau * (T)psi[p][j] - lq) * ex;                                         \
        psi[p][j] = (float)((T)psi[p][j] - dpsi);                                              \
        if constexpr (sizeof(T) == 8) acc[j] = fma((double)w[p], (double)dpsi, acc[j]);        \
        else acc[j] += (double)(w[p] * dpsi);                                                  \
      }                                                                                        \
      red_add_if(&phi[bc + et[j]], acc[j], flush && valid[j]);                                 \
      acc[j] = flush ? 0.0 : acc[j];                                                           \
    }                                                                                          \
  }

  int i = 0;
  for (; i + 4 < n; i += 4) {          /* strictly less: the last segment goes through the tail */
    B200_DEEP_STEP(R0, Q0, R1, R2, Q2, R3, false)
    B200_DEEP_STEP(R1, Q1, R2, R3, Q3, R0, false)
    B200_DEEP_STEP(R2, Q2, R3, R0, Q0, R1, false)
    B200_DEEP_STEP(R3, Q3, R0, R1, Q1, R2, false)
  }
  /* tail: 1..4 segments left (n > 0), the very last one always flushes */
  const int rem = n - i;
  if (rem > 0) { B200_DEEP_STEP(R0, Q0, R1, R2, Q2, R3, rem == 1) }
  if (rem > 1) { B200_DEEP_STEP(R1, Q1, R2, R3, Q3, R0, rem == 2) }
  if (rem > 2) { B200_DEEP_STEP(R2, Q2, R3, R0, Q0, R1, rem == 3) }
  if (rem > 3) { B200_DEEP_STEP(R3, Q3, R0, R1, Q1, R2, true) }
#undef B200_DEEP_STEP

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
  } else if (a.leakage != nullptr) {
    double lk = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
      for (int j = 0; j < GPL; j++)
        if (valid[j]) lk += (double)psi[p][j];
    atomicAdd(&a.leakage[t], (float)((double)a.cls_w[cls * NP] * lk));
  }
}

}  // namespace b200
