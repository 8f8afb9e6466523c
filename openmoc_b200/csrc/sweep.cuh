/*
 * sweep.cuh - the MOC transport sweep as one sm_100a kernel.
 *
 * Replaces CPUSolver::transportSweep -> TransportSweep::onTrack ->
 * tallyScalarFlux / accumulateScalarFluxContribution / transferBoundaryFlux
 * (src/CPUSolver.cpp:2338-2601, src/TrackTraversingAlgorithms.cpp:890-1052) and
 * the reference GPUSolver's transportSweepOnDevice (GPUSolver.cu:577-654).
 *
 * Work decomposition:
 *   item   = (track, direction).  The sweep is Jacobi across tracks AND across
 *            the two directions of one track (each reads its own start flux,
 *            src/CPUSolver.cpp:2351,2586), so 2*N_trk items are independent.
 *   lanes  = LPI consecutive threads own one item; thread `sub` owns energy groups
 *            sub, sub+LPI, ... (GPL of them) and all NP polar angles of each
 *            (psi in registers, fp32 like the reference's float track flux).
 *            G=7: LPI=7, 32 items per 224-thread CTA (items may straddle warps,
 *            so no lane idles); G=70: LPI=10 x GPL=7.
 *   stream = segments of a track are contiguous 16-byte records {f64 length,
 *            u32 FSR*G}; threads of an item read the same address (one L1
 *            broadcast, one LDG.128), the record two steps ahead and the
 *            {q, sigma_t} pair of the next segment's FSR (a single 16-byte
 *            gather) are in flight while the current segment is attenuated.
 *   tally  = per-lane register accumulation while consecutive segments share an
 *            FSR (mirrors TrackTraversingAlgorithms.cpp:982,1030), then one
 *            fire-and-forget RED.ADD.F64 per (FSR, group).
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

/* ---- (1-exp(-x))/x, 5/6-order rational of src/exponentials.h:156-192 ---- */
template <typename T> struct F1Coef;
template <> struct F1Coef<double> {
  static constexpr double p0 = 1.0, p1 = 2.4172687328033081e-1, p2 = 6.2804790965268531e-2,
      p3 = 1.0567595009016521e-2, p4 = 1.0059468082903561e-3, p5 = 1.9309063097411041e-4;
  static constexpr double d0 = 1.0, d1 = 7.4169266112320541e-1, d2 = 2.6722515319494311e-1,
      d3 = 6.1643725066901411e-2, d4 = 1.0590759992367811e-2, d5 = 1.0057980007137651e-3,
      d6 = 1.9309063097411041e-4;
};
template <> struct F1Coef<float> {
  static constexpr float p0 = 1.0f, p1 = 2.4172687328033081e-1f, p2 = 6.2804790965268531e-2f,
      p3 = 1.0567595009016521e-2f, p4 = 1.0059468082903561e-3f, p5 = 1.9309063097411041e-4f;
  static constexpr float d0 = 1.0f, d1 = 7.4169266112320541e-1f, d2 = 2.6722515319494311e-1f,
      d3 = 6.1643725066901411e-2f, d4 = 1.0590759992367811e-2f, d5 = 1.0057980007137651e-3f,
      d6 = 1.9309063097411041e-4f;
};

/* 1/d for d >= 1: hardware seed + two Newton steps (4 DFMA) instead of the
 * IEEE division slow path; the denominator polynomial is >= 1 for x >= 0. */
#ifndef B200_NR_STEPS
#define B200_NR_STEPS 2
#endif
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
#pragma unroll
  for (int it = 0; it < B200_NR_STEPS; it++) {
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
  }
  return r;
}
__device__ __forceinline__ float fast_rcp(float d) { return __frcp_rn(d); }
/* n/d: MUFU.RCP64H seed (error e ~ 2^-20) times (1 + e + e^2): relative error e^3 */
__device__ __forceinline__ double fast_div(double n, double d) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
  double e = fma(-d, r0, 1.0);
  const double n0 = n * r0;
  e = fma(e, e, e);
  return fma(n0, e, n0);
}
__device__ __forceinline__ float fast_div(float n, float d) { return n * __frcp_rn(d); }
__device__ __forceinline__ double fast_rcp_seed(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return r;
}
__device__ __forceinline__ float fast_rcp_seed(float d) { return __frcp_rn(d); }

/* The double coefficients live in constant memory so that every DFMA takes its
 * coefficient straight from the constant bank: as literals the compiler
 * re-materialises all 13 of them with 23 UMOV/IMAD.MOV per segment. */
__constant__ double c_F1[13] = {
    F1Coef<double>::d1, F1Coef<double>::d2, F1Coef<double>::d3, F1Coef<double>::d4,
    F1Coef<double>::d5, F1Coef<double>::d6, F1Coef<double>::p1, F1Coef<double>::p2,
    F1Coef<double>::p3, F1Coef<double>::p4, F1Coef<double>::p5, 1.0, 0.0};

__device__ __forceinline__ double expF1(double x) {
  double den = fma(c_F1[5], x, c_F1[4]);
  den = fma(den, x, c_F1[3]);
  den = fma(den, x, c_F1[2]);
  den = fma(den, x, c_F1[1]);
  den = fma(den, x, c_F1[0]);
  den = fma(den, x, 1.0);
  double num = fma(c_F1[10], x, c_F1[9]);
  num = fma(num, x, c_F1[8]);
  num = fma(num, x, c_F1[7]);
  num = fma(num, x, c_F1[6]);
  num = fma(num, x, 1.0);
  return fast_div(num, den);
}

template <typename T>
__device__ __forceinline__ T expF1(T x) {
  using C = F1Coef<T>;
  T den = fma(C::d6, x, C::d5);
  den = fma(den, x, C::d4);
  den = fma(den, x, C::d3);
  den = fma(den, x, C::d2);
  den = fma(den, x, C::d1);
  den = fma(den, x, C::d0);
  T num = fma(C::p5, x, C::p4);
  num = fma(num, x, C::p3);
  num = fma(num, x, C::p2);
  num = fma(num, x, C::p1);
  num = fma(num, x, C::p0);
  return fast_div(num, den);
}

/* One segment of the device stream: 16 bytes, read with a single LDG.128.
 * `base` is the FSR id premultiplied by G, so {q, sigma_t} and the tally slot of
 * group e sit at element base + e (32-bit index arithmetic in the hot loop). */
/* NP rationals at once, every Horner step issued for all of them before the next */
template <typename T, int NP>
__device__ __forceinline__ void expF1_batch(const T (&x)[NP], T (&out)[NP], const double (&cf)[11]) {
  T den[NP], num[NP];
  if constexpr (sizeof(T) == 8) {
#pragma unroll
    for (int p = 0; p < NP; p++) den[p] = fma(x[p], cf[5], cf[4]);
#pragma unroll
    for (int p = 0; p < NP; p++) num[p] = fma(x[p], cf[10], cf[9]);
#pragma unroll
    for (int k = 3; k >= 0; k--) {
#pragma unroll
      for (int p = 0; p < NP; p++) den[p] = fma(den[p], x[p], cf[k]);
#pragma unroll
      for (int p = 0; p < NP; p++) num[p] = fma(num[p], x[p], k > 0 ? cf[5 + k] : 1.0);
    }
#pragma unroll
    for (int p = 0; p < NP; p++) den[p] = fma(den[p], x[p], 1.0);
  } else {
    using C = F1Coef<float>;
    const float dc[6] = {C::d1, C::d2, C::d3, C::d4, C::d5, C::d6};
    const float pc[5] = {C::p1, C::p2, C::p3, C::p4, C::p5};
#pragma unroll
    for (int p = 0; p < NP; p++) den[p] = fma(dc[5], x[p], dc[4]);
#pragma unroll
    for (int p = 0; p < NP; p++) num[p] = fma(pc[4], x[p], pc[3]);
#pragma unroll
    for (int k = 3; k >= 0; k--) {
#pragma unroll
      for (int p = 0; p < NP; p++) den[p] = fma(den[p], x[p], dc[k]);
#pragma unroll
      for (int p = 0; p < NP; p++) num[p] = fma(num[p], x[p], k > 0 ? pc[k - 1] : 1.0f);
    }
#pragma unroll
    for (int p = 0; p < NP; p++) den[p] = fma(den[p], x[p], 1.0f);
  }
  if constexpr (sizeof(T) == 8) {
    /* num/den without a division: seed r0 = MUFU.RCP64H (relative error e ~ 2^-20), then
     * 1/den = r0/(1-e) = r0 (1 + e + e^2) (1 + O(e^3)), i.e. full double precision from four
     * DFMA/DMUL, two of them independent (two Newton steps + the product take five, in one
     * dependent chain) */
    T r0[NP], e[NP], n0[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) r0[p] = fast_rcp_seed(den[p]);
#pragma unroll
    for (int p = 0; p < NP; p++) e[p] = fma(-den[p], r0[p], (T)1);
#pragma unroll
    for (int p = 0; p < NP; p++) n0[p] = num[p] * r0[p];
#pragma unroll
    for (int p = 0; p < NP; p++) e[p] = fma(e[p], e[p], e[p]);
#pragma unroll
    for (int p = 0; p < NP; p++) out[p] = fma(n0[p], e[p], n0[p]);
  } else {
#pragma unroll
    for (int p = 0; p < NP; p++) out[p] = num[p] * fast_rcp_seed(den[p]);
  }
}

#ifndef B200_REC_HINT
#define B200_REC_HINT 0     /* 0: ld.global.nc  1: L1::evict_last */
#endif
#ifndef B200_QS_HINT
#define B200_QS_HINT 0      /* 0: ld.global.nc  1: L1::evict_first  2: L1::no_allocate */
#endif
__device__ __forceinline__ int4 ld_rec(const void* p) {
  int4 r;
#if B200_REC_HINT == 1
  asm volatile("ld.global.nc.L1::evict_last.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#else
  r = __ldg(reinterpret_cast<const int4*>(p));
#endif
  return r;
}
__device__ __forceinline__ double2 ld_qs(const double2* p) {
#if B200_QS_HINT == 1
  double2 r;
  asm volatile("ld.global.nc.L1::evict_first.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
#elif B200_QS_HINT == 2
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
#else
  return __ldg(p);
#endif
}

__device__ __forceinline__ void red_add_if(double* addr, double v, bool p) {
  asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.u32 pp, %2, 0;\n\t@pp red.global.add.f64 [%0], %1;\n\t}"
               ::"l"(addr), "d"(v), "r"((uint32_t)p) : "memory");
}

#ifndef B200_ACC_DIRECT
#define B200_ACC_DIRECT 1
#endif
#ifndef B200_SWEEP_UNROLL
#define B200_SWEEP_UNROLL 1
#endif

struct __align__(16) SegRec {
  double len;
  uint32_t base;
  uint32_t spare;
};
constexpr int SEG_PAD = 40;  /* sentinel records before and after the stream (covers look-ahead + L2 prefetch distance) */

/* Hand-offs that cross devices (a multi-device group whose tracks are sharded track by track,
 * group.cuh): the destination slot carries the owning shard + 1 in its top 16 bits, and the sweep
 * stores the outgoing flux straight into that shard's start-flux buffer through its peer pointer
 * (NVLink P2P stores from inside the sweep kernel: the role of CPUSolver::transferAllInterfaceFluxes,
 * src/CPUSolver.cpp:1063-1211, with no pack / send / unpack step). */
constexpr int PEER_SHIFT = 48;
constexpr int64_t PEER_SLOT_MASK = ((int64_t)1 << PEER_SHIFT) - 1;
struct PeerOut { float* p[16]; };

struct SweepArgs {
  /* segment stream, padded by SEG_PAD records at both ends; element i of the
   * logical stream is seg[i] (the pointer already skips the front padding) */
  const SegRec* __restrict__ seg;
  /* per track */
  const int64_t* __restrict__ trk_off;     /* n_trk + 1 */
  const int32_t* __restrict__ trk_class;   /* angle class -> rows of cls_w / cls_inv_sin */
  const int32_t* __restrict__ order;       /* item>>1 -> track id */
  /* per (track, dir) */
  const int64_t* __restrict__ out_slot;    /* start-flux slot fed by this end, -1: none (vacuum) */
  const uint8_t* __restrict__ carry;       /* 1: nobody feeds this slot -> copy psi_in through */
  /* angle-class tables [n_class][NP] */
  const double* __restrict__ cls_w;
  const double* __restrict__ cls_inv_sin;
  /* per (FSR, group): {q, sigma_t} */
  const double2* __restrict__ qst;
  /* fluxes */
  const float* __restrict__ psi_in;
  float* __restrict__ psi_out;
  PeerOut peer_out;                        /* psi_out of every shard of the group (unused entries NULL) */
  double* __restrict__ phi;                /* tally target [n_fsr*G] */
  /* Tally replicas: CTA b adds into copy (b & rep_mask) of the tally (copies rep_stride
   * elements apart, copy 0 = phi itself) and fold_replicas_kernel sums them after the
   * sweep.  With few FSRs (512 in the lattice decks) same-address RED contention, not
   * arithmetic, bounds the sweep; R copies divide it by R.  rep_mask = 0: one copy. */
  int64_t rep_stride;
  int rep_mask;
  /* deterministic mode: the tally is accumulated as 64-bit fixed point (integer adds are
   * associative, so the result does not depend on the order the atomics land in) */
  unsigned long long* __restrict__ phi_fx; /* [n_fsr*G], used by the DET kernels */
  const double* __restrict__ fx_scale;     /* device scalar: power-of-two scale */
  /* CMFD surface-current tally (Cmfd::tallyCurrent, src/Cmfd.h:572-670); NULL when CMFD is off */
  const int2* __restrict__ seg_cmfd;       /* {surface crossed at the forward end, at the backward end} or -1, padded like seg */
  const int32_t* __restrict__ cmfd_group;  /* MOC group -> CMFD group */
  double* __restrict__ currents;           /* [(cell*26 + surface)*ncg + g] */
  int ncg;
  float* __restrict__ leakage;             /* [n_trk] vacuum leakage tally, NULL unless k_eff from neutron balance */
  const int* __restrict__ done;            /* device convergence flag (may be NULL) */
  int64_t n_items;
  int G, lpi;
  int exact;                               /* G == GPL * LPI */
  /* expF1 coefficients d1..d6, p1..p5 (src/exponentials.h:159-173).  As kernel parameters
   * they sit in constant bank 0 and every Horner DFMA takes its coefficient as a c[0][..]
   * operand: two register pairs per DFMA, which the register file can feed every 2 cycles
   * (three distinct 64-bit register operands cost a third cycle). */
  double cf[11];
};

/* sums the tally replicas into copy 0 and clears the others for the next sweep */
template <typename V>
__global__ void fold_replicas_kernel(V* __restrict__ base, int64_t n, int64_t stride, int n_rep) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    V sum = base[i];
    for (int r = 1; r < n_rep; r++) {
      sum += base[r * stride + i];
      base[r * stride + i] = V(0);
    }
    base[i] = sum;
  }
}

/* CTAs are at most 224 threads (7 warps: 32 items of 7 lanes for G = 7).  For the
 * small-G shapes four CTAs per SM (28 warps) are worth more than registers: the
 * bound caps the kernel at 72 registers. */
template <typename T, int NP, int GPL, bool DET, bool CMFD>
#ifndef B200_LB_THREADS
#define B200_LB_THREADS 224
#define B200_LB_BLOCKS 4
#endif
#ifndef B200_LB_BLOCKS_3D
#define B200_LB_BLOCKS_3D B200_LB_BLOCKS     /* NP == 1 (3D tracks): experiments with more resident CTAs */
#endif
#ifndef B200_LB_BLOCKS_CMFD
#define B200_LB_BLOCKS_CMFD 4                /* resident CTAs the current-tally variants are compiled for: 72 registers, no spill (was 106 at 2 CTAs: 7.98 -> 6.86 ms on C3) */
#endif
__global__ void __launch_bounds__(B200_LB_THREADS, (GPL == 1 && NP <= 3) ? (CMFD ? B200_LB_BLOCKS_CMFD : (NP == 1 ? B200_LB_BLOCKS_3D : B200_LB_BLOCKS)) : 1)
sweep_kernel(const SweepArgs a) {
  if (a.done != nullptr && *a.done) return;
  /* flat mapping: LPI consecutive threads own one item; an item may straddle two
   * warps (no warp collectives are used), so every lane of every warp is busy */
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

  /* energy groups of this lane (clamped duplicates are computed but never stored) */
  uint32_t e[GPL];
  bool valid[GPL];
#pragma unroll
  for (int j = 0; j < GPL; j++) {
    int ej = sub + j * a.lpi;
    /* a.exact: G == GPL*LPI, every slot is a real group (G = 7, 70, ...): no clamping */
    valid[j] = a.exact || ej < G;
    e[j] = (uint32_t)(valid[j] ? ej : G - 1);
  }

  T w[NP], inv_sin[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    w[p] = (T)a.cls_w[cls * NP + p];
    inv_sin[p] = (T)a.cls_inv_sin[cls * NP + p];
  }

  /* incoming angular flux */
  const int F = G * NP;
  const int64_t slot_in = (t * 2 + dir) * (int64_t)F;
#ifdef B200_EXP_PSID
  double psi[NP][GPL];
#else
  float psi[NP][GPL];
#endif
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

  /* Software pipeline: the segment record two steps ahead and the {q, sigma_t}
   * pair one step ahead are in flight while the current segment is attenuated.
   * The stream is padded, so the look-ahead needs no bounds predicate: past the
   * end of the track it reads the neighbouring track's (or a sentinel) record,
   * whose only effect is an early flush of the tally. */
  const int step = dir ? -1 : 1;
  const SegRec* __restrict__ ps = a.seg + (dir ? s1 - 1 : s0);
  const int4 r0 = ld_rec(ps);
  const int4 r1 = ld_rec(ps + step);
  double L0 = __hiloint2double(r0.y, r0.x), L1 = __hiloint2double(r1.y, r1.x);
  uint32_t b0 = (uint32_t)r0.z, b1 = (uint32_t)r1.z;
  double2 qs0[GPL], qs1[GPL];
#pragma unroll
  for (int j = 0; j < GPL; j++) qs0[j] = ld_qs(&a.qst[b0 + e[j]]);
  ps += 2 * step;

  /* tally replica of this CTA, folded into the 32-bit group index (n_rep * N_FSR * G < 2^32 is
   * checked at finalize): the RED address stays one 32-bit add and one IMAD.WIDE off a uniform base */
  const uint32_t rep_idx = (uint32_t)(blockIdx.x & a.rep_mask) * (uint32_t)a.rep_stride;
  uint32_t et[GPL];
#pragma unroll
  for (int j = 0; j < GPL; j++) et[j] = e[j] + rep_idx;
  double* __restrict__ const phi = a.phi;
  [[maybe_unused]] unsigned long long* __restrict__ const phi_fx = DET ? a.phi_fx : nullptr;
  const double fx_scale = DET ? *a.fx_scale : 0.0;
  [[maybe_unused]] const int2* __restrict__ pc = CMFD ? a.seg_cmfd + (dir ? s1 - 1 : s0) : nullptr;
  [[maybe_unused]] int cg[GPL];
  if constexpr (CMFD) {
#pragma unroll
    for (int j = 0; j < GPL; j++) cg[j] = a.cmfd_group[e[j]];
  }
  constexpr int kUnroll = B200_SWEEP_UNROLL;
#pragma unroll kUnroll
  for (int i = 0; i < n; i++) {
    const int4 r2 = ld_rec(ps);
#pragma unroll
#ifndef B200_EXP_NOGATHER
    for (int j = 0; j < GPL; j++) qs1[j] = ld_qs(&a.qst[b1 + e[j]]);
#else
    for (int j = 0; j < GPL; j++) qs1[j] = make_double2(1.0 + 1e-9 * (double)b1, 0.5);
#endif

    const T len = (T)L0;
#pragma unroll
    for (int j = 0; j < GPL; j++) {
      const T tau = (T)qs0[j].y * len;        /* sigma_t * length */
      const T lq = len * (T)qs0[j].x;          /* length * q */
      /* ExpEvaluator::computeExponential (src/ExpEvaluator.h:170-183) for the NP polar
       * angles in lock-step: NP independent Horner chains keep the FP64 pipe fed
       * instead of one dependent chain after the other */
      T x[NP], f1[NP];
#pragma unroll
      for (int p = 0; p < NP; p++) x[p] = tau * inv_sin[p];
      expF1_batch<T, NP>(x, f1, a.cf);
#if B200_ACC_DIRECT
      /* fsr_flux[e] += weight * delta_psi, polar angle after polar angle as the reference does
       * (CPUSolver.cpp:2463-2470): NP DFMA on the accumulator, no separate sum */
#pragma unroll
      for (int p = 0; p < NP; p++) {
        const T ex = inv_sin[p] * f1[p];
        const T dpsi = (tau * (T)psi[p][j] - lq) * ex;
        psi[p][j] = (float)((T)psi[p][j] - dpsi);
        if constexpr (sizeof(T) == 8) acc[j] = fma((double)w[p], (double)dpsi, acc[j]);
        else acc[j] += (double)(w[p] * dpsi);
      }
#else
      T sum = (T)0;
#pragma unroll
      for (int p = 0; p < NP; p++) {
        const T ex = inv_sin[p] * f1[p];
        const T dpsi = (tau * (T)psi[p][j] - lq) * ex;
#ifdef B200_EXP_PSID
        psi[p][j] = psi[p][j] - dpsi;
#else
        psi[p][j] = (float)((T)psi[p][j] - dpsi);
#endif
        sum = fma(w[p], dpsi, sum);
      }
      acc[j] += (double)sum;
#endif
    }

    if constexpr (CMFD) {
      /* outgoing flux of this segment across a CMFD cell surface (src/Cmfd.h:572-670) */
      const int2 c = *pc;
      pc += step;
      const int surf = dir ? c.y : c.x;
      if (surf >= 0) {
#pragma unroll
        for (int j = 0; j < GPL; j++) {
          double cur = 0.0;
#pragma unroll
          for (int p = 0; p < NP; p++) cur = fma((double)w[p], (double)psi[p][j], cur);
          if (valid[j]) atomicAdd(&a.currents[(size_t)surf * a.ncg + cg[j]], cur);
        }
      }
    }

    /* flush before the FSR changes */
    {
      const bool flush = b1 != b0;
#pragma unroll
      for (int j = 0; j < GPL; j++) {
        if constexpr (DET) {
          if (flush && valid[j])
            atomicAdd(&phi_fx[b0 + et[j]], (unsigned long long)__double2ll_rn(acc[j] * fx_scale));
        } else {
          red_add_if(&phi[b0 + et[j]], acc[j], flush && valid[j]);   /* predicated RED, no branch */
        }
        acc[j] = flush ? 0.0 : acc[j];
      }
    }
    L0 = L1; b0 = b1;
    L1 = __hiloint2double(r2.y, r2.x); b1 = (uint32_t)r2.z;
#pragma unroll
    for (int j = 0; j < GPL; j++) qs0[j] = qs1[j];
    ps += step;
  }
  /* the look-ahead may have hidden the last FSR change: flush what is left.  b0 has
   * been rotated once past the last segment, so recompute its slot from the stream. */
  if (n > 0) {
    const uint32_t blast = a.seg[dir ? s0 : s1 - 1].base;
#pragma unroll
    for (int j = 0; j < GPL; j++)
      if (valid[j] && acc[j] != 0.0) {
        if constexpr (DET) atomicAdd(&phi_fx[blast + et[j]], (unsigned long long)__double2ll_rn(acc[j] * fx_scale));
        else atomicAdd(&phi[blast + et[j]], acc[j]);
      }
  }

  /* transferBoundaryFlux (src/CPUSolver.cpp:2560-2601): reflective / periodic
   * ends feed the next track's start flux; vacuum ends just drop it. */
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


/* ------------------------------------------------------------------------------------
 * B200_PRECISION_TABLE: the exponential from a shared-memory interpolation table.
 *
 * The reference's ExpEvaluator was built around a table of F1 with linear interpolation
 * (src/ExpEvaluator.cpp:190-330, now dead code behind the rational); the flat 2D sweep here is bound
 * by FP64 issue - 15 of its 22 FP64 instructions per integration are the rational and its quotient.
 * This optional mode trades them for one LDS.128 and two FFMA: F1 is a quadratic Taylor expansion
 * about the midpoint of intervals of width 0.02 over [0, 40) (fp32 {a, b, c}, 2 000 entries, 32 KB of
 * shared memory per CTA; truncation error < 4e-8, fp32 rounding 6e-8) and 1/x beyond.  Everything
 * else - tau, the source term, the update of psi and the tally - stays in double.  Held to the
 * north-star tolerance only (k_eff within 1 pcm, fluxes within 1e-4); the default stays bit-faithful.
 * ------------------------------------------------------------------------------------ */
constexpr int F1TAB_N = 2000;
constexpr float F1TAB_INV_H = 50.0f;           /* 1 / 0.02 */
constexpr float F1TAB_MAX = 40.0f;

template <int NP>
__global__ void __launch_bounds__(B200_LB_THREADS, B200_LB_BLOCKS)
sweep_kernel_tab(const SweepArgs a, const float4* __restrict__ f1tab) {
  __shared__ float4 tab[F1TAB_N];
  for (int i = threadIdx.x; i < F1TAB_N; i += blockDim.x) tab[i] = f1tab[i];
  __syncthreads();
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
  const uint32_t e = (uint32_t)sub;              /* one group per thread: lpi == G */

  double w[NP];
  float inv_sin[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    w[p] = a.cls_w[cls * NP + p];
    inv_sin[p] = (float)a.cls_inv_sin[cls * NP + p];
  }
  const int F = G * NP;
  const int64_t slot_in = (t * 2 + dir) * (int64_t)F;
  float psi[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) psi[p] = a.psi_in[slot_in + p * G + e];
  if (a.carry[t * 2 + dir]) {
#pragma unroll
    for (int p = 0; p < NP; p++) a.psi_out[slot_in + p * G + e] = psi[p];
  }
  double acc = 0.0;

  const int step = dir ? -1 : 1;
  const SegRec* __restrict__ ps = a.seg + (dir ? s1 - 1 : s0);
  const int4 r0 = ld_rec(ps);
  const int4 r1 = ld_rec(ps + step);
  double L0 = __hiloint2double(r0.y, r0.x), L1 = __hiloint2double(r1.y, r1.x);
  uint32_t b0 = (uint32_t)r0.z, b1 = (uint32_t)r1.z;
  double2 qs0 = ld_qs(&a.qst[b0 + e]), qs1;
  ps += 2 * step;
  const uint32_t et = e + (uint32_t)(blockIdx.x & a.rep_mask) * (uint32_t)a.rep_stride;
  double* __restrict__ const phi = a.phi;

  for (int i = 0; i < n; i++) {
    const int4 r2 = ld_rec(ps);
    qs1 = ld_qs(&a.qst[b1 + e]);
    const double tau = qs0.y * L0;
    const double lq = L0 * qs0.x;
    const float tau32 = (float)tau;
#pragma unroll
    for (int p = 0; p < NP; p++) {
      /* ExpEvaluator::computeExponential (src/ExpEvaluator.h:170-183): inv_sin * F1(tau * inv_sin) */
      const float x = tau32 * inv_sin[p];
      const float xi = fminf(x, F1TAB_MAX - 0.001f) * F1TAB_INV_H;
      const int k = (int)xi;
      const float4 c = tab[k];
      const float d = (xi - (float)k - 0.5f) * (1.0f / F1TAB_INV_H);       /* distance from the interval's midpoint */
      float f1 = fmaf(fmaf(c.z, d, c.y), d, c.x);
      f1 = x >= F1TAB_MAX ? __frcp_rn(x) : f1;
      const double ex = (double)(inv_sin[p] * f1);
      const double dpsi = (tau * (double)psi[p] - lq) * ex;
      psi[p] = (float)((double)psi[p] - dpsi);
      acc = fma(w[p], dpsi, acc);
    }
    const bool flush = b1 != b0;
    red_add_if(&phi[b0 + et], acc, flush);
    acc = flush ? 0.0 : acc;
    L0 = L1; b0 = b1;
    L1 = __hiloint2double(r2.y, r2.x); b1 = (uint32_t)r2.z;
    qs0 = qs1;
    ps += step;
  }
  if (n > 0 && acc != 0.0) atomicAdd(&phi[a.seg[dir ? s0 : s1 - 1].base + et], acc);

  const int64_t out = a.out_slot[t * 2 + dir];
  if (out >= 0) {
    const int peer = (int)(out >> PEER_SHIFT);
    float* __restrict__ dst = peer ? a.peer_out.p[peer - 1] : a.psi_out;
    const int64_t base = (out & PEER_SLOT_MASK) * (int64_t)F;
#pragma unroll
    for (int p = 0; p < NP; p++) dst[base + p * G + e] = psi[p];
  } else if (a.leakage != nullptr) {
    double lk = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++) lk += (double)psi[p];
    atomicAdd(&a.leakage[t], (float)(a.cls_w[cls * NP] * lk));
  }
}

/* ---- padded FSR rows ------------------------------------------------------------------
 * With G = 7 a row of {q, sigma_t} pairs is 112 bytes and a tally row 56 bytes: laid end to end, 87 %
 * of the gathers and 43 % of the tallies of a 7-lane item straddle two 128-byte lines, and every line a
 * warp instruction touches is one more pass through the L1 data stage (the 3D sweep keeps that stage
 * 81 % busy, profiles/r02_sweep3d.md).  The RED-bound sweeps (3D, linear source) therefore work on
 * private copies whose rows are padded to GP = 8 groups (128-byte / 64-byte aligned): pack_pad_kernel
 * copies the sources in before the sweep, unpack_pad_kernel sums the tally replicas back into the
 * solver's [r*G + e] arrays afterwards (and clears them for the next sweep). */
__global__ void pack_pad_kernel(const double2* __restrict__ qst, double2* __restrict__ qst_pad,
                                const double4* __restrict__ qxyz, double4* __restrict__ qxyz_pad,
                                int64_t n_fsr, int G, int GP, const int* __restrict__ done) {
  if (done != nullptr && *done) return;
  const int64_t n = n_fsr * GP;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / GP;
    const int e = (int)(i - r * GP);
    if (e < G) {
      qst_pad[i] = qst[r * G + e];
      if (qxyz != nullptr) qxyz_pad[i] = qxyz[r * G + e];
    }
  }
}
/* dst[plane][r*G + e] = sum over replicas of pad[rep][plane][r*GP + e]; the padded copies are cleared */
__global__ void unpack_pad_kernel(double* __restrict__ pad, double* __restrict__ dst, int64_t n_fsr, int G, int GP,
                                  int n_planes, int n_rep, const int* __restrict__ done) {
  if (done != nullptr && *done) return;
  const int64_t nphi = n_fsr * G, nphi_pad = n_fsr * GP, n = nphi * n_planes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / nphi, k = i - c * nphi, r = k / G;
    const int e = (int)(k - r * G);
    const int64_t j = c * nphi_pad + r * GP + e;
    double sum = 0.0;
    for (int rep = 0; rep < n_rep; rep++) {
      sum += pad[(int64_t)rep * n_planes * nphi_pad + j];
      pad[(int64_t)rep * n_planes * nphi_pad + j] = 0.0;
    }
    dst[i] = sum;
  }
}

/* builds the padded SegRec stream from the uploaded SoA arrays (once per upload) */
__global__ void build_segrec_kernel(SegRec* __restrict__ out, const double* __restrict__ len,
                                    const int32_t* __restrict__ fsr, int64_t n_seg, int G) {
  const int64_t total = n_seg + 2 * SEG_PAD;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i - SEG_PAD;
    SegRec r;
    if (s >= 0 && s < n_seg) { r.len = len[s]; r.base = (uint32_t)fsr[s] * (uint32_t)G; }
    else { r.len = 0.0; r.base = 0u; }
    r.spare = 0u;
    out[i] = r;
  }
}

/* padded {fwd, bwd} CMFD surface stream */
__global__ void build_segcmfd_kernel(int2* __restrict__ out, const int32_t* __restrict__ fwd,
                                     const int32_t* __restrict__ bwd, int64_t n_seg) {
  const int64_t total = n_seg + 2 * SEG_PAD;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i - SEG_PAD;
    out[i] = (s >= 0 && s < n_seg) ? make_int2(fwd[s], bwd[s]) : make_int2(-1, -1);
  }
}

}  // namespace b200
