/*
 * microbench.cuh - the two machine ceilings that bound the sweep kernels besides HBM, measured on
 * the spot (b200_measure_ceilings) so that bench.py can report the roofline of the BINDING
 * resource next to the HBM one:
 *   FP64 pipe     sustained DFMA issue rate with the sweep's own shape (3 independent chains per
 *                 thread, coefficient as a uniform operand, 28 warps per SM)
 *   RED.ADD.F64   fire-and-forget tally rate into an L2-resident table, 7 consecutive doubles per
 *                 item walking through neighbouring rows (what sweep_kernel issues per FSR change)
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

__global__ void mb_fp64_kernel(double* out, int iters, double c0) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1., a2 = a0 + 2.;
  const double b0 = 1.0 + 1e-9 * threadIdx.x, b1 = b0 + 1e-3, b2 = b0 + 2e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      a0 = fma(a0, b0, c0); a1 = fma(a1, b1, c0); a2 = fma(a2, b2, c0);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2;
}

__device__ __forceinline__ uint32_t mb_hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

__global__ void mb_red_kernel(double* __restrict__ table, int n_rows, int iters) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int item = gtid / 7, sub = gtid % 7;
  uint32_t row = mb_hash(item) % n_rows;
  for (int i = 0; i < iters; i++) {
    row = (row + 1 + (mb_hash(row + i) & 1)) % n_rows;
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(table + (size_t)row * 7 + sub), "d"(1.0 + sub) : "memory");
  }
}

}  // namespace b200
