// Microbenchmark: cost of the FSR tally on B200.
//   mode 0: per-lane RED.ADD.F64, groups of 7 consecutive doubles (what the sweep does)
//   mode 1: same, groups padded/aligned to 8 doubles (64 B rows)
//   mode 2: TMA bulk reduction cp.reduce.async.bulk ... .add.f64 of one 64-byte row per group
//   mode 3: per-lane RED.ADD.F64, 32 consecutive doubles per warp
// Every "op" adds 7 (or 8/32) doubles into a table of n_rows rows chosen pseudo-randomly
// with locality (neighbouring groups hit the same or adjacent rows, like adjacent tracks).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int MODE>
__global__ void bench(double* __restrict__ table, int n_rows, int iters, int stride) {
  extern __shared__ __align__(128) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
  if (MODE == 0 || MODE == 1) {
    const int grp = lane / 7, sub = lane % 7;
    if (grp >= 4) return;
    uint32_t row = hash32(gwarp * 4 + grp) % n_rows;
    for (int i = 0; i < iters; i++) {
      row = (row + 1 + (hash32(row + i) & 1)) % n_rows;   // walk through neighbouring rows
      atomicAdd(&table[(size_t)row * stride + sub], 1.0 + sub);
    }
  } else if (MODE == 3) {
    uint32_t row = hash32(gwarp) % (n_rows / 4);
    for (int i = 0; i < iters; i++) {
      row = (row + 1 + (hash32(row + i) & 1)) % (n_rows / 4);
      atomicAdd(&table[(size_t)row * 32 + lane], 1.0 + lane);
    }
  } else {
    // 4 groups per warp; each group stages 8 doubles in its own double-buffered 64-byte slot
    const int grp = lane >> 3, sub = lane & 7;
    double* slot = sm + ((warp * 4 + grp) * 2) * 8;
    uint32_t row = hash32(gwarp * 4 + grp) % n_rows;
    for (int i = 0; i < iters; i++) {
      row = (row + 1 + (hash32(row + i) & 1)) % n_rows;
      double* buf = slot + (i & 1) * 8;
      // make sure the bulk op that read this buffer two iterations ago is done
      if (sub == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      buf[sub] = (sub < 7) ? 1.0 + sub : 0.0;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (sub == 0) {
        uint32_t s = (uint32_t)__cvta_generic_to_shared(buf);
        double* g = table + (size_t)row * 8;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 64;"
                     ::"l"(g), "r"(s) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (sub == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main(int argc, char** argv) {
  int n_rows = argc > 1 ? atoi(argv[1]) : 23869;
  int iters = argc > 2 ? atoi(argv[2]) : 2000;
  double* table;
  cudaMalloc(&table, (size_t)n_rows * 32 * 8);
  cudaMemset(table, 0, (size_t)n_rows * 32 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = 224, blocks = 148 * 3;
  const int warps = threads / 32;
  for (int mode = 0; mode < 4; mode++) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      size_t smem = (size_t)warps * 4 * 2 * 8 * 8;
      if (mode == 0) bench<0><<<blocks, threads>>>(table, n_rows, iters, 7);
      if (mode == 1) bench<1><<<blocks, threads>>>(table, n_rows, iters, 8);
      if (mode == 2) bench<2><<<blocks, threads, smem>>>(table, n_rows, iters, 8);
      if (mode == 3) bench<3><<<blocks, threads>>>(table, n_rows, iters, 32);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    double groups = (double)blocks * warps * (mode == 3 ? 1 : 4) * iters;
    double cyc_per_warp_op = best * 1e-3 * 1.965e9 / ((double)blocks * warps * iters / 148.0);
    printf("mode %d: %.3f ms  %.3e row-adds/s  %.1f SM-cycles per warp-op  (%s)\n", mode, best,
           groups / (best * 1e-3), cyc_per_warp_op, cudaGetErrorString(err));
  }
  // checksum so the work is not optimised away
  double h[8]; cudaMemcpy(h, table, 64, cudaMemcpyDeviceToHost); printf("table[0..2]=%g %g %g\n", h[0], h[1], h[2]);
  return 0;
}
