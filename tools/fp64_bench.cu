// FP64 pipe microbenchmark on B200: sustained DFMA rate vs ILP / operand form / occupancy.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int FORM>
__global__ void k(double* out, int iters, double c0, double c1) {
  double a[ILP], b[ILP];
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 1e-3 + i; b[i] = 1.0 + i * 1e-3; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (FORM == 0) a[i] = fma(a[i], b[i], c0);        // 2 register pairs + uniform
        else a[i] = fma(a[i], b[i], b[(i + 1) % ILP]);    // 3 distinct register pairs
      }
  }
  double s = 0; for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int FORM> void run(int warps_per_sm, double* out) {
  int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
  int blocks = 148 * (warps_per_sm / (threads / 32));
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, FORM><<<blocks, threads>>>(out, 16, 1.0000001, 0.5);
  cudaEventRecord(e0);
  k<ILP, FORM><<<blocks, threads>>>(out, iters, 1.0000001, 0.5);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double dfma = (double)blocks * threads * iters * 8.0 * ILP;
  printf("ILP=%d form=%d warps/SM=%2d : %.1f DFMA/clk/SM (at 1.965 GHz)  %.2f TFLOP/s\n", ILP, FORM, warps_per_sm,
         dfma / (ms * 1e-3) / 148 / 1.965e9, 2 * dfma / (ms * 1e-3) / 1e12);
}
int main() {
  double* out; cudaMalloc(&out, 148 * 64 * 32 * 8 * 8);
  run<1, 0>(8, out); run<1, 0>(16, out); run<1, 0>(28, out); run<1, 0>(32, out); run<1, 0>(64, out);
  run<3, 0>(8, out); run<3, 0>(16, out); run<3, 0>(28, out);
  run<4, 0>(32, out); run<8, 0>(32, out);
  run<3, 1>(28, out); run<4, 1>(32, out); run<8, 1>(32, out);
  return 0;
}
