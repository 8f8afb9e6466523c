/* Factory over the reference's own (unmodified) GPUSolver, compiled for sm_100a into
 * oracle/_ref/libopenmoc_refgpu.so.  TEST / BENCH INFRASTRUCTURE ONLY: ref_driver dlopens
 * it for `--solver refgpu` so that the reference's Kepler-era CUDA kernels, recompiled for
 * the B200, can be timed beside B200Solver on the same tracks. */
#include "accel/cuda/GPUSolver.h"

extern "C" Solver* make_ref_gpu_solver(TrackGenerator* tg, int blocks, int threads) {
  GPUSolver* s = new GPUSolver(tg);
  if (blocks > 0) s->setNumThreadBlocks(blocks);
  if (threads > 0) s->setNumThreadsPerBlock(threads);
  /* Solver::computeEigenvalue (Solver.cpp:1567-1568) calls initializeMaterials() before
   * initializeFSRs(), and GPUSolver::initializeMaterials (GPUSolver.cu:1262) rejects the
   * still-zero group count.  The public initializeSolver() (Solver.cpp:1074-1081) runs the
   * same steps in the other order, so calling it first sets the counts - no source change. */
  s->initializeSolver(FORWARD);
  return s;
}
