/*
 * group.cuh - several GPUs (or several shards on one GPU) behind ONE b200_solver handle.
 *
 * b200_set_devices() turns a freshly created solver into a group: the uploads are kept on the
 * host, b200_finalize() shards the tracks by whole chains (connected components of the boundary
 * hand-off graph, so no angular flux ever crosses a shard: the role of the reference's
 * azimuthal / domain decomposition, src/CPUSolver.cpp:545-1211, without its interface exchange)
 * and builds one complete single-device solver per shard with replicated FSR, material and
 * quadrature data.  Every entry point of the C ABI then forwards to the shards:
 *   replicated steps (sources, closure, k_eff, normalisation, residual, ...) run on every shard and
 *     stay bit-identical, because they see bit-identical inputs;
 *   the transport sweep runs on every shard's tracks, then the shards' tallies (scalar flux, flux
 *     moments, CMFD currents, fixed-point tally) are summed ACROSS DEVICES BY OUR OWN KERNELS over
 *     peer memory (no NCCL, no host copy): a two-shot all-reduce - every shard sums its slice of all
 *     tallies in a fixed order (reduce-scatter through P2P loads), then gathers the other slices -
 *     ordered by CUDA events between the shards' streams.  This replaces the MPI reductions of
 *     src/CPUSolver.cpp:1900, 2224, 2317.
 * The reference-facing plug-in (B200Solver : Solver) therefore gets multi-GPU, with CMFD and the
 * linear source, from one call: B200Solver::setNumDevices().
 */
#pragma once
#include <cstdint>
#include <numeric>
#include <vector>
#include <cuda_runtime.h>

namespace b200 {

/* ---- chain partition (host) ---- */
struct UnionFind {
  std::vector<int64_t> p;
  explicit UnionFind(int64_t n) : p(n) { std::iota(p.begin(), p.end(), (int64_t)0); }
  int64_t find(int64_t x) {
    while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
    return x;
  }
  void unite(int64_t a, int64_t b) {
    a = find(a); b = find(b);
    if (a != b) p[std::max(a, b)] = std::min(a, b);
  }
};

/* owner[t] in [0, world): whole chains per owner, balanced by `load`.  Longest-processing-time first
 * for up to 200 k chains, a snake deal in order of decreasing load beyond (same balance to within one
 * chain).  Returns the number of chains, or -1 when there are fewer chains than owners. */
static int64_t partition_chains(int64_t n, const int64_t* next_fwd, const int64_t* next_bwd, const uint8_t* bc_fwd,
                                const uint8_t* bc_bwd, const double* load, int world, std::vector<int32_t>& owner) {
  UnionFind uf(n);
  for (int64_t t = 0; t < n; t++) {
    if ((bc_fwd[t] == 1 || bc_fwd[t] == 2) && next_fwd[t] >= 0 && next_fwd[t] < n) uf.unite(t, next_fwd[t]);
    if ((bc_bwd[t] == 1 || bc_bwd[t] == 2) && next_bwd[t] >= 0 && next_bwd[t] < n) uf.unite(t, next_bwd[t]);
  }
  std::vector<int64_t> label(n, -1), root_label(n, -1);
  int64_t n_comp = 0;
  for (int64_t t = 0; t < n; t++) {
    const int64_t r = uf.find(t);
    if (root_label[r] < 0) root_label[r] = n_comp++;
    label[t] = root_label[r];
  }
  if (n_comp < world) return -1;
  std::vector<double> cl(n_comp, 0.);
  for (int64_t t = 0; t < n; t++) cl[label[t]] += load[t] + 1e-3;
  std::vector<int64_t> order(n_comp);
  std::iota(order.begin(), order.end(), (int64_t)0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return cl[a] > cl[b]; });
  std::vector<int32_t> comp_owner(n_comp);
  if (n_comp <= 200000) {
    std::vector<double> tot(world, 0.);
    for (int64_t c : order) {
      int r = 0;
      for (int k = 1; k < world; k++) if (tot[k] < tot[r]) r = k;
      comp_owner[c] = r;
      tot[r] += cl[c];
    }
  } else {
    for (int64_t i = 0; i < n_comp; i++) {
      const int64_t pos = i % (2 * world);
      comp_owner[order[i]] = (int32_t)(pos < world ? pos : 2 * world - 1 - pos);
    }
  }
  owner.resize(n);
  for (int64_t t = 0; t < n; t++) owner[t] = comp_owner[label[t]];
  return n_comp;
}

/* ---- the all-reduce over peer memory ---- */
constexpr int GROUP_MAX = 16;
struct PeerPtrs { void* p[GROUP_MAX]; };

/* stage 1 (reduce-scatter): this shard sums elements [lo, hi) of all shards' tallies, in shard order,
 * into its own staging buffer (peer loads over NVLink when the shards sit on different GPUs) */
template <typename V>
__global__ void group_reduce_kernel(const PeerPtrs src, int n_shards, V* __restrict__ stage, int64_t lo, int64_t hi,
                                    const int* __restrict__ done) {
  if (*done) return;        /* converged device-side loop: the tallies already hold the final (summed) flux */
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    V sum = __ldcg(reinterpret_cast<const V*>(src.p[0]) + i);
    for (int j = 1; j < n_shards; j++) sum += __ldcg(reinterpret_cast<const V*>(src.p[j]) + i);
    stage[i] = sum;
  }
}
/* stage 2 (all-gather): every shard copies all slices from their owners' staging buffers into its tally */
template <typename V>
__global__ void group_gather_kernel(const PeerPtrs stage, int n_shards, int64_t n, V* __restrict__ dst,
                                    const int* __restrict__ done) {
  if (*done) return;
  const int64_t chunk = (n + n_shards - 1) / n_shards;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int owner = (int)(i / chunk);
    dst[i] = __ldcg(reinterpret_cast<const V*>(stage.p[owner]) + i);
  }
}

}  // namespace b200
