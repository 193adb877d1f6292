// reduce.cuh — deterministic map-reduce over a vector index space.
//
// Every O(n) scalar of the path (dot products of line_search/mod.rs:35,47,55, norms of
// bfgs.rs:97,99, the inf-norm of number.rs:27-31, MoreThuenteB's t_max scan morethuente_b.rs:185-197)
// is produced by one kernel: grid-stride map, per-thread sequential accumulation, warp-shuffle tree,
// fixed-order cross-warp sum, one partial per CTA, and the LAST CTA (atomic ticket) folds the
// partials in CTA order and writes the K results.  The order depends only on (count, grid), and
// the grid is a pure function of count, so results are run-to-run and rank-to-rank identical.
#pragma once
#include "common.cuh"

namespace osb {

enum RedOp : int { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };

template <int K>
struct RedOps {
  int op[K];
};

HD double red_identity(int op) { return op == RED_SUM ? 0.0 : (op == RED_MAX ? -INFINITY : INFINITY); }
HD double red_combine(int op, double a, double b) { return op == RED_SUM ? a + b : (op == RED_MAX ? fmax(a, b) : fmin(a, b)); }

constexpr int RED_THREADS = 256;
constexpr int64_t RED_SEQ_MAX = 7;

#ifdef __CUDACC__
__device__ __forceinline__ double warp_red(int op, double v) {
  return op == RED_SUM ? warp_sum(v) : (op == RED_MAX ? warp_max(v) : warp_min(v));
}

// CTA-wide reduction of acc[K]; result valid in ALL threads.  smem must hold K*32 doubles.
template <int K>
__device__ __forceinline__ void cta_reduce(double (&acc)[K], const RedOps<K>& ops, double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  __syncthreads();  // protect smem reuse across consecutive calls
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = warp_red(ops.op[k], acc[k]);
    if (lane == 0) smem[k * 32 + warp] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = red_identity(ops.op[k]);
    for (int w = 0; w < nwarp; ++w) v = red_combine(ops.op[k], v, smem[k * 32 + w]);
    acc[k] = v;
  }
}

// default finalizer: write the K folded values
struct FinStore {
  template <int K>
  __device__ __forceinline__ void operator()(const double (&v)[K], double* out) const {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = v[k];
  }
};

// F:   __device__ void operator()(int64_t i, double (&acc)[K]) — folds item i into acc.
// Fin: __device__ void operator()(const double (&v)[K], double* out) — run by ONE thread on the K folded values.
template <int K, class F, class Fin>
__global__ void __launch_bounds__(RED_THREADS) mapreduce_kernel(F f, Fin fin, int64_t count, RedOps<K> ops, double* partials,
                                                                unsigned int* ticket, double* out) {
  __shared__ double smem[K * 32];
  __shared__ bool is_last;
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = red_identity(ops.op[k]);
  if (count <= RED_SEQ_MAX) {
    // nalgebra's dot for fewer than 8 entries is a plain left-to-right sum: replay it in one thread so
    // that the reference's own tiny tests (n = 2, 3) reproduce bit-for-bit
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      for (int64_t i = 0; i < count; ++i) f(i, acc);
      fin(acc, out);
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) f(i, acc);
  cta_reduce<K>(acc, ops, smem);
  if (gridDim.x == 1) {
    if (threadIdx.x == 0) fin(acc, out);
    return;
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) partials[(int64_t)blockIdx.x * K + k] = acc[k];
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < K) {
    const int op = ops.op[warp];
    double v = red_identity(op);
    for (int c = lane; c < (int)gridDim.x; c += 32) v = red_combine(op, v, __ldcg(&partials[(int64_t)c * K + warp]));
    v = warp_red(op, v);
    if (lane == 0) smem[warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = smem[k];
    fin(v, out);
    *ticket = 0u;
  }
}
#endif

}  // namespace osb
