// qn_device.cu — the device-resident "head" of one dense quasi-Newton outer iteration.
//
// One single-CTA kernel performs, without any host round trip, everything of
// LineSearchSolver::minimize (src/ls_solver.rs:78-107) that is not an O(n^2) pass over H:
//   evaluate_x_k            cached (f, g) of the accepted point; NaN/inf f -> OutOfDomain   ls_solver.rs:32-42
//   has_converged           s_norm / y_norm / ||g||_2 tests                                 bfgs.rs:64-76
//   compute_direction       d = -u  or  d = P(x - u) - x  with u = H g from the last H pass bfgs.rs:47, bfgs_b.rs:72-75
//   compute_step_len        the complete line search (ls_automaton.cuh), objective evaluated
//                           by the block functor, every thread runs the scalar automaton on
//                           CTA-reduced (hence identical) values                            backtracking.rs, morethuente.rs
//   update_next_iterate     x+ = x + t d, s, y, norms, y.s, skip flag, k += 1               bfgs.rs:94-112
// The O(n) vectors (128 KiB each at n = 16384) live in L2; at that size one CTA is the right
// shape: a trial is ~2 us, whereas a multi-CTA launch per trial costs more in launch latency
// than the arithmetic.  All later launches of the iteration read DevState.done / .skip.
#include <cooperative_groups.h>

#include "engine.cuh"
#include "functors.cuh"

namespace osb {

constexpr int HEAD_T = 1024;

template <class Fn>
__device__ __forceinline__ void head_eval_at(const Fn& fn, int64_t n, const double* __restrict__ x, const double* __restrict__ d,
                                             double t, bool project, const double* __restrict__ plb,
                                             const double* __restrict__ pub, double* __restrict__ xt, double* __restrict__ gt,
                                             double (&acc)[3]) {
  acc[0] = acc[1] = acc[2] = 0.0;
  const int64_t nb = n / Fn::BS;
  for (int64_t b = threadIdx.x; b < nb; b += HEAD_T) {
    double xb[Fn::BS], gb[Fn::BS], db[Fn::BS];
    const int64_t i0 = b * Fn::BS;
#pragma unroll
    for (int j = 0; j < Fn::BS; ++j) {
      const double xi = x[i0 + j];
      db[j] = d[i0 + j];
      const double td = t * db[j];
      double v = xi + td;
      if (project) v = fmin(fmax(v, plb[i0 + j]), pub[i0 + j]);
      xb[j] = v;
      xt[i0 + j] = v;
      const double df = v - xi;
      acc[2] = acc[2] + df * df;
    }
    const double fb = fn.block(i0, xb, gb);
#pragma unroll
    for (int j = 0; j < Fn::BS; ++j) {
      gt[i0 + j] = gb[j];
      acc[1] = fma(gb[j], db[j], acc[1]);
    }
    acc[0] = acc[0] + fb;
  }
}

template <class Fn, bool BOUNDED>
__global__ void __launch_bounds__(HEAD_T, 1)
qn_head_kernel(Fn fn, LSParams* __restrict__ lsp, int64_t n, double tol, int64_t max_ls, DevState* __restrict__ st, double* __restrict__ x,
               double* __restrict__ g, double* __restrict__ d, double* __restrict__ xt, double* __restrict__ gt,
               double* __restrict__ s, double* __restrict__ y, const double* __restrict__ u, const double* __restrict__ lb,
               const double* __restrict__ ub, const double* __restrict__ ls_lb, const double* __restrict__ ls_ub) {
  __shared__ double smem[3 * 32];
  const RedOps<3> sum3{{RED_SUM, RED_SUM, RED_SUM}};
  if (st->done) return;
  const int tid = threadIdx.x;
  // ---- evaluate_x_k: cached
  const double f0 = st->f;
  if (is_bad(f0)) {
    if (tid == 0) {
      st->done = 1;
      st->status = OSB_OUT_OF_DOMAIN;
    }
    return;
  }
  // ---- has_converged (bfgs.rs:64-76) and direction in one sweep
  int why = OSB_REASON_NONE;
  if (st->has_s && st->s_norm < tol) why = OSB_REASON_S_NORM;
  else if (st->has_y && st->y_norm < tol) why = OSB_REASON_Y_NORM;
  double tmaxc = INFINITY;
  double gd0 = 0.0;
  if (why == OSB_REASON_NONE) {
    const bool need_tmax = lsp->kind == LS_MORETHUENTE_B;
    double acc[3] = {0.0, 0.0, 0.0};
    double tm = INFINITY;
    for (int64_t i = tid; i < n; i += HEAD_T) {
      const double gi = g[i], xi = x[i];
      double di;
      if (BOUNDED) di = fmin(fmax(xi - u[i], lb[i]), ub[i]) - xi;
      else di = -u[i];
      d[i] = di;
      acc[0] = fma(gi, gi, acc[0]);
      acc[1] = fma(gi, di, acc[1]);
      if (need_tmax) {
        double cand;
        if (di > 0.0) cand = (ls_ub[i] - xi) / di;
        else if (di < 0.0) cand = (ls_lb[i] - xi) / di;
        else cand = INFINITY;
        tm = fmin(cand, tm);
      }
    }
    double accs[3] = {acc[0], acc[1], 0.0};
    cta_reduce<3>(accs, sum3, smem);
    if (need_tmax) {
      double mm[1] = {tm};
      cta_reduce<1>(mm, RedOps<1>{{RED_MIN}}, smem);
      tmaxc = mm[0];
    }
    if (sqrt(accs[0]) < tol) why = OSB_REASON_GRAD_TOL;
    gd0 = accs[1];
  }
  if (why != OSB_REASON_NONE) {
    if (tid == 0) {
      st->done = 1;
      st->status = OSB_OK;
      st->reason = why;
    }
    return;
  }
  // ---- compute_step_len: the whole search on device
  LSParams p = *lsp;
  LSMachine m;
  m.begin(p, f0, gd0, max_ls, tmaxc);
  int evals = 0;
  double ft = f0;
  while (!m.done) {
    const double t = m.request(p);
    double acc[3];
    head_eval_at(fn, n, x, d, t, m.wants_projection(p), ls_lb, ls_ub, xt, gt, acc);
    cta_reduce<3>(acc, sum3, smem);
    ft = acc[0];
    m.feed(p, acc[0], acc[1], acc[2]);
    ++evals;
  }
  const double t = m.result;
  // ---- update_next_iterate: next = x + t d; the extra oracle call of bfgs.rs:98 is the cached trial
  if (!m.last_eval_is_result) {
    double acc[3];
    __syncthreads();
    head_eval_at(fn, n, x, d, t, false, ls_lb, ls_ub, xt, gt, acc);
    cta_reduce<3>(acc, sum3, smem);
    ft = acc[0];
    ++evals;
  }
  __syncthreads();  // xt / gt written by other threads' blocks when BS > 1 are read below
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = tid; i < n; i += HEAD_T) {
    const double xn = xt[i], gn = gt[i];
    const double si = xn - x[i];
    const double yi = gn - g[i];
    s[i] = si;
    y[i] = yi;
    x[i] = xn;
    g[i] = gn;
    acc[0] = fma(si, si, acc[0]);
    acc[1] = fma(yi, yi, acc[1]);
    acc[2] = fma(yi, si, acc[2]);
  }
  cta_reduce<3>(acc, sum3, smem);
  if (tid == 0) {
    st->f = ft;
    st->ft = ft;
    st->gd0 = gd0;
    st->ss = acc[0];
    st->yy = acc[1];
    st->ys = acc[2];
    const double sn = sqrt(acc[0]), yn = sqrt(acc[1]);
    st->s_norm = sn;
    st->y_norm = yn;
    st->has_s = 1;
    st->has_y = 1;
    st->skip = (sn < tol || yn < tol) ? 1 : 0;  // bfgs.rs:106-112
    st->t_last = t;
    st->k += 1;
    st->ls_evals += evals;
    *lsp = p;  // GLLQuadratic.f_previous / MoreThuenteB.t_max persist across outer iterations
  }
}

// ---- fast head: n <= HEADF_T * HEADF_EPT -------------------------------------------------------
// x_k is staged once in shared memory (128 KiB at n = 16384) and d_k lives in registers, so a
// line-search trial touches no global memory at all: it is FP64 arithmetic on the block functor
// plus one CTA reduction (~1 us), however many trials the search needs.  Only the accepted point
// is written back (x, g, s, y).
constexpr int HEADF_T = 512;
constexpr int HEADF_EPT = 32;  // elements per thread

template <class Fn, bool BOUNDED>
__global__ void __launch_bounds__(HEADF_T, 1)
qn_head_fast_kernel(Fn fn, LSParams* __restrict__ lsp, int64_t n, double tol, int64_t max_ls, DevState* __restrict__ st,
                    double* __restrict__ x, double* __restrict__ g, double* __restrict__ s, double* __restrict__ y,
                    const double* __restrict__ u, const double* __restrict__ lb, const double* __restrict__ ub,
                    const double* __restrict__ ls_lb, const double* __restrict__ ls_ub) {
  constexpr int BS = Fn::BS;
  constexpr int KPT = HEADF_EPT / BS;  // blocks per thread
  extern __shared__ double sx[];       // n doubles: x_k
  __shared__ double smem[4 * 32];
  const RedOps<3> sum3{{RED_SUM, RED_SUM, RED_SUM}};
  const RedOps<4> sum4{{RED_SUM, RED_SUM, RED_SUM, RED_SUM}};
  if (st->done) return;
  const int tid = threadIdx.x;
  const int nb = (int)(n / BS);
  const double f0 = st->f;
  if (is_bad(f0)) {  // ls_solver.rs:37-40
    if (tid == 0) {
      st->done = 1;
      st->status = OSB_OUT_OF_DOMAIN;
    }
    return;
  }
  int why = OSB_REASON_NONE;
  if (st->has_s && st->s_norm < tol) why = OSB_REASON_S_NORM;        // bfgs.rs:67-69
  else if (st->has_y && st->y_norm < tol) why = OSB_REASON_Y_NORM;   // bfgs.rs:70-72
  if (why != OSB_REASON_NONE) {
    if (tid == 0) {
      st->done = 1;
      st->status = OSB_OK;
      st->reason = why;
    }
    return;
  }
  // ---- stage x, form d in registers, ||g||^2 and g.d
  double dreg[KPT][BS];
  const bool need_tmax = lsp->kind == LS_MORETHUENTE_B;
  double tm = INFINITY;
  {
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      const int b = tid + k * HEADF_T;
#pragma unroll
      for (int j = 0; j < BS; ++j) {
        dreg[k][j] = 0.0;
        if (b < nb) {
          const int i = b * BS + j;
          const double xi = x[i], gi = g[i], ui = u[i];
          double di;
          if (BOUNDED) di = fmin(fmax(xi - ui, lb[i]), ub[i]) - xi;  // bfgs_b.rs:72-75
          else di = -ui;                                             // bfgs.rs:47
          sx[i] = xi;
          dreg[k][j] = di;
          acc[0] = acc[0] + gi * gi;
          acc[1] = acc[1] + gi * di;
          if (need_tmax) {  // morethuente_b.rs:185-197
            double cand;
            if (di > 0.0) cand = (ls_ub[i] - xi) / di;
            else if (di < 0.0) cand = (ls_lb[i] - xi) / di;
            else cand = INFINITY;
            tm = fmin(cand, tm);
          }
        }
      }
    }
    cta_reduce<3>(acc, sum3, smem);
    if (sqrt(acc[0]) < tol) {  // bfgs.rs:74
      if (tid == 0) {
        st->done = 1;
        st->status = OSB_OK;
        st->reason = OSB_REASON_GRAD_TOL;
      }
      return;
    }
    // gd0 kept in acc[1]
    tm = need_tmax ? tm : INFINITY;
    double gd0 = acc[1];
    double tmaxc = INFINITY;
    if (need_tmax) {
      double mm[1] = {tm};
      cta_reduce<1>(mm, RedOps<1>{{RED_MIN}}, smem);
      tmaxc = mm[0];
    }
    // ---- compute_step_len: every thread runs the automaton on CTA-reduced (identical) scalars
    LSParams p = *lsp;
    LSMachine m;
    m.begin(p, f0, gd0, max_ls, tmaxc);
    int evals = 0;
    while (!m.done) {
      const double t = m.request(p);
      const bool proj = m.wants_projection(p);
      double a3[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < KPT; ++k) {
        const int b = tid + k * HEADF_T;
        if (b < nb) {
          double xb[BS], gb[BS];
#pragma unroll
          for (int j = 0; j < BS; ++j) {
            const int i = b * BS + j;
            const double xi = sx[i];
            const double td = t * dreg[k][j];
            double v = xi + td;
            if (proj) v = fmin(fmax(v, ls_lb[i]), ls_ub[i]);  // backtracking_b.rs:65-67
            xb[j] = v;
            const double df = v - xi;
            a3[2] = a3[2] + df * df;
          }
          const double fb = fn.block((int64_t)b * BS, xb, gb);
#pragma unroll
          for (int j = 0; j < BS; ++j) a3[1] = a3[1] + gb[j] * dreg[k][j];
          a3[0] = a3[0] + fb;
        }
      }
      cta_reduce<3>(a3, sum3, smem);
      m.feed(p, a3[0], a3[1], a3[2]);
      ++evals;
    }
    const double t = m.result;
    // ---- update_next_iterate: next = x + t d (ls_solver.rs:60); oracle(next) (bfgs.rs:98); s, y, norms
    double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      const int b = tid + k * HEADF_T;
      if (b < nb) {
        double xb[BS], gb[BS];
#pragma unroll
        for (int j = 0; j < BS; ++j) {
          const double td = t * dreg[k][j];
          xb[j] = sx[b * BS + j] + td;
        }
        const double fb = fn.block((int64_t)b * BS, xb, gb);
        a4[3] = a4[3] + fb;
#pragma unroll
        for (int j = 0; j < BS; ++j) {
          const int i = b * BS + j;
          const double si = xb[j] - sx[i];
          const double yi = gb[j] - g[i];
          s[i] = si;
          y[i] = yi;
          x[i] = xb[j];
          g[i] = gb[j];
          a4[0] = a4[0] + si * si;
          a4[1] = a4[1] + yi * yi;
          a4[2] = a4[2] + yi * si;
        }
      }
    }
    cta_reduce<4>(a4, sum4, smem);
    if (tid == 0) {
      st->f = a4[3];
      st->ft = a4[3];
      st->gd0 = gd0;
      st->ss = a4[0];
      st->yy = a4[1];
      st->ys = a4[2];
      const double sn = sqrt(a4[0]), yn = sqrt(a4[1]);
      st->s_norm = sn;
      st->y_norm = yn;
      st->has_s = 1;
      st->has_y = 1;
      st->skip = (sn < tol || yn < tol) ? 1 : 0;  // bfgs.rs:106-112
      st->t_last = t;
      st->k += 1;
      st->ls_evals += evals + 1;
      *lsp = p;
    }
  }
}

template <class Fn>
static bool launch_head_fast(Ctx* ctx, Fn fn, bool bounded, LSParams* d_ls, int64_t n, double tol, int64_t max_ls, DevState* st,
                             double* x, double* g, double* s, double* y, const double* u, const double* lb, const double* ub,
                             const double* ls_lb, const double* ls_ub) {
  if (n > (int64_t)HEADF_T * HEADF_EPT) return false;
  const size_t smem = sizeof(double) * (size_t)n;
  static bool attr_set[2] = {false, false};
  if (bounded) {
    if (!attr_set[1]) {
      OSB_CUDA(cudaFuncSetAttribute(qn_head_fast_kernel<Fn, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set[1] = true;
    }
    qn_head_fast_kernel<Fn, true><<<1, HEADF_T, smem, ctx->stream>>>(fn, d_ls, n, tol, max_ls, st, x, g, s, y, u, lb, ub, ls_lb, ls_ub);
  } else {
    if (!attr_set[0]) {
      OSB_CUDA(cudaFuncSetAttribute(qn_head_fast_kernel<Fn, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set[0] = true;
    }
    qn_head_fast_kernel<Fn, false><<<1, HEADF_T, smem, ctx->stream>>>(fn, d_ls, n, tol, max_ls, st, x, g, s, y, u, lb, ub, ls_lb, ls_ub);
  }
  ctx->counters[0]++;
  return true;
}

// ---- cluster head: the same head on a thread-block cluster of 8 CTAs ---------------------------
// One SM's FP64 pipe bounds a trial of the single-CTA head (~9 us per trial at n = 16384, measured
// with ncu); a cluster of 8 CTAs holds x_k and d_k entirely in registers (EPT elements per
// thread), evaluates a trial in well under a microsecond and reduces across the cluster through
// distributed shared memory (one barrier.cluster per reduction, double-buffered partials, rank-
// ordered sum => every CTA sees the same bits).
namespace cg = cooperative_groups;
constexpr int HC_CTAS = 8;
constexpr int HC_T = 512;

// Cluster-wide sum of K <= 12 accumulators, result in all threads of all CTAs, ~1.5k cycles whatever K:
//   1. warp shuffle trees for the K values (independent chains, interleaved);
//   2. the K x 16 warp partials are folded by K x 16 threads with a 16-lane segmented shuffle tree;
//   3. one barrier.cluster; the K x 8 CTA partials are fetched through DSMEM by K x 8 threads in parallel
//      and folded with an 8-lane segmented tree.
// Every tree has a fixed shape, so all CTAs (and all ranks of a sharded run) obtain identical bits.
// (An earlier variant let every thread fold the 16 warp partials sequentially from shared memory:
//  clock64 stamps showed 17.7k cycles per reduction at K = 12, two thirds of the head kernel.)
// Warp sums of P (a power of two) values with P - 1 + log2(32 / P) double shuffles instead of 5 P (the head's
// reductions were bound by the SM's shuffle throughput: 12 values x 5 steps x 16 warps).  Recursive halving: at
// offset o a lane keeps half of its values and adds the partner's copies of them — the same pairs (l, l ^ o) at
// the same levels as the plain butterfly, so the sums are bit-identical to warp_sum().  On return the lanes
// [i * 32 / P, (i + 1) * 32 / P) all hold the total of value i.
template <int P>
__device__ __forceinline__ double warp_sum_multi(double (&v)[P]) {
  static_assert(P == 1 || P == 2 || P == 4 || P == 8 || P == 16, "P must be a power of two <= 16");
  const int lane = threadIdx.x & 31;
  int o = 16;
#pragma unroll
  for (int hs = P / 2; hs >= 1; hs >>= 1, o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < hs; ++i) {
      const double send = upper ? v[i] : v[i + hs];
      const double keep = upper ? v[i + hs] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  double r = v[0];
#pragma unroll
  for (int oo = 16 / P; oo > 0; oo >>= 1) r = r + __shfl_xor_sync(0xffffffffu, r, oo);
  return r;
}

template <int K>
__device__ __forceinline__ void cluster_sum(double (&acc)[K], double* smem_cta /* K x 16 */, double* part /* 2 x 16 */,
                                            double* res /* 16 */, int& phase) {
  static_assert(K <= 12 && HC_T == 512, "layout below assumes 16 warps and K <= 12");
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int P = K <= 1 ? 1 : K <= 2 ? 2 : K <= 4 ? 4 : K <= 8 ? 8 : 16;
  double vp[P];
#pragma unroll
  for (int k = 0; k < P; ++k) vp[k] = k < K ? acc[k] : 0.0;
  const double wsum = warp_sum_multi<P>(vp);
  __syncthreads();  // smem_cta / res may still be read from the previous call
  if (lane % (32 / P) == 0 && lane / (32 / P) < K) smem_cta[(lane / (32 / P)) * 16 + warp] = wsum;
  __syncthreads();
  double* mine = part + (phase & 1) * 16;
  if (threadIdx.x < ((K * 16 + 31) / 32) * 32) {  // thread = k * 16 + w ; whole warps execute the shuffles
    double v = threadIdx.x < K * 16 ? smem_cta[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o, 16);
    if (threadIdx.x < K * 16 && (threadIdx.x & 15) == 0) mine[threadIdx.x >> 4] = v;
  }
  cluster.sync();
  if (threadIdx.x < ((K * HC_CTAS + 31) / 32) * 32) {  // thread = k * 8 + r
    const int k = threadIdx.x >> 3, r = threadIdx.x & 7;
    double v = threadIdx.x < K * HC_CTAS ? cluster.map_shared_rank(mine, r)[k] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o, 8);
    if (threadIdx.x < K * HC_CTAS && r == 0) res[k] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = res[k];
  phase += 1;
}

__device__ __noinline__ double cluster_min_rt(double v, double* smem_cta, double* part, double* res, double* gat, int* phase) {
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  __syncthreads();
  v = warp_min(v);
  if (lane == 0) smem_cta[warp] = v;
  __syncthreads();
  double m = INFINITY;
#pragma unroll 1
  for (int w = 0; w < nwarp; ++w) m = fmin(m, smem_cta[w]);
  double* mine = part + (*phase & 1) * 16;
  if (threadIdx.x == 0) mine[0] = m;
  cluster.sync();
  if (threadIdx.x < HC_CTAS) gat[threadIdx.x] = cluster.map_shared_rank(mine, threadIdx.x)[0];
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = INFINITY;
#pragma unroll 1
    for (int q = 0; q < HC_CTAS; ++q) r = fmin(r, gat[q]);
    res[0] = r;
  }
  __syncthreads();
  *phase += 1;
  return res[0];
}

// ---- deferred lazy-schedule epilogue on the cluster -------------------------------------------
// y.h, s.g, h.g with element i = gt + k * NT (a mapping that does not depend on the functor's block size, so the
// head and the stand-alone finisher add in the same order), then the coefficients of qn_kernels.cu's
// lazy_epilogue_body.  Returns ca, cb of  u = w + (s ca + h cb); the leader publishes the coefficients.
// h or w of the deferred epilogue: one vector, or (sharded packed storage) the sum of the per-rank slots in rank order —
// every rank adds the same numbers in the same order, so all ranks hold the same bits
__device__ __forceinline__ double epi_ld(const double* __restrict__ base, int nslots, int64_t stride, int64_t i) {
  double v = __ldcg(base + i);
  for (int r = 1; r < nslots; ++r) v = v + __ldcg(base + (int64_t)r * stride + i);
  return v;
}

template <int EPT>
__device__ __forceinline__ void cluster_epilogue_coefs(int kind, int nslots, int64_t slot_stride, const double* __restrict__ hh, const double* s, const double* y,
                                                       const double* g, int64_t n, DevState* st, bool leader, int gt,
                                                       double* smem_cta, double* part, double* res, int& phase, double& ca,
                                                       double& cb) {
  constexpr int NT = HC_CTAS * HC_T;
  double hv[EPT], gv[EPT], yv[EPT], sv[EPT];
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int64_t i = gt + (int64_t)k * NT;
    const bool ok = i < n;
    hv[k] = ok ? epi_ld(hh, nslots, slot_stride, i) : 0.0;
    gv[k] = ok ? g[i] : 0.0;
    yv[k] = ok ? y[i] : 0.0;
    sv[k] = ok ? s[i] : 0.0;
  }
  double e3[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    e3[0] = fma(yv[k], hv[k], e3[0]);  // y.h
    e3[1] = fma(sv[k], gv[k], e3[1]);  // s.g
    e3[2] = fma(hv[k], gv[k], e3[2]);  // h.g
  }
  cluster_sum<3>(e3, smem_cta, part, res, phase);
  const double yh = e3[0], sg = e3[1], hg = e3[2], ys = st->ys;
  double c0, c1, c2;
  if (kind == QN_BFGS) {
    const double rho = 1.0 / ys;
    c0 = rho * rho * yh + rho;
    c1 = -rho;
    c2 = 0.0;
  } else {  // DFP
    c0 = 1.0 / ys;
    c1 = 0.0;
    c2 = -1.0 / yh;
  }
  ca = c0 * sg + c1 * hg;
  cb = c1 * sg + c2 * hg;
  if (leader) {
    st->yh = yh;
    st->c0 = c0;
    st->c1 = c1;
    st->c2 = c2;
    st->pc0 = c0;
    st->pc1 = c1;
    st->pc2 = c2;
    st->pending = 1;
    st->epi = 0;
  }
}

// stand-alone finisher: runs the owed epilogue when no head follows (end of minimize(), before a host callback)
__global__ void __cluster_dims__(HC_CTAS, 1, 1) __launch_bounds__(HC_T, 1)
qn_epilogue_cluster_kernel(HeadEpi e, int64_t n, DevState* __restrict__ st, const double* s, const double* y, const double* g, int force) {
  constexpr int EPT = 4;
  constexpr int NT = HC_CTAS * HC_T;
  __shared__ double smem_cta[3 * 16];
  __shared__ double part[32];
  __shared__ double res[16];
  // force: the caller has just completed h, w itself (NCCL all-gather) and nobody raised st->epi
  const int epi = st->epi ? st->epi : ((force && !st->done) ? 1 : 0);
  if (!epi) return;
  cg::cluster_group cluster = cg::this_cluster();
  const int gt = (int)cluster.block_rank() * HC_T + threadIdx.x;
  const bool leader = gt == 0;
  const double* hh = epi == 2 ? e.h2 : e.h;
  const double* ww = epi == 2 ? e.w2 : e.w;
  const bool skip = st->skip != 0;
  int phase = 0;
  double ca = 0.0, cb = 0.0;
  if (!skip) cluster_epilogue_coefs<EPT>(e.kind, e.nslots, e.slot_stride, hh, s, y, g, n, st, leader, gt, smem_cta, part, res, phase, ca, cb);
  else if (leader) {  // bfgs.rs:106-112: no new update; the stored matrix is exact once the pending one is applied
    st->pending = 0;
    st->pc0 = st->pc1 = st->pc2 = 0.0;
    st->epi = 0;
  }
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int64_t i = gt + (int64_t)k * NT;
    if (i < n) {
      const double wi = epi_ld(ww, e.nslots, e.slot_stride, i);
      if (skip) {
        e.u_out[i] = wi;
      } else {
        const double si = s[i], hi = epi_ld(hh, e.nslots, e.slot_stride, i);
        e.u_out[i] = wi + (si * ca + hi * cb);
        e.ps_out[i] = si;
        e.ph_out[i] = hi;
      }
    }
  }
  cluster.sync();  // keep every CTA's shared memory alive until all peers have read it
}

template <class Fn, bool BOUNDED, int EPT, int LSK>
__global__ void __cluster_dims__(HC_CTAS, 1, 1) __launch_bounds__(HC_T, 1)
qn_head_cluster_kernel(Fn fn, LSParams* __restrict__ lsp, int64_t n, double tol, int64_t max_ls, DevState* __restrict__ st,
                       double* __restrict__ x, double* __restrict__ g, double* __restrict__ s, double* __restrict__ y,
                       const double* __restrict__ u, const double* __restrict__ lb, const double* __restrict__ ub,
                       const double* __restrict__ ls_lb, const double* __restrict__ ls_ub, int spec_on, HeadEpi epi) {
  pdl_wait();  // (launched with programmatic stream serialization: the previous kernel's results are complete from here on)
  pdl_launch_dependents();  // (8 CTAs)
  constexpr int BS = Fn::BS;
  constexpr int KPT = EPT / BS;
  constexpr int NT = HC_CTAS * HC_T;
  constexpr int SPEC = 4;  // backtracking trials evaluated per reduction round (speculatively)
  constexpr bool IS_BT = LSK == LS_BACKTRACKING || LSK == LS_BACKTRACKING_B;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double smem_cta[3 * SPEC * 32];
  __shared__ double part[32];
  __shared__ double res[16];
  __shared__ double gat[16 * HC_CTAS];
  if (st->done) return;
  const int gt = (int)cluster.block_rank() * HC_T + threadIdx.x;
  const bool leader = gt == 0;
  const int nb = (int)(n / BS);
  const double f0 = st->f;
  if (is_bad(f0)) {  // ls_solver.rs:37-40
    if (leader) {
      st->done = 1;
      st->status = OSB_OUT_OF_DOMAIN;
    }
    return;
  }
  int phase = 0;
  // the lazy pass of the previous iteration left h = H y, w = H g and owes their O(n) epilogue (st->epi): do it here,
  // on 8 SMs, before anything can terminate the run (the update it defines belongs to the iteration that has ended)
  const int epi_flag = epi.kind >= 0 ? st->epi : 0;
  const bool epi_skip = st->skip != 0;
  const double* __restrict__ epi_h = epi_flag == 2 ? epi.h2 : epi.h;
  const double* __restrict__ epi_w = epi_flag == 2 ? epi.w2 : epi.w;
  double epi_ca = 0.0, epi_cb = 0.0;
  if (epi_flag) {
    // (skip: the leader clears st->epi / pending only AFTER a cluster barrier below — every CTA reads st->epi above, and
    //  a CTA arriving late must not see the flag already cleared)
    if (!epi_skip) cluster_epilogue_coefs<EPT>(epi.kind, epi.nslots, epi.slot_stride, epi_h, s, y, g, n, st, leader, gt, smem_cta, part, res, phase, epi_ca, epi_cb);
  }
  int why = OSB_REASON_NONE;
  if (st->has_s && st->s_norm < tol) why = OSB_REASON_S_NORM;        // bfgs.rs:67-69
  else if (st->has_y && st->y_norm < tol) why = OSB_REASON_Y_NORM;   // bfgs.rs:70-72
  double xreg[KPT][BS], dreg[KPT][BS], greg[KPT][BS];
  constexpr bool need_tmax = LSK == LS_MORETHUENTE_B;
  double tm = INFINITY;
  double acc[4];
  acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
#pragma unroll
  for (int k = 0; k < KPT; ++k) {
    const int b = gt + k * NT;
#pragma unroll
    for (int j = 0; j < BS; ++j) {
      xreg[k][j] = dreg[k][j] = greg[k][j] = 0.0;
      if (b < nb) {
        const int i = b * BS + j;
        const double xi = x[i], gi = g[i];
        double ui;
        if (epi_flag) {  // u = H+ g = w + (s ca + h cb); the vectors of the now pending update are saved
          const double wi = epi_ld(epi_w, epi.nslots, epi.slot_stride, i);
          if (epi_skip) {
            ui = wi;
          } else {
            const double si = s[i], hi = epi_ld(epi_h, epi.nslots, epi.slot_stride, i);
            ui = wi + (si * epi_ca + hi * epi_cb);
            epi.ps_out[i] = si;
            epi.ph_out[i] = hi;
          }
          epi.u_out[i] = ui;
        } else {
          ui = u[i];
        }
        double di;
        if (BOUNDED) di = fmin(fmax(xi - ui, lb[i]), ub[i]) - xi;  // bfgs_b.rs:72-75
        else di = -ui;                                             // bfgs.rs:47
        xreg[k][j] = xi;
        dreg[k][j] = di;
        greg[k][j] = gi;
        acc[0] = acc[0] + gi * gi;
        acc[1] = acc[1] + gi * di;
        if (need_tmax) {  // morethuente_b.rs:185-197
          double cand;
          if (di > 0.0) cand = (ls_ub[i] - xi) / di;
          else if (di < 0.0) cand = (ls_lb[i] - xi) / di;
          else cand = INFINITY;
          tm = fmin(cand, tm);
        }
      }
    }
  }
  if (why != OSB_REASON_NONE) {
    if (leader) {
      st->done = 1;
      st->status = OSB_OK;
      st->reason = why;
    }
    cluster.sync();  // peers may still be reading this CTA's partials of the epilogue reduction
    if (leader && epi_flag && epi_skip) {  // bfgs.rs:106-112: no new update; the stored matrix is exact once the pending one is applied
      st->pending = 0;
      st->pc0 = st->pc1 = st->pc2 = 0.0;
      st->epi = 0;
    }
    return;
  }
  {
    double a2[2] = {acc[0], acc[1]};
    cluster_sum<2>(a2, smem_cta, part, res, phase);
    acc[0] = a2[0];
    acc[1] = a2[1];
  }
  if (leader && epi_flag && epi_skip) {  // (unreachable today: skip implies a norm below tol, i.e. termination above)
    st->pending = 0;
    st->pc0 = st->pc1 = st->pc2 = 0.0;
    st->epi = 0;
  }
  if (sqrt(acc[0]) < tol) {  // bfgs.rs:74
    if (leader) {
      st->done = 1;
      st->status = OSB_OK;
      st->reason = OSB_REASON_GRAD_TOL;
    }
    cluster.sync();  // peers may still be reading this CTA's partials
    return;
  }
  const double gd0 = acc[1];
  double tmaxc = INFINITY;
  if (need_tmax) tmaxc = cluster_min_rt(tm, smem_cta, part, res, gat, &phase);
  LSParams p = *lsp;
  LSMachine m;
  m.template begin<LSK>(p, f0, gd0, max_ls, tmaxc);
  int evals = 0;
  // Backtracking visits t, t*beta, t*beta^2, ... whatever the outcome of a trial (backtracking.rs:37-55
  // multiplies by beta on both the NaN and the rejection branch), so up to SPEC consecutive trials are
  // evaluated in one sweep and ONE cluster reduction, then fed to the automaton in order.  The objective
  // has no side effects: the extra evaluations change nothing but the latency.
  (void)spec_on;
  constexpr int NSPEC = IS_BT ? SPEC : 1;  // everything below is unrolled over NSPEC: arrays stay in registers
  while (!m.done) {                        // (local memory misses L1 after every barrier.cluster, which invalidates it)
    constexpr bool proj = LSK == LS_BACKTRACKING_B;
    double ts[NSPEC];
    ts[0] = m.request(p);
#pragma unroll
    for (int q = 1; q < NSPEC; ++q) ts[q] = ts[q - 1] * p.beta;
    double aS[3 * NSPEC];
#pragma unroll
    for (int q = 0; q < 3 * NSPEC; ++q) aS[q] = 0.0;
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      const int b = gt + k * NT;
      if (b < nb) {
#pragma unroll
        for (int q = 0; q < NSPEC; ++q) {
          double xb[BS], gb[BS];
#pragma unroll
          for (int j = 0; j < BS; ++j) {
            const double td = ts[q] * dreg[k][j];
            double v = xreg[k][j] + td;
            if (proj) v = fmin(fmax(v, ls_lb[b * BS + j]), ls_ub[b * BS + j]);  // backtracking_b.rs:65-67
            xb[j] = v;
            const double df = v - xreg[k][j];
            aS[3 * q + 2] = aS[3 * q + 2] + df * df;
          }
          const double fb = fn.block((int64_t)b * BS, xb, gb);
#pragma unroll
          for (int j = 0; j < BS; ++j) aS[3 * q + 1] = aS[3 * q + 1] + gb[j] * dreg[k][j];
          aS[3 * q] = aS[3 * q] + fb;
        }
      }
    }
    cluster_sum<3 * NSPEC>(aS, smem_cta, part, res, phase);
#pragma unroll
    for (int q = 0; q < NSPEC; ++q) {
      if (!m.done && m.request(p) == ts[q]) {
        m.template feed<LSK>(p, aS[3 * q], aS[3 * q + 1], aS[3 * q + 2]);
        ++evals;
      }
    }
  }
  const double t = m.result;
  // ---- next = x + t d (ls_solver.rs:60); oracle(next) (bfgs.rs:98); s, y, norms, y.s
  acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
#pragma unroll
  for (int k = 0; k < KPT; ++k) {
    const int b = gt + k * NT;
    if (b < nb) {
      double xb[BS], gb[BS];
#pragma unroll
      for (int j = 0; j < BS; ++j) {
        const double td = t * dreg[k][j];
        xb[j] = xreg[k][j] + td;
      }
      const double fb = fn.block((int64_t)b * BS, xb, gb);
      acc[3] = acc[3] + fb;
#pragma unroll
      for (int j = 0; j < BS; ++j) {
        const int i = b * BS + j;
        const double si = xb[j] - xreg[k][j];
        const double yi = gb[j] - greg[k][j];
        s[i] = si;
        y[i] = yi;
        x[i] = xb[j];
        g[i] = gb[j];
        acc[0] = acc[0] + si * si;
        acc[1] = acc[1] + yi * yi;
        acc[2] = acc[2] + yi * si;
      }
    }
  }
  cluster_sum<4>(acc, smem_cta, part, res, phase);
  if (leader) {
    st->f = acc[3];
    st->ft = acc[3];
    st->gd0 = gd0;
    st->ss = acc[0];
    st->yy = acc[1];
    st->ys = acc[2];
    const double sn = sqrt(acc[0]), yn = sqrt(acc[1]);
    st->s_norm = sn;
    st->y_norm = yn;
    st->has_s = 1;
    st->has_y = 1;
    st->skip = (sn < tol || yn < tol) ? 1 : 0;  // bfgs.rs:106-112
    st->t_last = t;
    st->k += 1;
    st->ls_evals += evals + 1;
    // only GLLQuadratic.f_previous and MoreThuenteB.t_max persist across outer iterations
    if (LSK == LS_GLL || LSK == LS_MORETHUENTE_B) *lsp = p;
  }
  cluster.sync();  // keep every CTA's shared memory alive until all peers have read it
}

template <class Fn, int LSK>
static void launch_head_cluster_k(Ctx* ctx, Fn fn, bool bounded, LSParams* d_ls, int64_t n, double tol, int64_t max_ls, DevState* st,
                                  double* x, double* g, double* s, double* y, const double* u, const double* lb, const double* ub,
                                  const double* ls_lb, const double* ls_ub, int spec_on, const HeadEpi& epi) {
  if (bounded)
    launch_pdl(qn_head_cluster_kernel<Fn, true, 4, LSK>, dim3(HC_CTAS), dim3(HC_T), 0, ctx->stream, fn, d_ls, n, tol, max_ls, st, x, g, s, y, u, lb, ub,
               ls_lb, ls_ub, spec_on, epi);
  else
    launch_pdl(qn_head_cluster_kernel<Fn, false, 4, LSK>, dim3(HC_CTAS), dim3(HC_T), 0, ctx->stream, fn, d_ls, n, tol, max_ls, st, x, g, s, y, u, lb, ub,
               ls_lb, ls_ub, spec_on, epi);
}

template <class Fn>
static bool launch_head_cluster(Ctx* ctx, Fn fn, bool bounded, LSParams* d_ls, int ls_kind, int64_t n, double tol, int64_t max_ls,
                                DevState* st, double* x, double* g, double* s, double* y, const double* u, const double* lb,
                                const double* ub, const double* ls_lb, const double* ls_ub, int spec_on, const HeadEpi& epi) {
  constexpr int NT = HC_CTAS * HC_T;
  if (n > (int64_t)NT * 4) return false;
#define OSB_HC(K) launch_head_cluster_k<Fn, K>(ctx, fn, bounded, d_ls, n, tol, max_ls, st, x, g, s, y, u, lb, ub, ls_lb, ls_ub, spec_on, epi)
  switch (ls_kind) {
    case LS_BACKTRACKING: OSB_HC(LS_BACKTRACKING); break;
    case LS_BACKTRACKING_B: OSB_HC(LS_BACKTRACKING_B); break;
    case LS_MORETHUENTE: OSB_HC(LS_MORETHUENTE); break;
    case LS_MORETHUENTE_B: OSB_HC(LS_MORETHUENTE_B); break;
    case LS_GLL: OSB_HC(LS_GLL); break;
    default: OSB_HC(LS_NOSEARCH); break;
  }
#undef OSB_HC
  ctx->counters[0]++;
  return true;
}

template <class Fn>
static void launch_head(Ctx* ctx, Fn fn, bool bounded, LSParams* d_ls, int64_t n, double tol, int64_t max_ls, DevState* st,
                        double* x, double* g, double* d, double* xt, double* gt, double* s, double* y, const double* u,
                        const double* lb, const double* ub, const double* ls_lb, const double* ls_ub) {
  if (bounded)
    qn_head_kernel<Fn, true><<<1, HEAD_T, 0, ctx->stream>>>(fn, d_ls, n, tol, max_ls, st, x, g, d, xt, gt, s, y, u, lb, ub, ls_lb, ls_ub);
  else
    qn_head_kernel<Fn, false><<<1, HEAD_T, 0, ctx->stream>>>(fn, d_ls, n, tol, max_ls, st, x, g, d, xt, gt, s, y, u, lb, ub, ls_lb, ls_ub);
  ctx->counters[0]++;
}

bool qn_device_head_is_cluster(int functor_kind, int64_t n, int head_variant) {
  return (functor_kind == FN_ROSENBROCK || functor_kind == FN_SEPQUAD) && (head_variant == 0 || head_variant == 3) &&
         n <= (int64_t)HC_CTAS * HC_T * 4;
}
void qn_launch_epilogue_cluster(Ctx* ctx, const HeadEpi& e, int64_t n, DevState* st, const double* s, const double* y, const double* g,
                                bool force) {
  qn_epilogue_cluster_kernel<<<HC_CTAS, HC_T, 0, ctx->stream>>>(e, n, st, s, y, g, force ? 1 : 0);
  ctx->counters[0]++;
}

void qn_device_launch_head(Ctx* ctx, int functor_kind, const double* fn_a, const double* fn_b, bool bounded, LSParams* d_ls,
                           int64_t n, double tol, int64_t max_ls, DevState* st, double* x, double* g, double* d, double* xt,
                           double* gt, double* s, double* y, const double* u, const double* lb, const double* ub,
                           const double* ls_lb, const double* ls_ub, int head_variant, int ls_kind, const HeadEpi& epi) {
  if (functor_kind == FN_ROSENBROCK) {
    if ((head_variant == 0 || head_variant == 3) && launch_head_cluster(ctx, RosenbrockFn{}, bounded, d_ls, ls_kind, n, tol, max_ls, st, x, g, s, y, u, lb, ub, ls_lb, ls_ub, head_variant == 0, epi)) return;
    if (head_variant <= 1 && launch_head_fast(ctx, RosenbrockFn{}, bounded, d_ls, n, tol, max_ls, st, x, g, s, y, u, lb, ub, ls_lb, ls_ub)) return;
    launch_head(ctx, RosenbrockFn{}, bounded, d_ls, n, tol, max_ls, st, x, g, d, xt, gt, s, y, u, lb, ub, ls_lb, ls_ub);
  } else if (functor_kind == FN_SEPQUAD) {
    SepQuadFn fn;
    fn.c = fn_a;
    fn.a = fn_b;
    fn.index0 = 0;
    if ((head_variant == 0 || head_variant == 3) && launch_head_cluster(ctx, fn, bounded, d_ls, ls_kind, n, tol, max_ls, st, x, g, s, y, u, lb, ub, ls_lb, ls_ub, head_variant == 0, epi)) return;
    if (head_variant <= 1 && launch_head_fast(ctx, fn, bounded, d_ls, n, tol, max_ls, st, x, g, s, y, u, lb, ub, ls_lb, ls_ub)) return;
    launch_head(ctx, fn, bounded, d_ls, n, tol, max_ls, st, x, g, d, xt, gt, s, y, u, lb, ub, ls_lb, ls_ub);
  } else {
    throw Error(OSB_ERR_UNSUPPORTED, "objective has no block functor for the device-resident engine");
  }
}

}  // namespace osb
