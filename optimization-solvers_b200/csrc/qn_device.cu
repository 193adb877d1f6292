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

void qn_device_launch_head(Ctx* ctx, int functor_kind, const double* fn_a, const double* fn_b, bool bounded, LSParams* d_ls,
                           int64_t n, double tol, int64_t max_ls, DevState* st, double* x, double* g, double* d, double* xt,
                           double* gt, double* s, double* y, const double* u, const double* lb, const double* ub,
                           const double* ls_lb, const double* ls_ub) {
  if (functor_kind == FN_ROSENBROCK) {
    launch_head(ctx, RosenbrockFn{}, bounded, d_ls, n, tol, max_ls, st, x, g, d, xt, gt, s, y, u, lb, ub, ls_lb, ls_ub);
  } else if (functor_kind == FN_SEPQUAD) {
    SepQuadFn fn;
    fn.c = fn_a;
    fn.a = fn_b;
    launch_head(ctx, fn, bounded, d_ls, n, tol, max_ls, st, x, g, d, xt, gt, s, y, u, lb, ub, ls_lb, ls_ub);
  } else {
    throw Error(OSB_ERR_UNSUPPORTED, "objective has no block functor for the device-resident engine");
  }
}

}  // namespace osb
