// objectives.cu — device objectives: the counterpart of the user closure
// `FnMut(&DVector<f64>) -> FuncEvalMultivariate` (src/ls_solver.rs:34, src/func_eval.rs:5-41).
// Built-ins: dense quadratic (examples/quadratic.rs:10-14 pattern), extended Rosenbrock,
// separable box quadratic; plus host-closure and user-device-functor adapters.
#include "engine.cuh"
#include "functors.cuh"

namespace osb {

// default trial: x + t d [projected] -> eval -> g.d      (three launches; built-ins fuse them)
void Objective::trial(const double* x, const double* d, double t, bool project, const double* lb, const double* ub, double* xt,
                      double* gt, double* d_out3) {
  vec_axpy_project(ctx, n, x, d, t, project, lb, ub, xt, d_out3 + 2);
  eval(xt, d_out3, gt, nullptr);
  vec_dot(ctx, n, gt, d, d_out3 + 1);
}

// ---- block-functor objectives: one fused kernel per evaluation / per line-search trial -----
template <class Fn>
struct FunctorObjective : Objective {
  Fn fn;
  int fkind;
  DBuf pa, pb;
  FunctorObjective(Ctx* c, int64_t n_, int kind) : Objective(c, n_), fkind(kind) {}
  int functor_kind() const override { return fkind; }
  const double* functor_ptr(int i) const override { return i == 0 ? pa.p : pb.p; }

  void eval(const double* x, double* d_f, double* g, double* hess) override {
    OSB_REQUIRE(hess == nullptr, OSB_PANIC_NO_HESSIAN, "Hessian not available in the oracle");
    calls++;
    ctx->counters[1]++;
    const Fn f_ = fn;
    auto f = [=] __device__(int64_t b, double(&acc)[1]) {
      double xb[Fn::BS], gb[Fn::BS];
      const int64_t i0 = b * Fn::BS;
#pragma unroll
      for (int j = 0; j < Fn::BS; ++j) xb[j] = x[i0 + j];
      const double fb = f_.block(i0, xb, gb);
#pragma unroll
      for (int j = 0; j < Fn::BS; ++j) g[i0 + j] = gb[j];
      acc[0] = acc[0] + fb;
    };
    launch_mapreduce<1>(ctx, f, n / Fn::BS, RedOps<1>{{RED_SUM}}, d_f);
  }

  // "one trial step is one kernel": axpy, projection, objective, gradient, g.d and ||xt-x||^2 fused
  void trial(const double* x, const double* d, double t, bool project, const double* lb, const double* ub, double* xt,
             double* gt, double* d_out3) override {
    calls++;
    ctx->counters[1]++;
    const Fn f_ = fn;
    auto f = [=] __device__(int64_t b, double(&acc)[3]) {
      double xb[Fn::BS], gb[Fn::BS], db[Fn::BS];
      const int64_t i0 = b * Fn::BS;
#pragma unroll
      for (int j = 0; j < Fn::BS; ++j) {
        const double xi = x[i0 + j];
        db[j] = d[i0 + j];
        const double td = t * db[j];
        double v = xi + td;
        if (project) v = fmin(fmax(v, lb[i0 + j]), ub[i0 + j]);
        xb[j] = v;
        xt[i0 + j] = v;
        const double df = v - xi;
        acc[2] = acc[2] + df * df;
      }
      const double fb = f_.block(i0, xb, gb);
#pragma unroll
      for (int j = 0; j < Fn::BS; ++j) {
        gt[i0 + j] = gb[j];
        acc[1] = acc[1] + gb[j] * db[j];
      }
      acc[0] = acc[0] + fb;
    };
    launch_mapreduce<3>(ctx, f, n / Fn::BS, RedOps<3>{{RED_SUM, RED_SUM, RED_SUM}}, d_out3);
  }
  // spg.rs:81-84 / projected_gradient_descent.rs:56-59 (direction), ls_solver.rs:60 + backtracking_b.rs:65-67 (trial point),
  // spg.rs:129-141 (s.s, s.y), ls_solver.rs:121-133 + number.rs:27-31 (projected gradient), morethuente_b.rs:185-197
  // (feasible step): the arithmetic of each coordinate is the one of vec_projected_direction + trial + vec_sy, in the
  // same order, so the iterates (and the active set, defined by exact ==) do not change — only the number of passes:
  // 4 reads + 2 writes per trial instead of ~14 vector passes per iteration.
  bool has_stream_trial() const override { return true; }
  // One work item = VW coordinates (whole functor blocks), moved with 128-bit loads and stores: 4 vectors x 32 B of loads
  // in flight per thread.  (With one coordinate per work item the kernel had 32 KB in flight per SM and ran at half the
  // HBM rate: 4.1 ms per trial at n = 2^28 for 12.9 GB.)
  template <int VW>
  void stream_trial_vw(const double* x, const double* g, const double* lb, const double* ub, double lam, bool scale, double t, bool project,
                       const double* ls_lb, const double* ls_ub, double* xt, double* gt, double* d_out8) {
    static_assert(VW % Fn::BS == 0 && (VW == 1 || VW % 2 == 0), "whole blocks, whole 16-byte vectors");
    const Fn f_ = fn;
    auto f = [=] __device__(int64_t w, double(&acc)[8]) {
      const int64_t i0 = w * VW;
      double xi[VW], gi[VW], lo[VW], hi[VW], xv[VW], gv[VW], di[VW];
      if (VW >= 2) {
#pragma unroll
        for (int j = 0; j < VW; j += 2) {
          const double2 a = *reinterpret_cast<const double2*>(x + i0 + j), b = *reinterpret_cast<const double2*>(g + i0 + j);
          const double2 c = *reinterpret_cast<const double2*>(lb + i0 + j), d = *reinterpret_cast<const double2*>(ub + i0 + j);
          xi[j] = a.x, xi[j + 1] = a.y, gi[j] = b.x, gi[j + 1] = b.y, lo[j] = c.x, lo[j + 1] = c.y, hi[j] = d.x, hi[j + 1] = d.y;
        }
      } else {
        xi[0] = x[i0], gi[0] = g[i0], lo[0] = lb[i0], hi[0] = ub[i0];
      }
#pragma unroll
      for (int j = 0; j < VW; ++j) {
        const double wi = scale ? lam * gi[j] : gi[j];
        double v = xi[j] - wi;
        v = fmin(fmax(v, lo[j]), hi[j]);
        di[j] = v - xi[j];
        acc[4] = acc[4] + gi[j] * di[j];
        double pg = gi[j];
        if ((xi[j] == lo[j] && pg > 0.0) || (xi[j] == hi[j] && pg < 0.0)) pg = 0.0;
        acc[7] = fmax(acc[7], fabs(pg));
        const double td = t * di[j];
        double xn = xi[j] + td;
        if (project) xn = fmin(fmax(xn, ls_lb[i0 + j]), ls_ub[i0 + j]);
        xv[j] = xn;
        const double df = xn - xi[j];
        acc[2] = acc[2] + df * df;
      }
#pragma unroll
      for (int b = 0; b < VW; b += Fn::BS) {
        const double fb = f_.block(i0 + b, xv + b, gv + b);
        acc[0] = acc[0] + fb;
      }
#pragma unroll
      for (int j = 0; j < VW; ++j) {
        acc[1] = acc[1] + gv[j] * di[j];
        const double yi = gv[j] - gi[j], si = xv[j] - xi[j];
        acc[3] = acc[3] + yi * si;
      }
      if (VW >= 2) {
#pragma unroll
        for (int j = 0; j < VW; j += 2) {
          *reinterpret_cast<double2*>(xt + i0 + j) = make_double2(xv[j], xv[j + 1]);
          *reinterpret_cast<double2*>(gt + i0 + j) = make_double2(gv[j], gv[j + 1]);
        }
      } else {
        xt[i0] = xv[0];
        gt[i0] = gv[0];
      }
    };
    // {f_t, g_t.d, ||x_t - x||^2, s.y, g.d, -, -, ||proj grad||_inf}: the feasible-step candidate and ||d||_inf of the
    // separate direction kernel belong to MoreThuenteB / the SPG constructor, which do not come through here
    launch_mapreduce<8>(ctx, f, n / VW, RedOps<8>{{RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_MIN, RED_MAX, RED_MAX}}, d_out8);
  }
  // spg.rs:81-84 / projected_gradient_descent.rs:56-59 (direction), ls_solver.rs:60 + backtracking_b.rs:65-67 (trial point),
  // spg.rs:129-141 (s.s, s.y), ls_solver.rs:121-133 + number.rs:27-31 (projected gradient): per coordinate the arithmetic
  // of vec_projected_direction + trial + vec_sy in the same order, so the iterates (and the active set, defined by exact
  // ==) are the ones of the separate kernels; the dot products are summed in a different grouping (4 coordinates per
  // work item), i.e. equal to rounding.  4 vector reads + 2 writes per trial instead of ~14 vector passes per iteration.
  bool stream_trial(const double* x, const double* g, const double* lb, const double* ub, double lam, bool scale, double t,
                    bool project, const double* ls_lb, const double* ls_ub, double* xt, double* gt, double* d_out8) override {
    calls++;
    ctx->counters[1]++;
    if (n % 4 == 0) stream_trial_vw<4>(x, g, lb, ub, lam, scale, t, project, ls_lb, ls_ub, xt, gt, d_out8);
    else stream_trial_vw<Fn::BS>(x, g, lb, ub, lam, scale, t, project, ls_lb, ls_ub, xt, gt, d_out8);
    return true;
  }
};

Objective* make_rosenbrock(Ctx* ctx, int64_t n) {
  OSB_REQUIRE(n >= 2 && n % 2 == 0, OSB_ERROR_INPUT_PARAMS, "extended Rosenbrock needs an even n >= 2");
  return new FunctorObjective<RosenbrockFn>(ctx, n, FN_ROSENBROCK);
}

// the generated separable problem keeps no coefficient vectors: SepQuadFn recomputes c_i, a_i from the integer hash
// (4 GiB less memory and a third less traffic per trial at n = 2^28); index0 = global index of local coordinate 0
Objective* make_sepquad_generated(Ctx* ctx, int64_t n, int64_t index0) {
  auto* o = new FunctorObjective<SepQuadFn>(ctx, n, FN_SEPQUAD);
  o->fn.c = nullptr;
  o->fn.a = nullptr;
  o->fn.index0 = index0;
  return o;
}

// ---- dense quadratic ----------------------------------------------------------------------
__global__ void gen_quad_kernel(int64_t n, int64_t ld, double sc, double* A, int64_t nrows, int64_t row0) {
  const int64_t total = nrows * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = row0 + e / ld, j = e % ld;
    double v = 0.0;
    if (j < n) {
      if (i == j) v = 2.0 + (double)(i % 7);
      else {
        const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
        v = (double)h16(1, (uint64_t)lo, (uint64_t)hi) * sc;
      }
    }
    A[e] = v;
  }
}
__global__ void scale_copy_kernel(int64_t total, const double* in, double* out, double sc) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) out[e] = sc * in[e];
}

// Multi-GPU: A is row-block sharded like H — rank p owns rows [p n/P, (p+1) n/P) — and x, g are
// replicated; (A x)_p is all-gathered, then f and g are formed redundantly (and identically) on every rank.  A row's
// dot product is formed inside one CTA whatever the sharding, so sharded runs are bit-identical to one GPU.
struct DenseQuadratic : Objective {
  DBuf A, b, Ax;
  int64_t ld, nrows, row0;
  bool shifted = false;
  DenseQuadratic(Ctx* c, int64_t n_) : Objective(c, n_), ld(qn_ld(n_)), nrows(n_), row0(0) {
    OSB_REQUIRE(!c->vec_sharded, OSB_ERR_UNSUPPORTED, "index-range sharding applies to block-functor objectives only");
    if (c->world > 1) {
      OSB_REQUIRE(n_ % (c->world * 8) == 0, OSB_ERROR_INPUT_PARAMS, "row-sharded A needs n divisible by 8 * world");
      nrows = n_ / c->world;
      row0 = nrows * c->rank;
    }
    A.alloc(qn_rows_padded(nrows) * ld);
    A.zero(c->stream);
    Ax.alloc(ld);
    Ax.zero(c->stream);
  }
  bool provides_hessian() const override { return ctx->world == 1; }
  void eval(const double* x, double* d_f, double* g, double* hess) override {
    calls++;
    ctx->counters[1]++;
    OSB_REQUIRE(hess == nullptr || ctx->world == 1, OSB_ERR_UNSUPPORTED, "the Hessian of a row-sharded quadratic is not assembled");
    // one read of A yields A x, hence f and g  (n^2 * 8 B per evaluation, / world when sharded)
    qn_launch_gemv(ctx, A.p, ld, nrows, row0, nullptr, x, Ax.p, nullptr, nullptr, 0);
    if (ctx->world > 1) ctx->all_gather_inplace(Ax.p, nrows);
    const double* ax = Ax.p;
    const double* bb = shifted ? b.p : nullptr;
    auto f = [=] __device__(int64_t i, double(&acc)[2]) {
      const double xi = x[i], axi = ax[i];
      acc[0] = acc[0] + xi * axi;
      if (bb) {
        acc[1] = acc[1] + bb[i] * xi;
        g[i] = 2.0 * (axi - bb[i]);
      } else {
        g[i] = 2.0 * axi;
      }
    };
    auto fin = [=] __device__(const double(&v)[2], double* out) { out[0] = bb ? v[0] - 2.0 * v[1] : v[0]; };
    launch_mapreduce_fin<2>(ctx, f, fin, n, RedOps<2>{{RED_SUM, RED_SUM}}, d_f);
    if (hess) {  // constant Hessian 2A (same padded layout as A)
      scale_copy_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n * ld, A.p, hess, 2.0);
      ctx->counters[0]++;
    }
  }
};

Objective* make_dense_quadratic(Ctx* ctx, int64_t n, const double* A_host, const double* b_host) {
  auto* o = new DenseQuadratic(ctx, n);
  OSB_CUDA(cudaMemcpy2DAsync(o->A.p, o->ld * sizeof(double), A_host + o->row0 * n, n * sizeof(double), n * sizeof(double), o->nrows,
                             cudaMemcpyHostToDevice, ctx->stream));
  if (b_host) {
    o->shifted = true;
    o->b.alloc(o->ld);
    o->b.zero(ctx->stream);
    o->b.upload(b_host, n, ctx->stream);
  }
  ctx->sync();
  return o;
}

Objective* make_dense_quadratic_generated(Ctx* ctx, int64_t n, bool shifted, double* x0_host) {
  auto* o = new DenseQuadratic(ctx, n);
  int lg = 0;
  while (((int64_t)1 << lg) < n) ++lg;
  const double sc = std::ldexp(1.0, -(15 + lg));
  gen_quad_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(n, o->ld, sc, o->A.p, o->nrows, o->row0);
  ctx->counters[0]++;
  if (shifted) {
    std::vector<double> b(n);
    for (int64_t i = 0; i < n; ++i) b[i] = (double)h16(9, (uint64_t)i, 0) * std::ldexp(1.0, -13);
    o->shifted = true;
    o->b.alloc(o->ld);
    o->b.zero(ctx->stream);
    o->b.upload(b.data(), n, ctx->stream);
  }
  if (x0_host)
    for (int64_t i = 0; i < n; ++i) x0_host[i] = (double)h16(2, (uint64_t)i, 0) * std::ldexp(1.0, -13);
  ctx->sync();
  return o;
}

// ---- host closure (compatibility path) ----------------------------------------------------
// Staging buffers are PINNED, so both copies are real asynchronous DMA transfers, and the call returns without waiting
// for the upload: the next call's download of x is ordered after it on the same stream and is followed by the only
// synchronisation of the path, which is also what makes the staging buffers safe to overwrite.
struct HostObjective : Objective {
  osb_host_eval_fn fn;
  void* user;
  bool with_h;
  double* hx = nullptr;  // n
  double* hg = nullptr;  // n + 1: gradient, then f
  double* hh = nullptr;  // n * n when the closure provides the Hessian
  HostObjective(Ctx* c, int64_t n_, osb_host_eval_fn f, void* u, bool wh) : Objective(c, n_), fn(f), user(u), with_h(wh) {
    OSB_CUDA(cudaHostAlloc(&hx, sizeof(double) * (size_t)n_, cudaHostAllocDefault));
    OSB_CUDA(cudaHostAlloc(&hg, sizeof(double) * (size_t)(n_ + 1), cudaHostAllocDefault));
    if (wh) OSB_CUDA(cudaHostAlloc(&hh, sizeof(double) * (size_t)n_ * (size_t)n_, cudaHostAllocDefault));
  }
  ~HostObjective() override {
    cudaStreamSynchronize(ctx->stream);
    cudaFreeHost(hx);
    cudaFreeHost(hg);
    cudaFreeHost(hh);
  }
  bool provides_hessian() const override { return with_h; }
  void eval(const double* x, double* d_f, double* g, double* hess) override {
    calls++;
    ctx->counters[1]++;
    OSB_CUDA(cudaMemcpyAsync(hx, x, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    double f = NAN;
    int got = fn(user, hx, n, &f, hg, (hess && with_h) ? hh : nullptr);
    OSB_REQUIRE(!(hess && !got), OSB_PANIC_NO_HESSIAN, "Hessian not available in the oracle");
    hg[n] = f;
    OSB_CUDA(cudaMemcpyAsync(d_f, hg + n, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    OSB_CUDA(cudaMemcpyAsync(g, hg, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (hess) {
      const int64_t ld = qn_ld(n);
      OSB_CUDA(cudaMemcpy2DAsync(hess, ld * sizeof(double), hh, n * sizeof(double), n * sizeof(double), n,
                                 cudaMemcpyHostToDevice, ctx->stream));
    }
  }
};
Objective* make_host_objective(Ctx* ctx, int64_t n, osb_host_eval_fn fn, void* user, bool with_h) {
  return new HostObjective(ctx, n, fn, user, with_h);
}

// ---- user-supplied device functor ---------------------------------------------------------
struct UserObjective : Objective {
  osb_device_eval_fn fn;
  void* user;
  bool with_h;
  UserObjective(Ctx* c, int64_t n_, osb_device_eval_fn f, void* u, bool wh) : Objective(c, n_), fn(f), user(u), with_h(wh) {}
  bool provides_hessian() const override { return with_h; }
  void eval(const double* x, double* d_f, double* g, double* hess) override {
    calls++;
    ctx->counters[1]++;
    OSB_REQUIRE(!(hess && !with_h), OSB_PANIC_NO_HESSIAN, "Hessian not available in the oracle");
    int rc = fn(user, x, n, d_f, g, hess, (void*)ctx->stream);
    OSB_REQUIRE(rc == 0, OSB_ABNORMAL_TERMINATION, "user device functor returned an error");
  }
};
Objective* make_user_objective(Ctx* ctx, int64_t n, osb_device_eval_fn fn, void* user, bool with_h) {
  return new UserObjective(ctx, n, fn, user, with_h);
}

}  // namespace osb
