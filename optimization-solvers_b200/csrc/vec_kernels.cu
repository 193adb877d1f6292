// vec_kernels.cu — O(n) streaming kernels of the path: every one is a single fused
// map(+reduce) launch over the vectors (reduce.cuh), no temporaries.
//   x + t*d, projection           ls_solver.rs:60; number.rs:13-21; backtracking_b.rs:65-67
//   projected directions          projected_gradient_descent.rs:56-59; spg.rs:81-84; bfgs_b.rs:72-75
//   convergence scalars           bfgs.rs:74; gradient_descent.rs:46-53; ls_solver.rs:121-133 + number.rs:27-31
//   s, y, norms, dots             bfgs.rs:96-99,115; spg.rs:129-141
//   MoreThuenteB t_max scan       morethuente_b.rs:185-197
// Elementwise arithmetic is bit-exact w.r.t. the reference (separate roundings, NaN-dropping
// max/min): the active set is defined by exact == on these values.
#include "engine.cuh"

namespace osb {

void vec_axpy_project(Ctx* ctx, int64_t n, const double* x, const double* d, double t, bool project, const double* lb,
                      const double* ub, double* out, double* d_dn) {
  auto f = [=] __device__(int64_t i, double(&acc)[1]) {
    const double xi = x[i];
    const double td = t * d[i];
    double v = xi + td;
    if (project) v = fmin(fmax(v, lb[i]), ub[i]);
    out[i] = v;
    const double df = v - xi;
    acc[0] = acc[0] + df * df;
  };
  if (d_dn == nullptr) d_dn = ctx->d_dummy;
  launch_mapreduce<1>(ctx, f, n, RedOps<1>{{RED_SUM}}, d_dn);
}

void vec_dot(Ctx* ctx, int64_t n, const double* a, const double* b, double* d_out) {
  auto f = [=] __device__(int64_t i, double(&acc)[1]) { acc[0] = acc[0] + a[i] * b[i]; };
  launch_mapreduce<1>(ctx, f, n, RedOps<1>{{RED_SUM}}, d_out);
}

__global__ void project_kernel(int64_t n, double* x, const double* lb, const double* ub) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = fmin(fmax(x[i], lb[i]), ub[i]);
}
void vec_project_inplace(Ctx* ctx, int64_t n, double* x, const double* lb, const double* ub) {
  project_kernel<<<ctx->red_grid(n), RED_THREADS, 0, ctx->stream>>>(n, x, lb, ub);
  ctx->counters[0]++;
}

// d = -a ; gd0 = g . d
void vec_neg(Ctx* ctx, int64_t n, const double* a, double* out, const double* g, double* d_gd0) {
  auto f = [=] __device__(int64_t i, double(&acc)[1]) {
    const double di = -a[i];
    out[i] = di;
    acc[0] = acc[0] + g[i] * di;
  };
  launch_mapreduce<1>(ctx, f, n, RedOps<1>{{RED_SUM}}, d_gd0);
}

// d = P(x - [lam *] w) - x ; out3 = {g.d, min feasible step, ||d||_inf}
void vec_projected_direction(Ctx* ctx, int64_t n, const double* x, const double* w, double lam, bool scale, const double* lb,
                             const double* ub, const double* g, double* d, double* d_out3) {
  auto f = [=] __device__(int64_t i, double(&acc)[3]) {
    const double xi = x[i];
    const double wi = scale ? lam * w[i] : w[i];
    double v = xi - wi;
    v = fmin(fmax(v, lb[i]), ub[i]);
    const double di = v - xi;
    d[i] = di;
    acc[0] = acc[0] + g[i] * di;
    double cand;  // morethuente_b.rs:185-197
    if (di > 0.0) cand = (ub[i] - xi) / di;
    else if (di < 0.0) cand = (lb[i] - xi) / di;
    else cand = INFINITY;
    acc[1] = fmin(cand, acc[1]);
    acc[2] = fmax(acc[2], fabs(di));
  };
  launch_mapreduce<3>(ctx, f, n, RedOps<3>{{RED_SUM, RED_MIN, RED_MAX}}, d_out3);
}

void vec_tmax_candidate(Ctx* ctx, int64_t n, const double* x, const double* d, const double* lb, const double* ub,
                        double* d_out) {
  auto f = [=] __device__(int64_t i, double(&acc)[1]) {
    const double di = d[i], xi = x[i];
    double cand;
    if (di > 0.0) cand = (ub[i] - xi) / di;
    else if (di < 0.0) cand = (lb[i] - xi) / di;
    else cand = INFINITY;
    acc[0] = fmin(cand, acc[0]);
  };
  launch_mapreduce<1>(ctx, f, n, RedOps<1>{{RED_MIN}}, d_out);
}

void vec_conv_gnorm2(Ctx* ctx, int64_t n, const double* g, double* d_out) {
  auto f = [=] __device__(int64_t i, double(&acc)[1]) { acc[0] = acc[0] + g[i] * g[i]; };
  launch_mapreduce<1>(ctx, f, n, RedOps<1>{{RED_SUM}}, d_out);
}
// gradient_descent.rs:46-53: fold from -inf with NaN-dropping max
void vec_conv_gmax(Ctx* ctx, int64_t n, const double* g, double* d_out) {
  auto f = [=] __device__(int64_t i, double(&acc)[1]) { acc[0] = fmax(fabs(g[i]), acc[0]); };
  launch_mapreduce<1>(ctx, f, n, RedOps<1>{{RED_MAX}}, d_out);
}
// ls_solver.rs:121-133 + number.rs:27-31 (host applies the fold's 0.0 seed)
void vec_conv_pginf(Ctx* ctx, int64_t n, const double* x, const double* g, const double* lb, const double* ub, double* d_out) {
  auto f = [=] __device__(int64_t i, double(&acc)[1]) {
    const double xi = x[i];
    double pg = g[i];
    if ((xi == lb[i] && pg > 0.0) || (xi == ub[i] && pg < 0.0)) pg = 0.0;
    acc[0] = fmax(acc[0], fabs(pg));
  };
  launch_mapreduce<1>(ctx, f, n, RedOps<1>{{RED_MAX}}, d_out);
}

void vec_sy(Ctx* ctx, int64_t n, const double* xn, const double* x, const double* gn, const double* g, double* s, double* y,
            double* d_out4) {
  auto f = [=] __device__(int64_t i, double(&acc)[4]) {
    const double si = xn[i] - x[i];
    const double yi = gn[i] - g[i];
    s[i] = si;
    y[i] = yi;
    acc[0] = acc[0] + si * si;
    acc[1] = acc[1] + yi * yi;
    acc[2] = acc[2] + yi * si;
  };
  launch_mapreduce<4>(ctx, f, n, RedOps<4>{{RED_SUM, RED_SUM, RED_SUM, RED_SUM}}, d_out4);
}

__global__ void active_set_kernel(int64_t n, const double* x, const double* lb, const double* ub, uint8_t* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (uint8_t)((x[i] == lb[i] ? 1 : 0) | (x[i] == ub[i] ? 2 : 0));
}
void vec_active_set(Ctx* ctx, int64_t n, const double* x, const double* lb, const double* ub, uint8_t* out) {
  active_set_kernel<<<ctx->red_grid(n), RED_THREADS, 0, ctx->stream>>>(n, x, lb, ub, out);
  ctx->counters[0]++;
}

__global__ void finish_sy_kernel(DevState* st, double tol) {
  st->s_norm = sqrt(st->ss);  // bfgs.rs:97
  st->y_norm = sqrt(st->yy);  // bfgs.rs:99
  st->has_s = 1;
  st->has_y = 1;
  st->skip = (st->s_norm < tol || st->y_norm < tol) ? 1 : 0;  // bfgs.rs:106-112
}
void state_finish_sy(Ctx* ctx, DevState* st, double tol) {
  finish_sy_kernel<<<1, 1, 0, ctx->stream>>>(st, tol);
  ctx->counters[0]++;
}

}  // namespace osb
