// qn_small.cu — dense quasi-Newton update for n <= 5 in the REFERENCE'S OWN operation order.
//
// For all dimensions <= 5 nalgebra does not call matrixmultiply: every product of
// src/quasi_newton/bfgs.rs:115-124 (dfp.rs:115-120, broyden.rs:115-118, sr1_b.rs:143-146) is a
// column-by-column gemv whose entries are strict left-to-right sums of separately rounded
// products, and `dot` degenerates to a sequential sum.  At that size there is nothing to
// parallelise, so ONE thread replays exactly those operations (the statement-for-statement twin
// of oracle/oracle.cpp `QuasiNewton::update_next_iterate`).  This makes the reference's own unit
// tests and examples (all n = 2 or 3) reproduce bit-for-bit on the device — including the only
// exact assert of the crate, examples/quadratic.rs:43 `assert_eq!(eval.f(), &0.0)`.
#include "engine.cuh"

namespace osb {

constexpr int SN = 5;

struct SMat {
  double a[SN][SN];
};

__device__ static void s_matmul(int n, const SMat& A, const SMat& B, SMat& C) {  // nalgebra gemm -> gemv per column
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) C.a[i][j] = A.a[i][0] * B.a[0][j];
    for (int k = 1; k < n; ++k)
      for (int i = 0; i < n; ++i) C.a[i][j] = A.a[i][k] * B.a[k][j] + C.a[i][j];
  }
}
__device__ static void s_gemv(int n, const SMat& A, const double* x, double* y) {
  for (int i = 0; i < n; ++i) y[i] = A.a[i][0] * x[0];
  for (int k = 1; k < n; ++k)
    for (int i = 0; i < n; ++i) y[i] = A.a[i][k] * x[k] + y[i];
}
__device__ static double s_dot(int n, const double* a, const double* b) {
  double r = 0.0;
  for (int i = 0; i < n; ++i) r += a[i] * b[i];
  return r;
}

__global__ void qn_small_kernel(int kind, int n, int64_t ld, double* __restrict__ Hg, DevState* __restrict__ st,
                                const double* __restrict__ sv, const double* __restrict__ yv, const double* __restrict__ gv,
                                double* __restrict__ u_out) {
  if (st->done) return;
  SMat H;
  double s[SN], y[SN], g[SN], u[SN];
  for (int i = 0; i < n; ++i) {
    s[i] = sv[i];
    y[i] = yv[i];
    g[i] = gv[i];
    for (int j = 0; j < n; ++j) H.a[i][j] = Hg[i * ld + j];
  }
  if (!st->skip) {
    SMat T1, T2, T3;
    if (kind == QN_BFGS) {
      const double ys = s_dot(n, y, s);
      const double rho = 1.0 / ys;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          const double idt = (i == j) ? 1.0 : 0.0;
          T1.a[i][j] = idt - (s[i] * y[j]) * rho;  // left  = I - w_a rho
          T2.a[i][j] = idt - (y[i] * s[j]) * rho;  // right = I - w_b rho
        }
      s_matmul(n, T1, H, T3);
      s_matmul(n, T3, T2, T1);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) H.a[i][j] = T1.a[i][j] + (s[i] * s[j]) * rho;
    } else if (kind == QN_DFP) {
      const double sy = s_dot(n, s, y);
      double hy[SN];
      s_gemv(n, H, y, hy);
      const double yhy = s_dot(n, y, hy);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) T1.a[i][j] = y[i] * y[j];
      s_matmul(n, H, T1, T2);
      s_matmul(n, T2, H, T3);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) H.a[i][j] = H.a[i][j] + ((s[i] * s[j]) / sy - T3.a[i][j] / yhy);
    } else if (kind == QN_BROYDEN) {
      double hy[SN];
      s_gemv(n, H, y, hy);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) T1.a[i][j] = (s[i] - hy[i]) * s[j];
      s_matmul(n, T1, H, T2);
      const double den = s_dot(n, s, y);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) H.a[i][j] = H.a[i][j] + T2.a[i][j] / den;
    } else {  // SR1
      double hy[SN], p[SN];
      s_gemv(n, H, y, hy);
      for (int i = 0; i < n; ++i) p[i] = s[i] - hy[i];
      const double den = s_dot(n, p, y);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) H.a[i][j] = H.a[i][j] + (p[i] * p[j]) / den;
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) Hg[i * ld + j] = H.a[i][j];
  }
  s_gemv(n, H, g, u);  // next direction's u = H g   ((-H) g == -(H g), bfgs.rs:47)
  for (int i = 0; i < n; ++i) u_out[i] = u[i];
}

// u = H g only (first iteration / after set_x)
__global__ void qn_small_gemv_kernel(int n, int64_t ld, const double* __restrict__ Hg, const double* __restrict__ gv,
                                     double* __restrict__ u_out) {
  SMat H;
  double g[SN], u[SN];
  for (int i = 0; i < n; ++i) {
    g[i] = gv[i];
    for (int j = 0; j < n; ++j) H.a[i][j] = Hg[i * ld + j];
  }
  s_gemv(n, H, g, u);
  for (int i = 0; i < n; ++i) u_out[i] = u[i];
}

void qn_small_step(Ctx* ctx, int kind, int64_t n, int64_t ld, double* H, DevState* st, const double* s, const double* y,
                   const double* g, double* u) {
  qn_small_kernel<<<1, 1, 0, ctx->stream>>>(kind, (int)n, ld, H, st, s, y, g, u);
  ctx->counters[0]++;
}
void qn_small_gemv(Ctx* ctx, int64_t n, int64_t ld, const double* H, const double* g, double* u) {
  qn_small_gemv_kernel<<<1, 1, 0, ctx->stream>>>((int)n, ld, H, g, u);
  ctx->counters[0]++;
}

}  // namespace osb
