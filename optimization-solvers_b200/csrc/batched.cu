// batched.cu — batched mode: many independent small dense-BFGS problems, ONE WARP PER PROBLEM.
//
// Each warp runs the complete LineSearchSolver::minimize loop (src/ls_solver.rs:66-111) of
// BFGS (src/quasi_newton/bfgs.rs) + BackTracking (src/line_search/backtracking.rs:19-59) on an
// extended-Rosenbrock problem of dimension n <= 32, in a single launch: lane i owns x_i, g_i, d_i
// and ROW i of the inverse-Hessian approximation in registers; the vectors that every lane needs
// (g or y, s, h) are mirrored in shared memory and read as broadcasts.  HBM traffic is the
// start points in and the results out; the kernel is bound by the FP64 pipe (ncu: 45.7 % of its issue slots busy at 16
// warps per SM; every a * b + c is two separately rounded instructions by design, so the FMA-counted flop fraction is
// 13.6 %).  Forcing 5 or 6 CTAs per SM (96 / 80 registers) is SLOWER: 175 / 199 ms against 163 for 262,144 problems.
//
// Operation order is the oracle's (oracle/oracle.cpp, update form RANK2) statement by statement:
//   * row products  (H v)_i : strict left-to-right sum of separately rounded products (nalgebra gemv);
//   * dot products          : nalgebra's 8-accumulator pattern, replayed with warp shuffles;
//   * objective             : sequential sum over the n/2 Rosenbrock blocks;
//   * update                : H_ij = (H_ij - rho (s_i h_j + h_i s_j)) + c (s_i s_j), no contraction.
// The results are therefore BIT-IDENTICAL to the oracle's rank-2 form (tests/test_gpu_batched.py),
// which makes iteration counts and termination reasons comparable problem by problem even though
// BFGS on Rosenbrock is chaotic in the rounding.
#include "engine.cuh"
#include "functors.cuh"

namespace osb {

constexpr int BB_WARPS = 4;  // warps (problems) per CTA
constexpr int BB_N = 32;

__device__ __forceinline__ double bshfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// nalgebra `dot` (blas.rs dotx) for a vector spread over the lanes: p = a_l * b_l on lane l (0 beyond n)
__device__ __forceinline__ double warp_dot_nalgebra(double p, int n) {
  const int lane = threadIdx.x & 31;
  const int nb8 = n >> 3;
  // acc_k = ((0 + p_k) + p_{k+8}) + ... on lanes k < 8
  double acc = 0.0;
  for (int m = 0; m < nb8; ++m) {
    const double v = bshfl(p, (lane & 7) + 8 * m);
    acc = acc + v;
  }
  double res = 0.0;
  const double a0 = bshfl(acc, 0), a1 = bshfl(acc, 1), a2 = bshfl(acc, 2), a3 = bshfl(acc, 3);
  const double a4 = bshfl(acc, 4), a5 = bshfl(acc, 5), a6 = bshfl(acc, 6), a7 = bshfl(acc, 7);
  res = res + (a0 + a4);
  res = res + (a1 + a5);
  res = res + (a2 + a6);
  res = res + (a3 + a7);
  for (int i = nb8 * 8; i < n; ++i) res = res + bshfl(p, i);
  return res;
}

// Rosenbrock at the point held one coordinate per lane: returns f (all lanes), writes this lane's gradient entry
__device__ __forceinline__ double warp_rosenbrock(double xi, int n, double& gi) {
  const int lane = threadIdx.x & 31;
  const double other = __shfl_xor_sync(0xffffffffu, xi, 1);
  const bool even = (lane & 1) == 0;
  const double a = even ? xi : other, b = even ? other : xi;
  const double t1 = b - a * a;
  const double t2 = 1.0 - a;
  gi = even ? (-400.0 * (a * t1) - 2.0 * t2) : (200.0 * t1);
  const double fb = 100.0 * (t1 * t1) + t2 * t2;
  double f = 0.0;
  for (int pidx = 0; pidx + 1 < n; pidx += 2) f = f + bshfl(fb, pidx);
  if (lane >= n) gi = 0.0;
  return f;
}

__global__ void __launch_bounds__(BB_WARPS * 32)
batched_bfgs_kernel(int n, int64_t np, const double* __restrict__ x0, int generated, int64_t problem0, double tol, int max_iter,
                    int max_ls, double c1, double beta, double* __restrict__ x_out, double* __restrict__ f_out,
                    int32_t* __restrict__ k_out, int32_t* __restrict__ st_out, int32_t* __restrict__ reason_out) {
  __shared__ double sh_v[BB_WARPS][BB_N];  // g or y
  __shared__ double sh_s[BB_WARPS][BB_N];
  __shared__ double sh_h[BB_WARPS][BB_N];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t prob = (int64_t)blockIdx.x * BB_WARPS + w;
  if (prob >= np) return;
  const bool act = lane < n;
  double* vv = sh_v[w];
  double* vs = sh_s[w];
  double* vh = sh_h[w];
  // ---- BFGS::new (bfgs.rs:27-40): H = I
  double Hrow[BB_N];
#pragma unroll
  for (int j = 0; j < BB_N; ++j) Hrow[j] = (j == lane && act) ? 1.0 : 0.0;
  double x;
  if (generated) x = act ? ((lane & 1) ? 1.0 : -1.2) + (double)h16(3, (uint64_t)(problem0 + prob), (uint64_t)lane) * 1.52587890625e-05 : 0.0;
  else x = act ? x0[prob * n + lane] : 0.0;
  double g;
  double f = warp_rosenbrock(x, n, g);
  int k = 0, status = OSB_MAX_ITER_REACHED, reason = OSB_REASON_NONE;
  bool has_sy = false;
  double s_norm = 0.0, y_norm = 0.0;
  while (max_iter > k) {  // ls_solver.rs:78
    if (is_bad(f)) {      // ls_solver.rs:37-40
      status = OSB_OUT_OF_DOMAIN;
      break;
    }
    // has_converged (bfgs.rs:64-76)
    if (has_sy && s_norm < tol) {
      status = OSB_OK;
      reason = OSB_REASON_S_NORM;
      break;
    }
    if (has_sy && y_norm < tol) {
      status = OSB_OK;
      reason = OSB_REASON_Y_NORM;
      break;
    }
    if (sqrt(warp_dot_nalgebra(g * g, n)) < tol) {
      status = OSB_OK;
      reason = OSB_REASON_GRAD_TOL;
      break;
    }
    // compute_direction (bfgs.rs:47): d = -(H g), row i = strict left-to-right sum
    __syncwarp();
    vv[lane] = g;
    __syncwarp();
    double hg = Hrow[0] * vv[0];
#pragma unroll
    for (int j = 1; j < BB_N; ++j)
      if (j < n) hg = Hrow[j] * vv[j] + hg;
    const double d = act ? -hg : 0.0;
    // BackTracking::compute_step_len (backtracking.rs:19-59)
    const double gd0 = warp_dot_nalgebra(g * d, n);
    double t = 1.0;
    int i = 0;
    while (max_ls > i) {
      const double td = t * d;
      const double xn = x + td;
      double gn;
      const double fn = warp_rosenbrock(xn, n, gn);
      if (is_bad(fn)) {  // :37-41
        t *= beta;
        continue;
      }
      if (fn - f <= c1 * t * gd0) break;  // line_search/mod.rs:35
      t *= beta;
      i += 1;
    }
    // update_next_iterate (bfgs.rs:86-112)
    const double td = t * d;
    const double xn = x + td;
    double gn;
    const double fn = warp_rosenbrock(xn, n, gn);  // bfgs.rs:98 (same bits as the accepted trial)
    const double s = xn - x;
    const double y = gn - g;
    s_norm = sqrt(warp_dot_nalgebra(s * s, n));
    y_norm = sqrt(warp_dot_nalgebra(y * y, n));
    has_sy = true;
    x = xn;
    g = gn;
    f = fn;
    if (!(s_norm < tol) && !(y_norm < tol)) {
      // rank-2 form of bfgs.rs:115-124 (oracle RANK2): h = H y, rho = 1/(y.s), c = rho^2 (y.h) + rho
      __syncwarp();
      vv[lane] = y;
      vs[lane] = s;
      __syncwarp();
      double h = Hrow[0] * vv[0];
#pragma unroll
      for (int j = 1; j < BB_N; ++j)
        if (j < n) h = Hrow[j] * vv[j] + h;
      if (!act) h = 0.0;
      vh[lane] = h;
      __syncwarp();
      const double ys = warp_dot_nalgebra(y * s, n);
      const double rho = 1.0 / ys;
      const double yh = warp_dot_nalgebra(y * h, n);
      const double c = rho * rho * yh + rho;
#pragma unroll
      for (int j = 0; j < BB_N; ++j) {
        if (j < n && act) {
          const double cross = s * vh[j] + h * vs[j];
          const double ssq = s * vs[j];
          Hrow[j] = (Hrow[j] - rho * cross) + c * ssq;
        }
      }
    }
    k += 1;  // ls_solver.rs:104
  }
  if (act) x_out[prob * n + lane] = x;
  if (lane == 0) {
    f_out[prob] = f;
    k_out[prob] = k;
    st_out[prob] = status;
    reason_out[prob] = reason;
  }
}

int batched_bfgs_rosenbrock(Ctx* ctx, int64_t n, int64_t np, const double* x0_host, bool generated, int64_t problem0, double tol,
                            int64_t max_iter, int64_t max_ls, double c1, double beta, double* x_out, double* f_out, int32_t* k_out,
                            int32_t* st_out, int32_t* reason_out, double* ms_out) {
  OSB_REQUIRE(n >= 2 && n <= BB_N && n % 2 == 0, OSB_ERROR_INPUT_PARAMS, "batched BFGS supports even n in [2, 32]");
  OSB_REQUIRE(np >= 1, OSB_ERROR_INPUT_PARAMS, "n_problems >= 1");
  cudaStream_t stm = ctx->stream;
  DBuf dx0, dx(np * n), df(np);
  int32_t *dk = nullptr, *dst = nullptr, *dre = nullptr;
  OSB_CUDA(cudaMalloc(&dk, sizeof(int32_t) * np));
  OSB_CUDA(cudaMalloc(&dst, sizeof(int32_t) * np));
  OSB_CUDA(cudaMalloc(&dre, sizeof(int32_t) * np));
  if (!generated) {
    dx0.alloc(np * n);
    dx0.upload(x0_host, np * n, stm);
  }
  cudaEvent_t e0, e1;
  OSB_CUDA(cudaEventCreate(&e0));
  OSB_CUDA(cudaEventCreate(&e1));
  const unsigned grid = (unsigned)((np + BB_WARPS - 1) / BB_WARPS);
  OSB_CUDA(cudaEventRecord(e0, stm));
  batched_bfgs_kernel<<<grid, BB_WARPS * 32, 0, stm>>>((int)n, np, dx0.p, generated ? 1 : 0, problem0, tol, (int)max_iter, (int)max_ls,
                                                       c1, beta, dx.p, df.p, dk, dst, dre);
  ctx->counters[0]++;
  OSB_CUDA(cudaEventRecord(e1, stm));
  OSB_CUDA(cudaGetLastError());
  if (x_out) dx.download(x_out, np * n, stm);
  if (f_out) df.download(f_out, np, stm);
  if (k_out) OSB_CUDA(cudaMemcpyAsync(k_out, dk, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, stm));
  if (st_out) OSB_CUDA(cudaMemcpyAsync(st_out, dst, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, stm));
  if (reason_out) OSB_CUDA(cudaMemcpyAsync(reason_out, dre, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, stm));
  ctx->sync();
  float ms = 0.f;
  OSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  if (ms_out) *ms_out = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dk);
  cudaFree(dst);
  cudaFree(dre);
  return OSB_OK;
}

}  // namespace osb
