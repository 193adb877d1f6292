// newton.cu — dense SPD factor + solve for the Newton family
// (newton/mod.rs:31-47, projected_newton.rs:70-78, spn.rs:82-89).
//
// Small systems (n <= CHOL_SMALL_MAX) use a single-CTA right-looking Cholesky whose per-element
// operation order equals nalgebra's left-looking Cholesky::new (each entry receives the updates
// k = 0..j-1 in order, `a + (-L_jk) * L_ik`), so the factor is bit-identical to the reference's;
// the forward solve is the same column-oriented axpy sweep.  Larger systems go through the blocked
// path (chol_blocked.cu).
#include "engine.cuh"

namespace osb {

constexpr int CHOL_T = 1024;

__global__ void __launch_bounds__(CHOL_T) chol_small_kernel(int64_t n, int64_t ld, const double* __restrict__ A, double* __restrict__ L,
                                                            int* __restrict__ fail) {
  const int tid = threadIdx.x;
  __shared__ double s_diag;
  __shared__ int s_fail;
  if (tid == 0) s_fail = 0;
  for (int64_t e = tid; e < n * n; e += CHOL_T) {
    const int64_t i = e / n, j = e % n;
    L[i * ld + j] = (j <= i) ? A[i * ld + j] : 0.0;
  }
  __syncthreads();
  for (int64_t j = 0; j < n; ++j) {
    if (tid == 0) {
      const double diag = L[j * ld + j];
      if (diag != 0.0 && diag >= 0.0) {  // try_sqrt: NaN and negatives fail (cholesky.rs)
        s_diag = sqrt(diag);
        L[j * ld + j] = s_diag;
      } else {
        s_fail = 1;
      }
    }
    __syncthreads();
    if (s_fail) {
      if (tid == 0) *fail = 1;
      return;
    }
    const double denom = s_diag;
    for (int64_t i = j + 1 + tid; i < n; i += CHOL_T) L[i * ld + j] = L[i * ld + j] / denom;
    __syncthreads();
    // trailing update of the lower triangle: (i, k) with j < k <= i
    const int64_t m = n - j - 1;
    for (int64_t e = tid; e < m * m; e += CHOL_T) {
      const int64_t i = j + 1 + e / m, k = j + 1 + e % m;
      if (k <= i) {
        const double f = -L[k * ld + j];
        L[i * ld + k] = f * L[i * ld + j] + L[i * ld + k];
      }
    }
    __syncthreads();
  }
  if (tid == 0) *fail = 0;
}

__global__ void __launch_bounds__(CHOL_T) chol_solve_small_kernel(int64_t n, int64_t ld, const double* __restrict__ L,
                                                                  const double* __restrict__ rhs, double* __restrict__ b,
                                                                  const int* __restrict__ fail) {
  if (*fail) return;
  const int tid = threadIdx.x;
  __shared__ double smem[32];
  __shared__ double s_coeff;
  for (int64_t i = tid; i < n; i += CHOL_T) b[i] = rhs[i];
  __syncthreads();
  for (int64_t i = 0; i < n; ++i) {  // solve_lower_triangular: column-oriented
    if (tid == 0) {
      s_coeff = b[i] / L[i * ld + i];
      b[i] = s_coeff;
    }
    __syncthreads();
    const double nc = -s_coeff;
    for (int64_t r = i + 1 + tid; r < n; r += CHOL_T) b[r] = nc * L[r * ld + i] + b[r];
    __syncthreads();
  }
  for (int64_t i = n; i-- > 0;) {  // ad_solve_lower_triangular: dot-oriented
    double acc[1] = {0.0};
    for (int64_t r = i + 1 + tid; r < n; r += CHOL_T) acc[0] = acc[0] + L[r * ld + i] * b[r];
    cta_reduce<1>(acc, RedOps<1>{{RED_SUM}}, smem);
    if (tid == 0) b[i] = (b[i] - acc[0]) / L[i * ld + i];
    __syncthreads();
  }
}

int chol_blocked_factor(Ctx* ctx, int64_t n, int64_t ld, const double* hess, double* chol, int* d_fail);
void chol_blocked_solve(Ctx* ctx, int64_t n, int64_t ld, const double* chol, const double* rhs, double* out, const int* d_fail);

constexpr int64_t CHOL_SMALL_MAX = 256;

// hess != nullptr: factor hess into chol (lower Cholesky), then solve chol chol^T w = rhs;
// hess == nullptr: reuse the factor already in chol.
int newton_solve(Ctx* ctx, int64_t n, int64_t ld, const double* hess, double* chol, const double* rhs, double* w_out) {
  static thread_local int* d_fail = nullptr;
  if (!d_fail) OSB_CUDA(cudaMalloc(&d_fail, sizeof(int)));
  if (n <= CHOL_SMALL_MAX) {
    if (hess) {
      chol_small_kernel<<<1, CHOL_T, 0, ctx->stream>>>(n, ld, hess, chol, d_fail);
      ctx->counters[0]++;
    }
    chol_solve_small_kernel<<<1, CHOL_T, 0, ctx->stream>>>(n, ld, chol, rhs, w_out, d_fail);
    ctx->counters[0]++;
  } else {
    if (hess) chol_blocked_factor(ctx, n, ld, hess, chol, d_fail);
    chol_blocked_solve(ctx, n, ld, chol, rhs, w_out, d_fail);
  }
  if (hess) {
    int fail = 0;
    OSB_CUDA(cudaMemcpyAsync(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    if (fail) return OSB_PANIC_NOT_SPD;
  }
  return OSB_OK;
}

// placeholders until the blocked DMMA path lands
int chol_blocked_factor(Ctx*, int64_t, int64_t, const double*, double*, int*) {
  throw Error(OSB_ERR_UNSUPPORTED, "blocked Cholesky (n > 256) not built yet");
}
void chol_blocked_solve(Ctx*, int64_t, int64_t, const double*, const double*, double*, const int*) {
  throw Error(OSB_ERR_UNSUPPORTED, "blocked Cholesky (n > 256) not built yet");
}

}  // namespace osb
