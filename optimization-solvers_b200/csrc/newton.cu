// newton.cu — dense SPD factor + solve for the Newton family
// (newton/mod.rs:31-47, projected_newton.rs:70-78, spn.rs:82-89).
//
// Small systems (n <= CHOL_SMALL_MAX) use a single-CTA right-looking Cholesky whose per-element
// operation order equals nalgebra's left-looking Cholesky::new (each entry receives the updates
// k = 0..j-1 in order, `a + (-L_jk) * L_ik`), so the factor is bit-identical to the reference's;
// the forward solve is the same column-oriented axpy sweep.  Larger systems go through the blocked
// path (chol_blocked.cu).
#include <algorithm>

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

// ---- LU with partial pivoting: plain Newton on a Hessian that is not SPD ---------------------
// newton/mod.rs:31-47 inverts with `try_inverse` (nalgebra: Gauss elimination with partial pivoting for n >= 5), so an
// indefinite but invertible Hessian is a normal iteration there, and a singular one (an exactly zero pivot) falls back
// to d = -g.  The Cholesky above is the fast path for the SPD Hessians of the configs; when it reports a non-positive
// pivot, OSB_NEWTON comes here.  Right-looking elimination, two launches per column (pivot search + row swap on one CTA,
// rank-1 update of the trailing block on the grid): a fallback, not a tuned path.  Pivot rule = nalgebra's icamax
// (first row of maximal |a|).  `lu` holds L (unit lower, multipliers) and U; perm[i] = source row of row i.
constexpr int LU_T = 1024;
__global__ void __launch_bounds__(256) lu_copy_kernel(int64_t n, int64_t ld, const double* __restrict__ A, double* __restrict__ LU, int* __restrict__ perm,
                                                      int* __restrict__ fail) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * ld; e += (int64_t)gridDim.x * blockDim.x) LU[e] = A[e];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) perm[i] = (int)i;
  if (blockIdx.x == 0 && threadIdx.x == 0) *fail = 0;
}
__global__ void __launch_bounds__(LU_T) lu_pivot_kernel(int64_t n, int64_t ld, double* __restrict__ LU, int* __restrict__ perm, int64_t k,
                                                        int* __restrict__ fail) {
  if (*fail) return;
  __shared__ double sv[LU_T / 32];
  __shared__ int64_t si[LU_T / 32];
  __shared__ int64_t s_piv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double best = -1.0;
  int64_t bi = n;
  for (int64_t r = k + tid; r < n; r += LU_T) {
    const double v = fabs(LU[r * ld + k]);
    if (v > best) {  // ascending r per thread: ties keep the smaller row
      best = v;
      bi = r;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if (lane == 0) {
    sv[warp] = best;
    si[warp] = bi;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < LU_T / 32; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) {
        best = sv[w];
        bi = si[w];
      }
    // NaN columns: every compare is false, bi stays n -> treated like the reference's `diag == 0` failure is NOT right
    // (NaN != 0 there and the inverse is full of NaN); keep row k as the pivot in that case
    if (bi >= n) bi = k;
    s_piv = bi;
    if (LU[bi * ld + k] == 0.0) *fail = 1;  // lu.rs: `if diag.is_zero() { return false }`
    else if (bi != k) {
      const int t = perm[k];
      perm[k] = perm[bi];
      perm[bi] = t;
    }
  }
  __syncthreads();
  const int64_t piv = s_piv;
  if (piv == k || LU[piv * ld + k] == 0.0) return;
  for (int64_t c = tid; c < n; c += LU_T) {
    const double a = LU[k * ld + c];
    LU[k * ld + c] = LU[piv * ld + c];
    LU[piv * ld + c] = a;
  }
}
__global__ void __launch_bounds__(256) lu_update_kernel(int64_t n, int64_t ld, double* __restrict__ LU, int64_t k, const int* __restrict__ fail) {
  if (*fail) return;
  const double inv_diag = 1.0 / LU[k * ld + k];  // lu.rs gauss_step: multipliers = column * (1 / pivot)
  for (int64_t r = k + 1 + blockIdx.x; r < n; r += gridDim.x) {
    const double m = LU[r * ld + k] * inv_diag;
    const double nm = -m;
    for (int64_t c = k + 1 + threadIdx.x; c < n; c += blockDim.x) LU[r * ld + c] = nm * LU[k * ld + c] + LU[r * ld + c];
    __syncthreads();
    if (threadIdx.x == 0) LU[r * ld + k] = m;
    __syncthreads();
  }
}
// out = U^-1 L^-1 P rhs (column-oriented substitutions, like solve_lower/upper_triangular_mut)
__global__ void __launch_bounds__(LU_T) lu_solve_kernel(int64_t n, int64_t ld, const double* __restrict__ LU, const int* __restrict__ perm,
                                                        const double* __restrict__ rhs, double* __restrict__ b, double* __restrict__ tmp,
                                                        const int* __restrict__ fail) {
  if (*fail) return;
  const int tid = threadIdx.x;
  __shared__ double s_coeff;
  for (int64_t i = tid; i < n; i += LU_T) tmp[i] = rhs[perm[i]];
  __syncthreads();
  for (int64_t i = tid; i < n; i += LU_T) b[i] = tmp[i];
  __syncthreads();
  for (int64_t i = 0; i + 1 < n; ++i) {
    const double nc = -b[i];
    __syncthreads();
    for (int64_t r = i + 1 + tid; r < n; r += LU_T) b[r] = nc * LU[r * ld + i] + b[r];
    __syncthreads();
  }
  for (int64_t i = n; i-- > 0;) {
    if (tid == 0) {
      s_coeff = b[i] / LU[i * ld + i];
      b[i] = s_coeff;
    }
    __syncthreads();
    const double nc = -s_coeff;
    for (int64_t r = tid; r < i; r += LU_T) b[r] = nc * LU[r * ld + i] + b[r];
    __syncthreads();
  }
}

// factor (hess != nullptr) and/or solve with the LU in `lu`; returns OSB_OK, or OSB_PANIC_NOT_SPD standing for
// "singular" (the caller maps it to the reference's d = -g fallback)
int newton_solve_lu(Ctx* ctx, int64_t n, int64_t ld, const double* hess, double* lu, int* perm, double* tmp, const double* rhs, double* w_out) {
  static thread_local int* d_fail = nullptr;
  if (!d_fail) OSB_CUDA(cudaMalloc(&d_fail, sizeof(int)));
  cudaStream_t st = ctx->stream;
  if (hess) {
    lu_copy_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(n, ld, hess, lu, perm, d_fail);
    ctx->counters[0]++;
    for (int64_t k = 0; k < n; ++k) {
      lu_pivot_kernel<<<1, LU_T, 0, st>>>(n, ld, lu, perm, k, d_fail);
      ctx->counters[0]++;
      if (k + 1 < n) {
        const int grid = (int)std::min<int64_t>(n - k - 1, (int64_t)ctx->num_sms * 4);
        lu_update_kernel<<<grid, 256, 0, st>>>(n, ld, lu, k, d_fail);
        ctx->counters[0]++;
      }
    }
    int fail = 0;
    OSB_CUDA(cudaMemcpyAsync(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, st));
    ctx->sync();
    if (fail) return OSB_PANIC_NOT_SPD;
  }
  lu_solve_kernel<<<1, LU_T, 0, st>>>(n, ld, lu, perm, rhs, w_out, tmp, d_fail);
  ctx->counters[0]++;
  return OSB_OK;
}

// ---- blocked right-looking Cholesky (n > CHOL_SMALL_MAX) -----------------------------------
// per block column k: POTRF of the 64x64 diagonal block in shared memory, TRSM of the panel below
// (one thread per row, the row held in registers), SYRK update of the trailing lower triangle with
// 64x64 register-tiled output tiles.  n^3/3 flops in total — three orders of magnitude below the
// Hessian assembly of the logistic config, so plain FP64 FMA is the right tool here.
constexpr int CB = 64;

__global__ void __launch_bounds__(256) chol_copy_lower_kernel(int64_t n, int64_t ld, const double* __restrict__ A, double* __restrict__ L) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * ld; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / ld, j = e % ld;
    L[e] = (j <= i && j < n) ? A[e] : 0.0;
  }
}

__global__ void __launch_bounds__(256) chol_potrf_block_kernel(int64_t n, int64_t ld, double* __restrict__ L, int64_t k0, int* __restrict__ fail) {
  __shared__ double T[CB][CB + 1];
  __shared__ int s_fail;
  if (*fail) return;
  const int nb = (int)(n - k0 < CB ? n - k0 : CB);
  const int tid = threadIdx.x;
  if (tid == 0) s_fail = 0;
  for (int e = tid; e < nb * nb; e += 256) {
    const int i = e / nb, j = e % nb;
    T[i][j] = (j <= i) ? L[(k0 + i) * ld + k0 + j] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (tid == 0) {
      const double d = T[j][j];
      if (d != 0.0 && d >= 0.0) T[j][j] = sqrt(d);
      else s_fail = 1;
    }
    __syncthreads();
    if (s_fail) {
      if (tid == 0) *fail = 1;
      return;
    }
    const double dj = T[j][j];
    for (int i = j + 1 + tid; i < nb; i += 256) T[i][j] = T[i][j] / dj;
    __syncthreads();
    const int mrem = nb - j - 1;
    for (int e = tid; e < mrem * mrem; e += 256) {
      const int i = j + 1 + e / mrem, k = j + 1 + e % mrem;
      if (k <= i) T[i][k] = (-T[k][j]) * T[i][j] + T[i][k];
    }
    __syncthreads();
  }
  for (int e = tid; e < nb * nb; e += 256) {
    const int i = e / nb, j = e % nb;
    if (j <= i) L[(k0 + i) * ld + k0 + j] = T[i][j];
  }
}

// panel rows i >= k0 + nb: L[i, k0:k0+nb] <- A[i, k0:k0+nb] L11^{-T}
__global__ void __launch_bounds__(128) chol_trsm_kernel(int64_t n, int64_t ld, double* __restrict__ L, int64_t k0, const int* __restrict__ fail) {
  __shared__ double T[CB][CB + 1];
  if (*fail) return;
  const int nb = (int)(n - k0 < CB ? n - k0 : CB);
  for (int e = threadIdx.x; e < nb * nb; e += 128) {
    const int i = e / nb, j = e % nb;
    T[i][j] = (j <= i) ? L[(k0 + i) * ld + k0 + j] : 0.0;
  }
  __syncthreads();
  const int64_t i = k0 + nb + (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  double xr[CB];
  double* row = L + i * ld + k0;
#pragma unroll
  for (int c = 0; c < CB; ++c) xr[c] = c < nb ? row[c] : 0.0;
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    if (c < nb) {
      double v = xr[c];
#pragma unroll
      for (int p = 0; p < c; ++p) v = fma(-xr[p], T[c][p], v);
      xr[c] = v / T[c][c];
    }
  }
#pragma unroll
  for (int c = 0; c < CB; ++c)
    if (c < nb) row[c] = xr[c];
}

// trailing update, lower-triangular 64x64 tiles: A[i][j] -= sum_p L[i][k0+p] L[j][k0+p]
__global__ void __launch_bounds__(256) chol_syrk_kernel(int64_t n, int64_t ld, double* __restrict__ L, int64_t k0, int nb, int64_t t0,
                                                        const int* __restrict__ fail) {
  constexpr int PH = 32;  // the 64-wide panel is consumed in two halves (48 KB static shared memory limit)
  __shared__ double Pi[CB][PH + 2];
  __shared__ double Pj[CB][PH + 2];
  if (*fail) return;
  int t = blockIdx.x, ti = 0;
  while (t > ti) {
    t -= ti + 1;
    ++ti;
  }
  const int tj = t;
  const int64_t i0 = t0 + (int64_t)ti * CB, j0 = t0 + (int64_t)tj * CB;
  const int tid = threadIdx.x;
  const int tr = (tid / 16) * 4, tc = (tid % 16) * 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int ph = 0; ph < CB; ph += PH) {
    for (int e = tid; e < CB * PH; e += 256) {
      const int r = e / PH, p = e % PH;
      Pi[r][p] = (i0 + r < n && ph + p < nb) ? L[(i0 + r) * ld + k0 + ph + p] : 0.0;
      Pj[r][p] = (j0 + r < n && ph + p < nb) ? L[(j0 + r) * ld + k0 + ph + p] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int p = 0; p < PH; ++p) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = Pi[tr + a][p];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Pj[tc + b][p];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t i = i0 + tr + a, j = j0 + tc + b;
      if (i < n && j < n && j <= i) L[i * ld + j] = L[i * ld + j] - acc[a][b];
    }
}

int chol_blocked_factor(Ctx* ctx, int64_t n, int64_t ld, const double* hess, double* chol, int* d_fail) {
  cudaStream_t st = ctx->stream;
  OSB_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), st));
  chol_copy_lower_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(n, ld, hess, chol);
  ctx->counters[0]++;
  for (int64_t k0 = 0; k0 < n; k0 += CB) {
    const int nb = (int)(n - k0 < CB ? n - k0 : CB);
    chol_potrf_block_kernel<<<1, 256, 0, st>>>(n, ld, chol, k0, d_fail);
    const int64_t rem = n - k0 - nb;
    ctx->counters[0]++;
    if (rem <= 0) break;
    chol_trsm_kernel<<<(unsigned)((rem + 127) / 128), 128, 0, st>>>(n, ld, chol, k0, d_fail);
    const int nt = (int)((rem + CB - 1) / CB);
    chol_syrk_kernel<<<(unsigned)(nt * (nt + 1) / 2), 256, 0, st>>>(n, ld, chol, k0, nb, k0 + nb, d_fail);
    ctx->counters[0] += 2;
  }
  return OSB_OK;
}

// blocked triangular solves: L y = rhs (forward), L^T w = y (backward)
__global__ void __launch_bounds__(CB) chol_fwd_diag_kernel(int64_t n, int64_t ld, const double* __restrict__ L, double* __restrict__ b, int64_t k0,
                                                           const int* __restrict__ fail) {
  if (*fail) return;
  __shared__ double y[CB];
  const int nb = (int)(n - k0 < CB ? n - k0 : CB);
  const int t = threadIdx.x;
  if (t < nb) y[t] = b[k0 + t];
  __syncthreads();
  for (int c = 0; c < nb; ++c) {
    if (t == c) y[c] = y[c] / L[(k0 + c) * ld + k0 + c];
    __syncthreads();
    if (t > c && t < nb) y[t] = (-y[c]) * L[(k0 + t) * ld + k0 + c] + y[t];
    __syncthreads();
  }
  if (t < nb) b[k0 + t] = y[t];
}
__global__ void __launch_bounds__(128) chol_fwd_update_kernel(int64_t n, int64_t ld, const double* __restrict__ L, double* __restrict__ b, int64_t k0,
                                                              int nb, const int* __restrict__ fail) {
  if (*fail) return;
  __shared__ double y[CB];
  if (threadIdx.x < nb) y[threadIdx.x] = b[k0 + threadIdx.x];
  __syncthreads();
  const int64_t i = k0 + nb + (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  const double* row = L + i * ld + k0;
  double acc = 0.0;
  for (int p = 0; p < nb; ++p) acc = fma(row[p], y[p], acc);
  b[i] = b[i] - acc;
}
__global__ void __launch_bounds__(CB) chol_bwd_diag_kernel(int64_t n, int64_t ld, const double* __restrict__ L, double* __restrict__ b, int64_t k0,
                                                           const int* __restrict__ fail) {
  if (*fail) return;
  __shared__ double y[CB];
  const int nb = (int)(n - k0 < CB ? n - k0 : CB);
  const int t = threadIdx.x;
  if (t < nb) y[t] = b[k0 + t];
  __syncthreads();
  for (int c = nb - 1; c >= 0; --c) {
    if (t == c) y[c] = y[c] / L[(k0 + c) * ld + k0 + c];
    __syncthreads();
    if (t < c) y[t] = y[t] - L[(k0 + c) * ld + k0 + t] * y[c];
    __syncthreads();
  }
  if (t < nb) b[k0 + t] = y[t];
}
__global__ void __launch_bounds__(128) chol_bwd_update_kernel(int64_t n, int64_t ld, const double* __restrict__ L, double* __restrict__ b, int64_t k0,
                                                              int nb, const int* __restrict__ fail) {
  if (*fail) return;
  __shared__ double y[CB];
  if (threadIdx.x < nb) y[threadIdx.x] = b[k0 + threadIdx.x];
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (j >= k0) return;
  double acc = 0.0;
  for (int p = 0; p < nb; ++p) acc = fma(L[(k0 + p) * ld + j], y[p], acc);
  b[j] = b[j] - acc;
}

void chol_blocked_solve(Ctx* ctx, int64_t n, int64_t ld, const double* chol, const double* rhs, double* out, const int* d_fail) {
  cudaStream_t st = ctx->stream;
  if (out != rhs) OSB_CUDA(cudaMemcpyAsync(out, rhs, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  for (int64_t k0 = 0; k0 < n; k0 += CB) {
    const int nb = (int)(n - k0 < CB ? n - k0 : CB);
    chol_fwd_diag_kernel<<<1, CB, 0, st>>>(n, ld, chol, out, k0, d_fail);
    const int64_t rem = n - k0 - nb;
    ctx->counters[0]++;
    if (rem > 0) {
      chol_fwd_update_kernel<<<(unsigned)((rem + 127) / 128), 128, 0, st>>>(n, ld, chol, out, k0, nb, d_fail);
      ctx->counters[0]++;
    }
  }
  const int64_t last = ((n - 1) / CB) * CB;
  for (int64_t k0 = last; k0 >= 0; k0 -= CB) {
    const int nb = (int)(n - k0 < CB ? n - k0 : CB);
    chol_bwd_diag_kernel<<<1, CB, 0, st>>>(n, ld, chol, out, k0, d_fail);
    ctx->counters[0]++;
    if (k0 > 0) {
      chol_bwd_update_kernel<<<(unsigned)((k0 + 127) / 128), 128, 0, st>>>(n, ld, chol, out, k0, nb, d_fail);
      ctx->counters[0]++;
    }
  }
}

}  // namespace osb
