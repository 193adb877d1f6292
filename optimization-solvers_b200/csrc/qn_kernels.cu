// qn_kernels.cu — the dense quasi-Newton hot path on the device-resident inverse-Hessian
// approximation H (row-major, leading dimension ld = n rounded up to 16 doubles, zero padded).
//
// Replaces, per outer iteration of the reference:
//   compute_direction          d = (-H) g                     src/quasi_newton/bfgs.rs:43-48
//   update_next_iterate        two dense n^3 GEMMs            src/quasi_newton/bfgs.rs:115-124
//                              (dfp.rs:115-120, broyden.rs:115-118, sr1_b.rs:143-146)
// with the algebraically equal O(n^2) schedule:
//   pass 1  qn_gemv_kernel     h = H y                        reads  n^2 * 8 B
//   pass 2  qn_update_kernel   H <- H + rank-2(s, h) and, in the same read-modify-write,
//                              u = H' g_new                   reads + writes n^2 * 8 B
// so one iteration moves 3 * n^2 * 8 B of HBM traffic.  Both kernels are HBM-bound streaming
// kernels: 128-bit coalesced loads with no L1 allocation, R rows per CTA tile so that the O(n)
// vectors are re-read from L1/L2 only once per R rows, fused-multiply-add only in the dot
// accumulators, warp-shuffle + fixed-order cross-warp reduction (deterministic; a row's sum is
// formed entirely inside one CTA, so the result does not depend on the grid or on the number of
// GPUs the rows are sharded over).
#include <cstdlib>

#include "engine.cuh"
#include "qn_sym.cuh"

namespace osb {

int64_t qn_ld(int64_t n) { return (n + 15) / 16 * 16; }
// H is allocated with its row count rounded up to a whole tile (zero rows), so that the kernels
// need no row clamping (a CTA's last, partial tile still LOADS whole tiles: hence the extra tile of slack);
// vectors are allocated with ld entries (zero padded).
int64_t qn_rows_padded(int64_t nrows) { return (nrows + QN_R - 1) / QN_R * QN_R + QN_R; }

// Tile height: tiles of `rt` <= QN_R rows are dealt round-robin to the CTAs (round-robin keeps concurrently
// running CTAs on ADJACENT rows, which the memory system rewards: profiles/r01_rmw_pattern_experiments.txt;
// contiguous per-CTA row ranges were measured 2 % slower).  With rt = 8, 256 tiles on 148 CTAs (the 8-GPU
// shard of n = 16384) take two full waves for 1.73 waves of work; rt = 7 gives 293 tiles = 1.98 waves.
// The host picks the rt in 5..8 that minimises waves * rt.
inline int qn_pick_tile_rows(int64_t nrows, int grid) {
  int best = QN_R;
  int64_t best_cost = -1;
  for (int rt = QN_R; rt >= 5; --rt) {
    const int64_t tiles = (nrows + rt - 1) / rt;
    const int64_t cost = ((tiles + grid - 1) / grid) * rt;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = rt;
    }
  }
  return best;
}

__device__ __forceinline__ void tile_reduce_store(double (&acc)[QN_R], double (*red)[QN_T / 32], double* out, int64_t row_base,
                                                  int64_t r0, int64_t rend) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < QN_R; ++r) {
    double v = warp_sum(acc[r]);
    if (lane == 0) red[r][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < QN_R) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < QN_T / 32; ++w) v = v + red[threadIdx.x][w];
    if (r0 + threadIdx.x < rend) out[row_base + r0 + threadIdx.x] = v;
    __threadfence();  // the fused coefficient epilogue of another CTA may read this row sum
  }
  __syncthreads();
}

template <int NTHREADS>
__device__ __forceinline__ void qn_coef_body(int kind, int64_t n, DevState* st, const double* __restrict__ s,
                                             const double* __restrict__ y, const double* h, double* __restrict__ p_out,
                                             double* smem);

// ---- pass 1: out = H v ---------------------------------------------------------------------
__global__ void __launch_bounds__(QN_T, 2)
qn_gemv_kernel(const double* __restrict__ H, int64_t ld, int64_t nrows, int64_t row0, DevState* st,
               const double* __restrict__ v, double* out, const double* __restrict__ v_skip,
               double* __restrict__ out_skip, QNCoefArgs ca, int rt) {
  bool fuse_coef = ca.ticket != nullptr;
  if (st != nullptr) {
    if (st->done) return;
    if (st->skip) {  // bfgs.rs:106-112: H unchanged; only the next direction's u = H g is needed
      v = v_skip;
      out = out_skip;
      fuse_coef = false;
    }
  }
  if (v == nullptr) return;
  const unsigned long long pol = l2_evict_first_policy();
  __shared__ double red[QN_R][QN_T / 32];
  __shared__ bool is_last;
  for (int64_t r0 = (int64_t)blockIdx.x * rt; r0 < nrows; r0 += (int64_t)gridDim.x * rt) {
    const int64_t re = r0 + rt < nrows ? r0 + rt : nrows;
    const int rows_here = (int)(re - r0);
    double acc[QN_R];
#pragma unroll
    for (int r = 0; r < QN_R; ++r) acc[r] = 0.0;
    const double* __restrict__ base = H + r0 * ld;
    for (int col = 2 * threadIdx.x; col < (int)ld; col += QN_CHUNK) {
      const double2 vv = ld_vec2(v + col);
      double2 hv[QN_R];
#pragma unroll
      for (int r = 0; r < QN_R; ++r) hv[r] = r < rows_here ? ld_stream_nc_ef(base + r * ld + col, pol) : make_double2(0.0, 0.0);
#pragma unroll
      for (int r = 0; r < QN_R; ++r) {
        acc[r] = fma(hv[r].x, vv.x, acc[r]);
        acc[r] = fma(hv[r].y, vv.y, acc[r]);
      }
    }
    tile_reduce_store(acc, red, out, row0, r0, re);
  }
  if (!fuse_coef) return;
  // fused epilogue (single GPU): the last CTA to finish owns the complete h and computes y.h and the
  // update coefficients, saving a launch on the critical path
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(ca.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  qn_coef_body<QN_T>(ca.kind, ca.n, st, ca.s, ca.y, out, ca.p_out, &red[0][0]);
  if (threadIdx.x == 0) *ca.ticket = 0u;
}

void qn_launch_gemv(Ctx* ctx, const double* H, int64_t ld, int64_t nrows, int64_t row0, const DevState* st, const double* v,
                    double* out, const double* v_skip, double* out_skip, int variant, const QNCoefArgs* coef) {
  (void)variant;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nrows + QN_R - 1) / QN_R, (int64_t)ctx->num_sms * 2));
  const int rt = qn_pick_tile_rows(nrows, grid);
  QNCoefArgs ca{};
  if (coef) ca = *coef;
  // rows are addressed relative to the local block; `out` is indexed by global row (row0 + local row)
  qn_gemv_kernel<<<grid, QN_T, 0, ctx->stream>>>(H, ld, nrows, row0, const_cast<DevState*>(st), v, out, v_skip, out_skip, ca, rt);
  ctx->counters[0]++;
}

// ---- Broyden: out[j] = sum_i H_ij s_i  (column sums; two deterministic stages) -------------
constexpr int GT_T = 256, GT_ROWS = 64;
__global__ void __launch_bounds__(GT_T) qn_gemvT_stage1(const double* __restrict__ H, int64_t ld, int64_t nrows, int64_t row0,
                                                        const DevState* __restrict__ st, const double* __restrict__ s,
                                                        double* __restrict__ partial) {
  if (st != nullptr && (st->done || st->skip)) return;
  const int64_t col = ((int64_t)blockIdx.x * GT_T + threadIdx.x) * 2;
  if (col >= ld) return;
  const int64_t rb = (int64_t)blockIdx.y * GT_ROWS;
  const int64_t re = rb + GT_ROWS < nrows ? rb + GT_ROWS : nrows;
  double a0 = 0.0, a1 = 0.0;
  for (int64_t i = rb; i < re; ++i) {
    const double si = s[row0 + i];
    const double2 hv = ld_stream_nc(H + i * ld + col);
    a0 = fma(hv.x, si, a0);
    a1 = fma(hv.y, si, a1);
  }
  *reinterpret_cast<double2*>(partial + (int64_t)blockIdx.y * ld + col) = make_double2(a0, a1);
}
__global__ void __launch_bounds__(GT_T) qn_gemvT_stage2(int64_t ld, int64_t nsplit, const DevState* __restrict__ st,
                                                        const double* __restrict__ partial, double* __restrict__ out) {
  if (st != nullptr && (st->done || st->skip)) return;
  const int64_t col = (int64_t)blockIdx.x * GT_T + threadIdx.x;
  if (col >= ld) return;
  double a = 0.0;
  for (int64_t k = 0; k < nsplit; ++k) a = a + partial[k * ld + col];
  out[col] = a;
}
void qn_launch_gemvT(Ctx* ctx, const double* H, int64_t ld, int64_t nrows, int64_t row0, const DevState* st, const double* s,
                     double* out, double* scratch) {
  int64_t nsplit = (nrows + GT_ROWS - 1) / GT_ROWS;
  dim3 g1((unsigned)((ld / 2 + GT_T - 1) / GT_T), (unsigned)nsplit);
  qn_gemvT_stage1<<<g1, GT_T, 0, ctx->stream>>>(H, ld, nrows, row0, st, s, scratch);
  qn_gemvT_stage2<<<(unsigned)((ld + GT_T - 1) / GT_T), GT_T, 0, ctx->stream>>>(ld, nsplit, st, scratch, out);
  ctx->counters[0] += 2;
}

// ---- update coefficients -------------------------------------------------------------------
// BFGS  (bfgs.rs:115-124)     H' = H - rho (s h^T + h s^T) + (rho^2 y.h + rho) s s^T,  rho = 1/(y.s), h = H y
// DFP   (dfp.rs:115-120)      H' = H + s s^T/(s.y) - h h^T/(y.h)
// SR1   (sr1_b.rs:143-146)    H' = H + p p^T/(p.y),            p = s - h
// Broyden (broyden.rs:115-118) H' = H + p v^T/(s.y),           p = s - h, v = H^T s
// Divisions by the scalar denominators are applied as multiplications by their reciprocals
// (<= 1 ulp per term away from the reference's elementwise division; documented in DESIGN.md).
template <int NTHREADS>
__device__ __forceinline__ void qn_coef_body(int kind, int64_t n, DevState* st, const double* __restrict__ s,
                                             const double* __restrict__ y, const double* h, double* __restrict__ p_out,
                                             double* smem) {
  double acc[2] = {0.0, 0.0};
  for (int64_t i = threadIdx.x; i < n; i += NTHREADS) {
    const double yi = y[i], hi = h[i];
    acc[0] = fma(yi, hi, acc[0]);
    if (kind == QN_SR1 || kind == QN_BROYDEN) {
      const double pi = s[i] - hi;
      p_out[i] = pi;
      acc[1] = fma(pi, yi, acc[1]);
    }
  }
  RedOps<2> ops{{RED_SUM, RED_SUM}};
  cta_reduce<2>(acc, ops, smem);
  if (threadIdx.x == 0) {
    const double yh = acc[0], ys = st->ys;
    st->yh = yh;
    if (kind == QN_BFGS) {
      const double rho = 1.0 / ys;
      st->c0 = rho * rho * yh + rho;
      st->c1 = -rho;
      st->c2 = 0.0;
    } else if (kind == QN_DFP) {
      st->c0 = 1.0 / ys;
      st->c1 = 0.0;
      st->c2 = -1.0 / yh;
    } else if (kind == QN_SR1) {
      st->c0 = 1.0 / acc[1];
      st->c1 = st->c2 = 0.0;
    } else {
      st->c0 = 1.0 / ys;
      st->c1 = st->c2 = 0.0;
    }
  }
}

// standalone launch (row-sharded H: y.h needs the all-gathered h).  Same thread count as the fused
// epilogue, hence the same summation order: results are bit-identical for every number of GPUs.
__global__ void __launch_bounds__(QN_T) qn_coef_kernel(int kind, int64_t n, DevState* st, const double* __restrict__ s,
                                                       const double* __restrict__ y, const double* __restrict__ h,
                                                       double* __restrict__ p_out) {
  if (st->done || st->skip) return;
  __shared__ double smem[2 * 32];
  qn_coef_body<QN_T>(kind, n, st, s, y, h, p_out, smem);
}
void qn_launch_coef(Ctx* ctx, int kind, int64_t n, DevState* st, const double* s, const double* y, const double* h,
                    double* p_out) {
  qn_coef_kernel<<<1, QN_T, 0, ctx->stream>>>(kind, n, st, s, y, h, p_out);
  ctx->counters[0]++;
}

// ---- pass 2: fused rank-2 read-modify-write + u = H' g -------------------------------------
template <int KIND>
__global__ void __launch_bounds__(QN_T, 1)
qn_update_kernel(double* __restrict__ H, int64_t ld, int64_t nrows, int64_t row0, const DevState* __restrict__ st,
                 const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ rv,
                 const double* __restrict__ g, double* __restrict__ u_out, int rt) {
  if (st->done || st->skip) return;
  const double c0 = st->c0, c1 = st->c1, c2 = st->c2;
  const unsigned long long pol = l2_evict_first_policy();
  __shared__ double red[QN_R][QN_T / 32];
  __shared__ double2 rowpq[QN_R];
  for (int64_t r0 = (int64_t)blockIdx.x * rt; r0 < nrows; r0 += (int64_t)gridDim.x * rt) {
    const int64_t re = r0 + rt < nrows ? r0 + rt : nrows;
    const int rows_here = (int)(re - r0);
    double acc[QN_R];
#pragma unroll
    for (int r = 0; r < QN_R; ++r) acc[r] = 0.0;
    // per-row scalars (p_i, q_i) are CTA-uniform: kept in shared memory and read as broadcasts
    if (threadIdx.x < QN_R) {
      const bool ok = (int)threadIdx.x < rows_here;
      rowpq[threadIdx.x] = make_double2(ok ? p[row0 + r0 + threadIdx.x] : 0.0,
                                        (ok && (KIND == QN_BFGS || KIND == QN_DFP)) ? q[row0 + r0 + threadIdx.x] : 0.0);
    }
    __syncthreads();
    double* __restrict__ base = H + r0 * ld;
    for (int col = 2 * threadIdx.x; col < (int)ld; col += QN_CHUNK) {
      const double2 gj = ld_vec2(g + col);
      double2 pj = make_double2(0.0, 0.0), qj = make_double2(0.0, 0.0);
      if (KIND != QN_BROYDEN) pj = ld_vec2(p + col);
      if (KIND == QN_BFGS || KIND == QN_DFP) qj = ld_vec2(q + col);
      if (KIND == QN_BROYDEN) pj = ld_vec2(rv + col);  // column vector v = H^T s
      double2 hv[QN_R];
#pragma unroll
      for (int r = 0; r < QN_R; ++r) hv[r] = r < rows_here ? ld_stream_ef(base + r * ld + col, pol) : make_double2(0.0, 0.0);
#pragma unroll
      for (int r = 0; r < QN_R; ++r) {
        double2 hn;
        const double2 pq = rowpq[r];
        const double pi = pq.x, qi = pq.y;
        if (KIND == QN_BFGS) {
          // symmetric in (row, col): products are commutative and the cross term is a plain sum
          const double cx = pi * qj.x + qi * pj.x, cy = pi * qj.y + qi * pj.y;
          hn.x = fma(c0, pi * pj.x, fma(c1, cx, hv[r].x));
          hn.y = fma(c0, pi * pj.y, fma(c1, cy, hv[r].y));
        } else if (KIND == QN_DFP) {
          hn.x = fma(c2, qi * qj.x, fma(c0, pi * pj.x, hv[r].x));
          hn.y = fma(c2, qi * qj.y, fma(c0, pi * pj.y, hv[r].y));
        } else {  // SR1: p p^T ; Broyden: p v^T
          hn.x = fma(c0, pi * pj.x, hv[r].x);
          hn.y = fma(c0, pi * pj.y, hv[r].y);
        }
        acc[r] = fma(hn.x, gj.x, acc[r]);
        acc[r] = fma(hn.y, gj.y, acc[r]);
        if (r < rows_here) st_stream_ef(base + r * ld + col, hn, pol);
      }
    }
    tile_reduce_store(acc, red, u_out, row0, r0, re);
  }
}

// ---- lazy schedule -------------------------------------------------------------------------
// Stored matrix M = H_k minus the update of iteration k-1 (pending).  One read-modify-write:
//     M_ij <- M_ij + pc0 p_i p_j + pc1 (p_i q_j + q_i p_j) + pc2 q_i q_j      (now M = H_k exactly)
//     h_i = sum_j M_ij y_j ,  w_i = sum_j M_ij g_j
// The update of iteration k, H_{k+1} = H_k + rank2(s, h; c), is NOT applied: it becomes the new pending
// update, and the next direction needs only u = H_{k+1} g = w + s (c0 s.g + c1 h.g) + h (c1 s.g + c2 h.g),
// an O(n) epilogue.  HBM traffic per iteration: 2 n^2 8 B instead of 3 n^2 8 B.
template <int KIND>
// (not inlined, and the kernels pass their __grid_constant__ parameter block by address: the streaming loops keep
//  the whole register budget — inlined, this body cost qn_lazy_kernel 40 us per launch in spills and scheduling)
__device__ __noinline__ void lazy_epilogue_body(const QNLazyArgs& a, const double* hsrc, const double* wsrc, double* smem) {
  DevState* st = a.st;
  const int64_t n = a.n;
  if (st->skip) {  // bfgs.rs:106-112: no new update; the stored matrix is now exact
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a.u[i] = wsrc[i];
    if (threadIdx.x == 0) {
      st->pending = 0;
      st->pc0 = st->pc1 = st->pc2 = 0.0;
    }
    return;
  }
  // One CTA walks the O(n) vectors out of L2: the loop is latency bound, so loads are issued in batches of EPI_U
  // before any use (the stores of the second loop would otherwise serialise it: the pointers may alias as far as
  // the compiler knows).  The order of the additions is unchanged: element i goes to thread i % blockDim.x, ascending.
  constexpr int EPI_U = 8;
  const int64_t bd = blockDim.x;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i0 = threadIdx.x; i0 < n; i0 += bd * EPI_U) {
    double hv[EPI_U], gv[EPI_U], yv[EPI_U], sv[EPI_U];
#pragma unroll
    for (int k = 0; k < EPI_U; ++k) {
      const int64_t i = i0 + k * bd;
      const bool ok = i < n;
      hv[k] = ok ? __ldcg(hsrc + i) : 0.0;
      gv[k] = ok ? __ldcg(a.g + i) : 0.0;
      yv[k] = ok ? __ldcg(a.y + i) : 0.0;
      sv[k] = ok ? __ldcg(a.s + i) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < EPI_U; ++k) {
      if (i0 + k * bd < n) {
        acc[0] = fma(yv[k], hv[k], acc[0]);  // y.h
        acc[1] = fma(sv[k], gv[k], acc[1]);  // s.g
        acc[2] = fma(hv[k], gv[k], acc[2]);  // h.g
      }
    }
  }
  RedOps<3> ops{{RED_SUM, RED_SUM, RED_SUM}};
  cta_reduce<3>(acc, ops, smem);
  const double yh = acc[0], sg = acc[1], hg = acc[2], ys = st->ys;
  double c0, c1, c2;
  if (KIND == QN_BFGS) {
    const double rho = 1.0 / ys;
    c0 = rho * rho * yh + rho;
    c1 = -rho;
    c2 = 0.0;
  } else {  // DFP
    c0 = 1.0 / ys;
    c1 = 0.0;
    c2 = -1.0 / yh;
  }
  const double ca = c0 * sg + c1 * hg, cb = c1 * sg + c2 * hg;
  for (int64_t i0 = threadIdx.x; i0 < n; i0 += bd * EPI_U) {
    double hv[EPI_U], sv[EPI_U], wv[EPI_U];
#pragma unroll
    for (int k = 0; k < EPI_U; ++k) {
      const int64_t i = i0 + k * bd;
      const bool ok = i < n;
      hv[k] = ok ? __ldcg(hsrc + i) : 0.0;
      sv[k] = ok ? __ldcg(a.s + i) : 0.0;
      wv[k] = ok ? __ldcg(wsrc + i) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < EPI_U; ++k) {
      const int64_t i = i0 + k * bd;
      if (i < n) {
        a.u[i] = wv[k] + (sv[k] * ca + hv[k] * cb);
        a.ps_out[i] = sv[k];
        a.ph_out[i] = hv[k];
      }
    }
  }
  if (threadIdx.x == 0) {
    st->yh = yh;
    st->c0 = c0;
    st->c1 = c1;
    st->c2 = c2;
    st->pc0 = c0;
    st->pc1 = c1;
    st->pc2 = c2;
    st->pending = 1;
  }
}

// DEFER (compile time, = a.defer_epi): the epilogue is left to the next cluster head.  A run-time test of the flag
// in this kernel was enough to push the streaming loop into spills (128 registers, 1 CTA of 512 threads per SM).
// P2P (compile time, = P2P): the single-GPU instantiation carries no exchange state through the loop.
template <int KIND, bool DEFER, bool P2P>
__global__ void __launch_bounds__(QN_T, 1) qn_lazy_kernel(const __grid_constant__ QNLazyArgs a) {
  DevState* st = a.st;
  if (st->done) return;
  const double c0 = st->pc0, c1 = st->pc1, c2 = st->pc2;
  const unsigned long long pol = l2_evict_first_policy();
  __shared__ double red[2 * QN_R][QN_T / 32];
  __shared__ double2 rowpq[QN_R];
  __shared__ bool is_last;
  const int64_t ld = a.ld, nrows = a.nrows, row0 = a.row0;
  const double* __restrict__ p = a.ps;
  const double* __restrict__ q = a.ph;
  const double* __restrict__ yv = a.y;
  const double* __restrict__ gv = a.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long seq = P2P ? *a.seq + 1ULL : 0ULL;  // this exchange's sequence number
  const int par = (int)(seq & 1ULL);
  const int rt = a.tile_rows;
  for (int64_t r0 = (int64_t)blockIdx.x * rt; r0 < nrows; r0 += (int64_t)gridDim.x * rt) {
    const int64_t re = r0 + rt < nrows ? r0 + rt : nrows;
    const int rows_here = (int)(re - r0);
    double ah[QN_R], aw[QN_R];
#pragma unroll
    for (int r = 0; r < QN_R; ++r) ah[r] = aw[r] = 0.0;
    if (threadIdx.x < QN_R) {
      const bool ok = (int)threadIdx.x < rows_here;
      rowpq[threadIdx.x] = make_double2(ok ? p[row0 + r0 + threadIdx.x] : 0.0, ok ? q[row0 + r0 + threadIdx.x] : 0.0);
    }
    __syncthreads();
    double* __restrict__ base = a.M + r0 * ld;
    for (int col = 2 * threadIdx.x; col < (int)ld; col += QN_CHUNK) {
      const double2 gj = ld_vec2(gv + col);
      const double2 yj = ld_vec2(yv + col);
      const double2 pj = ld_vec2(p + col);
      const double2 qj = ld_vec2(q + col);
      double2 hv[QN_R];
#pragma unroll
      for (int r = 0; r < QN_R; ++r) hv[r] = r < rows_here ? ld_stream_ef(base + r * ld + col, pol) : make_double2(0.0, 0.0);
#pragma unroll
      for (int r = 0; r < QN_R; ++r) {
        const double2 pq = rowpq[r];
        const double pi = pq.x, qi = pq.y;
        double2 hn;
        if (KIND == QN_BFGS) {
          const double cx = pi * qj.x + qi * pj.x, cy = pi * qj.y + qi * pj.y;
          hn.x = fma(c0, pi * pj.x, fma(c1, cx, hv[r].x));
          hn.y = fma(c0, pi * pj.y, fma(c1, cy, hv[r].y));
        } else {
          hn.x = fma(c2, qi * qj.x, fma(c0, pi * pj.x, hv[r].x));
          hn.y = fma(c2, qi * qj.y, fma(c0, pi * pj.y, hv[r].y));
        }
        ah[r] = fma(hn.x, yj.x, ah[r]);
        ah[r] = fma(hn.y, yj.y, ah[r]);
        aw[r] = fma(hn.x, gj.x, aw[r]);
        aw[r] = fma(hn.y, gj.y, aw[r]);
        if (r < rows_here) st_stream_ef(base + r * ld + col, hn, pol);
      }
    }
#pragma unroll
    for (int r = 0; r < QN_R; ++r) {
      const double v1 = warp_sum(ah[r]), v2 = warp_sum(aw[r]);
      if (lane == 0) {
        red[r][warp] = v1;
        red[QN_R + r][warp] = v2;
      }
    }
    __syncthreads();
    if (threadIdx.x < 2 * QN_R) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < QN_T / 32; ++w) v = v + red[threadIdx.x][w];
      const int r = threadIdx.x % QN_R;
      if (r < rows_here) {
        if (P2P) {
          // fused all-gather: the row sum goes straight into every rank's exchange region (NVLink peer stores,
          // `par` double-buffers the region across iterations).  The stores are posted here and fenced ONCE per
          // CTA below, so their latency overlaps with the streaming of the other CTAs.
          const int64_t off = (int64_t)(par * 2 + (threadIdx.x < QN_R ? 0 : 1)) * XCHG_LD + row0 + r0 + r;
          for (int pr = 0; pr < a.world; ++pr) a.peers[pr][off] = v;
        } else {
          if (threadIdx.x < QN_R) a.h[row0 + r0 + r] = v;
          else a.w[row0 + r0 + r] = v;
        }
      }
      if (!P2P) __threadfence();
    }
    __syncthreads();
  }
  if (DEFER && !P2P) {  // the next cluster head runs the epilogue; the kernel boundary orders h, w before it
    if (blockIdx.x == 0 && threadIdx.x == 0) st->epi = 1;
    return;
  }
  if (a.ticket == nullptr) return;
  if (P2P && threadIdx.x < 2 * QN_R) __threadfence_system();  // this CTA's peer stores are performed
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(a.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (P2P) {
    // every local CTA has pushed and fenced its rows: publish "rank `a.rank` reached `seq`" on every rank,
    // then wait until all ranks have published the same sequence number here
    __threadfence_system();
    unsigned long long* myflags = reinterpret_cast<unsigned long long*>(a.peers[a.rank] + 4 * XCHG_LD);
    if (threadIdx.x < a.world) {
      unsigned long long* f = reinterpret_cast<unsigned long long*>(a.peers[threadIdx.x] + 4 * XCHG_LD) + a.rank;
      st_release_sys(f, seq);
      if (!wait_flag_sys(myflags + threadIdx.x, seq)) peer_timeout(st);
    }
    __syncthreads();
    if (DEFER) {  // all ranks' rows have arrived in this rank's exchange buffers of parity `par`
      if (threadIdx.x == 0) {
        st->epi = 1 + par;
        *a.seq = seq;
        *a.ticket = 0u;
      }
      return;
    }
    if (!DEFER)
      lazy_epilogue_body<KIND>(a, a.peers[a.rank] + (int64_t)(par * 2 + 0) * XCHG_LD,
                               a.peers[a.rank] + (int64_t)(par * 2 + 1) * XCHG_LD, &red[0][0]);
    if (threadIdx.x == 0) {
      *a.seq = seq;
      *a.ticket = 0u;
    }
    return;
  }
  if (!DEFER) lazy_epilogue_body<KIND>(a, a.h, a.w, &red[0][0]);
  if (threadIdx.x == 0) *a.ticket = 0u;
}

// ---- lazy pass, TMA-staged variant (qn_kernel = 1) ----------------------------------------------
// Same arithmetic as qn_lazy_kernel, different data movement: a producer thread streams row tiles of H
// into a ring of shared-memory stages with bulk asynchronous copies (cp.async.bulk + mbarrier
// complete_tx, SASS UBLKCP), so the bytes in flight per SM (3 stages x 64 KiB) no longer depend on
// registers; 8 consumer warps read the stage with conflict-free 128-bit LDS, apply the pending update,
// accumulate the two dot products for 16 rows per tile (the O(n) vectors are re-read once per 16 rows
// instead of once per 8) and store the new rows straight to global memory.
constexpr int TM_R = 16;              // rows per tile
constexpr int TM_CW = 512;            // columns per stage (4 KiB per row segment)
constexpr int TM_STAGES = 3;
constexpr int TM_CONS = 256;          // consumer threads: one column pair each
constexpr int TM_T = TM_CONS + 32;    // + producer warp
constexpr size_t TM_SMEM = (size_t)TM_STAGES * TM_R * TM_CW * sizeof(double) + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(TM_T, 1) qn_lazy_tma_kernel(const __grid_constant__ QNLazyArgs a) {
  DevState* st = a.st;
  if (st->done) return;
  extern __shared__ __align__(128) unsigned char tm_smem[];
  double* stage = reinterpret_cast<double*>(tm_smem);                          // TM_STAGES x TM_R x TM_CW
  uint64_t* full = reinterpret_cast<uint64_t*>(tm_smem + (size_t)TM_STAGES * TM_R * TM_CW * sizeof(double));
  uint64_t* empty = full + TM_STAGES;
  __shared__ double red[2 * TM_R][TM_CONS / 32];
  __shared__ double2 rowpq[TM_R];
  __shared__ bool is_last;
  const double c0 = st->pc0, c1 = st->pc1, c2 = st->pc2;
  const unsigned long long pol = l2_evict_first_policy();
  const int64_t ld = a.ld, nrows = a.nrows, row0 = a.row0;
  const double* __restrict__ p = a.ps;
  const double* __restrict__ q = a.ph;
  const double* __restrict__ yv = a.y;
  const double* __restrict__ gv = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long seq = a.peers != nullptr ? *a.seq + 1ULL : 0ULL;
  const int par = (int)(seq & 1ULL);
  if (tid == 0) {
    for (int s_ = 0; s_ < TM_STAGES; ++s_) {
      mbar_init(&full[s_], 1);
      mbar_init(&empty[s_], TM_CONS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t ntiles = (nrows + TM_R - 1) / TM_R;  // H is allocated with rows padded to a multiple of 8; TM_R = 16 needs a guard
  const int nchunks = (int)((ld + TM_CW - 1) / TM_CW);
  if (warp == TM_CONS / 32) {
    // ===== producer warp: lane 0 streams (tile, chunk) stages; the warp stays converged for the barriers =====
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t r0 = tile * TM_R;
      const int rows_here = (int)((nrows - r0) < TM_R ? (nrows - r0) : TM_R);
      for (int c = 0; c < nchunks; ++c, ++it) {
        if (lane == 0) {
          const int s_ = it % TM_STAGES;
          const uint32_t ph = (uint32_t)((it / TM_STAGES) & 1);
          mbar_wait(&empty[s_], ph ^ 1u);  // fresh barrier: the wait for parity 1 passes immediately
          const int cw = (int)((ld - (int64_t)c * TM_CW) < TM_CW ? (ld - (int64_t)c * TM_CW) : TM_CW);
          const uint32_t bytes_row = (uint32_t)(cw * sizeof(double));
          mbar_expect_tx(&full[s_], bytes_row * (uint32_t)rows_here);
          double* dst = stage + (size_t)s_ * TM_R * TM_CW;
          for (int r = 0; r < rows_here; ++r)
            bulk_load(dst + (size_t)r * TM_CW, a.M + (r0 + r) * ld + (int64_t)c * TM_CW, bytes_row, &full[s_]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== consumers =====
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t r0 = tile * TM_R;
      const int rows_here = (int)((nrows - r0) < TM_R ? (nrows - r0) : TM_R);
      double ah[TM_R], aw[TM_R];
#pragma unroll
      for (int r = 0; r < TM_R; ++r) ah[r] = aw[r] = 0.0;
      asm volatile("bar.sync 1, %0;" ::"r"(TM_CONS));  // rowpq / red reuse across tiles
      if (tid < TM_R) {
        const bool ok = tid < rows_here;
        rowpq[tid] = make_double2(ok ? p[row0 + r0 + tid] : 0.0, ok ? q[row0 + r0 + tid] : 0.0);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(TM_CONS));
      // the O(n) vectors come from L2 (they never survive in L1 next to the H stream): their loads are
      // issued one chunk ahead so that the latency hides behind the previous chunk's arithmetic
      double2 gn = make_double2(0.0, 0.0), yn = gn, pn = gn, qn = gn;
      {
        const int col0 = 2 * tid;
        if (col0 < (int)ld) {
          gn = *reinterpret_cast<const double2*>(gv + col0);
          yn = *reinterpret_cast<const double2*>(yv + col0);
          pn = *reinterpret_cast<const double2*>(p + col0);
          qn = *reinterpret_cast<const double2*>(q + col0);
        }
      }
      for (int c = 0; c < nchunks; ++c, ++it) {
        const int s_ = it % TM_STAGES;
        const uint32_t ph = (uint32_t)((it / TM_STAGES) & 1);
        const int col = c * TM_CW + 2 * tid;
        const bool colok = col < (int)ld;
        const double2 gj = gn, yj = yn, pj = pn, qj = qn;
        {
          const int coln = col + TM_CW;
          if (c + 1 < nchunks && coln < (int)ld) {
            gn = *reinterpret_cast<const double2*>(gv + coln);
            yn = *reinterpret_cast<const double2*>(yv + coln);
            pn = *reinterpret_cast<const double2*>(p + coln);
            qn = *reinterpret_cast<const double2*>(q + coln);
          }
        }
        mbar_wait(&full[s_], ph);
        const double* src = stage + (size_t)s_ * TM_R * TM_CW + 2 * tid;
        if (colok) {
#pragma unroll
          for (int r = 0; r < TM_R; ++r) {
            if (r < rows_here) {
              const double2 hv = *reinterpret_cast<const double2*>(src + (size_t)r * TM_CW);
              const double2 pq = rowpq[r];
              const double pi = pq.x, qi = pq.y;
              double2 hn;
              if (KIND == QN_BFGS) {
                const double cx = pi * qj.x + qi * pj.x, cy = pi * qj.y + qi * pj.y;
                hn.x = fma(c0, pi * pj.x, fma(c1, cx, hv.x));
                hn.y = fma(c0, pi * pj.y, fma(c1, cy, hv.y));
              } else {
                hn.x = fma(c2, qi * qj.x, fma(c0, pi * pj.x, hv.x));
                hn.y = fma(c2, qi * qj.y, fma(c0, pi * pj.y, hv.y));
              }
              ah[r] = fma(hn.x, yj.x, ah[r]);
              ah[r] = fma(hn.y, yj.y, ah[r]);
              aw[r] = fma(hn.x, gj.x, aw[r]);
              aw[r] = fma(hn.y, gj.y, aw[r]);
              st_stream_ef(a.M + (r0 + r) * ld + col, hn, pol);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s_]);
      }
      // row sums of the tile: shuffle trees, then fixed-order cross-warp fold
#pragma unroll
      for (int r = 0; r < TM_R; ++r) {
        const double v1 = warp_sum(ah[r]), v2 = warp_sum(aw[r]);
        if (lane == 0) {
          red[r][warp] = v1;
          red[TM_R + r][warp] = v2;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(TM_CONS));
      if (tid < 2 * TM_R) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < TM_CONS / 32; ++w) v = v + red[tid][w];
        const int r = tid % TM_R;
        if (r < rows_here) {
          if (tid < TM_R) a.h[row0 + r0 + r] = v;
          else a.w[row0 + r0 + r] = v;
        }
        __threadfence();
      }
    }
  }
  __syncthreads();
  if (a.ticket == nullptr) return;
  if (tid == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(a.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (a.peers != nullptr) {
    for (int pr = 0; pr < a.world; ++pr) {
      double* dst_h = a.peers[pr] + (int64_t)(par * 2 + 0) * XCHG_LD + row0;
      double* dst_w = a.peers[pr] + (int64_t)(par * 2 + 1) * XCHG_LD + row0;
      for (int64_t i = 2 * tid; i < nrows; i += 2 * TM_T) {
        *reinterpret_cast<double2*>(dst_h + i) = *reinterpret_cast<const double2*>(a.h + row0 + i);
        *reinterpret_cast<double2*>(dst_w + i) = *reinterpret_cast<const double2*>(a.w + row0 + i);
      }
    }
    __threadfence_system();
    __syncthreads();
    unsigned long long* myflags = reinterpret_cast<unsigned long long*>(a.peers[a.rank] + 4 * XCHG_LD);
    if (tid < a.world) {
      unsigned long long* f = reinterpret_cast<unsigned long long*>(a.peers[tid] + 4 * XCHG_LD) + a.rank;
      st_release_sys(f, seq);
      if (!wait_flag_sys(myflags + tid, seq)) peer_timeout(st);
    }
    __syncthreads();
    lazy_epilogue_body<KIND>(a, a.peers[a.rank] + (int64_t)(par * 2 + 0) * XCHG_LD,
                             a.peers[a.rank] + (int64_t)(par * 2 + 1) * XCHG_LD, &red[0][0]);
    if (tid == 0) {
      *a.seq = seq;
      *a.ticket = 0u;
    }
    return;
  }
  lazy_epilogue_body<KIND>(a, a.h, a.w, &red[0][0]);
  if (tid == 0) *a.ticket = 0u;
}

// ---- lazy pass on PACKED SYMMETRIC storage (qn_storage = 1) ---------------------------------------
// BFGS / DFP keep H exactly symmetric (every update term is symmetric in (i, j)), so only the lower
// triangle needs to live in HBM: the pass reads and writes n^2/2 elements, i.e. n^2 * 8 B of traffic per
// iteration instead of 2 n^2 * 8 B.
// Layout: row tiles of 8 rows; tile T (rows 8T .. 8T+7) stores, for each of its rows, the columns
// 0 .. 8T+7 (the 8x8 diagonal block is stored in full), padded to a multiple of 16 doubles; tiles are
// consecutive.  A stored element M_ij contributes to the row sum of i (as before) and — when it lies
// strictly left of the diagonal block — to the row sum of j (M_ji = M_ij).  The second contribution is
// a COLUMN sum: the thread that owns the column accumulates it over the 8 rows of the tile in
// registers and adds it to a per-CTA partial vector (L2-resident, 148 x n x 2 doubles); a fold kernel
// adds the per-CTA partials in CTA order (deterministic) to the row sums.
int64_t qn_sym_doubles(int64_t n) { return sym_tile_offset((n + 7) / 8) + 8 * sym_lpad((n + 7) / 8); }

int64_t qn_sym_doubles_sharded(int64_t n, int world, int rank) {
  const int64_t T = (n + QN_R - 1) / QN_R;
  return symsh_local_pairs(T, world, rank) * symsh_pair_doubles(T);
}

// host-side description of the packed layouts (tests, tools): where tile `tile` of an n x n matrix lives
void qn_sym_layout(int64_t n, int world, int64_t tile, int* owner, int64_t* offset, int64_t* lpad) {
  const int64_t T = (n + QN_R - 1) / QN_R;
  *lpad = sym_lpad(tile);
  if (world <= 1) {
    *owner = 0;
    *offset = sym_tile_offset(tile);
  } else {
    const int64_t pairi = tile < T / 2 ? tile : T - 1 - tile;
    *owner = (int)(pairi % world);
    *offset = symsh_tile_offset(tile, T, world);
  }
}

// the streaming pass as its own launch (host-driven engine, profiling, pass variants)
template <int KIND, bool SHARDED, int NT, bool OOP, bool ZERO, bool IDENT = false>
__global__ void __launch_bounds__(NT, 512 / NT) qn_lazy_sym_kernel(QNLazyArgs a, QNSymArgs sa) {
  pdl_wait();
  pdl_launch_dependents();  // (one CTA per SM, one wave)
  DevState* st = a.st;
  if (st->done) return;
  const int pp = OOP ? st->pp : 0;  // which buffer holds the current matrix (toggled by the fold kernel)
  sym_pass_body<KIND, SHARDED, NT, OOP, ZERO, IDENT>(a, sa, st->pc0, st->pc1, st->pc2, pp, (int)gridDim.x, (int)blockIdx.x);
}

// ---- the packed pass with a shared-memory ring (qn_kernel bit 3) ----------------------------------------------------
// Same tiles, same thread -> column mapping and the same order of every sum as sym_pass_body (hence the same bits); what
// changes is who waits for HBM.  In the register-staged pass a warp issues the 8 row loads of a column step, waits for
// them, computes, stores and only then issues the next step's loads: with 16 warps per SM and nothing in flight across a
// step border the SM idles for one memory latency per step (~0.8 us against ~2.9 us of transfer).  Here thread 0
// streams the CTA's (tile, column step) sequence into a ring of 3 stages of 8 rows x 1024 columns (64 KiB each) with bulk
// asynchronous copies (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP), two steps ahead of the arithmetic and
// straight across tile borders; the 16 warps read a stage with conflict-free 128-bit LDS, apply the pending update,
// accumulate row and column sums and store the new elements to global memory from registers.
// MEASURED (profiles/r02_packed_pass_experiments.md): 0.416 ms against 0.397 ms for the register-staged pass at n = 16384 —
// bytes in flight are not what limits the pass; the variant stays selectable (bit-identical, tested) as the evidence.
constexpr int SR_T = 512;                // one column pair of the stage per thread; thread 0 doubles as the producer
                                         // (a 17th warp would cap the kernel at 96 registers)
constexpr int SR_CW = 2 * SR_T;          // columns per stage
constexpr int SR_STAGES = 3;
constexpr size_t SR_STAGE_DOUBLES = (size_t)QN_R * SR_CW;
constexpr size_t SR_SMEM = SR_STAGES * SR_STAGE_DOUBLES * sizeof(double) + 128;

__device__ __forceinline__ void bulk_load_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}

template <int KIND, bool SHARDED>
__global__ void __launch_bounds__(SR_T, 1) qn_sym_ring_kernel(QNLazyArgs a, QNSymArgs sa) {
  DevState* st = a.st;
  if (st->done) return;
  extern __shared__ __align__(128) unsigned char sr_smem[];
  double* stage = reinterpret_cast<double*>(sr_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(sr_smem + SR_STAGES * SR_STAGE_DOUBLES * sizeof(double));
  uint64_t* empty = full + SR_STAGES;
  __shared__ double red2[2][SR_T / 32][16];
  __shared__ double4 rowv2[2][QN_R];
  const double c0 = st->pc0, c1 = st->pc1, c2 = st->pc2;
  const unsigned long long pol = l2_evict_first_policy();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
  const int64_t n = sa.n, ld = sa.ld;
  const int64_t ntiles = (n + QN_R - 1) / QN_R;
  if (tid == 0) {
    for (int s_ = 0; s_ < SR_STAGES; ++s_) {
      mbar_init(&full[s_], 1);
      mbar_init(&empty[s_], SR_T / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // ===== producer state (thread 0): walks the same (tile, column step) sequence, SR_STAGES - 1 steps ahead =====
  int p_step = -1, p_cb = 0, p_lpad = 0, p_rows = 0, p_it = 0;
  const double* p_base = nullptr;
  bool p_done = false;
  auto produce = [&]() {
    if (p_done) return;
    if (p_cb >= p_lpad) {  // next tile
      for (;;) {
        ++p_step;
        const int64_t tile = sym_cta_tile<SHARDED>(ntiles, sa.world, sa.rank, grid, cta, p_step);
        if (tile < 0) {
          if (SHARDED || (p_step & 1) == 0) {
            p_done = true;
            return;
          }
          continue;
        }
        const int64_t r0 = tile * QN_R;
        p_rows = (int)((n - r0) < QN_R ? (n - r0) : QN_R);
        p_lpad = (int)sym_lpad(tile);
        p_base = sa.P + (SHARDED ? symsh_tile_offset(tile, ntiles, sa.world) : sym_tile_offset(tile));
        p_cb = 0;
        break;
      }
    }
    const int s_ = p_it % SR_STAGES;
    mbar_wait(&empty[s_], (uint32_t)(((p_it / SR_STAGES) & 1) ^ 1));  // (fresh barrier: parity 1 passes at once)
    const int cw = p_lpad - p_cb < SR_CW ? p_lpad - p_cb : SR_CW;
    const uint32_t bytes_row = (uint32_t)(cw * sizeof(double));
    mbar_expect_tx(&full[s_], bytes_row * (uint32_t)p_rows);
    double* dst = stage + (size_t)s_ * SR_STAGE_DOUBLES;
    for (int r = 0; r < p_rows; ++r) bulk_load_hint(dst + (size_t)r * SR_CW, p_base + (int64_t)r * p_lpad + p_cb, bytes_row, &full[s_], pol);
    p_cb += SR_CW;
    ++p_it;
  };
  if (tid == 0) {
#pragma unroll 1
    for (int k = 0; k < SR_STAGES - 1; ++k) produce();
  }
  // ===== the arithmetic =====
  const double* __restrict__ p = a.ps;
  const double* __restrict__ q = a.ph;
  const double* __restrict__ yv = a.y;
  const double* __restrict__ gv = a.g;
  double* __restrict__ cph = sa.colpart + (int64_t)cta * 2 * ld;
  double* __restrict__ cpw = cph + ld;
  int tpar = 0, it = 0;
  bool first = true;
  for (int64_t step = 0;; ++step) {
    const int64_t tile = sym_cta_tile<SHARDED>(ntiles, sa.world, sa.rank, grid, cta, step);
    if (tile < 0) {
      if (SHARDED || (step & 1) == 0) break;
      continue;
    }
    const int64_t r0 = tile * QN_R;
    const int rows_here = (int)((n - r0) < QN_R ? (n - r0) : QN_R);
    const int lpad = (int)sym_lpad(tile);
    const int ncols = (int)(r0 + QN_R < n ? r0 + QN_R : n);
    double ah[QN_R], aw[QN_R];
#pragma unroll
    for (int r = 0; r < QN_R; ++r) ah[r] = aw[r] = 0.0;
    tpar ^= 1;
    double4* rowv = rowv2[tpar];
    double (*red)[16] = red2[tpar];
    if (tid < QN_R) {
      const bool ok = tid < rows_here;
      const int64_t i = r0 + tid;
      rowv[tid] = ok ? make_double4(p[i], q[i], yv[i], gv[i]) : make_double4(0.0, 0.0, 0.0, 0.0);
    }
    __syncthreads();  // (A)
    double* __restrict__ obase = sa.P + (SHARDED ? symsh_tile_offset(tile, ntiles, sa.world) : sym_tile_offset(tile));
    // the O(n) vectors of a step come from L2: loaded one step ahead, behind the previous step's arithmetic
    double2 gn = make_double2(0.0, 0.0), yn = gn, pn = gn, qn = gn;
    if (2 * tid < ncols) {
      gn = ld_vec2(gv + 2 * tid);
      yn = ld_vec2(yv + 2 * tid);
      pn = ld_vec2(p + 2 * tid);
      qn = ld_vec2(q + 2 * tid);
    }
    for (int cb = 0; cb < lpad; cb += SR_CW, ++it) {
      const int s_ = it % SR_STAGES;
      const int col = cb + 2 * tid;
      const bool v0 = col < ncols, v1 = col + 1 < ncols;
      const bool cok = v0 && col < (int)r0;
      const double2 gj = gn, yj = yn, pj = pn, qj = qn;
      double2 oh = make_double2(0.0, 0.0), ow = make_double2(0.0, 0.0);
      if (cok && !first) {
        oh = *reinterpret_cast<double2*>(cph + col);
        ow = *reinterpret_cast<double2*>(cpw + col);
      }
      if (col + SR_CW < ncols) {
        gn = ld_vec2(gv + col + SR_CW);
        yn = ld_vec2(yv + col + SR_CW);
        pn = ld_vec2(p + col + SR_CW);
        qn = ld_vec2(q + col + SR_CW);
      }
      if (tid == 0) produce();  // the stage the CTA finished one step ago is refilled with the step after next
      __syncwarp();
      mbar_wait(&full[s_], (uint32_t)((it / SR_STAGES) & 1));
      if (v0) {
        const double* src = stage + (size_t)s_ * SR_STAGE_DOUBLES + 2 * tid;
        double ch0 = 0.0, ch1 = 0.0, cw0 = 0.0, cw1 = 0.0;
#pragma unroll
        for (int r = 0; r < QN_R; ++r) {
          if (r < rows_here) {
            const double2 hv = *reinterpret_cast<const double2*>(src + (size_t)r * SR_CW);
            const double4 rv = rowv[r];
            const double pi = rv.x, qi = rv.y;
            double2 hn;
            if (KIND == QN_BFGS) {
              const double cx = pi * qj.x + qi * pj.x, cy = pi * qj.y + qi * pj.y;
              hn.x = fma(c0, pi * pj.x, fma(c1, cx, hv.x));
              hn.y = fma(c0, pi * pj.y, fma(c1, cy, hv.y));
            } else {
              hn.x = fma(c2, qi * qj.x, fma(c0, pi * pj.x, hv.x));
              hn.y = fma(c2, qi * qj.y, fma(c0, pi * pj.y, hv.y));
            }
            if (!v1) hn.y = 0.0;
            ah[r] = fma(hn.x, yj.x, ah[r]);
            ah[r] = fma(hn.y, yj.y, ah[r]);
            aw[r] = fma(hn.x, gj.x, aw[r]);
            aw[r] = fma(hn.y, gj.y, aw[r]);
            ch0 = fma(hn.x, rv.z, ch0);
            ch1 = fma(hn.y, rv.z, ch1);
            cw0 = fma(hn.x, rv.w, cw0);
            cw1 = fma(hn.y, rv.w, cw1);
            st_stream_ef(obase + (int64_t)r * lpad + col, hn, pol);
          }
        }
        if (cok) {
          oh.x += ch0;
          ow.x += cw0;
          oh.y += ch1;
          ow.y += cw1;
          *reinterpret_cast<double2*>(cph + col) = oh;
          *reinterpret_cast<double2*>(cpw + col) = ow;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s_]);  // this warp's reads of the stage are done
    }
    first = false;
    {
      double v16[16];
#pragma unroll
      for (int r = 0; r < QN_R; ++r) {
        v16[r] = ah[r];
        v16[QN_R + r] = aw[r];
      }
      const double ws = warp_sum16(v16);
      if ((lane & 1) == 0) red[warp][lane >> 1] = ws;
    }
    __syncthreads();  // (B)
    if (tid < 2 * QN_R) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < SR_T / 32; ++w) v = v + red[w][tid];
      const int r = tid % QN_R;
      if (r < rows_here) {
        if (tid < QN_R) a.h[r0 + r] = v;
        else a.w[r0 + r] = v;
      }
    }
  }
}

// h_j += sum over CTAs of the column partials; then, depending on the mode, the O(n) epilogue by the last CTA (host
// engine), nothing (the next cluster head runs it), or — sharded — the push of this rank's contribution to every peer.
// Fixed summation shape: FOLD_G groups of consecutive CTAs summed in order, then the groups in order.
// CTA = 16 warps = {h, w} x FOLD_G groups, lane = 2 columns of a block of 64: every load is a coalesced 512-byte piece
// of one partial vector, 8 loads in flight per thread (a one-thread-per-column version took 58 us, this one 6).
constexpr int FOLD_T = 512;   // = QN_T: the fused epilogue then adds in the same order as the full-storage kernel's
constexpr int FOLD_G = FOLD_T / 64;  // groups of partial rows per vector
constexpr int FOLD_MAXPARTS = 1024;  // >= the largest pass grid (2 CTAs per SM)
template <int KIND>
__global__ void __launch_bounds__(FOLD_T) qn_sym_fold_kernel(const __grid_constant__ QNLazyArgs a, QNSymArgs sa, int nparts, unsigned int* ticket) {
  pdl_wait();  // (no early trigger here: this grid runs in two waves)
  DevState* st = a.st;
  if (st->done) return;
  if (sa.Pout != sa.P && blockIdx.x == 0 && threadIdx.x == 0) st->pp ^= 1;  // ping-pong: the pass wrote the other buffer
  __shared__ double2 part[2][FOLD_G][32];
  __shared__ double smem[3 * 32];
  __shared__ bool is_last;
  const int64_t ld = sa.ld;
  const unsigned long long seq = sa.world > 1 ? *sa.seq + 1ULL : 0ULL;  // this exchange's sequence number
  const int par = (int)(seq & 1ULL);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int vec = warp / FOLD_G, grp = warp % FOLD_G;
  const int per = (nparts + FOLD_G - 1) / FOLD_G;
  const int cb = grp * per, ce = (cb + per < nparts) ? cb + per : nparts;
  // partial vector c is valid on the columns [0, ext_s[c]) only (the pass never wrote the rest)
  __shared__ int ext_s[FOLD_MAXPARTS];
  {
    const int64_t T = (sa.n + QN_R - 1) / QN_R;
    for (int c = threadIdx.x; c < nparts; c += FOLD_T)
      ext_s[c] = sa.zeroed ? (int)ld
                           : (int)(sa.world > 1 ? sym_first_row<true>(T, sa.world, sa.rank, nparts, c) : sym_first_row<false>(T, 1, 0, nparts, c));
    __syncthreads();
  }
  // a warp reads 512 contiguous bytes of one partial row per load and keeps 8 loads in flight; the order of the
  // additions is fixed by (nparts, FOLD_G groups), not by timing
  for (int64_t j0 = (int64_t)blockIdx.x * 64; j0 < sa.n; j0 += (int64_t)gridDim.x * 64) {
    const int64_t j = j0 + 2 * lane;  // ld is a multiple of 8 and the pad columns of colpart are zero, so j + 1 < ld is readable
    double2 acc = make_double2(0.0, 0.0);
    if (j < sa.n) {
      const double* src = sa.colpart + (int64_t)vec * ld + j;
      int c = cb;
      for (; c + 8 <= ce; c += 8) {
        double2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          v[k] = j < ext_s[c + k] ? __ldcg(reinterpret_cast<const double2*>(src + (int64_t)(c + k) * 2 * ld)) : make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc.x = acc.x + v[k].x;
          acc.y = acc.y + v[k].y;
        }
      }
      for (; c < ce; ++c) {
        const double2 v = j < ext_s[c] ? __ldcg(reinterpret_cast<const double2*>(src + (int64_t)c * 2 * ld)) : make_double2(0.0, 0.0);
        acc.x = acc.x + v.x;
        acc.y = acc.y + v.y;
      }
    }
    __syncthreads();
    part[vec][grp][lane] = acc;
    __syncthreads();
    if (grp == 0 && j < sa.n) {
      double2 tot = part[vec][0][lane];
#pragma unroll
      for (int g = 1; g < FOLD_G; ++g) {
        tot.x = tot.x + part[vec][g][lane].x;
        tot.y = tot.y + part[vec][g][lane].y;
      }
      double* dst = vec == 0 ? a.h : a.w;
      const double vx = dst[j] + tot.x;
      const double vy = (j + 1 < sa.n) ? dst[j + 1] + tot.y : 0.0;
      if (sa.world > 1) {
        // this rank's contribution (its tiles' row sums + its column sums) goes straight into slot `rank` of every
        // rank's exchange region; the head adds the slots in rank order
        const int64_t off = XSLOT_OFF + ((int64_t)(par * sa.world + sa.rank) * 2 + vec) * XSLOT_LD + j;
        for (int pr = 0; pr < sa.world; ++pr) *reinterpret_cast<double2*>(sa.peers[pr] + off) = make_double2(vx, vy);
      } else {
        dst[j] = vx;
        if (j + 1 < sa.n) dst[j + 1] = vy;
      }
    }
  }
  if (sa.world > 1) {
    // same protocol as qn_lazy_kernel<.., P2P>: fence this CTA's peer stores, the last CTA publishes the sequence
    // number on every rank and waits for every rank's; the epilogue is always left to the head
    if (grp == 0) __threadfence_system();  // only the two warps that stored to the peers
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int t = atomicAdd(ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence_system();
    unsigned long long* myflags = reinterpret_cast<unsigned long long*>(sa.peers[sa.rank] + 4 * XCHG_LD);
    if (threadIdx.x < sa.world) {
      unsigned long long* f = reinterpret_cast<unsigned long long*>(sa.peers[threadIdx.x] + 4 * XCHG_LD) + sa.rank;
      st_release_sys(f, seq);
      if (!wait_flag_sys(myflags + threadIdx.x, seq)) peer_timeout(st);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      st->epi = 1 + par;
      *sa.seq = seq;
      *ticket = 0u;
    }
    return;
  }
  if (a.defer_epi) {  // the next cluster head forms the coefficients and u on 8 SMs; the kernel boundary orders h, w before it
    if (blockIdx.x == 0 && threadIdx.x == 0) st->epi = 1;
    return;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  lazy_epilogue_body<KIND>(a, a.h, a.w, smem);
  if (threadIdx.x == 0) *ticket = 0u;
}

// full (row-major, ld) <-> packed conversions
__global__ void __launch_bounds__(256) qn_sym_pack_kernel(const double* __restrict__ H, int64_t ld, int64_t n, double* __restrict__ P) {
  const int64_t ntiles = (n + QN_R - 1) / QN_R;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * QN_R, lpad = sym_lpad(tile);
    const int64_t ncols = r0 + QN_R < n ? r0 + QN_R : n;
    double* base = P + sym_tile_offset(tile);
    for (int64_t e = threadIdx.x; e < QN_R * lpad; e += blockDim.x) {
      const int64_t r = e / lpad, c = e % lpad;
      base[e] = (r0 + r < n && c < ncols) ? H[(r0 + r) * ld + c] : 0.0;
    }
  }
}
__global__ void __launch_bounds__(256) qn_sym_unpack_kernel(const double* __restrict__ P, int64_t ld, int64_t n, double* __restrict__ H) {
  const int64_t ntiles = (n + QN_R - 1) / QN_R;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * QN_R, lpad = sym_lpad(tile);
    const int64_t ncols = r0 + QN_R < n ? r0 + QN_R : n;
    const double* base = P + sym_tile_offset(tile);
    for (int64_t e = threadIdx.x; e < QN_R * lpad; e += blockDim.x) {
      const int64_t r = e / lpad, c = e % lpad;
      if (r0 + r < n && c < ncols) {
        const double v = base[e];
        H[(r0 + r) * ld + c] = v;
        if (c < r0) H[c * ld + r0 + r] = v;  // mirror (the diagonal block is stored in full)
      }
    }
  }
}

__global__ void qn_sym_identity_kernel(int64_t n, double* __restrict__ P) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tile = i / QN_R;
    P[sym_tile_offset(tile) + (i % QN_R) * sym_lpad(tile) + i] = 1.0;  // bfgs.rs:30-33: H_0 = I, straight into the packed layout
  }
}
__global__ void qn_symsh_identity_kernel(int64_t n, double* __restrict__ P, int world, int rank) {
  const int64_t T = (n + QN_R - 1) / QN_R;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tile = i / QN_R;
    const int64_t pairi = tile < T / 2 ? tile : T - 1 - tile;
    if (pairi % world == rank) P[symsh_tile_offset(tile, T, world) + (i % QN_R) * sym_lpad(tile) + i] = 1.0;
  }
}
void qn_sym_set_identity_sharded(Ctx* ctx, int64_t n, double* P) {
  OSB_CUDA(cudaMemsetAsync(P, 0, sizeof(double) * (size_t)qn_sym_doubles_sharded(n, ctx->world, ctx->rank), ctx->stream));
  qn_symsh_identity_kernel<<<ctx->red_grid(n), RED_THREADS, 0, ctx->stream>>>(n, P, ctx->world, ctx->rank);
  ctx->counters[0]++;
}
// a FULL n x n matrix (every rank holds all of it) -> this rank's tile pairs
__global__ void __launch_bounds__(256) qn_symsh_pack_kernel(const double* __restrict__ H, int64_t ld, int64_t n, double* __restrict__ P, int world,
                                                           int rank) {
  const int64_t T = (n + QN_R - 1) / QN_R;
  const int64_t nlp = symsh_local_pairs(T, world, rank);
  for (int64_t q = blockIdx.x; q < 2 * nlp; q += gridDim.x) {
    const int64_t pairi = (q >> 1) * world + rank;
    const int64_t tile = (q & 1) == 0 ? pairi : T - 1 - pairi;
    const int64_t r0 = tile * QN_R, lpad = sym_lpad(tile);
    const int64_t ncols = r0 + QN_R < n ? r0 + QN_R : n;
    double* base = P + symsh_tile_offset(tile, T, world);
    for (int64_t e = threadIdx.x; e < QN_R * lpad; e += blockDim.x) {
      const int64_t r = e / lpad, c = e % lpad;
      base[e] = (r0 + r < n && c < ncols) ? H[(r0 + r) * ld + c] : 0.0;
    }
  }
}
void qn_sym_pack_sharded(Ctx* ctx, const double* Hfull, int64_t ld, int64_t n, double* P) {
  qn_symsh_pack_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(Hfull, ld, n, P, ctx->world, ctx->rank);
  ctx->counters[0]++;
}
// this rank's tiles -> a FULL n x n matrix (both triangles of every owned element; everything else untouched, i.e.
// zero in a zeroed buffer): the sum over ranks of these matrices is H
__global__ void __launch_bounds__(256) qn_symsh_unpack_kernel(const double* __restrict__ P, int64_t ld, int64_t n, double* __restrict__ H,
                                                             int world, int rank) {
  const int64_t T = (n + QN_R - 1) / QN_R;
  const int64_t nlp = symsh_local_pairs(T, world, rank);
  for (int64_t q = blockIdx.x; q < 2 * nlp; q += gridDim.x) {
    const int64_t pairi = (q >> 1) * world + rank;
    const int64_t tile = (q & 1) == 0 ? pairi : T - 1 - pairi;
    const int64_t r0 = tile * QN_R, lpad = sym_lpad(tile);
    const int64_t ncols = r0 + QN_R < n ? r0 + QN_R : n;
    const double* base = P + symsh_tile_offset(tile, T, world);
    for (int64_t e = threadIdx.x; e < QN_R * lpad; e += blockDim.x) {
      const int64_t r = e / lpad, c = e % lpad;
      if (r0 + r < n && c < ncols) {
        const double v = base[e];
        H[(r0 + r) * ld + c] = v;
        if (c < r0) H[c * ld + r0 + r] = v;
      }
    }
  }
}
void qn_sym_unpack_sharded(Ctx* ctx, const double* P, int64_t ld, int64_t n, double* Hfull_zeroed) {
  qn_symsh_unpack_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(P, ld, n, Hfull_zeroed, ctx->world, ctx->rank);
  ctx->counters[0]++;
}
void qn_sym_set_identity(Ctx* ctx, int64_t n, double* P) {
  OSB_CUDA(cudaMemsetAsync(P, 0, sizeof(double) * (size_t)qn_sym_doubles(n), ctx->stream));
  qn_sym_identity_kernel<<<ctx->red_grid(n), RED_THREADS, 0, ctx->stream>>>(n, P);
  ctx->counters[0]++;
}
void qn_sym_pack(Ctx* ctx, const double* H, int64_t ld, int64_t n, double* P) {
  qn_sym_pack_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(H, ld, n, P);
  ctx->counters[0]++;
}
void qn_sym_unpack(Ctx* ctx, const double* P, int64_t ld, int64_t n, double* H) {
  qn_sym_unpack_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(P, ld, n, H);
  ctx->counters[0]++;
}
// pass variants (option "qn_kernel" with packed storage): bit 0 = two 256-thread CTAs per SM instead of one of 512,
// bit 1 = ping-pong storage (read P, write Pout), bit 2 = legacy zero-first column partials
int qn_sym_grid(Ctx* ctx, int64_t n, int variant) {
  const int64_t T = (n + QN_R - 1) / QN_R;
  const int64_t units = ctx->world > 1 ? 2 * symsh_local_pairs(T, ctx->world, ctx->rank) : T;  // sharded: local tiles
  const int per_sm = (variant & 1) ? 2 : 1;
  return (int)std::max<int64_t>(1, std::min<int64_t>(units, (int64_t)ctx->num_sms * per_sm));
}

template <int KIND, bool SHARDED>
static void launch_sym_pass(int grid, cudaStream_t stream, const QNLazyArgs& a, const QNSymArgs& sa, int variant) {
  if (variant & 8) {  // shared-memory ring fed by bulk asynchronous copies
    static bool attr = false;
    if (!attr) {
      OSB_CUDA(cudaFuncSetAttribute(qn_sym_ring_kernel<KIND, SHARDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SR_SMEM));
      attr = true;
    }
    qn_sym_ring_kernel<KIND, SHARDED><<<grid, SR_T, SR_SMEM, stream>>>(a, sa);
    return;
  }
  switch (variant & 7) {
    case 0:
      if (sa.zeroed == 2) {  // the stored matrix is a still unwritten identity: generate instead of load (one GPU only)
        QNSymArgs sb = sa;
        sb.zeroed = 0;
        launch_pdl(qn_lazy_sym_kernel<KIND, false, 512, false, false, true>, dim3(grid), dim3(512), 0, stream, a, sb);
      } else {
        launch_pdl(qn_lazy_sym_kernel<KIND, SHARDED, 512, false, false>, dim3(grid), dim3(512), 0, stream, a, sa);
      }
      break;
    case 1: qn_lazy_sym_kernel<KIND, SHARDED, 256, false, false><<<grid, 256, 0, stream>>>(a, sa); break;
    case 2: qn_lazy_sym_kernel<KIND, SHARDED, 512, true, false><<<grid, 512, 0, stream>>>(a, sa); break;
    case 3: qn_lazy_sym_kernel<KIND, SHARDED, 256, true, false><<<grid, 256, 0, stream>>>(a, sa); break;
    default: qn_lazy_sym_kernel<KIND, SHARDED, 512, false, true><<<grid, 512, 0, stream>>>(a, sa); break;
  }
}

void qn_launch_lazy_sym(Ctx* ctx, const QNLazyArgs& a, double* P, double* Pout, double* colpart, int64_t n, int64_t ld, int phase, int variant,
                        bool identity_unwritten) {
  const bool sharded = ctx->world > 1;
  const int grid = qn_sym_grid(ctx, n, variant);
  OSB_REQUIRE(grid <= FOLD_MAXPARTS, OSB_ERR_UNSUPPORTED, "pass grid exceeds the fold's partial table");
  QNSymArgs sa{P, Pout, colpart, n, ld, sharded ? ctx->world : 1, sharded ? ctx->rank : 0, sharded ? ctx->d_peers : nullptr, ctx->d_seq,
               grid, (variant & 4) ? 1 : (identity_unwritten ? 2 : 0)};
  const int fgrid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 63) / 64, (int64_t)ctx->num_sms * 2));
  if (phase == 0) {  // the streaming pass over the packed triangle
    if (sharded) {
      if (a.kind == QN_BFGS) launch_sym_pass<QN_BFGS, true>(grid, ctx->stream, a, sa, variant);
      else launch_sym_pass<QN_DFP, true>(grid, ctx->stream, a, sa, variant);
    } else {
      if (a.kind == QN_BFGS) launch_sym_pass<QN_BFGS, false>(grid, ctx->stream, a, sa, variant);
      else launch_sym_pass<QN_DFP, false>(grid, ctx->stream, a, sa, variant);
    }
  } else {  // fold of the per-CTA column partials + coefficient epilogue
    if (a.kind == QN_BFGS) launch_pdl(qn_sym_fold_kernel<QN_BFGS>, dim3(fgrid), dim3(FOLD_T), 0, ctx->stream, a, sa, grid, a.ticket);
    else launch_pdl(qn_sym_fold_kernel<QN_DFP>, dim3(fgrid), dim3(FOLD_T), 0, ctx->stream, a, sa, grid, a.ticket);
  }
  ctx->counters[0] += 1;
}

template <int KIND>
__global__ void __launch_bounds__(QN_T) qn_lazy_epilogue_kernel(const __grid_constant__ QNLazyArgs a) {
  if (a.st->done) return;
  __shared__ double smem[3 * 32];
  lazy_epilogue_body<KIND>(a, a.h, a.w, smem);
}

void qn_launch_lazy(Ctx* ctx, const QNLazyArgs& a_in, int variant) {
  QNLazyArgs a = a_in;
  a.tile_rows = qn_pick_tile_rows(a.nrows, (int)std::min<int64_t>((a.nrows + QN_R - 1) / QN_R, (int64_t)ctx->num_sms));
  if (variant == 1) {
    static bool attr = false;
    if (!attr) {
      OSB_CUDA(cudaFuncSetAttribute(qn_lazy_tma_kernel<QN_BFGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM));
      OSB_CUDA(cudaFuncSetAttribute(qn_lazy_tma_kernel<QN_DFP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM));
      attr = true;
    }
    int64_t nt = (a.nrows + TM_R - 1) / TM_R;
    int g = (int)std::min<int64_t>(nt, (int64_t)ctx->num_sms);
    if (a.kind == QN_BFGS) qn_lazy_tma_kernel<QN_BFGS><<<g, TM_T, TM_SMEM, ctx->stream>>>(a);
    else qn_lazy_tma_kernel<QN_DFP><<<g, TM_T, TM_SMEM, ctx->stream>>>(a);
    ctx->counters[0]++;
    return;
  }
  int64_t ntiles = (a.nrows + QN_R - 1) / QN_R;
  int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx->num_sms);
  const bool p2p = a.peers != nullptr;
#define OSB_LAZY(K, D, P) qn_lazy_kernel<K, D, P><<<grid, QN_T, 0, ctx->stream>>>(a)
#define OSB_LAZY_K(K)                                  \
  do {                                                 \
    if (a.defer_epi) {                                 \
      if (p2p) OSB_LAZY(K, true, true);                \
      else OSB_LAZY(K, true, false);                   \
    } else {                                           \
      if (p2p) OSB_LAZY(K, false, true);               \
      else OSB_LAZY(K, false, false);                  \
    }                                                  \
  } while (0)
  if (a.kind == QN_BFGS) OSB_LAZY_K(QN_BFGS);
  else OSB_LAZY_K(QN_DFP);
#undef OSB_LAZY_K
#undef OSB_LAZY
  ctx->counters[0]++;
}
void qn_launch_lazy_epilogue(Ctx* ctx, const QNLazyArgs& a) {
  if (a.kind == QN_BFGS) qn_lazy_epilogue_kernel<QN_BFGS><<<1, QN_T, 0, ctx->stream>>>(a);
  else qn_lazy_epilogue_kernel<QN_DFP><<<1, QN_T, 0, ctx->stream>>>(a);
  ctx->counters[0]++;
}

// apply the pending update only (no products): M <- M + rank2(ps, ph; pc*)
template <int KIND>
__global__ void __launch_bounds__(QN_T, 1) qn_flush_kernel(double* __restrict__ M, int64_t ld, int64_t nrows, int64_t row0, DevState* st,
                                                          const double* __restrict__ p, const double* __restrict__ q, int rt) {
  if (!st->pending) return;
  const double c0 = st->pc0, c1 = st->pc1, c2 = st->pc2;
  __shared__ double2 rowpq[QN_R];
  for (int64_t r0 = (int64_t)blockIdx.x * rt; r0 < nrows; r0 += (int64_t)gridDim.x * rt) {
    const int64_t re = r0 + rt < nrows ? r0 + rt : nrows;
    const int rows_here = (int)(re - r0);
    __syncthreads();
    if (threadIdx.x < QN_R) {
      const bool ok = (int)threadIdx.x < rows_here;
      rowpq[threadIdx.x] = make_double2(ok ? p[row0 + r0 + threadIdx.x] : 0.0, ok ? q[row0 + r0 + threadIdx.x] : 0.0);
    }
    __syncthreads();
    double* __restrict__ base = M + r0 * ld;
    for (int col = 2 * threadIdx.x; col < (int)ld; col += QN_CHUNK) {
      const double2 pj = ld_vec2(p + col);
      const double2 qj = ld_vec2(q + col);
#pragma unroll
      for (int r = 0; r < QN_R; ++r) {
        const double2 pq = rowpq[r];
        const double pi = pq.x, qi = pq.y;
        if (r >= rows_here) continue;
        double2 hv = ld_stream(base + r * ld + col), hn;
        if (KIND == QN_BFGS) {
          const double cx = pi * qj.x + qi * pj.x, cy = pi * qj.y + qi * pj.y;
          hn.x = fma(c0, pi * pj.x, fma(c1, cx, hv.x));
          hn.y = fma(c0, pi * pj.y, fma(c1, cy, hv.y));
        } else {
          hn.x = fma(c2, qi * qj.x, fma(c0, pi * pj.x, hv.x));
          hn.y = fma(c2, qi * qj.y, fma(c0, pi * pj.y, hv.y));
        }
        st_stream(base + r * ld + col, hn);
      }
    }
  }
}
__global__ void qn_clear_pending_kernel(DevState* st) {
  st->pending = 0;
  st->pc0 = st->pc1 = st->pc2 = 0.0;
}
void qn_launch_flush(Ctx* ctx, int kind, double* M, int64_t ld, int64_t nrows, int64_t row0, DevState* st, const double* ps,
                     const double* ph) {
  int64_t ntiles = (nrows + QN_R - 1) / QN_R;
  int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx->num_sms);
  const int rt = qn_pick_tile_rows(nrows, grid);
  if (kind == QN_BFGS) qn_flush_kernel<QN_BFGS><<<grid, QN_T, 0, ctx->stream>>>(M, ld, nrows, row0, st, ps, ph, rt);
  else qn_flush_kernel<QN_DFP><<<grid, QN_T, 0, ctx->stream>>>(M, ld, nrows, row0, st, ps, ph, rt);
  qn_clear_pending_kernel<<<1, 1, 0, ctx->stream>>>(st);
  ctx->counters[0] += 2;
}

void qn_launch_update(Ctx* ctx, int kind, double* H, int64_t ld, int64_t nrows, int64_t row0, const DevState* st,
                      const double* p, const double* q, const double* r, const double* g, double* u_out, int variant) {
  (void)variant;
  int64_t ntiles = (nrows + QN_R - 1) / QN_R;
  int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx->num_sms);
  const int rt = qn_pick_tile_rows(nrows, grid);
  switch (kind) {
    case QN_BFGS: qn_update_kernel<QN_BFGS><<<grid, QN_T, 0, ctx->stream>>>(H, ld, nrows, row0, st, p, q, r, g, u_out, rt); break;
    case QN_DFP: qn_update_kernel<QN_DFP><<<grid, QN_T, 0, ctx->stream>>>(H, ld, nrows, row0, st, p, q, r, g, u_out, rt); break;
    case QN_SR1: qn_update_kernel<QN_SR1><<<grid, QN_T, 0, ctx->stream>>>(H, ld, nrows, row0, st, p, q, r, g, u_out, rt); break;
    default: qn_update_kernel<QN_BROYDEN><<<grid, QN_T, 0, ctx->stream>>>(H, ld, nrows, row0, st, p, q, r, g, u_out, rt); break;
  }
  ctx->counters[0]++;
}

}  // namespace osb
