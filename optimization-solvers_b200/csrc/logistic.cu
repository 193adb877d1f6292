// logistic.cu — synthetic l2-regularised logistic regression (BASELINE.json configs[4]):
//     f(w) = sum_i log(1 + exp(-y_i x_i.w)) + (lambda/2) ||w||^2,   X: m x n, generated on device.
// The Newton-family solvers (src/newton/mod.rs:26-49, projected_newton.rs:64-80, spn.rs:76-91) take
// f, g and the Hessian from the oracle; here the oracle is a device functor with three kernels:
//   logit_margin_kernel   one pass over X: z = X w, per-sample loss, g-coefficients, Hessian weights
//   logit_grad_*          one pass over X: g = X^T c + lambda w   (two-stage deterministic column sums)
//   syrk_dmma_kernel      H = X^T D X + lambda I as a genuine dense contraction on the FP64 tensor
//                         cores (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; tcgen05 has no FP64 kind),
//                         lower-triangular 128x128 output tiles, register-prefetched shared-memory
//                         panels, m*n^2 multiply-adds (SYRK convention).
// Labels are decided in exact integer arithmetic so every implementation sees the same problem.
#include "engine.cuh"

namespace osb {

void ctx_all_reduce_sum(Ctx* ctx, double* buf, int64_t count);

// ---- generator ------------------------------------------------------------------------------
__global__ void logit_gen_kernel(int64_t m, int64_t n, int64_t row0, double* __restrict__ X, double* __restrict__ ysign) {
  // one warp per sample row
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < m; i += nwarps) {
    const uint64_t gi = (uint64_t)(row0 + i);
    long long acc = 0;
    for (int64_t j = lane; j < n; j += 32) {
      const int xi = h16(4, gi, (uint64_t)j);
      const int wj = h16(5, (uint64_t)j, 0);
      X[i * n + j] = (double)xi * 3.0517578125e-05;  // 2^-15
      acc += (long long)xi * (long long)wj;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      acc += (long long)h16(6, gi, 0) * 8192LL;  // noise 2^-17 in units of 2^-30
      ysign[i] = acc > 0 ? 1.0 : -1.0;
    }
  }
}

// ---- margins: z_i = x_i . w ; loss_i ; c_i = -y_i sigma(-y_i z_i) ; d_i = sigma (1 - sigma) -------
__global__ void __launch_bounds__(256) logit_margin_kernel(int64_t m, int64_t n, const double* __restrict__ X,
                                                          const double* __restrict__ ysign, const double* __restrict__ w,
                                                          double* __restrict__ loss, double* __restrict__ gc,
                                                          double* __restrict__ dc) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < m; i += nwarps) {
    const double* row = X + i * n;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t j = 2 * lane; j < n; j += 64) {
      const double2 xv = ld_stream_nc(row + j);
      const double2 wv = *reinterpret_cast<const double2*>(w + j);
      a0 = fma(xv.x, wv.x, a0);
      a1 = fma(xv.y, wv.y, a1);
    }
    const double z = warp_sum(a0 + a1);
    if (lane == 0) {
      const double ys = ysign[i];
      const double u = -ys * z;  // loss = log(1 + exp(u))
      loss[i] = u > 0 ? u + log1p(exp(-u)) : log1p(exp(u));
      const double sig = 1.0 / (1.0 + exp(-u));
      gc[i] = -ys * sig;
      dc[i] = sig * (1.0 - sig);
    }
  }
}

// ---- gradient: column sums of c_i x_i over row splits, then a fixed-order fold -----------------
constexpr int LG_T = 256;
__global__ void __launch_bounds__(LG_T) logit_grad_stage1(int64_t m, int64_t n, int64_t rows_per_split, const double* __restrict__ X,
                                                         const double* __restrict__ gc, double* __restrict__ partial) {
  const int64_t col = ((int64_t)blockIdx.x * LG_T + threadIdx.x) * 2;
  if (col >= n) return;
  const int64_t rb = (int64_t)blockIdx.y * rows_per_split;
  const int64_t re = rb + rows_per_split < m ? rb + rows_per_split : m;
  double a0 = 0.0, a1 = 0.0;
  for (int64_t i = rb; i < re; ++i) {
    const double c = gc[i];
    const double2 xv = ld_stream_nc(X + i * n + col);
    a0 = fma(xv.x, c, a0);
    a1 = fma(xv.y, c, a1);
  }
  *reinterpret_cast<double2*>(partial + (int64_t)blockIdx.y * n + col) = make_double2(a0, a1);
}
__global__ void __launch_bounds__(LG_T) logit_grad_stage2(int64_t n, int64_t nsplit, double lambda, const double* __restrict__ partial,
                                                         const double* __restrict__ w, double* __restrict__ g) {
  const int64_t col = (int64_t)blockIdx.x * LG_T + threadIdx.x;
  if (col >= n) return;
  double a = 0.0;
  for (int64_t k = 0; k < nsplit; ++k) a = a + partial[k * n + col];
  g[col] = a + lambda * w[col];
}

// ---- Hessian: C = X^T diag(d) X (+ lambda I) on DMMA ------------------------------------------
// mma.sync m8n8k4 f64 is the hardware shape (SASS DMMA.8x8x4; the larger PTX shapes m16n8k{4,8,16} assemble to
// sequences of the same instruction on sm_100a — checked with cuobjdump — and tcgen05 has no FP64 kind).
// Persistent CTAs walk a list of (K split, output tile) units, split-major, so that at any moment the whole GPU streams
// the same rows of X (they are then L2 hits for all but the first reader) and the last wave is 1/56 of a CTA's work
// instead of 1/15 (2080 lower-triangular 128 x 128 tiles do not divide by 148 SMs).  Each unit accumulates its K range
// for one tile and stores a partial tile; a second kernel adds the SY_SPLIT partials of every element in split order
// (deterministic), adds lambda on the diagonal and mirrors into the upper triangle.
// Operands move global -> shared with 16-byte cp.async in a ring of SY_STAGES stages of SY_KC samples (no register staging); the Hessian weights d_k
// multiply the A fragments as they are read from shared memory, so both panels are raw rows of X and a diagonal tile
// loads only one.
constexpr int SY_TILE = 128;   // output tile edge
constexpr int SY_LD = SY_TILE + 4;  // padded row stride (conflict-free fragment reads)
constexpr int SY_SPLIT = 4;    // K splits per tile

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int KC, int STAGES>
struct SyrkSmem {
  double A[STAGES][KC][SY_LD];
  double B[STAGES][KC][SY_LD];
  double dk[STAGES][KC];
};

// NWC: warps along the columns of the tile: 2 (8 warps, 32 x 64 per warp) or 4 (16 warps, 32 x 32 per warp);
// KC samples per pipeline stage, STAGES stages
template <int NWC, int KC, int STAGES>
__global__ void __launch_bounds__(128 * NWC, 1)
syrk_dmma_kernel(int64_t m, int64_t n, const double* __restrict__ X, const double* __restrict__ dcoef, double* __restrict__ Cpart, int ntiles) {
  extern __shared__ __align__(16) unsigned char sy_raw[];
  using Smem = SyrkSmem<KC, STAGES>;
  Smem& sm = *reinterpret_cast<Smem*>(sy_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NT = 128 * NWC, WCOLS = SY_TILE / NWC, NB = WCOLS / 8;
  const int wr = warp / NWC, wc = warp % NWC;  // warp tile origin: rows wr * 32, cols wc * WCOLS
  const int64_t kper = ((m + SY_SPLIT - 1) / SY_SPLIT + KC - 1) / KC * KC;  // samples per split (whole stages)
  const int nunits = ntiles * SY_SPLIT;
  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int split = u / ntiles;
    int t = u % ntiles, ti = 0;
    while (t > ti) {  // lower-triangular tile index -> (ti, tj), tj <= ti
      t -= ti + 1;
      ++ti;
    }
    const int tj = t;
    const bool diag = ti == tj;
    const int64_t i0 = (int64_t)ti * SY_TILE, j0 = (int64_t)tj * SY_TILE;
    const int64_t kbeg = (int64_t)split * kper, kend = kbeg + kper < m ? kbeg + kper : m;
    const int npanels = kend > kbeg ? (int)((kend - kbeg + KC - 1) / KC) : 0;
    double acc[4][NB][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    // one stage = KC rows x 128 doubles per operand = 64 KC 16-byte chunks per operand
    auto issue = [&](int p) {
      if (p < npanels) {
        const int st = p % STAGES;
        const int64_t k0 = kbeg + (int64_t)p * KC;
#pragma unroll
        for (int c = 0; c < KC * 64 / NT; ++c) {
          const int chunk = tid + c * NT;  // 0 .. 64 KC - 1
          const int kk = chunk >> 6;         // 64 chunks per row
          const int cc = (chunk & 63) * 2;
          const int64_t k = k0 + kk;
          const bool kok = k < kend;
          const int64_t krow = kok ? k : kbeg;  // (a valid address even when the copy is zero-filled)
          cp_async16(&sm.A[st][kk][cc], X + krow * n + i0 + cc, kok && i0 + cc < n);
          if (!diag) cp_async16(&sm.B[st][kk][cc], X + krow * n + j0 + cc, kok && j0 + cc < n);
        }
        if (tid < KC) sm.dk[st][tid] = (k0 + tid < kend) ? dcoef[k0 + tid] : 0.0;
      }
      cp_async_commit();  // (an empty group keeps the group count in step with the panel count)
    };
    __syncthreads();  // the previous unit's readers are done with the ring
#pragma unroll
    for (int p = 0; p < STAGES - 1; ++p) issue(p);
    for (int p = 0; p < npanels; ++p) {
      cp_async_wait<STAGES - 2>();  // panel p has landed (this thread's copies) ...
      __syncthreads();                 // ... and everybody's; stage (p - 1) % STAGES is free again
      issue(p + STAGES - 1);
      const int st = p % STAGES;
      const double (*As)[SY_LD] = sm.A[st];
      const double (*Bs)[SY_LD] = diag ? sm.A[st] : sm.B[st];
#pragma unroll
      for (int k4 = 0; k4 < KC / 4; ++k4) {
        double af[4], bf[NB];
        const int kr = k4 * 4 + (lane & 3);
        const double dkv = sm.dk[st][kr];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = As[kr][wr * 32 + a * 8 + (lane >> 2)] * dkv;
#pragma unroll
        for (int b = 0; b < NB; ++b) bf[b] = Bs[kr][wc * WCOLS + b * 8 + (lane >> 2)];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < NB; ++b) dmma_884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
      }
    }
    // this unit's partial tile (row-major 128 x 128)
    double* out = Cpart + ((int64_t)split * ntiles + (u % ntiles)) * (SY_TILE * SY_TILE);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int r = wr * 32 + a * 8 + (lane >> 2);
        const int c = wc * WCOLS + b * 8 + (lane & 3) * 2;
        *reinterpret_cast<double2*>(out + r * SY_TILE + c) = make_double2(acc[a][b][0], acc[a][b][1]);
      }
  }
}

// C = sum over the splits (in split order) of the partial tiles, + lambda on the diagonal, mirrored
__global__ void __launch_bounds__(256) syrk_reduce_kernel(int64_t n, const double* __restrict__ Cpart, int ntiles, double lambda, int add_lambda,
                                                          double* __restrict__ C, int64_t ldc) {
  int t = blockIdx.x, ti = 0;
  while (t > ti) {
    t -= ti + 1;
    ++ti;
  }
  const int tj = t;
  const int64_t i0 = (int64_t)ti * SY_TILE, j0 = (int64_t)tj * SY_TILE;
  const double* base = Cpart + (int64_t)blockIdx.x * (SY_TILE * SY_TILE);
  for (int e = threadIdx.x; e < SY_TILE * SY_TILE / 2; e += 256) {
    const int r = e / (SY_TILE / 2), c = (e % (SY_TILE / 2)) * 2;
    double2 v = *reinterpret_cast<const double2*>(base + r * SY_TILE + c);
#pragma unroll
    for (int sp = 1; sp < SY_SPLIT; ++sp) {
      const double2 w = *reinterpret_cast<const double2*>(base + (int64_t)sp * ntiles * (SY_TILE * SY_TILE) + r * SY_TILE + c);
      v.x = v.x + w.x;
      v.y = v.y + w.y;
    }
    const int64_t i = i0 + r;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int64_t jx = j0 + c + q;
      double val = q == 0 ? v.x : v.y;
      if (i < n && jx < n && jx <= i) {
        if (add_lambda && i == jx) val += lambda;
        C[i * ldc + jx] = val;
        C[jx * ldc + i] = val;
      }
    }
  }
}

constexpr int SY_KC = 48, SY_STAGES = 2;  // measured at m = 262144: 16 x 4 stages 28.5, 32 x 3 29.8, 48 x 2 30.6 TFLOP/s (one CTA barrier per stage)
constexpr size_t SY_SMEM = sizeof(SyrkSmem<SY_KC, SY_STAGES>);
static void syrk_set_smem() {
  static bool done = false;
  if (!done) {
    OSB_CUDA(cudaFuncSetAttribute(syrk_dmma_kernel<2, SY_KC, SY_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SY_SMEM));
    done = true;
  }
}
static int syrk_ntiles(int64_t n) {
  const int nt = (int)((n + SY_TILE - 1) / SY_TILE);
  return nt * (nt + 1) / 2;
}
static int64_t syrk_part_doubles(int64_t n) { return (int64_t)SY_SPLIT * syrk_ntiles(n) * SY_TILE * SY_TILE; }
// H = X^T D X (+ lambda I when add_lambda) into hess (row-major, ldc), both triangles
static void syrk_launch(Ctx* ctx, int64_t m, int64_t n, const double* X, const double* dc, double lambda, int add_lambda, double* part,
                        double* hess, int64_t ldc) {
  syrk_set_smem();
  const int ntiles = syrk_ntiles(n);
  const int grid = std::min(ctx->num_sms, ntiles * SY_SPLIT);
  syrk_dmma_kernel<2, SY_KC, SY_STAGES><<<grid, 256, SY_SMEM, ctx->stream>>>(m, n, X, dc, part, ntiles);
  syrk_reduce_kernel<<<ntiles, 256, 0, ctx->stream>>>(n, part, ntiles, lambda, add_lambda, hess, ldc);
  ctx->counters[0] += 2;
}

__global__ void logit_finish_f_kernel(const double* __restrict__ tmp2, double lambda, double* __restrict__ d_f) {
  d_f[0] = tmp2[0] + 0.5 * lambda * tmp2[1];
}

struct LogisticObjective : Objective {
  int64_t m;        // local sample count (sample-sharded across ranks)
  int64_t m_total;
  double lambda;
  DBuf X, ysign, loss, gc, dc, partial, tmp, hpart;  // hpart: SY_SPLIT partial tiles per output tile of the Hessian
  int64_t rows_per_split, nsplit;
  LogisticObjective(Ctx* c, int64_t m_, int64_t n_, double lam) : Objective(c, n_), m_total(m_), lambda(lam) {
    OSB_REQUIRE(n_ % 2 == 0, OSB_ERROR_INPUT_PARAMS, "logistic regression needs an even feature count");
    OSB_REQUIRE(m_ % c->world == 0, OSB_ERROR_INPUT_PARAMS, "sample count must divide by the number of ranks");
    m = m_ / c->world;
    const int64_t row0 = m * c->rank;
    X.alloc(m * n);
    ysign.alloc(m);
    loss.alloc(m);
    gc.alloc(m);
    dc.alloc(m);
    nsplit = m < 512 ? m : 512;
    rows_per_split = (m + nsplit - 1) / nsplit;
    nsplit = (m + rows_per_split - 1) / rows_per_split;
    partial.alloc(nsplit * n);
    tmp.alloc(8);
    logit_gen_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(m, n, row0, X.p, ysign.p);
    c->counters[0]++;
    c->sync();
  }
  bool provides_hessian() const override { return true; }
  void eval(const double* w, double* d_f, double* g, double* hess) override {
    calls++;
    ctx->counters[1]++;
    cudaStream_t st = ctx->stream;
    logit_margin_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(m, n, X.p, ysign.p, w, loss.p, gc.p, dc.p);
    dim3 g1((unsigned)((n / 2 + LG_T - 1) / LG_T), (unsigned)nsplit);
    logit_grad_stage1<<<g1, LG_T, 0, st>>>(m, n, rows_per_split, X.p, gc.p, partial.p);
    const double lam_local = ctx->world > 1 ? 0.0 : lambda;  // sharded: lambda w is added after the all-reduce
    logit_grad_stage2<<<(unsigned)((n + LG_T - 1) / LG_T), LG_T, 0, st>>>(n, nsplit, lam_local, partial.p, w, g);
    ctx->counters[0] += 3;
    const double* lp = loss.p;
    auto fl = [=] __device__(int64_t i, double(&acc)[1]) { acc[0] = acc[0] + lp[i]; };
    launch_mapreduce<1>(ctx, fl, m, RedOps<1>{{RED_SUM}}, tmp.p);
    auto fw = [=] __device__(int64_t i, double(&acc)[1]) { acc[0] = acc[0] + w[i] * w[i]; };
    launch_mapreduce<1>(ctx, fw, n, RedOps<1>{{RED_SUM}}, tmp.p + 1);
    if (ctx->world > 1) {
      ctx_all_reduce_sum(ctx, tmp.p, 1);  // loss sum over shards
      ctx_all_reduce_sum(ctx, g, n);
      const double lam = lambda;
      double* gg = g;
      auto fa = [=] __device__(int64_t i, double(&acc)[1]) { gg[i] = gg[i] + lam * w[i]; };
      launch_mapreduce<1>(ctx, fa, n, RedOps<1>{{RED_SUM}}, ctx->d_dummy);
    }
    logit_finish_f_kernel<<<1, 1, 0, st>>>(tmp.p, lambda, d_f);
    ctx->counters[0]++;
    if (hess) {
      const int64_t ldc = qn_ld(n);
      if (hpart.p == nullptr) hpart.alloc(syrk_part_doubles(n));
      syrk_launch(ctx, m, n, X.p, dc.p, lambda, ctx->world > 1 ? 0 : 1, hpart.p, hess, ldc);
      if (ctx->world > 1) {
        ctx_all_reduce_sum(ctx, hess, n * ldc);
        const double lam = lambda;
        double* hh = hess;
        auto fd = [=] __device__(int64_t i, double(&acc)[1]) { hh[i * ldc + i] = hh[i * ldc + i] + lam; };
        launch_mapreduce<1>(ctx, fd, n, RedOps<1>{{RED_SUM}}, ctx->d_dummy);
      }
    }
  }
};

Objective* make_logistic_generated(Ctx* ctx, int64_t m, int64_t n, double lambda) { return new LogisticObjective(ctx, m, n, lambda); }

// stand-alone timing of the Hessian assembly for bench.py: returns ms per launch
double bench_syrk_dmma(Ctx* ctx, Objective* obj, int reps) {
  auto* o = dynamic_cast<LogisticObjective*>(obj);
  OSB_REQUIRE(o != nullptr, OSB_ERROR_INPUT_PARAMS, "not a logistic objective");
  const int64_t n = o->n, ldc = qn_ld(n);
  DBuf hess(qn_rows_padded(n) * ldc), w(ldc);
  w.zero(ctx->stream);
  logit_margin_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(o->m, n, o->X.p, o->ysign.p, w.p, o->loss.p, o->gc.p, o->dc.p);
  if (o->hpart.p == nullptr) o->hpart.alloc(syrk_part_doubles(n));
  cudaEvent_t e0, e1;
  OSB_CUDA(cudaEventCreate(&e0));
  OSB_CUDA(cudaEventCreate(&e1));
  syrk_launch(ctx, o->m, n, o->X.p, o->dc.p, o->lambda, 1, o->hpart.p, hess.p, ldc);
  OSB_CUDA(cudaEventRecord(e0, ctx->stream));
  for (int r = 0; r < reps; ++r) syrk_launch(ctx, o->m, n, o->X.p, o->dc.p, o->lambda, 1, o->hpart.p, hess.p, ldc);
  OSB_CUDA(cudaEventRecord(e1, ctx->stream));
  OSB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  OSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx->counters[0] += 1;
  return (double)ms / reps;
}

}  // namespace osb
