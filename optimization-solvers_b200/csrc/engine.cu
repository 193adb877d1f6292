// engine.cu — context, buffers and the solver state machines.
//
// Solver::minimize is LineSearchSolver::minimize (src/ls_solver.rs:66-111) for every solver struct
// on the hot path.  Two control engines drive the same kernels:
//   * host-driven   — every solver x line search x objective; the scalar automaton
//                     (ls_automaton.cuh) runs on the host, one small D2H fetch per decision;
//   * device-resident — dense quasi-Newton + block-functor objective: convergence test, direction,
//                     the whole line search and the s/y bookkeeping run in one single-CTA kernel,
//                     the H passes are predicated on device flags, and the host only polls a
//                     `done` flag a few iterations behind (no host round trip on the critical path).
#include "engine.cuh"

#include <atomic>
#include <mutex>
#include <vector>

#include <algorithm>
#include <cstring>

namespace osb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
const std::string& get_last_error() { return g_last_error; }

static osb_log_fn g_log_fn = nullptr;
static void* g_log_user = nullptr;
void set_log_callback(osb_log_fn fn, void* user) {
  g_log_fn = fn;
  g_log_user = user;
}
void log_event(int level, const char* target, const std::string& message) {
  if (g_log_fn) g_log_fn(g_log_user, level, target, message.c_str());
}

// ---- Ctx / DBuf ---------------------------------------------------------------------------
Ctx::Ctx(int dev) : device(dev) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error(OSB_ERR_CUDA, "no CUDA device available: this library has no CPU fallback");
  OSB_REQUIRE(dev >= 0 && dev < count, OSB_ERROR_INPUT_PARAMS, "bad device index");
  OSB_CUDA(cudaSetDevice(dev));
  cudaDeviceProp prop;
  OSB_CUDA(cudaGetDeviceProperties(&prop, dev));
  num_sms = prop.multiProcessorCount;
  OSB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  red_max_grid = num_sms * 4;
  OSB_CUDA(cudaMalloc(&red_partials, sizeof(double) * 8 * red_max_grid));
  OSB_CUDA(cudaMalloc(&red_ticket, sizeof(unsigned int)));
  OSB_CUDA(cudaMemsetAsync(red_ticket, 0, sizeof(unsigned int), stream));
  OSB_CUDA(cudaMalloc(&gemv_ticket, sizeof(unsigned int)));
  OSB_CUDA(cudaMemsetAsync(gemv_ticket, 0, sizeof(unsigned int), stream));
  OSB_CUDA(cudaMalloc(&d_dummy, sizeof(double) * 8));
  OSB_CUDA(cudaHostAlloc(&h_pinned, sizeof(double) * 4096, cudaHostAllocDefault));
  OSB_CUDA(cudaStreamSynchronize(stream));
}
Ctx::~Ctx() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  cudaFree(red_partials);
  cudaFree(red_ticket);
  cudaFree(gemv_ticket);
  cudaFree(d_dummy);
  cudaFree(shard_scratch);
  pool_trim(device);
  cudaFreeHost(h_pinned);
  if (stream) cudaStreamDestroy(stream);
}

// ---- allocation pool -----------------------------------------------------------------------
// Every device / pinned buffer of a solver goes back to a per-device free list when the solver is destroyed and is
// handed out again for an equal-sized request.  Measured on B200 with another solver alive on the device: constructing
// a solver straight from the driver costs 2-9 ms and destroying it 2-6 ms (15 cudaMalloc + 1 cudaHostAlloc, and their
// frees); 1-2 GiB matrices cost 2-7 ms to allocate and up to 90 ms to free.  From the pool: 0.1 ms and 0.05 ms.
// Contents are neither preserved nor cleared.  osb_ctx_trim_memory / context destruction return everything.
namespace {
struct PoolEntry {
  int device;
  int kind;  // 0 = device memory, 1 = pinned host memory
  size_t bytes;
  void* p;
};
std::mutex g_pool_mutex;
std::vector<PoolEntry> g_pool;
constexpr size_t POOL_MAX_ENTRIES = 256;
constexpr size_t POOL_MAX_BYTES = (size_t)8 << 30;
}  // namespace

constexpr int64_t SNAP_CHUNK = 32;  // iterations per launch of the fused kernel when it publishes callback snapshots (ring: 2 halves)

// flag values of the in-kernel callback snapshots: unique over the process, so a recycled pinned ring never holds a
// value a later launch waits for
static std::atomic<unsigned long long> g_snap_seq{1ULL << 20};

void* pool_get(int kind, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    for (size_t i = 0; i < g_pool.size(); ++i)
      if (g_pool[i].device == dev && g_pool[i].kind == kind && g_pool[i].bytes == bytes) {
        void* p = g_pool[i].p;
        g_pool.erase(g_pool.begin() + i);
        return p;
      }
  }
  void* p = nullptr;
  cudaError_t e = kind == 0 ? cudaMalloc(&p, bytes) : cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    pool_trim(dev);  // give the pooled memory back and try once more
    e = kind == 0 ? cudaMalloc(&p, bytes) : cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
  }
  if (e != cudaSuccess)
    throw Error(OSB_ERR_ALLOC, std::string(kind == 0 ? "cudaMalloc of " : "cudaHostAlloc of ") + std::to_string(bytes) +
                                   " bytes failed: " + cudaGetErrorString(e));
  return p;
}
void pool_put(int kind, size_t bytes, void* p) {
  if (!p) return;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    size_t total = 0;
    for (const PoolEntry& e : g_pool) total += e.bytes;
    if (g_pool.size() < POOL_MAX_ENTRIES && total + bytes <= POOL_MAX_BYTES) {
      g_pool.push_back(PoolEntry{dev, kind, bytes, p});
      return;
    }
  }
  if (kind == 0) cudaFree(p);
  else cudaFreeHost(p);
}
void pool_trim(int device) {
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  for (size_t i = 0; i < g_pool.size();) {
    if (g_pool[i].device == device) {
      if (g_pool[i].kind == 0) cudaFree(g_pool[i].p);
      else cudaFreeHost(g_pool[i].p);
      g_pool.erase(g_pool.begin() + i);
    } else {
      ++i;
    }
  }
}

void DBuf::alloc(int64_t n_) {
  release();
  n = n_;
  if (n > 0) p = (double*)pool_get(0, sizeof(double) * (size_t)n);
}
void DBuf::alloc_pooled(int64_t n_) { alloc(n_); }
void DBuf::release() {
  if (p) pool_put(0, sizeof(double) * (size_t)n, p);
  p = nullptr;
  n = 0;
}
void DBuf::zero(cudaStream_t s) {
  if (p) OSB_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * (size_t)n, s));
}
void DBuf::upload(const double* h, int64_t cnt, cudaStream_t s) {
  OSB_CUDA(cudaMemcpyAsync(p, h, sizeof(double) * (size_t)cnt, cudaMemcpyHostToDevice, s));
}
void DBuf::download(double* h, int64_t cnt, cudaStream_t s) const {
  OSB_CUDA(cudaMemcpyAsync(h, p, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, s));
}

// ---- small state kernels ------------------------------------------------------------------
__global__ void set_identity_kernel(double* H, int64_t ld, int64_t nrows, int64_t row0) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += (int64_t)gridDim.x * blockDim.x)
    H[i * ld + row0 + i] = 1.0;
}
__global__ void accept_trial_kernel(DevState* st) { st->f = st->ft; }
// run-ahead callbacks, one launch per phase: x, g and the control block of the iteration that has just been enqueued go
// into a device ring slot (one small kernel on the compute stream, ~2 us, instead of three device-to-host copies that
// held the stream for ~21 us per iteration); the host copies the slot out on a side stream when it delivers the callback
__global__ void __launch_bounds__(256) snap_copy_kernel(const double* __restrict__ x, const double* __restrict__ g, const DevState* __restrict__ st,
                                                        int64_t n, int64_t ld, double* __restrict__ slot_xg, DevState* __restrict__ slot_st) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    slot_xg[i] = x[i];
    slot_xg[ld + i] = g[i];
  }
  if (blockIdx.x == 0 && threadIdx.x < sizeof(DevState) / sizeof(int))
    reinterpret_cast<int*>(slot_st)[threadIdx.x] = reinterpret_cast<const int*>(st)[threadIdx.x];
}

__global__ void reset_run_flags_kernel(DevState* st) {  // ls_solver.rs:74: k = 0; a fresh run
  st->k = 0;
  st->done = 0;
  st->status = OSB_MAX_ITER_REACHED;
  st->reason = OSB_REASON_NONE;
  st->skip = 0;
  st->ls_evals = 0;
}

// ---- Solver -------------------------------------------------------------------------------
static bool kind_is_qn(int k) { return k >= OSB_BFGS && k <= OSB_SR1B; }
static bool kind_is_bounded(int k) {
  return k == OSB_PGD || k == OSB_SPG || k == OSB_BFGSB || k == OSB_DFPB || k == OSB_BROYDENB || k == OSB_SR1B ||
         k == OSB_PROJ_NEWTON || k == OSB_SPN;
}
static bool kind_needs_hessian(int k) { return k == OSB_NEWTON || k == OSB_PROJ_NEWTON || k == OSB_SPN; }
static bool kind_needs_y(int k) { return kind_is_qn(k) || k == OSB_SPG || k == OSB_SPN || k == OSB_PROJ_NEWTON; }

Solver::Solver(Ctx* c, int kind_, int64_t n_, double tol_, const double* x0, const double* lb_h, const double* ub_h,
               Objective* obj0)
    : ctx(c), kind(kind_), n(n_), tol(tol_) {
  ctx->use();
  OSB_REQUIRE(kind >= OSB_GD && kind <= OSB_PNORM, OSB_ERROR_INPUT_PARAMS, "unknown solver kind");
  OSB_REQUIRE(n >= 1 && x0 != nullptr, OSB_ERROR_INPUT_PARAMS, "n >= 1 and x0 required");
  bounded = kind_is_bounded(kind);
  is_qn = kind_is_qn(kind);
  OSB_REQUIRE(!ctx->vec_sharded || kind == OSB_GD || kind == OSB_PGD || kind == OSB_SPG, OSB_ERR_UNSUPPORTED,
              "index-range sharding supports GradientDescent, ProjectedGradientDescent and SpectralProjectedGradient");
  OSB_REQUIRE(!bounded || (lb_h && ub_h), OSB_ERROR_INPUT_PARAMS, "bounded solver needs lower and upper bounds");
  ld = qn_ld(n);
  cudaStream_t st = ctx->stream;
  for (DBuf* b : {&x, &g, &d, &xt, &gt, &s, &y, &w}) {
    b->alloc(ld);
    b->zero(st);
  }
  x.upload(x0, n, st);
  if (bounded) {
    lb.alloc(ld);
    ub.alloc(ld);
    lb.zero(st);
    ub.zero(st);
    lb.upload(lb_h, n, st);
    ub.upload(ub_h, n, st);
    vec_project_inplace(ctx, n, x.p, lb.p, ub.p);  // constructors project x0: bfgs_b.rs:50, spg.rs:35
  }
  d_state = (DevState*)pool_get(0, sizeof(DevState));
  OSB_CUDA(cudaMemsetAsync(d_state, 0, sizeof(DevState), st));
  h_state = (DevState*)pool_get(1, sizeof(DevState));
  std::memset(h_state, 0, sizeof(DevState));
  OSB_CUDA(cudaEventCreate(&ev0));
  OSB_CUDA(cudaEventCreate(&ev1));
  if (is_qn) {
    qn_kind = (kind == OSB_BFGS || kind == OSB_BFGSB) ? QN_BFGS
              : (kind == OSB_DFP || kind == OSB_DFPB) ? QN_DFP
              : (kind == OSB_SR1B)                    ? QN_SR1
                                                      : QN_BROYDEN;
    if (ctx->world > 1) {
      OSB_REQUIRE(n % (ctx->world * 8) == 0, OSB_ERROR_INPUT_PARAMS, "row-sharded H needs n divisible by 8 * world");
      nrows = n / ctx->world;
      row0 = nrows * ctx->rank;
    } else {
      nrows = n;
      row0 = 0;
    }
    // bfgs.rs:30-33: H_0 = I.  At large n the identity is kept virtual until something needs the n x n buffer: with packed
    // symmetric storage nothing ever does (1 GiB instead of 3 at n = 16384, and construction costs no 2 GiB memset)
    if (n >= 2048) H_virtual_identity = true;
    else ensure_full();
    for (DBuf* b : {&u, &h, &pvec, &vvec, &wv, &ps, &ph}) {
      b->alloc(ld);
      b->zero(st);
    }
    if (qn_kind == QN_BROYDEN) scratch.alloc(((nrows + 63) / 64) * ld);
  }
  if (kind == OSB_PNORM) {  // pnorm_descent.rs:15-30: a constant n x n matrix, row-block sharded like H
    if (ctx->world > 1) {
      OSB_REQUIRE(n % (ctx->world * 8) == 0, OSB_ERROR_INPUT_PARAMS, "row-sharded inverse_p needs n divisible by 8 * world");
      nrows = n / ctx->world;
      row0 = nrows * ctx->rank;
    } else {
      nrows = n;
      row0 = 0;
    }
    H.alloc(qn_rows_padded(nrows) * ld);
    H.zero(st);
    set_identity_kernel<<<ctx->red_grid(nrows), RED_THREADS, 0, st>>>(H.p, ld, nrows, row0);
    ctx->counters[0]++;
    u.alloc(ld);
    u.zero(st);
  }
  if (kind_needs_hessian(kind)) {
    hess.alloc(qn_rows_padded(n) * ld);
    hess.zero(st);
    chol.alloc(qn_rows_padded(n) * ld);
    chol.zero(st);
  }
  if (kind == OSB_SPG || kind == OSB_SPN) {
    // spg.rs:40-46: lambda0 = clamp(1 / ||P(x0 - g0) - x0||_inf, lambda_min, lambda_max) — one oracle call
    OSB_REQUIRE(obj0 != nullptr, OSB_ERROR_INPUT_PARAMS, "SPG/SPN constructors need the oracle");
    obj0->eval(x.p, &d_state->f, g.p, nullptr);
    vec_projected_direction(ctx, n, x.p, g.p, 1.0, false, lb.p, ub.p, g.p, d.p, &d_state->gd0);
    fetch_state();
    double dinf = rmax(0.0, h_state->dinf);
    lambda = rmax(rmin(1. / dinf, lambda_max), lambda_min);
    have_eval = true;
  }
  ctx->sync();
}

Solver::~Solver() {
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  pool_put(0, sizeof(DevState), d_state);
  pool_put(1, sizeof(DevState), h_state);
  pool_put(1, 2 * sizeof(DevState), cb_snap);
  pool_put(1, 2 * sizeof(double) * (size_t)ld, cb_xsnap);
  pool_put(1, 2 * sizeof(double) * (size_t)ld, cb_gsnap);
  pool_put(0, 2 * SNAP_CHUNK * 2 * sizeof(double) * (size_t)ld, snap_x);
  pool_put(0, 2 * SNAP_CHUNK * sizeof(DevState), snap_st);
  pool_put(1, 2 * SNAP_CHUNK * sizeof(unsigned long long), snap_flag);
  if (snap_stream) cudaStreamDestroy(snap_stream);
  pool_put(0, sizeof(LSParams), d_ls_buf);
  pool_put(1, 2 * sizeof(DevState), poll_snap);
  if (d_iter_prof) cudaFree(d_iter_prof);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
}

void Solver::fetch_state() {
  OSB_CUDA(cudaMemcpyAsync(h_state, d_state, sizeof(DevState), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
}
void Solver::push_state() {
  OSB_CUDA(cudaMemcpyAsync(d_state, h_state, sizeof(DevState), cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
}

void Solver::compute_conv_scalar(Objective*) {
  double* out = &d_state->conv;
  if (is_qn) vec_conv_gnorm2(ctx, n, g.p, out);                        // bfgs.rs:74
  else if (kind == OSB_GD || kind == OSB_PNORM) vec_conv_gmax(ctx, n, g.p, out);  // gradient_descent.rs:46-53, pnorm_descent.rs:52-58
  else if (kind == OSB_NEWTON) return;                                 // newton/mod.rs:64-69 ignores the eval
  else vec_conv_pginf(ctx, n, x.p, g.p, lb.p, ub.p, out);              // projected_gradient_descent.rs:76-83
}

// dense SPD solve for the Newton family (newton.cu)
int newton_solve(Ctx* ctx, int64_t n, int64_t ld, const double* hess, double* chol, const double* g, double* w_out);
int newton_solve_lu(Ctx* ctx, int64_t n, int64_t ld, const double* hess, double* lu, int* perm, double* tmp, const double* rhs, double* w_out);

int Solver::compute_direction(Objective*, LineSearch* ls) {
  double* out3 = &d_state->gd0;  // {gd0, tmaxc, dinf}
  switch (kind) {
    case OSB_GD:
      vec_neg(ctx, n, g.p, d.p, g.p, out3);  // gradient_descent.rs:29
      break;
    case OSB_PNORM:  // pnorm_descent.rs:35: (-inverse_p) g == -(inverse_p g), one read of the matrix
      if (n <= QN_SMALL_N && ctx->world == 1) qn_small_gemv(ctx, n, ld, H.p, g.p, u.p);
      else qn_launch_gemv(ctx, H.p, ld, nrows, row0, nullptr, g.p, u.p, nullptr, nullptr, 0);
      if (ctx->world > 1) ctx->all_gather_inplace(u.p, nrows);
      vec_neg(ctx, n, u.p, d.p, g.p, out3);
      break;
    case OSB_PGD:
      vec_projected_direction(ctx, n, x.p, g.p, 1.0, false, lb.p, ub.p, g.p, d.p, out3);  // projected_gradient_descent.rs:56-59
      break;
    case OSB_SPG:
      vec_projected_direction(ctx, n, x.p, g.p, lambda, true, lb.p, ub.p, g.p, d.p, out3);  // spg.rs:81-84
      break;
    case OSB_BFGS:
    case OSB_DFP:
    case OSB_BROYDEN:
      vec_neg(ctx, n, u.p, d.p, g.p, out3);  // bfgs.rs:47: (-H) g == -(H g)
      break;
    case OSB_BFGSB:
    case OSB_DFPB:
    case OSB_BROYDENB:
    case OSB_SR1B:
      vec_projected_direction(ctx, n, x.p, u.p, 1.0, false, lb.p, ub.p, g.p, d.p, out3);  // bfgs_b.rs:72-75
      break;
    case OSB_NEWTON: {
      // newton/mod.rs:31-47.  The reference inverts with LU (try_inverse); for the SPD Hessians of the configs a
      // Cholesky solve gives the same direction to O(cond * eps).  A Hessian that is not SPD goes through LU with
      // partial pivoting like the reference's; a singular one (exact zero pivot) takes d = -g and leaves the
      // decrement untouched (newton/mod.rs:43-46).  cholesky().unwrap() panics belong to ProjectedNewton / SPN only.
      newton_singular = false;
      bool use_lu = false;
      int rc = newton_solve(ctx, n, ld, hess.p, chol.p, g.p, w.p);
      if (rc == OSB_PANIC_NOT_SPD) {
        use_lu = true;
        if (!lu_perm.p) lu_perm.alloc(ld);  // n ints + n doubles of scratch
        int* perm = reinterpret_cast<int*>(lu_perm.p);
        rc = newton_solve_lu(ctx, n, ld, hess.p, chol.p, perm, xt.p, g.p, w.p);
        if (rc == OSB_PANIC_NOT_SPD) {  // singular
          log_event(2, "newton", "Hessian is singular. Using gradient descent direction.");  // newton/mod.rs:44
          newton_singular = true;
          vec_neg(ctx, n, g.p, d.p, g.p, out3);
          break;
        }
      }
      if (rc != OSB_OK) return rc;
      vec_neg(ctx, n, w.p, d.p, g.p, out3);  // d = -(H^-1 g)
      // decrement^2 = (H^-1 d) . d
      if (use_lu) rc = newton_solve_lu(ctx, n, ld, nullptr, chol.p, reinterpret_cast<int*>(lu_perm.p), xt.p, d.p, w.p);
      else rc = newton_solve(ctx, n, ld, nullptr, chol.p, d.p, w.p);
      if (rc != OSB_OK) return rc;
      vec_dot(ctx, n, w.p, d.p, &d_state->dinf);
      break;
    }
    case OSB_PROJ_NEWTON:
    case OSB_SPN: {
      int rc = newton_solve(ctx, n, ld, hess.p, chol.p, g.p, w.p);  // projected_newton.rs:75, spn.rs:86
      if (rc != OSB_OK) return rc;
      vec_projected_direction(ctx, n, x.p, w.p, lambda, kind == OSB_SPN, lb.p, ub.p, g.p, d.p, out3);
      break;
    }
  }
  if (ls->p.kind == LS_MORETHUENTE_B)  // morethuente_b.rs:185-197 with the line search's own bounds
    vec_tmax_candidate(ctx, n, x.p, d.p, ls->lb.p, ls->ub.p, &d_state->tmaxc);
  return OSB_OK;
}

// after the step: s, y are formed, x/g hold the NEW iterate; all launches are predicated on the
// device flags (done / skip), so no host decision is needed here.
void Solver::prof_mark() {
  if (!profile_kernels || prof_events.size() >= 4 * 4096) return;
  cudaEvent_t e;
  OSB_CUDA(cudaEventCreate(&e));
  OSB_CUDA(cudaEventRecord(e, ctx->stream));
  prof_events.push_back(e);
}
void Solver::prof_collect() {
  prof_ms[0] = prof_ms[1] = prof_ms[2] = 0.0;
  const size_t iters = prof_events.size() / 4;
  for (size_t i = 0; i < iters; ++i) {
    float a = 0.f, b = 0.f;
    OSB_CUDA(cudaEventElapsedTime(&a, prof_events[4 * i], prof_events[4 * i + 1]));
    OSB_CUDA(cudaEventElapsedTime(&b, prof_events[4 * i + 2], prof_events[4 * i + 3]));
    prof_ms[0] += a;
    prof_ms[1] += b;
  }
  if (iters) {
    prof_ms[0] /= iters;
    prof_ms[1] /= iters;
  }
  prof_ms[2] = (double)iters;
  for (cudaEvent_t e : prof_events) cudaEventDestroy(e);
  prof_events.clear();
}

// lazy schedule: the stored matrix lags the true H by one rank-2 update; apply it (predicated on the device flag)
// packed symmetric copy -> full matrix (getters, engine / schedule switches)
void Solver::ensure_full() {
  if (!H.p) {
    H.alloc_pooled(qn_rows_padded(nrows) * ld);
    H.zero(ctx->stream);
    if (H_virtual_identity || !sym_current) {
      set_identity_kernel<<<ctx->red_grid(nrows), RED_THREADS, 0, ctx->stream>>>(H.p, ld, nrows, row0);
      ctx->counters[0]++;
    }
  }
  H_virtual_identity = false;
}
void ctx_all_reduce_sum(Ctx* ctx, double* buf, int64_t count);  // dist.cu
void Solver::sym_settle_pingpong() {
  if (!sym_pingpong_dirty) return;
  fetch_state();
  if (h_state->pp) {
    std::swap(Hsym.p, Hsym2.p);
    h_state->pp = 0;
    OSB_CUDA(cudaMemcpyAsync(&d_state->pp, &h_state->pp, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ctx->sync();
  }
  sym_pingpong_dirty = false;
}
void Solver::sym_to_full() {
  if (!sym_current) return;
  ensure_packed();  // (materialises a still unwritten identity)
  sym_settle_pingpong();
  ensure_full();
  if (ctx->world > 1) {
    // tile pairs are spread over the ranks: every rank unpacks its tiles into a zeroed n x n scratch (both triangles),
    // the scratches are summed over the ranks, and the local row block is copied out.  Only getters and engine /
    // schedule switches come here.
    DBuf full;
    full.alloc(qn_rows_padded(n) * ld);
    full.zero(ctx->stream);
    qn_sym_unpack_sharded(ctx, Hsym.p, ld, n, full.p);
    ctx_all_reduce_sum(ctx, full.p, n * ld);
    OSB_CUDA(cudaMemcpyAsync(H.p, full.p + row0 * ld, sizeof(double) * (size_t)(nrows * ld), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->sync();
    Hsym.release();  // the local tile set is only meaningful together with H_virtual_identity / sym_current
    colpart.release();
  } else {
    qn_sym_unpack(ctx, Hsym.p, ld, n, H.p);
  }
  sym_current = false;
}

HeadEpi Solver::head_epi() const {
  if (!defer_epi) return HeadEpi{-1, nullptr, nullptr, nullptr, nullptr, 1, 0, nullptr, nullptr, nullptr};
  if (sym_sharded) {  // per-rank slots of the exchange region: [parity][rank][h | w][XSLOT_LD]
    const double* b0 = ctx->xchg + XSLOT_OFF;
    const double* b1 = b0 + (int64_t)ctx->world * 2 * XSLOT_LD;
    return HeadEpi{qn_kind, b0, b0 + XSLOT_LD, b1, b1 + XSLOT_LD, ctx->world, 2 * XSLOT_LD, u.p, ps.p, ph.p};
  }
  const double* h2 = nullptr;
  const double* w2 = nullptr;
  const double* h1 = h.p;
  const double* w1 = wv.p;
  if (epi_p2p) {  // qn_lazy_kernel's exchange buffers: parity * 2 + {h, w}
    h1 = ctx->xchg;
    w1 = ctx->xchg + XCHG_LD;
    h2 = ctx->xchg + 2 * XCHG_LD;
    w2 = ctx->xchg + 3 * XCHG_LD;
  }
  return HeadEpi{qn_kind, h1, w1, h2, w2, 1, 0, u.p, ps.p, ph.p};
}
void Solver::finish_epilogue() {
  if (iter_path) {  // the same kernel, epilogue only: the bits do not depend on who ran the epilogue
    QNIterArgs a = iter_args;
    a.iters = 0;
    a.epi_only = 1;
    qn_launch_iter(ctx, iter_fn_kind, iter_fn_a, iter_fn_b, bounded, iter_ls_kind, a);
    return;
  }
  if (!defer_epi) return;
  qn_launch_epilogue_cluster(ctx, head_epi(), n, d_state, s.p, y.p, g.p);
}

void Solver::flush_pending() {
  sym_to_full();
  if (!lazy_used) return;
  ensure_full();
  qn_launch_flush(ctx, qn_kind, H.p, ld, nrows, row0, d_state, ps.p, ph.p);
  lazy_used = false;
}

// A pending update means "the stored matrix lags by one rank-2 term": packing the lagging matrix keeps that meaning.
void Solver::ensure_packed(bool allow_unwritten_identity) {
  if (sym_current) {
    if (sym_ident_unwritten && !allow_unwritten_identity) {  // somebody is about to READ the packed memory
      qn_sym_set_identity(ctx, n, Hsym.p);
      sym_ident_unwritten = false;
    }
    return;
  }
  if (Hsym.p == nullptr) Hsym.alloc_pooled(sym_sharded ? qn_sym_doubles_sharded(n, ctx->world, ctx->rank) : qn_sym_doubles(n));
  if (H_virtual_identity) {
    if (sym_sharded) qn_sym_set_identity_sharded(ctx, n, Hsym.p);
    else if (allow_unwritten_identity) sym_ident_unwritten = true;  // (no 1 GiB memset: the first pass writes the triangle)
    else qn_sym_set_identity(ctx, n, Hsym.p);
    H_virtual_identity = false;
  } else if (sym_sharded) {
    // The matrix exists as row blocks (set_inv_hessian, or an earlier run on the row-block layout); a rank's tile pairs
    // are spread over every row block, so row blocks -> tile pairs is an all-to-all.  It happens once per solve, not per
    // iteration: gather the blocks into an n x n scratch on every rank (NCCL all-gather, 2 GiB at n = 16384), pack the
    // local pairs out of it, drop the scratch and the row block.
    OSB_REQUIRE(n % ctx->world == 0, OSB_ERR_UNSUPPORTED, "row blocks of unequal height");  // (guarded by the caller)
    DBuf full;
    full.alloc(qn_rows_padded(n) * ld);
    OSB_CUDA(cudaMemcpyAsync(full.p + row0 * ld, H.p, sizeof(double) * (size_t)(nrows * ld), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->all_gather_inplace(full.p, nrows * ld);
    qn_sym_pack_sharded(ctx, full.p, ld, n, Hsym.p);
    ctx->sync();
    sym_ident_unwritten = false;
  } else {
    qn_sym_pack(ctx, H.p, ld, n, Hsym.p);
    sym_ident_unwritten = false;
  }
  sym_current = true;
}

void Solver::qn_after_step() {
  const DevState* st = d_state;
  if (n <= QN_SMALL_N && ctx->world == 1) {  // reference operation order, bit-for-bit (qn_small.cu)
    qn_small_step(ctx, qn_kind, n, ld, H.p, d_state, s.p, y.p, g.p, u.p);
    u_valid = true;
    return;
  }
  if (qn_schedule == 1 && qn_storage == 1 && h_symmetric && (ctx->world == 1 || sym_sharded) && (qn_kind == QN_BFGS || qn_kind == QN_DFP)) {
    // packed symmetric storage: the pass moves n^2 * 8 B (read + write of the lower triangle).  A pending
    // update means "the stored matrix lags by one rank-2 term": packing the lagging matrix keeps that meaning.
    ensure_packed(/*allow_unwritten_identity=*/ctx->world == 1 && (qn_variant & 15) == 0);
    QNLazyArgs a{nullptr, ld, nrows, row0, n, d_state, ps.p, ph.p, y.p, g.p, s.p, h.p, wv.p, u.p, ps.p, ph.p,
                 ctx->gemv_ticket, qn_kind, nullptr, ctx->d_seq, 1, 0};
    a.defer_epi = defer_epi ? 1 : 0;
    if (sym_sharded) {  // rows this rank does not own contribute nothing to its slot
      OSB_CUDA(cudaMemsetAsync(h.p, 0, sizeof(double) * (size_t)ld, ctx->stream));
      OSB_CUDA(cudaMemsetAsync(wv.p, 0, sizeof(double) * (size_t)ld, ctx->stream));
      ctx->counters[4]++;  // one fused exchange
      ctx->counters[5]++;  // passes over the sharded packed triangle
    }
    const int pgrid = qn_sym_grid(ctx, n, qn_variant);
    if (colpart.p == nullptr || colpart_grid < pgrid) {
      colpart.alloc_pooled((int64_t)pgrid * 2 * ld);
      colpart_grid = pgrid;
    }
    const bool pingpong = (qn_variant & 2) != 0;
    if (pingpong && Hsym2.n != Hsym.n) Hsym2.alloc_pooled(Hsym.n);
    prof_mark();  // slot 0: the streaming pass, slot 1: the column fold + epilogue
    qn_launch_lazy_sym(ctx, a, Hsym.p, pingpong ? Hsym2.p : Hsym.p, colpart.p, n, ld, 0, qn_variant, sym_ident_unwritten);
    sym_ident_unwritten = false;  // (the pass has written every stored element)
    prof_mark();
    prof_mark();
    qn_launch_lazy_sym(ctx, a, Hsym.p, pingpong ? Hsym2.p : Hsym.p, colpart.p, n, ld, 1, qn_variant);
    prof_mark();
    sym_pingpong_dirty = pingpong;  // which buffer is current is DevState.pp until sym_settle_pingpong()
    lazy_used = true;
    u_valid = true;
    return;
  }
  ensure_full();
  if (qn_schedule == 1 && (qn_kind == QN_BFGS || qn_kind == QN_DFP)) {
    sym_to_full();  // (the pending update, if any, now refers to the full matrix)
    // ONE read-modify-write per iteration (2 n^2 8 B): pending update + h = H y + w = H g, epilogue forms u
    const bool p2p = ctx->world > 1 && ctx->p2p_ready && ld <= XCHG_LD && use_p2p;
    QNLazyArgs a{H.p, ld, nrows, row0, n, d_state, ps.p, ph.p, y.p, g.p, s.p, h.p, wv.p, u.p, ps.p, ph.p,
                 (ctx->world == 1 || p2p) ? ctx->gemv_ticket : nullptr, qn_kind,
                 p2p ? ctx->d_peers : nullptr, ctx->d_seq, ctx->world, ctx->rank};
    a.defer_epi = (defer_epi && qn_variant == 0 && (ctx->world == 1 || p2p)) ? 1 : 0;
    prof_mark();
    prof_mark();
    prof_mark();
    qn_launch_lazy(ctx, a, qn_variant);
    prof_mark();
    if (p2p) ctx->counters[4]++;  // one fused exchange
    if (ctx->world > 1 && !p2p) {
      ctx->all_gather_inplace(h.p, nrows);
      ctx->all_gather_inplace(wv.p, nrows);
      // same arithmetic (cluster order) as the head's deferred epilogue: sharded runs stay bit-identical to one GPU
      if (defer_epi) qn_launch_epilogue_cluster(ctx, head_epi(), n, d_state, s.p, y.p, g.p, true);
      else qn_launch_lazy_epilogue(ctx, a);
    }
    lazy_used = true;
    u_valid = true;
    return;
  }
  // pass 1: h = H y   (skip: u = H g_new, H unchanged — bfgs.rs:106-112)
  const bool fuse_coef = ctx->world == 1;  // sharded: y.h needs the all-gathered h
  QNCoefArgs ca{ctx->gemv_ticket, qn_kind, n, s.p, y.p, pvec.p};
  prof_mark();
  qn_launch_gemv(ctx, H.p, ld, nrows, row0, st, y.p, h.p, g.p, u.p, qn_variant, fuse_coef ? &ca : nullptr);
  prof_mark();
  if (ctx->world > 1) {
    ctx->all_gather_inplace(h.p, nrows);
    ctx->all_gather_inplace(u.p, nrows);
  }
  if (qn_kind == QN_BROYDEN) {  // v = H^T s (broyden.rs:115-117): column sums over the local rows, summed over the row blocks
    qn_launch_gemvT(ctx, H.p, ld, nrows, row0, st, s.p, vvec.p, scratch.p);
    if (ctx->world > 1) ctx_all_reduce_sum(ctx, vvec.p, n);
  }
  if (!fuse_coef) qn_launch_coef(ctx, qn_kind, n, d_state, s.p, y.p, h.p, pvec.p);
  const double* p = (qn_kind == QN_BFGS || qn_kind == QN_DFP) ? s.p : pvec.p;
  // pass 2: fused rank-2 read-modify-write + u = H' g_new
  prof_mark();
  qn_launch_update(ctx, qn_kind, H.p, ld, nrows, row0, st, p, h.p, vvec.p, g.p, u.p, qn_variant);
  prof_mark();
  if (ctx->world > 1) ctx->all_gather_inplace(u.p, nrows);
  u_valid = true;
}

bool Solver::device_engine_supported(const LineSearch* ls, const Objective* obj) const {
  if (!is_qn) return false;
  if (obj->functor_kind() == FN_NONE) return false;
  (void)ls;
  return true;
}

// Options in force for this minimize() call.  Auto (-1) = the fast path wherever it applies: the lazy schedule on the
// packed lower triangle for BFGS / DFP (their H stays exactly symmetric), the eager schedule for Broyden / SR1.
void Solver::resolve_options() {
  if (!is_qn) return;
  const bool sym_kind = (qn_kind == QN_BFGS || qn_kind == QN_DFP) && h_symmetric && n > QN_SMALL_N;
  qn_schedule = opt_schedule < 0 ? (sym_kind ? 1 : 0) : opt_schedule;
  qn_storage = opt_storage < 0 ? ((qn_schedule == 1 && sym_kind) ? 1 : 0) : opt_storage;
}

// u = H g from scratch: only when no H pass has produced it (first iteration, after set_x / set_inv_hessian)
void Solver::recompute_u() {
  flush_pending();  // needs the exact H
  if (H_virtual_identity) {  // H = I: u = g
    OSB_CUDA(cudaMemcpyAsync(u.p, g.p, sizeof(double) * (size_t)ld, cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    ensure_full();
    if (n <= QN_SMALL_N && ctx->world == 1) qn_small_gemv(ctx, n, ld, H.p, g.p, u.p);
    else qn_launch_gemv(ctx, H.p, ld, nrows, row0, nullptr, g.p, u.p, nullptr, nullptr, qn_variant);
    if (ctx->world > 1) ctx->all_gather_inplace(u.p, nrows);
  }
  u_valid = true;
}

int Solver::minimize(LineSearch* ls, Objective* obj, int64_t max_iter, int64_t max_ls, osb_callback_fn cb, void* user) {
  ctx->use();
  resolve_options();
  OSB_REQUIRE(obj->n == n, OSB_ERROR_INPUT_PARAMS, "objective dimension does not match the solver");
  OSB_REQUIRE(!kind_needs_hessian(kind) || obj->provides_hessian(), OSB_PANIC_NO_HESSIAN, "Hessian not available in the oracle");
  if ((ls->p.kind == LS_BACKTRACKING_B || ls->p.kind == LS_MORETHUENTE_B))
    OSB_REQUIRE(ls->n == n, OSB_ERROR_INPUT_PARAMS, "line-search bounds dimension does not match the solver");
  bool dev_ok = device_engine_supported(ls, obj);
  OSB_REQUIRE(engine != 2 || dev_ok, OSB_ERR_UNSUPPORTED,
              "device-resident engine needs a quasi-Newton solver and a block-functor objective");
  OSB_CUDA(cudaEventRecord(ev0, ctx->stream));
  int rc;
  last_sym_sharded = last_p2p = last_fused = false;
  if (dev_ok && engine != 1) {
    last_engine = 2;
    rc = minimize_device(ls, obj, max_iter, max_ls, cb, user);
  } else {
    last_engine = 1;
    rc = minimize_host(ls, obj, max_iter, max_ls, cb, user);
  }
  OSB_CUDA(cudaEventRecord(ev1, ctx->stream));
  OSB_CUDA(cudaEventSynchronize(ev1));
  sym_settle_pingpong();
  float ms = 0.f;
  OSB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  last_ms = ms;
  last_iters = k;
  // the reference's events, same targets and messages (ls_solver.rs:38,82-86,109; bfgs.rs:68,71 and siblings)
  if (rc == OSB_OK) {
    static const char* qn_targets[] = {"bfgs", "dfp", "broyden", "bfgs_b", "dfp_b", "broyden_b", "sr1_b"};
    const char* tgt = is_qn ? qn_targets[kind - OSB_BFGS] : (kind == OSB_PROJ_NEWTON ? "projected_newton" : "solver");
    if (reason == OSB_REASON_S_NORM) log_event(2, tgt, "Minimization completed: next iterate too close");
    else if (reason == OSB_REASON_Y_NORM) log_event(2, tgt, "Minimization completed: gradient next iterate too close");
    log_event(3, "solver", "Minimization completed: convergence in " + std::to_string(k) + " iterations");
  } else if (rc == OSB_MAX_ITER_REACHED) {
    log_event(2, "solver", "Minimization completed: max iter reached during minimization");
  } else if (rc == OSB_OUT_OF_DOMAIN) {
    log_event(1, "solver", "Minimization completed: next iterate is out of domain");
  }
  if (profile_kernels) prof_collect();
  return rc;
}

__global__ void accept_stream_kernel(DevState* st) { st->f = st->fz[0]; }

// ProjectedGradientDescent / SpectralProjectedGradient on a block-functor objective: the whole iteration is ONE fused
// kernel per line-search trial (Objective::stream_trial): 4 vector reads + 2 writes, against ~14 vector passes with one
// launch per vector expression.  The first trial step (every supported search starts at t = 1) is issued together with
// the convergence scalar and g.d of x_k, so one host fetch serves has_converged, the direction scalars and the first
// Armijo test.  projected_gradient_descent.rs:50-109, spg.rs:76-145, ls_solver.rs:66-111.
int Solver::minimize_stream(LineSearch* ls, Objective* obj, int64_t max_iter, int64_t max_ls, osb_callback_fn cb, void* user) {
  k = 0;  // ls_solver.rs:74
  reason = OSB_REASON_NONE;
  trace.clear();
  LSParams& lp = ls->p;
  const bool scale = kind == OSB_SPG;
  cudaStream_t stm = ctx->stream;
  if (!have_eval) {
    obj->eval(x.p, &d_state->f, g.p, nullptr);
    have_eval = true;
  }
  const double t0 = 1.0;  // BackTracking(B), GLLQuadratic and NoSearch all start from t = 1
  const bool proj0 = lp.kind == LS_BACKTRACKING_B;
  double* fz = d_state->fz;
  while (max_iter > k) {
    obj->stream_trial(x.p, g.p, lb.p, ub.p, lambda, scale, t0, proj0, ls->lb.p, ls->ub.p, xt.p, gt.p, fz);
    ctx->counters[2]++;
    fetch_state();
    const double f = h_state->f;
    if (is_bad(f)) return OSB_OUT_OF_DOMAIN;  // ls_solver.rs:37-40
    if (rmax(0.0, h_state->fz[7]) < tol) {    // projected_gradient_descent.rs:76-83 (number.rs:27-31 folds from 0.0)
      reason = OSB_REASON_PROJ_GRAD_TOL;
      return OSB_OK;
    }
    LSMachine m;
    m.begin(lp, f, h_state->fz[4], max_ls, h_state->fz[5]);
    bool have_first = true;
    while (!m.done) {
      const double t = m.request(lp);
      const bool pj = m.wants_projection(lp);
      if (!(have_first && t == t0 && pj == proj0)) {
        obj->stream_trial(x.p, g.p, lb.p, ub.p, lambda, scale, t, pj, ls->lb.p, ls->ub.p, xt.p, gt.p, fz);
        ctx->counters[2]++;
        fetch_state();
      }
      have_first = false;
      m.feed(lp, h_state->fz[0], h_state->fz[1], h_state->fz[2]);
    }
    const double t = m.result;
    if (!m.last_eval_is_result) {  // next = x + t d, un-projected (ls_solver.rs:60; projected_gradient_descent.rs:103)
      obj->stream_trial(x.p, g.p, lb.p, ub.p, lambda, scale, t, false, ls->lb.p, ls->ub.p, xt.p, gt.p, fz);
      fetch_state();
    }
    if (kind == OSB_SPG) {  // spg.rs:134-143: s.y <= 0 -> lambda_max, else clamp(s.s / s.y)
      const double sy = h_state->fz[3], ss = h_state->fz[2];
      if (sy <= 0.) lambda = lambda_max;
      else lambda = rmax(rmin(ss / sy, lambda_max), lambda_min);
    }
    std::swap(x.p, xt.p);
    std::swap(g.p, gt.p);
    accept_stream_kernel<<<1, 1, 0, stm>>>(d_state);
    ctx->counters[0]++;
    have_eval = true;
    if (record_trace) trace.push_back(TraceRec{f, t, NAN, NAN});
    k += 1;  // ls_solver.rs:104
    if (cb) {
      ctx->sync();
      cb(user, reinterpret_cast<osb_solver*>(this));
    }
  }
  return OSB_MAX_ITER_REACHED;  // ls_solver.rs:109-110
}

int Solver::minimize_host(LineSearch* ls, Objective* obj, int64_t max_iter, int64_t max_ls, osb_callback_fn cb, void* user) {
  last_stream = false;
  if ((kind == OSB_PGD || kind == OSB_SPG) && opt_stream != 0 && obj->has_stream_trial() &&
      (ls->p.kind == LS_BACKTRACKING || ls->p.kind == LS_BACKTRACKING_B || ls->p.kind == LS_GLL || ls->p.kind == LS_NOSEARCH)) {
    last_stream = true;
    return minimize_stream(ls, obj, max_iter, max_ls, cb, user);
  }
  if (is_qn && qn_schedule != 1) {  // eager: works on the exact, full matrix (lazy: the stored matrix may lag / be packed)
    flush_pending();
    ensure_full();
  }
  k = 0;  // ls_solver.rs:74
  reason = OSB_REASON_NONE;
  trace.clear();
  const bool needs_h = kind_needs_hessian(kind);
  const bool needs_y = kind_needs_y(kind);
  bool have_hess = false;
  LSParams& lp = ls->p;
  cudaStream_t stm = ctx->stream;
  while (max_iter > k) {
    // ---- evaluate_x_k (ls_solver.rs:32-42); cached when the point was already evaluated
    if (!have_eval || (needs_h && !have_hess)) {
      obj->eval(x.p, &d_state->f, g.p, needs_h ? hess.p : nullptr);
      have_eval = true;
      have_hess = needs_h;
      u_valid = false;
    }
    if (is_qn && !u_valid) recompute_u();
    compute_conv_scalar(obj);
    bool dir_done = false;
    if (!needs_h) {  // cheap directions are issued speculatively so that one fetch serves both decisions
      compute_direction(obj, ls);
      dir_done = true;
    }
    fetch_state();
    const double f = h_state->f;
    if (is_qn || kind == OSB_PROJ_NEWTON) {
      has_s = h_state->has_s != 0;
      has_y = h_state->has_y != 0;
      s_norm = h_state->s_norm;
      y_norm = h_state->y_norm;
    }
    if (is_bad(f)) return OSB_OUT_OF_DOMAIN;  // ls_solver.rs:37-40
    // ---- has_converged
    bool conv = false;
    if (is_qn || kind == OSB_PROJ_NEWTON) {
      if (has_s && s_norm < tol) { reason = OSB_REASON_S_NORM; conv = true; }        // bfgs.rs:67-69
      else if (has_y && y_norm < tol) { reason = OSB_REASON_Y_NORM; conv = true; }   // bfgs.rs:70-72
      else if (is_qn) {
        if (sqrt(h_state->conv) < tol) { reason = OSB_REASON_GRAD_TOL; conv = true; }  // bfgs.rs:74
      } else if (rmax(0.0, h_state->conv) < tol) { reason = OSB_REASON_PROJ_GRAD_TOL; conv = true; }
    } else if (kind == OSB_GD || kind == OSB_PNORM) {
      if (h_state->conv < tol) { reason = OSB_REASON_GRAD_TOL; conv = true; }
    } else if (kind == OSB_NEWTON) {
      if (has_dec && decrement_squared * 0.5 < tol) { reason = OSB_REASON_NEWTON_DECREMENT; conv = true; }
    } else {
      if (rmax(0.0, h_state->conv) < tol) { reason = OSB_REASON_PROJ_GRAD_TOL; conv = true; }
    }
    if (conv) return OSB_OK;
    if (!dir_done) {
      int rc = compute_direction(obj, ls);
      if (rc != OSB_OK) return rc;
      fetch_state();
      if (kind == OSB_NEWTON && !newton_singular) {
        decrement_squared = h_state->dinf;
        has_dec = true;
      }
    }
    // ---- line search: scalar automaton on the host, one fused trial kernel per requested step
    LSMachine m;
    m.begin(lp, f, h_state->gd0, max_ls, h_state->tmaxc);
    while (!m.done) {
      const double t = m.request(lp);
      obj->trial(x.p, d.p, t, m.wants_projection(lp), ls->lb.p, ls->ub.p, xt.p, gt.p, &d_state->ft);
      ctx->counters[2]++;
      fetch_state();
      m.feed(lp, h_state->ft, h_state->gdt, h_state->dn);
    }
    const double t = m.result;
    const bool current = m.last_eval_is_result;
    // ---- update_next_iterate: next = x + t*d (ls_solver.rs:60)
    if (!current) vec_axpy_project(ctx, n, x.p, d.p, t, false, nullptr, nullptr, xt.p, nullptr);
    if (needs_y) {
      if (!current) obj->eval(xt.p, &d_state->ft, gt.p, nullptr);  // bfgs.rs:98 (re-used when already evaluated)
      vec_sy(ctx, n, xt.p, x.p, gt.p, g.p, s.p, y.p, &d_state->ss);
      if (is_qn || kind == OSB_PROJ_NEWTON) state_finish_sy(ctx, d_state, tol);
      accept_trial_kernel<<<1, 1, 0, stm>>>(d_state);
      ctx->counters[0]++;
      std::swap(x.p, xt.p);
      std::swap(g.p, gt.p);
      have_eval = true;
      have_hess = false;
      if (is_qn) qn_after_step();
      if (kind == OSB_SPG || kind == OSB_SPN) {  // spg.rs:134-143
        fetch_state();
        const double sy = h_state->ys;
        if (sy <= 0.) lambda = lambda_max;
        else lambda = rmax(rmin(h_state->ss / sy, lambda_max), lambda_min);
      }
    } else {
      std::swap(x.p, xt.p);
      if (current) {
        std::swap(g.p, gt.p);
        accept_trial_kernel<<<1, 1, 0, stm>>>(d_state);
        ctx->counters[0]++;
        have_eval = true;
      } else {
        have_eval = false;
      }
      have_hess = false;
    }
    if (record_trace) {
      if (is_qn || kind == OSB_PROJ_NEWTON) fetch_state();
      trace.push_back(TraceRec{f, t, (is_qn || kind == OSB_PROJ_NEWTON) ? h_state->s_norm : NAN,
                               (is_qn || kind == OSB_PROJ_NEWTON) ? h_state->y_norm : NAN});
    }
    k += 1;  // ls_solver.rs:104
    if (cb) {
      fetch_state();
      if (is_qn || kind == OSB_PROJ_NEWTON) {
        has_s = has_y = true;
        s_norm = h_state->s_norm;
        y_norm = h_state->y_norm;
      }
      cb(user, reinterpret_cast<osb_solver*>(this));
    }
  }
  fetch_state();
  if (is_qn || kind == OSB_PROJ_NEWTON) {
    has_s = h_state->has_s != 0;
    has_y = h_state->has_y != 0;
    s_norm = h_state->s_norm;
    y_norm = h_state->y_norm;
  }
  return OSB_MAX_ITER_REACHED;  // ls_solver.rs:109-110
}

int Solver::minimize_device(LineSearch* ls, Objective* obj, int64_t max_iter, int64_t max_ls, osb_callback_fn cb, void* user) {
  if (qn_schedule != 1) flush_pending();
  k = 0;
  reason = OSB_REASON_NONE;
  trace.clear();
  cudaStream_t stm = ctx->stream;
  if (!have_eval) {
    obj->eval(x.p, &d_state->f, g.p, nullptr);
    have_eval = true;
    u_valid = false;
  }
  if (!u_valid) recompute_u();
  // control block: keep f / norms, reset the run flags (on the device: no host round trip before the first launch)
  reset_run_flags_kernel<<<1, 1, 0, stm>>>(d_state);
  ctx->counters[0]++;
  if (cb != nullptr || record_trace) fetch_state();  // (the trace wants f before the first iteration)
  if (!d_ls_buf) d_ls_buf = (LSParams*)pool_get(0, sizeof(LSParams));
  LSParams* d_ls = d_ls_buf;
  OSB_CUDA(cudaMemcpyAsync(d_ls, &ls->p, sizeof(LSParams), cudaMemcpyHostToDevice, stm));
  // polling: a snapshot of the control block every POLL iterations, at most two in flight
  const int POLL = 4;
  const int64_t ITER_CHUNK = 32;  // fused iteration kernel: outer iterations per cooperative launch
  if (!poll_snap) poll_snap = (DevState*)pool_get(1, 2 * sizeof(DevState));
  DevState* snap = poll_snap;
  std::memset(snap, 0, 2 * sizeof(DevState));
  cudaEvent_t sev[2];
  OSB_CUDA(cudaEventCreateWithFlags(&sev[0], cudaEventDisableTiming));
  OSB_CUDA(cudaEventCreateWithFlags(&sev[1], cudaEventDisableTiming));
  bool pending[2] = {false, false};
  int slot = 0;
  bool stop = false;
  const bool ls_bounded = ls->p.kind == LS_BACKTRACKING_B || ls->p.kind == LS_MORETHUENTE_B;
  // lazy schedule + cluster head: the O(n) epilogue of the pass runs at the top of the next head (8 SMs instead of 1)
  epi_p2p = ctx->world > 1 && ctx->p2p_ready && ld <= XCHG_LD && use_p2p;
  // packed symmetric storage sharded by tile pairs: needs the fused peer-memory exchange, the cluster head (it sums the
  // per-rank slots) and an even tile count; a matrix that exists as row blocks is redistributed once (ensure_packed)
  sym_sharded = ctx->world > 1 && qn_storage == 1 && qn_schedule == 1 && epi_p2p && h_symmetric && (qn_kind == QN_BFGS || qn_kind == QN_DFP) &&
                n % 16 == 0 && n <= XSLOT_LD && (n / 16) >= ctx->world && (H_virtual_identity || sym_current || n % ctx->world == 0) &&
                qn_device_head_is_cluster(obj->functor_kind(), n, head_variant);
  if (!sym_sharded && sym_current && ctx->world > 1) sym_to_full();
  last_sym_sharded = sym_sharded;
  last_p2p = epi_p2p;
  // whole iterations in one cooperative kernel (qn_iter.cu): packed lazy schedule on one GPU or sharded by tile pairs
  // (auto: on several GPUs, where it removes the fold / exchange kernel, the replicated cluster head and two launch
  //  boundaries per iteration.  On one GPU the three-launch path is 1-3 % faster without callbacks — the stand-alone pass
  //  kernel compiles to a tighter loop — and 4 % slower with a per-iteration callback (0.462 against 0.445 ms, the fused
  //  kernel publishes its snapshots from inside the launch for free, profiles/r02_callbacks.md); auto does NOT switch on
  //  the presence of a callback, because the two paths sum in different orders and a callback must not change the bits)
  const bool fused_wanted = opt_fused > 0 || (opt_fused < 0 && ctx->world > 1);
  iter_path = fused_wanted && !profile_kernels && qn_schedule == 1 && qn_storage == 1 && h_symmetric &&
              (qn_kind == QN_BFGS || qn_kind == QN_DFP) && (ctx->world == 1 || sym_sharded) && head_variant == 0 &&
              (qn_variant & 7) == 0 && qn_iter_supported(ctx, obj->functor_kind(), n, ctx->world);
  if (iter_path) {
    if (sym_current) sym_settle_pingpong();
    ensure_packed();
    const int g_it = qn_iter_grid(ctx);
    if (colpart.p == nullptr || colpart_grid < g_it) {
      colpart.alloc_pooled((int64_t)g_it * 2 * ld);
      colpart_grid = g_it;
    }
    if (gpart.p == nullptr) {  // (zeroed once: sequence numbers of the flagged reductions start at 1 and never repeat)
      gpart.alloc_pooled(qn_iter_gpart_doubles(ctx));
      gpart.zero(stm);
    }
    if (profile_iter && !d_iter_prof) {
      OSB_CUDA(cudaMalloc(&d_iter_prof, 16 * sizeof(long long)));
    }
    if (profile_iter) OSB_CUDA(cudaMemsetAsync(d_iter_prof, 0, 16 * sizeof(long long), stm));
    QNIterArgs a{};
    a.n = n;
    a.ld = ld;
    a.tol = tol;
    a.max_ls = max_ls;
    a.kind = qn_kind;
    a.iters = 0;
    a.epi_only = 0;
    a.st = d_state;
    a.lsp = d_ls;
    a.x = x.p;
    a.g = g.p;
    a.s = s.p;
    a.y = y.p;
    a.u = u.p;
    a.ps = ps.p;
    a.ph = ph.p;
    a.h = h.p;
    a.w = wv.p;
    a.lb = bounded ? lb.p : nullptr;
    a.ub = bounded ? ub.p : nullptr;
    a.ls_lb = ls_bounded ? ls->lb.p : nullptr;
    a.ls_ub = ls_bounded ? ls->ub.p : nullptr;
    a.P = Hsym.p;
    a.colpart = colpart.p;
    a.gpart = gpart.p;
    a.world = sym_sharded ? ctx->world : 1;
    a.rank = sym_sharded ? ctx->rank : 0;
    a.peers = sym_sharded ? ctx->d_peers : nullptr;
    a.seq = ctx->d_seq;
    a.prof = profile_iter ? d_iter_prof : nullptr;
    iter_args = a;
    iter_fn_kind = obj->functor_kind();
    iter_fn_a = obj->functor_ptr(0);
    iter_fn_b = obj->functor_ptr(1);
    iter_ls_kind = ls->p.kind;
    defer_epi = false;  // the epilogue is the fused kernel's own business
  }
  last_fused = iter_path;
  defer_epi = !iter_path && qn_schedule == 1 && (qn_kind == QN_BFGS || qn_kind == QN_DFP) && n > QN_SMALL_N &&
              (qn_storage == 1 || qn_variant == 0) &&
              qn_device_head_is_cluster(obj->functor_kind(), n, head_variant);
  // ---- run-ahead delivery of callbacks / trace records (see engine.cuh: callback_run_ahead)
  const bool run_ahead = (cb != nullptr || record_trace) && callback_run_ahead != 0;
  cudaEvent_t cbev[2] = {nullptr, nullptr};
  int cb_slot = 0, cb_prev = -1;
  double cb_f_before = 0.0;
  if (run_ahead) {
    if (!cb_snap) {
      cb_snap = (DevState*)pool_get(1, 2 * sizeof(DevState));
      cb_xsnap = (double*)pool_get(1, 2 * sizeof(double) * (size_t)ld);
      cb_gsnap = (double*)pool_get(1, 2 * sizeof(double) * (size_t)ld);
    }
    OSB_CUDA(cudaEventCreateWithFlags(&cbev[0], cudaEventDisableTiming));
    OSB_CUDA(cudaEventCreateWithFlags(&cbev[1], cudaEventDisableTiming));
    cb_f_before = h_state->f;
  }
  // Fused iteration kernel + run-ahead: the kernel itself writes every iteration's snapshot into a device ring and raises
  // a flag in pinned host memory (QNIterArgs.snap_*); the host copies the slot out on a side stream.  A launch therefore
  // still runs SNAP_CHUNK iterations while the host delivers the callbacks behind it.
  struct SnapLaunch {
    int half, count;
    unsigned long long seq0;
  };
  const bool snap_mode = run_ahead && iter_path;
  SnapLaunch snap_prev{0, 0, 0ULL};
  int snap_half = 0;
  if (run_ahead && !snap_x) {
    snap_x = (double*)pool_get(0, 2 * SNAP_CHUNK * 2 * sizeof(double) * (size_t)ld);
    snap_st = (DevState*)pool_get(0, 2 * SNAP_CHUNK * sizeof(DevState));
    snap_flag = (unsigned long long*)pool_get(1, 2 * SNAP_CHUNK * sizeof(unsigned long long));
    OSB_CUDA(cudaStreamCreateWithFlags(&snap_stream, cudaStreamNonBlocking));
  }
  // delivers the iterations of one launch in order; false = the solve ended inside it (or before it)
  auto deliver_launch = [&](const SnapLaunch& L) -> bool {
    for (int j = 0; j < L.count; ++j) {
      volatile unsigned long long* fl = snap_flag + (size_t)L.half * SNAP_CHUNK + j;
      const unsigned long long want = L.seq0 + (unsigned long long)j + 1ULL;
      bool launch_over = false;
      for (unsigned spins = 0; *fl != want; ++spins) {
        if ((spins & 255u) != 255u) continue;
        if (launch_over) return false;  // the kernel has ended and never published slot j: it found `done` earlier
        const cudaError_t q = cudaEventQuery(cbev[L.half]);
        if (q == cudaSuccess) launch_over = true;  // (one more look at the flag before giving up)
        else if (q != cudaErrorNotReady) OSB_CUDA(q);
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      ctx->counters[3]++;
      // the slot is complete in device memory: copy it out beside the running kernel (copy engine, side stream)
      const size_t slot = (size_t)L.half * SNAP_CHUNK + j;
      OSB_CUDA(cudaMemcpyAsync(&cb_snap[0], snap_st + slot, sizeof(DevState), cudaMemcpyDeviceToHost, snap_stream));
      OSB_CUDA(cudaMemcpyAsync(cb_xsnap, snap_x + slot * 2 * ld, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, snap_stream));
      if (cb) OSB_CUDA(cudaMemcpyAsync(cb_gsnap, snap_x + slot * 2 * ld + ld, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, snap_stream));
      OSB_CUDA(cudaStreamSynchronize(snap_stream));
      const DevState& sn = cb_snap[0];
      if (sn.done) return false;
      k = sn.k;
      has_s = has_y = true;
      s_norm = sn.s_norm;
      y_norm = sn.y_norm;
      if (record_trace) trace.push_back(TraceRec{cb_f_before, sn.t_last, s_norm, y_norm});
      cb_f_before = sn.f;
      if (cb) {
        cb_x_mirror = cb_xsnap;
        cb_g_mirror = cb_gsnap;
        cb_state_mirror = &cb_snap[0];
        cb(user, reinterpret_cast<osb_solver*>(this));
        cb_x_mirror = nullptr;
        cb_g_mirror = nullptr;
        cb_state_mirror = nullptr;
      }
    }
    return true;
  };
  // returns false when the snapshot says the head found convergence at the START of that iteration (no k += 1, no callback)
  auto deliver = [&](int sl) -> bool {
    OSB_CUDA(cudaStreamWaitEvent(snap_stream, cbev[sl], 0));
    OSB_CUDA(cudaMemcpyAsync(&cb_snap[sl], snap_st + sl, sizeof(DevState), cudaMemcpyDeviceToHost, snap_stream));
    OSB_CUDA(cudaMemcpyAsync(cb_xsnap + (size_t)sl * ld, snap_x + (size_t)sl * 2 * ld, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, snap_stream));
    if (cb) OSB_CUDA(cudaMemcpyAsync(cb_gsnap + (size_t)sl * ld, snap_x + (size_t)sl * 2 * ld + ld, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, snap_stream));
    OSB_CUDA(cudaStreamSynchronize(snap_stream));
    ctx->counters[3]++;
    const DevState& sn = cb_snap[sl];
    if (sn.done) return false;
    k = sn.k;
    has_s = has_y = true;
    s_norm = sn.s_norm;
    y_norm = sn.y_norm;
    if (record_trace) trace.push_back(TraceRec{cb_f_before, sn.t_last, s_norm, y_norm});
    cb_f_before = sn.f;
    if (cb) {
      cb_x_mirror = cb_xsnap + (size_t)sl * ld;
      cb_g_mirror = cb_gsnap + (size_t)sl * ld;
      cb_state_mirror = &cb_snap[sl];
      cb(user, reinterpret_cast<osb_solver*>(this));
      cb_x_mirror = nullptr;
      cb_g_mirror = nullptr;
      cb_state_mirror = nullptr;
    }
    return true;
  };
  int64_t chunk = 1;
  for (int64_t it = 0; it < max_iter && !stop; it += chunk) {
    if (iter_path) {
      // up to POLL iterations per launch; one per launch when a callback / trace wants every iteration's state
      // (a launch that finds convergence simply ends, and later launches return at once: the chunk only bounds how far
      //  the host runs ahead of the device)
      chunk = snap_mode ? std::min<int64_t>(SNAP_CHUNK, max_iter - it)
                        : (cb != nullptr || record_trace) ? 1 : std::min<int64_t>(ITER_CHUNK, max_iter - it);
      QNIterArgs a = iter_args;
      a.iters = (int)chunk;
      if (snap_mode) {  // this launch publishes its iterations into one half of the pinned ring
        a.snap_x = snap_x + (size_t)snap_half * SNAP_CHUNK * 2 * ld;
        a.snap_st = snap_st + (size_t)snap_half * SNAP_CHUNK;
        a.snap_flag = snap_flag + (size_t)snap_half * SNAP_CHUNK;
        a.snap_seq0 = g_snap_seq.fetch_add((unsigned long long)chunk);
      }
      qn_launch_iter(ctx, iter_fn_kind, iter_fn_a, iter_fn_b, bounded, iter_ls_kind, a);
      if (snap_mode) {
        OSB_CUDA(cudaEventRecord(cbev[snap_half], stm));
        SnapLaunch cur{snap_half, (int)chunk, a.snap_seq0};
        lazy_used = true;
        u_valid = true;
        if (sym_sharded) {
          ctx->counters[4] += chunk;
          ctx->counters[5] += chunk;
        }
        // deliver the PREVIOUS launch's iterations while this one runs (its ring half is the other one)
        if (snap_prev.count > 0 && !deliver_launch(snap_prev)) {
          snap_prev.count = 0;
          break;
        }
        snap_prev = cur;
        snap_half ^= 1;
        continue;
      }
      if (sym_sharded) {
        ctx->counters[4] += chunk;  // fused exchanges
        ctx->counters[5] += chunk;  // passes over the sharded packed triangle
      }
      lazy_used = true;
      u_valid = true;
    } else {
      qn_device_launch_head(ctx, obj->functor_kind(), obj->functor_ptr(0), obj->functor_ptr(1), bounded, d_ls, n, tol, max_ls,
                            d_state, x.p, g.p, d.p, xt.p, gt.p, s.p, y.p, u.p, bounded ? lb.p : nullptr, bounded ? ub.p : nullptr,
                            ls_bounded ? ls->lb.p : nullptr, ls_bounded ? ls->ub.p : nullptr, head_variant, ls->p.kind, head_epi());
      qn_after_step();
    }
    if (run_ahead) {
      // snapshot of this iteration, then keep going: the previous iteration's callback runs while the device works
      snap_copy_kernel<<<16, 256, 0, stm>>>(x.p, g.p, d_state, n, ld, snap_x + (size_t)cb_slot * 2 * ld, snap_st + cb_slot);
      ctx->counters[0]++;
      OSB_CUDA(cudaEventRecord(cbev[cb_slot], stm));
      if (cb_prev >= 0 && !deliver(cb_prev)) {
        cb_prev = -1;
        break;
      }
      cb_prev = cb_slot;
      cb_slot ^= 1;
      continue;
    }
    if (cb != nullptr || record_trace) {
      finish_epilogue();  // the callback may read H or restart: leave no epilogue owed
      // a host callback (ls_solver.rs:105-107) or a trace needs the state after every iteration: one
      // synchronisation per outer iteration, still none inside the line search
      const double f_before = h_state->f;
      fetch_state();
      if (h_state->done) break;  // the head found convergence at the start of this iteration: no k += 1, no callback
      k = h_state->k;
      has_s = has_y = true;
      s_norm = h_state->s_norm;
      y_norm = h_state->y_norm;
      if (record_trace) trace.push_back(TraceRec{f_before, h_state->t_last, s_norm, y_norm});
      if (cb) cb(user, reinterpret_cast<osb_solver*>(this));
      continue;
    }
    if (iter_path || (it + 1) % POLL == 0) {
      if (pending[slot]) {  // bound the run-ahead: wait for the older snapshot of this slot
        OSB_CUDA(cudaEventSynchronize(sev[slot]));
        ctx->counters[3]++;
        if (snap[slot].done) stop = true;
        pending[slot] = false;
      }
      if (!stop) {
        OSB_CUDA(cudaMemcpyAsync(&snap[slot], d_state, sizeof(DevState), cudaMemcpyDeviceToHost, stm));
        OSB_CUDA(cudaEventRecord(sev[slot], stm));
        pending[slot] = true;
        slot ^= 1;
        // opportunistic early look (single GPU only: ranks must take identical stop decisions)
        if (ctx->world == 1 && pending[slot] && cudaEventQuery(sev[slot]) == cudaSuccess) {
          if (snap[slot].done) stop = true;
          pending[slot] = false;
        }
      }
    }
  }
  if (run_ahead) {
    if (snap_prev.count > 0) deliver_launch(snap_prev);
    if (cb_prev >= 0) deliver(cb_prev);
    cudaEventDestroy(cbev[0]);
    cudaEventDestroy(cbev[1]);
  }
  finish_epilogue();
  defer_epi = false;
  sym_sharded = false;
  iter_path = false;
  fetch_state();
  OSB_CUDA(cudaMemcpyAsync(&ls->p, d_ls, sizeof(LSParams), cudaMemcpyDeviceToHost, stm));
  ctx->sync();

  cudaEventDestroy(sev[0]);
  cudaEventDestroy(sev[1]);
  k = h_state->k;
  has_s = h_state->has_s != 0;
  has_y = h_state->has_y != 0;
  s_norm = h_state->s_norm;
  y_norm = h_state->y_norm;
  ctx->counters[2] += h_state->ls_evals;
  if (h_state->done) {
    reason = h_state->reason;
    return h_state->status;
  }
  return OSB_MAX_ITER_REACHED;
}

}  // namespace osb
