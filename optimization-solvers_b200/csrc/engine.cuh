// engine.cuh — host-side objects behind the C ABI: context, objectives, line searches, solvers.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "ls_automaton.cuh"
#include "reduce.cuh"

namespace osb {

// ---- device-resident control block of one solver ------------------------------------------
// Written by kernels, mirrored to pinned host memory when the host needs a decision.  Groups of
// fields that one reduction kernel produces are contiguous (the kernel writes K consecutive doubles).
struct DevState {
  double f;       // f(x_k) of the cached evaluate_x_k (ls_solver.rs:32-42)
  double conv;    // kind-specific convergence scalar: ||g||^2 | max|g_i| | ||proj grad||_inf
  double gd0;     // g(x_k) . d_k
  double tmaxc;   // MoreThuenteB feasible-step candidate (morethuente_b.rs:185-197)
  double dinf;    // spare (SPG lambda0: ||P(x-g)-x||_inf)
  // trial outputs (3 contiguous)
  double ft, gdt, dn;
  // post-step dots (4 contiguous): s.s, y.y, y.s, spare
  double ss, yy, ys, sp;
  // update dots (2 contiguous): y.h, spare
  double yh, sp2;
  double s_norm, y_norm;
  double c0, c1, c2;  // rank-2 update coefficients (kind-specific, see qn_kernels.cu)
  double pc0, pc1, pc2;  // lazy schedule: coefficients of the update that is still PENDING on the stored matrix
  double t_last;
  // fused streaming trial of PGD / SPG (8 contiguous): f_t, g_t.d, ||x_t - x||^2, s.y, g.d, feasible-step candidate,
  // ||d||_inf, ||projected gradient||_inf
  double fz[8];
  long long k;
  int has_s, has_y;
  int skip;    // bfgs.rs:106-112: s_norm < tol || y_norm < tol -> H not updated
  int done;    // device-resident engine: minimize() has terminated
  int status;  // OSB_* status when done
  int reason;  // OSB_REASON_*
  int ls_evals;
  int pending;  // lazy schedule: stored matrix = H - rank2(ps, ph; pc0..pc2) still to be applied
  int pp;       // ping-pong packed storage (pass variant bit 1): 0 = the current matrix is in Hsym, 1 = in Hsym2; toggled on
                // the device by the fold kernel, so that predicated (no-op) launches after `done` do not flip it
  int ll_seq;   // fused iteration kernel: sequence number of its flagged grid reductions (never reused: stale words of an
                // earlier launch can then not be mistaken for fresh ones)
  int epi;      // lazy schedule: h = H y and w = H g are fresh and their O(n) epilogue is still owed (deferred to the next
                // cluster head, or to qn_launch_epilogue_cluster when no head follows)
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;
  // peer-memory exchange region (CUDA IPC): the fused all-gather of the lazy quasi-Newton pass writes
  // its row sums straight into every peer's region over NVLink and synchronises with flags
  double* xchg = nullptr;            // local region: 4 vectors of XCHG_LD doubles, then `world` 64-bit flags
  double** d_peers = nullptr;        // device array[world] of region base pointers (own entry = xchg)
  unsigned long long* d_seq = nullptr;  // exchange sequence number (device)
  bool p2p_ready = false;
  std::vector<void*> peer_opened;
  // reduction scratch
  double* red_partials = nullptr;
  unsigned int* red_ticket = nullptr;
  unsigned int* gemv_ticket = nullptr;
  int red_max_grid = 0;
  // counters: launches, objective evals, ls trials, host syncs, collectives
  int64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // pinned staging for small transfers
  double* h_pinned = nullptr;  // 4096 doubles
  double* d_dummy = nullptr;   // 8 doubles: sink for reductions whose result is unused
  // Index-range sharding of the O(n) solvers (GD / PGD / SPG with block-functor objectives, SURVEY 8e "SPG / PGD
  // streaming"): every rank owns a contiguous slice of x, g, lb, ub; every map-reduce of the path then ends in an
  // all-gather of the K per-rank values and a rank-ordered combine (same bits on every rank).
  bool vec_sharded = false;
  double* shard_scratch = nullptr;  // world x 8 doubles

  explicit Ctx(int dev);
  ~Ctx();
  void sync() {
    OSB_CUDA(cudaStreamSynchronize(stream));
    counters[3]++;
  }
  void all_gather_inplace(double* buf, int64_t count_per_rank);  // dist.cu
  void use() { OSB_CUDA(cudaSetDevice(device)); }
  int red_grid(int64_t count) const {
    int64_t g = (count + RED_THREADS * 8 - 1) / (RED_THREADS * 8);
    if (g < 1) g = 1;
    if (g > red_max_grid) g = red_max_grid;
    return (int)g;
  }
};

// RAII device buffer
struct DBuf {
  double* p = nullptr;
  int64_t n = 0;
  DBuf() {}
  explicit DBuf(int64_t n_) { alloc(n_); }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }
  void alloc(int64_t n_);
  // all buffers come from / return to the per-device pool (engine.cu); contents are NOT preserved or cleared
  void alloc_pooled(int64_t n_);
  void release();
  void zero(cudaStream_t s);
  void upload(const double* h, int64_t cnt, cudaStream_t s);
  void download(double* h, int64_t cnt, cudaStream_t s) const;
};

#ifdef __CUDACC__
template <int K, class Fin>
__global__ void shard_combine_kernel(Fin fin, RedOps<K> ops, const double* gathered, int world, double* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double v[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    v[k] = red_identity(ops.op[k]);
    for (int r = 0; r < world; ++r) v[k] = red_combine(ops.op[k], v[k], gathered[r * 8 + k]);
  }
  fin(v, out);
}
#endif

template <int K, class F, class Fin>
inline void launch_mapreduce_fin(Ctx* ctx, F f, Fin fin, int64_t count, RedOps<K> ops, double* out) {
  static_assert(K <= RED_THREADS / 32, "final fold uses one warp per output");
  static_assert(K <= 8, "shard_scratch holds 8 values per rank");
  int grid = ctx->red_grid(count);
  if (ctx->vec_sharded && ctx->world > 1) {
    // local fold -> K raw values per rank -> all-gather -> rank-ordered combine + the caller's finalizer
    mapreduce_kernel<K, F, FinStore><<<grid, RED_THREADS, 0, ctx->stream>>>(f, FinStore{}, count, ops, ctx->red_partials, ctx->red_ticket,
                                                                           ctx->shard_scratch + ctx->rank * 8);
    ctx->all_gather_inplace(ctx->shard_scratch, 8);
    shard_combine_kernel<K, Fin><<<1, 32, 0, ctx->stream>>>(fin, ops, ctx->shard_scratch, ctx->world, out);
    ctx->counters[0] += 2;
    return;
  }
  mapreduce_kernel<K, F, Fin><<<grid, RED_THREADS, 0, ctx->stream>>>(f, fin, count, ops, ctx->red_partials, ctx->red_ticket, out);
  ctx->counters[0]++;
}
template <int K, class F>
inline void launch_mapreduce(Ctx* ctx, F f, int64_t count, RedOps<K> ops, double* out) {
  launch_mapreduce_fin<K>(ctx, f, FinStore{}, count, ops, out);
}

// ---- objectives ---------------------------------------------------------------------------
enum FunctorKind : int { FN_NONE = 0, FN_ROSENBROCK = 1, FN_SEPQUAD = 2 };

struct Objective {
  Ctx* ctx;
  int64_t n;
  int64_t calls = 0;
  Objective(Ctx* c, int64_t n_) : ctx(c), n(n_) {}
  virtual ~Objective() {}
  virtual bool provides_hessian() const { return false; }
  virtual int functor_kind() const { return FN_NONE; }
  virtual const double* functor_ptr(int) const { return nullptr; }
  // f -> *d_f (device), g -> d_g[n], Hessian -> d_hess[n*n] (row-major; symmetric) when non-null
  virtual void eval(const double* d_x, double* d_f, double* d_g, double* d_hess) = 0;
  // one line-search trial: xt = [P](x + t d); (ft, gt) = eval(xt); out3 = {ft, gt.d, ||xt - x||^2}
  virtual void trial(const double* x, const double* d, double t, bool project, const double* lb, const double* ub,
                     double* xt, double* gt, double* d_out3);
  // Fused trial step of the projected-gradient family (PGD / SPG, block-functor objectives): ONE kernel reads x, g, lb, ub
  // and writes x_t, g_t; the direction d = P(x - lam g) - x is formed on the fly and never stored.
  //   x_t = [P_ls](x + t d),  (f_t, g_t) = eval(x_t)
  //   out8 = {f_t, g_t.d, ||x_t - x||^2, (x_t - x).(g_t - g), g.d, min feasible step, ||d||_inf, ||proj grad(x)||_inf}
  // Returns false when the objective has no fused form (the caller then takes one launch per vector expression).
  virtual bool has_stream_trial() const { return false; }
  virtual bool stream_trial(const double*, const double*, const double*, const double*, double, bool, double, bool, const double*,
                            const double*, double*, double*, double*) {
    return false;
  }
};

Objective* make_dense_quadratic(Ctx*, int64_t n, const double* A_host, const double* b_host);
Objective* make_dense_quadratic_generated(Ctx*, int64_t n, bool shifted, double* x0_host);
Objective* make_rosenbrock(Ctx*, int64_t n);
Objective* make_sepquad_generated(Ctx*, int64_t n, int64_t index0 = 0);
Objective* make_logistic_generated(Ctx*, int64_t m, int64_t n, double lambda);
Objective* make_host_objective(Ctx*, int64_t n, osb_host_eval_fn fn, void* user, bool with_h);
Objective* make_user_objective(Ctx*, int64_t n, osb_device_eval_fn fn, void* user, bool with_h);

// ---- line search handle -------------------------------------------------------------------
struct LineSearch {
  LSParams p;
  Ctx* ctx = nullptr;
  DBuf lb, ub;  // bounded kinds
  int64_t n = 0;
};

// ---- dense quasi-Newton kernels (qn_kernels.cu) -------------------------------------------
enum QNKind : int { QN_BFGS = 0, QN_DFP = 1, QN_BROYDEN = 2, QN_SR1 = 3 };
// out = H v over the local row block.  `sel` chooses the operand at run time ON DEVICE from
// DevState.skip: skip ? (v_skip -> out_skip) : (v -> out); a null `st` means unconditional (v -> out).
// optional epilogue of the gemv launch (single GPU): the last CTA computes y.h and c0..c2
struct QNCoefArgs {
  unsigned int* ticket;  // null => no fused epilogue
  int kind;
  int64_t n;
  const double* s;
  const double* y;
  double* p_out;
};
void qn_launch_gemv(Ctx* ctx, const double* H, int64_t ld, int64_t nrows, int64_t row0, const DevState* st, const double* v,
                    double* out, const double* v_skip, double* out_skip, int variant, const QNCoefArgs* coef = nullptr);
// out[j] = sum_i H_ij s_i over the local row block (Broyden's H^T s), two-stage deterministic
void qn_launch_gemvT(Ctx* ctx, const double* H, int64_t ld, int64_t nrows, int64_t row0, const DevState* st,
                     const double* s, double* out, double* scratch);
// coefficients c0..c2 (and p = s - h for SR1/Broyden) from y.h etc.; single CTA
void qn_launch_coef(Ctx* ctx, int kind, int64_t n, DevState* st, const double* s, const double* y, const double* h,
                    double* p_out);
// fused: H <- H + rank-2(kind; p, q, r; c0..c2) and u = H' g over the local row block
void qn_launch_update(Ctx* ctx, int kind, double* H, int64_t ld, int64_t nrows, int64_t row0, const DevState* st,
                      const double* p, const double* q, const double* r, const double* g, double* u_out, int variant);
// lazy (2-pass-less) schedule: ONE read-modify-write per iteration — applies the pending update while
// computing h = H y and w = H g with the updated rows; the epilogue forms the new pending update and u.
constexpr int64_t XCHG_LD = 65536;  // capacity (doubles) of one exchanged vector
// sharded packed-symmetric storage: after the flags, 2 parities x world ranks x {h, w} slots of XSLOT_LD doubles; every
// rank's contribution (row sums of its tiles + its column sums) lands in slot `rank` of every rank's region
constexpr int64_t XSLOT_LD = 16384;
constexpr int64_t XSLOT_OFF = 4 * XCHG_LD + 64;  // (flags: world + 8 <= 64 64-bit words)
// Deferred lazy-schedule epilogue, executed by the 8-CTA cluster head (a single CTA is bound by one SM's L2 port:
// 1.3 MB of O(n) vectors took 16 us of a 450 us iteration at n = 16384).
struct HeadEpi {
  int kind;          // QN_BFGS / QN_DFP, or -1: the head never runs an epilogue
  const double* h;   // H y
  const double* w;   // H g
  const double* h2;  // peer-memory exchange: h, w of the odd-parity exchange buffers (st->epi == 2), else null
  const double* w2;
  int nslots;        // > 1: h and w are the rank-ordered sums of `nslots` vectors `slot_stride` doubles apart
  int64_t slot_stride;  //      (sharded packed storage: one slot per rank in the exchange region)
  double* u_out;     // u = H+ g (also kept for getters)
  double* ps_out;    // pending p <- s
  double* ph_out;    // pending q <- h
};
struct QNLazyArgs {
  double* M;            // stored matrix (local row block)
  int64_t ld, nrows, row0, n;
  DevState* st;
  const double* ps;     // pending p (= s of the previous iteration)
  const double* ph;     // pending q (= h of the previous iteration)
  const double* y;
  const double* g;      // new gradient
  const double* s;      // new s
  double* h;            // out: H y
  double* w;            // out: H g
  double* u;            // out (epilogue): H+ g
  double* ps_out;       // epilogue: ps <- s
  double* ph_out;       // epilogue: ph <- h
  unsigned int* ticket; // null => no fused epilogue (sharded: separate launch after the all-gather)
  int kind;
  // fused peer-memory all-gather (world > 1 with an IPC-connected context); null => NCCL path
  double* const* peers;
  unsigned long long* seq;
  int world, rank;
  int tile_rows;  // filled by qn_launch_lazy
  int defer_epi;  // 1 => do not run the epilogue in the pass: raise st->epi and let the cluster head do it on 8 SMs
};
void qn_launch_lazy(Ctx* ctx, const QNLazyArgs& a, int variant = 0);
void qn_launch_lazy_epilogue(Ctx* ctx, const QNLazyArgs& a);
// packed symmetric storage (lower triangle in 8-row tiles): one pass moves n^2 * 8 B
int64_t qn_sym_doubles(int64_t n);
int qn_sym_grid(Ctx* ctx, int64_t n, int variant);
void qn_sym_pack(Ctx* ctx, const double* H, int64_t ld, int64_t n, double* P);
void qn_sym_set_identity(Ctx* ctx, int64_t n, double* P);
int64_t qn_sym_doubles_sharded(int64_t n, int world, int rank);
void qn_sym_layout(int64_t n, int world, int64_t tile, int* owner, int64_t* offset, int64_t* lpad);
void qn_sym_set_identity_sharded(Ctx* ctx, int64_t n, double* P);
void qn_sym_pack_sharded(Ctx* ctx, const double* Hfull, int64_t ld, int64_t n, double* P);  // every rank holds the full matrix
void qn_sym_unpack_sharded(Ctx* ctx, const double* P, int64_t ld, int64_t n, double* Hfull_zeroed);
void qn_sym_unpack(Ctx* ctx, const double* P, int64_t ld, int64_t n, double* H);
void qn_launch_lazy_sym(Ctx* ctx, const QNLazyArgs& a, double* P, double* Pout, double* colpart, int64_t n, int64_t ld, int phase, int variant,
                        bool identity_unwritten = false);  // identity_unwritten: P is H_0 = I and has not been written yet
// whole outer iterations in one cooperative kernel (qn_iter.cu): head + line search + H pass + fold + exchange
constexpr int64_t XFLAG2_LD = 256;  // per-source-rank chunk flags of the fused iteration kernel (one per CTA)
HD int64_t xflag2_off(int world) { return XSLOT_OFF + 4 * (int64_t)world * XSLOT_LD; }  // doubles; after the {h, w} slots
struct QNIterArgs {
  int64_t n, ld;
  double tol;
  int64_t max_ls;
  int kind;      // QN_BFGS / QN_DFP
  int iters;     // outer iterations to run in this launch
  int epi_only;  // 1: only run the epilogue an earlier pass left owed (end of minimize, before a stalling callback)
  DevState* st;
  LSParams* lsp;
  double *x, *g, *s, *y, *u, *ps, *ph;
  double *h, *w;  // row sums of the pass; after the fold (one GPU) h = H y and w = H g
  const double *lb, *ub, *ls_lb, *ls_ub;
  double* P;        // packed matrix (this rank's tiles)
  double* colpart;  // grid x 2 x ld
  double* gpart;    // 2 x grid x IT_GPK
  int world, rank;
  double* const* peers;
  unsigned long long* seq;
  long long* prof;  // optional [16]: ns in head / pass / fold+exchange (CTA 0), iterations, head sub-phases
  // Run-ahead snapshots for host callbacks / traces (null: none), written from inside the kernel so that a launch keeps
  // running many iterations while the host delivers ls_solver.rs:104-107's callback behind it.  Iteration j of this
  // launch: x then g at snap_x + j * 2 * ld and the scalars a callback may read at snap_st + j (DEVICE ring), then
  // (st.release.sys) the flag snap_flag[j] = snap_seq0 + j + 1 in PINNED HOST memory; the host copies the slot out.
  double* snap_x;
  DevState* snap_st;
  unsigned long long* snap_flag;
  unsigned long long snap_seq0;
};

bool qn_iter_supported(Ctx* ctx, int functor_kind, int64_t n, int world);
int qn_iter_grid(Ctx* ctx);
int64_t qn_iter_gpart_doubles(Ctx* ctx);
void qn_launch_iter(Ctx* ctx, int functor_kind, const double* fn_a, const double* fn_b, bool bounded, int ls_kind, const QNIterArgs& a);
// apply a pending update to the stored matrix (getters, engine switches)
void qn_launch_flush(Ctx* ctx, int kind, double* M, int64_t ld, int64_t nrows, int64_t row0, DevState* st, const double* ps,
                     const double* ph);
int64_t qn_ld(int64_t n);
// n <= QN_SMALL_N: single-thread replay of the reference's own operation order (qn_small.cu)
constexpr int64_t QN_SMALL_N = 5;
void qn_small_step(Ctx* ctx, int kind, int64_t n, int64_t ld, double* H, DevState* st, const double* s, const double* y,
                   const double* g, double* u);
void qn_small_gemv(Ctx* ctx, int64_t n, int64_t ld, const double* H, const double* g, double* u);
int64_t qn_rows_padded(int64_t nrows);

// ---- solver -------------------------------------------------------------------------------
struct TraceRec {
  double f, t, s_norm, y_norm;
};

struct Solver {
  Ctx* ctx;
  int kind;
  int64_t n;
  double tol;
  int64_t k = 0;
  int reason = OSB_REASON_NONE;
  // options
  int engine = 0;
  int record_trace = 0;
  // callback_run_ahead = 1: the per-iteration callback / trace of the device-resident engine no longer stalls the device.
  // After iteration k the control block and x_k+1 are copied into pinned snapshots, iteration k + 1 is enqueued, and
  // only then is the callback for iteration k delivered; inside it x(), f(), k(), s_norm(), y_norm() read the
  // snapshot.  Any other getter (or a mutation) inside the callback would see the solver one iteration ahead:
  // hence an option and not the default.
  int callback_run_ahead = -1;  // -1 = auto (on), 0 = stalling delivery, 1 = on
  LSParams* d_ls_buf = nullptr;      // device copy of the line-search parameters (minimize_device)
  DevState* poll_snap = nullptr;     // 2 pinned slots of the polling snapshots
  DevState* cb_snap = nullptr;       // 2 pinned slots
  double* cb_xsnap = nullptr;        // 2 x ld doubles, pinned
  const double* cb_x_mirror = nullptr;  // non-null while a run-ahead callback is being delivered
  const double* cb_g_mirror = nullptr;
  double* cb_gsnap = nullptr;           // 2 x ld doubles, pinned
  const DevState* cb_state_mirror = nullptr;
  int qn_variant = 0;
  int head_variant = 0;  // 0 = cluster head, 1 = single-CTA smem head, 2 = generic single-CTA head
  // state vectors (device)
  DBuf x, g, d, xt, gt, s, y, lb, ub, w;
  bool bounded = false;
  // quasi-Newton
  bool is_qn = false;
  int qn_kind = QN_BFGS;
  int64_t ld = 0, row0 = 0, nrows = 0;  // local row block of H
  DBuf H, u, h, pvec, vvec, scratch;
  DBuf wv, ps, ph;        // lazy schedule: w = H g, pending p and q
  // Options as the caller set them (-1 = auto) and the values in force for the current minimize() call.  Auto picks the
  // lazy schedule and the packed symmetric storage whenever they apply (BFGS / DFP, symmetric H, n > QN_SMALL_N): a caller
  // that writes `BFGS::new(tol, x0)` + `minimize(...)` like examples/bfgs_example.rs:46-52 gets the n^2 8 B path.
  int opt_schedule = -1, opt_storage = -1;
  int last_engine = 0;         // what the last minimize() ran: 1 = host-driven, 2 = device-resident control
  bool last_sym_sharded = false, last_p2p = false, last_fused = false;
  void resolve_options();
  void recompute_u();     // u = H g from the exact matrix (first iteration, after set_x / set_inv_hessian)
  int qn_schedule = 0;    // 0 = eager (h = H y, then fused update: 3 n^2 8 B), 1 = lazy (one RMW: 2 n^2 8 B)
  bool lazy_used = false;
  bool H_virtual_identity = false;  // H = I and not materialised yet (large n: the 2 GiB buffer is allocated on first need;
                                    // with packed storage it never is)
  void ensure_full();               // materialise the full row-major H
  bool defer_epi = false;  // minimize_device with the cluster head: the lazy pass leaves its epilogue to the next head
  bool epi_p2p = false;    // ... and h, w live in the peer-memory exchange buffers
  bool sym_sharded = false;  // packed symmetric storage sharded by tile pairs over the ranks (this minimize call)
  HeadEpi head_epi() const;
  void finish_epilogue();  // run an owed epilogue now (no head follows)
  int qn_storage = 0;     // 0 = full n x n, 1 = packed symmetric lower triangle (lazy schedule, BFGS/DFP, single GPU)
  DBuf Hsym, Hsym2, colpart;  // packed matrix (+ its ping-pong twin, pass variant bit 1) and per-CTA column partials
  int colpart_grid = 0;
  bool sym_pingpong_dirty = false;
  // fused iteration kernel (qn_iter.cu): -1 = auto (on whenever it applies), 0 = off (one launch per phase)
  int opt_fused = -1;
  double* snap_x = nullptr;      // device ring of in-kernel callback snapshots (2 halves x 32 iterations x {x, g})
  DevState* snap_st = nullptr;   // device ring of their scalars
  unsigned long long* snap_flag = nullptr;  // pinned host flags
  cudaStream_t snap_stream = nullptr;       // side stream of the copies out of the ring (runs beside the kernel)
  int opt_stream = -1;           // PGD / SPG: one fused kernel per trial step (-1 / 1 = on whenever it applies, 0 = off)
  bool last_stream = false;
  bool iter_path = false;        // this minimize() runs whole iterations in one cooperative kernel
  QNIterArgs iter_args{};        // last launch parameters (the epilogue-only launch re-uses them)
  int iter_fn_kind = 0, iter_ls_kind = 0;
  const double* iter_fn_a = nullptr;
  const double* iter_fn_b = nullptr;
  DBuf gpart;                    // grid-reduction partials
  long long* d_iter_prof = nullptr;  // option "profile_iter": ns in head / pass / fold (CTA 0) and iterations
  int profile_iter = 0;
  void ensure_packed(bool allow_unwritten_identity = false);  // the packed copy holds the current matrix (identity / pack on first use)
  bool sym_ident_unwritten = false;  // the packed copy IS the current matrix, H_0 = I, but its memory has not been written:
                                     // the first pass generates the elements (one GPU, default pass variant)
  bool sym_current = false;  // the packed copy (not H) holds the current matrix
  bool h_symmetric = true;   // false after set_inv_hessian with a non-symmetric matrix (then full storage is used)
  void sym_to_full();
  void sym_settle_pingpong();  // make Hsym the buffer that holds the current matrix (host pointers swap, DevState.pp = 0)
  int use_p2p = 1;        // lazy schedule, world > 1: fused peer-memory all-gather when the context is IPC-connected
  void flush_pending();
  // Newton family
  DBuf hess, chol, lu_perm;
  bool newton_singular = false;  // this iteration's Hessian had an exactly zero pivot: d = -g (newton/mod.rs:43-46)
  bool has_dec = false;
  double decrement_squared = NAN;
  // spectral
  double lambda = 1.0, lambda_min = 1e-3, lambda_max = 1e3;
  // norms (host mirror)
  bool has_s = false, has_y = false;
  double s_norm = NAN, y_norm = NAN;
  // control block
  DevState* d_state = nullptr;
  DevState* h_state = nullptr;  // pinned
  bool have_eval = false;       // (f, g) valid for the current x
  bool u_valid = false;         // u == H g for the current (H, g)
  std::vector<TraceRec> trace;
  double last_ms = 0.0;
  int64_t last_iters = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // optional per-kernel timing (option "profile_kernels"): CUDA-event pairs around the two H passes
  int profile_kernels = 0;
  std::vector<cudaEvent_t> prof_events;  // 4 per iteration: gemv begin/end, update begin/end
  double prof_ms[3] = {0.0, 0.0, 0.0};   // mean ms: pass 1 (gemv), pass 2 (update), count
  void prof_mark();
  void prof_collect();

  Solver(Ctx* c, int kind_, int64_t n_, double tol_, const double* x0, const double* lb_h, const double* ub_h,
         Objective* obj_for_lambda0);
  ~Solver();
  int minimize(LineSearch* ls, Objective* obj, int64_t max_iter, int64_t max_ls, osb_callback_fn cb, void* user);
  void fetch_state();
  void push_state();

 private:
  int minimize_host(LineSearch* ls, Objective* obj, int64_t max_iter, int64_t max_ls, osb_callback_fn cb, void* user);
  int minimize_device(LineSearch* ls, Objective* obj, int64_t max_iter, int64_t max_ls, osb_callback_fn cb, void* user);
  int minimize_stream(LineSearch* ls, Objective* obj, int64_t max_iter, int64_t max_ls, osb_callback_fn cb, void* user);
  bool device_engine_supported(const LineSearch* ls, const Objective* obj) const;
  void compute_conv_scalar(Objective* obj);
  int compute_direction(Objective* obj, LineSearch* ls);
  void qn_after_step();
};

// vector kernels (vec_kernels.cu)
void vec_axpy_project(Ctx*, int64_t n, const double* x, const double* d, double t, bool project, const double* lb,
                      const double* ub, double* out, double* d_dn /*nullable*/);
void vec_dot(Ctx*, int64_t n, const double* a, const double* b, double* d_out);
void vec_project_inplace(Ctx*, int64_t n, double* x, const double* lb, const double* ub);
void vec_neg(Ctx*, int64_t n, const double* a, double* out, const double* g, double* d_gd0);
// d = P(x - lam*w) - x ; gd0 = g.d ; tmaxc = min feasible step (only if lb given) ; dinf = ||d||_inf
void vec_projected_direction(Ctx*, int64_t n, const double* x, const double* w, double lam, bool scale, const double* lb,
                             const double* ub, const double* g, double* d, double* d_gd0_tmaxc_dinf);
void vec_tmax_candidate(Ctx*, int64_t n, const double* x, const double* d, const double* lb, const double* ub,
                        double* d_out);
void vec_conv_gnorm2(Ctx*, int64_t n, const double* g, double* d_out);
void vec_conv_gmax(Ctx*, int64_t n, const double* g, double* d_out);
void vec_conv_pginf(Ctx*, int64_t n, const double* x, const double* g, const double* lb, const double* ub, double* d_out);
// s = xn - x ; y = gn - g ; out4 = {s.s, y.y, y.s, 0}
void vec_sy(Ctx*, int64_t n, const double* xn, const double* x, const double* gn, const double* g, double* s, double* y,
            double* d_out4);
void vec_active_set(Ctx*, int64_t n, const double* x, const double* lb, const double* ub, uint8_t* out);
// device-side: s_norm = sqrt(ss), y_norm = sqrt(yy), skip flag, has_s/has_y
void state_finish_sy(Ctx*, DevState* st, double tol);

// device-resident engine (qn_device.cu)
bool qn_device_head_is_cluster(int functor_kind, int64_t n, int head_variant);
void qn_launch_epilogue_cluster(Ctx* ctx, const HeadEpi& e, int64_t n, DevState* st, const double* s, const double* y, const double* g,
                                bool force = false);
void qn_device_launch_head(Ctx* ctx, int functor_kind, const double* fn_a, const double* fn_b, bool bounded,
                           LSParams* d_ls, int64_t n, double tol, int64_t max_ls, DevState* st, double* x, double* g,
                           double* d, double* xt, double* gt, double* s, double* y, const double* u, const double* lb,
                           const double* ub, const double* ls_lb, const double* ls_ub, int head_variant, int ls_kind,
                           const HeadEpi& epi);

// batched (batched.cu)
int batched_bfgs_rosenbrock(Ctx* ctx, int64_t n, int64_t np, const double* x0_host, bool generated, int64_t problem0,
                            double tol, int64_t max_iter, int64_t max_ls, double c1, double beta, double* x_out,
                            double* f_out, int32_t* k_out, int32_t* st_out, int32_t* reason_out, double* ms_out);

void* pool_get(int kind, size_t bytes);            // kind 0 = device, 1 = pinned host
void pool_put(int kind, size_t bytes, void* p);
void pool_trim(int device);  // return every pooled buffer of the device to the driver
void set_last_error(const std::string& s);
// Tracer events (tracer.rs): level 1 error .. 5 trace, the reference's targets and messages
void log_event(int level, const char* target, const std::string& message);
void set_log_callback(osb_log_fn fn, void* user);

}  // namespace osb
