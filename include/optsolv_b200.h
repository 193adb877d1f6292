/* optsolv_b200.h — C ABI of the B200-native (sm_100a) line-search-solver library.
 *
 * This is the drop-in boundary for the hot path of fedemagnani/optimization-solvers
 * (reference paths below are relative to the crate root).  The crate has no FFI for this path
 * today (it is pure Rust, single-threaded, nalgebra f64); these entry points are what a `gpu`
 * backend module of the crate binds with an `extern "C"` block (see INTEGRATION.md and
 * rust/src/gpu/ffi.rs).  Plain pointers, sizes and opaque handles only — no torch/C++ types.
 *
 * Conventions
 *   - every function returns an `int` status unless stated otherwise:
 *       0 OK; 1..4 mirror `SolverError` (src/ls_solver.rs:10-20);
 *       >= 100 are errors the reference expresses as panics or that are new to a device backend.
 *   - all host buffers are caller-owned; all device memory is library-owned.
 *   - dense matrices crossing the boundary are ROW-major n*n doubles.
 *   - a context is single-host-thread-affine; calls are synchronous at `osb_minimize` granularity.
 *   - there is NO CPU fallback: every entry point that computes needs a CUDA device.
 */
#ifndef OPTSOLV_B200_H
#define OPTSOLV_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes -------------------------------------------------------------------------- */
enum {
  OSB_OK = 0,
  OSB_MAX_ITER_REACHED = 1,     /* SolverError::MaxIterReached      src/ls_solver.rs:12-13 */
  OSB_OUT_OF_DOMAIN = 2,        /* SolverError::OutOfDomain         src/ls_solver.rs:14-15 */
  OSB_ERROR_INPUT_PARAMS = 3,   /* SolverError::ErrorInputParams    src/ls_solver.rs:16-17 */
  OSB_ABNORMAL_TERMINATION = 4, /* SolverError::AbnormalTermination src/ls_solver.rs:18-19 */
  OSB_PANIC_NO_HESSIAN = 101,   /* `.expect("Hessian not available in the oracle")` src/newton/mod.rs:34 */
  OSB_PANIC_NOT_SPD = 102,      /* `.cholesky().unwrap()`  src/newton/projected_newton.rs:75, spn.rs:86 */
  OSB_ERR_CUDA = 110,
  OSB_ERR_NCCL = 111,
  OSB_ERR_ALLOC = 112,
  OSB_ERR_UNSUPPORTED = 113,
  OSB_ERR_BAD_HANDLE = 114
};

/* why `minimize` returned Ok(()): the reference only reveals it through warn!/info! logs
 * (src/quasi_newton/bfgs.rs:68,71; src/ls_solver.rs:82-86) */
enum {
  OSB_REASON_NONE = 0,
  OSB_REASON_GRAD_TOL = 1,         /* ||g||_2 < tol (bfgs.rs:74) or max|g_i| < tol (gradient_descent.rs:46-53) */
  OSB_REASON_S_NORM = 2,           /* "next iterate too close"           bfgs.rs:67-69 */
  OSB_REASON_Y_NORM = 3,           /* "gradient next iterate too close"  bfgs.rs:70-72 */
  OSB_REASON_PROJ_GRAD_TOL = 4,    /* projected-gradient inf-norm        projected_gradient_descent.rs:76-83 */
  OSB_REASON_NEWTON_DECREMENT = 5  /* lambda^2/2 < tol                   newton/mod.rs:64-69 */
};

/* solver kinds: one per struct implementing LineSearchSolver on the hot path */
enum {
  OSB_GD = 0,           /* GradientDescent            src/steepest_descent/gradient_descent.rs */
  OSB_PGD = 1,          /* ProjectedGradientDescent   src/steepest_descent/projected_gradient_descent.rs */
  OSB_SPG = 2,          /* SpectralProjectedGradient  src/steepest_descent/spg.rs */
  OSB_BFGS = 3,         /* src/quasi_newton/bfgs.rs */
  OSB_DFP = 4,          /* src/quasi_newton/dfp.rs */
  OSB_BROYDEN = 5,      /* src/quasi_newton/broyden.rs */
  OSB_BFGSB = 6,        /* src/quasi_newton/bfgs_b.rs */
  OSB_DFPB = 7,         /* src/quasi_newton/dfp_b.rs */
  OSB_BROYDENB = 8,     /* src/quasi_newton/broyden_b.rs */
  OSB_SR1B = 9,         /* src/quasi_newton/sr1_b.rs */
  OSB_NEWTON = 10,      /* src/newton/mod.rs */
  OSB_PROJ_NEWTON = 11, /* src/newton/projected_newton.rs */
  OSB_SPN = 12,         /* src/newton/spn.rs */
  OSB_PNORM = 13        /* PnormDescent               src/steepest_descent/pnorm_descent.rs: direction
                           -(inverse_p g); the matrix is set with osb_solver_set_inv_hessian (identity until then) */
};

typedef struct osb_ctx osb_ctx;
typedef struct osb_objective osb_objective;
typedef struct osb_linesearch osb_linesearch;
typedef struct osb_solver osb_solver;

/* thread-local description of the last non-zero status returned on this thread */
const char* osb_last_error_string(void);
/* library / build identification, e.g. "optsolv_b200 0.1 sm_100a" */
const char* osb_version(void);

/* ---- Tracer (src/tracer.rs:5-63): the library emits the reference's own events — same targets, levels and messages —
 * through ONE callback, on the thread that called osb_minimize; the host side routes them into its tracing subscriber
 * (Rust: the crate's `Tracer` unchanged; Python mirror: `Tracer.build()`).  Events: "solver" error "Minimization
 * completed: next iterate is out of domain" (ls_solver.rs:38), "solver" info "Minimization completed: convergence in {k}
 * iterations" (:82-86), "solver" warn "Minimization completed: max iter reached during minimization" (:109), "bfgs" warn
 * "Minimization completed: next iterate too close" / "... gradient next iterate too close" (bfgs.rs:68,71 and siblings),
 * "newton" warn "Hessian is singular. Using gradient descent direction." (newton/mod.rs:44).  level: 1 error, 2 warn,
 * 3 info, 4 debug, 5 trace.  NULL removes the callback (default: no events). */
typedef void (*osb_log_fn)(void* user, int level, const char* target, const char* message);
int osb_set_log_callback(osb_log_fn fn, void* user);

/* ---- context: one per GPU (one process per GPU) -------------------------------------------- */
int osb_ctx_create(int device, osb_ctx** out);
/* Row-block sharded multi-GPU context.  `nccl_unique_id` is the 128-byte ncclUniqueId produced
 * by osb_nccl_unique_id() on rank 0 and broadcast by the caller (torch.distributed, MPI, ...). */
int osb_nccl_unique_id(void* out128);
int osb_ctx_create_dist(int device, int rank, int world, const void* nccl_unique_id, osb_ctx** out);
/* Peer-memory exchange (optional, world > 1): every rank exports the 64-byte CUDA IPC handle of its exchange
 * region, the caller all-gathers the handles (rank order, 64 bytes each) and every rank connects.  The
 * lazy quasi-Newton pass then all-gathers its row sums with NVLink peer stores fused into the kernel
 * instead of NCCL calls. */
int osb_ctx_ipc_handle(osb_ctx* ctx, void* out64);
int osb_ctx_ipc_connect(osb_ctx* ctx, const void* handles_world_x_64);
/* Drops the peer mappings and the exchange region again (p2p off): EVERY rank must call it when the connect did not
 * succeed on every rank, otherwise the connected ranks would wait in the fused exchange for a rank that sits in
 * NCCL.  (The in-kernel waits are bounded all the same: ~5 s without progress end the solve with
 * OSB_ABNORMAL_TERMINATION.) */
int osb_ctx_ipc_close(osb_ctx* ctx);
/* Index-range sharding of the O(n) solvers (world > 1): when on, GradientDescent / ProjectedGradientDescent /
 * SpectralProjectedGradient created on this context own a contiguous SLICE of the variables (their n, x0, lb, ub are
 * the local ones; objectives must be block-functor objectives created for the same slice), and every scalar of
 * the path (f, g.d, s.y, s.s, the inf-norms of number.rs:27-31 and projected_gradient_descent.rs:76-83) is
 * combined across ranks in rank order.  Replaces nothing in the reference (single-threaded); SURVEY 8e. */
int osb_ctx_set_vector_sharding(osb_ctx* ctx, int on);
/* Host-only description of the packed symmetric layouts (qn_storage = 1; no GPU needed): the 8-row tile `tile` of an
 * n x n matrix is owned by rank *owner and starts *offset_doubles into that rank's packed array, rows *row_stride_doubles
 * apart; *rank_total_doubles is the size of rank `rank`'s array.  world = 1: all tiles consecutive.  world > 1: tile
 * pairs (p, T-1-p) dealt round-robin over the ranks, see DESIGN.md section 6. */
int osb_sym_layout(int64_t n, int world, int rank, int64_t tile, int* owner, int64_t* offset_doubles, int64_t* row_stride_doubles,
                   int64_t* rank_total_doubles);
/* The n x n matrices of destroyed solvers are kept in a small per-device pool (at most 6 buffers / 6 GiB) and reused by
 * the next solver of the same size; this returns them to the driver (context destruction does it too). */
int osb_ctx_trim_memory(osb_ctx* ctx);
void osb_ctx_destroy(osb_ctx* ctx);
int osb_ctx_rank(const osb_ctx* ctx);
int osb_ctx_world(const osb_ctx* ctx);
int osb_ctx_synchronize(osb_ctx* ctx);
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on */
void* osb_ctx_stream(osb_ctx* ctx);
/* counters since context creation: [0] kernel launches, [1] objective evaluations,
 * [2] line-search trials, [3] host<->device synchronisations, [4] collectives,
 * [5] passes over the sharded packed triangle (multi-GPU, qn_storage = 1) */
int osb_ctx_counters(const osb_ctx* ctx, int64_t out[8]);

/* ---- objectives: the device counterpart of `FnMut(&DVector<f64>) -> FuncEvalMultivariate` ---
 * (src/ls_solver.rs:34, src/func_eval.rs:5-41).  f, g and (optionally) the Hessian stay on the
 * device; `osb_objective_eval` is the host-visible probe used by tests and examples. */

/* f = x.(A x) [- 2 b.x], g = 2 A x [- 2 b] — the oracle pattern of examples/quadratic.rs:10-14.
 * A is row-major n*n (symmetric use), b may be NULL. */
int osb_objective_create_dense_quadratic(osb_ctx* ctx, int64_t n, const double* A_host, const double* b_host,
                                         osb_objective** out);
/* synthetic SPD quadratic generated on device from the integer hash of DESIGN.md §inputs;
 * x0_host (n doubles, may be NULL) receives the matching start point. */
int osb_objective_create_dense_quadratic_generated(osb_ctx* ctx, int64_t n, int shifted, double* x0_host,
                                                   osb_objective** out);
/* extended Rosenbrock, n even */
int osb_objective_create_rosenbrock(osb_ctx* ctx, int64_t n, osb_objective** out);
/* separable f = sum 0.5 c_i (x_i - a_i)^2 with hash-generated c, a (box-constrained SPG config) */
int osb_objective_create_separable_quadratic_generated(osb_ctx* ctx, int64_t n, osb_objective** out);
/* the slice [index0, index0 + n_local) of the same generated problem (index-range sharding) */
int osb_objective_create_separable_quadratic_generated_shard(osb_ctx* ctx, int64_t n_local, int64_t index0, osb_objective** out);
/* synthetic l2-regularised logistic regression, m samples x n features generated on device */
int osb_objective_create_logistic_generated(osb_ctx* ctx, int64_t m, int64_t n, double lambda, osb_objective** out);

/* A host closure (compatibility path: H2D of x and D2H of f,g,[H] per call).  The callback fills
 * *f, g[n] and, when `hess` is non-NULL and it has a Hessian, hess[n*n] row-major, returning 1 if
 * it wrote the Hessian and 0 otherwise. */
typedef int (*osb_host_eval_fn)(void* user, const double* x, int64_t n, double* f, double* g, double* hess);
int osb_objective_create_host(osb_ctx* ctx, int64_t n, osb_host_eval_fn fn, void* user, int with_hessian,
                              osb_objective** out);

/* A user-supplied DEVICE functor: the callback must enqueue, on `stream` (a cudaStream_t), work
 * that reads d_x[n] and writes *d_f, d_g[n] and, if d_hess != NULL, the Hessian (device pointers).
 * The Hessian buffer is row-major with leading dimension osb_hessian_ld(n) (n rounded up to 16
 * doubles; the padding columns must be left zero).  It must not synchronise.  Returns 0 on success. */
int64_t osb_hessian_ld(int64_t n);
typedef int (*osb_device_eval_fn)(void* user, const double* d_x, int64_t n, double* d_f, double* d_g,
                                  double* d_hess, void* stream);
int osb_objective_create_user(osb_ctx* ctx, int64_t n, osb_device_eval_fn fn, void* user, int with_hessian,
                              osb_objective** out);

int osb_objective_eval(osb_objective* obj, const double* x_host, double* f, double* g_host, double* hess_host);
int64_t osb_objective_calls(const osb_objective* obj);
int64_t osb_objective_dim(const osb_objective* obj);
void osb_objective_destroy(osb_objective* obj);

/* ---- line searches (src/line_search/) ------------------------------------------------------ */
int osb_linesearch_create_backtracking(double c1, double beta, osb_linesearch** out);           /* backtracking.rs:8-10 */
int osb_linesearch_create_backtracking_b(osb_ctx* ctx, double c1, double beta, int64_t n, const double* lb_host,
                                         const double* ub_host, osb_linesearch** out);          /* backtracking_b.rs:11-23 */
int osb_linesearch_create_morethuente(double c1, double c2, double t_min, double t_max, double delta_min,
                                      double delta, double delta_max, osb_linesearch** out);    /* morethuente.rs:16-62 */
int osb_linesearch_create_morethuente_b(osb_ctx* ctx, double c1, double c2, double t_min, double t_max,
                                        double delta_min, double delta, double delta_max, int64_t n,
                                        const double* lb_host, const double* ub_host,
                                        osb_linesearch** out);                                  /* morethuente_b.rs:17-40 */
int osb_linesearch_create_gll_quadratic(double c1, int64_t m, double sigma1, double sigma2,
                                        osb_linesearch** out);                                  /* gll_quadratic.rs:13-28 */
int osb_linesearch_create_nosearch(osb_linesearch** out);                                       /* nosearch.rs:3 */
/* current t_max (MoreThuenteB shrinks it permanently, morethuente_b.rs:201); NaN for other kinds */
double osb_linesearch_t_max(const osb_linesearch* ls);
void osb_linesearch_destroy(osb_linesearch* ls);
/* LineSearch::compute_step_len (line_search/mod.rs:14-23) as a standalone call: evaluates the
 * objective at x itself, then searches along d.  Used by tests mirroring backtracking.rs:63-113. */
int osb_linesearch_compute_step_len(osb_ctx* ctx, osb_linesearch* ls, osb_objective* obj, const double* x_host,
                                    const double* d_host, int64_t max_iter, double* t_out);

/* Host-only run of the line-search scalar automaton on a 1-D model phi(t) = f(x + t d): the
 * callback returns f, g.d and ||P(x + t d) - x||^2 for the requested step (`projected` is 1 when
 * the search evaluates at the projected trial, backtracking_b.rs:65-67).  No GPU is involved;
 * this is the same automaton code the device engines execute (used by CPU tests and by callers
 * that bring their own evaluation).  `*last_eval_is_result` tells whether the last evaluated
 * step is the returned one. */
typedef void (*osb_phi_fn)(void* user, double t, int projected, double* f, double* gd, double* dn);
int osb_linesearch_step_len_scalar(osb_linesearch* ls, osb_phi_fn phi, void* user, double f0, double gd0,
                                   double tmax_candidate, int64_t max_iter, double* t_out, int* last_eval_is_result);

/* ---- solvers ------------------------------------------------------------------------------- */
/* `X::new(tol, x0[, lb, ub])`.  lb/ub are NULL for unbounded kinds.  `objective_for_lambda0` is
 * only used by OSB_SPG / OSB_SPN, whose constructors call the oracle once (spg.rs:28-46). */
int osb_solver_create(osb_ctx* ctx, int kind, int64_t n, double tol, const double* x0_host, const double* lb_host,
                      const double* ub_host, osb_objective* objective_for_lambda0, osb_solver** out);
void osb_solver_destroy(osb_solver* s);

/* LineSearchSolver::minimize (src/ls_solver.rs:66-111).  `callback` (may be NULL) is invoked on
 * the calling thread after k += 1, like `Option<&mut dyn FnMut(&Self)>`; a non-NULL callback
 * forces one host synchronisation per outer iteration. */
typedef void (*osb_callback_fn)(void* user, osb_solver* s);
int osb_minimize(osb_solver* s, osb_linesearch* ls, osb_objective* obj, int64_t max_iter_solver,
                 int64_t max_iter_line_search, osb_callback_fn callback, void* user);

/* options (set before minimize; every one has a default that needs no call):
 *   "engine"             0 = auto (default: device-resident control whenever solver, line search and objective allow it),
 *                        1 = host-driven control only, 2 = device-resident only (OSB_ERR_UNSUPPORTED if it does not apply)
 *   "qn_schedule"       -1 = auto (default: lazy for BFGS / DFP with n > 5, eager otherwise),
 *                        0 = eager (h = H y, then fused update: 3 n^2 8 B per iteration),
 *                        1 = lazy (one read-modify-write per iteration, the update of iteration k is applied by the pass of
 *                            iteration k + 1: 2 n^2 8 B; BFGS / DFP)
 *   "qn_storage"        -1 = auto (default: packed whenever the lazy schedule runs), 0 = full n x n row-major,
 *                        1 = packed lower triangle in 8-row tiles (n^2 8 B per iteration; on several GPUs sharded by tile
 *                            pairs, which needs the peer-memory exchange)
 *   "callback_run_ahead" -1 / 1 = (default) the callback of iteration k is delivered from pinned snapshots of x, g, f, k,
 *                        s_norm, y_norm while the device already runs iteration k + 1; any other getter inside such a
 *                        callback returns OSB_ERR_UNSUPPORTED.  0 = the device waits for the callback (every getter works)
 *   "record_trace"       1 = keep per-iteration (f, t, s_norm, y_norm) for osb_solver_trace
 *   "use_p2p"            multi-GPU: 1 = (default) fused peer-memory exchange when the context is IPC-connected, 0 = NCCL
 *   "head_kernel"        device engine head: 0 = 8-CTA cluster (default), 1 = single CTA + shared memory, 2 = generic
 *   "qn_kernel"          kernel variant of the H pass (diagnostics, default 0).  Full storage: 1 = TMA-staged ring.  Packed
 *                        storage: bit 0 = two 256-thread CTAs per SM, bit 1 = ping-pong (out-of-place) storage,
 *                        bit 2 = zero-first column partials (round-1 behaviour), bit 3 = shared-memory ring fed by
 *                        cp.async.bulk (bit-identical to 0; measured 5 % slower, profiles/r02_packed_pass_experiments.md)
 *   "fused_iteration"    whole outer iterations in ONE cooperative kernel (lazy schedule on packed storage under
 *                        device-resident control: line search on every SM, H pass, fold and multi-GPU exchange separated by
 *                        grid barriers only): -1 = auto (default: on with several GPUs, off with one), 1 = on, 0 = one launch
 *                        per phase
 *   "fused_stream"       ProjectedGradientDescent / SpectralProjectedGradient on a block-functor objective: -1 / 1 = (default)
 *                        ONE kernel per line-search trial (direction, projection, objective, all dot products and the
 *                        projected-gradient norm: 4 vector reads + 2 writes), 0 = one launch per vector expression
 *   "profile_kernels"    1 = one launch per phase, the H pass(es) bracketed with CUDA events (osb_solver_kernel_timing)
 *   "profile_iter"       1 = the fused kernel records where its time goes (osb_solver_iter_profile) */
int osb_solver_set_option(osb_solver* s, const char* name, int64_t value);
int osb_solver_set_lambdas(osb_solver* s, double lambda_min, double lambda_max); /* spg.rs:23-27 */

/* getters (derive_getters accessors of the reference structs) */
int64_t osb_solver_k(const osb_solver* s);
int64_t osb_solver_dim(const osb_solver* s);
int osb_solver_termination_reason(const osb_solver* s);
int osb_solver_x(osb_solver* s, double* out_host);            /* x() / xk() */
int osb_solver_set_x(osb_solver* s, const double* x_host);    /* xk_mut() */
int osb_solver_f(osb_solver* s, double* f_out);               /* f at the last evaluate_x_k */
int osb_solver_grad(osb_solver* s, double* out_host);         /* g at the last evaluate_x_k */
double osb_solver_s_norm(const osb_solver* s);                /* NaN when None */
double osb_solver_y_norm(const osb_solver* s);                /* NaN when None */
int osb_solver_clear_norms(osb_solver* s);                    /* s_norm = y_norm = None */
double osb_solver_lambda(const osb_solver* s);                /* SPG / SPN */
double osb_solver_decrement_squared(const osb_solver* s);     /* Newton; NaN when None */
int osb_solver_inv_hessian(osb_solver* s, double* out_host);  /* approx_inv_hessian(), row-major n*n */
/* also PnormDescent::new's inverse_p (pnorm_descent.rs:23-30) / its getter inverse_p() */
int osb_solver_set_inv_hessian(osb_solver* s, const double* in_host);
/* one byte per coordinate: bit0 = (x_i == lb_i), bit1 = (x_i == ub_i) — exact compares, the
 * active-set definition of HasProjectedGradient::projected_gradient (src/ls_solver.rs:121-133) */
int osb_solver_active_set(osb_solver* s, uint8_t* out_host);
int64_t osb_solver_trace_len(const osb_solver* s);
int osb_solver_trace(const osb_solver* s, double* f, double* t, double* s_norm, double* y_norm);
/* with option "profile_kernels" = 1: mean device ms (CUDA-event pairs on the launching stream) of
 * out[0] pass 1 (h = H y), out[1] pass 2 (fused update) over the last minimize; out[2] = #iterations timed */
int osb_solver_kernel_timing(const osb_solver* s, double out[3]);
/* which path the last minimize() took (the defaults are "auto"): out[0] engine (1 host-driven, 2 device-resident control),
 * out[1] schedule in force (0 eager, 1 lazy), out[2] storage in force (0 full n x n, 1 packed lower triangle), out[3] packed
 * triangle sharded over the ranks, out[4] fused peer-memory exchange used, out[5] ranks, out[6] kernel variant,
 * out[7] bit 0: whole iterations ran in the fused cooperative kernel, bit 1: PGD / SPG ran one fused kernel per trial */
int osb_solver_path_info(const osb_solver* s, int64_t out[8]);
/* option "profile_iter": mean ms per iteration in out[0] head (epilogue, line search, next iterate), out[1] H pass,
 * out[2] fold + exchange of the fused iteration kernel (globaltimer stamps of CTA 0); out[3] = iterations covered;
 * out[4..14] = sub-phases of the head (ms per iteration): 4 epilogue loads, 5 its grid sum, 6 u + direction, 7 trial steps,
 * 8 their grid sum, 9 line-search automaton, 10 next iterate, 11 its grid sum */
int osb_solver_iter_profile(osb_solver* s, double out[16]);
/* device time (ms, CUDA events on the context stream) and outer iterations of the last minimize */
int osb_solver_last_timing(const osb_solver* s, double* ms, int64_t* iterations);

/* ---- batched mode: one small problem per warp/CTA, many independent problems ---------------
 * BFGS (bfgs.rs) + BackTracking (backtracking.rs) on extended Rosenbrock, all problems resident
 * on one GPU; shard `n_problems` across GPUs on the caller side (no collective).
 * x0_host / x_out_host are n_problems*n row-major; k_out, status_out, reason_out have n_problems
 * entries.  ms_out (may be NULL) receives the device time of the solve kernel. */
int osb_batched_bfgs_rosenbrock(osb_ctx* ctx, int64_t n, int64_t n_problems, const double* x0_host, double tol,
                                int64_t max_iter_solver, int64_t max_iter_line_search, double c1, double beta,
                                double* x_out_host, double* f_out_host, int32_t* k_out, int32_t* status_out,
                                int32_t* reason_out, double* ms_out);
/* same, with x0 generated on device: x0 = (-1.2, 1, ...) + int16(hash(3, problem0 + p, i)) * 2^-16 */
int osb_batched_bfgs_rosenbrock_generated(osb_ctx* ctx, int64_t n, int64_t n_problems, int64_t problem0, double tol,
                                          int64_t max_iter_solver, int64_t max_iter_line_search, double c1,
                                          double beta, double* x_out_host, double* f_out_host, int32_t* k_out,
                                          int32_t* status_out, int32_t* reason_out, double* ms_out);

/* ---- micro-benchmark hooks for bench.py / ncu (device-resident inputs, no host traffic) ---- */
/* runs `reps` back-to-back launches of one hot kernel on an n*n H and returns the mean ms/launch:
 *   which = 0: h = H y (read n^2)   1: fused rank-2 update + u = H' g (read + write n^2)
 *           2: Broyden H^T s        3: plain device copy of H (cudaMemcpyAsync D2D) for calibration
 *           4: the lazy pass over the packed lower triangle alone (read + write n^2/2), variant = option "qn_kernel" */
int osb_bench_qn_kernel(osb_ctx* ctx, int which, int64_t n, int reps, int variant, double* ms_out);
/* mean microseconds of one grid barrier of the fused iteration kernel's shape (one 512-thread CTA per SM, cooperative
 * launch): the fixed cost the kernel pays four times per iteration */
int osb_bench_grid_sync(osb_ctx* ctx, int reps, double* us_out);
/* mean ms per launch of the DMMA Hessian assembly X^T D X of a logistic-regression objective (m n^2 MACs on the lower triangle) */
int osb_bench_syrk(osb_ctx* ctx, osb_objective* logistic, int reps, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* OPTSOLV_B200_H */
