// api.cu — the extern "C" boundary declared in include/optsolv_b200.h.
#include <cstring>

#include "engine.cuh"

namespace osb {
const std::string& get_last_error();
void nccl_unique_id(void* out128);
void ctx_init_dist(Ctx* ctx, int rank, int world, const void* uid);
void ctx_destroy_dist(Ctx* ctx);
void ctx_ipc_export(Ctx* ctx, void* out64);
void ctx_ipc_connect(Ctx* ctx, const void* handles);
void ctx_ipc_close(Ctx* ctx);
}  // namespace osb

using namespace osb;
namespace osb { double bench_syrk_dmma(Ctx* ctx, Objective* obj, int reps); double bench_grid_sync(Ctx* ctx, int reps); }

#define OSB_TRY try {
#define OSB_CATCH                                  \
  }                                                \
  catch (const osb::Error& e) {                    \
    osb::set_last_error(e.what());                 \
    return e.code;                                 \
  }                                                \
  catch (const std::exception& e) {                \
    osb::set_last_error(e.what());                 \
    return OSB_ABNORMAL_TERMINATION;               \
  }

static Ctx* C(osb_ctx* c) { return reinterpret_cast<Ctx*>(c); }
static const Ctx* C(const osb_ctx* c) { return reinterpret_cast<const Ctx*>(c); }
static Objective* O(osb_objective* o) { return reinterpret_cast<Objective*>(o); }
static LineSearch* L(osb_linesearch* l) { return reinterpret_cast<LineSearch*>(l); }
static Solver* S(osb_solver* s) { return reinterpret_cast<Solver*>(s); }
static const Solver* S(const osb_solver* s) { return reinterpret_cast<const Solver*>(s); }

extern "C" {

const char* osb_last_error_string(void) { return osb::get_last_error().c_str(); }
const char* osb_version(void) { return "optsolv_b200 0.2 sm_100a"; }
int osb_set_log_callback(osb_log_fn fn, void* user) {
  osb::set_log_callback(fn, user);
  return OSB_OK;
}

int osb_ctx_create(int device, osb_ctx** out) {
  OSB_TRY
  *out = reinterpret_cast<osb_ctx*>(new Ctx(device));
  return OSB_OK;
  OSB_CATCH
}
int osb_nccl_unique_id(void* out128) {
  OSB_TRY
  nccl_unique_id(out128);
  return OSB_OK;
  OSB_CATCH
}
int osb_ctx_create_dist(int device, int rank, int world, const void* uid, osb_ctx** out) {
  OSB_TRY
  OSB_REQUIRE(world >= 1 && rank >= 0 && rank < world, OSB_ERROR_INPUT_PARAMS, "bad rank/world");
  Ctx* c = new Ctx(device);
  if (world > 1) {
    try {
      ctx_init_dist(c, rank, world, uid);
    } catch (...) {
      delete c;
      throw;
    }
  }
  *out = reinterpret_cast<osb_ctx*>(c);
  return OSB_OK;
  OSB_CATCH
}
int osb_ctx_ipc_handle(osb_ctx* ctx, void* out64) {
  OSB_TRY
  ctx_ipc_export(C(ctx), out64);
  return OSB_OK;
  OSB_CATCH
}
int osb_ctx_ipc_connect(osb_ctx* ctx, const void* handles) {
  OSB_TRY
  ctx_ipc_connect(C(ctx), handles);
  return OSB_OK;
  OSB_CATCH
}
int osb_ctx_ipc_close(osb_ctx* ctx) {
  OSB_TRY
  C(ctx)->use();
  C(ctx)->sync();
  ctx_ipc_close(C(ctx));
  return OSB_OK;
  OSB_CATCH
}
void osb_ctx_destroy(osb_ctx* ctx) {
  if (!ctx) return;
  try {
    ctx_ipc_close(C(ctx));
    ctx_destroy_dist(C(ctx));
  } catch (...) {
  }
  delete C(ctx);
}
int osb_ctx_rank(const osb_ctx* ctx) { return C(ctx)->rank; }
int osb_ctx_world(const osb_ctx* ctx) { return C(ctx)->world; }
int osb_ctx_synchronize(osb_ctx* ctx) {
  OSB_TRY
  C(ctx)->use();
  C(ctx)->sync();
  return OSB_OK;
  OSB_CATCH
}
int osb_ctx_set_vector_sharding(osb_ctx* ctx, int on) {
  OSB_TRY
  Ctx* c = C(ctx);
  c->use();
  OSB_REQUIRE(!on || c->world > 1, OSB_ERROR_INPUT_PARAMS, "index-range sharding needs a distributed context (osb_ctx_create_dist)");
  if (on && !c->shard_scratch) {
    OSB_CUDA(cudaMalloc(&c->shard_scratch, sizeof(double) * 8 * (size_t)c->world));
    OSB_CUDA(cudaMemset(c->shard_scratch, 0, sizeof(double) * 8 * (size_t)c->world));
  }
  c->vec_sharded = on != 0;
  return OSB_OK;
  OSB_CATCH
}
int osb_sym_layout(int64_t n, int world, int rank, int64_t tile, int* owner, int64_t* offset_doubles, int64_t* row_stride_doubles,
                   int64_t* rank_total_doubles) {
  OSB_TRY
  const int64_t T = (n + 7) / 8;
  OSB_REQUIRE(n >= 1 && world >= 1 && rank >= 0 && rank < world && tile >= 0 && tile < T, OSB_ERROR_INPUT_PARAMS, "bad layout query");
  OSB_REQUIRE(world == 1 || (n % 16 == 0 && n / 16 >= world), OSB_ERROR_INPUT_PARAMS,
              "the sharded packed layout needs n divisible by 16 and at least one tile pair per rank");
  qn_sym_layout(n, world, tile, owner, offset_doubles, row_stride_doubles);
  *rank_total_doubles = world == 1 ? qn_sym_doubles(n) : qn_sym_doubles_sharded(n, world, rank);
  return OSB_OK;
  OSB_CATCH
}
int osb_ctx_trim_memory(osb_ctx* ctx) {
  OSB_TRY
  C(ctx)->use();
  C(ctx)->sync();
  pool_trim(C(ctx)->device);
  return OSB_OK;
  OSB_CATCH
}
void* osb_ctx_stream(osb_ctx* ctx) { return (void*)C(ctx)->stream; }
int osb_ctx_counters(const osb_ctx* ctx, int64_t out[8]) {
  std::memcpy(out, C(ctx)->counters, sizeof(int64_t) * 8);
  return OSB_OK;
}

// ---- objectives
#define MAKE_OBJ(expr)                                   \
  OSB_TRY                                                \
  C(ctx)->use();                                         \
  *out = reinterpret_cast<osb_objective*>(expr);         \
  return OSB_OK;                                         \
  OSB_CATCH

int osb_objective_create_dense_quadratic(osb_ctx* ctx, int64_t n, const double* A, const double* b, osb_objective** out) {
  MAKE_OBJ(make_dense_quadratic(C(ctx), n, A, b))
}
int osb_objective_create_dense_quadratic_generated(osb_ctx* ctx, int64_t n, int shifted, double* x0, osb_objective** out) {
  MAKE_OBJ(make_dense_quadratic_generated(C(ctx), n, shifted != 0, x0))
}
int osb_objective_create_rosenbrock(osb_ctx* ctx, int64_t n, osb_objective** out) { MAKE_OBJ(make_rosenbrock(C(ctx), n)) }
int osb_objective_create_separable_quadratic_generated(osb_ctx* ctx, int64_t n, osb_objective** out) {
  MAKE_OBJ(make_sepquad_generated(C(ctx), n))
}
int osb_objective_create_separable_quadratic_generated_shard(osb_ctx* ctx, int64_t n_local, int64_t index0, osb_objective** out) {
  MAKE_OBJ(make_sepquad_generated(C(ctx), n_local, index0))
}
int osb_objective_create_logistic_generated(osb_ctx* ctx, int64_t m, int64_t n, double lambda, osb_objective** out) {
  MAKE_OBJ(make_logistic_generated(C(ctx), m, n, lambda))
}
int osb_objective_create_host(osb_ctx* ctx, int64_t n, osb_host_eval_fn fn, void* user, int with_h, osb_objective** out) {
  MAKE_OBJ(make_host_objective(C(ctx), n, fn, user, with_h != 0))
}
int64_t osb_hessian_ld(int64_t n) { return qn_ld(n); }
int osb_objective_create_user(osb_ctx* ctx, int64_t n, osb_device_eval_fn fn, void* user, int with_h, osb_objective** out) {
  MAKE_OBJ(make_user_objective(C(ctx), n, fn, user, with_h != 0))
}
int osb_objective_eval(osb_objective* obj, const double* x_host, double* f, double* g_host, double* hess_host) {
  OSB_TRY
  Objective* o = O(obj);
  Ctx* ctx = o->ctx;
  ctx->use();
  const int64_t n = o->n, ld = qn_ld(n);
  DBuf x(ld), g(ld), fbuf(8), hess;
  x.zero(ctx->stream);
  g.zero(ctx->stream);
  x.upload(x_host, n, ctx->stream);
  const bool want_h = hess_host != nullptr && o->provides_hessian();
  if (want_h) {
    hess.alloc(qn_rows_padded(n) * ld);
    hess.zero(ctx->stream);
  }
  o->eval(x.p, fbuf.p, g.p, want_h ? hess.p : nullptr);
  fbuf.download(f, 1, ctx->stream);
  if (g_host) g.download(g_host, n, ctx->stream);
  if (want_h)
    OSB_CUDA(cudaMemcpy2DAsync(hess_host, n * sizeof(double), hess.p, ld * sizeof(double), n * sizeof(double), n,
                               cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  return OSB_OK;
  OSB_CATCH
}
int64_t osb_objective_calls(const osb_objective* obj) { return reinterpret_cast<const Objective*>(obj)->calls; }
int64_t osb_objective_dim(const osb_objective* obj) { return reinterpret_cast<const Objective*>(obj)->n; }
void osb_objective_destroy(osb_objective* obj) {
  if (!obj) return;
  cudaSetDevice(O(obj)->ctx->device);
  delete O(obj);
}

// ---- line searches
static int make_ls(int kind, osb_linesearch** out, LineSearch** raw) {
  LineSearch* l = new LineSearch();
  l->p = ls_defaults(kind);
  *out = reinterpret_cast<osb_linesearch*>(l);
  *raw = l;
  return OSB_OK;
}
static void ls_bounds(LineSearch* l, osb_ctx* ctx, int64_t n, const double* lb, const double* ub) {
  if (ctx == nullptr) {  // host-only handle (osb_linesearch_step_len_scalar): no device vectors
    l->n = n;
    return;
  }
  OSB_REQUIRE(lb && ub && n >= 1, OSB_ERROR_INPUT_PARAMS, "bounded line search needs n, lb, ub");
  Ctx* c = C(ctx);
  c->use();
  l->ctx = c;
  l->n = n;
  l->lb.alloc(qn_ld(n));
  l->ub.alloc(qn_ld(n));
  l->lb.zero(c->stream);
  l->ub.zero(c->stream);
  l->lb.upload(lb, n, c->stream);
  l->ub.upload(ub, n, c->stream);
  c->sync();
}
int osb_linesearch_create_backtracking(double c1, double beta, osb_linesearch** out) {
  OSB_TRY
  LineSearch* l;
  make_ls(LS_BACKTRACKING, out, &l);
  l->p.c1 = c1;
  l->p.beta = beta;
  return OSB_OK;
  OSB_CATCH
}
int osb_linesearch_create_backtracking_b(osb_ctx* ctx, double c1, double beta, int64_t n, const double* lb, const double* ub,
                                         osb_linesearch** out) {
  OSB_TRY
  LineSearch* l;
  make_ls(LS_BACKTRACKING_B, out, &l);
  l->p.c1 = c1;
  l->p.beta = beta;
  try {
    ls_bounds(l, ctx, n, lb, ub);
  } catch (...) {
    delete l;
    *out = nullptr;
    throw;
  }
  return OSB_OK;
  OSB_CATCH
}
static void mt_params(LineSearch* l, double c1, double c2, double t_min, double t_max, double dmin, double d, double dmax) {
  l->p.c1 = c1;
  l->p.c2 = c2;
  l->p.t_min = t_min;
  l->p.t_max = t_max;
  l->p.delta_min = dmin;
  l->p.delta = d;
  l->p.delta_max = dmax;
}
int osb_linesearch_create_morethuente(double c1, double c2, double t_min, double t_max, double dmin, double d, double dmax,
                                      osb_linesearch** out) {
  OSB_TRY
  LineSearch* l;
  make_ls(LS_MORETHUENTE, out, &l);
  mt_params(l, c1, c2, t_min, t_max, dmin, d, dmax);
  return OSB_OK;
  OSB_CATCH
}
int osb_linesearch_create_morethuente_b(osb_ctx* ctx, double c1, double c2, double t_min, double t_max, double dmin, double d,
                                        double dmax, int64_t n, const double* lb, const double* ub, osb_linesearch** out) {
  OSB_TRY
  LineSearch* l;
  make_ls(LS_MORETHUENTE_B, out, &l);
  mt_params(l, c1, c2, t_min, t_max, dmin, d, dmax);
  try {
    ls_bounds(l, ctx, n, lb, ub);
  } catch (...) {
    delete l;
    *out = nullptr;
    throw;
  }
  return OSB_OK;
  OSB_CATCH
}
int osb_linesearch_create_gll_quadratic(double c1, int64_t m, double sigma1, double sigma2, osb_linesearch** out) {
  OSB_TRY
  OSB_REQUIRE(m >= 1 && m <= GLL_MAX_M, OSB_ERROR_INPUT_PARAMS, "GLL window m must be in [1, 64]");
  LineSearch* l;
  make_ls(LS_GLL, out, &l);
  l->p.c1 = c1;
  l->p.m = (int)m;
  l->p.sigma1 = sigma1;
  l->p.sigma2 = sigma2;
  return OSB_OK;
  OSB_CATCH
}
int osb_linesearch_create_nosearch(osb_linesearch** out) {
  OSB_TRY
  LineSearch* l;
  make_ls(LS_NOSEARCH, out, &l);
  return OSB_OK;
  OSB_CATCH
}
double osb_linesearch_t_max(const osb_linesearch* ls) {
  const LineSearch* l = reinterpret_cast<const LineSearch*>(ls);
  return (l->p.kind == LS_MORETHUENTE || l->p.kind == LS_MORETHUENTE_B) ? l->p.t_max : NAN;
}
void osb_linesearch_destroy(osb_linesearch* ls) {
  if (!ls) return;
  if (L(ls)->ctx) cudaSetDevice(L(ls)->ctx->device);
  delete L(ls);
}

int osb_linesearch_compute_step_len(osb_ctx* ctxh, osb_linesearch* lsh, osb_objective* objh, const double* x_host,
                                    const double* d_host, int64_t max_iter, double* t_out) {
  OSB_TRY
  Ctx* ctx = C(ctxh);
  ctx->use();
  LineSearch* ls = L(lsh);
  Objective* obj = O(objh);
  const int64_t n = obj->n, ld = qn_ld(n);
  DBuf x(ld), d(ld), g(ld), xt(ld), gt(ld), sc(16);
  for (DBuf* b : {&x, &d, &g, &xt, &gt, &sc}) b->zero(ctx->stream);
  x.upload(x_host, n, ctx->stream);
  d.upload(d_host, n, ctx->stream);
  // sc: [0] f0, [1] gd0, [2] tmaxc, [4..6] trial outputs
  obj->eval(x.p, sc.p, g.p, nullptr);
  vec_dot(ctx, n, g.p, d.p, sc.p + 1);
  if (ls->p.kind == LS_MORETHUENTE_B) vec_tmax_candidate(ctx, n, x.p, d.p, ls->lb.p, ls->ub.p, sc.p + 2);
  double h[16];
  sc.download(h, 16, ctx->stream);
  ctx->sync();
  LSMachine m;
  m.begin(ls->p, h[0], h[1], max_iter, h[2]);
  while (!m.done) {
    const double t = m.request(ls->p);
    obj->trial(x.p, d.p, t, m.wants_projection(ls->p), ls->lb.p, ls->ub.p, xt.p, gt.p, sc.p + 4);
    ctx->counters[2]++;
    sc.download(h, 16, ctx->stream);
    ctx->sync();
    m.feed(ls->p, h[4], h[5], h[6]);
  }
  *t_out = m.result;
  return OSB_OK;
  OSB_CATCH
}

int osb_linesearch_step_len_scalar(osb_linesearch* lsh, osb_phi_fn phi, void* user, double f0, double gd0, double tmaxc,
                                   int64_t max_iter, double* t_out, int* last_eval_is_result) {
  OSB_TRY
  LineSearch* ls = L(lsh);
  LSMachine m;
  m.begin(ls->p, f0, gd0, max_iter, tmaxc);
  while (!m.done) {
    double f = NAN, gd = NAN, dn = 0.0;
    phi(user, m.request(ls->p), m.wants_projection(ls->p) ? 1 : 0, &f, &gd, &dn);
    m.feed(ls->p, f, gd, dn);
  }
  *t_out = m.result;
  if (last_eval_is_result) *last_eval_is_result = m.last_eval_is_result ? 1 : 0;
  return OSB_OK;
  OSB_CATCH
}

// ---- solvers
int osb_solver_create(osb_ctx* ctx, int kind, int64_t n, double tol, const double* x0, const double* lb, const double* ub,
                      osb_objective* obj0, osb_solver** out) {
  OSB_TRY
  *out = reinterpret_cast<osb_solver*>(new Solver(C(ctx), kind, n, tol, x0, lb, ub, obj0 ? O(obj0) : nullptr));
  return OSB_OK;
  OSB_CATCH
}
void osb_solver_destroy(osb_solver* s) {
  if (s) delete S(s);
}
int osb_minimize(osb_solver* s, osb_linesearch* ls, osb_objective* obj, int64_t max_iter_solver, int64_t max_iter_ls,
                 osb_callback_fn cb, void* user) {
  OSB_TRY
  OSB_REQUIRE(s && ls && obj, OSB_ERR_BAD_HANDLE, "null handle");
  int rc = S(s)->minimize(L(ls), O(obj), max_iter_solver, max_iter_ls, cb, user);
  if (rc != OSB_OK) {
    static const char* names[] = {"", "Max iter reached", "Out of domain", "Error in input parameters", "Abnormal termination"};
    if (rc >= 1 && rc <= 4) osb::set_last_error(names[rc]);  // thiserror messages, ls_solver.rs:12-19
    else if (rc == OSB_PANIC_NOT_SPD) osb::set_last_error("Hessian is not positive definite (reference: cholesky().unwrap() panics)");
  }
  return rc;
  OSB_CATCH
}
int osb_solver_set_option(osb_solver* s, const char* name, int64_t value) {
  OSB_TRY
  std::string nm(name);
  if (nm == "engine") S(s)->engine = (int)value;
  else if (nm == "record_trace") S(s)->record_trace = (int)value;
  else if (nm == "callback_run_ahead") S(s)->callback_run_ahead = (int)value;
  else if (nm == "qn_kernel") S(s)->qn_variant = (int)value;
  else if (nm == "qn_schedule") S(s)->opt_schedule = (int)value;
  else if (nm == "qn_storage") S(s)->opt_storage = (int)value;
  else if (nm == "use_p2p") S(s)->use_p2p = (int)value;
  else if (nm == "head_kernel") S(s)->head_variant = (int)value;
  else if (nm == "profile_kernels") S(s)->profile_kernels = (int)value;
  else if (nm == "fused_iteration") S(s)->opt_fused = (int)value;
  else if (nm == "fused_stream") S(s)->opt_stream = (int)value;
  else if (nm == "profile_iter") S(s)->profile_iter = (int)value;
  else throw Error(OSB_ERROR_INPUT_PARAMS, "unknown option " + nm);
  return OSB_OK;
  OSB_CATCH
}
int osb_solver_set_lambdas(osb_solver* s, double lmin, double lmax) {
  S(s)->lambda_min = lmin;
  S(s)->lambda_max = lmax;
  return OSB_OK;
}
int64_t osb_solver_k(const osb_solver* s) { return S(s)->k; }
int64_t osb_solver_dim(const osb_solver* s) { return S(s)->n; }
int osb_solver_termination_reason(const osb_solver* s) { return S(s)->reason; }
int osb_solver_x(osb_solver* s, double* out) {
  OSB_TRY
  Solver* p = S(s);
  if (p->cb_x_mirror) {  // inside a run-ahead callback: the snapshot of this iteration (the device is already one ahead)
    std::memcpy(out, p->cb_x_mirror, sizeof(double) * (size_t)p->n);
    return OSB_OK;
  }
  p->ctx->use();
  p->x.download(out, p->n, p->ctx->stream);
  p->ctx->sync();
  return OSB_OK;
  OSB_CATCH
}
int osb_solver_set_x(osb_solver* s, const double* in) {
  OSB_TRY
  Solver* p = S(s);
  OSB_REQUIRE(p->cb_x_mirror == nullptr, OSB_ERR_UNSUPPORTED,
              "inside a run-ahead callback only x, f, grad, k, s_norm, y_norm are available (the device is already one "
              "iteration ahead): set option callback_run_ahead = 0 for callbacks that need anything else");
  p->ctx->use();
  p->x.upload(in, p->n, p->ctx->stream);
  p->ctx->sync();
  p->have_eval = false;
  p->u_valid = false;
  return OSB_OK;
  OSB_CATCH
}
int osb_solver_f(osb_solver* s, double* f_out) {
  OSB_TRY
  Solver* p = S(s);
  if (p->cb_state_mirror) {
    *f_out = p->cb_state_mirror->f;
    return OSB_OK;
  }
  p->ctx->use();
  p->fetch_state();
  *f_out = p->h_state->f;
  return OSB_OK;
  OSB_CATCH
}
int osb_solver_grad(osb_solver* s, double* out) {
  OSB_TRY
  Solver* p = S(s);
  if (p->cb_g_mirror) {
    std::memcpy(out, p->cb_g_mirror, sizeof(double) * (size_t)p->n);
    return OSB_OK;
  }
  p->ctx->use();
  p->g.download(out, p->n, p->ctx->stream);
  p->ctx->sync();
  return OSB_OK;
  OSB_CATCH
}
double osb_solver_s_norm(const osb_solver* s) { return S(s)->has_s ? S(s)->s_norm : NAN; }
double osb_solver_y_norm(const osb_solver* s) { return S(s)->has_y ? S(s)->y_norm : NAN; }
int osb_solver_clear_norms(osb_solver* s) {
  OSB_TRY
  Solver* p = S(s);
  OSB_REQUIRE(p->cb_x_mirror == nullptr, OSB_ERR_UNSUPPORTED,
              "inside a run-ahead callback only x, f, grad, k, s_norm, y_norm are available (the device is already one "
              "iteration ahead): set option callback_run_ahead = 0 for callbacks that need anything else");
  p->ctx->use();
  p->fetch_state();
  p->h_state->has_s = p->h_state->has_y = 0;
  p->push_state();
  p->has_s = p->has_y = false;
  return OSB_OK;
  OSB_CATCH
}
double osb_solver_lambda(const osb_solver* s) { return S(s)->lambda; }
double osb_solver_decrement_squared(const osb_solver* s) { return S(s)->has_dec ? S(s)->decrement_squared : NAN; }
int osb_solver_inv_hessian(osb_solver* s, double* out) {
  OSB_TRY
  Solver* p = S(s);
  OSB_REQUIRE(p->cb_x_mirror == nullptr, OSB_ERR_UNSUPPORTED,
              "inside a run-ahead callback only x, f, grad, k, s_norm, y_norm are available (the device is already one "
              "iteration ahead): set option callback_run_ahead = 0 for callbacks that need anything else");
  OSB_REQUIRE(p->is_qn || p->kind == OSB_PNORM, OSB_ERROR_INPUT_PARAMS, "solver holds no n x n matrix");
  p->ctx->use();
  p->flush_pending();
  p->ensure_full();
  // local row block [row0, row0 + nrows)
  OSB_CUDA(cudaMemcpy2DAsync(out + p->row0 * p->n, p->n * sizeof(double), p->H.p, p->ld * sizeof(double), p->n * sizeof(double),
                             p->nrows, cudaMemcpyDeviceToHost, p->ctx->stream));
  p->ctx->sync();
  return OSB_OK;
  OSB_CATCH
}
int osb_solver_set_inv_hessian(osb_solver* s, const double* in) {
  OSB_TRY
  Solver* p = S(s);
  OSB_REQUIRE(p->cb_x_mirror == nullptr, OSB_ERR_UNSUPPORTED,
              "inside a run-ahead callback only x, f, grad, k, s_norm, y_norm are available (the device is already one "
              "iteration ahead): set option callback_run_ahead = 0 for callbacks that need anything else");
  OSB_REQUIRE(p->is_qn || p->kind == OSB_PNORM, OSB_ERROR_INPUT_PARAMS, "solver holds no n x n matrix");
  p->ctx->use();
  p->flush_pending();
  p->ensure_full();
  {
    bool sym = true;  // packed symmetric storage is only valid for a symmetric matrix
    for (int64_t i = 0; i < p->n && sym; ++i)
      for (int64_t j = 0; j < i; ++j)
        if (in[i * p->n + j] != in[j * p->n + i]) {
          sym = false;
          break;
        }
    // BFGS / DFP / SR1 are implemented in their symmetric rank-2 / rank-1 forms (H y serves as row AND column factor):
    // with a non-symmetric H the reference's products (bfgs.rs:121-124, dfp.rs:120) would need H^T y as well.  The
    // crate has no setter for approx_inv_hessian (it starts from I and stays symmetric), so this is an input error.
    OSB_REQUIRE(sym || !p->is_qn || p->qn_kind == QN_BROYDEN, OSB_ERROR_INPUT_PARAMS,
                "BFGS / DFP / SR1 need a symmetric inverse-Hessian approximation (Broyden and PnormDescent accept any matrix)");
    p->h_symmetric = sym;
  }
  OSB_CUDA(cudaMemcpy2DAsync(p->H.p, p->ld * sizeof(double), in + p->row0 * p->n, p->n * sizeof(double), p->n * sizeof(double),
                             p->nrows, cudaMemcpyHostToDevice, p->ctx->stream));
  p->ctx->sync();
  p->u_valid = false;
  return OSB_OK;
  OSB_CATCH
}
int osb_solver_active_set(osb_solver* s, uint8_t* out) {
  OSB_TRY
  Solver* p = S(s);
  OSB_REQUIRE(p->cb_x_mirror == nullptr, OSB_ERR_UNSUPPORTED,
              "inside a run-ahead callback only x, f, grad, k, s_norm, y_norm are available (the device is already one "
              "iteration ahead): set option callback_run_ahead = 0 for callbacks that need anything else");
  OSB_REQUIRE(p->bounded, OSB_ERROR_INPUT_PARAMS, "solver has no bounds");
  p->ctx->use();
  uint8_t* d_out = nullptr;
  OSB_CUDA(cudaMalloc(&d_out, p->n));
  vec_active_set(p->ctx, p->n, p->x.p, p->lb.p, p->ub.p, d_out);
  OSB_CUDA(cudaMemcpyAsync(out, d_out, p->n, cudaMemcpyDeviceToHost, p->ctx->stream));
  p->ctx->sync();
  cudaFree(d_out);
  return OSB_OK;
  OSB_CATCH
}
int64_t osb_solver_trace_len(const osb_solver* s) { return (int64_t)S(s)->trace.size(); }
int osb_solver_trace(const osb_solver* s, double* f, double* t, double* sn, double* yn) {
  const auto& tr = S(s)->trace;
  for (size_t i = 0; i < tr.size(); ++i) {
    f[i] = tr[i].f;
    t[i] = tr[i].t;
    sn[i] = tr[i].s_norm;
    yn[i] = tr[i].y_norm;
  }
  return OSB_OK;
}
int osb_solver_kernel_timing(const osb_solver* s, double out[3]) {
  out[0] = S(s)->prof_ms[0];
  out[1] = S(s)->prof_ms[1];
  out[2] = S(s)->prof_ms[2];
  return OSB_OK;
}
int osb_solver_iter_profile(osb_solver* s, double out[16]) {
  OSB_TRY
  Solver* p = S(s);
  for (int q = 0; q < 16; ++q) out[q] = 0.0;
  if (!p->d_iter_prof) return OSB_OK;
  p->ctx->use();
  p->ctx->sync();
  long long v[16];
  OSB_CUDA(cudaMemcpy(v, p->d_iter_prof, sizeof(v), cudaMemcpyDeviceToHost));
  const double its = v[3] > 0 ? (double)v[3] : 1.0;
  for (int q = 0; q < 16; ++q) out[q] = (double)v[q] * 1e-6 / its;
  out[3] = (double)v[3];
  return OSB_OK;
  OSB_CATCH
}
int osb_solver_path_info(const osb_solver* s, int64_t out[8]) {
  const Solver* p = S(s);
  out[0] = p->last_engine;
  out[1] = p->qn_schedule;
  out[2] = p->qn_storage;
  out[3] = p->last_sym_sharded ? 1 : 0;
  out[4] = p->last_p2p ? 1 : 0;
  out[5] = p->ctx->world;
  out[6] = p->qn_variant;
  out[7] = (p->last_fused ? 1 : 0) | (p->last_stream ? 2 : 0);
  return OSB_OK;
}
int osb_solver_last_timing(const osb_solver* s, double* ms, int64_t* iters) {
  *ms = S(s)->last_ms;
  *iters = S(s)->last_iters;
  return OSB_OK;
}

// ---- batched
int osb_batched_bfgs_rosenbrock(osb_ctx* ctx, int64_t n, int64_t np, const double* x0, double tol, int64_t max_iter,
                                int64_t max_ls, double c1, double beta, double* x_out, double* f_out, int32_t* k_out,
                                int32_t* st_out, int32_t* reason_out, double* ms_out) {
  OSB_TRY
  C(ctx)->use();
  return batched_bfgs_rosenbrock(C(ctx), n, np, x0, false, 0, tol, max_iter, max_ls, c1, beta, x_out, f_out, k_out, st_out,
                                 reason_out, ms_out);
  OSB_CATCH
}
int osb_batched_bfgs_rosenbrock_generated(osb_ctx* ctx, int64_t n, int64_t np, int64_t problem0, double tol, int64_t max_iter,
                                          int64_t max_ls, double c1, double beta, double* x_out, double* f_out, int32_t* k_out,
                                          int32_t* st_out, int32_t* reason_out, double* ms_out) {
  OSB_TRY
  C(ctx)->use();
  return batched_bfgs_rosenbrock(C(ctx), n, np, nullptr, true, problem0, tol, max_iter, max_ls, c1, beta, x_out, f_out, k_out,
                                 st_out, reason_out, ms_out);
  OSB_CATCH
}

int osb_bench_grid_sync(osb_ctx* ctx, int reps, double* us_out) {
  OSB_TRY
  C(ctx)->use();
  *us_out = bench_grid_sync(C(ctx), reps);
  return OSB_OK;
  OSB_CATCH
}
int osb_bench_syrk(osb_ctx* ctx, osb_objective* logistic, int reps, double* ms_out) {
  OSB_TRY
  C(ctx)->use();
  *ms_out = bench_syrk_dmma(C(ctx), O(logistic), reps);
  return OSB_OK;
  OSB_CATCH
}

// ---- kernel micro-benchmarks (device-resident inputs)
int osb_bench_qn_kernel(osb_ctx* ctxh, int which, int64_t n, int reps, int variant, double* ms_out) {
  OSB_TRY
  Ctx* ctx = C(ctxh);
  ctx->use();
  const int64_t ld = qn_ld(n);
  DBuf H(which == 4 ? ld : qn_rows_padded(n) * ld), H2, a(ld), b(ld), c(ld), out(ld), scratch;
  DevState* st = nullptr;
  OSB_CUDA(cudaMalloc(&st, sizeof(DevState)));
  DevState hs;
  std::memset(&hs, 0, sizeof(hs));
  hs.c0 = 1e-9;
  hs.c1 = -1e-9;
  hs.c2 = 1e-9;
  hs.ys = 1.0;
  OSB_CUDA(cudaMemcpyAsync(st, &hs, sizeof(hs), cudaMemcpyHostToDevice, ctx->stream));
  H.zero(ctx->stream);
  {
    std::vector<double> v(ld, 0.0);
    for (int64_t i = 0; i < n; ++i) v[i] = (double)h16(21, (uint64_t)i, 0) / 32768.0;
    a.upload(v.data(), ld, ctx->stream);
    for (int64_t i = 0; i < n; ++i) v[i] = (double)h16(22, (uint64_t)i, 0) / 32768.0;
    b.upload(v.data(), ld, ctx->stream);
    for (int64_t i = 0; i < n; ++i) v[i] = (double)h16(23, (uint64_t)i, 0) / 32768.0;
    c.upload(v.data(), ld, ctx->stream);
    ctx->sync();
  }
  if (which == 4) {  // the packed-triangle pass alone (variant = option "qn_kernel"), on a synthetic matrix
    OSB_REQUIRE(ctx->world == 1, OSB_ERR_UNSUPPORTED, "kernel micro-benchmarks run on a single-GPU context");
    DBuf P(qn_sym_doubles(n)), P2, colpart((int64_t)2 * ctx->num_sms * 2 * ld);
    if (variant & 2) P2.alloc(qn_sym_doubles(n));
    hs.pc0 = 1e-9;
    hs.pc1 = -1e-9;
    hs.pc2 = 1e-9;
    OSB_CUDA(cudaMemcpyAsync(st, &hs, sizeof(hs), cudaMemcpyHostToDevice, ctx->stream));
    qn_sym_set_identity(ctx, n, P.p);
    colpart.zero(ctx->stream);
    QNLazyArgs la{};
    la.ld = ld;
    la.n = n;
    la.nrows = n;
    la.st = st;
    la.ps = a.p;
    la.ph = b.p;
    la.y = c.p;
    la.g = a.p;
    la.s = b.p;
    la.h = out.p;
    la.w = H.p;  // (any n-vector scratch)
    la.kind = QN_BFGS;
    cudaEvent_t f0, f1;
    OSB_CUDA(cudaEventCreate(&f0));
    OSB_CUDA(cudaEventCreate(&f1));
    for (int i = 0; i < 3; ++i) qn_launch_lazy_sym(ctx, la, P.p, (variant & 2) ? P2.p : P.p, colpart.p, n, ld, 0, variant);
    OSB_CUDA(cudaEventRecord(f0, ctx->stream));
    for (int i = 0; i < reps; ++i) qn_launch_lazy_sym(ctx, la, P.p, (variant & 2) ? P2.p : P.p, colpart.p, n, ld, 0, variant);
    OSB_CUDA(cudaEventRecord(f1, ctx->stream));
    OSB_CUDA(cudaEventSynchronize(f1));
    float fms = 0.f;
    OSB_CUDA(cudaEventElapsedTime(&fms, f0, f1));
    *ms_out = (double)fms / reps;
    cudaEventDestroy(f0);
    cudaEventDestroy(f1);
    cudaFree(st);
    return OSB_OK;
  }
  if (which == 2) scratch.alloc(((n + 63) / 64) * ld);
  if (which == 3) H2.alloc(qn_rows_padded(n) * ld);
  cudaEvent_t e0, e1;
  OSB_CUDA(cudaEventCreate(&e0));
  OSB_CUDA(cudaEventCreate(&e1));
  auto run = [&]() {
    if (which == 0) qn_launch_gemv(ctx, H.p, ld, n, 0, st, a.p, out.p, b.p, out.p, variant);
    else if (which == 1) qn_launch_update(ctx, QN_BFGS, H.p, ld, n, 0, st, a.p, b.p, c.p, c.p, out.p, variant);
    else if (which == 2) qn_launch_gemvT(ctx, H.p, ld, n, 0, st, a.p, out.p, scratch.p);
    else OSB_CUDA(cudaMemcpyAsync(H2.p, H.p, sizeof(double) * (size_t)(n * ld), cudaMemcpyDeviceToDevice, ctx->stream));
  };
  for (int i = 0; i < 3; ++i) run();
  OSB_CUDA(cudaEventRecord(e0, ctx->stream));
  for (int i = 0; i < reps; ++i) run();
  OSB_CUDA(cudaEventRecord(e1, ctx->stream));
  OSB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  OSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_out = (double)ms / reps;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(st);
  return OSB_OK;
  OSB_CATCH
}

}  // extern "C"
