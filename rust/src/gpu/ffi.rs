//! Raw bindings of `include/optsolv_b200.h` (libosb_b200.so, sm_100a).
//! Source only: this image has no Rust toolchain, so the module is written against the header and
//! exercised through the same C ABI by the Python/ctypes tests.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)] pub struct osb_ctx { _p: [u8; 0] }
#[repr(C)] pub struct osb_objective { _p: [u8; 0] }
#[repr(C)] pub struct osb_linesearch { _p: [u8; 0] }
#[repr(C)] pub struct osb_solver { _p: [u8; 0] }

pub const OSB_OK: c_int = 0;
pub const OSB_MAX_ITER_REACHED: c_int = 1;
pub const OSB_OUT_OF_DOMAIN: c_int = 2;
pub const OSB_ERROR_INPUT_PARAMS: c_int = 3;
pub const OSB_ABNORMAL_TERMINATION: c_int = 4;

pub const OSB_GD: c_int = 0;
pub const OSB_PGD: c_int = 1;
pub const OSB_SPG: c_int = 2;
pub const OSB_BFGS: c_int = 3;
pub const OSB_DFP: c_int = 4;
pub const OSB_BROYDEN: c_int = 5;
pub const OSB_BFGSB: c_int = 6;
pub const OSB_DFPB: c_int = 7;
pub const OSB_BROYDENB: c_int = 8;
pub const OSB_SR1B: c_int = 9;
pub const OSB_NEWTON: c_int = 10;
pub const OSB_PROJ_NEWTON: c_int = 11;
pub const OSB_SPN: c_int = 12;
pub const OSB_PNORM: c_int = 13;

pub type osb_host_eval_fn = unsafe extern "C" fn(user: *mut c_void, x: *const c_double, n: i64, f: *mut c_double,
                                                 g: *mut c_double, hess: *mut c_double) -> c_int;
pub type osb_device_eval_fn = unsafe extern "C" fn(user: *mut c_void, d_x: *const c_double, n: i64, d_f: *mut c_double,
                                                   d_g: *mut c_double, d_hess: *mut c_double, stream: *mut c_void) -> c_int;
pub type osb_callback_fn = unsafe extern "C" fn(user: *mut c_void, s: *mut osb_solver);

#[link(name = "osb_b200")]
extern "C" {
    pub fn osb_last_error_string() -> *const c_char;
    pub fn osb_ctx_create(device: c_int, out: *mut *mut osb_ctx) -> c_int;
    pub fn osb_ctx_create_dist(device: c_int, rank: c_int, world: c_int, nccl_unique_id: *const c_void,
                               out: *mut *mut osb_ctx) -> c_int;
    pub fn osb_nccl_unique_id(out128: *mut c_void) -> c_int;
    pub fn osb_ctx_ipc_handle(ctx: *mut osb_ctx, out64: *mut c_void) -> c_int;
    pub fn osb_ctx_ipc_connect(ctx: *mut osb_ctx, handles_world_x_64: *const c_void) -> c_int;
    pub fn osb_ctx_set_vector_sharding(ctx: *mut osb_ctx, on: c_int) -> c_int;
    pub fn osb_ctx_trim_memory(ctx: *mut osb_ctx) -> c_int;
    pub fn osb_ctx_destroy(ctx: *mut osb_ctx);

    pub fn osb_objective_create_dense_quadratic(ctx: *mut osb_ctx, n: i64, a: *const c_double, b: *const c_double,
                                                out: *mut *mut osb_objective) -> c_int;
    pub fn osb_objective_create_rosenbrock(ctx: *mut osb_ctx, n: i64, out: *mut *mut osb_objective) -> c_int;
    pub fn osb_objective_create_separable_quadratic_generated(ctx: *mut osb_ctx, n: i64, out: *mut *mut osb_objective) -> c_int;
    pub fn osb_objective_create_separable_quadratic_generated_shard(ctx: *mut osb_ctx, n_local: i64, index0: i64,
                                                                    out: *mut *mut osb_objective) -> c_int;
    pub fn osb_objective_create_logistic_generated(ctx: *mut osb_ctx, m: i64, n: i64, lambda: c_double,
                                                   out: *mut *mut osb_objective) -> c_int;
    pub fn osb_objective_create_host(ctx: *mut osb_ctx, n: i64, f: osb_host_eval_fn, user: *mut c_void,
                                     with_hessian: c_int, out: *mut *mut osb_objective) -> c_int;
    pub fn osb_objective_create_user(ctx: *mut osb_ctx, n: i64, f: osb_device_eval_fn, user: *mut c_void,
                                     with_hessian: c_int, out: *mut *mut osb_objective) -> c_int;
    pub fn osb_objective_destroy(o: *mut osb_objective);

    pub fn osb_linesearch_create_backtracking(c1: c_double, beta: c_double, out: *mut *mut osb_linesearch) -> c_int;
    pub fn osb_linesearch_create_backtracking_b(ctx: *mut osb_ctx, c1: c_double, beta: c_double, n: i64,
                                                lb: *const c_double, ub: *const c_double,
                                                out: *mut *mut osb_linesearch) -> c_int;
    pub fn osb_linesearch_create_morethuente(c1: c_double, c2: c_double, t_min: c_double, t_max: c_double,
                                             delta_min: c_double, delta: c_double, delta_max: c_double,
                                             out: *mut *mut osb_linesearch) -> c_int;
    pub fn osb_linesearch_create_morethuente_b(ctx: *mut osb_ctx, c1: c_double, c2: c_double, t_min: c_double,
                                               t_max: c_double, delta_min: c_double, delta: c_double,
                                               delta_max: c_double, n: i64, lb: *const c_double, ub: *const c_double,
                                               out: *mut *mut osb_linesearch) -> c_int;
    pub fn osb_linesearch_create_gll_quadratic(c1: c_double, m: i64, sigma1: c_double, sigma2: c_double,
                                               out: *mut *mut osb_linesearch) -> c_int;
    pub fn osb_linesearch_create_nosearch(out: *mut *mut osb_linesearch) -> c_int;
    pub fn osb_linesearch_destroy(ls: *mut osb_linesearch);

    pub fn osb_solver_create(ctx: *mut osb_ctx, kind: c_int, n: i64, tol: c_double, x0: *const c_double,
                             lb: *const c_double, ub: *const c_double, objective_for_lambda0: *mut osb_objective,
                             out: *mut *mut osb_solver) -> c_int;
    pub fn osb_solver_destroy(s: *mut osb_solver);
    pub fn osb_minimize(s: *mut osb_solver, ls: *mut osb_linesearch, obj: *mut osb_objective, max_iter_solver: i64,
                        max_iter_line_search: i64, callback: Option<osb_callback_fn>, user: *mut c_void) -> c_int;
    pub fn osb_solver_k(s: *const osb_solver) -> i64;
    pub fn osb_solver_x(s: *mut osb_solver, out: *mut c_double) -> c_int;
    pub fn osb_solver_s_norm(s: *const osb_solver) -> c_double;
    pub fn osb_solver_y_norm(s: *const osb_solver) -> c_double;
    pub fn osb_solver_lambda(s: *const osb_solver) -> c_double;
    pub fn osb_solver_decrement_squared(s: *const osb_solver) -> c_double;
    pub fn osb_solver_inv_hessian(s: *mut osb_solver, out: *mut c_double) -> c_int;
    pub fn osb_solver_termination_reason(s: *const osb_solver) -> c_int;
}
