//! `optimization_solvers::gpu` — the B200 (sm_100a) backend behind the crate's own traits.
//!
//! Integration (see INTEGRATION.md for the exact patch):
//!   * `src/lib.rs`:              `#[cfg(feature = "gpu")] pub mod gpu;`
//!   * `Cargo.toml`:              `[features] gpu = []`, `build.rs` prints `cargo:rustc-link-search=<dir of libosb_b200.so>`
//!   * `src/line_search/mod.rs`:  ONE provided method on the `LineSearch` trait,
//!                                `fn gpu_spec(&self) -> Option<crate::gpu::LineSearchSpec> { None }`,
//!                                implemented by the six searches of the crate (they own the parameters); nothing else
//!                                of the reference changes.
//!
//! What stays: the traits `ComputeDirection`, `LineSearchSolver`, `HasBounds`, `HasProjectedGradient`, `LineSearch`,
//! `FuncEvalMultivariate`, `SolverError`, `Tracer`, and the call a user writes (examples/bfgs_example.rs:46-52):
//!
//! ```ignore
//! let mut solver = gpu::BFGS::new(tol, x0);
//! solver.minimize(&mut ls, oracle, max_iter_solver, max_iter_line_search, None)?;
//! ```
//!
//! Every solver struct of the hot path exists here under its own name with the reference's constructor arguments and
//! getters.  `LineSearchSolver::minimize` is OVERRIDDEN ("Methods that are already implemented can be freely overriden",
//! src/ls_solver.rs:22): the whole loop of src/ls_solver.rs:66-111 runs inside `osb_minimize`, with iterate, gradient and
//! the dense inverse-Hessian approximation resident on the device.
//!
//! The oracle keeps its type, `impl FnMut(&DVector<f64>) -> FuncEvalMultivariate`:
//!   * a device objective (`ExtendedRosenbrock`, `DenseQuadratic`, `SeparableQuadratic`, `LogisticRegression`,
//!     `UserDeviceObjective`) hands out such a closure with `.oracle()`.  The closure really evaluates on the GPU when it is
//!     called, and it leaves a note in a thread-local: "device objective H was just evaluated at the vector stored at P".
//!     The overridden `minimize` starts, like the reference (ls_solver.rs:79), by calling the oracle at `x_k`; when the
//!     note names that very vector, the oracle IS the device objective H and the solve is dispatched to the device-resident
//!     path with H.  No sentinel values, no extra trait bound, one redundant evaluation.
//!   * any other closure is an ordinary host oracle: it is wrapped (`osb_objective_create_host`) and called back from
//!     the library with host copies of x (H2D / D2H per call — functional, not fast).
pub mod ffi;

use crate::{
    ComputeDirection, Floating, FuncEvalMultivariate, HasBounds, LineSearch, LineSearchSolver, SolverError,
};
use nalgebra::{DMatrix, DVector};
use std::cell::Cell;
use std::ffi::CStr;
use std::os::raw::{c_int, c_void};
use std::ptr;
use std::rc::Rc;
use tracing::{error, info, warn};

// ---------------------------------------------------------------------------------------------------------------------
// context: one per GPU (one process per GPU); a process-wide default like the reference has none to configure
// ---------------------------------------------------------------------------------------------------------------------
pub struct GpuContext {
    raw: *mut ffi::osb_ctx,
}
impl GpuContext {
    pub fn new(device: i32) -> Result<Rc<Self>, SolverError> {
        install_log_bridge();
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::osb_ctx_create(device, &mut raw) })?;
        Ok(Rc::new(Self { raw }))
    }
    /// One rank of a multi-GPU solve (H sharded over the ranks).  `nccl_unique_id` comes from rank 0
    /// (`GpuContext::nccl_unique_id`) through whatever the application uses to talk between its processes.
    pub fn new_dist(device: i32, rank: i32, world: i32, nccl_unique_id: &[u8; 128]) -> Result<Rc<Self>, SolverError> {
        install_log_bridge();
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::osb_ctx_create_dist(device, rank, world, nccl_unique_id.as_ptr() as *const c_void, &mut raw) })?;
        Ok(Rc::new(Self { raw }))
    }
    pub fn nccl_unique_id() -> Result<[u8; 128], SolverError> {
        let mut id = [0u8; 128];
        check(unsafe { ffi::osb_nccl_unique_id(id.as_mut_ptr() as *mut c_void) })?;
        Ok(id)
    }
    /// CUDA IPC handle of this rank's exchange region; all-gather the handles and call `ipc_connect` on every rank
    /// (the exchange is then fused into the kernels over NVLink peer memory instead of NCCL calls).
    pub fn ipc_handle(&self) -> Result<[u8; 64], SolverError> {
        let mut h = [0u8; 64];
        check(unsafe { ffi::osb_ctx_ipc_handle(self.raw, h.as_mut_ptr() as *mut c_void) })?;
        Ok(h)
    }
    pub fn ipc_connect(&self, handles_in_rank_order: &[[u8; 64]]) -> Result<(), SolverError> {
        let flat: Vec<u8> = handles_in_rank_order.iter().flat_map(|h| h.iter().copied()).collect();
        check(unsafe { ffi::osb_ctx_ipc_connect(self.raw, flat.as_ptr() as *const c_void) })
    }
    /// Every rank calls this when the connect did not succeed on every rank (falls back to NCCL together).
    pub fn ipc_close(&self) -> Result<(), SolverError> {
        check(unsafe { ffi::osb_ctx_ipc_close(self.raw) })
    }
    pub fn set_vector_sharding(&self, on: bool) -> Result<(), SolverError> {
        check(unsafe { ffi::osb_ctx_set_vector_sharding(self.raw, on as c_int) })
    }
    pub fn rank(&self) -> i32 {
        unsafe { ffi::osb_ctx_rank(self.raw) }
    }
    pub fn world(&self) -> i32 {
        unsafe { ffi::osb_ctx_world(self.raw) }
    }
    pub fn synchronize(&self) -> Result<(), SolverError> {
        check(unsafe { ffi::osb_ctx_synchronize(self.raw) })
    }
    pub fn raw(&self) -> *mut ffi::osb_ctx {
        self.raw
    }
}
impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { ffi::osb_ctx_destroy(self.raw) }
    }
}

thread_local! {
    static DEFAULT_CTX: std::cell::RefCell<Option<Rc<GpuContext>>> = std::cell::RefCell::new(None);
    /// (objective handle, address of the vector it was last evaluated at) — see the module documentation
    static LAST_DEVICE_CALL: Cell<(*mut ffi::osb_objective, *const f64)> = Cell::new((ptr::null_mut(), ptr::null()));
}
/// The context `X::new(tol, x0)` uses when none is given: device `OSB_DEVICE` / `LOCAL_RANK` / 0.
pub fn default_context() -> Rc<GpuContext> {
    DEFAULT_CTX.with(|c| {
        c.borrow_mut()
            .get_or_insert_with(|| {
                let dev = std::env::var("OSB_DEVICE").or_else(|_| std::env::var("LOCAL_RANK")).ok().and_then(|s| s.parse().ok()).unwrap_or(0);
                GpuContext::new(dev).expect("no CUDA device: the gpu backend has no CPU fallback")
            })
            .clone()
    })
}
pub fn set_default_context(ctx: Rc<GpuContext>) {
    DEFAULT_CTX.with(|c| *c.borrow_mut() = Some(ctx));
}

// ---------------------------------------------------------------------------------------------------------------------
// Tracer (src/tracer.rs): the crate's subscriber is used as it is; the library's events are forwarded into `tracing`
// ---------------------------------------------------------------------------------------------------------------------
unsafe extern "C" fn log_bridge(_user: *mut c_void, level: c_int, target: *const std::os::raw::c_char, message: *const std::os::raw::c_char) {
    let target = CStr::from_ptr(target).to_string_lossy();
    let message = CStr::from_ptr(message).to_string_lossy();
    // `tracing` wants the target as a literal: one arm per target the reference uses on this path
    macro_rules! emit {
        ($t:literal) => {
            match level {
                1 => error!(target: $t, "{}", message),
                2 => warn!(target: $t, "{}", message),
                3 => info!(target: $t, "{}", message),
                4 => tracing::debug!(target: $t, "{}", message),
                _ => tracing::trace!(target: $t, "{}", message),
            }
        };
    }
    match target.as_ref() {
        "bfgs" => emit!("bfgs"),
        "dfp" => emit!("dfp"),
        "broyden" => emit!("broyden"),
        "bfgs_b" => emit!("bfgs_b"),
        "dfp_b" => emit!("dfp_b"),
        "broyden_b" => emit!("broyden_b"),
        "sr1_b" => emit!("sr1_b"),
        "newton" => emit!("newton"),
        "projected_newton" => emit!("projected_newton"),
        _ => emit!("solver"),
    }
}
fn install_log_bridge() {
    static ONCE: std::sync::Once = std::sync::Once::new();
    ONCE.call_once(|| unsafe {
        ffi::osb_set_log_callback(Some(log_bridge), ptr::null_mut());
    });
}

// ---------------------------------------------------------------------------------------------------------------------
// objectives: device functors behind the crate's oracle type
// ---------------------------------------------------------------------------------------------------------------------
pub struct ObjectiveHandle {
    raw: *mut ffi::osb_objective,
    n: usize,
    with_hessian: bool,
    _ctx: Rc<GpuContext>,
}
impl Drop for ObjectiveHandle {
    fn drop(&mut self) {
        unsafe { ffi::osb_objective_destroy(self.raw) }
    }
}
impl ObjectiveHandle {
    /// f, g (and the Hessian when the objective has one) at a host vector: `osb_objective_eval`
    fn eval(&self, x: &DVector<Floating>) -> FuncEvalMultivariate {
        assert_eq!(x.len(), self.n, "oracle called with a vector of the wrong dimension");
        let mut f = 0.0;
        let mut g = DVector::zeros(self.n);
        let mut h = if self.with_hessian { vec![0.0; self.n * self.n] } else { Vec::new() };
        let rc = unsafe {
            ffi::osb_objective_eval(self.raw, x.as_ptr(), &mut f, g.as_mut_ptr(), if self.with_hessian { h.as_mut_ptr() } else { ptr::null_mut() })
        };
        if rc != ffi::OSB_OK {
            panic!("device objective: {}", last_error());
        }
        LAST_DEVICE_CALL.with(|c| c.set((self.raw, x.as_ptr())));
        let eval = FuncEvalMultivariate::new(f, g);
        if self.with_hessian {
            // row-major from the library; the Hessians of the path are symmetric, from_row_slice is exact anyway
            eval.with_hessian(DMatrix::from_row_slice(self.n, self.n, &h))
        } else {
            eval
        }
    }
}

/// Common surface of the device objectives.
pub trait DeviceObjective {
    fn handle(&self) -> &Rc<ObjectiveHandle>;
    /// The objective as the crate's oracle type (`FnMut(&DVector<f64>) -> FuncEvalMultivariate`, ls_solver.rs:34).
    fn oracle(&self) -> Box<dyn FnMut(&DVector<Floating>) -> FuncEvalMultivariate> {
        let h = self.handle().clone();
        Box::new(move |x: &DVector<Floating>| h.eval(x))
    }
    fn calls(&self) -> i64 {
        unsafe { ffi::osb_objective_calls(self.handle().raw) }
    }
    fn dim(&self) -> usize {
        self.handle().n
    }
}
macro_rules! device_objective {
    ($name:ident) => {
        pub struct $name {
            h: Rc<ObjectiveHandle>,
        }
        impl DeviceObjective for $name {
            fn handle(&self) -> &Rc<ObjectiveHandle> {
                &self.h
            }
        }
    };
}
fn wrap(raw: *mut ffi::osb_objective, n: usize, with_hessian: bool, ctx: Rc<GpuContext>) -> Rc<ObjectiveHandle> {
    Rc::new(ObjectiveHandle { raw, n, with_hessian, _ctx: ctx })
}

device_objective!(ExtendedRosenbrock);
impl ExtendedRosenbrock {
    /// f = sum_{i < n/2} 100 (x_{2i+1} - x_{2i}^2)^2 + (1 - x_{2i})^2
    pub fn new(n: usize) -> Result<Self, SolverError> {
        Self::new_in(default_context(), n)
    }
    pub fn new_in(ctx: Rc<GpuContext>, n: usize) -> Result<Self, SolverError> {
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::osb_objective_create_rosenbrock(ctx.raw, n as i64, &mut raw) })?;
        Ok(Self { h: wrap(raw, n, false, ctx) })
    }
}

device_objective!(DenseQuadratic);
impl DenseQuadratic {
    /// f = x.(A x) [- 2 b.x], g = 2 A x [- 2 b] (the pattern of examples/quadratic.rs:10-14); the Hessian is 2A
    pub fn new(a: &DMatrix<Floating>, b: Option<&DVector<Floating>>) -> Result<Self, SolverError> {
        let ctx = default_context();
        let n = a.nrows();
        let row_major: Vec<f64> = (0..n).flat_map(|i| (0..n).map(move |j| (i, j))).map(|(i, j)| a[(i, j)]).collect();
        let mut raw = ptr::null_mut();
        check(unsafe {
            ffi::osb_objective_create_dense_quadratic(ctx.raw, n as i64, row_major.as_ptr(), b.map_or(ptr::null(), |v| v.as_ptr()), &mut raw)
        })?;
        Ok(Self { h: wrap(raw, n, true, ctx) })
    }
    /// the synthetic SPD quadratic of SURVEY 8d (config C2), generated on the device; returns the objective and its x0
    pub fn generated(n: usize, shifted: bool) -> Result<(Self, DVector<Floating>), SolverError> {
        let ctx = default_context();
        let mut x0 = DVector::zeros(n);
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::osb_objective_create_dense_quadratic_generated(ctx.raw, n as i64, shifted as c_int, x0.as_mut_ptr(), &mut raw) })?;
        Ok((Self { h: wrap(raw, n, true, ctx) }, x0))
    }
}

device_objective!(SeparableQuadratic);
impl SeparableQuadratic {
    /// f = sum 0.5 c_i (x_i - a_i)^2 with the generated c, a of SURVEY 8d (config C5b)
    pub fn generated(n: usize) -> Result<Self, SolverError> {
        let ctx = default_context();
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::osb_objective_create_separable_quadratic_generated(ctx.raw, n as i64, &mut raw) })?;
        Ok(Self { h: wrap(raw, n, false, ctx) })
    }
    /// coordinates [index0, index0 + n_local) of the generated problem (index-range sharded GD / PGD / SPG)
    pub fn generated_shard(ctx: Rc<GpuContext>, n_local: usize, index0: usize) -> Result<Self, SolverError> {
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::osb_objective_create_separable_quadratic_generated_shard(ctx.raw, n_local as i64, index0 as i64, &mut raw) })?;
        Ok(Self { h: wrap(raw, n_local, false, ctx) })
    }
}

device_objective!(LogisticRegression);
impl LogisticRegression {
    /// l2-regularised logistic regression on m generated samples (config C5a); samples are sharded over the ranks of `ctx`
    pub fn generated(m: usize, n: usize, lambda: Floating) -> Result<Self, SolverError> {
        Self::generated_in(default_context(), m, n, lambda)
    }
    pub fn generated_in(ctx: Rc<GpuContext>, m: usize, n: usize, lambda: Floating) -> Result<Self, SolverError> {
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::osb_objective_create_logistic_generated(ctx.raw, m as i64, n as i64, lambda, &mut raw) })?;
        Ok(Self { h: wrap(raw, n, true, ctx) })
    }
}

/// A user-supplied device functor: `enqueue(d_x, n, d_f, d_g, d_hess, stream)` launches the user's own kernels on the
/// library's stream (`osb_device_eval_fn`).
pub struct UserDeviceObjective {
    h: Rc<ObjectiveHandle>,
    _f: Box<Box<dyn FnMut(*const f64, i64, *mut f64, *mut f64, *mut f64, *mut c_void) -> i32>>,
}
unsafe extern "C" fn user_trampoline(user: *mut c_void, d_x: *const f64, n: i64, d_f: *mut f64, d_g: *mut f64, d_h: *mut f64, stream: *mut c_void) -> c_int {
    let f = &mut *(user as *mut Box<dyn FnMut(*const f64, i64, *mut f64, *mut f64, *mut f64, *mut c_void) -> i32>);
    f(d_x, n, d_f, d_g, d_h, stream) as c_int
}
impl UserDeviceObjective {
    pub fn new(n: usize, with_hessian: bool, enqueue: impl FnMut(*const f64, i64, *mut f64, *mut f64, *mut f64, *mut c_void) -> i32 + 'static) -> Result<Self, SolverError> {
        let ctx = default_context();
        let mut boxed: Box<Box<dyn FnMut(*const f64, i64, *mut f64, *mut f64, *mut f64, *mut c_void) -> i32>> = Box::new(Box::new(enqueue));
        let mut raw = ptr::null_mut();
        check(unsafe {
            ffi::osb_objective_create_user(ctx.raw, n as i64, user_trampoline, &mut *boxed as *mut _ as *mut c_void, with_hessian as c_int, &mut raw)
        })?;
        Ok(Self { h: wrap(raw, n, with_hessian, ctx), _f: boxed })
    }
}
impl DeviceObjective for UserDeviceObjective {
    fn handle(&self) -> &Rc<ObjectiveHandle> {
        &self.h
    }
}

/// A host closure wrapped for `osb_minimize` (compatibility path): lives for the duration of one `minimize` call.
struct HostClosure<'a> {
    raw: *mut ffi::osb_objective,
    _f: Box<&'a mut dyn FnMut(&DVector<Floating>) -> FuncEvalMultivariate>,
}
unsafe extern "C" fn host_trampoline(user: *mut c_void, x: *const f64, n: i64, f: *mut f64, g: *mut f64, hess: *mut f64) -> c_int {
    let closure = &mut *(user as *mut &mut dyn FnMut(&DVector<Floating>) -> FuncEvalMultivariate);
    let n = n as usize;
    let xv = DVector::from_column_slice(std::slice::from_raw_parts(x, n));
    let eval = closure(&xv);
    *f = *eval.f();
    std::slice::from_raw_parts_mut(g, n).copy_from_slice(eval.g().as_slice());
    match (hess.is_null(), eval.hessian()) {
        (false, Some(h)) => {
            let out = std::slice::from_raw_parts_mut(hess, n * n);  // row-major out
            for i in 0..n {
                for j in 0..n {
                    out[i * n + j] = h[(i, j)];
                }
            }
            1
        }
        _ => 0,
    }
}
impl<'a> HostClosure<'a> {
    fn new(ctx: &GpuContext, n: usize, f: &'a mut dyn FnMut(&DVector<Floating>) -> FuncEvalMultivariate, with_hessian: bool) -> Result<Self, SolverError> {
        let mut boxed = Box::new(f);
        let mut raw = ptr::null_mut();
        check(unsafe {
            ffi::osb_objective_create_host(ctx.raw, n as i64, host_trampoline, &mut *boxed as *mut _ as *mut c_void, with_hessian as c_int, &mut raw)
        })?;
        Ok(Self { raw, _f: boxed })
    }
}
impl Drop for HostClosure<'_> {
    fn drop(&mut self) {
        unsafe { ffi::osb_objective_destroy(self.raw) }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// line searches: the crate's own structs, described to the library by ONE provided method of the `LineSearch` trait
// ---------------------------------------------------------------------------------------------------------------------
/// What `LineSearch::gpu_spec` returns (the constructor parameters of the reference's searches, src/line_search/*.rs).
#[derive(Debug, Clone)]
pub enum LineSearchSpec {
    BackTracking { c1: Floating, beta: Floating },
    BackTrackingB { c1: Floating, beta: Floating, lower_bound: DVector<Floating>, upper_bound: DVector<Floating> },
    MoreThuente { c1: Floating, c2: Floating, t_min: Floating, t_max: Floating, delta_min: Floating, delta: Floating, delta_max: Floating },
    MoreThuenteB {
        c1: Floating, c2: Floating, t_min: Floating, t_max: Floating, delta_min: Floating, delta: Floating, delta_max: Floating,
        lower_bound: DVector<Floating>, upper_bound: DVector<Floating>,
    },
    GLLQuadratic { c1: Floating, m: usize, sigma1: Floating, sigma2: Floating },
    NoSearch,
}
struct LsHandle(*mut ffi::osb_linesearch);
impl Drop for LsHandle {
    fn drop(&mut self) {
        unsafe { ffi::osb_linesearch_destroy(self.0) }
    }
}
fn build_line_search(ctx: &GpuContext, spec: &LineSearchSpec) -> Result<LsHandle, SolverError> {
    let mut raw = ptr::null_mut();
    let rc = unsafe {
        match spec {
            LineSearchSpec::BackTracking { c1, beta } => ffi::osb_linesearch_create_backtracking(*c1, *beta, &mut raw),
            LineSearchSpec::BackTrackingB { c1, beta, lower_bound, upper_bound } => {
                ffi::osb_linesearch_create_backtracking_b(ctx.raw, *c1, *beta, lower_bound.len() as i64, lower_bound.as_ptr(), upper_bound.as_ptr(), &mut raw)
            }
            LineSearchSpec::MoreThuente { c1, c2, t_min, t_max, delta_min, delta, delta_max } => {
                ffi::osb_linesearch_create_morethuente(*c1, *c2, *t_min, *t_max, *delta_min, *delta, *delta_max, &mut raw)
            }
            LineSearchSpec::MoreThuenteB { c1, c2, t_min, t_max, delta_min, delta, delta_max, lower_bound, upper_bound } => {
                ffi::osb_linesearch_create_morethuente_b(ctx.raw, *c1, *c2, *t_min, *t_max, *delta_min, *delta, *delta_max, lower_bound.len() as i64,
                                                         lower_bound.as_ptr(), upper_bound.as_ptr(), &mut raw)
            }
            LineSearchSpec::GLLQuadratic { c1, m, sigma1, sigma2 } => ffi::osb_linesearch_create_gll_quadratic(*c1, *m as i64, *sigma1, *sigma2, &mut raw),
            LineSearchSpec::NoSearch => ffi::osb_linesearch_create_nosearch(&mut raw),
        }
    };
    check(rc)?;
    Ok(LsHandle(raw))
}

// ---------------------------------------------------------------------------------------------------------------------
// solvers
// ---------------------------------------------------------------------------------------------------------------------
/// Why `minimize` returned `Ok(())` — the reference only reveals it through `warn!` / `info!` (bfgs.rs:68,71).
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum TerminationReason {
    GradTol,
    SNormTooSmall,
    YNormTooSmall,
    ProjGradTol,
    NewtonDecrement,
}

/// State shared by every solver struct of this module (the device handle and the lazily refreshed host mirrors).
pub struct GpuSolverCore {
    raw: *mut ffi::osb_solver,
    ctx: Rc<GpuContext>,
    kind: c_int,
    x: DVector<Floating>,
    x_dirty: bool,  // `xk_mut()` handed out a mutable borrow: upload before the next minimize
    k: usize,
    tol: Floating,
    lower_bound: DVector<Floating>,
    upper_bound: DVector<Floating>,
    ls_cache: Option<LsHandle>,  // MoreThuenteB.t_max / GLLQuadratic.f_previous persist inside the line-search object
}
impl Drop for GpuSolverCore {
    fn drop(&mut self) {
        self.ls_cache = None;
        unsafe { ffi::osb_solver_destroy(self.raw) }
    }
}
struct CallbackCtx<'a, S> {
    solver: *mut S,
    callback: &'a mut dyn FnMut(&S),
}
unsafe extern "C" fn callback_trampoline<S: GpuSolver>(user: *mut c_void, _s: *mut ffi::osb_solver) {
    let c = &mut *(user as *mut CallbackCtx<S>);
    let solver = &mut *c.solver;
    solver.core_mut().refresh_after_iteration();
    (c.callback)(&*c.solver);
}
impl GpuSolverCore {
    fn create(ctx: Rc<GpuContext>, kind: c_int, tol: Floating, x0: DVector<Floating>, bounds: Option<(DVector<Floating>, DVector<Floating>)>,
              lambda0_oracle: Option<*mut ffi::osb_objective>) -> Self {
        let (lb, ub) = bounds.unwrap_or((DVector::zeros(0), DVector::zeros(0)));
        let mut raw = ptr::null_mut();
        let rc = unsafe {
            ffi::osb_solver_create(ctx.raw, kind, x0.len() as i64, tol, x0.as_ptr(), if lb.len() > 0 { lb.as_ptr() } else { ptr::null() },
                                   if ub.len() > 0 { ub.as_ptr() } else { ptr::null() }, lambda0_oracle.unwrap_or(ptr::null_mut()), &mut raw)
        };
        if rc != ffi::OSB_OK {
            panic!("gpu solver construction failed: {}", last_error());
        }
        let mut core = Self { raw, ctx, kind, x: x0, x_dirty: false, k: 0, tol, lower_bound: lb, upper_bound: ub, ls_cache: None };
        core.download_x();  // constructors of the bounded solvers project x0 (bfgs_b.rs:50)
        core
    }
    fn download_x(&mut self) {
        unsafe { ffi::osb_solver_x(self.raw, self.x.as_mut_ptr()) };
    }
    fn refresh_after_iteration(&mut self) {
        self.download_x();
        self.k = unsafe { ffi::osb_solver_k(self.raw) } as usize;
    }
    pub fn set_option(&mut self, name: &str, value: i64) -> Result<(), SolverError> {
        let c = std::ffi::CString::new(name).map_err(|_| SolverError::ErrorInputParams)?;
        check(unsafe { ffi::osb_solver_set_option(self.raw, c.as_ptr(), value) })
    }
    pub fn termination_reason(&self) -> Option<TerminationReason> {
        match unsafe { ffi::osb_solver_termination_reason(self.raw) } {
            ffi::OSB_REASON_GRAD_TOL => Some(TerminationReason::GradTol),
            ffi::OSB_REASON_S_NORM => Some(TerminationReason::SNormTooSmall),
            ffi::OSB_REASON_Y_NORM => Some(TerminationReason::YNormTooSmall),
            ffi::OSB_REASON_PROJ_GRAD_TOL => Some(TerminationReason::ProjGradTol),
            ffi::OSB_REASON_NEWTON_DECREMENT => Some(TerminationReason::NewtonDecrement),
            _ => None,
        }
    }
    fn opt(v: f64) -> Option<f64> {
        if v.is_nan() { None } else { Some(v) }
    }
    fn s_norm(&self) -> Option<Floating> {
        Self::opt(unsafe { ffi::osb_solver_s_norm(self.raw) })
    }
    fn y_norm(&self) -> Option<Floating> {
        Self::opt(unsafe { ffi::osb_solver_y_norm(self.raw) })
    }
    fn matrix(&self) -> DMatrix<Floating> {
        let n = self.x.len();
        let mut row_major = vec![0.0; n * n];
        unsafe { ffi::osb_solver_inv_hessian(self.raw, row_major.as_mut_ptr()) };
        DMatrix::from_row_slice(n, n, &row_major)
    }

    /// `LineSearchSolver::minimize` (src/ls_solver.rs:66-111) for every solver of this module.
    fn minimize<S: GpuSolver, LS: LineSearch>(solver: &mut S, line_search: &mut LS,
                                              mut oracle: impl FnMut(&DVector<Floating>) -> FuncEvalMultivariate, max_iter_solver: usize,
                                              max_iter_line_search: usize, callback: Option<&mut dyn FnMut(&S)>) -> Result<(), SolverError> {
        let needs_hessian = S::NEEDS_HESSIAN;
        let spec = line_search.gpu_spec().expect("this line search has no device form: implement LineSearch::gpu_spec (INTEGRATION.md)");
        let solver_ptr: *mut S = solver;
        let core = solver.core_mut();
        if core.x_dirty {
            check(unsafe { ffi::osb_solver_set_x(core.raw, core.x.as_ptr()) })?;
            core.x_dirty = false;
        }
        core.k = 0;  // ls_solver.rs:74
        // bounded searches and the stateful ones keep their device object across minimize() calls of the same solver
        let ls = match core.ls_cache.take() {
            Some(h) => h,
            None => build_line_search(&core.ctx, &spec)?,
        };
        // ls_solver.rs:79: the first thing the loop does is evaluate the oracle at x_k.  A device objective's closure
        // notes its handle and the address it was called with: if that is x_k itself, the oracle is that objective.
        LAST_DEVICE_CALL.with(|c| c.set((ptr::null_mut(), ptr::null())));
        let first = oracle(&core.x);
        let (noted, at) = LAST_DEVICE_CALL.with(|c| c.get());
        let device_objective = if !noted.is_null() && at == core.x.as_ptr() { Some(noted) } else { None };
        if device_objective.is_none() && (first.f().is_nan() || first.f().is_infinite()) {
            error!(target: "solver", "Minimization completed: next iterate is out of domain");
            core.ls_cache = Some(ls);
            return Err(SolverError::OutOfDomain);
        }
        let mut host_oracle = oracle;
        let host_wrapper;
        let objective = match device_objective {
            Some(h) => h,
            None => {
                host_wrapper = HostClosure::new(&core.ctx, core.x.len(), &mut host_oracle, needs_hessian)?;
                host_wrapper.raw
            }
        };
        let rc = match callback {
            Some(cb) => {
                let mut cctx = CallbackCtx { solver: solver_ptr, callback: cb };
                unsafe {
                    ffi::osb_minimize(core.raw, ls.0, objective, max_iter_solver as i64, max_iter_line_search as i64, Some(callback_trampoline::<S>),
                                      &mut cctx as *mut _ as *mut c_void)
                }
            }
            None => unsafe { ffi::osb_minimize(core.raw, ls.0, objective, max_iter_solver as i64, max_iter_line_search as i64, None, ptr::null_mut()) },
        };
        let core = unsafe { &mut *solver_ptr }.core_mut();
        core.ls_cache = Some(ls);
        core.refresh_after_iteration();
        // (the events of ls_solver.rs:38,82-86,109, bfgs.rs:68,71 and newton/mod.rs:44 are emitted by the library through
        //  the bridge below: same targets, levels and messages as the reference)
        check(rc)
    }
}

/// Implemented by every solver struct below (gives the shared `minimize` access to the core).
pub trait GpuSolver: Sized {
    const NEEDS_HESSIAN: bool;
    fn core(&self) -> &GpuSolverCore;
    fn core_mut(&mut self) -> &mut GpuSolverCore;
}

macro_rules! gpu_solver {
    ($name:ident, $kind:expr, needs_hessian = $nh:expr, converged = $conv:expr) => {
        impl GpuSolver for $name {
            const NEEDS_HESSIAN: bool = $nh;
            fn core(&self) -> &GpuSolverCore {
                &self.core
            }
            fn core_mut(&mut self) -> &mut GpuSolverCore {
                &mut self.core
            }
        }
        impl $name {
            pub fn x(&self) -> &DVector<Floating> {
                &self.core.x
            }
            pub fn k(&self) -> &usize {
                &self.core.k
            }
            pub fn termination_reason(&self) -> Option<TerminationReason> {
                self.core.termination_reason()
            }
            /// tuning knobs of the backend (include/optsolv_b200.h, "options"); every one has a default
            pub fn set_option(mut self, name: &str, value: i64) -> Self {
                self.core.set_option(name, value).expect("unknown option");
                self
            }
            /// f at the last evaluate_x_k
            pub fn f(&self) -> Floating {
                let mut f = 0.0;
                unsafe { ffi::osb_solver_f(self.core.raw, &mut f) };
                f
            }
        }
        impl ComputeDirection for $name {
            /// Template hook of the host loop (ls_solver.rs:3-8).  `minimize` is overridden and computes the direction on the
            /// device from the device-resident state; asking for it for an arbitrary host evaluation is not a device operation.
            fn compute_direction(&mut self, _eval_x_k: &FuncEvalMultivariate) -> Result<DVector<Floating>, SolverError> {
                Err(SolverError::AbnormalTermination)
            }
        }
        impl LineSearchSolver for $name {
            fn xk(&self) -> &DVector<Floating> {
                &self.core.x
            }
            fn xk_mut(&mut self) -> &mut DVector<Floating> {
                self.core.x_dirty = true;
                &mut self.core.x
            }
            fn k(&self) -> &usize {
                &self.core.k
            }
            fn k_mut(&mut self) -> &mut usize {
                &mut self.core.k
            }
            fn has_converged(&self, eval_x_k: &FuncEvalMultivariate) -> bool {
                let f: fn(&$name, &FuncEvalMultivariate) -> bool = $conv;
                f(self, eval_x_k)
            }
            fn minimize<LS: LineSearch>(&mut self, line_search: &mut LS, oracle: impl FnMut(&DVector<Floating>) -> FuncEvalMultivariate,
                                        max_iter_solver: usize, max_iter_line_search: usize, callback: Option<&mut dyn FnMut(&Self)>)
                                        -> Result<(), SolverError> {
                GpuSolverCore::minimize(self, line_search, oracle, max_iter_solver, max_iter_line_search, callback)
            }
        }
    };
}
macro_rules! gpu_bounds {
    ($name:ident) => {
        impl HasBounds for $name {
            fn lower_bound(&self) -> &DVector<Floating> {
                &self.core.lower_bound
            }
            fn upper_bound(&self) -> &DVector<Floating> {
                &self.core.upper_bound
            }
            /// (bounds are fixed at construction on the device: rebuild the solver to change them)
            fn set_lower_bound(&mut self, _lower_bound: DVector<Floating>) {
                panic!("gpu backend: bounds are fixed at construction");
            }
            fn set_upper_bound(&mut self, _upper_bound: DVector<Floating>) {
                panic!("gpu backend: bounds are fixed at construction");
            }
        }
        impl $name {
            /// one byte per coordinate: bit 0 = (x_i == lb_i), bit 1 = (x_i == ub_i) — the active set of ls_solver.rs:121-133
            pub fn active_set(&self) -> Vec<u8> {
                let mut out = vec![0u8; self.core.x.len()];
                unsafe { ffi::osb_solver_active_set(self.core.raw, out.as_mut_ptr()) };
                out
            }
        }
    };
}

// ---- has_converged of the reference, on host mirrors (the device applies the same tests inside minimize) ------------
fn qn_converged(core: &GpuSolverCore, eval: &FuncEvalMultivariate) -> bool {
    // bfgs.rs:64-76 (and siblings)
    if let Some(s) = core.s_norm() {
        if s < core.tol {
            return true;
        }
    }
    if let Some(y) = core.y_norm() {
        if y < core.tol {
            return true;
        }
    }
    eval.g().norm() < core.tol
}
fn inf_norm_ignoring_nan(v: &DVector<Floating>, from: Floating) -> Floating {
    v.iter().fold(from, |acc, x| x.abs().max(acc))
}
fn projected_gradient_inf_norm(core: &GpuSolverCore, eval: &FuncEvalMultivariate) -> Floating {
    // ls_solver.rs:121-133 + number.rs:27-31
    let mut pg = eval.g().clone();
    for (i, x) in core.x.iter().enumerate() {
        if (x == &core.lower_bound[i] && pg[i] > 0.0) || (x == &core.upper_bound[i] && pg[i] < 0.0) {
            pg[i] = 0.0;
        }
    }
    inf_norm_ignoring_nan(&pg, 0.0)
}

macro_rules! quasi_newton {
    ($name:ident, $kind:expr, $doc:expr) => {
        #[doc = $doc]
        pub struct $name {
            core: GpuSolverCore,
        }
        impl $name {
            pub fn new(tol: Floating, x0: DVector<Floating>) -> Self {
                Self { core: GpuSolverCore::create(default_context(), $kind, tol, x0, None, None) }
            }
            pub fn new_in(ctx: Rc<GpuContext>, tol: Floating, x0: DVector<Floating>) -> Self {
                Self { core: GpuSolverCore::create(ctx, $kind, tol, x0, None, None) }
            }
            pub fn tol(&self) -> &Floating {
                &self.core.tol
            }
            pub fn s_norm(&self) -> Option<Floating> {
                self.core.s_norm()
            }
            pub fn y_norm(&self) -> Option<Floating> {
                self.core.y_norm()
            }
            /// the dense inverse-Hessian approximation, downloaded (n x n; 2 GiB at n = 16384)
            pub fn approx_inv_hessian(&self) -> DMatrix<Floating> {
                self.core.matrix()
            }
        }
        gpu_solver!($name, $kind, needs_hessian = false, converged = |s, e| qn_converged(&s.core, e));
    };
}
macro_rules! quasi_newton_bounded {
    ($name:ident, $kind:expr, $doc:expr) => {
        #[doc = $doc]
        pub struct $name {
            core: GpuSolverCore,
        }
        impl $name {
            pub fn new(tol: Floating, x0: DVector<Floating>, lower_bound: DVector<Floating>, upper_bound: DVector<Floating>) -> Self {
                Self { core: GpuSolverCore::create(default_context(), $kind, tol, x0, Some((lower_bound, upper_bound)), None) }
            }
            pub fn tol(&self) -> &Floating {
                &self.core.tol
            }
            pub fn s_norm(&self) -> Option<Floating> {
                self.core.s_norm()
            }
            pub fn y_norm(&self) -> Option<Floating> {
                self.core.y_norm()
            }
            pub fn approx_inv_hessian(&self) -> DMatrix<Floating> {
                self.core.matrix()
            }
        }
        gpu_solver!($name, $kind, needs_hessian = false, converged = |s, e| qn_converged(&s.core, e));
        gpu_bounds!($name);
    };
}

quasi_newton!(BFGS, ffi::OSB_BFGS, "`BFGS` (src/quasi_newton/bfgs.rs): device-resident H, packed lower triangle, lazy rank-2 schedule");
quasi_newton!(DFP, ffi::OSB_DFP, "`DFP` (src/quasi_newton/dfp.rs)");
quasi_newton!(Broyden, ffi::OSB_BROYDEN, "`Broyden` (src/quasi_newton/broyden.rs): non-symmetric H, full storage, eager schedule");
quasi_newton_bounded!(BFGSB, ffi::OSB_BFGSB, "`BFGSB` (src/quasi_newton/bfgs_b.rs)");
quasi_newton_bounded!(DFPB, ffi::OSB_DFPB, "`DFPB` (src/quasi_newton/dfp_b.rs)");
quasi_newton_bounded!(BroydenB, ffi::OSB_BROYDENB, "`BroydenB` (src/quasi_newton/broyden_b.rs)");
quasi_newton_bounded!(SR1B, ffi::OSB_SR1B, "`SR1B` (src/quasi_newton/sr1_b.rs)");

/// `GradientDescent` (src/steepest_descent/gradient_descent.rs)
pub struct GradientDescent {
    core: GpuSolverCore,
}
impl GradientDescent {
    pub fn new(grad_tol: Floating, x0: DVector<Floating>) -> Self {
        Self { core: GpuSolverCore::create(default_context(), ffi::OSB_GD, grad_tol, x0, None, None) }
    }
    pub fn new_in(ctx: Rc<GpuContext>, grad_tol: Floating, x0: DVector<Floating>) -> Self {
        Self { core: GpuSolverCore::create(ctx, ffi::OSB_GD, grad_tol, x0, None, None) }
    }
    pub fn grad_tol(&self) -> &Floating {
        &self.core.tol
    }
}
// gradient_descent.rs:46-53: max |g_i|, folded from -inf with f64::max (NaNs ignored)
gpu_solver!(GradientDescent, ffi::OSB_GD, needs_hessian = false, converged = |s, e| inf_norm_ignoring_nan(e.g(), Floating::NEG_INFINITY) < s.core.tol);

/// `PnormDescent` (src/steepest_descent/pnorm_descent.rs): d = -(inverse_p g)
pub struct PnormDescent {
    core: GpuSolverCore,
}
impl PnormDescent {
    pub fn new(grad_tol: Floating, x0: DVector<Floating>, inverse_p: DMatrix<Floating>) -> Self {
        let n = x0.len();
        let core = GpuSolverCore::create(default_context(), ffi::OSB_PNORM, grad_tol, x0, None, None);
        let row_major: Vec<f64> = (0..n).flat_map(|i| (0..n).map(move |j| (i, j))).map(|(i, j)| inverse_p[(i, j)]).collect();
        if unsafe { ffi::osb_solver_set_inv_hessian(core.raw, row_major.as_ptr()) } != ffi::OSB_OK {
            panic!("PnormDescent: {}", last_error());
        }
        Self { core }
    }
    pub fn grad_tol(&self) -> &Floating {
        &self.core.tol
    }
    pub fn inverse_p(&self) -> DMatrix<Floating> {
        self.core.matrix()
    }
}
gpu_solver!(PnormDescent, ffi::OSB_PNORM, needs_hessian = false, converged = |s, e| inf_norm_ignoring_nan(e.g(), Floating::NEG_INFINITY) < s.core.tol);

/// `ProjectedGradientDescent` (src/steepest_descent/projected_gradient_descent.rs)
pub struct ProjectedGradientDescent {
    core: GpuSolverCore,
}
impl ProjectedGradientDescent {
    pub fn new(grad_tol: Floating, x0: DVector<Floating>, lower_bound: DVector<Floating>, upper_bound: DVector<Floating>) -> Self {
        Self { core: GpuSolverCore::create(default_context(), ffi::OSB_PGD, grad_tol, x0, Some((lower_bound, upper_bound)), None) }
    }
    pub fn new_in(ctx: Rc<GpuContext>, grad_tol: Floating, x0: DVector<Floating>, lower_bound: DVector<Floating>, upper_bound: DVector<Floating>) -> Self {
        Self { core: GpuSolverCore::create(ctx, ffi::OSB_PGD, grad_tol, x0, Some((lower_bound, upper_bound)), None) }
    }
    pub fn grad_tol(&self) -> &Floating {
        &self.core.tol
    }
}
gpu_solver!(ProjectedGradientDescent, ffi::OSB_PGD, needs_hessian = false, converged = |s, e| projected_gradient_inf_norm(&s.core, e) < s.core.tol);
gpu_bounds!(ProjectedGradientDescent);

/// The constructors of SPG / SPN call the oracle once for lambda_0 (spg.rs:40-46).  The reference takes
/// `&mut impl FnMut(..)`; a device objective's `.oracle()` closure is recognised the same way as in `minimize`, any
/// other closure goes through the host wrapper.
fn lambda0_oracle<R>(ctx: &Rc<GpuContext>, x0: &DVector<Floating>, with_hessian: bool,
                     oracle: &mut impl FnMut(&DVector<Floating>) -> FuncEvalMultivariate, build: impl FnOnce(*mut ffi::osb_objective) -> R) -> R {
    LAST_DEVICE_CALL.with(|c| c.set((ptr::null_mut(), ptr::null())));
    let _ = oracle(x0);
    let (noted, at) = LAST_DEVICE_CALL.with(|c| c.get());
    if !noted.is_null() && at == x0.as_ptr() {
        return build(noted);
    }
    let mut dynref: &mut dyn FnMut(&DVector<Floating>) -> FuncEvalMultivariate = oracle;
    let wrapper = HostClosure::new(ctx, x0.len(), &mut dynref, with_hessian).expect("host oracle wrapper");
    build(wrapper.raw)
}

/// `SpectralProjectedGradient` (src/steepest_descent/spg.rs)
pub struct SpectralProjectedGradient {
    core: GpuSolverCore,
}
impl SpectralProjectedGradient {
    pub fn new(grad_tol: Floating, x0: DVector<Floating>, oracle: &mut impl FnMut(&DVector<Floating>) -> FuncEvalMultivariate,
               lower_bound: DVector<Floating>, upper_bound: DVector<Floating>) -> Self {
        let ctx = default_context();
        let core = lambda0_oracle(&ctx, &x0.clone(), false, oracle, |obj| {
            GpuSolverCore::create(ctx.clone(), ffi::OSB_SPG, grad_tol, x0, Some((lower_bound, upper_bound)), Some(obj))
        });
        Self { core }
    }
    pub fn with_lambdas(self, lambda_min: Floating, lambda_max: Floating) -> Self {
        unsafe { ffi::osb_solver_set_lambdas(self.core.raw, lambda_min, lambda_max) };
        self
    }
    pub fn grad_tol(&self) -> &Floating {
        &self.core.tol
    }
    pub fn lambda(&self) -> Floating {
        unsafe { ffi::osb_solver_lambda(self.core.raw) }
    }
}
gpu_solver!(SpectralProjectedGradient, ffi::OSB_SPG, needs_hessian = false, converged = |s, e| projected_gradient_inf_norm(&s.core, e) < s.core.tol);
gpu_bounds!(SpectralProjectedGradient);

/// `Newton` (src/newton/mod.rs): DMMA Hessian assembly for the logistic objective, blocked Cholesky, LU fallback
pub struct Newton {
    core: GpuSolverCore,
}
impl Newton {
    pub fn new(tol: Floating, x0: DVector<Floating>) -> Self {
        Self { core: GpuSolverCore::create(default_context(), ffi::OSB_NEWTON, tol, x0, None, None) }
    }
    pub fn new_in(ctx: Rc<GpuContext>, tol: Floating, x0: DVector<Floating>) -> Self {
        Self { core: GpuSolverCore::create(ctx, ffi::OSB_NEWTON, tol, x0, None, None) }
    }
    pub fn tol(&self) -> &Floating {
        &self.core.tol
    }
    pub fn decrement_squared(&self) -> Option<Floating> {
        GpuSolverCore::opt(unsafe { ffi::osb_solver_decrement_squared(self.core.raw) })
    }
}
// newton/mod.rs:64-69: ignores the eval, tests the decrement stored by the previous compute_direction
gpu_solver!(Newton, ffi::OSB_NEWTON, needs_hessian = true, converged = |s, _e| match s.decrement_squared() {
    Some(d) => d * 0.5 < s.core.tol,
    None => false,
});

/// `ProjectedNewton` (src/newton/projected_newton.rs)
pub struct ProjectedNewton {
    core: GpuSolverCore,
}
impl ProjectedNewton {
    pub fn new(grad_tol: Floating, x0: DVector<Floating>, lower_bound: DVector<Floating>, upper_bound: DVector<Floating>) -> Self {
        Self { core: GpuSolverCore::create(default_context(), ffi::OSB_PROJ_NEWTON, grad_tol, x0, Some((lower_bound, upper_bound)), None) }
    }
    pub fn grad_tol(&self) -> &Floating {
        &self.core.tol
    }
    pub fn s_norm(&self) -> Option<Floating> {
        self.core.s_norm()
    }
    pub fn y_norm(&self) -> Option<Floating> {
        self.core.y_norm()
    }
}
// projected_newton.rs:95-110
gpu_solver!(ProjectedNewton, ffi::OSB_PROJ_NEWTON, needs_hessian = true, converged = |s, e| {
    if let Some(v) = s.core.s_norm() {
        if v < s.core.tol {
            return true;
        }
    }
    if let Some(v) = s.core.y_norm() {
        if v < s.core.tol {
            return true;
        }
    }
    projected_gradient_inf_norm(&s.core, e) < s.core.tol
});
gpu_bounds!(ProjectedNewton);

/// `SpectralProjectedNewton` (src/newton/spn.rs)
pub struct SpectralProjectedNewton {
    core: GpuSolverCore,
}
impl SpectralProjectedNewton {
    pub fn new(grad_tol: Floating, x0: DVector<Floating>, oracle: &mut impl FnMut(&DVector<Floating>) -> FuncEvalMultivariate,
               lower_bound: DVector<Floating>, upper_bound: DVector<Floating>) -> Self {
        let ctx = default_context();
        let core = lambda0_oracle(&ctx, &x0.clone(), true, oracle, |obj| {
            GpuSolverCore::create(ctx.clone(), ffi::OSB_SPN, grad_tol, x0, Some((lower_bound, upper_bound)), Some(obj))
        });
        Self { core }
    }
    pub fn with_lambdas(self, lambda_min: Floating, lambda_max: Floating) -> Self {
        unsafe { ffi::osb_solver_set_lambdas(self.core.raw, lambda_min, lambda_max) };
        self
    }
    pub fn grad_tol(&self) -> &Floating {
        &self.core.tol
    }
    pub fn lambda(&self) -> Floating {
        unsafe { ffi::osb_solver_lambda(self.core.raw) }
    }
}
gpu_solver!(SpectralProjectedNewton, ffi::OSB_SPN, needs_hessian = true, converged = |s, e| projected_gradient_inf_norm(&s.core, e) < s.core.tol);
gpu_bounds!(SpectralProjectedNewton);

// ---------------------------------------------------------------------------------------------------------------------
// batched mode: many small independent problems, one launch
// ---------------------------------------------------------------------------------------------------------------------
pub struct BatchedResult {
    pub x: Vec<Floating>,  // n_problems * n, row-major
    pub f: Vec<Floating>,
    pub k: Vec<i32>,
    pub status: Vec<i32>,
    pub reason: Vec<i32>,
    pub device_ms: f64,
}
/// BFGS + BackTracking on `n_problems` extended-Rosenbrock problems of dimension n, x0 generated on the device
/// (`problem0` offsets the generator: split the problems over the GPUs on the caller side, no collective).
pub fn batched_bfgs_rosenbrock(ctx: &GpuContext, n: usize, n_problems: usize, problem0: usize, tol: Floating, max_iter_solver: usize,
                               max_iter_line_search: usize, c1: Floating, beta: Floating) -> Result<BatchedResult, SolverError> {
    let mut r = BatchedResult {
        x: vec![0.0; n * n_problems], f: vec![0.0; n_problems], k: vec![0; n_problems], status: vec![0; n_problems], reason: vec![0; n_problems],
        device_ms: 0.0,
    };
    check(unsafe {
        ffi::osb_batched_bfgs_rosenbrock_generated(ctx.raw, n as i64, n_problems as i64, problem0 as i64, tol, max_iter_solver as i64,
                                                   max_iter_line_search as i64, c1, beta, r.x.as_mut_ptr(), r.f.as_mut_ptr(), r.k.as_mut_ptr(),
                                                   r.status.as_mut_ptr(), r.reason.as_mut_ptr(), &mut r.device_ms)
    })?;
    Ok(r)
}

// ---------------------------------------------------------------------------------------------------------------------
fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::osb_last_error_string()) }.to_string_lossy().into_owned()
}
fn check(rc: c_int) -> Result<(), SolverError> {
    match rc {
        ffi::OSB_OK => Ok(()),
        ffi::OSB_MAX_ITER_REACHED => Err(SolverError::MaxIterReached),
        ffi::OSB_OUT_OF_DOMAIN => Err(SolverError::OutOfDomain),
        ffi::OSB_ERROR_INPUT_PARAMS => Err(SolverError::ErrorInputParams),
        // the reference panics here: `.expect("Hessian not available in the oracle")` (newton/mod.rs:34),
        // `.cholesky().unwrap()` (projected_newton.rs:75, spn.rs:86)
        ffi::OSB_PANIC_NO_HESSIAN | ffi::OSB_PANIC_NOT_SPD => panic!("{}", last_error()),
        _ => {
            error!(target: "solver", "gpu backend: {}", last_error());
            Err(SolverError::AbnormalTermination)
        }
    }
}
