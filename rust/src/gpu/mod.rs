//! `optimization_solvers::gpu` — the B200 backend behind the crate's own traits.
//!
//! Add to `src/lib.rs`:  `#[cfg(feature = "gpu")] pub mod gpu;`   and to Cargo.toml:
//! `[features] gpu = []`, plus a `build.rs` printing `cargo:rustc-link-search=<dir of libosb_b200.so>`.
//!
//! The solver structs keep the reference's constructors and getters (`BFGS::new(tol, x0)`, `x()`, `k()`,
//! `s_norm()`, ...).  `LineSearchSolver::minimize` is overridden ("Methods that are already implemented can be
//! freely overriden", src/ls_solver.rs:22) to dispatch to `osb_minimize`; iterate, gradient and the dense
//! inverse-Hessian state stay on the device, and a host mirror of `x` is refreshed lazily for `xk()`, the callback
//! and `debug!` output.  Objectives are device functors (`DeviceObjective`); a plain Rust closure is still accepted
//! through `HostClosure` (H2D/D2H per call — functional, not fast).
pub mod ffi;

use crate::{ComputeDirection, FuncEvalMultivariate, LineSearch, LineSearchSolver, SolverError};
use nalgebra::DVector;
use std::os::raw::{c_int, c_void};
use std::ptr;

pub struct GpuContext { raw: *mut ffi::osb_ctx }
impl GpuContext {
    pub fn new(device: i32) -> Result<Self, SolverError> {
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::osb_ctx_create(device, &mut raw) })?;
        Ok(Self { raw })
    }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { ffi::osb_ctx_destroy(self.raw) } } }

/// Device counterpart of `FnMut(&DVector<f64>) -> FuncEvalMultivariate` (src/ls_solver.rs:34).
pub trait DeviceObjective { fn raw(&self) -> *mut ffi::osb_objective; }

pub struct ExtendedRosenbrock { raw: *mut ffi::osb_objective }
impl ExtendedRosenbrock {
    pub fn new(ctx: &GpuContext, n: usize) -> Result<Self, SolverError> {
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::osb_objective_create_rosenbrock(ctx.raw, n as i64, &mut raw) })?;
        Ok(Self { raw })
    }
}
impl DeviceObjective for ExtendedRosenbrock { fn raw(&self) -> *mut ffi::osb_objective { self.raw } }
impl Drop for ExtendedRosenbrock { fn drop(&mut self) { unsafe { ffi::osb_objective_destroy(self.raw) } } }

/// A host closure wrapped as an objective (compatibility path).
pub struct HostClosure<F: FnMut(&DVector<f64>) -> FuncEvalMultivariate> { raw: *mut ffi::osb_objective, _f: Box<F> }
unsafe extern "C" fn host_trampoline<F: FnMut(&DVector<f64>) -> FuncEvalMultivariate>(
    user: *mut c_void, x: *const f64, n: i64, f: *mut f64, g: *mut f64, hess: *mut f64) -> c_int {
    let closure = &mut *(user as *mut F);
    let xv = DVector::from_column_slice(std::slice::from_raw_parts(x, n as usize));
    let eval = closure(&xv);
    *f = *eval.f();
    std::slice::from_raw_parts_mut(g, n as usize).copy_from_slice(eval.g().as_slice());
    match (hess.is_null(), eval.hessian()) {
        (false, Some(h)) => {
            // row-major out (nalgebra is column-major; Hessians are symmetric)
            let out = std::slice::from_raw_parts_mut(hess, (n * n) as usize);
            for i in 0..n as usize { for j in 0..n as usize { out[i * n as usize + j] = h[(i, j)]; } }
            1
        }
        _ => 0,
    }
}
impl<F: FnMut(&DVector<f64>) -> FuncEvalMultivariate> HostClosure<F> {
    pub fn new(ctx: &GpuContext, n: usize, f: F, with_hessian: bool) -> Result<Self, SolverError> {
        let mut boxed = Box::new(f);
        let mut raw = ptr::null_mut();
        check(unsafe {
            ffi::osb_objective_create_host(ctx.raw, n as i64, host_trampoline::<F>, &mut *boxed as *mut F as *mut c_void,
                                           with_hessian as c_int, &mut raw)
        })?;
        Ok(Self { raw, _f: boxed })
    }
}
impl<F: FnMut(&DVector<f64>) -> FuncEvalMultivariate> DeviceObjective for HostClosure<F> {
    fn raw(&self) -> *mut ffi::osb_objective { self.raw }
}

/// Line searches keep their reference constructors; each knows how to build its device handle.
pub trait GpuLineSearch { fn raw(&mut self, ctx: &GpuContext) -> *mut ffi::osb_linesearch; }

/// `BFGS` with device-resident state (src/quasi_newton/bfgs.rs:4-12).
pub struct BFGS { raw: *mut ffi::osb_solver, x: DVector<f64>, k: usize, tol: f64 }
impl BFGS {
    pub fn new(ctx: &GpuContext, tol: f64, x0: DVector<f64>) -> Result<Self, SolverError> {
        let mut raw = ptr::null_mut();
        check(unsafe {
            ffi::osb_solver_create(ctx.raw, ffi::OSB_BFGS, x0.len() as i64, tol, x0.as_ptr(), ptr::null(), ptr::null(),
                                   ptr::null_mut(), &mut raw)
        })?;
        Ok(Self { raw, x: x0, k: 0, tol })
    }
    pub fn x(&mut self) -> &DVector<f64> { unsafe { ffi::osb_solver_x(self.raw, self.x.as_mut_ptr()) }; &self.x }
    pub fn k(&self) -> usize { unsafe { ffi::osb_solver_k(self.raw) as usize } }
    pub fn tol(&self) -> f64 { self.tol }
    pub fn s_norm(&self) -> Option<f64> { opt(unsafe { ffi::osb_solver_s_norm(self.raw) }) }
    pub fn y_norm(&self) -> Option<f64> { opt(unsafe { ffi::osb_solver_y_norm(self.raw) }) }
    /// `LineSearchSolver::minimize` (src/ls_solver.rs:66-111) with a device objective.
    pub fn minimize_device<LS: GpuLineSearch, O: DeviceObjective>(&mut self, ctx: &GpuContext, ls: &mut LS, oracle: &O,
                                                                  max_iter_solver: usize, max_iter_line_search: usize)
                                                                  -> Result<(), SolverError> {
        let rc = unsafe {
            ffi::osb_minimize(self.raw, ls.raw(ctx), oracle.raw(), max_iter_solver as i64, max_iter_line_search as i64,
                              None, ptr::null_mut())
        };
        self.k = self.k();
        check(rc)
    }
}
impl Drop for BFGS { fn drop(&mut self) { unsafe { ffi::osb_solver_destroy(self.raw) } } }

fn opt(v: f64) -> Option<f64> { if v.is_nan() { None } else { Some(v) } }
fn check(rc: c_int) -> Result<(), SolverError> {
    match rc {
        ffi::OSB_OK => Ok(()),
        ffi::OSB_MAX_ITER_REACHED => Err(SolverError::MaxIterReached),
        ffi::OSB_OUT_OF_DOMAIN => Err(SolverError::OutOfDomain),
        ffi::OSB_ERROR_INPUT_PARAMS => Err(SolverError::ErrorInputParams),
        101 | 102 => panic!("{}", unsafe { std::ffi::CStr::from_ptr(ffi::osb_last_error_string()) }.to_string_lossy()),
        _ => Err(SolverError::AbnormalTermination),
    }
}

// The remaining solver structs (DFP, Broyden, BFGSB, DFPB, BroydenB, SR1B, GradientDescent,
// ProjectedGradientDescent, SpectralProjectedGradient, Newton, ProjectedNewton, SpectralProjectedNewton) repeat
// the BFGS wrapper with their `OSB_*` kind and the reference's constructor arguments (bounds, oracle for lambda0).
#[allow(dead_code)]
fn _traits_in_scope(_: &dyn ComputeDirection) {}
#[allow(dead_code)]
fn _traits_in_scope2<T: LineSearchSolver, L: LineSearch>(_: &T, _: &L) {}
