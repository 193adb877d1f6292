//! The crate-side half of the integration: what `src/line_search/*.rs` gains under `--features gpu`.
//!
//! 1. `src/line_search/mod.rs` — one provided method on the trait (default: no device form):
//!
//! ```ignore
//! pub trait LineSearch {
//!     fn compute_step_len(&mut self, ...) -> Floating;                       // unchanged (line_search/mod.rs:14-23)
//!     #[cfg(feature = "gpu")]
//!     fn gpu_spec(&self) -> Option<crate::gpu::LineSearchSpec> { None }      // NEW
//! }
//! ```
//!
//! 2. the six searches of the crate describe themselves (they own their private fields); this file is included from
//!    `src/line_search/mod.rs` with `#[cfg(feature = "gpu")] mod gpu_spec;` and adds the method through the same trait
//!    impls by the usual pattern below (shown for each struct; in the crate these bodies go INTO the existing
//!    `impl LineSearch for X` blocks).
use super::*;
use crate::gpu::LineSearchSpec;

impl BackTracking {
    pub(crate) fn gpu_spec_impl(&self) -> Option<LineSearchSpec> {
        Some(LineSearchSpec::BackTracking { c1: self.c1, beta: self.beta })  // backtracking.rs:3-6
    }
}
impl BackTrackingB {
    pub(crate) fn gpu_spec_impl(&self) -> Option<LineSearchSpec> {
        Some(LineSearchSpec::BackTrackingB {
            c1: self.c1, beta: self.beta, lower_bound: self.lower_bound.clone(), upper_bound: self.upper_bound.clone(),  // backtracking_b.rs:4-9
        })
    }
}
impl MoreThuente {
    pub(crate) fn gpu_spec_impl(&self) -> Option<LineSearchSpec> {
        Some(LineSearchSpec::MoreThuente {  // morethuente.rs:6-14
            c1: self.c1, c2: self.c2, t_min: self.t_min, t_max: self.t_max, delta_min: self.delta_min, delta: self.delta, delta_max: self.delta_max,
        })
    }
}
impl MoreThuenteB {
    pub(crate) fn gpu_spec_impl(&self) -> Option<LineSearchSpec> {
        Some(LineSearchSpec::MoreThuenteB {  // morethuente_b.rs:6-16
            c1: self.c1, c2: self.c2, t_min: self.t_min, t_max: self.t_max, delta_min: self.delta_min, delta: self.delta, delta_max: self.delta_max,
            lower_bound: self.lower_bound.clone(), upper_bound: self.upper_bound.clone(),
        })
    }
}
impl GLLQuadratic {
    pub(crate) fn gpu_spec_impl(&self) -> Option<LineSearchSpec> {
        Some(LineSearchSpec::GLLQuadratic { c1: self.c1, m: self.m, sigma1: self.sigma1, sigma2: self.sigma2 })  // gll_quadratic.rs:3-10
    }
}
impl NoSearch {
    pub(crate) fn gpu_spec_impl(&self) -> Option<LineSearchSpec> {
        Some(LineSearchSpec::NoSearch)  // nosearch.rs:3-15
    }
}
// in each `impl LineSearch for X { ... }` of the crate:
//     #[cfg(feature = "gpu")]
//     fn gpu_spec(&self) -> Option<crate::gpu::LineSearchSpec> { self.gpu_spec_impl() }
//
// State that the reference keeps inside the line-search object across outer iterations (MoreThuenteB.t_max shrinks
// permanently, morethuente_b.rs:201; GLLQuadratic.f_previous, gll_quadratic.rs:30-43) lives in the device handle the
// solver caches between `minimize` calls (gpu/mod.rs: GpuSolverCore::ls_cache).
