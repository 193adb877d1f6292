//! The twin of examples/bfgs_example.rs on the gpu backend: same solver name, same constructor, same `minimize` call.
//! The only difference a user sees is where the oracle comes from: a device objective hands out the closure
//! (`.oracle()`), an ordinary Rust closure still works (host path).
//!
//!     cargo run --release --features gpu --example bfgs_gpu
use nalgebra::DVector;
use optimization_solvers::gpu::{DeviceObjective, ExtendedRosenbrock, BFGS};
use optimization_solvers::{BackTracking, FuncEvalMultivariate, LineSearchSolver, LogFormat, MoreThuente, Tracer};

fn main() {
    std::env::set_var("RUST_LOG", "info");
    let _tracer = Tracer::default().with_stdout_layer(Some(LogFormat::Normal)).build();

    // ---- 1. the headline configuration: dense BFGS, extended Rosenbrock, n = 16384, H (1 GiB packed) on the device
    let n = 16384;
    let x0 = DVector::from_fn(n, |i, _| if i % 2 == 0 { -1.2 } else { 1.0 });
    let objective = ExtendedRosenbrock::new(n).expect("CUDA device");
    let mut ls = BackTracking::new(1e-4, 0.5);
    let mut solver = BFGS::new(1e-8, x0);
    let mut iterates = 0usize;
    let mut callback = |s: &BFGS| {
        iterates += 1;
        if *s.k() % 100 == 0 {
            println!("k = {:5}  f = {:.6e}  |s| = {:?}", s.k(), s.f(), s.s_norm());
        }
    };
    let res = solver.minimize(&mut ls, objective.oracle(), 500, 20, Some(&mut callback));
    println!("{:?} after {} iterations ({} callbacks), reason {:?}", res, solver.k(), iterates, solver.termination_reason());

    // ---- 2. examples/bfgs_example.rs verbatim (3-D quadratic, host closure oracle, More-Thuente)
    let oracle = |x: &DVector<f64>| -> FuncEvalMultivariate {
        let (x1, x2, x3) = (x[0], x[1], x[2]);
        let f = x1 * x1 + 2.0 * (x2 * x2) + 3.0 * (x3 * x3) + x1 * x2 + x2 * x3;
        let g = DVector::from_vec(vec![2.0 * x1 + x2, 4.0 * x2 + x1 + x3, 6.0 * x3 + x2]);
        FuncEvalMultivariate::new(f, g)
    };
    let mut ls = MoreThuente::default();
    let mut solver = BFGS::new(1e-8, DVector::from_vec(vec![1.0, 1.0, 1.0]));
    solver.minimize(&mut ls, oracle, 50, 20, None).unwrap();
    println!("bfgs_example: k = {}  x = {:?}", solver.k(), solver.x().as_slice());
}
