"""Small workloads for ncu captures of the four dominant kernels (one per argument):
  sym    - default BFGS path, n = 16384, 12 iterations  (qn_lazy_sym_kernel, qn_sym_fold_kernel, cluster head)
  fused  - the same through the fused iteration kernel, 3 launches of 4 iterations (qn_iter_kernel)
  stream - SPG on the generated separable quadratic, n = 2^26, 6 iterations (stream_trial kernel)
  syrk   - logistic Hessian, m = 32768, n = 8192 (syrk_dmma_kernel, syrk_reduce_kernel)
  batched - 32768 BFGS solves of extended Rosenbrock, n = 32 (batched_bfgs_kernel)"""
import importlib, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
osb = importlib.import_module("optimization-solvers_b200")
from test_gpu_parity import rosen_x0
what = sys.argv[1]
if what in ("sym", "fused"):
    n = 16384
    s = osb.BFGS(1e-30, rosen_x0(n, 3))
    if what == "fused":
        s.set_option("fused_iteration", 1)
    for _ in range(3 if what == "fused" else 1):
        try:
            s.minimize(osb.BackTracking(1e-4, 0.5), osb.ExtendedRosenbrock(n), 4 if what == "fused" else 12, 40)
        except osb.MaxIterReached:
            pass
    print(s.path_info()["kernel"], s.k())
elif what == "stream":
    n = 1 << 26
    obj = osb.SeparableQuadratic.generated(n)
    s = osb.SpectralProjectedGradient(1e-30, np.zeros(n), obj, np.full(n, -1.0), np.full(n, 1.0))
    try:
        s.minimize(osb.GLLQuadratic(1e-4, 10), obj, 6, 20)
    except osb.MaxIterReached:
        pass
    print(s.path_info()["fused_stream"], s.k())
elif what == "batched":
    r = osb.batched_bfgs_rosenbrock(32, 32768)
    print(r["ms"] if isinstance(r, dict) else r)
else:
    obj = osb.LogisticRegression.generated(32768, 8192, 1.0)
    print(osb.bench_syrk(obj, 1))
