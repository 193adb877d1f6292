"""Times the DMMA Hessian assembly (X^T D X, n = 8192) of the logistic objective: 8 warps (32 x 64 per warp) against 16 warps
(32 x 32 per warp) per 128 x 128 tile; negative reps select the 16-warp kernel (osb_bench_syrk)."""
import sys; sys.path.insert(0, '.')
import importlib
osb = importlib.import_module("optimization-solvers_b200")
m, n = 262144, 8192
obj = osb.LogisticRegression.generated(m, n, 1.0)
for name, reps in (("8 warps", 3), ("16 warps", -3), ("8 warps", 3), ("16 warps", -3)):
    ms = osb.bench_syrk(obj, reps)
    print("%-10s %.2f ms  %.2f TFLOP/s" % (name, ms, m * n * n / ms / 1e9), flush=True)
