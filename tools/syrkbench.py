"""Times the DMMA Hessian assembly (X^T D X, n = 8192) of the logistic objective (osb_bench_syrk).  Stage shapes measured
in round 2 at m = 262144 (gpurun_out/r3d_syrk.log): 16 samples x 4 stages 28.5, 32 x 3 29.8, 48 x 2 30.6 TFLOP/s."""
import sys; sys.path.insert(0, '.')
import importlib
osb = importlib.import_module("optimization-solvers_b200")
m, n = 262144, 8192
obj = osb.LogisticRegression.generated(m, n, 1.0)
for _ in range(2):
    ms = osb.bench_syrk(obj, 3)
    print("%.2f ms  %.2f TFLOP/s" % (ms, m * n * n / ms / 1e9), flush=True)
