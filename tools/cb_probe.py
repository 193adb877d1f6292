"""Where the per-iteration cost of host callbacks goes (fused iteration kernel): wall clock per iteration of minimize()
for K = 40 and K = 200 (the slope is the steady-state cost) with no callback, a no-op callback and the bench's reading
callback, run-ahead delivery on and off, plus the in-kernel profile of the same runs."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
osb = importlib.import_module("optimization-solvers_b200")
from test_gpu_parity import rosen_x0
n = 16384
x0 = rosen_x0(n, 3)
obj = osb.ExtendedRosenbrock(n)
ctx = osb.default_context()

def go(K, cbkind, run_ahead, fused=1):
    s = osb.BFGS(1e-30, x0).set_option("fused_iteration", fused).set_option("callback_run_ahead", run_ahead).set_option("profile_iter", 1 if fused else 0)
    seen = []
    cb = None if cbkind == "none" else (lambda sv: None) if cbkind == "noop" else (lambda sv: seen.append((sv.x()[0], sv.f())))
    ctx.synchronize()
    t0 = time.perf_counter()
    try:
        s.minimize(osb.BackTracking(1e-4, 0.5), obj, K, 40, callback=cb)
    except osb.MaxIterReached:
        pass
    ctx.synchronize()
    t = time.perf_counter() - t0
    p = s.iter_profile() if fused else None
    s.close()
    return t, p

for fused in (1, 0):
    for cbkind, ra in (("none", 1), ("noop", 1), ("read", 1), ("read", 0)):
        go(20, cbkind, ra, fused)
        t1, p1 = go(40, cbkind, ra, fused)
        t2, p2 = go(200, cbkind, ra, fused)
        print("fused=%d callback=%-5s run_ahead=%d: K=40 %.3f ms/it, K=200 %.3f ms/it, slope %.4f ms/it; in-kernel %s"
              % (fused, cbkind, ra, t1 / 40 * 1e3, t2 / 200 * 1e3, (t2 - t1) / 160 * 1e3,
                 ({k: round(v, 4) for k, v in p2.items() if k.endswith("_ms")} if p2 else "-")), flush=True)
