"""Where the fixed cost of a short solve goes: construction, first minimize (allocations), later minimize calls."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
osb = importlib.import_module("optimization-solvers_b200")
from bench import rosen_x0
n = 16384
x0 = rosen_x0(n, 0)
obj = osb.ExtendedRosenbrock(n)
ctx = osb.default_context()
def T():
    ctx.synchronize(); return time.perf_counter()
for rep in range(3):
    for storage, ra in ((1, 1), (1, 0), (0, 0)):
        t0 = T()
        s = osb.BFGS(1e-8, x0).set_option("engine", 2).set_option("qn_schedule", 1).set_option("qn_storage", storage).set_option("callback_run_ahead", ra)
        t1 = T()
        seen = []
        def cb(sv): seen.append(sv.f())
        ts = [t1]
        for iters in (1, 1, 20, 20):
            try: s.minimize(osb.BackTracking(1e-4, 0.5), obj, iters, 20, callback=cb)
            except osb.MaxIterReached: pass
            ts.append(T())
        s.close(); t_close = T()
        print("rep %d storage=%d run_ahead=%d construct %.2f ms | minimize(1) %.2f | minimize(1) %.2f | minimize(20) %.2f | minimize(20) %.2f | close %.2f"
              % (rep, storage, ra, (t1 - t0) * 1e3, *[(ts[i + 1] - ts[i]) * 1e3 for i in range(4)], (t_close - ts[-1]) * 1e3), flush=True)
