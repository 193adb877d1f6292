import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
osb = importlib.import_module("optimization-solvers_b200")
from bench import rosen_x0
n = 16384
ctx = osb.default_context()
obj = osb.ExtendedRosenbrock(n)
x0 = rosen_x0(n, 0)
for rep in range(2):
    t0 = time.perf_counter()
    s = osb.BFGS(1e-8, x0).set_option("qn_schedule", 1)
    ctx.synchronize()
    t1 = time.perf_counter()
    ts = []
    def cb(sv):
        a = time.perf_counter()
        v = sv.x()[0]
        ts.append(time.perf_counter() - a)
    try:
        s.minimize(osb.BackTracking(1e-4, 0.5), obj, 200, 20, callback=cb)
    except osb.MaxIterReached:
        pass
    t2 = time.perf_counter()
    xf = s.x(); ff = s.f()
    t3 = time.perf_counter()
    print("construct %.2f ms, minimize(200, cb) %.2f ms (%.3f ms/it), cb x() mean %.1f us, final %.2f ms, device ms %.2f" %
          ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t1) * 5, np.mean(ts) * 1e6, (t3 - t2) * 1e3, s.last_timing()[0]))
    try:
        s2 = osb.BFGS(1e-8, x0).set_option("qn_schedule", 1)
        t4 = time.perf_counter()
        s2.minimize(osb.BackTracking(1e-4, 0.5), obj, 200, 20, callback=lambda sv: None)
    except osb.MaxIterReached:
        print("minimize(200, empty cb) %.2f ms" % ((time.perf_counter() - t4) * 1e3))
    s.close(); s2.close()
