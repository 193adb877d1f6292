import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
torch.cuda.set_device(0)
osb = importlib.import_module("optimization-solvers_b200")
from bench import rosen_x0, ClockSampler
n = 16384
ctx = osb.Context(0)
osb.set_default_context(ctx)
obj = osb.ExtendedRosenbrock(n, ctx=ctx)
x0 = rosen_x0(n, 0)
main = osb.BFGS(1e-8, x0, ctx=ctx).set_option("engine", 2).set_option("qn_schedule", 1)
def steps(s, k, cb=None):
    t = time.perf_counter()
    try: s.minimize(osb.BackTracking(1e-4, 0.5), obj, k, 20, callback=cb)
    except osb.MaxIterReached: pass
    return (time.perf_counter() - t) * 1e3
print("main 200:", steps(main, 200))
sam = ClockSampler(0); sam.start(); time.sleep(0.3); sam.stop_flag = True; sam.join(timeout=2)
print("sampler stopped, alive:", sam.is_alive())
main.set_option("profile_kernels", 1); print("main prof 200:", steps(main, 200), main.kernel_timing()); main.set_option("profile_kernels", 0)
x0p = torch.from_numpy(x0).pin_memory()
for rep in range(3):
    xs = []
    t0 = time.perf_counter()
    s2 = osb.BFGS(1e-8, x0p.numpy(), ctx=ctx).set_option("qn_schedule", 1)
    ms = steps(s2, 200, cb=lambda s: xs.append(s.x()[0]))
    xf = s2.x(); ff = s2.f(); ctx.synchronize()
    print("e2e total %.1f ms, minimize %.1f ms, device %.1f ms" % ((time.perf_counter() - t0) * 1e3, ms, s2.last_timing()[0]))
    s2.close()
