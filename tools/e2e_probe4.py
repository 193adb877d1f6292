"""Constructor cost under bench.py's conditions (torch loaded, NVML initialised, pinned x0, a live main solver)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
osb = importlib.import_module("optimization-solvers_b200")
from bench import rosen_x0
n = 16384
x0 = rosen_x0(n, 0)
ctx = osb.default_context()
obj = osb.ExtendedRosenbrock(n)
def T():
    ctx.synchronize(); return time.perf_counter()
def ctor(tag, x):
    for _ in range(4):
        t0 = T(); s = osb.BFGS(1e-8, x); t1 = T(); s.close(); t2 = T()
        print("%-40s ctor %.2f ms close %.2f ms" % (tag, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
ctor("plain", x0)
import torch
torch.cuda.synchronize()
ctor("torch imported + cuda init", x0)
xp = torch.from_numpy(x0).pin_memory()
ctor("pinned x0", xp.numpy())
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0); pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
ctor("nvml initialised", xp.numpy())
main = osb.BFGS(1e-8, x0).set_option("engine", 2).set_option("qn_schedule", 1).set_option("qn_storage", 1)
try: main.minimize(osb.BackTracking(1e-4, 0.5), obj, 10, 20)
except osb.MaxIterReached: pass
ctor("live main solver", xp.numpy())
