import importlib, ctypes as C, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
osb=importlib.import_module("optimization-solvers_b200")
from bench import rosen_x0
n=16384
s=osb.BFGS(1e-8, rosen_x0(n,0)).set_option("engine",2).set_option("qn_schedule",1).set_option("head_debug",1).set_option("qn_storage", int(os.environ.get("SYM","1")))
obj=osb.ExtendedRosenbrock(n)
L=osb.lib(); L.osb_debug_head_stamps.argtypes=[C.POINTER(C.c_longlong)]
for it in (20, 1, 1, 1, 60, 1, 1):
    try: s.minimize(osb.BackTracking(1e-4,0.5), obj, it, 20)
    except osb.MaxIterReached: pass
    out=(C.c_longlong*32)(); L.osb_debug_head_stamps(out)
    v=list(out); nst=v[31]
    d=[v[i+1]-v[i] for i in range(nst-1)]
    print("iters",it,"stamps",nst,"total cycles",v[nst-1]-v[0],"deltas",d, "ls_evals", osb.default_context().counters()["ls_trials"])
