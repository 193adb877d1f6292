import sys; sys.path.insert(0, '.')
import importlib
osb = importlib.import_module("optimization-solvers_b200")
n = 16384
for name, v in (("register-staged", 0), ("register-staged, column partials in shared memory", 128), ("ring", 8), ("ring, no column partials", 8 | 16), ("ring, no stores", 8 | 32), ("ring, copy only (no math, no reduction)", 8 | 64),
                ("ring, copy only, no colpart", 8 | 80), ("ring, nothing but loads", 8 | 112), ("register-staged again", 0), ("register-staged, smem column partials again", 128), ("d2d copy of 2 GiB (x0.5)", -1)):
    if v < 0:
        ms = osb.bench_qn_kernel(3, n, 20, 0) * 0.5
    else:
        ms = osb.bench_qn_kernel(4, n, 50, v)
    print("%-55s %.4f ms  %.0f GB/s" % (name, ms, n * n * 8 / ms / 1e6), flush=True)
