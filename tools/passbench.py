"""Times the lazy pass over the packed triangle alone (osb_bench_qn_kernel which = 4) for each selectable variant, next to a
device-to-device copy of the same number of bytes.  The experiment variants measured in round 2 (no column partials, no
stores, copy only, column partials in shared memory) lived in commit 1fc1bd8; their output is profiles/r02_packed_pass_experiments.md."""
import sys; sys.path.insert(0, '.')
import importlib
osb = importlib.import_module("optimization-solvers_b200")
n = 16384
for name, v in (("register-staged (default)", 0), ("2 x 256 threads", 1), ("ping-pong storage", 2), ("zero-first column partials", 4),
                ("shared-memory ring (cp.async.bulk)", 8), ("register-staged again", 0), ("d2d copy of 2 GiB (x0.5)", -1)):
    if v < 0:
        ms = osb.bench_qn_kernel(3, n, 20, 0) * 0.5
    else:
        ms = osb.bench_qn_kernel(4, n, 50, v)
    print("%-45s %.4f ms  %.0f GB/s" % (name, ms, n * n * 8 / ms / 1e6), flush=True)
