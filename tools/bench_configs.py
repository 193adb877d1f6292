"""Measured runs of the other BASELINE.json configs (the headline C3 lives in bench.py).
Prints one JSON object per config; results are copied to profiles/ by hand.

    python tools/bench_configs.py [c2] [c4] [c5a] [c5b] [--m 262144]
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
osb = importlib.import_module("optimization-solvers_b200")
PEAK = 6549.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def run(solver, ls, obj, mi, ml):
    try:
        solver.minimize(ls, obj, mi, ml)
        return "Ok"
    except osb.SolverError as e:
        return type(e).__name__


def c2():
    # GradientDescent + BackTracking(1e-4, 0.5), dense SPD quadratic n = 16384 (GEMV-bound: n^2 * 8 B per oracle call)
    n = 16384
    obj = osb.DenseQuadratic.generated(n, True)
    ctx = osb.default_context()
    s = osb.GradientDescent(1e-6, obj.x0)
    run(s, osb.BackTracking(1e-4, 0.5), obj, 3, 100)  # warm-up
    s = osb.GradientDescent(1e-6, obj.x0)
    c0 = obj.calls()
    st = run(s, osb.BackTracking(1e-4, 0.5), obj, 1000, 100)
    ms, it = s.last_timing()
    evals = obj.calls() - c0
    return {"config": "C2 GD + BackTracking, dense SPD quadratic n=16384", "status": st, "iterations": it, "reason": s.termination_reason(),
            "ms_total": ms, "iterations_per_s": it / ms * 1e3, "oracle_evals": evals, "evals_per_iteration": evals / max(it, 1),
            "bytes_per_eval": n * n * 8, "achieved_GBps": evals * n * n * 8 / (ms * 1e-3) / 1e9,
            "frac_of_measured_hbm_peak": evals * n * n * 8 / (ms * 1e-3) / 1e9 / PEAK,
            "note": "host-driven engine: one D2H fetch per trial is inside the time"}


def c4():
    r = osb.batched_bfgs_rosenbrock(32, 262144)
    its = int(r["k"].sum())
    return {"config": "C4 batched BFGS, 262144 Rosenbrock problems n=32, one warp per problem", "ms": r["ms"],
            "solves_per_s": 262144 / (r["ms"] * 1e-3), "mean_iterations": float(r["k"].mean()),
            "status_counts": np.bincount(r["status"]).tolist(), "reason_counts": np.bincount(r["reason"]).tolist(),
            "iteration_histogram_p5_p50_p95": [float(np.percentile(r["k"], q)) for q in (5, 50, 95)],
            "approx_fp64_GFLOPs": its * 10 * 32 * 32 / (r["ms"] * 1e-3) / 1e9}


def c5a(m):
    n = 8192
    t0 = time.time()
    obj = osb.LogisticRegression.generated(m, n, 1.0)
    gen_s = time.time() - t0
    ctx = osb.default_context()
    L = osb.lib()
    syrk_ms = osb.bench_syrk(obj, 2)
    nt = (n + 127) // 128
    macs = float(m) * 128 * 128 * (nt * (nt + 1) // 2)  # lower-triangular 128x128 tiles actually computed
    s = osb.Newton(1e-8, np.zeros(n))
    t0 = time.time()
    st = run(s, osb.BackTracking(1e-4, 0.5), obj, 50, 20)
    wall = time.time() - t0
    ms, it = s.last_timing()
    calls = obj.calls()
    return {"config": "C5a Newton + BackTracking, logistic regression m=%d n=8192 (DMMA Hessian + blocked Cholesky)" % m,
            "status": st, "iterations": it, "reason": s.termination_reason(), "ms_total": ms, "s_per_iteration": ms / 1e3 / max(it, 1),
            "oracle_calls": calls, "generate_s": gen_s, "wall_s": wall,
            "hessian_ms": syrk_ms, "hessian_macs_computed": macs, "hessian_TFLOPs": 2 * macs / (syrk_ms * 1e-3) / 1e12,
            "hessian_flops_syrk_convention": float(m) * n * n, "x_bytes": float(m) * n * 8}


def c5b():
    n = 1 << 28
    lb, ub = np.full(n, -1.0), np.full(n, 1.0)
    out = []
    for name, ls in (("GLLQuadratic(1e-4,10)", osb.GLLQuadratic(1e-4, 10)), ("BackTracking(1e-4,0.5)", osb.BackTracking(1e-4, 0.5))):
        obj = osb.SeparableQuadratic.generated(n)
        s = osb.SpectralProjectedGradient(1e-6 if "GLL" in name else 1e-5, np.zeros(n), obj, lb, ub)
        st = run(s, ls, obj, 500, 50)
        ms, it = s.last_timing()
        aset = s.active_set()
        evals = obj.calls()
        out.append({"line_search": name, "status": st, "iterations": it, "reason": s.termination_reason(), "ms_total": ms,
                    "iterations_per_s": it / ms * 1e3, "oracle_evals": evals, "active_fraction": float(np.mean(aset != 0)),
                    "vector_bytes": n * 8, "approx_vector_passes_per_iteration": 14,
                    "approx_GBps": it * 14 * n * 8 / (ms * 1e-3) / 1e9})
    return {"config": "C5b SPG box-constrained separable quadratic n=2^28", "runs": out}


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a in ("c2", "c4", "c5a", "c5b")] or ["c2", "c4", "c5b"]
    m = 262144
    if "--m" in sys.argv:
        m = int(sys.argv[sys.argv.index("--m") + 1])
    for w in which:
        r = {"c2": c2, "c4": c4, "c5b": c5b}[w]() if w != "c5a" else c5a(m)
        print(json.dumps(r), flush=True)
