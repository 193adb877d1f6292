"""Measured runs of the other BASELINE.json configs (the headline C3 lives in bench.py).
Prints one JSON object per config; results are copied to profiles/ by hand.

    python tools/bench_configs.py [c2] [c4] [c5a] [c5b] [--m 262144]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_configs.py c2 c5b        # C2 with A row-sharded, C5b with index-range sharded vectors
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


RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
_DIST = None


def dist_ctx(vector_sharding=False):
    """Per-rank library context under torchrun (a fresh NCCL communicator per call); the default context otherwise."""
    global _DIST
    if WORLD == 1:
        return osb.default_context()
    import torch
    import torch.distributed as dist
    if _DIST is None:
        torch.cuda.set_device(LOCAL)
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
        _DIST = dist
    uid = [osb.Context.nccl_unique_id() if RANK == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = osb.Context(LOCAL, RANK, WORLD, uid[0])
    if vector_sharding:
        ctx.set_vector_sharding(True)
    return ctx


def max_over_ranks(v):
    if WORLD == 1:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    _DIST.all_reduce(t, op=_DIST.ReduceOp.MAX)
    return float(t.item())


def run(solver, ls, obj, mi, ml):
    try:
        solver.minimize(ls, obj, mi, ml)
        return "Ok"
    except osb.SolverError as e:
        return type(e).__name__


def c2():
    # GradientDescent + BackTracking(1e-4, 0.5), dense SPD quadratic n = 16384 (GEMV-bound: n^2 * 8 B per oracle call)
    n = 16384
    ctx = dist_ctx()
    obj = osb.DenseQuadratic.generated(n, True, ctx=ctx)
    s = osb.GradientDescent(1e-6, obj.x0, ctx=ctx)
    run(s, osb.BackTracking(1e-4, 0.5), obj, 3, 100)  # warm-up
    s = osb.GradientDescent(1e-6, obj.x0, ctx=ctx)
    c0 = obj.calls()
    st = run(s, osb.BackTracking(1e-4, 0.5), obj, 1000, 100)
    ms, it = s.last_timing()
    ms = max_over_ranks(ms)
    evals = obj.calls() - c0
    return {"config": "C2 GD + BackTracking, dense SPD quadratic n=16384, A row-sharded over %d GPU(s)" % WORLD, "n_gpus": WORLD,
            "status": st, "iterations": it, "reason": s.termination_reason(),
            "ms_total": ms, "iterations_per_s": it / ms * 1e3, "oracle_evals": evals, "evals_per_iteration": evals / max(it, 1),
            "bytes_per_eval": n * n * 8, "achieved_GBps": evals * n * n * 8 / (ms * 1e-3) / 1e9,
            "frac_of_measured_hbm_peak": evals * n * n * 8 / (ms * 1e-3) / 1e9 / PEAK / WORLD,
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
    ctx = dist_ctx(vector_sharding=True)
    nl, i0 = n // WORLD, RANK * (n // WORLD)  # index-range sharding: this rank owns coordinates [i0, i0 + nl)
    lb, ub = np.full(nl, -1.0), np.full(nl, 1.0)
    out = []
    for name, ls in (("GLLQuadratic(1e-4,10)", osb.GLLQuadratic(1e-4, 10)), ("BackTracking(1e-4,0.5)", osb.BackTracking(1e-4, 0.5))):
        obj = osb.SeparableQuadratic.generated_shard(nl, i0, ctx) if WORLD > 1 else osb.SeparableQuadratic.generated(n, ctx=ctx)
        s = osb.SpectralProjectedGradient(1e-6 if "GLL" in name else 1e-5, np.zeros(nl), obj, lb, ub, ctx=ctx)
        st = run(s, ls, obj, 500, 50)
        ms, it = s.last_timing()
        ms = max_over_ranks(ms)
        aset = s.active_set()
        evals = obj.calls()
        act = float(np.mean(aset != 0))
        if WORLD > 1:
            act = max_over_ranks(act)  # (per-rank fractions are equal to 1e-4 on this problem; reported as the max)
        out.append({"line_search": name, "status": st, "iterations": it, "reason": s.termination_reason(), "ms_total": ms,
                    "iterations_per_s": it / ms * 1e3, "oracle_evals": evals, "active_fraction": act, "n_gpus": WORLD,
                    "vector_bytes": n * 8, "approx_vector_passes_per_iteration": 14,
                    "approx_GBps": it * 14 * n * 8 / (ms * 1e-3) / 1e9})
    return {"config": "C5b SPG box-constrained separable quadratic n=2^28, vectors index-range sharded over %d GPU(s)" % WORLD, "runs": out}


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a in ("c2", "c4", "c5a", "c5b")] or ["c2", "c4", "c5b"]
    m = 262144
    if "--m" in sys.argv:
        m = int(sys.argv[sys.argv.index("--m") + 1])
    for w in which:
        r = {"c2": c2, "c4": c4, "c5b": c5b}[w]() if w != "c5a" else c5a(m)
        if RANK == 0:
            print(json.dumps(r), flush=True)
    if _DIST is not None:
        _DIST.barrier()
        _DIST.destroy_process_group()
