#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native line-search-solver path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU):
dense BFGS + BackTracking(1e-4, 0.5) on extended Rosenbrock, n = 16384, f64, H (2 GiB) device-resident.
One "step" = one BFGS outer iteration = one pass of the hot path (device line search + h = H y +
fused rank-2 update with u = H' g).  N > 1 shards H by row blocks over the ranks (one process per GPU),
NCCL all-gather of the h / u slices: the total work is fixed, hence "scaling": "strong".

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the reference's own
O(n^3) update on the host cores instead (the reference itself is Rust and cannot be built here).
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DIM = 16384
TOL = 1e-8
MAX_LS = 20
METRIC = "bfgs_iterations_per_second_n16384_f64"
UNIT = "iterations/s"
WORKLOAD = "C3: dense BFGS + BackTracking(1e-4,0.5), extended Rosenbrock n=16384 f64, H row-sharded over N GPUs"


def rosen_x0(n, problem=0):
    """(-1.2, 1, ...) + int16(hash(3, problem, i)) * 2^-16   — vectorised splitmix64."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = np.uint64(3) ^ (np.uint64(problem) * np.uint64(0x9E3779B97F4A7C15) + i)
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
        x = x ^ (x >> np.uint64(31))
    v = (x & np.uint64(0xFFFF)).astype(np.int64)
    v = np.where(v >= 32768, v - 65536, v).astype(np.float64)
    base = np.where(np.arange(n) % 2 == 0, -1.2, 1.0)
    return base + v * 2.0 ** -16


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.reasons, self.max_mhz, self.ok = set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(lazy=True, sym=False):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "qn_lazy_sym_ncu_summary.json" if sym else
                     "qn_lazy_ncu_summary.json" if lazy else "qn_update_ncu_summary.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


def cpu_baseline(threads, n_sample=2048):
    """The oracle port of the reference's update (two dense n^3 products, bfgs.rs:115-124) timed on the
    host cores at n_sample and scaled by (N_DIM / n_sample)^3 — a bounded sample of the same workload."""
    from oracle import oracle as O
    O.build()
    t = O.time_bfgs_update_rowsample(n_sample, n_sample, threads)
    scale = (N_DIM / n_sample) ** 3
    t_iter = t * scale
    return {"value": 1.0 / t_iter, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "oracle/oracle.cpp restatement of bfgs.rs:115-124 (two dense n^3 products) timed once at n=%d "
                      "(%.2f s) and scaled by (16384/%d)^3 = %.0f; O(n^2) passes not counted" % (n_sample, t, n_sample, scale),
            "seconds_per_iteration_extrapolated": t_iter}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # torchrun exports OMP_NUM_THREADS=1 to its workers; libgomp reads the variable when the oracle library is loaded
    # (measured: with the variable at 1, asking for 8 threads afterwards runs 10x slower than 8 threads should)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import oracle as O
    O.build()
    threads = min(threads, max(1, O.lib().orc_max_threads()))
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(threads)
        if i >= args.warmup:
            vals.append(cb)
    v = float(np.mean([c["value"] for c in vals])) if vals else float("nan")
    cb = vals[-1] if vals else cpu_baseline(threads)
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / v if v > 0 else None, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n": N_DIM, "line_search": "BackTracking(1e-4,0.5)",
                       "note": "the reference is a Rust crate (no Rust toolchain in this image): CPU oracle port, "
                               "all host threads over the dgemm columns (the reference itself is single-threaded)"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=N_DIM, help="problem dimension (the benchmark line is only valid at 16384)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--head", type=int, default=0, help="device-engine head variant (0 cluster+speculation, 3 cluster, 1 single CTA)")
    ap.add_argument("--storage", default="auto", choices=["auto", "full", "sym"],
                    help="sym: packed lower triangle of H (n^2 8 B per iteration, lazy schedule; sharded by tile pairs over "
                         "N GPUs with the peer-memory exchange); full: n x n row-major (row-block sharded); auto: sym "
                         "whenever it is available (one GPU, or N GPUs with P2P)")
    ap.add_argument("--qn-kernel", type=int, default=0, help="lazy-pass kernel: 0 = register-staged LDG, 1 = TMA-staged (cp.async.bulk + mbarrier)")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL all-gathers instead of the fused peer-memory exchange")
    ap.add_argument("--schedule", default="lazy", choices=["lazy", "eager"],
                    help="lazy: one read-modify-write of H per iteration (2 n^2 8 B); eager: h = H y then fused update (3 n^2 8 B)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    osb = importlib.import_module("optimization-solvers_b200")
    torch.cuda.set_device(local_rank)
    n = args.n
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = [osb.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx = osb.Context(local_rank, rank, world, uid[0])
        if not args.no_p2p:
            # CUDA IPC: the exchange is fused into the kernels over NVLink peer memory; without it (no peer access,
            # IPC disabled) every rank falls back to the NCCL path together
            if not ctx.connect_peers(strict=False):
                args.no_p2p = True
                if rank == 0:
                    print("bench: peer-memory exchange unavailable, using NCCL all-gathers and full storage", file=sys.stderr)
    else:
        ctx = osb.Context(local_rank)
    osb.set_default_context(ctx)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    x0 = rosen_x0(n, 0)
    obj = osb.ExtendedRosenbrock(n, ctx=ctx)
    ls = osb.BackTracking(1e-4, 0.5)
    lazy = args.schedule == "lazy"
    solver = osb.BFGS(TOL, x0, ctx=ctx).set_option("engine", 2).set_option("qn_schedule", 1 if lazy else 0)
    solver.set_option("use_p2p", 0 if args.no_p2p else 1)
    solver.set_option("head_kernel", args.head)
    solver.set_option("qn_kernel", args.qn_kernel)
    sym = lazy and args.storage in ("sym", "auto") and (world == 1 or not args.no_p2p)
    solver.set_option("qn_storage", 1 if sym else 0)

    def run_steps(k):
        try:
            solver.minimize(ls, obj, k, MAX_LS)
            raise SystemExit("bench: the solve converged inside the timed window; the step count is not what was asked")
        except osb.MaxIterReached:
            pass
        ms, iters = solver.last_timing()
        assert iters == k, (iters, k)
        return ms

    # ---- device-timed region: W warm-up steps, then exactly K steps
    W = max(args.warmup, 3)
    barrier()
    run_steps(W)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    c0 = ctx.counters()
    ms = run_steps(args.steps)
    barrier()
    c1 = ctx.counters()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = args.steps / (ms / 1000.0)

    # ---- roofline of the dominant kernel: per-launch CUDA-event pairs over a second K-step window
    solver.set_option("profile_kernels", 1)
    run_steps(args.steps)
    kt = solver.kernel_timing()
    solver.set_option("profile_kernels", 0)
    rows_local = n // world
    upd_bytes = 2.0 * rows_local * n * 8.0  # read H + write H' (local row block); O(n) vectors excluded
    gemv_bytes = 1.0 * rows_local * n * 8.0
    iter_bytes = (2.0 if lazy else 3.0) * rows_local * n * 8.0
    if sym:
        upd_bytes = iter_bytes = 1.0 * n * n * 8.0 / world  # read + write of (this rank's share of) the lower triangle
    peak, peak_src = hbm_peak()
    sym_parts = None
    if sym:  # slot 0 = streaming pass, slot 1 = column fold + epilogue; the roofline is quoted on their sum
        sym_parts = {"pass_ms": kt["gemv_ms"], "fold_ms": kt["update_ms"]}
        kt = dict(kt, update_ms=kt["gemv_ms"] + kt["update_ms"])
    ach = upd_bytes / (kt["update_ms"] * 1e-3) / 1e9 if kt["update_ms"] > 0 else None
    kname = ("qn_lazy_sym_kernel<BFGS> + fold (packed lower triangle: pending rank-2 RMW + row and column sums)" if sym else
             "qn_lazy_kernel<BFGS> (pending rank-2 RMW + h = H y + w = H g in one pass)" if lazy
             else "qn_update_kernel<BFGS> (fused rank-2 RMW + u = H' g)")
    roofline = {"bound": "hbm", "kernel": kname, "achieved": ach,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": ncu_traffic(lazy, sym) if world == 1 else None, "algorithmic_bytes_per_launch": upd_bytes, "ms_per_launch": kt["update_ms"], "sym_parts": sym_parts,
                "gemv_kernel": None if lazy else {
                    "achieved": gemv_bytes / (kt["gemv_ms"] * 1e-3) / 1e9 if kt["gemv_ms"] > 0 else None,
                    "ms_per_launch": kt["gemv_ms"], "algorithmic_bytes_per_launch": gemv_bytes},
                "iteration_bytes": iter_bytes,
                "iteration_frac_of_peak": (iter_bytes / (ms / args.steps * 1e-3) / 1e9) / peak,
                # SURVEY 8d counts the 3-pass form, 3 n^2 8 B per iteration (per GPU: / N): the same step time expressed on
                # that byte count (> 1 means faster than ANY implementation that moves those bytes could be at peak)
                "survey_8d_iteration": {"bytes": 3.0 * n * n * 8.0 / world,
                                        "achieved": 3.0 * n * n * 8.0 / world / (ms / args.steps * 1e-3) / 1e9,
                                        "frac_of_measured_peak": 3.0 * n * n * 8.0 / world / (ms / args.steps * 1e-3) / 1e9 / peak,
                                        "frac_of_nominal_8TBps": 3.0 * n * n * 8.0 / world / (ms / args.steps * 1e-3) / 1e9 / 8000.0}}

    # ---- end to end through the public API with HOST buffers: construction from a pinned host x0 (H2D),
    # minimize with a per-iteration host callback that reads the iterate back (D2H), final x() (D2H)
    e2e = None
    if world == 1:
        k_e2e = args.steps
        x0_pinned = torch.from_numpy(x0).pin_memory()
        runs = []
        for _rep in range(3):  # host-side timing is noisy on shared boxes: median of three identical runs
            xs = []
            barrier()
            t0 = time.perf_counter()
            s2 = osb.BFGS(TOL, x0_pinned.numpy(), ctx=ctx).set_option("qn_schedule", 1 if lazy else 0)
            s2.set_option("qn_storage", 1 if sym else 0)
            s2.set_option("callback_run_ahead", 1)  # the callback reads pinned snapshots of (x, f, k); the device is not stalled
            t_c = time.perf_counter()

            def cb(s):
                xs.append((s.x()[0], s.f()))
            try:
                s2.minimize(osb.BackTracking(1e-4, 0.5), obj, k_e2e, MAX_LS, callback=cb)
            except osb.MaxIterReached:
                pass
            t_m = time.perf_counter()
            xf = s2.x()
            ff = s2.f()
            ctx.synchronize()
            t1 = time.perf_counter()
            assert len(xs) == k_e2e and np.isfinite(ff) and xf.shape == (n,)
            runs.append((t1 - t0, t_c - t0, t_m - t_c, t1 - t_m))
            s2.close()
        runs.sort()
        tt, tc_, tm_, tr_ = runs[1]
        e2e = {"value": k_e2e / tt, "unit": UNIT, "h2d_bytes_per_step": int(n * 8 / k_e2e),
               "d2h_bytes_per_step": int(n * 8 + 8 + n * 8 / k_e2e), "steps": k_e2e,
               "construct_ms": tc_ * 1e3, "minimize_ms": tm_ * 1e3, "readback_ms": tr_ * 1e3,
               "all_runs_it_per_s": [k_e2e / r[0] for r in runs],
               "what": "BFGS::new(tol, host x0) + minimize(K iterations, host callback reading x() and f() every iteration from "
                       "the pinned per-iteration snapshot, option callback_run_ahead) + x(), f(): wall clock around the calls, "
                       "median of 3 runs; construction amortised over K"}
    else:
        # sharded: the public API call itself (host x0 in, host x out), wall clock, max over ranks
        barrier()
        t0 = time.perf_counter()
        s2 = osb.BFGS(TOL, x0, ctx=ctx).set_option("engine", 2).set_option("qn_schedule", 1 if lazy else 0)
        s2.set_option("use_p2p", 0 if args.no_p2p else 1)
        s2.set_option("qn_storage", 1 if sym else 0)
        try:
            s2.minimize(osb.BackTracking(1e-4, 0.5), obj, args.steps, MAX_LS)
        except osb.MaxIterReached:
            pass
        xf = s2.x()
        ctx.synchronize()
        t1 = time.perf_counter()
        t = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": args.steps / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": int(n * 8 / args.steps),
               "d2h_bytes_per_step": int(n * 8 / args.steps), "steps": args.steps,
               "what": "BFGS::new(tol, host x0) + minimize(K) + x() per rank, wall clock, max over ranks"}
        s2.close()

    if rank == 0:
        cb_line = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb_line = cpu_baseline(1)
            except Exception as e:  # the checker is optional for the product line
                cb_line = {"error": str(e)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "n": n, "line_search": "BackTracking(1e-4,0.5)", "tol": TOL,
                           "max_iter_line_search": MAX_LS, "engine": "device-resident control",
                           "schedule": "lazy + packed symmetric storage: 1 RMW pass of the lower triangle per iteration (n^2 8 B)" if sym else
                           "lazy: 1 RMW pass of H per iteration (2 n^2 8 B)" if lazy else "eager: gemv + fused update (3 n^2 8 B)",
                           "storage": "packed lower triangle, 8-row tiles" if sym else "full n x n row-major",
                           "l2": "inputs larger than L2 (H = %.2f GiB per GPU, streamed every step)" % (
                               (n * (n + 8) / 2 / world if sym else rows_local * n) * 8 / 2 ** 30),
                           "parallelism": ("packed triangle sharded by tile pairs over %d GPU(s); exchange: %s" if sym else
                                           "row-block sharded H over %d GPU(s); exchange: %s") % (
                               world, "none" if world == 1 else (
                                   "per-rank {h, w} contributions stored into every peer's slot by the fold kernel (NVLink stores + flags), summed in rank order by the head" if sym
                                   else "NCCL all-gather of the h / w slices" if (args.no_p2p or not lazy)
                                   else "peer-memory all-gather fused into the lazy kernel (NVLink stores + flags)"))},
                "roofline": roofline, "cpu_baseline": cb_line, "e2e": e2e,
                "gpu_launches": int(c1["launches"] - c0["launches"]),
                "ls_trials_per_step": (c1["ls_trials"] - c0["ls_trials"]) / args.steps,
                "collectives_per_step": (c1["collectives"] - c0["collectives"]) / args.steps,
                "clocks": sampler.summary()}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
