"""Multi-GPU check (run under torchrun on a GPU box, one process per GPU):
row-block sharded BFGS/DFP must reproduce the single-GPU result BIT-FOR-BIT (a row's dot product is
formed inside one CTA whatever the sharding; every O(n) vector and scalar is replicated).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _second_uid(osb, dist, rank):
    uid = [osb.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    return uid[0]


def main():
    import torch
    import torch.distributed as dist
    from test_gpu_parity import rosen_x0
    osb = importlib.import_module("optimization-solvers_b200")
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    uid = [osb.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = osb.Context(lr, rank, world, uid[0])
    ctx.connect_peers()  # CUDA IPC exchange regions: fused NVLink all-gather in the lazy pass
    solo = osb.Context(lr)
    ok = True
    # (kind, n, iterations, schedule 0 = eager / 1 = lazy, use_p2p)
    for kind, n, iters, sched, p2p in (("BFGS", 2048, 40, 0, 0), ("DFP", 1024, 25, 0, 0), ("BFGS", 16384, 12, 0, 0),
                                       ("BFGS", 2048, 40, 1, 0), ("BFGS", 2048, 40, 1, 1), ("DFP", 1024, 25, 1, 1),
                                       ("BFGS", 16384, 12, 1, 1)):
        x0 = rosen_x0(n, 5)
        res = []
        for c in (ctx, solo):
            # full n x n storage, row-block sharded (the packed triangle, the default storage, is checked below)
            s = getattr(osb, kind)(1e-8, x0, ctx=c).set_option("engine", 2).set_option("qn_schedule", sched).set_option("qn_storage", 0)
            s.set_option("use_p2p", p2p)
            try:
                s.minimize(osb.BackTracking(1e-4, 0.5), osb.ExtendedRosenbrock(n, ctx=c), iters, 20)
            except osb.MaxIterReached:
                pass
            H = s.approx_inv_hessian()
            res.append((s.k(), s.x(), s.s_norm(), H))
            s.close()
        rows = slice(rank * n // world, (rank + 1) * n // world)
        same = (res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]
                and np.array_equal(res[0][3][rows], res[1][3][rows]))
        print("rank %d %s n=%d k=%d schedule=%d p2p=%d bit-identical to single GPU: %s"
              % (rank, kind, n, res[0][0], sched, p2p, same), flush=True)
        ok &= bool(same)
    # ---- packed symmetric storage sharded by tile pairs (fused peer-memory exchange of per-rank slots, summed in rank
    # order by the head): same trajectory as one GPU to rounding of the summation order, identical on all ranks
    for kind, n, iters in (("BFGS", 2048, 40), ("DFP", 1024, 25), ("BFGS", 16384, 12)):
        x0 = rosen_x0(n, 5)
        res = []
        for c in (ctx, solo):
            s = getattr(osb, kind)(1e-8, x0, ctx=c).set_option("engine", 2).set_option("qn_schedule", 1).set_option("qn_storage", 1)
            try:
                s.minimize(osb.BackTracking(1e-4, 0.5), osb.ExtendedRosenbrock(n, ctx=c), iters, 20)
            except osb.MaxIterReached:
                pass
            launches = c.counters()["sharded_packed_passes"]
            H = s.approx_inv_hessian()  # (collective on the sharded context: all ranks call it)
            res.append((s.k(), s.x(), s.s_norm(), H, launches))
            s.close()
        rows = slice(rank * n // world, (rank + 1) * n // world)
        xs = [None] * world
        dist.all_gather_object(xs, res[0][1].tobytes())
        same_on_ranks = all(b == xs[0] for b in xs)
        dx = float(np.max(np.abs(res[0][1] - res[1][1])))
        dH = float(np.max(np.abs(res[0][3][rows] - res[1][3][rows])))
        good = res[0][0] == res[1][0] and res[0][4] > 0 and res[1][4] == 0 and same_on_ranks and dx <= 1e-9 and dH <= 1e-8 * max(1.0, float(np.max(np.abs(res[1][3][rows]))))
        print("rank %d %s n=%d k=%d/%d packed storage sharded by tile pairs: ranks identical %s, max|dx| vs one GPU %.2e, max|dH| %.2e -> %s"
              % (rank, kind, n, res[0][0], res[1][0], same_on_ranks, dx, dH, good), flush=True)
        ok &= bool(good)

    # ---- packed sharded storage starting from a user matrix (set_approx_inv_hessian: the matrix arrives as row blocks and is
    # redistributed into tile pairs once)
    for kind, n, iters in (("BFGS", 2048, 15), ("BFGS", 8192, 8)):
        rng = np.random.default_rng(17)
        x0 = rosen_x0(n, 7)
        dsc = 0.5 + rng.random(n)
        lowrank = rng.standard_normal((n, 3)) * 0.05
        H0 = np.diag(dsc) + lowrank @ lowrank.T  # symmetric positive definite, dense
        H0 = 0.5 * (H0 + H0.T)
        res = []
        for c in (ctx, solo):
            s = getattr(osb, kind)(1e-8, x0, ctx=c)  # library defaults
            s.set_approx_inv_hessian(H0)
            try:
                s.minimize(osb.BackTracking(1e-4, 0.5), osb.ExtendedRosenbrock(n, ctx=c), iters, 20)
            except osb.MaxIterReached:
                pass
            info = s.path_info()
            H = s.approx_inv_hessian()
            res.append((s.k(), s.x(), H, info))
            s.close()
        rows = slice(rank * n // world, (rank + 1) * n // world)
        xs = [None] * world
        dist.all_gather_object(xs, res[0][1].tobytes())
        same_on_ranks = all(b == xs[0] for b in xs)
        dx = float(np.max(np.abs(res[0][1] - res[1][1])))
        dH = float(np.max(np.abs(res[0][2][rows] - res[1][2][rows])))
        good = (res[0][0] == res[1][0] and res[0][3]["sharded_packed"] and res[0][3]["fused"] and same_on_ranks and dx <= 1e-9
                and dH <= 1e-8 * max(1.0, float(np.max(np.abs(res[1][2][rows])))))
        print("rank %d %s n=%d k=%d/%d packed sharded storage from a user matrix (sharded_packed=%s fused=%s): ranks identical %s, max|dx| %.2e, max|dH| %.2e -> %s"
              % (rank, kind, n, res[0][0], res[1][0], res[0][3]["sharded_packed"], res[0][3]["fused"], same_on_ranks, dx, dH, good), flush=True)
        ok &= bool(good)

    # ---- Broyden on row blocks: v = H^T s is summed over the ranks (all-reduce), H is not symmetric
    n = 2048
    res = []
    for c in (ctx, solo):
        obj = osb.DenseQuadratic.generated(n, True, ctx=c)
        s = osb.Broyden(1e-6, obj.x0, ctx=c)
        try:
            s.minimize(osb.BackTracking(1e-4, 0.5), obj, 30, 30)
        except osb.MaxIterReached:
            pass
        H = s.approx_inv_hessian()
        res.append((s.k(), s.x(), s.f(), H))
        s.close()
    rows = slice(rank * n // world, (rank + 1) * n // world)
    dx = float(np.max(np.abs(res[0][1] - res[1][1])))
    dH = float(np.max(np.abs(res[0][3][rows] - res[1][3][rows])))
    good = res[0][0] == res[1][0] and dx <= 1e-9 and dH <= 1e-9 * max(1.0, float(np.max(np.abs(res[1][3]))))
    print("rank %d Broyden n=%d k=%d/%d row-sharded H with all-reduced H^T s: max|dx| %.2e max|dH| %.2e -> %s" % (rank, n, res[0][0], res[1][0], dx, dH, good), flush=True)
    ok &= bool(good)

    # ---- C2 shape: GradientDescent on a row-sharded dense quadratic (A row-block sharded, x / g replicated, all-gather
    # of (A x)_p): bit-identical to one GPU
    n = 4096
    res = []
    for c in (ctx, solo):
        obj = osb.DenseQuadratic.generated(n, True, ctx=c)
        s = osb.GradientDescent(1e-6, obj.x0, ctx=c)
        try:
            s.minimize(osb.BackTracking(1e-4, 0.5), obj, 400, 100)
        except osb.MaxIterReached:
            pass
        res.append((s.k(), s.x(), s.f(), s.termination_reason()))
        s.close()
    same = res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2] and res[0][3] == res[1][3]
    print("rank %d GD dense quadratic n=%d k=%d reason=%s row-sharded A bit-identical to single GPU: %s"
          % (rank, n, res[0][0], res[0][3], same), flush=True)
    ok &= bool(same)

    # ---- C5b shape: SPG / PGD / GD with index-range sharded vectors (every scalar combined across ranks in rank order):
    # same iteration count and termination reason, bit-exact active set, x and f to rounding of the summation order
    from importlib import import_module
    shard_indices = import_module("optimization-solvers_b200.dist").shard_indices
    n = 1 << 20
    i0, cnt = shard_indices(n, rank, world)
    vctx = osb.Context(lr, rank, world, _second_uid(osb, dist, rank)).set_vector_sharding(True)
    for name, mk_ls, tol in (("SPG+GLL", lambda: osb.GLLQuadratic(1e-4, 10), 1e-6), ("SPG+BackTracking", lambda: osb.BackTracking(1e-4, 0.5), 1e-5),
                             ("PGD+BackTracking", lambda: osb.BackTracking(1e-4, 0.5), 1e-5)):
        out = []
        for c, nn, off in ((vctx, cnt, i0), (solo, n, 0)):
            obj = (osb.SeparableQuadratic.generated_shard(nn, off, c) if c is vctx else osb.SeparableQuadratic.generated(nn, ctx=c))
            lb, ub = np.full(nn, -1.0), np.full(nn, 1.0)
            if name.startswith("SPG"):
                s = osb.SpectralProjectedGradient(tol, np.zeros(nn), obj, lb, ub, ctx=c)
            else:
                s = osb.ProjectedGradientDescent(tol, np.zeros(nn), lb, ub, ctx=c)
            try:
                s.minimize(mk_ls(), obj, 500, 50)
            except osb.MaxIterReached:
                pass
            out.append((s.k(), s.termination_reason(), s.x(), s.f(), s.active_set()))
            s.close()
        (k1, r1, x1, f1, a1), (k0, r0, x0_, f0, a0) = out
        sl = slice(i0, i0 + cnt)
        good = (k1 == k0 and r1 == r0 and np.array_equal(a1, a0[sl]) and abs(f1 - f0) <= 1e-12 * abs(f0)
                and np.max(np.abs(x1 - x0_[sl])) <= 1e-9)
        print("rank %d %s n=2^20 index-range sharded: k=%d/%d reason=%s/%s f rel diff %.2e max|dx| %.2e active sets equal %s -> %s"
              % (rank, name, k1, k0, r1, r0, abs(f1 - f0) / abs(f0), np.max(np.abs(x1 - x0_[sl])), np.array_equal(a1, a0[sl]), good), flush=True)
        ok &= bool(good)

    # ---- C5a shape: Newton on the logistic regression with the SAMPLES sharded over the ranks (all-reduce of f, g and the
    # n x n Hessian, newton/mod.rs:26-49 replicated): same iteration count and reason, x within 1e-9, identical on all ranks
    m_, n = 4096, 96
    out = []
    for c in (ctx, solo):
        obj = osb.LogisticRegression.generated(m_, n, 1.0, ctx=c)
        s = osb.Newton(1e-8, np.zeros(n), ctx=c)
        st = "Ok"
        try:
            s.minimize(osb.BackTracking(1e-4, 0.5), obj, 50, 20)
        except osb.SolverError as e:
            st = type(e).__name__
        out.append((st, s.k(), s.termination_reason(), s.x(), s.decrement_squared()))
        s.close()
    xs = [None] * world
    dist.all_gather_object(xs, out[0][3].tobytes())
    same_on_ranks = all(b == xs[0] for b in xs)
    dx = float(np.max(np.abs(out[0][3] - out[1][3])) / max(1.0, float(np.max(np.abs(out[1][3])))))
    good = out[0][:3] == out[1][:3] and out[0][0] == "Ok" and same_on_ranks and dx <= 1e-9
    print("rank %d Newton logistic m=%d n=%d sample-sharded: %s k=%d/%d reason=%s ranks identical %s max rel|dx| %.2e -> %s"
          % (rank, m_, n, out[0][0], out[0][1], out[1][1], out[0][2], same_on_ranks, dx, good), flush=True)
    ok &= bool(good)

    # ---- C4 shape: the batched mode split over the ranks (problem0 offsets, no collective): every rank's slice is
    # bit-identical to the same problems solved in one single-GPU batch
    np_, nb = 1024, 32
    per = np_ // world
    mine = osb.batched_bfgs_rosenbrock(nb, per, problem0=rank * per, ctx=ctx)
    full = osb.batched_bfgs_rosenbrock(nb, np_, problem0=0, ctx=solo)
    sl = slice(rank * per, (rank + 1) * per)
    good = (np.array_equal(mine["x"], full["x"][sl]) and np.array_equal(mine["k"], full["k"][sl]) and
            np.array_equal(mine["status"], full["status"][sl]) and np.array_equal(mine["reason"], full["reason"][sl]))
    print("rank %d batched BFGS %d problems split over %d ranks: slice bit-identical to the single-GPU batch: %s" % (rank, np_, world, good), flush=True)
    ok &= bool(good)

    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(t.item()) != 1:
        raise SystemExit("dist_check FAILED")
    if rank == 0:
        print("dist_check ok")


if __name__ == "__main__":
    main()
