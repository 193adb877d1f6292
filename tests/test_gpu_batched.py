"""GPU (-m gpu): batched mode (BASELINE.json configs[3]: many independent Rosenbrock problems, n = 32,
BFGS + BackTracking(1e-4, 0.5), tol 1e-8, max 2000/20) — one warp per problem.

The kernel replays the oracle's rank-2 operation order exactly, so the comparison is BIT-EXACT:
iterates, objective, iteration count, status and termination reason per problem."""
import numpy as np
import pytest

from test_gpu_parity import rosen_x0

pytestmark = pytest.mark.gpu


def oracle_solve(orc, x0, tol=1e-8, mi=2000, ml=20):
    s = orc.BFGS(tol, x0).set_update_form("rank2")
    obj = orc.ExtendedRosenbrock()
    try:
        s.minimize(orc.BackTracking(1e-4, 0.5), obj, mi, ml)
        st = 0
    except orc.MaxIterReached:
        st = 1
    except orc.OutOfDomain:
        st = 2
    reason = {None: 0, "grad_tol": 1, "s_norm": 2, "y_norm": 3}[s.termination_reason()]
    return s.x(), obj(s.x()).f(), s.k(), st, reason


@pytest.mark.parametrize("n", [32, 8, 2, 20])
def test_batched_bit_exact_vs_oracle(osb, orc, n):
    npb = 48 if n == 32 else 12
    x0 = np.stack([rosen_x0(n, p) for p in range(npb)])
    r = osb.batched_bfgs_rosenbrock(n, npb, x0=x0)
    for p in range(npb):
        x, f, k, st, reason = oracle_solve(orc, x0[p])
        assert r["k"][p] == k and r["status"][p] == st and r["reason"][p] == reason, (p, r["k"][p], k)
        assert np.array_equal(r["x"][p], x), (p, np.max(np.abs(r["x"][p] - x)))
        assert r["f"][p] == f
    if n == 32:
        assert np.all(r["status"] == 0) and r["k"].min() > 50


def test_batched_generated_matches_explicit_and_is_sharded_consistently(osb):
    # device-generated start points == the host replay of the same hash; a shard (problem0 offset) of the
    # batch gives the same per-problem results as the whole batch (how N GPUs split the work, no collective)
    n, npb = 32, 64
    x0 = np.stack([rosen_x0(n, p) for p in range(npb)])
    a = osb.batched_bfgs_rosenbrock(n, npb, x0=x0)
    b = osb.batched_bfgs_rosenbrock(n, npb)
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["k"], b["k"])
    c = osb.batched_bfgs_rosenbrock(n, 16, problem0=32)
    assert np.array_equal(c["x"], a["x"][32:48]) and np.array_equal(c["k"], a["k"][32:48])


def test_batched_edge_cases(osb):
    n = 32
    # already at the minimiser: converges at k = 0 by the gradient test
    r = osb.batched_bfgs_rosenbrock(n, 3, x0=np.ones((3, n)))
    assert np.all(r["k"] == 0) and np.all(r["status"] == 0) and np.all(r["reason"] == 1)
    # max_iter = 0 -> MaxIterReached with k = 0 (ls_solver.rs:78,109-110)
    r = osb.batched_bfgs_rosenbrock(n, 2, max_iter_solver=0)
    assert np.all(r["k"] == 0) and np.all(r["status"] == 1)
    # a NaN start is out of domain (ls_solver.rs:37-40)
    x0 = np.ones((1, n))
    x0[0, 3] = np.nan
    r = osb.batched_bfgs_rosenbrock(n, 1, x0=x0)
    assert r["status"][0] == 2
    with pytest.raises(osb.ErrorInputParams):
        osb.batched_bfgs_rosenbrock(64, 1)


def test_batched_full_size_statistics(osb):
    # 65,536 problems (a quarter of the BASELINE batch): all converge; iteration counts in the expected range
    r = osb.batched_bfgs_rosenbrock(32, 1 << 16)
    assert np.all(r["status"] == 0)
    assert 100 < np.median(r["k"]) < 400
    # s_norm / y_norm exits (bfgs.rs:67-72) can stop a little short of the gradient tolerance
    # (a handful of the 65,536 starts stall early through those exits, exactly as the oracle does)
    near = np.max(np.abs(r["x"] - 1.0), axis=1) < 1e-3
    assert np.mean(near) > 0.995 and np.all(r["f"][near] < 1e-8)
