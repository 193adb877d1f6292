"""GPU (-m gpu): Newton family (src/newton/mod.rs, projected_newton.rs, spn.rs) on the synthetic logistic
regression of BASELINE.json configs[4] and on dense quadratics: DMMA Hessian assembly, blocked Cholesky."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def close(a, b, rtol=RTOL, atol=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(1.0, float(np.max(np.abs(b))))
    return bool(np.all(np.abs(a - b) <= atol + rtol * scale))


def run(m, solver, ls, oracle, mi, ml):
    try:
        solver.minimize(ls, oracle, mi, ml)
        return "Ok"
    except m.SolverError as e:
        return type(e).__name__


@pytest.mark.parametrize("m,n", [(512, 64), (300, 200), (1000, 130), (37, 258), (5000, 384)])
def test_logistic_objective_matches_oracle(osb, orc, m, n):
    # f, g (two passes over X) and the DMMA Hessian X^T D X + lambda I: ragged tiles, K splits that are empty or partial
    rng = np.random.default_rng(m + n)
    w = rng.standard_normal(n) * 0.3
    a = osb.LogisticRegression.generated(m, n, 1.0)(w)
    b = orc.LogisticRegression.generated(m, n, 1.0)(w)
    assert abs(a.f() - b.f()) <= 1e-12 * abs(b.f())
    assert close(a.g(), b.g(), rtol=1e-12)
    assert close(a.hessian(), b.hessian(), rtol=1e-12)
    assert np.array_equal(a.hessian(), a.hessian().T)


@pytest.mark.parametrize("ls", ["bt", "mt"])
def test_newton_logistic_vs_oracle(osb, orc, ls):
    m_, n = 2048, 64

    def script(m):
        obj = m.LogisticRegression.generated(m_, n, 1.0)
        s = m.Newton(1e-8, np.zeros(n))
        lsearch = m.BackTracking(1e-4, 0.5) if ls == "bt" else m.MoreThuente.default()
        st = run(m, s, lsearch, obj, 50, 20)
        return st, s.k(), s.termination_reason(), s.x(), s.decrement_squared()

    ref, got = script(orc), script(osb)
    assert got[:3] == ref[:3], (got[:3], ref[:3])
    assert ref[0] == "Ok" and ref[2] == "newton_decrement"
    assert close(got[3], ref[3]) and close(got[4], ref[4], atol=1e-18)


def test_newton_logistic_many_tiles_and_splits_vs_oracle(osb, orc):
    """C5a at a size where the Hessian assembly runs many 128 x 128 tiles and full K splits (m = 8192, n = 1024: 36 lower
    tiles x 4 splits of 43 stages) and the blocked Cholesky takes several panels: 3 Newton iterations against the oracle,
    x and the decrement within 1e-9 relative."""
    m_, n = 8192, 1024

    def script(m):
        obj = m.LogisticRegression.generated(m_, n, 1.0)
        s = m.Newton(1e-12, np.zeros(n))
        st = run(m, s, m.BackTracking(1e-4, 0.5), obj, 3, 20)
        return st, s.k(), s.x(), s.decrement_squared()

    ref, got = script(orc), script(osb)
    assert got[:2] == ref[:2]
    assert close(got[2], ref[2], rtol=1e-9) and abs(got[3] - ref[3]) <= 1e-9 * abs(ref[3])


def test_projected_and_spectral_newton_logistic_vs_oracle(osb, orc):
    m_, n = 1024, 48
    lb, ub = np.full(n, -0.05), np.full(n, 0.05)
    for cls in ("ProjectedNewton", "SpectralProjectedNewton"):
        def script(m):
            obj = m.LogisticRegression.generated(m_, n, 1.0)
            if cls == "ProjectedNewton":
                s = m.ProjectedNewton(1e-7, np.zeros(n), lb, ub)
                lsearch = m.BackTrackingB(1e-4, 0.5, lb, ub)
            else:
                s = m.SpectralProjectedNewton(1e-7, np.zeros(n), obj, lb, ub)
                lsearch = m.GLLQuadratic(1e-4, 10)
            st = run(m, s, lsearch, obj, 12, 30)
            return st, s.k(), s.termination_reason(), s.x(), s.active_set()

        ref, got = script(orc), script(osb)
        assert got[:3] == ref[:3], (cls, got[:3], ref[:3])
        assert close(got[3], ref[3])
        assert np.array_equal(got[4], ref[4])


@pytest.mark.parametrize("n", [300, 1000, 2050])
def test_blocked_cholesky_newton_step(osb, n):
    # one full Newton step on a dense SPD quadratic lands on the minimiser: x1 = x0 - (2A)^-1 g(x0) = A^-1 b
    obj = osb.DenseQuadratic.generated(n, True)
    s = osb.Newton(1e-30, obj.x0)
    assert run(osb, s, osb.NoSearch(), obj, 1, 1) == "MaxIterReached"
    x1 = s.x()
    g1 = obj(x1).g()
    g0 = obj(obj.x0).g()
    assert np.max(np.abs(g1)) <= 1e-11 * np.max(np.abs(g0))
    H = obj(obj.x0).hessian()
    ref = obj.x0 - np.linalg.solve(H, g0)
    assert close(x1, ref, rtol=1e-11)
    assert s.decrement_squared() is not None and s.decrement_squared() > 0


def test_not_spd_is_reported_like_the_reference_panic(osb):
    # projected_newton.rs:75 `.cholesky().unwrap()` panics on a non-SPD Hessian
    H = np.array([[1.0, 0.0], [0.0, -1.0]])

    def orac(x):
        return osb.FuncEvalMultivariate(0.5 * x @ H @ x, H @ x).with_hessian(H)
    s = osb.ProjectedNewton(1e-8, [1.0, 1.0], [-5.0, -5.0], [5.0, 5.0])
    with pytest.raises(osb.ReferencePanic):
        s.minimize(osb.BackTrackingB(1e-4, 0.5, [-5.0, -5.0], [5.0, 5.0]), osb.HostOracle(orac, 2, True), 5, 5)
    # missing Hessian: newton/mod.rs:34 `.expect("Hessian not available in the oracle")`
    s = osb.Newton(1e-8, [1.0, 1.0])
    with pytest.raises(osb.ReferencePanic):
        s.minimize(osb.NoSearch(), osb.ExtendedRosenbrock(2), 5, 5)


def _nonconvex_oracle(m, n, singular):
    """f = 1/2 x^T H x + 1/4 sum x_i^4 with an indefinite (or exactly singular) constant part H."""
    rng = np.random.default_rng(7)
    Q = np.round(rng.standard_normal((n, n)) * 4) / 8      # small dyadic entries: identical bits on both sides
    H = Q + Q.T
    H[np.diag_indices(n)] = np.where(np.arange(n) % 2 == 0, 3.0, -2.0)

    def fn(x):
        if singular:
            # Hessian with an exactly zero first row / column: every pivot candidate of column 0 is 0.0
            Hs = H.copy()
            Hs[0, :] = 0.0
            Hs[:, 0] = 0.0
            return m.FuncEvalMultivariate(0.5 * x @ Hs @ x + x[0], Hs @ x + np.eye(n)[0]).with_hessian(Hs)
        hess = H + np.diag(3.0 * x * x)
        return m.FuncEvalMultivariate(0.5 * x @ H @ x + 0.25 * np.sum(x ** 4), H @ x + x ** 3).with_hessian(hess)
    return fn


@pytest.mark.parametrize("n", [3, 6, 40])
def test_newton_indefinite_hessian_goes_through_lu_like_try_inverse(osb, orc, n):
    # newton/mod.rs:36: try_inverse (LU with partial pivoting) accepts an indefinite Hessian; only ProjectedNewton / SPN
    # panic on a failed Cholesky
    def script(m, host):
        fn = _nonconvex_oracle(m, n, False)
        o = m.HostOracle(fn, n, True) if host else m.HostOracle(fn, True)
        s = m.Newton(1e-10, np.linspace(-1.0, 1.0, n) * 0.5 + 0.25)
        st = run(m, s, m.NoSearch(), o, 4, 5)
        return st, s.k(), s.termination_reason(), s.x(), s.decrement_squared()

    ref, got = script(orc, False), script(osb, True)
    assert got[:3] == ref[:3], (got[:3], ref[:3])
    assert close(got[3], ref[3], rtol=1e-9)
    assert close(got[4], ref[4], rtol=1e-8)


def test_newton_singular_hessian_falls_back_to_the_gradient(osb, orc):
    # newton/mod.rs:43-46: try_inverse fails -> direction = -g, decrement_squared untouched (stays None here)
    n = 6

    def script(m, host):
        fn = _nonconvex_oracle(m, n, True)
        o = m.HostOracle(fn, n, True) if host else m.HostOracle(fn, True)
        s = m.Newton(1e-10, np.full(n, 0.5))
        st = run(m, s, m.BackTracking(1e-4, 0.5), o, 2, 30)
        return st, s.k(), s.x(), s.decrement_squared()

    ref, got = script(orc, False), script(osb, True)
    assert got[:2] == ref[:2], (got[:2], ref[:2])
    assert close(got[2], ref[2])
    assert got[3] is None and ref[3] is None
