"""CPU: pins the oracle (oracle/oracle.cpp) against every known-answer check the reference holds
for the path — the exact assert of examples/quadratic.rs:43 and the |f| < 1e-6 / |x0| < 1e-6 asserts
of the inline unit tests — plus the hand-traced expectations table of SURVEY.md §4 (iteration
counts, termination reasons, step lengths)."""
import numpy as np
import pytest

from problems import INF, X0_TESTS, bfgs_example_3d, quad2


def run(solver, ls, oracle, mi, ml, O):
    try:
        solver.minimize(ls, oracle, mi, ml)
        return "Ok"
    except O.SolverError as e:
        return type(e).__name__


def test_examples_quadratic_rs_exact(orc):
    # examples/quadratic.rs:10-43: BFGS + MoreThuente on f = x^T I x, x0 = (1,1); assert_eq!(f, 0.0)
    obj = orc.DenseQuadratic(np.eye(2))
    s = orc.BFGS(1e-6, [1.0, 1.0])
    assert run(s, orc.MoreThuente.default(), obj, 100, 10, orc) == "Ok"
    assert s.k() == 2 and s.termination_reason() == "grad_tol"
    assert obj(s.x()).f() == 0.0
    assert np.all(s.x() == 0.0)
    assert s.trace()["t"][0] == 0.49995


def test_examples_bfgs_example_rs(orc):
    s = orc.BFGS(1e-8, [1.0, 1.0, 1.0])
    assert run(s, orc.MoreThuente.default(), bfgs_example_3d, 50, 20, orc) == "Ok"
    assert s.k() == 4 and s.termination_reason() == "grad_tol"
    t = s.trace()["t"]
    assert t[0] == 0.16317812500000003 and t[1] == 0.27995646245544603 and t[2] == 1.0 and t[3] == 1.0
    assert bfgs_example_3d(s.x())[0] < 1e-20


@pytest.mark.parametrize("cls", ["BFGS", "DFP", "Broyden"])
@pytest.mark.parametrize("ls", ["mt", "bt"])
def test_qn_unit_tests(orc, cls, ls):
    # bfgs.rs:141-188,190-239; dfp.rs:136-183,185-234; broyden.rs:134-181,183-232 : |f| < 1e-6
    f = quad2(1.0, shifted=True)
    s = getattr(orc, cls)(1e-12, X0_TESTS)
    lsearch = orc.MoreThuente.default() if ls == "mt" else orc.BackTracking(1e-4, 0.5)
    assert run(s, lsearch, f, 1000, 100000, orc) == "Ok"
    assert abs(f(s.x())[0]) < 1e-6
    assert s.k() == 1 and np.all(s.x() == np.array([-1.0, 1.0]))


@pytest.mark.parametrize("cls", ["DFPB", "BroydenB", "SR1B", "BFGSB"])
def test_bounded_qn_unit_tests(orc, cls):
    # dfp_b.rs:216, broyden_b.rs:215, sr1_b.rs:211 (|f| < 1e-6); bfgs_b.rs:160-212 (print only)
    lb, ub = [-INF, -INF], [INF, INF]
    gamma = 999.0 if cls == "BFGSB" else 1.0
    f = quad2(gamma)
    s = getattr(orc, cls)(1e-12, X0_TESTS, lb, ub)
    assert run(s, orc.BackTrackingB(1e-4, 0.5, lb, ub), f, 10000, 1000, orc) == "Ok"
    assert abs(f(s.x())[0]) < 1e-6
    assert s.k() == (4 if cls == "BFGSB" else 1)


def test_gradient_descent_unit_tests(orc):
    # gradient_descent.rs:86-130 (MoreThuente, |f| < 1e-6 holds); :133-179 (BackTracking): the
    # reference's own .unwrap() would panic — MaxIterReached after 1000 iterations
    s = orc.GradientDescent(1e-12, X0_TESTS)
    assert run(s, orc.MoreThuente.default(), quad2(90.0), 1000, 100, orc) == "Ok"
    assert s.k() == 6 and abs(quad2(90.0)(s.x())[0]) < 1e-6
    s = orc.GradientDescent(1e-12, X0_TESTS)
    assert run(s, orc.BackTracking(1e-4, 0.5), quad2(90.0), 1000, 100, orc) == "MaxIterReached"
    assert s.k() == 1000 and abs(quad2(90.0)(s.x())[0]) < 1e-6


def test_pnorm_descent_unit_tests(orc):
    # pnorm_descent.rs:90-141 (MoreThuente) and :144-194 (BackTracking): inverse_p is the exact inverse Hessian of
    # 0.5 (x0^2 + 90 x1^2), so one unit step lands on the minimiser exactly; both asserts |f| < 1e-6 hold
    P = [[1.0, 0.0], [0.0, 1.0 / 90.0]]
    for ls in (orc.MoreThuente.default(), orc.BackTracking(1e-4, 0.5)):
        s = orc.PnormDescent(1e-12, X0_TESTS, P)
        assert run(s, ls, quad2(90.0), 1000, 100, orc) == "Ok"
        assert s.k() == 1 and s.termination_reason() == "grad_tol" and quad2(90.0)(s.x())[0] == 0.0
    # with inverse_p = I the solver is GradientDescent (pnorm_descent.rs:9): same 6 iterations as gradient_descent.rs:86-130
    s = orc.PnormDescent(1e-12, X0_TESTS, [[1.0, 0.0], [0.0, 1.0]])
    g = orc.GradientDescent(1e-12, X0_TESTS)
    assert run(s, orc.MoreThuente.default(), quad2(90.0), 1000, 100, orc) == run(g, orc.MoreThuente.default(), quad2(90.0), 1000, 100, orc)
    assert s.k() == g.k() == 6 and (s.x() == g.x()).all()


def test_projected_and_spectral_unit_tests(orc):
    lb, ub = [-INF, -INF], [INF, INF]
    s = orc.ProjectedGradientDescent(1e-6, X0_TESTS, lb, ub)  # projected_gradient_descent.rs:114-165
    assert run(s, orc.BackTrackingB(1e-4, 0.5, lb, ub), quad2(999.0), 10000, 1000, orc) == "Ok"
    assert s.k() == 9340 and s.termination_reason() == "proj_grad_tol"
    lb = [-1.0, 47.0]
    s = orc.SpectralProjectedGradient(1e-12, X0_TESTS, quad2(1e9), lb, ub)  # spg.rs:151-204
    assert run(s, orc.GLLQuadratic(1e-4, 10), quad2(1e9), 10000, 1000, orc) == "Ok"
    assert s.k() == 3 and np.all(s.x() == np.array([0.0, 47.0])) and quad2(1e9)(s.x())[0] == 1.1045e12
    assert np.all(s.active_set() == np.array([0, 1]))
    qh = quad2(1e9, with_hessian=True, FE=orc.FuncEvalMultivariate)
    s = orc.SpectralProjectedNewton(1e-12, X0_TESTS, qh, lb, ub)  # spn.rs:156-210
    assert run(s, orc.GLLQuadratic(1e-4, 10), orc.HostOracle(qh, True), 10000, 1000, orc) == "Ok"
    assert s.k() == 1171 and np.all(s.x() == np.array([0.0, 47.0]))
    qh = quad2(999.0, with_hessian=True, FE=orc.FuncEvalMultivariate)
    s = orc.ProjectedNewton(1e-6, X0_TESTS, [-INF, -INF], ub)  # projected_newton.rs:147-198
    assert run(s, orc.BackTrackingB(1e-4, 0.5, [-INF, -INF], ub), orc.HostOracle(qh, True), 10000, 1000, orc) == "Ok"
    assert s.k() == 1 and abs(s.x()[0]) < 1e-9 and abs(s.x()[1]) < 1e-9


@pytest.mark.parametrize("ls", ["mt", "bt"])
def test_newton_unit_tests(orc, ls):
    # newton/mod.rs:77-118,121-163: |f| < 1e-6; converges by decrement at k = 2 (never at k = 0)
    qh = quad2(1222.0, with_hessian=True, FE=orc.FuncEvalMultivariate)
    s = orc.Newton(1e-8, [1.0, 1.0])
    lsearch = orc.MoreThuente.default() if ls == "mt" else orc.BackTracking(1e-4, 0.5)
    assert run(s, lsearch, orc.HostOracle(qh, True), 1000, 100, orc) == "Ok"
    assert s.k() == 2 and s.termination_reason() == "newton_decrement"
    assert qh(s.x()).f() == 0.0


def test_line_search_unit_tests(orc):
    # backtracking.rs:63-113, morethuente.rs:303-352, morethuente_b.rs:330-379: hand-rolled GD loop, |x0| < 1e-6
    f = quad2(90.0)
    for ls in (orc.BackTracking(1e-4, 0.5), orc.MoreThuente.default(), orc.MoreThuenteB(2)):
        x = np.array(X0_TESTS)
        k = 1
        while 1000 > k:
            val, g = f(x)
            if g @ g < 1e-12:
                break
            d = -g
            t = ls.compute_step_len(x, d, f, 1000)
            x = x + t * d
            k += 1
        assert abs(x[0]) < 1e-6


def test_rank2_form_matches_faithful_on_convex(orc):
    # the O(n^2) rank-2 form and the reference's triple product agree on convex problems
    # (tol above the objective's rounding floor, so that termination is not decided by noise)
    n = 48
    res = []
    for form in ("faithful", "rank2"):
        obj = orc.DenseQuadratic.generated(n, shifted=True)
        s = orc.BFGS(1e-6, obj.x0).set_update_form(form)
        assert run(s, orc.BackTracking(1e-4, 0.5), obj, 500, 50, orc) == "Ok"
        res.append((s.k(), s.termination_reason(), s.x()))
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    assert np.allclose(res[0][2], res[1][2], rtol=1e-9, atol=1e-12)


def test_small_inverses(orc):
    # try_inverse closed forms (n <= 4) and LU (n >= 5) of the nalgebra restatement vs numpy
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 4, 5, 9):
        A = rng.standard_normal((n, n))
        H = A @ A.T + n * np.eye(n)
        g = rng.standard_normal(n)

        def orac(x, H=H, g=g):
            return orc.FuncEvalMultivariate(0.5 * x @ H @ x - g @ x, H @ x - g).with_hessian(H)
        s = orc.Newton(1e-20, np.zeros(n))
        try:
            s.minimize(orc.NoSearch(), orc.HostOracle(orac, True), 1, 1)
        except orc.MaxIterReached:
            pass
        assert np.allclose(s.x(), np.linalg.solve(H, g), rtol=1e-10, atol=1e-12)


def _load_golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_trajectories.json")) as f:
        return json.load(f)


def test_oracle_reproduces_the_committed_vectors(orc):
    """tests/golden/oracle_trajectories.json (tools/make_golden.py): the oracle of today gives the committed outcomes —
    bit for bit where only n <= 5 vector arithmetic is involved, to 1e-12 where the host's dot / GEMV code runs."""
    import numpy as np
    from golden_cases import BIT_EXACT, CASES
    gold = _load_golden()
    assert set(gold) == set(CASES)
    for name, script in CASES.items():
        r, g = script(orc), gold[name]
        assert (r["status"], int(r["k"]), r["reason"]) == (g["status"], g["k"], g["reason"]), name
        x, gx = np.asarray(r["x"]), np.asarray(g["x"])
        if name in BIT_EXACT:
            assert np.array_equal(x, gx), name
        else:
            assert np.all(np.abs(x - gx) <= 1e-12 * max(1.0, float(np.max(np.abs(gx))))), name
        if g["active_set"] is not None:
            assert np.array_equal(np.asarray(r["active_set"]), np.asarray(g["active_set"])), name
