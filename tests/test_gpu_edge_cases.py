"""GPU (-m gpu): edge cases of the path — ragged dimensions (n not a multiple of the 16-double row padding,
of the 8-row tile or of the 1024-column sweep), n = 1, non-finite values, exhausted line searches,
user-supplied device functors, re-entrant minimize, input validation."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_parity import both, close, run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 6, 7, 9, 17, 33, 250, 1001, 1030, 2049])
@pytest.mark.parametrize("kind", ["BFGS", "DFP", "Broyden"])
def test_ragged_dimensions_vs_oracle(osb, orc, n, kind):
    # separable convex quadratic (block size 1: any n), dense kernels with padded columns / partial row tiles
    def script(m):
        obj = m.SeparableQuadratic.generated(n)
        # (Broyden's count is rounding-sensitive near 1e-7, see test_free_running_convex_vs_faithful_oracle)
        s = getattr(m, kind)(1e-5 if kind == "Broyden" else 1e-7, np.linspace(-1.0, 1.0, n) if n > 1 else np.array([0.7]))
        st = run(m, s, m.BackTracking(1e-4, 0.5), obj, 400, 40)
        return st, s.k(), s.termination_reason(), s.x(), s.approx_inv_hessian()

    ref, got = both(osb, orc, script)
    assert got[:3] == ref[:3], (got[:3], ref[:3])
    assert close(got[3], ref[3])
    assert got[4].shape == (n, n)


@pytest.mark.parametrize("n", [7, 250, 1030])
def test_ragged_dimensions_device_engine_and_lazy(osb, n):
    out = []
    for engine, sched, variant in ((1, 0, 0), (2, 0, 0), (2, 1, 0), (2, 1, 1)):
        obj = osb.SeparableQuadratic.generated(n)
        s = osb.BFGS(1e-7, np.linspace(-1.0, 1.0, n)).set_option("engine", engine).set_option("qn_schedule", sched)
        s.set_option("qn_kernel", variant)
        st = run(osb, s, osb.BackTracking(1e-4, 0.5), obj, 400, 40)
        out.append((st, s.k(), s.termination_reason(), s.x()))
    for o in out[1:]:
        assert o[:3] == out[0][:3], (o[:3], out[0][:3])
        assert close(o[3], out[0][3])


def test_dense_quadratic_ragged_and_gd(osb, orc):
    for n in (5, 37, 130):
        def script(m):
            obj = m.DenseQuadratic.generated(n, True)
            s = m.GradientDescent(1e-5, obj.x0)
            st = run(m, s, m.MoreThuente.default(), obj, 500, 20)
            return st, s.k(), s.termination_reason(), s.x()
        ref, got = both(osb, orc, script)
        assert got[:3] == ref[:3] and close(got[3], ref[3])


def test_out_of_domain_and_nonfinite_handling(osb, orc):
    # ls_solver.rs:37-40: NaN / inf objective at x_k -> Err(OutOfDomain)
    def bad(x):
        return (float("nan"), np.array([1.0, 1.0]))
    for m in (orc, osb):
        s = m.BFGS(1e-8, [1.0, 2.0])
        assert run(m, s, m.BackTracking(1e-4, 0.5), bad, 10, 10) == "OutOfDomain"
        assert s.k() == 0
    # backtracking.rs:37-41: an inf trial shrinks the step without consuming an iteration
    def wall(x):
        f = float("inf") if x[0] < 0.25 else (x[0] - 0.3) ** 2 + x[1] * x[1]
        return (f, np.array([2.0 * (x[0] - 0.3), 2.0 * x[1]]))
    res = []
    for m in (orc, osb):
        s = m.GradientDescent(1e-9, [2.0, 1.0])
        st = run(m, s, m.BackTracking(1e-4, 0.5), wall, 200, 3)
        res.append((st, s.k(), s.x()))
    assert res[0][:2] == res[1][:2] and np.array_equal(res[0][2], res[1][2])
    # device functor path: a NaN start is OutOfDomain on both engines
    for engine in (1, 2):
        x0 = np.ones(64)
        x0[5] = np.nan
        s = osb.BFGS(1e-8, x0).set_option("engine", engine)
        assert run(osb, s, osb.BackTracking(1e-4, 0.5), osb.ExtendedRosenbrock(64), 5, 5) == "OutOfDomain"


def test_exhausted_line_search_still_steps(osb, orc):
    # backtracking.rs:53-54: on exhaustion the last (unaccepted, already shrunk) t is returned and the step is taken
    f = lambda x: (x[0] ** 4 + 1e6 * x[1] ** 2, np.array([4.0 * x[0] ** 3, 2e6 * x[1]]))
    res = []
    for m in (orc, osb):
        s = m.GradientDescent(1e-12, [1.0, 1.0])
        st = run(m, s, m.BackTracking(1e-4, 0.5), f, 3, 2)
        res.append((st, s.k(), s.x()))
    assert res[0][0] == res[1][0] == "MaxIterReached" and np.array_equal(res[0][2], res[1][2])


def test_minimize_is_reentrant_and_keeps_state(osb):
    # ls_solver.rs:74: k resets to 0, H / norms carry over; two half runs == one full run
    n = 512
    x0 = np.linspace(-1.2, 1.0, n)
    a = osb.BFGS(1e-8, x0).set_option("engine", 2)
    run(osb, a, osb.BackTracking(1e-4, 0.5), osb.ExtendedRosenbrock(n), 20, 20)
    b = osb.BFGS(1e-8, x0).set_option("engine", 2)
    run(osb, b, osb.BackTracking(1e-4, 0.5), osb.ExtendedRosenbrock(n), 9, 20)
    assert b.k() == 9
    run(osb, b, osb.BackTracking(1e-4, 0.5), osb.ExtendedRosenbrock(n), 11, 20)
    assert b.k() == 11
    assert np.array_equal(a.x(), b.x()) and np.array_equal(a.approx_inv_hessian(), b.approx_inv_hessian())


def test_user_device_functor(osb):
    """osb_objective_create_user: a user-supplied DEVICE functor (here the user's kernel is torch code enqueued on
    the library's stream) drives GD, BFGS and Newton exactly like a built-in objective."""
    import torch
    n = 96
    cvec = torch.linspace(1.0, 3.0, n, dtype=torch.float64, device="cuda")  # (c = 4 makes GD with t = 1/2 oscillate forever)
    avec = torch.linspace(-0.5, 0.5, n, dtype=torch.float64, device="cuda")
    ctx = osb.default_context()
    ext = torch.cuda.ExternalStream(ctx.stream())

    def enqueue(d_x, n_, d_f, d_g, d_h, stream):
        with torch.cuda.stream(ext):
            x = _as_tensor(d_x, n_)
            g = _as_tensor(d_g, n_)
            f = _as_tensor(d_f, 1)
            dlt = x - avec
            g.copy_(cvec * dlt)
            f.copy_((0.5 * cvec * dlt * dlt).sum().reshape(1))
            if d_h:
                ld = (n_ + 15) // 16 * 16
                h = _as_tensor(d_h, n_ * ld).view(n_, ld)
                h.zero_()
                h[:, :n_].copy_(torch.diag(cvec))
        return 0

    def _as_tensor(ptr, count):
        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
        return torch.as_tensor(h, device="cuda")

    obj = osb.UserDeviceObjective(enqueue, n, with_hessian=True)
    e = obj(np.zeros(n))
    assert close(e.g(), (-cvec * avec).cpu().numpy(), rtol=1e-14)
    for cls, tol in (("GradientDescent", 1e-8), ("BFGS", 1e-8), ("Newton", 1e-10)):
        s = getattr(osb, cls)(tol, np.zeros(n))
        assert run(osb, s, osb.BackTracking(1e-4, 0.5), obj, 500, 40) == "Ok"
        assert close(s.x(), avec.cpu().numpy(), rtol=1e-7, atol=1e-7)


def test_input_validation(osb):
    with pytest.raises(osb.ErrorInputParams):
        osb.ExtendedRosenbrock(7)
    with pytest.raises(osb.ErrorInputParams):
        osb.BFGSB(1e-8, [1.0, 2.0], None, None)
    s = osb.BFGS(1e-8, [1.0, 2.0, 3.0])
    with pytest.raises(osb.ErrorInputParams):
        s.minimize(osb.BackTracking(1e-4, 0.5), osb.ExtendedRosenbrock(4), 5, 5)
    with pytest.raises(osb.ErrorInputParams):
        osb.GLLQuadratic(1e-4, 1000)._h()
    with pytest.raises(osb.DeviceError):
        osb.GradientDescent(1e-8, [1.0, 2.0]).set_option("engine", 2).minimize(osb.BackTracking(1e-4, 0.5), osb.ExtendedRosenbrock(2), 5, 5)
