"""Objectives and settings of the reference's own inline unit tests and examples (cited per item)."""
import numpy as np

INF = float("inf")


def quad2(gamma, shifted=False, with_hessian=False, FE=None):
    """0.5 (x0^2 + gamma x1^2)  e.g. src/quasi_newton/bfgs_b.rs:167-171, or the shifted
    0.5 ((x0+1)^2 + gamma (x1-1)^2) of src/quasi_newton/bfgs.rs:149-153.  powi(2) == x*x."""
    def f(x):
        if shifted:
            a, b = x[0] + 1.0, x[1] - 1.0
        else:
            a, b = x[0], x[1]
        val = 0.5 * (a * a + gamma * (b * b))
        g = np.array([a, gamma * b])
        if with_hessian:
            return FE(val, g).with_hessian(np.array([[1.0, 0.0], [0.0, gamma]]))
        return (val, g)
    return f


def bfgs_example_3d(x):
    """examples/bfgs_example.rs:11-27"""
    x1, x2, x3 = x
    f = x1 * x1 + 2.0 * (x2 * x2) + 3.0 * (x3 * x3) + x1 * x2 + x2 * x3
    return (f, np.array([2.0 * x1 + x2, 4.0 * x2 + x1 + x3, 6.0 * x3 + x2]))


X0_TESTS = [180.0, 152.0]  # every inline unit test of the reference starts here
