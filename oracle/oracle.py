"""ctypes front-end of the CPU oracle (oracle/oracle.cpp) — TEST INFRASTRUCTURE ONLY.

Mirrors the reference crate's public names (BFGS::new(tol, x0), MoreThuente::default().with_c1(..),
solver.minimize(&mut ls, oracle, max_iter_solver, max_iter_line_search, callback), solver.x(), ...;
ls_solver.rs:23-112, line_search/mod.rs:14-23) so that a parity test can run the same script against
this oracle and against the CUDA library.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
HOST_EVAL = C.CFUNCTYPE(C.c_int, C.c_void_p, _dp, C.c_int64, _dp, _dp, _dp)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, i64, dbl, ci = C.c_void_p, C.c_int64, C.c_double, C.c_int
        sig = {
            "orc_set_threads": (None, [ci]), "orc_max_threads": (ci, []),
            "orc_obj_host": (vp, [HOST_EVAL, vp, ci]),
            "orc_obj_dense_quadratic": (vp, [i64, _dp, _dp]),
            "orc_obj_dense_quadratic_generated": (vp, [i64, ci, _dp]),
            "orc_obj_rosenbrock": (vp, []),
            "orc_obj_separable_quadratic_generated": (vp, [i64]),
            "orc_obj_logistic_generated": (vp, [i64, i64, dbl, ci]),
            "orc_obj_destroy": (None, [vp]), "orc_obj_calls": (i64, [vp]),
            "orc_obj_eval": (ci, [vp, i64, _dp, _dp, _dp, _dp]),
            "orc_ls_backtracking": (vp, [dbl, dbl]),
            "orc_ls_backtracking_b": (vp, [dbl, dbl, i64, _dp, _dp]),
            "orc_ls_morethuente": (vp, [dbl, dbl, dbl, dbl, dbl]),
            "orc_ls_morethuente_b": (vp, [dbl, dbl, dbl, dbl, dbl, i64, _dp, _dp]),
            "orc_ls_gll": (vp, [dbl, i64, dbl, dbl]), "orc_ls_nosearch": (vp, []),
            "orc_ls_destroy": (None, [vp]), "orc_ls_t_max": (dbl, [vp]),
            "orc_ls_compute_step_len": (dbl, [vp, vp, i64, _dp, _dp, i64]),
            "orc_solver_create": (vp, [ci, i64, dbl, _dp, _dp, _dp, vp]),
            "orc_solver_destroy": (None, [vp]), "orc_solver_set_form": (None, [vp, ci]),
            "orc_solver_set_lambdas": (None, [vp, dbl, dbl]),
            "orc_solver_record_iterates": (None, [vp, ci]),
            "orc_minimize": (ci, [vp, vp, vp, i64, i64]),
            "orc_solver_k": (i64, [vp]), "orc_solver_reason": (ci, [vp]),
            "orc_solver_x": (None, [vp, _dp]), "orc_solver_set_x": (None, [vp, _dp]),
            "orc_solver_s_norm": (dbl, [vp]), "orc_solver_y_norm": (dbl, [vp]),
            "orc_solver_clear_norms": (None, [vp]),
            "orc_solver_lambda": (dbl, [vp]), "orc_solver_decrement_squared": (dbl, [vp]),
            "orc_solver_inv_hessian": (ci, [vp, _dp]), "orc_solver_set_inv_hessian": (ci, [vp, _dp]),
            "orc_solver_active_set": (ci, [vp, C.POINTER(C.c_uint8)]),
            "orc_solver_trace_len": (i64, [vp]), "orc_solver_trace": (None, [vp, _dp, _dp, _dp, _dp]),
            "orc_solver_iterate": (None, [vp, i64, _dp]),
            "orc_time_bfgs_update_rowsample": (dbl, [i64, i64, ci]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _arr(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


# ---- ls_solver.rs:10-20
class SolverError(Exception):
    pass


class MaxIterReached(SolverError):
    pass


class OutOfDomain(SolverError):
    pass


class ErrorInputParams(SolverError):
    pass


class AbnormalTermination(SolverError):
    pass


_STATUS = {1: MaxIterReached, 2: OutOfDomain, 3: ErrorInputParams, 4: AbnormalTermination}
REASONS = {0: None, 1: "grad_tol", 2: "s_norm", 3: "y_norm", 4: "proj_grad_tol", 5: "newton_decrement"}


class FuncEvalMultivariate:
    """func_eval.rs:5-41"""

    def __init__(self, f, g, hessian=None):
        self._f, self._g, self._h = float(f), _arr(g), hessian

    @staticmethod
    def new(f, g):
        return FuncEvalMultivariate(f, g)

    def with_hessian(self, h):
        self._h = np.asarray(h, dtype=np.float64)
        return self

    def f(self):
        return self._f

    def g(self):
        return self._g

    def hessian(self):
        return self._h


# ---- objectives
class _Objective:
    handle = None

    def calls(self):
        return lib().orc_obj_calls(self.handle)

    def __call__(self, x):
        x = _arr(x)
        n = x.size
        f = C.c_double()
        g = np.empty(n)
        h = np.empty((n, n))
        has_h = lib().orc_obj_eval(self.handle, n, _p(x), C.byref(f), _p(g), _p(h))
        e = FuncEvalMultivariate(f.value, g)
        if has_h:
            e.with_hessian(h.T.copy())  # column-major -> numpy row-major
        return e

    def __del__(self):
        if self.handle and _lib is not None:
            _lib.orc_obj_destroy(self.handle)
            self.handle = None


class HostOracle(_Objective):
    """A user closure FnMut(&DVector) -> FuncEvalMultivariate (ls_solver.rs:34)."""

    def __init__(self, fn, with_hessian=False):
        self.fn = fn

        def tramp(_user, xp, n, fp, gp, hp):
            x = np.ctypeslib.as_array(xp, shape=(n,)).copy()
            r = fn(x)
            if isinstance(r, tuple):
                r = FuncEvalMultivariate(*r)
            fp[0] = r.f()
            np.ctypeslib.as_array(gp, shape=(n,))[:] = r.g()
            if hp and r.hessian() is not None:
                # column-major out
                np.ctypeslib.as_array(hp, shape=(n, n))[:] = np.asarray(r.hessian()).T
                return 1
            return 0

        self._cb = HOST_EVAL(tramp)
        self.handle = lib().orc_obj_host(self._cb, None, 1 if with_hessian else 0)


class DenseQuadratic(_Objective):
    """f = x.(A x) [- 2 b.x], g = 2 A x [- 2 b]  (examples/quadratic.rs:10-14 pattern)."""

    def __init__(self, A, b=None):
        A = np.asarray(A, dtype=np.float64)
        n = A.shape[0]
        Ac = np.ascontiguousarray(A.T)  # column-major bytes
        bb = _arr(b) if b is not None else None
        self.handle = lib().orc_obj_dense_quadratic(n, _p(Ac), _p(bb))

    @classmethod
    def generated(cls, n, shifted=True):
        self = cls.__new__(cls)
        self.x0 = np.empty(n)
        self.handle = lib().orc_obj_dense_quadratic_generated(n, 1 if shifted else 0, _p(self.x0))
        return self


class ExtendedRosenbrock(_Objective):
    def __init__(self, n=None):
        self.handle = lib().orc_obj_rosenbrock()


class SeparableQuadratic(_Objective):
    @classmethod
    def generated(cls, n):
        self = cls.__new__(cls)
        self.handle = lib().orc_obj_separable_quadratic_generated(n)
        return self


class LogisticRegression(_Objective):
    @classmethod
    def generated(cls, m, n, lam=1.0, with_hessian=True):
        self = cls.__new__(cls)
        self.handle = lib().orc_obj_logistic_generated(m, n, lam, 1 if with_hessian else 0)
        return self


def _as_objective(o, with_hessian=False):
    if isinstance(o, _Objective):
        return o
    return HostOracle(o, with_hessian=with_hessian)


# ---- line searches
class _LS:
    handle = None

    def _make(self):
        raise NotImplementedError

    def _h(self):
        if self.handle is None:
            self.handle = self._make()
        return self.handle

    def compute_step_len(self, x_k, direction_k, oracle, max_iter):
        """line_search/mod.rs:14-23 (eval_x_k is recomputed from the oracle here)."""
        o = _as_objective(oracle)
        x, d = _arr(x_k), _arr(direction_k)
        return lib().orc_ls_compute_step_len(self._h(), o.handle, x.size, _p(x), _p(d), max_iter)

    def __del__(self):
        if self.handle and _lib is not None:
            _lib.orc_ls_destroy(self.handle)
            self.handle = None


class BackTracking(_LS):
    def __init__(self, c1, beta):
        self.c1, self.beta = c1, beta

    @staticmethod
    def new(c1, beta):
        return BackTracking(c1, beta)

    def _make(self):
        return lib().orc_ls_backtracking(self.c1, self.beta)


class BackTrackingB(_LS):
    def __init__(self, c1, beta, lower_bound, upper_bound):
        self.c1, self.beta, self.lb, self.ub = c1, beta, _arr(lower_bound), _arr(upper_bound)

    new = classmethod(lambda cls, *a: cls(*a))

    def _make(self):
        return lib().orc_ls_backtracking_b(self.c1, self.beta, self.lb.size, _p(self.lb), _p(self.ub))


class MoreThuente(_LS):
    """morethuente.rs:16-62"""

    def __init__(self):
        self.c1, self.c2, self.t_min, self.t_max = 1e-4, 0.9, 0.0, float("inf")
        self.delta_min, self.delta, self.delta_max = 0.58333333, 0.66, 1.1

    default = classmethod(lambda cls: cls())

    def with_deltas(self, delta_min, delta, delta_max):
        self.delta_min, self.delta, self.delta_max = delta_min, delta, delta_max
        return self

    def with_t_min(self, t):
        self.t_min = t
        return self

    def with_t_max(self, t):
        self.t_max = t
        return self

    def with_c1(self, c1):
        assert c1 > 0.0, "c1 must be positive"
        assert c1 < self.c2, "c1 must be less than c2"
        self.c1 = c1
        return self

    def with_c2(self, c2):
        assert c2 > 0.0, "c2 must be positive"
        assert c2 < 1.0, "c2 must be less than 1"
        assert c2 > self.c1, "c2 must be greater than c1"
        self.c2 = c2
        return self

    def _make(self):
        return lib().orc_ls_morethuente(self.c1, self.c2, self.t_min, self.t_max, self.delta)


class MoreThuenteB(MoreThuente):
    """morethuente_b.rs:17-40"""

    def __init__(self, n):
        super().__init__()
        self.lb, self.ub = np.full(n, -np.inf), np.full(n, np.inf)

    new = classmethod(lambda cls, n: cls(n))

    def with_lower_bound(self, lb):
        self.lb = _arr(lb)
        return self

    def with_upper_bound(self, ub):
        self.ub = _arr(ub)
        return self

    def _make(self):
        return lib().orc_ls_morethuente_b(self.c1, self.c2, self.t_min, self.t_max, self.delta,
                                          self.lb.size, _p(self.lb), _p(self.ub))

    def current_t_max(self):
        return lib().orc_ls_t_max(self._h())


class GLLQuadratic(_LS):
    def __init__(self, c1, m):
        self.c1, self.m, self.sigma1, self.sigma2 = c1, m, 0.1, 0.9

    new = classmethod(lambda cls, *a: cls(*a))

    def with_sigmas(self, s1, s2):
        self.sigma1, self.sigma2 = s1, s2
        return self

    def _make(self):
        return lib().orc_ls_gll(self.c1, self.m, self.sigma1, self.sigma2)


class NoSearch(_LS):
    def _make(self):
        return lib().orc_ls_nosearch()


# ---- solvers
_KIND = dict(GD=0, PGD=1, SPG=2, BFGS=3, DFP=4, BROYDEN=5, BFGSB=6, DFPB=7, BROYDENB=8, SR1B=9, NEWTON=10,
             PROJ_NEWTON=11, SPN=12, PNORM=13)


class _Solver:
    KIND = None
    NEEDS_HESSIAN = False

    def __init__(self, tol, x0, lower_bound=None, upper_bound=None, oracle=None):
        x0 = _arr(x0)
        self.n = x0.size
        self._tol = tol
        self._lb = _arr(lower_bound) if lower_bound is not None else None
        self._ub = _arr(upper_bound) if upper_bound is not None else None
        self._keep = _as_objective(oracle, self.NEEDS_HESSIAN) if oracle is not None else None
        self.handle = lib().orc_solver_create(_KIND[self.KIND], self.n, tol, _p(x0), _p(self._lb), _p(self._ub),
                                              self._keep.handle if self._keep is not None else None)
        assert self.handle

    @classmethod
    def new(cls, *a, **k):
        return cls(*a, **k)

    def minimize(self, line_search, oracle, max_iter_solver, max_iter_line_search, callback=None):
        assert callback is None, "the oracle front-end has no callback; use record_iterates()"
        o = _as_objective(oracle, self.NEEDS_HESSIAN)
        st = lib().orc_minimize(self.handle, line_search._h(), o.handle, max_iter_solver, max_iter_line_search)
        self.status = st
        if st == 0:
            return None
        if st in _STATUS:
            raise _STATUS[st]()
        raise RuntimeError("oracle: reference would panic (status %d)" % st)

    def x(self):
        out = np.empty(self.n)
        lib().orc_solver_x(self.handle, _p(out))
        return out

    xk = x

    def set_x(self, x):
        x = _arr(x)
        lib().orc_solver_set_x(self.handle, _p(x))

    def k(self):
        return lib().orc_solver_k(self.handle)

    def tol(self):
        return self._tol

    grad_tol = tol

    def termination_reason(self):
        return REASONS[lib().orc_solver_reason(self.handle)]

    def _opt(self, v):
        return None if np.isnan(v) else v

    def s_norm(self):
        return self._opt(lib().orc_solver_s_norm(self.handle))

    def y_norm(self):
        return self._opt(lib().orc_solver_y_norm(self.handle))

    def clear_norms(self):
        lib().orc_solver_clear_norms(self.handle)

    def lambda_(self):
        return lib().orc_solver_lambda(self.handle)

    def with_lambdas(self, lmin, lmax):
        lib().orc_solver_set_lambdas(self.handle, lmin, lmax)
        return self

    def decrement_squared(self):
        return self._opt(lib().orc_solver_decrement_squared(self.handle))

    def approx_inv_hessian(self):
        out = np.empty((self.n, self.n))
        assert lib().orc_solver_inv_hessian(self.handle, _p(out)) == 0
        return out

    def set_approx_inv_hessian(self, H):
        H = _arr(H)
        assert lib().orc_solver_set_inv_hessian(self.handle, _p(H)) == 0

    def set_update_form(self, form):
        lib().orc_solver_set_form(self.handle, {"faithful": 0, "rank2": 1}[form])
        return self

    def active_set(self):
        out = np.zeros(self.n, dtype=np.uint8)
        assert lib().orc_solver_active_set(self.handle, out.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
        return out

    def lower_bound(self):
        return self._lb

    def upper_bound(self):
        return self._ub

    def projected_gradient(self, ev):
        """ls_solver.rs:121-133"""
        x, pg = self.x(), ev.g().copy()
        m = ((x == self._lb) & (pg > 0.0)) | ((x == self._ub) & (pg < 0.0))
        pg[m] = 0.0
        return pg

    def record_iterates(self, on=True):
        lib().orc_solver_record_iterates(self.handle, 1 if on else 0)
        return self

    def trace(self):
        m = lib().orc_solver_trace_len(self.handle)
        f, t, sn, yn = (np.empty(m) for _ in range(4))
        lib().orc_solver_trace(self.handle, _p(f), _p(t), _p(sn), _p(yn))
        return dict(f=f, t=t, s_norm=sn, y_norm=yn)

    def iterate(self, idx):
        out = np.empty(self.n)
        lib().orc_solver_iterate(self.handle, idx, _p(out))
        return out

    def __del__(self):
        if getattr(self, "handle", None) and _lib is not None:
            _lib.orc_solver_destroy(self.handle)
            self.handle = None


def _mk(name, kind, bounded=False, needs_oracle=False, needs_h=False):
    if needs_oracle:
        def __init__(self, tol, x0, oracle, lower_bound, upper_bound):
            _Solver.__init__(self, tol, x0, lower_bound, upper_bound, oracle)
    elif bounded:
        def __init__(self, tol, x0, lower_bound, upper_bound):
            _Solver.__init__(self, tol, x0, lower_bound, upper_bound)
    else:
        def __init__(self, tol, x0):
            _Solver.__init__(self, tol, x0)
    return type(name, (_Solver,), dict(KIND=kind, NEEDS_HESSIAN=needs_h, __init__=__init__))


GradientDescent = _mk("GradientDescent", "GD")
ProjectedGradientDescent = _mk("ProjectedGradientDescent", "PGD", bounded=True)
SpectralProjectedGradient = _mk("SpectralProjectedGradient", "SPG", needs_oracle=True)
BFGS = _mk("BFGS", "BFGS")
DFP = _mk("DFP", "DFP")
Broyden = _mk("Broyden", "BROYDEN")
BFGSB = _mk("BFGSB", "BFGSB", bounded=True)
DFPB = _mk("DFPB", "DFPB", bounded=True)
BroydenB = _mk("BroydenB", "BROYDENB", bounded=True)
SR1B = _mk("SR1B", "SR1B", bounded=True)
Newton = _mk("Newton", "NEWTON", needs_h=True)
ProjectedNewton = _mk("ProjectedNewton", "PROJ_NEWTON", bounded=True, needs_h=True)
SpectralProjectedNewton = _mk("SpectralProjectedNewton", "SPN", needs_oracle=True, needs_h=True)


class PnormDescent(_Solver):
    """PnormDescent::new(grad_tol, x0, inverse_p) (pnorm_descent.rs:23-30); inverse_p is an n x n array, [i, j] = row i, column j."""
    KIND = "PNORM"

    def __init__(self, tol, x0, inverse_p):
        _Solver.__init__(self, tol, x0)
        P = np.ascontiguousarray(np.asarray(inverse_p, dtype=np.float64))
        assert P.shape == (self.n, self.n)
        self.set_approx_inv_hessian(P)


def time_bfgs_update_rowsample(n, rows, threads=1):
    return lib().orc_time_bfgs_update_rowsample(n, rows, threads)
