"""optimization-solvers_b200 — B200-native (sm_100a) line-search solvers behind the crate's own API.

Host-side mirror (Python over the C ABI of include/optsolv_b200.h) of the reference crate's public
surface for the hot path: the solver structs with `new(tol, x0[, lb, ub])`, the line searches
(`BackTracking::new(c1, beta)`, `MoreThuente::default().with_c1(..)`, ...), `FuncEvalMultivariate`,
`SolverError`, `Tracer`, and `solver.minimize(&mut ls, oracle, max_iter_solver, max_iter_line_search,
callback)` (src/ls_solver.rs:66-111).  Names, argument meaning and error behaviour follow the crate, so
a test written against the crate reads the same here.  The directory name is not a Python identifier;
load it with `importlib.import_module("optimization-solvers_b200")`.

There is NO CPU fallback: every compute call goes through libosb_b200.so (hand-written CUDA for sm_100a)
and fails loudly when the library or a GPU is missing.
"""
import ctypes as C
import logging
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libosb_b200.so")
_lib = None
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
HOST_EVAL = C.CFUNCTYPE(C.c_int, _vp, _dp, C.c_int64, _dp, _dp, _dp)
DEVICE_EVAL = C.CFUNCTYPE(C.c_int, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp)
CALLBACK = C.CFUNCTYPE(None, _vp, _vp)
PHI_EVAL = C.CFUNCTYPE(None, _vp, C.c_double, C.c_int, _dp, _dp, _dp)
LOG_FN = C.CFUNCTYPE(None, _vp, C.c_int, C.c_char_p, C.c_char_p)

log = logging.getLogger("optimization_solvers")


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree (nvcc, sm_100a)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_osb_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


def lib():
    """The C-ABI library.  Raises if it has not been built — there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libosb_b200.so is missing: run `python optimization-solvers_b200/build.py` "
                               "(the CUDA extension is mandatory; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        i64, dbl, ci, pp = C.c_int64, C.c_double, C.c_int, C.POINTER(_vp)
        i32p = C.POINTER(C.c_int32)
        sig = {
            "osb_last_error_string": (C.c_char_p, []), "osb_version": (C.c_char_p, []),
            "osb_set_log_callback": (ci, [LOG_FN, _vp]),
            "osb_ctx_create": (ci, [ci, pp]), "osb_nccl_unique_id": (ci, [_vp]),
            "osb_ctx_create_dist": (ci, [ci, ci, ci, _vp, pp]), "osb_ctx_destroy": (None, [_vp]),
            "osb_ctx_ipc_handle": (ci, [_vp, _vp]), "osb_ctx_ipc_connect": (ci, [_vp, _vp]), "osb_ctx_ipc_close": (ci, [_vp]),
            "osb_ctx_rank": (ci, [_vp]), "osb_ctx_world": (ci, [_vp]), "osb_ctx_synchronize": (ci, [_vp]),
            "osb_ctx_stream": (_vp, [_vp]), "osb_ctx_counters": (ci, [_vp, C.POINTER(i64)]),
            "osb_objective_create_dense_quadratic": (ci, [_vp, i64, _dp, _dp, pp]),
            "osb_objective_create_dense_quadratic_generated": (ci, [_vp, i64, ci, _dp, pp]),
            "osb_objective_create_rosenbrock": (ci, [_vp, i64, pp]),
            "osb_objective_create_separable_quadratic_generated": (ci, [_vp, i64, pp]),
            "osb_objective_create_separable_quadratic_generated_shard": (ci, [_vp, i64, i64, pp]),
            "osb_ctx_set_vector_sharding": (ci, [_vp, ci]), "osb_ctx_trim_memory": (ci, [_vp]),
            "osb_sym_layout": (ci, [i64, ci, ci, i64, C.POINTER(ci), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
            "osb_objective_create_logistic_generated": (ci, [_vp, i64, i64, dbl, pp]),
            "osb_objective_create_host": (ci, [_vp, i64, HOST_EVAL, _vp, ci, pp]),
            "osb_objective_create_user": (ci, [_vp, i64, DEVICE_EVAL, _vp, ci, pp]),
            "osb_objective_eval": (ci, [_vp, _dp, _dp, _dp, _dp]),
            "osb_objective_calls": (i64, [_vp]), "osb_objective_dim": (i64, [_vp]),
            "osb_objective_destroy": (None, [_vp]),
            "osb_linesearch_create_backtracking": (ci, [dbl, dbl, pp]),
            "osb_linesearch_create_backtracking_b": (ci, [_vp, dbl, dbl, i64, _dp, _dp, pp]),
            "osb_linesearch_create_morethuente": (ci, [dbl, dbl, dbl, dbl, dbl, dbl, dbl, pp]),
            "osb_linesearch_create_morethuente_b": (ci, [_vp, dbl, dbl, dbl, dbl, dbl, dbl, dbl, i64, _dp, _dp, pp]),
            "osb_linesearch_create_gll_quadratic": (ci, [dbl, i64, dbl, dbl, pp]),
            "osb_linesearch_create_nosearch": (ci, [pp]),
            "osb_linesearch_t_max": (dbl, [_vp]), "osb_linesearch_destroy": (None, [_vp]),
            "osb_linesearch_compute_step_len": (ci, [_vp, _vp, _vp, _dp, _dp, i64, _dp]),
            "osb_linesearch_step_len_scalar": (ci, [_vp, PHI_EVAL, _vp, dbl, dbl, dbl, i64, _dp, C.POINTER(ci)]),
            "osb_solver_create": (ci, [_vp, ci, i64, dbl, _dp, _dp, _dp, _vp, pp]),
            "osb_solver_destroy": (None, [_vp]),
            "osb_minimize": (ci, [_vp, _vp, _vp, i64, i64, CALLBACK, _vp]),
            "osb_solver_set_option": (ci, [_vp, C.c_char_p, i64]),
            "osb_solver_set_lambdas": (ci, [_vp, dbl, dbl]),
            "osb_solver_k": (i64, [_vp]), "osb_solver_dim": (i64, [_vp]),
            "osb_solver_termination_reason": (ci, [_vp]),
            "osb_solver_x": (ci, [_vp, _dp]), "osb_solver_set_x": (ci, [_vp, _dp]),
            "osb_solver_f": (ci, [_vp, _dp]), "osb_solver_grad": (ci, [_vp, _dp]),
            "osb_solver_s_norm": (dbl, [_vp]), "osb_solver_y_norm": (dbl, [_vp]),
            "osb_solver_clear_norms": (ci, [_vp]),
            "osb_solver_lambda": (dbl, [_vp]), "osb_solver_decrement_squared": (dbl, [_vp]),
            "osb_solver_inv_hessian": (ci, [_vp, _dp]), "osb_solver_set_inv_hessian": (ci, [_vp, _dp]),
            "osb_solver_active_set": (ci, [_vp, C.POINTER(C.c_uint8)]),
            "osb_solver_trace_len": (i64, [_vp]), "osb_solver_trace": (ci, [_vp, _dp, _dp, _dp, _dp]),
            "osb_solver_last_timing": (ci, [_vp, _dp, C.POINTER(i64)]),
            "osb_solver_kernel_timing": (ci, [_vp, _dp]),
            "osb_solver_path_info": (ci, [_vp, C.POINTER(i64)]),
            "osb_solver_iter_profile": (ci, [_vp, _dp]),
            "osb_batched_bfgs_rosenbrock": (ci, [_vp, i64, i64, _dp, dbl, i64, i64, dbl, dbl, _dp, _dp, i32p, i32p,
                                                 i32p, _dp]),
            "osb_batched_bfgs_rosenbrock_generated": (ci, [_vp, i64, i64, i64, dbl, i64, i64, dbl, dbl, _dp, _dp,
                                                           i32p, i32p, i32p, _dp]),
            "osb_bench_qn_kernel": (ci, [_vp, ci, i64, ci, ci, _dp]),
            "osb_bench_syrk": (ci, [_vp, _vp, ci, _dp]),
            "osb_bench_grid_sync": (ci, [_vp, ci, _dp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


EXPORTED_SYMBOLS = None  # filled lazily by exported_symbols()


def exported_symbols():
    """Names every test may expect in the shared library (= the declarations of include/optsolv_b200.h)."""
    import re
    hdr = open(os.path.join(_HERE, "..", "include", "optsolv_b200.h")).read()
    return sorted(set(re.findall(r"\b(osb_[a-z0-9_]+)\s*\(", hdr)) - {"osb_host_eval_fn", "osb_device_eval_fn",
                                                                       "osb_callback_fn", "osb_phi_fn"})


def _arr(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


# ---- errors: src/ls_solver.rs:10-20 ---------------------------------------------------------
class SolverError(Exception):
    pass


class MaxIterReached(SolverError):
    def __init__(self):
        super().__init__("Max iter reached")


class OutOfDomain(SolverError):
    def __init__(self):
        super().__init__("Out of domain")


class ErrorInputParams(SolverError):
    def __init__(self, msg="Error in input parameters"):
        super().__init__(msg)


class AbnormalTermination(SolverError):
    def __init__(self, msg="Abnormal termination"):
        super().__init__(msg)


class ReferencePanic(RuntimeError):
    """Situations in which the reference panics (missing Hessian, non-SPD Cholesky)."""


class DeviceError(RuntimeError):
    """CUDA / NCCL / allocation / unsupported-configuration errors (new to a device backend)."""


REASONS = {0: None, 1: "grad_tol", 2: "s_norm", 3: "y_norm", 4: "proj_grad_tol", 5: "newton_decrement"}


def _check(rc):
    if rc == 0:
        return
    msg = lib().osb_last_error_string().decode()
    if rc == 1:
        raise MaxIterReached()
    if rc == 2:
        raise OutOfDomain()
    if rc == 3:
        raise ErrorInputParams(msg)
    if rc == 4:
        raise AbnormalTermination(msg)
    if rc in (101, 102):
        raise ReferencePanic(msg)
    raise DeviceError("status %d: %s" % (rc, msg))


# ---- Tracer / LogFormat: src/tracer.rs:5-63 (host-side logging only) --------------------------
class LogFormat:
    Pretty, Json, Normal = "pretty", "json", "normal"


class Tracer:
    """`Tracer::default().with_stdout_layer(Some(LogFormat::Normal)).build()` (src/tracer.rs:25-63).  The events themselves
    come from the compiled library (osb_set_log_callback: the reference's targets, levels and messages); this builder only
    installs the sink: one `logging` logger per target under "optimization_solvers", verbosity from RUST_LOG like
    EnvFilter::from_default_env()."""
    _LEVELS = {1: logging.ERROR, 2: logging.WARNING, 3: logging.INFO, 4: logging.DEBUG, 5: 5}
    _sink = None  # keeps the ctypes callback alive

    def __init__(self):
        self._fmt = None

    default = classmethod(lambda cls: cls())

    def with_stdout_layer(self, fmt=None):
        self._fmt = fmt or LogFormat.Normal
        return self

    def with_normal_stdout_layer(self):
        return self.with_stdout_layer(LogFormat.Normal)

    def build(self):
        level = {"trace": 5, "debug": logging.DEBUG, "info": logging.INFO, "warn": logging.WARNING,
                 "error": logging.ERROR}.get(os.environ.get("RUST_LOG", "").lower(), logging.ERROR)
        if self._fmt is not None and not log.handlers:
            h = logging.StreamHandler()
            if self._fmt == LogFormat.Json:
                h.setFormatter(logging.Formatter('{"level":"%(levelname)s","target":"%(target)s","fields":{"message":"%(message)s"}}'))
            elif self._fmt == LogFormat.Pretty:
                h.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(target)s\n    %(message)s"))
            else:
                h.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(target)s: %(message)s"))
            log.addHandler(h)
        log.setLevel(level)
        install_log_sink()
        return []  # the crate returns WorkerGuards


def install_log_sink():
    """Routes the library's Tracer events into the "optimization_solvers" logger (record attribute `target` = the crate's
    tracing target: "solver", "bfgs", "newton", ...).  Installed on first use of a solver as well: the events are part of
    the path, not of the builder."""
    if Tracer._sink is None:
        def sink(_user, level, target, message):
            log.log(Tracer._LEVELS.get(level, logging.INFO), message.decode(), extra={"target": target.decode()})
        Tracer._sink = LOG_FN(sink)
        lib().osb_set_log_callback(Tracer._sink, None)


# ---- FuncEval: src/func_eval.rs:5-41 --------------------------------------------------------
class FuncEvalMultivariate:
    def __init__(self, f, g, hessian=None):
        self._f, self._g, self._h = float(f), _arr(g), hessian

    @staticmethod
    def new(f, g):
        return FuncEvalMultivariate(f, g)

    def with_hessian(self, h):
        self._h = np.asarray(h, dtype=np.float64)
        return self

    def take_hessian(self):
        h, self._h = self._h, None
        if h is None:
            raise ReferencePanic("called `Option::unwrap()` on a `None` value")
        return h

    def f(self):
        return self._f

    def g(self):
        return self._g

    def hessian(self):
        return self._h


class FuncEvalUnivariate:
    def __init__(self, f, g):
        self._f, self._g = float(f), float(g)

    def f(self):
        return self._f

    def g(self):
        return self._g


def box_projection(x, lower_bound, upper_bound):
    """number.rs:13-21 (host helper for callers; the solvers project on device)."""
    return np.fmin(np.fmax(_arr(x), _arr(lower_bound)), _arr(upper_bound))


def infinity_norm(v):
    """number.rs:27-31"""
    acc = 0.0
    for a in np.abs(_arr(v)):
        acc = max(acc, a) if not np.isnan(a) else acc
    return acc


# ---- context --------------------------------------------------------------------------------
class Context:
    """One per GPU (one process per GPU)."""

    def __init__(self, device=0, rank=0, world=1, nccl_unique_id=None):
        h = _vp()
        if world > 1:
            buf = C.create_string_buffer(bytes(nccl_unique_id), 128)
            _check(lib().osb_ctx_create_dist(device, rank, world, buf, C.byref(h)))
        else:
            _check(lib().osb_ctx_create(device, C.byref(h)))
        self.handle = h
        self.device = device

    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        _check(lib().osb_nccl_unique_id(buf))
        return bytes(buf.raw)

    def ipc_handle(self):
        buf = C.create_string_buffer(64)
        _check(lib().osb_ctx_ipc_handle(self.handle, buf))
        return bytes(buf.raw)

    def ipc_connect(self, handles):
        """handles: list of the 64-byte IPC handles of all ranks, in rank order."""
        blob = b"".join(handles)
        buf = C.create_string_buffer(blob, len(blob))
        _check(lib().osb_ctx_ipc_connect(self.handle, buf))

    def connect_peers(self, strict=True):
        """All-gather the IPC handles over torch.distributed and connect (fused NVLink exchange).  Returns True when
        EVERY rank connected.  With strict=False a failure on any rank (no peer access between two GPUs, IPC disabled
        in a container) is not an error: all ranks return False together and keep using the NCCL path."""
        import torch.distributed as dist
        ok, err = 1, None
        try:
            mine = self.ipc_handle()
        except Exception as e:  # noqa: BLE001 - reported below, after the collective
            mine, ok, err = b"\0" * 64, 0, e
        allh = [None] * dist.get_world_size()
        dist.all_gather_object(allh, mine)
        oks = [None] * dist.get_world_size()
        dist.all_gather_object(oks, ok)
        if all(oks):
            try:
                self.ipc_connect(allh)
            except Exception as e:  # noqa: BLE001
                ok, err = 0, e
        else:
            ok = 0
        dist.all_gather_object(oks, ok)
        self.p2p = bool(all(oks))
        if not self.p2p:
            # not unanimous: the ranks that did connect must drop their mappings too, or they would take the fused
            # peer-memory path and wait for a rank that sits in an NCCL call
            lib().osb_ctx_ipc_close(self.handle)
        if not self.p2p and strict:
            raise err if err is not None else DeviceError("peer-memory exchange unavailable on another rank")
        return self.p2p

    def set_vector_sharding(self, on=True):
        """Index-range sharding of GD / PGD / SPG over the ranks of this context (see include/optsolv_b200.h)."""
        _check(lib().osb_ctx_set_vector_sharding(self.handle, 1 if on else 0))
        return self

    def trim_memory(self):
        """Return the pooled n x n buffers of closed solvers to the driver."""
        _check(lib().osb_ctx_trim_memory(self.handle))

    def rank(self):
        return lib().osb_ctx_rank(self.handle)

    def world(self):
        return lib().osb_ctx_world(self.handle)

    def synchronize(self):
        _check(lib().osb_ctx_synchronize(self.handle))

    def stream(self):
        return lib().osb_ctx_stream(self.handle)

    def counters(self):
        out = (C.c_int64 * 8)()
        lib().osb_ctx_counters(self.handle, out)
        v = list(out)
        return dict(launches=v[0], objective_evals=v[1], ls_trials=v[2], host_syncs=v[3], collectives=v[4],
                    sharded_packed_passes=v[5])

    def close(self):
        if self.handle:
            lib().osb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("OSB_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    return _default_ctx


def set_default_context(ctx):
    global _default_ctx
    _default_ctx = ctx


# ---- objectives -----------------------------------------------------------------------------
class _Objective:
    handle = None
    with_hessian = False

    def calls(self):
        return lib().osb_objective_calls(self.handle)

    def dim(self):
        return lib().osb_objective_dim(self.handle)

    def __call__(self, x):
        x = _arr(x)
        n = x.size
        f = C.c_double()
        g = np.empty(n)
        h = np.empty((n, n)) if self.with_hessian else None
        _check(lib().osb_objective_eval(self.handle, _p(x), C.byref(f), _p(g), _p(h)))
        e = FuncEvalMultivariate(f.value, g)
        if h is not None:
            e.with_hessian(h)
        return e

    def close(self):
        if self.handle:
            lib().osb_objective_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostOracle(_Objective):
    """A host closure `FnMut(&DVector) -> FuncEvalMultivariate` (compatibility path, H2D/D2H per call)."""

    def __init__(self, fn, n, with_hessian=False, ctx=None):
        self.ctx = ctx or default_context()
        self.fn = fn
        self.with_hessian = with_hessian

        def tramp(_user, xp, n_, fp, gp, hp):
            x = np.ctypeslib.as_array(xp, shape=(n_,)).copy()
            r = fn(x)
            if isinstance(r, tuple):
                r = FuncEvalMultivariate(*r)
            fp[0] = r.f()
            np.ctypeslib.as_array(gp, shape=(n_,))[:] = r.g()
            if hp and r.hessian() is not None:
                np.ctypeslib.as_array(hp, shape=(n_, n_))[:] = np.asarray(r.hessian())
                return 1
            return 0

        self._cb = HOST_EVAL(tramp)
        h = _vp()
        _check(lib().osb_objective_create_host(self.ctx.handle, n, self._cb, None, 1 if with_hessian else 0, C.byref(h)))
        self.handle = h


class UserDeviceObjective(_Objective):
    """A user-supplied device functor: `enqueue(d_x, n, d_f, d_g, d_hess, stream) -> 0` launches user kernels."""

    def __init__(self, enqueue, n, with_hessian=False, ctx=None):
        self.ctx = ctx or default_context()
        self.with_hessian = with_hessian

        def tramp(_user, d_x, n_, d_f, d_g, d_h, stream):
            return int(enqueue(d_x, n_, d_f, d_g, d_h, stream) or 0)

        self._cb = DEVICE_EVAL(tramp)
        h = _vp()
        _check(lib().osb_objective_create_user(self.ctx.handle, n, self._cb, None, 1 if with_hessian else 0, C.byref(h)))
        self.handle = h


class DenseQuadratic(_Objective):
    """f = x.(A x) [- 2 b.x], g = 2 A x [- 2 b]  (examples/quadratic.rs:10-14 pattern); Hessian 2A."""
    with_hessian = True

    def __init__(self, A, b=None, ctx=None):
        self.ctx = ctx or default_context()
        A = _arr(A)
        n = A.shape[0]
        bb = _arr(b) if b is not None else None
        h = _vp()
        _check(lib().osb_objective_create_dense_quadratic(self.ctx.handle, n, _p(A), _p(bb), C.byref(h)))
        self.handle = h

    @classmethod
    def generated(cls, n, shifted=True, ctx=None):
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self.x0 = np.empty(n)
        h = _vp()
        _check(lib().osb_objective_create_dense_quadratic_generated(self.ctx.handle, n, 1 if shifted else 0,
                                                                    _p(self.x0), C.byref(h)))
        self.handle = h
        return self


class ExtendedRosenbrock(_Objective):
    def __init__(self, n, ctx=None):
        self.ctx = ctx or default_context()
        h = _vp()
        _check(lib().osb_objective_create_rosenbrock(self.ctx.handle, n, C.byref(h)))
        self.handle = h


class SeparableQuadratic(_Objective):
    @classmethod
    def generated(cls, n, ctx=None):
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        h = _vp()
        _check(lib().osb_objective_create_separable_quadratic_generated(self.ctx.handle, n, C.byref(h)))
        self.handle = h
        return self

    @classmethod
    def generated_shard(cls, n_local, index0, ctx):
        """Coordinates [index0, index0 + n_local) of the generated problem (index-range sharded SPG / PGD / GD)."""
        self = cls.__new__(cls)
        self.ctx = ctx
        h = _vp()
        _check(lib().osb_objective_create_separable_quadratic_generated_shard(self.ctx.handle, n_local, index0, C.byref(h)))
        self.handle = h
        return self


class LogisticRegression(_Objective):
    with_hessian = True

    @classmethod
    def generated(cls, m, n, lam=1.0, ctx=None):
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        h = _vp()
        _check(lib().osb_objective_create_logistic_generated(self.ctx.handle, m, n, lam, C.byref(h)))
        self.handle = h
        return self


def _as_objective(o, n, with_hessian, ctx):
    if isinstance(o, _Objective):
        return o
    return HostOracle(o, n, with_hessian=with_hessian, ctx=ctx)


# ---- line searches: src/line_search/ ----------------------------------------------------------
class _LS:
    handle = None
    ctx = None

    def _make(self, ctx):
        raise NotImplementedError

    def _h(self, ctx=None):
        if self.handle is None:
            self.ctx = ctx or default_context()
            h = _vp()
            _check(self._make(self.ctx, C.byref(h)))
            self.handle = h
        return self.handle

    def compute_step_len(self, x_k, direction_k, oracle, max_iter, ctx=None):
        """LineSearch::compute_step_len (line_search/mod.rs:14-23); eval_x_k is recomputed from the oracle."""
        ctx = ctx or default_context()
        x, d = _arr(x_k), _arr(direction_k)
        o = _as_objective(oracle, x.size, False, ctx)
        t = C.c_double()
        _check(lib().osb_linesearch_compute_step_len(ctx.handle, self._h(ctx), o.handle, _p(x), _p(d), max_iter,
                                                     C.byref(t)))
        return t.value

    def step_len_scalar(self, phi, f0, gd0, max_iter, tmax_candidate=float("inf")):
        """Host-only run of the line-search automaton on a 1-D model: phi(t, projected) -> (f, g.d, ||dx||^2).
        Needs no GPU (the bounded kinds' vectors are not touched); returns (t, last_eval_is_result)."""
        def tramp(_u, t, proj, fp, gp, dp_):
            r = phi(t, bool(proj))
            fp[0], gp[0], dp_[0] = r[0], r[1], (r[2] if len(r) > 2 else 0.0)
        cb = PHI_EVAL(tramp)
        h = self.handle
        if h is None:
            hh = _vp()
            _check(self._make_host(C.byref(hh)))
            h = hh
        t, cur = C.c_double(), C.c_int()
        _check(lib().osb_linesearch_step_len_scalar(h, cb, None, f0, gd0, tmax_candidate, max_iter, C.byref(t),
                                                    C.byref(cur)))
        if self.handle is None:
            self._host_handle = h
        return t.value, bool(cur.value)

    def _make_host(self, out):
        return self._make(None, out)

    def close(self):
        if self.handle:
            lib().osb_linesearch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BackTracking(_LS):
    def __init__(self, c1, beta):
        self.c1, self.beta = c1, beta

    new = classmethod(lambda cls, *a: cls(*a))

    def _make(self, ctx, out):
        return lib().osb_linesearch_create_backtracking(self.c1, self.beta, out)


class BackTrackingB(_LS):
    def __init__(self, c1, beta, lower_bound, upper_bound):
        self.c1, self.beta, self.lb, self.ub = c1, beta, _arr(lower_bound), _arr(upper_bound)

    new = classmethod(lambda cls, *a: cls(*a))

    def lower_bound(self):
        return self.lb

    def upper_bound(self):
        return self.ub

    def _make(self, ctx, out):
        return lib().osb_linesearch_create_backtracking_b(ctx.handle if ctx else None, self.c1, self.beta, self.lb.size, _p(self.lb),
                                                          _p(self.ub), out)


class MoreThuente(_LS):
    """morethuente.rs:16-62 (asserts included)."""

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

    def _make(self, ctx, out):
        return lib().osb_linesearch_create_morethuente(self.c1, self.c2, self.t_min, self.t_max, self.delta_min,
                                                       self.delta, self.delta_max, out)


class MoreThuenteB(MoreThuente):
    """morethuente_b.rs:17-40; t_max shrinks permanently across outer iterations (:201)."""

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

    def _make(self, ctx, out):
        return lib().osb_linesearch_create_morethuente_b(ctx.handle if ctx else None, self.c1, self.c2, self.t_min, self.t_max,
                                                         self.delta_min, self.delta, self.delta_max, self.lb.size,
                                                         _p(self.lb), _p(self.ub), out)

    def current_t_max(self):
        return lib().osb_linesearch_t_max(self._h())


class GLLQuadratic(_LS):
    def __init__(self, c1, m):
        self.c1, self.m, self.sigma1, self.sigma2 = c1, m, 0.1, 0.9

    new = classmethod(lambda cls, *a: cls(*a))

    def with_sigmas(self, s1, s2):
        self.sigma1, self.sigma2 = s1, s2
        return self

    def _make(self, ctx, out):
        return lib().osb_linesearch_create_gll_quadratic(self.c1, self.m, self.sigma1, self.sigma2, out)


class NoSearch(_LS):
    def _make(self, ctx, out):
        return lib().osb_linesearch_create_nosearch(out)


# ---- solvers ----------------------------------------------------------------------------------
_KIND = dict(GD=0, PGD=1, SPG=2, BFGS=3, DFP=4, BROYDEN=5, BFGSB=6, DFPB=7, BROYDENB=8, SR1B=9, NEWTON=10,
             PROJ_NEWTON=11, SPN=12, PNORM=13)


class _Solver:
    KIND = None
    NEEDS_HESSIAN = False
    TARGET = "solver"

    def __init__(self, tol, x0, lower_bound=None, upper_bound=None, oracle=None, ctx=None):
        self.ctx = ctx or default_context()
        x0 = _arr(x0)
        self.n = x0.size
        self._tol = tol
        self._lb = _arr(lower_bound) if lower_bound is not None else None
        self._ub = _arr(upper_bound) if upper_bound is not None else None
        self._keep = _as_objective(oracle, self.n, self.NEEDS_HESSIAN, self.ctx) if oracle is not None else None
        h = _vp()
        _check(lib().osb_solver_create(self.ctx.handle, _KIND[self.KIND], self.n, tol, _p(x0), _p(self._lb),
                                       _p(self._ub), self._keep.handle if self._keep is not None else None,
                                       C.byref(h)))
        self.handle = h

    @classmethod
    def new(cls, *a, **k):
        return cls(*a, **k)

    def set_option(self, name, value):
        _check(lib().osb_solver_set_option(self.handle, name.encode(), int(value)))
        return self

    def minimize(self, line_search, oracle, max_iter_solver, max_iter_line_search, callback=None):
        """LineSearchSolver::minimize (ls_solver.rs:66-111): returns None for Ok(()), raises SolverError otherwise."""
        o = _as_objective(oracle, self.n, self.NEEDS_HESSIAN, self.ctx)
        install_log_sink()
        if callback is not None:
            cb = CALLBACK(lambda _u, _s: callback(self))
        else:
            cb = C.cast(None, CALLBACK)
        rc = lib().osb_minimize(self.handle, line_search._h(self.ctx), o.handle, max_iter_solver,
                                max_iter_line_search, cb, None)
        self.status = rc
        _check(rc)
        return None

    def x(self):
        out = np.empty(self.n)
        _check(lib().osb_solver_x(self.handle, _p(out)))
        return out

    xk = x

    def set_x(self, x):
        x = _arr(x)
        _check(lib().osb_solver_set_x(self.handle, _p(x)))

    def k(self):
        return lib().osb_solver_k(self.handle)

    def tol(self):
        return self._tol

    grad_tol = tol

    def f(self):
        v = C.c_double()
        _check(lib().osb_solver_f(self.handle, C.byref(v)))
        return v.value

    def grad(self):
        out = np.empty(self.n)
        _check(lib().osb_solver_grad(self.handle, _p(out)))
        return out

    def termination_reason(self):
        return REASONS[lib().osb_solver_termination_reason(self.handle)]

    @staticmethod
    def _opt(v):
        return None if np.isnan(v) else v

    def s_norm(self):
        return self._opt(lib().osb_solver_s_norm(self.handle))

    def y_norm(self):
        return self._opt(lib().osb_solver_y_norm(self.handle))

    def clear_norms(self):
        _check(lib().osb_solver_clear_norms(self.handle))

    def lambda_(self):
        return lib().osb_solver_lambda(self.handle)

    def with_lambdas(self, lmin, lmax):
        _check(lib().osb_solver_set_lambdas(self.handle, lmin, lmax))
        return self

    def decrement_squared(self):
        return self._opt(lib().osb_solver_decrement_squared(self.handle))

    def approx_inv_hessian(self):
        out = np.zeros((self.n, self.n))
        _check(lib().osb_solver_inv_hessian(self.handle, _p(out)))
        return out

    def set_approx_inv_hessian(self, H):
        H = _arr(H)
        _check(lib().osb_solver_set_inv_hessian(self.handle, _p(H)))

    def active_set(self):
        out = np.zeros(self.n, dtype=np.uint8)
        _check(lib().osb_solver_active_set(self.handle, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def lower_bound(self):
        return self._lb

    def upper_bound(self):
        return self._ub

    def projected_gradient(self, ev):
        """HasProjectedGradient::projected_gradient (ls_solver.rs:121-133), host helper on a FuncEval."""
        x, pg = self.x(), ev.g().copy()
        m = ((x == self._lb) & (pg > 0.0)) | ((x == self._ub) & (pg < 0.0))
        pg[m] = 0.0
        return pg

    def has_converged(self, ev):
        raise NotImplementedError("has_converged runs on the device inside minimize()")

    def record_trace(self, on=True):
        return self.set_option("record_trace", 1 if on else 0)

    def trace(self):
        m = lib().osb_solver_trace_len(self.handle)
        f, t, sn, yn = (np.empty(m) for _ in range(4))
        lib().osb_solver_trace(self.handle, _p(f), _p(t), _p(sn), _p(yn))
        return dict(f=f, t=t, s_norm=sn, y_norm=yn)

    def last_timing(self):
        ms, it = C.c_double(), C.c_int64()
        lib().osb_solver_last_timing(self.handle, C.byref(ms), C.byref(it))
        return ms.value, it.value

    def path_info(self):
        """Which path the last minimize() took (the options default to auto): see osb_solver_path_info."""
        out = (C.c_int64 * 8)()
        lib().osb_solver_path_info(self.handle, out)
        eng, sched, stor, shard, p2p, world, variant, fused = (int(out[i]) for i in range(8))
        return dict(engine={1: "host-driven control", 2: "device-resident control"}.get(eng, "none"), schedule=sched, storage=stor,
                    schedule_name=("lazy: 1 RMW pass of the stored matrix per iteration" if sched == 1 else "eager: gemv + fused update (3 n^2 8 B)"),
                    storage_name=("packed lower triangle, 8-row tiles (n^2 8 B per pass)" if stor == 1 else "full n x n row-major"),
                    sharded_packed=bool(shard), p2p=bool(p2p), world=world, variant=variant, fused=bool(fused & 1), fused_stream=bool(fused & 2),
                    kernel=("qn_iter_kernel: whole iterations (line search, H pass, fold, exchange) in one cooperative launch"
                            if fused & 1 else "one launch per phase (head, pass, fold)"),
                    parallelism=("1 GPU" if world == 1 else
                                 "packed triangle sharded by tile pairs over %d GPUs; per-rank {h, w} contributions stored into every "
                                 "peer's slot (NVLink stores + flags), summed in rank order" % world if shard else
                                 "row-block sharded H over %d GPUs; %s" % (world, "peer-memory all-gather fused into the pass kernel"
                                                                           if p2p else "NCCL all-gather of the h / w slices")))

    def iter_profile(self):
        """Option "profile_iter" = 1: mean ms per iteration spent in the head (epilogue, line search, step), the H pass and
        the fold + exchange of the fused iteration kernel (globaltimer stamps of CTA 0), and the iterations covered."""
        out = (C.c_double * 16)()
        _check(lib().osb_solver_iter_profile(self.handle, out))
        names = {4: "epilogue_loads", 5: "epilogue_grid_sum", 6: "u_and_direction", 7: "trial_steps", 8: "trials_grid_sum",
                 9: "automaton", 10: "next_iterate", 11: "step_grid_sum", 12: "barrier_wait_for_last_cta_sum", 13: "barrier_release_after_last_sum"}
        return dict(head_ms=out[0], pass_ms=out[1], fold_ms=out[2], iterations=int(out[3]),
                    head_parts_us={v: round(out[k] * 1e3, 3) for k, v in names.items()})

    def kernel_timing(self):
        out = (C.c_double * 3)()
        lib().osb_solver_kernel_timing(self.handle, out)
        return dict(gemv_ms=out[0], update_ms=out[1], iterations=int(out[2]))

    def close(self):
        if getattr(self, "handle", None):
            lib().osb_solver_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _mk(name, kind, bounded=False, needs_oracle=False, needs_h=False):
    if needs_oracle:
        def __init__(self, tol, x0, oracle, lower_bound, upper_bound, ctx=None):
            _Solver.__init__(self, tol, x0, lower_bound, upper_bound, oracle, ctx=ctx)
    elif bounded:
        def __init__(self, tol, x0, lower_bound, upper_bound, ctx=None):
            _Solver.__init__(self, tol, x0, lower_bound, upper_bound, ctx=ctx)
    else:
        def __init__(self, tol, x0, ctx=None):
            _Solver.__init__(self, tol, x0, ctx=ctx)
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
    """PnormDescent::new(grad_tol, x0, inverse_p) (src/steepest_descent/pnorm_descent.rs:23-30).  `inverse_p` is an
    n x n array ([i, j] = row i, column j); the direction is -(inverse_p g) (pnorm_descent.rs:35), convergence is the
    inf-norm test of GradientDescent (pnorm_descent.rs:52-58)."""
    KIND = "PNORM"

    def __init__(self, tol, x0, inverse_p, ctx=None):
        _Solver.__init__(self, tol, x0, ctx=ctx)
        P = np.ascontiguousarray(np.asarray(inverse_p, dtype=np.float64))
        if P.shape != (self.n, self.n):
            raise ErrorInputParams("inverse_p must be n x n")
        self.set_approx_inv_hessian(P)

    def inverse_p(self):
        return self.approx_inv_hessian()


def sym_layout(n, world, rank, tile):
    """(owner, offset, row_stride, rank_total) of an 8-row tile in the packed symmetric layout (host-only query)."""
    owner, off, ldp, tot = C.c_int(), C.c_int64(), C.c_int64(), C.c_int64()
    _check(lib().osb_sym_layout(n, world, rank, tile, C.byref(owner), C.byref(off), C.byref(ldp), C.byref(tot)))
    return owner.value, off.value, ldp.value, tot.value


# ---- batched mode -----------------------------------------------------------------------------
def batched_bfgs_rosenbrock(n, n_problems, x0=None, problem0=0, tol=1e-8, max_iter_solver=2000,
                            max_iter_line_search=20, c1=1e-4, beta=0.5, ctx=None):
    """Many independent BFGS + BackTracking solves of extended Rosenbrock (one warp per problem)."""
    ctx = ctx or default_context()
    x = np.empty((n_problems, n))
    f = np.empty(n_problems)
    k = np.empty(n_problems, dtype=np.int32)
    st = np.empty(n_problems, dtype=np.int32)
    rs = np.empty(n_problems, dtype=np.int32)
    ms = C.c_double()
    i32p = C.POINTER(C.c_int32)
    if x0 is not None:
        x0 = _arr(x0).reshape(n_problems, n)
        _check(lib().osb_batched_bfgs_rosenbrock(ctx.handle, n, n_problems, _p(x0), tol, max_iter_solver,
                                                 max_iter_line_search, c1, beta, _p(x), _p(f),
                                                 k.ctypes.data_as(i32p), st.ctypes.data_as(i32p),
                                                 rs.ctypes.data_as(i32p), C.byref(ms)))
    else:
        _check(lib().osb_batched_bfgs_rosenbrock_generated(ctx.handle, n, n_problems, problem0, tol, max_iter_solver,
                                                           max_iter_line_search, c1, beta, _p(x), _p(f),
                                                           k.ctypes.data_as(i32p), st.ctypes.data_as(i32p),
                                                           rs.ctypes.data_as(i32p), C.byref(ms)))
    return dict(x=x, f=f, k=k, status=st, reason=rs, ms=ms.value)


def bench_qn_kernel(which, n, reps=20, variant=0, ctx=None):
    """Mean ms/launch of one hot kernel on device-resident data (see include/optsolv_b200.h)."""
    ctx = ctx or default_context()
    ms = C.c_double()
    _check(lib().osb_bench_qn_kernel(ctx.handle, which, n, reps, variant, C.byref(ms)))
    return ms.value


def bench_grid_sync(reps=200, ctx=None):
    """Mean microseconds of one grid barrier of the fused iteration kernel's shape."""
    ctx = ctx or default_context()
    us = C.c_double()
    _check(lib().osb_bench_grid_sync(ctx.handle, reps, C.byref(us)))
    return us.value


def bench_syrk(logistic, reps=3, ctx=None):
    """Mean ms/launch of the DMMA Hessian assembly of a LogisticRegression objective."""
    ctx = ctx or default_context()
    ms = C.c_double()
    _check(lib().osb_bench_syrk(ctx.handle, logistic.handle, reps, C.byref(ms)))
    return ms.value
