"""CPU: the C-ABI shared library loads and exports every symbol include/optsolv_b200.h declares;
the host-only part of the boundary (line-search automaton on a 1-D model, error reporting) behaves
like the reference.  No compute call needs a GPU here."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from problems import quad2


def test_library_exports_every_declared_symbol(osb):
    assert os.path.exists(osb.LIB_PATH), "build the CUDA library first (python __graft_entry__.py)"
    L = C.CDLL(osb.LIB_PATH)
    names = osb.exported_symbols()
    assert len(names) >= 55
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in osb.lib().osb_version()


def test_no_cpu_fallback_without_gpu(osb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(osb.DeviceError) as e:
        osb.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_morethuente_builder_asserts(osb):
    # morethuente.rs:51-59
    with pytest.raises(AssertionError):
        osb.MoreThuente.default().with_c1(0.95)
    with pytest.raises(AssertionError):
        osb.MoreThuente.default().with_c2(1.5)
    ls = osb.MoreThuente.default().with_c1(1e-3).with_c2(0.5).with_t_min(0.0).with_t_max(10.0)
    assert (ls.c1, ls.c2, ls.t_max) == (1e-3, 0.5, 10.0)


def _phi_of(f, x, d):
    def phi(t, projected):
        val, g = f(x + t * d)
        return (val, float(g[0] * d[0] + g[1] * d[1]), 0.0)  # explicit: BLAS ddot may fuse
    return phi


@pytest.mark.parametrize("kind", ["bt", "mt", "gll", "nosearch"])
def test_host_automaton_matches_oracle_line_search(osb, orc, kind):
    """Same automaton code as the device engines, driven on the host by a 1-D model phi(t); the
    oracle runs the reference's line search on the vectors (backtracking.rs, morethuente.rs,
    gll_quadratic.rs).  Step lengths must be bit-identical (n = 2: same dot order)."""
    f = quad2(90.0)
    rng = np.random.default_rng(1)
    for trial in range(25):
        x = rng.uniform(-200, 200, 2)
        val, g = f(x)
        d = -g * (rng.uniform(0.01, 3.0) if trial % 2 else 1.0)
        mk = {"bt": lambda m: m.BackTracking(1e-4, 0.5), "mt": lambda m: m.MoreThuente.default(),
              "gll": lambda m: m.GLLQuadratic(1e-4, 10), "nosearch": lambda m: m.NoSearch()}[kind]
        t_ref = mk(orc).compute_step_len(x, d, f, 30)
        t_dev, _ = mk(osb).step_len_scalar(_phi_of(f, x, d), val, float(g[0] * d[0] + g[1] * d[1]), 30)
        assert t_dev == t_ref, (trial, t_dev, t_ref)


def test_host_automaton_quirks(osb):
    # backtracking.rs:37-41: a NaN/inf trial shrinks t without consuming an iteration
    calls = []

    def phi(t, projected):
        calls.append(t)
        if t > 0.2:
            return (float("nan"), 0.0, 0.0)
        return (10.0, -1.0, 0.0)  # never satisfies Armijo (f0 = 0)
    t, cur = osb.BackTracking(1e-4, 0.5).step_len_scalar(phi, 0.0, -1.0, 2)
    assert calls == [1.0, 0.5, 0.25, 0.125, 0.0625] and t == 0.03125 and cur is False
    # morethuente.rs:290: NaN from cubic_minimizer (negative discriminant) clamps to t_min = 0
    def phi2(t, projected):
        return (1.0 + t, 1.0, 0.0) if t > 0 else (0.0, -1.0, 0.0)
    t, cur = osb.MoreThuente.default().step_len_scalar(phi2, 0.0, -1.0, 10)
    assert t >= 0.0 and not math.isnan(t)
    # max_iter = 0: loops do not run, the initial step is returned unevaluated
    t, cur = osb.BackTracking(1e-4, 0.5).step_len_scalar(phi, 0.0, -1.0, 0)
    assert t == 1.0 and cur is False
    # MoreThuenteB: t_max shrinks permanently (morethuente_b.rs:201)
    ls = osb.MoreThuenteB(2)
    t, _ = ls.step_len_scalar(lambda t, p: (t * t - t, 2 * t - 1, 0.0), 0.0, -1.0, 10, tmax_candidate=0.25)
    assert t <= 0.25
    assert osb.lib().osb_linesearch_t_max(ls._host_handle) == 0.25


def _rosen2(x):
    a, b = x[0], x[1]
    t1 = b - a * a
    t2 = 1.0 - a
    return 100.0 * (t1 * t1) + t2 * t2, np.array([-400.0 * (a * t1) - 2.0 * t2, 200.0 * t1])


def _quartic(x):
    a, b = x[0], x[1]
    return (a * a) * (a * a) + 3.0 * (b * b) + a * b, np.array([4.0 * (a * a * a) + b, 6.0 * b + a])


def _logbarrier(x):
    # finite only for x0 < 1: trials that leave the domain return NaN (backtracking.rs:37-41, morethuente NaN clamps)
    a, b = x[0], x[1]
    if a >= 1.0:
        return float("nan"), np.array([float("nan"), float("nan")])
    return -math.log(1.0 - a) + 0.5 * (b * b) + 0.5 * (a * a), np.array([1.0 / (1.0 - a) + a, b])


@pytest.mark.parametrize("fname", ["rosen2", "quartic", "logbarrier"])
def test_host_automaton_matches_oracle_on_nonquadratic_models(osb, orc, fname):
    """The same comparison on non-quadratic 2-D models — curved valleys, quartic growth, a domain boundary that makes
    trials return NaN — over random points, descent and non-descent directions, large and tiny scalings, and
    non-default parameters: the automaton (the code both GPU engines run) and the oracle's restatement of
    backtracking.rs / morethuente.rs / gll_quadratic.rs must return bit-identical steps."""
    f = {"rosen2": _rosen2, "quartic": _quartic, "logbarrier": _logbarrier}[fname]
    rng = np.random.default_rng(11)
    makers = [
        lambda m: m.BackTracking(1e-4, 0.5),
        lambda m: m.BackTracking(0.3, 0.8),
        lambda m: m.MoreThuente.default(),
        lambda m: m.MoreThuente.default().with_c2(0.9).with_c1(1e-3).with_t_max(10.0),
        lambda m: m.MoreThuente.default().with_t_min(1e-6).with_deltas(0.5, 0.66, 4.0),
        lambda m: m.GLLQuadratic(1e-4, 5),
    ]
    checked = 0
    for trial in range(60):
        x = rng.uniform(-2.0, 0.9, 2) if fname == "logbarrier" else rng.uniform(-2.5, 2.5, 2)
        val, g = f(x)
        scale = [1.0, 1e-3, 37.0, 1e3][trial % 4]
        d = -g * scale
        if trial % 7 == 6:
            d = rng.uniform(-1, 1, 2) * scale  # arbitrary (possibly ascent) direction
        gd0 = float(g[0] * d[0] + g[1] * d[1])
        for mk in makers:
            t_ref = mk(orc).compute_step_len(x, d, f, 25)
            t_dev, _ = mk(osb).step_len_scalar(_phi_of(f, x, d), val, gd0, 25)
            assert (t_dev == t_ref) or (math.isnan(t_dev) and math.isnan(t_ref)), (fname, trial, t_dev, t_ref)
            checked += 1
    assert checked == 360


@pytest.mark.parametrize("fname", ["quad2", "rosen2", "quartic"])
def test_host_automaton_bounded_searches_match_oracle(osb, orc, fname):
    """BackTrackingB (objective at the PROJECTED trial, decrease measured with ||x_t - x||^2, backtracking_b.rs:24-34,
    52-90) and MoreThuenteB (feasible-step cap from the bounds, morethuente_b.rs:185-201) against the oracle's
    restatement, over random boxes, interior and boundary points: bit-identical steps."""
    f = {"quad2": quad2(9.0, True), "rosen2": _rosen2, "quartic": _quartic}[fname]
    rng = np.random.default_rng(5)
    n_bt = n_mt = 0
    for trial in range(50):
        lb = rng.uniform(-3.0, -0.5, 2)
        ub = rng.uniform(0.5, 3.0, 2)
        x = rng.uniform(lb, ub)
        if trial % 5 == 0:
            x[0] = ub[0]  # on a face
        val, g = f(x)
        d = -g * [1.0, 0.05, 20.0][trial % 3]
        gd0 = float(g[0] * d[0] + g[1] * d[1])

        def phi(t, projected):
            xt = x + t * d
            if projected:
                xt = np.minimum(np.maximum(xt, lb), ub)
            v, gg = f(xt)
            dx = xt - x
            return (v, float(gg[0] * d[0] + gg[1] * d[1]), float(dx[0] * dx[0] + dx[1] * dx[1]))

        t_ref = orc.BackTrackingB(1e-4, 0.5, lb, ub).compute_step_len(x, d, f, 30)
        t_dev, _ = osb.BackTrackingB(1e-4, 0.5, lb, ub).step_len_scalar(phi, val, gd0, 30)
        assert t_dev == t_ref, ("BackTrackingB", fname, trial, t_dev, t_ref)
        n_bt += 1
        # morethuente_b.rs:185-197: largest step that keeps x + t d inside the box
        cand = float("inf")
        for i in range(2):
            if d[i] > 0.0:
                cand = min(cand, (ub[i] - x[i]) / d[i])
            elif d[i] < 0.0:
                cand = min(cand, (lb[i] - x[i]) / d[i])
        mref = orc.MoreThuenteB(2).with_lower_bound(lb).with_upper_bound(ub)
        mdev = osb.MoreThuenteB(2).with_lower_bound(lb).with_upper_bound(ub)
        t_ref = mref.compute_step_len(x, d, f, 30)
        t_dev, _ = mdev.step_len_scalar(phi, val, gd0, 30, tmax_candidate=cand)
        assert (t_dev == t_ref) or (math.isnan(t_dev) and math.isnan(t_ref)), ("MoreThuenteB", fname, trial, t_dev, t_ref)
        n_mt += 1
    assert n_bt == n_mt == 50


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm) needs no GPU: one JSON line with the
    metric, unit, config and the cpu_baseline / e2e objects of the measurement contract."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "bfgs_iterations_per_second_n16384_f64"
    assert line["unit"] == "iterations/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"] == {"value": line["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["n"] == 16384 and "workload" in line["config"]


def test_rust_ffi_matches_the_header():
    """rust/src/gpu/ffi.rs (the `extern "C"` block the crate's `gpu` module binds) against include/optsolv_b200.h: the same
    symbol set, the same argument counts, pointer-ness of every argument, every enum constant with the same value.  No Rust
    toolchain exists in this image, so this is the check that keeps the two files from drifting apart."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "optsolv_b200.h")).read(), flags=re.S)
    rs = open(os.path.join(root, "rust", "src", "gpu", "ffi.rs")).read()
    c_protos = {}
    for m in re.finditer(r"\n\s*((?:const\s+)?[A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(osb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr):
        if "typedef" in m.group(1):
            continue
        args = re.sub(r"\s+", " ", m.group(3)).strip()
        alist = [] if args == "void" else [a.strip() for a in args.split(",")]
        c_protos[m.group(2)] = (["*" in a or "[" in a or a.split()[0].endswith("_fn") for a in alist], m.group(1).strip() != "void")
    r_protos = {}
    for m in re.finditer(r"pub fn (osb_[a-z0-9_]+)\((.*?)\)( -> [^;]+)?;", rs):
        args = m.group(2).strip()
        alist = [] if not args else [a.strip() for a in re.split(r",\s*(?=[a-zA-Z_#][a-zA-Z0-9_#]*\s*:)", args)]
        r_protos[m.group(1)] = (["*" in a.split(":", 1)[1] or "_fn" in a.split(":", 1)[1] for a in alist], m.group(3) is not None)
    assert set(c_protos) == set(r_protos), (sorted(set(c_protos) ^ set(r_protos)))
    assert len(c_protos) >= 70
    for name, (c_args, c_ret) in c_protos.items():
        r_args, r_ret = r_protos[name]
        assert len(c_args) == len(r_args), (name, len(c_args), len(r_args))
        assert c_args == r_args, (name, c_args, r_args)
        assert c_ret == r_ret, name
    c_enums = dict((k, int(v)) for k, v in re.findall(r"\b(OSB_[A-Z0-9_]+)\s*=\s*(\d+)", hdr))
    r_enums = dict((k, int(v)) for k, v in re.findall(r"pub const (OSB_[A-Z0-9_]+): c_int = (\d+);", rs))
    assert c_enums == r_enums and len(c_enums) >= 30
    # and the library exports every one of them
    import importlib
    osb = importlib.import_module("optimization-solvers_b200")
    assert set(osb.exported_symbols()) == set(c_protos)


def test_rust_gpu_module_covers_the_solver_and_line_search_matrix():
    """The `gpu` module mirrors the reference's public surface for the path: 13 solver structs + PnormDescent, each with
    `impl LineSearchSolver` whose `minimize` is overridden, the 6 line searches described through `LineSearch::gpu_spec`,
    and every extern symbol it calls is declared in ffi.rs."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mod = open(os.path.join(root, "rust", "src", "gpu", "mod.rs")).read()
    ffi = open(os.path.join(root, "rust", "src", "gpu", "ffi.rs")).read()
    spec = open(os.path.join(root, "rust", "src", "line_search", "gpu_spec.rs")).read()
    solvers = ["BFGS", "DFP", "Broyden", "BFGSB", "DFPB", "BroydenB", "SR1B", "GradientDescent", "ProjectedGradientDescent",
               "SpectralProjectedGradient", "Newton", "ProjectedNewton", "SpectralProjectedNewton", "PnormDescent"]
    for s in solvers:
        assert re.search(r"(quasi_newton(_bounded)?!\(%s,|pub struct %s \{)" % (s, s), mod), s
        assert re.search(r"(gpu_solver!\(%s,|quasi_newton(_bounded)?!\(%s,)" % (s, s), mod), s
    assert "impl LineSearchSolver for $name" in mod and "fn minimize<LS: LineSearch>" in mod and "GpuSolverCore::minimize(self" in mod
    for ls in ("BackTracking", "BackTrackingB", "MoreThuente", "MoreThuenteB", "GLLQuadratic", "NoSearch"):
        assert "LineSearchSpec::%s" % ls in mod and "impl %s {" % ls in spec, ls
    declared = set(re.findall(r"pub fn (osb_[a-z0-9_]+)\(", ffi))
    used = set(re.findall(r"ffi::(osb_[a-z0-9_]+)\(", mod))
    assert used <= declared, sorted(used - declared)
    assert len(used) >= 35


def test_every_solver_option_is_documented_in_the_header():
    """osb_solver_set_option: every name the library accepts (csrc/api.cu) is described next to the declaration in
    include/optsolv_b200.h, and the header describes no option the library would reject."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    api = open(os.path.join(root, "optimization-solvers_b200", "csrc", "api.cu")).read()
    hdr = open(os.path.join(root, "include", "optsolv_b200.h")).read()
    accepted = set(re.findall(r'nm == "([a-z0-9_]+)"', api))
    assert len(accepted) >= 10
    i = hdr.index("int osb_solver_set_option(")
    doc = hdr[hdr.rindex("/*", 0, i):i]
    documented = set(re.findall(r'^ \*   "([a-z0-9_]+)"', doc, flags=re.M))
    assert accepted == documented, (sorted(accepted - documented), sorted(documented - accepted))
