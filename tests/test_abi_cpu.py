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
