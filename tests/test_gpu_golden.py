"""GPU (-m gpu): the CUDA path against the COMMITTED vectors of tests/golden/oracle_trajectories.json — same status,
iteration count and termination reason, iterates within 1e-9 relative, bit-exact active sets.  Unlike
tests/test_gpu_parity.py this does not need the oracle to be built on the GPU box."""
import json
import os

import numpy as np
import pytest

from golden_cases import CASES

pytestmark = pytest.mark.gpu

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_trajectories.json")) as _f:
    GOLD = json.load(_f)


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_matches_committed_vector(osb, name):
    r, g = CASES[name](osb), GOLD[name]
    assert (r["status"], int(r["k"]), r["reason"]) == (g["status"], g["k"], g["reason"])
    x, gx = np.asarray(r["x"]), np.asarray(g["x"])
    assert np.all(np.abs(x - gx) <= 1e-12 + 1e-9 * max(1.0, float(np.max(np.abs(gx)))))
    if g["active_set"] is not None:
        assert np.array_equal(np.asarray(r["active_set"]), np.asarray(g["active_set"]))
