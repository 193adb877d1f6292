import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def osb():
    """The product package (directory name is not an identifier, hence importlib)."""
    return importlib.import_module("optimization-solvers_b200")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle — the checker, never the thing shipped."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def gpu_ctx(osb):
    return osb.default_context()
