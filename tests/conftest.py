import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ocs():
    """The product package (its directory name carries a hyphen, hence importlib)."""
    build = importlib.import_module("optimal-control-solvers_b200.build")
    build.build()
    return importlib.import_module("optimal-control-solvers_b200")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): NumPy restatement + C restatement."""
    from oracle import oracle_c, rk4_oracle
    oracle_c.build()
    rk4_oracle.C = oracle_c
    return rk4_oracle


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
