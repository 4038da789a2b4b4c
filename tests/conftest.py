import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes binding over libllz.so); built on demand."""
    p = entry.load_package()
    if not os.path.exists(p.LIB_PATH):
        p.build()
    return p


@pytest.fixture(scope="session")
def wl(pkg):
    import importlib

    return importlib.import_module("lambda_lanczos_b200.workloads")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def port(oracle_mod):
    """The plain-C restatement of the reference algorithm."""
    return oracle_mod.Restatement()


@pytest.fixture(scope="session")
def checker(oracle_mod):
    """Strongest CPU checker available: the compiled reference where oracle/_ref travelled, else the restatement."""
    return oracle_mod.best()


@pytest.fixture(scope="session")
def ctx(pkg):
    return pkg.Context(0)
