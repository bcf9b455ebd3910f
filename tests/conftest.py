"""Test configuration.

`-m "not gpu"`: oracle vs first-principles / documented properties / golden vectors, host logic, and the
C-ABI export checks (no compute on a GPU).  `-m gpu`: the parity tests proper -- the CUDA path, called
through the C ABI (liblqr-1.so -> libb200carve.so), against the oracle on the same seeded inputs.
"""
import importlib
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("gimp-lqr-plugin_b200")


@pytest.fixture(scope="session")
def oracle(pkg):
    if not os.path.exists(pkg.ORACLE_PATH):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")])
    return pkg.load_oracle()


@pytest.fixture(scope="session")
def product(pkg):
    """The shipped CUDA path.  Missing libraries are a hard error on a GPU box, never a skip."""
    return pkg.load_product()
