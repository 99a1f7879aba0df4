import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libphonic_oracle.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _build_oracle():
    if not os.path.exists(ORACLE_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return ORACLE_LIB


@pytest.fixture(scope="session")
def oracle_api():
    """The CPU restatement (test infrastructure) behind the same C-ABI, prefix `po_`."""
    from phonic_b200._capi import CApi
    return CApi(_build_oracle(), "po_")


@pytest.fixture(scope="session")
def oracle_lib():
    import ctypes
    return ctypes.CDLL(_build_oracle())


@pytest.fixture(scope="session")
def cuda_api():
    import phonic_b200
    return phonic_b200.load_api()
