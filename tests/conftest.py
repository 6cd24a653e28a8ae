import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement of the reference (oracle/sopot_oracle.c) -- the checker."""
    from oracle.oracle_py import Checker, build
    build()
    return Checker("port")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled in the build container; skipped where it was never built."""
    from oracle.oracle_py import Checker, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libsopot_ref.so not present (built only where /root/reference exists)")
    return Checker("reference")


@pytest.fixture(scope="session")
def engine():
    """CUDA engine through the C ABI; raises (never falls back) when the library or the GPU is missing."""
    from sopot_b200.capi import Engine
    eng = Engine(0)
    yield eng
    eng.close()


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(a, b, axis=0):
    """max-norm relative error per instance: ||a-b||_inf / max(||b||_inf, tiny) along `axis`."""
    a, b = np.asarray(a), np.asarray(b)
    num = np.max(np.abs(a - b), axis=axis)
    den = np.maximum(np.max(np.abs(b), axis=axis), 1e-300)
    return num / den
