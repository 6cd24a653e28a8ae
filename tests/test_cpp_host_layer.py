"""C++20 host layer (sopot_b200/include/sopot/...): the reference's plugin API on top of the C ABI.

CPU: the test programs under tests/cpp (plain main() programs in the style of the reference's tests/) compile
with g++ -std=c++20 and link against libsopot_b200.so, and without a GPU they fail loudly instead of
integrating on the host.  GPU: each program runs and passes -- same call shapes as the reference's tests
(wrapper lambda into RK4Solver::solve, makeLinearizer<4,1> with setControlForce in the dynamics lambda,
simulateRocket, makeGrid2DSystem) checked against known answers of the unmodified reference."""
import os
import subprocess

import pytest

from conftest import ROOT

CPP = os.path.join(ROOT, "tests", "cpp")
PROGRAMS = ["compile_time_test", "double_pendulum_test", "linearization_test", "rocket_test", "grid2d_test", "lqr_test"]


@pytest.fixture(scope="module")
def built():
    from sopot_b200 import build
    build.build()
    r = subprocess.run(["make", "-C", CPP, "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(CPP, "bin")


def test_host_layer_compiles_as_cxx20(built):
    for p in PROGRAMS:
        assert os.path.exists(os.path.join(built, p))


def test_host_layer_has_no_host_integrator(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([os.path.join(built, "compile_time_test")], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("prog", PROGRAMS)
def test_host_layer_program(built, prog):
    r = subprocess.run([os.path.join(built, prog)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert prog + ": OK" in r.stdout
