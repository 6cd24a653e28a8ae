"""The plugin path: user-defined components (the reference's TypedComponent contract with SOPOT_HD markers) compiled
by nvcc into an RK4 ensemble kernel through sopot/b200/generic_ensemble.cuh.

CPU: tests/cuda/user_components.cu -- the reference's own composition example (physics/coupled_oscillator/: two point
masses, a spring providing both forces through the registry, an energy monitor) -- cross-compiles for sm_100a with and
without FMA contraction, and the generic kernel keeps its whole state in registers (no local memory).
GPU: the program runs: reference call shapes on a user pack, ensemble == single solves, and the -fmad=false build
reproduces the reference's CPU result bit for bit."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "cuda", "user_components.cu")


def _nvcc(out, *extra):
    cmd = ["nvcc", "-std=c++20", "--expt-relaxed-constexpr", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-Xptxas", "-v",
           "-I" + os.path.join(ROOT, "sopot_b200", "include"), *extra, "-o", str(out), SRC,
           "-L" + os.path.join(ROOT, "sopot_b200", "lib"), "-lsopot_b200", "-Xlinker", "-rpath", "-Xlinker",
           os.path.join(ROOT, "sopot_b200", "lib")]
    return subprocess.run(cmd, capture_output=True, text=True)


@pytest.fixture(scope="module")
def programs(tmp_path_factory):
    from sopot_b200 import build
    build.build()
    d = tmp_path_factory.mktemp("plugin")
    logs = {}
    for name, extra in (("contract", ()), ("strict", ("-fmad=false", "-DSOPOT_TEST_STRICT"))):
        r = _nvcc(d / name, *extra)
        assert r.returncode == 0, r.stdout + r.stderr
        logs[name] = r.stdout + r.stderr
    return d, logs


def test_user_components_compile_to_register_resident_kernels(programs):
    """ptxas -v: the generic kernels of the coupled-oscillator pack (device factory and single-system instantiation) hold
    the whole state in registers -- no stack frame, no spills.  (The 6-link cart pendulum's 7x7 elimination is allowed a
    small stack frame; only its closed-loop instantiation -- 14 states, 14 gains and the elimination live at once at the
    255-register limit -- may spill, and then no more than a few dozen bytes.)"""
    _, logs = programs
    for log in logs.values():
        kernels = re.findall(r"Compiling entry function '(\S*rk4_kernel\S*)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores", log)
        assert len(kernels) >= 4
        coupled = [k for k in kernels if "coupled" in k[0] and "CartNPendulum" not in k[0]]
        assert len(coupled) >= 2                       # the factory ensemble and the single-system instantiation
        for name, stack, spill in coupled:
            assert int(stack) == 0 and int(spill) == 0, (name, stack, spill)
        for name, _, spill in kernels:
            assert int(spill) <= (128 if "CartNStateFeedback" in name and "ILm6E" in name else 0), (name, spill)


def test_coupled_oscillator_port_equals_reference(port, ref):
    """oracle side of the plugin check: C restatement of the coupled oscillator == the reference's components."""
    rng = np.random.default_rng(2)
    n = 16
    args = (rng.uniform(0.5, 2, n), rng.uniform(0.5, 2, n), rng.uniform(1, 10, n), rng.uniform(0.5, 1.5, n),
            np.where(np.arange(n) % 2, 0.0, rng.uniform(0, 0.3, n)), rng.uniform(-1, 1, n), rng.uniform(1, 3, n))
    a = port.coupled_rk4(*args, 0.0, 3.0, 0.01)
    b = ref.coupled_rk4(*args, 0.0, 3.0, 0.01)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    f, e = port.coupled_rk4(1.0, 1.5, np.array([4.0]), 1.0, 0.0, 0.0, 2.0, 0.0, 5.0, 0.01)      # the .cu test's known answer
    assert f[:, 0].tolist() == [0.1051977204129609, 1.5655568262665147, 1.929868186391363, -1.043704550844343]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["contract", "strict"])
def test_user_components_run_on_the_gpu(programs, mode):
    d, _ = programs
    r = subprocess.run([str(d / mode)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "user_components: OK" in r.stdout


# ---------------------------------------------------------------- lane-parallel Dual linearization, any number of links
LIN_SRC = os.path.join(ROOT, "tests", "cuda", "cartn_linearize.cu")


@pytest.fixture(scope="module")
def lin_program(tmp_path_factory):
    from sopot_b200 import build
    build.build()
    exe = tmp_path_factory.mktemp("cartn") / "cartn_linearize"
    cmd = ["nvcc", "-std=c++20", "--expt-relaxed-constexpr", "-O3", "-gencode", "arch=compute_100a,code=sm_100a",
           "-I" + os.path.join(ROOT, "sopot_b200", "include"), "-o", str(exe), LIN_SRC,
           "-L" + os.path.join(ROOT, "sopot_b200", "lib"), "-lsopot_b200", "-Xlinker", "-rpath", "-Xlinker",
           os.path.join(ROOT, "sopot_b200", "lib")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


def test_lane_parallel_linearization_compiles(lin_program):
    assert os.path.exists(lin_program)


@pytest.mark.gpu
@pytest.mark.parametrize("n_links", [1, 2, 6])
def test_lane_parallel_linearization_matches_reference(lin_program, tmp_path, n_links):
    """Dual partials spread over the lanes of a warp (Dual<double,1> per lane, component code unchanged) == the
    reference's makeLinearizer<2(N+1),1> around CartNPendulum<N, Dual<double,2N+3>> on the CPU (golden vectors),
    for 1, 2 and 6 links; per-point plant parameters."""
    from conftest import golden
    g = golden("cartn_linearize")
    k = f"n{n_links}_"
    P = g[k + "u"].size
    ns = 2 * (n_links + 1)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(np.array([n_links, P], dtype=np.int32).tobytes())
        for name in ("mc", "m", "L", "g", "x", "u"):
            f.write(np.ascontiguousarray(g[k + name], dtype=np.float64).tobytes())
    r = subprocess.run([str(lin_program), str(inp), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    res = np.fromfile(out, dtype=np.float64)
    A, B = res[:ns * ns * P].reshape(ns * ns, P), res[ns * ns * P:].reshape(ns, P)
    scale = max(np.abs(g[k + "A"]).max(), 1.0)
    assert np.abs(A - g[k + "A"]).max() <= 1e-11 * scale
    assert np.abs(B - g[k + "B"]).max() <= 1e-11 * max(np.abs(g[k + "B"]).max(), 1.0)
    # structure: d(position rates)/d(velocities) = identity block, exactly
    for i in range(ns // 2):
        for j in range(ns):
            assert np.all(A[i * ns + j] == (1.0 if j == ns // 2 + i else 0.0))
