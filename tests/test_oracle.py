"""CPU tests (-m "not gpu"): the oracle restatement against the reference's golden vectors.

Pins the checker: (1) known answers quoted from the reference run, (2) the
committed fixtures tests/golden/*.npz produced by the unmodified reference (make_golden.py), bit for
bit, (3) where oracle/_ref exists, the reference itself on fresh random inputs, bit for bit, and
(4) the physics invariants the reference's own tests assert (tests/compile_time_test.cpp:153,190,
tests/verify_grid_physics.cpp:76,116,
tests/rocket_flight_test.cpp:194-209, tests/autodiff_test.cpp:516-523).
"""
import os

import numpy as np
import pytest

from conftest import ROOT, golden
from sopot_b200 import workloads as W


def test_known_answers_ho_dpend(port):
    fin, _, tl = port.ho_rk4([2.0], [9.0], [0.1], [1.5], [0.0], 0.0, 10.0, 0.01)
    assert fin[0, 0] == -0.82129045056399141 and fin[1, 0] == -1.7419130577838189
    assert tl[0] == 9.9999999999998312
    fin, _, _ = port.dpend_rk4([1.0], [1.0], [1.0], [1.0], [9.81], [np.pi / 4], [np.pi / 6], [0.0], [0.0], 0.0, 10.0, 1e-3)
    assert list(fin[:, 0]) == [-2.0583664572549361, -9.8402182120618615, -1.5202026871720531, 0.27300742483572987]


def test_known_answers_linearization(port):
    A, B = port.cartpole_linearize(np.array([[0.0, 0.1], [0.0, 0.3], [0.0, -0.2], [0.0, 0.5]]), [0.0, 1.5], 1.0, 0.1, 0.5, 9.81)
    A = A.reshape(4, 4, 2)
    # upright: analytic cart-pole linearization (physics/control/lqr.hpp:578-681)
    np.testing.assert_allclose(A[:, :, 0], [[0, 0, 1, 0], [0, 0, 0, 1], [0, -0.981, 0, 0], [0, 21.582, 0, 0]], atol=1e-13)
    np.testing.assert_allclose(B[:, 0], [0, 0, 1, -2], atol=1e-13)
    assert A[2, 1, 1] == pytest.approx(-0.858878844120743, rel=1e-14)
    assert A[2, 3, 1] == pytest.approx(0.01464808539168254, rel=1e-14)
    assert A[3, 1, 1] == pytest.approx(21.103512372331434, rel=1e-14)
    assert A[3, 3, 1] == pytest.approx(-0.02798770094100415, rel=1e-14)
    assert B[2, 1] == pytest.approx(0.99134238955571397, rel=1e-14)
    assert B[3, 1] == pytest.approx(-1.8941311159190897, rel=1e-14)


def test_golden_ho(port):
    g = golden("ho")
    fin, traj, tl = port.ho_rk4(g["m"], g["k"], g["c"], g["x0"], g["v0"], 0.0, 10.0, 0.01, 100, 10, nthreads=2)
    assert np.array_equal(fin, g["ref_final"]) and np.array_equal(traj, g["ref_traj"]) and np.array_equal(tl, g["ref_t_last"])


def test_golden_dpend(port):
    g = golden("dpend")
    args = [g[k] for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")]
    fin, traj, _ = port.dpend_rk4(*args, 0.0, 10.0, 1e-3, 1000, 10, nthreads=2)
    assert np.array_equal(fin, g["ref_final"]) and np.array_equal(traj, g["ref_traj"])
    assert np.array_equal(port.dpend_rhs(1.3, 0.7, 0.9, 1.1, 9.81, g["rhs_state"]), g["ref_rhs"])


def test_golden_cartpole(port):
    g = golden("cartpole")
    A, B = port.cartpole_linearize(g["x"], g["u"], g["mc"], g["m"], g["L"], g["g"])
    assert np.array_equal(A, g["ref_A"]) and np.array_equal(B, g["ref_B"])
    fin = port.cartpole_rk4(g["mc"][:16], g["m"][:16], g["L"][:16], g["g"][:16], g["x"][:, :16], g["u"][:16], 0.0, 2.0, 0.01)
    assert np.array_equal(fin, g["ref_rk4_final"])


def test_golden_rocket(port):
    g = golden("rocket")
    fin, stats, traj = port.rocket_rk4(g["elev"], g["az"], g["diam"], g["g0"], g["engine"], g["mass_tab"], None, None,
                                       30.0, 0.01, 500, 6, nthreads=2)
    assert np.array_equal(fin, g["ref_final"]) and np.array_equal(stats, g["ref_stats"]) and np.array_equal(traj, g["ref_traj"])
    for got, key in zip(port.atmosphere(g["atm_h"]), ("ref_atm_T", "ref_atm_P", "ref_atm_rho", "ref_atm_a")):
        assert np.array_equal(got, g[key])
    assert np.array_equal(port.engine_thrust(g["engine"], g["thrust_t"], g["thrust_patm"]), g["ref_thrust"])
    assert np.array_equal(port.rocket_rhs(85.0, 0.0, 0.16, 9.80665, g["engine"], g["mass_tab"], g["rhs_state"]), g["ref_rhs"])
    fin2, stats2, _ = port.rocket_rk4(g["elev"][:4], g["az"][:4], g["diam"][:4], g["g0"][:4], g["engine"], g["mass_tab"],
                                      g["ix_tab"], g["iyz_tab"], 10.0, 0.01)
    assert np.array_equal(fin2, g["ref_final_inertia"]) and np.array_equal(stats2, g["ref_stats_inertia"])
    # integrated time state and apogee of the nominal trajectory as observed in the reference run
    assert fin[0, 0] == 30.00000000000189
    assert 2000.0 < stats[0, 0] < 2150.0


GRID_CASES = [(2, 2, 0), (3, 3, 0), (4, 5, 0), (2, 5, 0), (3, 3, 1), (4, 4, 1), (3, 4, 2), (4, 4, 2)]


@pytest.mark.parametrize("r,c,t", GRID_CASES)
def test_golden_grid2d(port, r, c, t):
    g = golden("grid2d")
    key = f"g{r}x{c}t{t}"
    assert np.array_equal(port.grid2d_initial_state(r, c, 0.5), g[key + "_init"])
    assert np.array_equal(port.grid2d_rhs(r, c, t, 1.3, 0.5, 50.0, 0.5, g[key + "_state"]), g[key + "_rhs"])
    for nt in (1, 2):
        assert np.array_equal(port.grid2d_rk4(r, c, t, 1.3, 0.5, 50.0, 0.5, 1e-3, 200, g[key + "_state"], nthreads=nt),
                              g[key + "_rk4"])


def test_grid2d_degenerate_and_printed_positions(port):
    g = golden("grid2d")
    out = port.grid2d_rhs(3, 3, 0, 1.0, 0.5, 50.0, 0.5, g["g3x3t0_degenerate_state"])
    assert np.array_equal(out, g["g3x3t0_degenerate_rhs"], equal_nan=True)
    # tests/grid_2d_test.cpp prints mass 4 -> (1, 1.05948) for the 3x3 case (k=10, c=0.5, centre +0.5 in y,
    # 2 s at dt=1e-3) -- produced there with the tests' own RK4 association, hence the loose tolerance
    s = port.grid2d_initial_state(3, 3, 1.0)
    s[4 * 4 + 1] += 0.5
    # that test loops `while (t < t_end) t += dt` (tests/grid_2d_test.cpp:78-95): 2001 steps, not 2000
    s = port.grid2d_rk4(3, 3, 0, 1.0, 1.0, 10.0, 0.5, 1e-3, 2001, s)
    assert abs(s[4 * 4 + 0] - 1.0) < 1e-9 and abs(s[4 * 4 + 1] - 1.05948) < 1e-5


def test_grid2d_invariants(port):
    # equilibrium force < 1e-10 and sum of internal forces < 1e-10 (tests/verify_grid_physics.cpp:76,116)
    s = port.grid2d_initial_state(6, 7, 0.5)
    d = port.grid2d_rhs(6, 7, 0, 1.0, 0.5, 50.0, 0.5, s).reshape(-1, 4)
    assert np.max(np.abs(d[:, 2:])) < 1e-10
    rng = np.random.default_rng(0)
    s2 = s + rng.normal(0, 0.05, s.shape)
    for topo in (0, 1, 2):
        d = port.grid2d_rhs(6, 7, topo, 1.0, 0.5, 50.0, 0.5, s2).reshape(-1, 4)
        assert np.max(np.abs(d[:, 2:].sum(axis=0))) < 1e-10


def test_reference_invariants(port):
    # undamped oscillator vs analytic solution < 1e-6, energy drift < 0.01 % (tests/compile_time_test.cpp:153,190)
    fin, traj, _ = port.ho_rk4([1.0], [4.0], [0.0], [1.0], [0.0], 0.0, 2.0, 0.01, 1, 200)
    t = 0.01 * np.arange(1, 201)
    assert np.max(np.abs(traj[:, 0, 0] - np.cos(2 * t))) < 1e-6
    e = 0.5 * traj[:, 1, 0] ** 2 + 0.5 * 4.0 * traj[:, 0, 0] ** 2
    assert np.max(np.abs(e - 2.0)) / 2.0 < 1e-4
    # (the reference's double-pendulum energy test, tests/double_pendulum_test.cpp:208, is not restated:
    # DoublePendulum::derivatives carries the centrifugal terms with the sign of double_pendulum.hpp:168-169,
    # which is what every checker and kernel here reproduces literally; the system is not energy conserving
    # at large angles and parity, not physics, is the contract.)
    # atmosphere at 0 / 10 km (tests/rocket_flight_test.cpp:194-209, loose)
    T, P, rho, a = port.atmosphere([0.0, 10000.0])
    assert abs(T[0] - 288.15) < 0.1 and abs(P[0] - 101325) < 1 and abs(rho[0] - 1.225) < 0.01
    assert abs(T[1] - 223.15) < 1.0 and abs(P[1] - 26436) < 300


def test_port_equals_reference_on_fresh_inputs(port, ref):
    """Only where the reference was compiled (this container): bit-for-bit on new random inputs."""
    w = W.ho_ensemble(48, seed=11)
    a = port.ho_rk4(w["m"], w["k"], w["c"], w["x0"], w["v0"], 0.0, 3.0, 0.01)
    b = ref.ho_rk4(w["m"], w["k"], w["c"], w["x0"], w["v0"], 0.0, 3.0, 0.01)
    assert np.array_equal(a[0], b[0])
    w = W.dpend_ensemble(12, seed=12)
    args = [w[k] for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")]
    assert np.array_equal(port.dpend_rk4(*args, 0.0, 2.0, 1e-3)[0], ref.dpend_rk4(*args, 0.0, 2.0, 1e-3)[0])
    w = W.cartpole_points(500, seed=13)
    a = port.cartpole_linearize(w["x"], w["u"], w["mc"], w["m"], w["L"], w["g"])
    b = ref.cartpole_linearize(w["x"], w["u"], w["mc"], w["m"], w["L"], w["g"])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    w = W.rocket_ensemble(4, seed=14)
    a = port.rocket_rk4(w["elev"], w["az"], w["diam"], w["g0"], w["engine"], w["mass_tab"], None, None, 12.0, 0.01)
    b = ref.rocket_rk4(w["elev"], w["az"], w["diam"], w["g0"], w["engine"], w["mass_tab"], None, None, 12.0, 0.01)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    rng = np.random.default_rng(15)
    for (r, c, t) in GRID_CASES:
        s = port.grid2d_initial_state(r, c, 0.7) + rng.normal(0, 0.2, 4 * r * c)
        assert np.array_equal(port.grid2d_rhs(r, c, t, 0.9, 0.7, 33.0, 0.2, s), ref.grid2d_rhs(r, c, t, 0.9, 0.7, 33.0, 0.2, s))


def test_rk4_step_count_rule(port):
    # core/solver.hpp:91: while (t < t_end - dt*0.5) with accumulated t
    assert port.rk4_num_steps(0.0, 10.0, 0.01) == 1000
    assert port.rk4_num_steps(0.0, 10.0, 1e-3) == 10000
    assert port.rk4_num_steps(0.0, 30.0, 0.01) == 3000
    assert port.rk4_num_steps(0.0, 0.0, 0.01) == 0
    assert port.rk4_num_steps(0.0, 0.014, 0.01) == 1 and port.rk4_num_steps(0.0, 0.016, 0.01) == 2


def test_lean_trig_accuracy(tmp_path):
    """csrc/trig.cuh (the FP_CONTRACT sincos and the small-increment rotation) against long double libm, on the
    host: <= 2 ulp for sincos over |x| <= 1e6, <= 2 ulp(1) absolute for the rotation.  Checks the kernel's own
    header, compiled for the host (same fma() sequence as the device)."""
    import subprocess
    exe = tmp_path / "trig_accuracy"
    src = os.path.join(ROOT, "tests", "cpp", "trig_accuracy.cpp")
    subprocess.run(["g++", "-std=c++17", "-O2", src, "-o", str(exe)], check=True)     # fma() is exact with or without -mfma
    r = subprocess.run([str(exe), "400000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_dpend_fast_step_on_host(tmp_path):
    """csrc/dpend_step.cuh (the FP_CONTRACT step of the double pendulum that carries its sin/cos pairs from stage to stage
    and step to step) compiled for the host -- the kernel's own fma() sequence -- against a long-double RK4 of the
    reference's equations, step by step from identical inputs along whole trajectories (so the rounding carried in the
    pairs counts): a few 1e-15 relative per step on the BASELINE ensemble (the mode promises 1e-12), every fall-back tier
    driven by fast spinners, coarse steps, g = 0 and m2 = 0; the two tiered rotations <= 1 ulp(1)."""
    import subprocess
    exe = tmp_path / "dpend_step_host"
    src = os.path.join(ROOT, "tests", "cpp", "dpend_step_host.cpp")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", src, "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(" ok") == 9, r.stdout


def test_dpend_fast_step_full_horizon_shadow_twin_on_host(tmp_path, port):
    """The timed configuration without a GPU: the kernel's own step (csrc/dpend_step.cuh compiled for the host, the loop of
    rk4_ensemble_kernel: load, refresh every 16th step, snapshots) over all 10 000 steps of the first 256 instances of
    bench.py's inputs, against the oracle by the rule of SURVEY.md section 8c -- error <= C * (separation of the oracle's run
    from its own run with theta1 nudged by one ulp) + 1e-9 at every stored second, <= 1e-9 outright at t = 1 s, C = 2 as in
    tests/test_gpu_parity.py::test_dpend_bench_inputs_full_horizon_shadow_twin (observed here: max (err - 1e-9) / twin = 0.09,
    5.2e-15 at t = 1 s -- the figure bench.py's parity_check reports from the GPU)."""
    import subprocess
    from sopot_b200 import workloads as W
    exe = tmp_path / "dpend_step_host"
    src = os.path.join(ROOT, "tests", "cpp", "dpend_step_host.cpp")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", src, "-o", str(exe)], check=True)
    n = 256
    w = W.dpend_ensemble(1 << 22, seed=2)
    args = [np.ascontiguousarray(w[k][:n]) for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")]
    _, ref, _ = port.dpend_rk4(*args, 0.0, 10.0, 1e-3, store_every=1000, n_snap=10, nthreads=4)
    nudged = list(args)
    nudged[5] = np.nextafter(args[5], np.inf)
    _, ref_twin, _ = port.dpend_rk4(*nudged, 0.0, 10.0, 1e-3, store_every=1000, n_snap=10, nthreads=4)
    np.concatenate(args).tofile(tmp_path / "in.bin")
    subprocess.run([str(exe), "integrate", str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(n), "10000", "1e-3", "1000"], check=True)
    out = np.fromfile(tmp_path / "out.bin").reshape(11, 4, n)
    assert np.array_equal(out[9], out[10])                       # last snapshot == final state

    def rel(a, b):
        return np.max(np.abs(a - b), axis=1) / np.maximum(np.max(np.abs(b), axis=1), 1e-300)

    twin, err = rel(ref_twin, ref), rel(out[:10], ref)
    assert err[0].max() <= 1e-9, err[0].max()
    assert np.all(err <= 2.0 * twin + 1e-9), (np.maximum(err - 1e-9, 0.0) / np.maximum(twin, 1e-300)).max()


def test_golden_lqr_and_closed_loop(port):
    """Section 8f-1: the C restatement of LQR<4,1>::solve and of the sampled-data closed loop against the reference's
    own classes (golden vectors from oracle/ref_harness.cpp): gains, Riccati solutions, status codes incl. the
    'singular' throw and the diverging iteration, closed-loop final states and running maxima -- bit for bit."""
    g = golden("control")
    for tag, r, dt, mi, tol in (("a", 0.01, 0.01, 1000, 1e-8), ("b", 0.5, 0.002, 300, 1e-10)):
        K, P, it, st = port.lqr4(g["A"], g["B"], g["Q"], r, dt, mi, tol, nthreads=4)
        assert np.array_equal(K, g["K_" + tag]) and np.array_equal(P, g["P_" + tag]) and np.array_equal(st, g["st_" + tag])
        assert it.min() >= 1 and it.max() <= mi
    _, _, _, st = port.lqr4(g["A2"], g["B2"], np.eye(4), 0.0)
    assert np.array_equal(st, g["st_sing"]) and list(st) == [2, 2]
    K, _, _, st = port.lqr4(g["A2"], g["B2"], np.eye(4), 1.0)
    assert np.array_equal(st, g["st_div"]) and np.array_equal(K, g["K_div"], equal_nan=True)
    n = g["cl_s0"].shape[1]
    pl = [g[k][:n] for k in ("mc", "m", "L", "g")]
    fin, stats = port.cartpole_feedback_rk4(*pl, g["cl_s0"], g["cl_K"], g["cl_xref"], -15.0, 15.0, 1, 0.0, 0.002, 1500, nthreads=4)
    assert np.array_equal(fin, g["cl_final"]) and np.array_equal(stats, g["cl_stats"])
    fin, stats = port.cartpole_feedback_rk4(*pl, g["cl_s0"], g["cl_K"], None, 0.0, 0.0, 0, 0.0, 0.002, 700, nthreads=4)
    assert np.array_equal(fin, g["cl_final_ns"]) and np.array_equal(stats, g["cl_stats_ns"])
    # the loop is stabilised: every pendulum ends near upright (tests/inverted_pendulum_test.cpp:288-289 style bounds)
    assert np.abs(g["cl_final"][1]).max() < 0.01 and g["cl_stats"][0].max() < 0.2


def test_lqr_port_equals_reference_on_fresh_inputs(port, ref):
    rng = np.random.default_rng(5)
    P = 24
    x = np.stack([rng.uniform(-1, 1, P), rng.uniform(-0.4, 0.4, P), rng.uniform(-1, 1, P), rng.uniform(-1, 1, P)])
    mc, m, L = rng.uniform(0.5, 2, P), rng.uniform(0.05, 0.3, P), rng.uniform(0.3, 1.0, P)
    A, B = ref.cartpole_linearize(x, rng.uniform(-3, 3, P), mc, m, L, np.full(P, 9.81))
    Q = np.diag(rng.uniform(1, 100, 4))
    a = port.lqr4(A, B, Q, 0.05, 0.005, 800, 1e-9)
    b = ref.lqr4(A, B, Q, 0.05, 0.005, 800, 1e-9)
    assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
    s0 = np.zeros((4, P)); s0[1] = rng.uniform(-0.1, 0.1, P)
    fa = port.cartpole_feedback_rk4(mc, m, L, np.full(P, 9.81), s0, a[0], None, -30.0, 30.0, 1, 0.0, 0.001, 400)
    fb = ref.cartpole_feedback_rk4(mc, m, L, np.full(P, 9.81), s0, a[0], None, -30.0, 30.0, 1, 0.0, 0.001, 400)
    assert np.array_equal(fa[0], fb[0]) and np.array_equal(fa[1], fb[1])


def test_golden_dpend_jacobian(port):
    """Forward-mode Jacobian of the double pendulum (computeJacobian on DoublePendulum<Dual<double,4>>): the C dual numbers of
    the port == the reference's, bit for bit; the structure the reference's own test asserts
    (tests/double_pendulum_test.cpp:405-409) and a central-difference cross-check of the angular rows."""
    g = golden("dpend_jacobian")
    J = port.dpend_jacobian(g["m1"], g["m2"], g["L1"], g["L2"], g["g"], g["state"])
    assert np.array_equal(J, g["J"])
    J0 = J[:, 0].reshape(4, 4)
    assert abs(J0[0, 2] - 1.0) < 1e-10 and abs(J0[1, 3] - 1.0) < 1e-10 and abs(J0[0, 0]) < 1e-10 and abs(J0[0, 1]) < 1e-10
    h = 1e-6
    for j in range(4):
        sp, sm = g["state"].copy(), g["state"].copy()
        sp[j] += h; sm[j] -= h
        fd = (port.dpend_rhs(g["m1"], g["m2"], g["L1"], g["L2"], g["g"], sp) - port.dpend_rhs(g["m1"], g["m2"], g["L1"], g["L2"], g["g"], sm)) / (2 * h)
        col = J.reshape(4, 4, -1)[:, j, :]
        assert np.abs(fd - col).max() <= 1e-5 * max(1.0, np.abs(col).max())


def test_golden_rocket_all_atmosphere_layers(port):
    """One derivative per layer of the standard atmosphere and at each layer boundary (the trajectory fixtures stay below
    11 km): isothermal layers, both signs of the lapse rate, the h <= 0 and h >= 100 km clamps -- port == reference."""
    g, ga = golden("rocket"), golden("rocket_atmosphere")
    assert ga["state"][3].min() < 0 and ga["state"][3].max() > 1e5
    assert np.array_equal(port.rocket_rhs(85.0, 0.0, 0.16, 9.80665, g["engine"], g["mass_tab"], ga["state"]), ga["rhs"])


def test_golden_rocket_aero_tables(port):
    """Table-driven aerodynamics (bilinear Cd/Cl/Cs/cmq tables, io/interpolation.hpp:137-181 through
    rocket/axisymmetric_aero.hpp:133-167,276-281): C restatement == the reference's components, bit for bit."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    aero = mg.rocket_aero_tables()
    g = golden("rocket_aero")
    for tag, on in (("on", True), ("off", False)):
        fin, stats, _ = port.rocket_rk4(g["elev"], g["az"], g["diam"], g["g0"], g["engine"], g["mass_tab"], None, None, 10.0, 0.01,
                                        aero=aero, engine_on=on, nthreads=4)
        assert np.array_equal(fin, g["final_" + tag]) and np.array_equal(stats, g["stats_" + tag])
    assert not np.array_equal(g["final_on"], g["final_off"])          # the power-on / power-off tables are both used
    rhs = port.rocket_rhs(84.0, 30.0, 0.16, 9.80665, g["engine"], g["mass_tab"], g["rhs_state"], aero=aero)
    assert np.array_equal(rhs, g["rhs"])


def test_golden_lqr_and_closed_loop_n_links(port):
    """LQR<6,1>, LQR<14,1> and the 2- and 6-link closed loops: C restatement == the reference classes, bit for bit."""
    g = golden("control_n")
    for N in (2, 6):
        k, ns = f"n{N}_", 2 * (N + 1)
        K, P, it, st = port.lqr_n(ns, g[k + "A"], g[k + "B"], g[k + "Q"], 0.001, 0.005, 1000, 1e-8)
        assert np.array_equal(K, g[k + "K"]) and np.array_equal(P, g[k + "P"]) and np.array_equal(st, g[k + "st"])
        fin, stats = port.cartn_feedback_rk4(N, g[k + "mc"], g[k + "m"], g[k + "L"], g[k + "g"], g[k + "s0"], g[k + "K"], None,
                                             -500.0, 500.0, 1, 0.0, 0.001, 1500)
        assert np.array_equal(fin, g[k + "final"]) and np.array_equal(stats, g[k + "stats"])
        assert g[k + "stats"][0].max() < 0.2            # never falls (tests/inverted_pendulum_test.cpp:288)


def _ugrid_params():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    return mg.UGRID_PARAMS


def test_golden_unified_grid(port):
    """Section 8f-2: runtime-sized restatement of the unified 6-state grid == the reference's ValidatedSystem on
    makeGridTopology<R,C> (golden vectors), bit for bit: initial state, derivative with rotation / torque / repulsion, RK4."""
    g = golden("ugrid")
    for rows, cols in ((3, 3), (4, 5), (2, 7), (6, 6), (10, 10)):
        for tag, prm in _ugrid_params().items():
            key = f"u{rows}x{cols}_{tag}"
            assert np.array_equal(port.ugrid(rows, cols, prm, mode="init"), g[key + "_init"])
            assert np.array_equal(port.ugrid(rows, cols, prm, g[key + "_state"]), g[key + "_rhs"])
            assert np.array_equal(port.ugrid(rows, cols, prm, g[key + "_state"], 1e-3, 50, "rk4"), g[key + "_rk4"])
    # Newton's third law on forces; a rest grid with rest_length = attachment distance feels nothing
    prm = [1.0, 0.05, 0.5, 50.0, 0.4, 0.5, 0.0, -1.0, 0.0, 3.14159265358979323846, 3.14159265358979323846 / 2, -3.14159265358979323846 / 2]
    s0 = port.ugrid(7, 9, prm, mode="init")
    assert np.abs(port.ugrid(7, 9, prm, s0)).max() < 1e-12           # spacing 0.5 - 2 * radius 0.05 = rest length 0.4
    d = port.ugrid(4, 5, _ugrid_params()["wasm"], g["u4x5_wasm_state"])
    assert np.abs(d[:, 2:4].sum(axis=0)).max() < 1e-10


def test_golden_unified_grid_triangle_topology(port):
    """The simulator's "triangle" grid type (makeTriangleGridTopology, constexpr_topology.hpp:436-505, populated as
    createTriangleSystem does): diagonal springs with their own rest length and attachment angles, accumulated per mass in
    spring-node order -- port == reference bit for bit (derivative, RK4, velocity Verlet)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    g, gt = golden("ugrid"), golden("ugrid_triangle")
    for rows, cols in ((3, 3), (4, 5), (2, 7), (6, 6)):
        st = g[f"u{rows}x{cols}_wasm_state"]
        key = f"x{rows}x{cols}"
        assert np.array_equal(port.ugrid(rows, cols, mg.UGRID_TRIANGLE, st), gt[key + "_rhs"])
        assert np.array_equal(port.ugrid(rows, cols, mg.UGRID_TRIANGLE, st, 1e-3, 50, "rk4"), gt[key + "_rk4"])
        assert np.array_equal(port.ugrid(rows, cols, mg.UGRID_TRIANGLE, st, 1e-3, 50, "velocity_verlet"), gt[key + "_verlet"])
        assert not np.array_equal(gt[key + "_rhs"], g[f"u{rows}x{cols}_wasm_rhs"])      # the diagonals do act
    # forces still sum to zero (Newton's third law over all eight spring directions)
    d = port.ugrid(4, 5, mg.UGRID_TRIANGLE, g["u4x5_wasm_state"])
    assert np.abs(d[:, 2:4].sum(axis=0)).max() < 1e-9


def test_golden_unified_grid_integrators(port):
    """Grid2DSimulator's symplectic Euler and velocity Verlet (wasm/wasm_grid2d.cpp:329-413) around the reference's
    computeDerivatives: the port repeats them bit for bit; both stay within O(dt) / O(dt^2) of the RK4 trajectory."""
    g, gi = golden("ugrid"), golden("ugrid_integrators")
    prm = _ugrid_params()["wasm"]
    for rows, cols in ((4, 5), (10, 10)):
        st = g[f"u{rows}x{cols}_wasm_state"]
        for mode in ("symplectic_euler", "velocity_verlet"):
            assert np.array_equal(port.ugrid(rows, cols, prm, st, 1e-3, 80, mode), gi[f"u{rows}x{cols}_{mode}"])
        rk = port.ugrid(rows, cols, prm, st, 1e-3, 50, "rk4")
        err = [np.abs(port.ugrid(rows, cols, prm, st, 1e-3, 50, m) - rk).max() for m in ("symplectic_euler", "velocity_verlet")]
        assert err[0] < 0.05 and err[1] < 0.05


def test_golden_generator_reproduces_the_committed_fixtures(ref, tmp_path):
    """tests/golden/make_golden.py, run as a script against the reference harness, regenerates every committed fixture bit for
    bit (needs /root/reference-built oracle/_ref: skipped where the `ref` fixture is unavailable)."""
    import subprocess
    import sys
    gold = os.path.join(ROOT, "tests", "golden")
    env = dict(os.environ, SOPOT_GOLDEN_OUT=str(tmp_path))
    # the atmosphere fixture is derived from rocket.npz of the same output directory: generated in order by main()
    r = subprocess.run([sys.executable, os.path.join(gold, "make_golden.py")], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    names = sorted(f for f in os.listdir(gold) if f.endswith(".npz"))
    assert names == sorted(f for f in os.listdir(tmp_path) if f.endswith(".npz"))
    for f in names:
        a, b = np.load(os.path.join(gold, f)), np.load(tmp_path / f)
        assert sorted(a.keys()) == sorted(b.keys()), f
        for k in a.keys():
            assert np.array_equal(a[k], b[k], equal_nan=True), (f, k)

