"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle.

Tolerances (north_star): <= 1e-12 relative per step, <= 1e-9 relative on the final state, relative =
max-norm over the state vector of one instance.  FP_STRICT is additionally required to be
BIT-IDENTICAL wherever the path contains no libm transcendental (oscillator, grid2d): there the only
possible difference to the reference is FMA contraction, which strict mode forbids.
The chaotic double pendulum uses the shadow-twin rule of SURVEY.md section 8c beyond t = 1 s.
"""
import numpy as np
import pytest

from conftest import golden, rel_err
from sopot_b200 import workloads as W
from sopot_b200.capi import FP_CONTRACT, FP_STRICT, MEM_DEVICE, MEM_HOST, ST_NONFINITE, ST_OK

pytestmark = pytest.mark.gpu

PER_STEP = 1e-12
FINAL = 1e-9


# ------------------------------------------------------------------------------------------------ config 1
def test_ho_golden_bit_exact_strict(engine):
    g = golden("ho")
    st, traj, status = engine.ho_rk4(g["m"], g["k"], g["c"], g["x0"], g["v0"], 64, 0.0, 0.01, 1000, fp_mode=FP_STRICT,
                                     store_every=100)
    assert np.array_equal(st, g["ref_final"]) and np.array_equal(traj, g["ref_traj"])
    assert st[0, 0] == -0.82129045056399141 and st[1, 0] == -1.7419130577838189      # README known answer
    assert np.all(status == ST_OK)


def test_ho_contract_within_tolerance(engine, port):
    g = golden("ho")
    one, _, _ = engine.ho_rk4(g["m"], g["k"], g["c"], g["x0"], g["v0"], 64, 0.0, 0.01, 1, fp_mode=FP_CONTRACT)
    ref1, _, _ = port.ho_rk4(g["m"], g["k"], g["c"], g["x0"], g["v0"], 0.0, 0.01, 0.01)
    assert rel_err(one, ref1).max() <= PER_STEP
    st, traj, _ = engine.ho_rk4(g["m"], g["k"], g["c"], g["x0"], g["v0"], 64, 0.0, 0.01, 1000, fp_mode=FP_CONTRACT,
                                store_every=100)
    assert rel_err(st, g["ref_final"]).max() <= FINAL
    assert rel_err(traj, g["ref_traj"], axis=1).max() <= FINAL


@pytest.mark.parametrize("mem", [MEM_HOST, MEM_DEVICE])
def test_ho_full_config_vs_oracle(engine, port, mem):
    """BASELINE config 1 at full size (2^20 instances, 1000 steps): strict == oracle bit for bit on a
    contiguous + strided sample, contract within 1e-9; host and device pointer paths agree exactly."""
    n = 1 << 20
    w = W.ho_ensemble(n)
    st, _, status = engine.ho_rk4(w["m"], w["k"], w["c"], w["x0"], w["v0"], n, 0.0, 0.01, 1000, mem=mem, fp_mode=FP_STRICT)
    st = st.download() if mem == MEM_DEVICE else st
    status = status.download() if mem == MEM_DEVICE else status
    idx = np.concatenate([np.arange(2048), np.arange(0, n, 997), [n - 1]])
    ref, _, _ = port.ho_rk4(w["m"][idx], w["k"][idx], w["c"][idx], w["x0"][idx], w["v0"][idx], 0.0, 10.0, 0.01, nthreads=4)
    assert np.array_equal(st[:, idx], ref)
    assert np.all(status == ST_OK)
    fast, _, _ = engine.ho_rk4(w["m"], w["k"], w["c"], w["x0"], w["v0"], n, 0.0, 0.01, 1000, mem=mem, fp_mode=FP_CONTRACT)
    fast = fast.download() if mem == MEM_DEVICE else fast
    assert rel_err(fast, st).max() <= FINAL


def test_ho_edge_cases(engine):
    # empty ensemble, single instance, ragged size (not a multiple of the block), broadcast parameters
    st, _, _ = engine.ho_rk4(np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), 0, 0.0, 0.01, 10)
    assert st.shape == (2, 0)
    a, _, _ = engine.ho_rk4([2.0], [9.0], [0.1], [1.5], [0.0], 1, 0.0, 0.01, 1000, fp_mode=FP_STRICT)
    assert a[0, 0] == -0.82129045056399141
    n = 1000 + 37
    k = 9.0 * (1 + 0.01 * np.arange(n) / n)
    b, _, _ = engine.ho_rk4(2.0, k, 0.1, 1.5, 0.0, n, 0.0, 0.01, 100, fp_mode=FP_STRICT)            # scalars broadcast
    c, _, _ = engine.ho_rk4(np.full(n, 2.0), k, np.full(n, 0.1), np.full(n, 1.5), np.zeros(n), n, 0.0, 0.01, 100, fp_mode=FP_STRICT)
    assert np.array_equal(b, c)
    # zero steps returns the initial condition; non-finite input is flagged, never thrown
    z, _, _ = engine.ho_rk4(2.0, k, 0.1, 1.5, 0.25, n, 0.0, 0.01, 0)
    assert np.all(z[0] == 1.5) and np.all(z[1] == 0.25)
    bad, _, status = engine.ho_rk4(2.0, np.where(np.arange(n) == 5, np.nan, k), 0.1, 1.5, 0.0, n, 0.0, 0.01, 10)
    assert status[5] == ST_NONFINITE and status.sum() == ST_NONFINITE
    # linearity of the (linear) oscillator flow: scaling the IC scales the solution (size-independent property)
    s1, _, _ = engine.ho_rk4(2.0, k, 0.1, 1.5, 0.0, n, 0.0, 0.01, 500, fp_mode=FP_STRICT)
    s2, _, _ = engine.ho_rk4(2.0, k, 0.1, 3.0, 0.0, n, 0.0, 0.01, 500, fp_mode=FP_STRICT)
    assert np.array_equal(2.0 * s1, s2)     # power-of-two scaling commutes with every rounding


def test_bad_arguments_report_errors(engine):
    from sopot_b200.capi import SopotB200Error
    with pytest.raises(SopotB200Error, match="dt must be positive"):
        engine.ho_rk4([2.0], [9.0], [0.1], [1.5], [0.0], 1, 0.0, -0.01, 10)
    with pytest.raises(SopotB200Error):
        engine.grid2d_rk4(engine.grid2d_params(1, 5, 0, 1.0, 0.5, 50.0, 0.5), 1e-3, 1, np.zeros((5, 4)))
    # every entry point validates opts (mem, fp_mode in {0, 1}), not only the ensemble integrators
    st = np.zeros((4, 3))
    with pytest.raises(SopotB200Error, match="bad opts"):
        engine.dpend_rhs(1.0, 1.0, 1.0, 1.0, 9.81, st, fp_mode=7)
    with pytest.raises(SopotB200Error, match="bad opts"):
        engine.grid2d_rhs(engine.grid2d_params(4, 4, 0, 1.0, 0.5, 50.0, 0.5), np.zeros((16, 4)), fp_mode=2)
    # a derivative evaluation without initial-condition arrays, all parameters broadcast, device-resident state: the
    # missing IC slots alias `state` as arrays (never dereferenced on the host)
    from sopot_b200.capi import DpendParams, RK4Opts, _ptr
    import ctypes
    d_state, d_out = engine.to_device(np.full((4, 3), 0.25)), engine.empty((4, 3))
    prm = [np.array([v]) for v in (1.0, 1.0, 1.0, 1.0, 9.81)]
    p = DpendParams(*[a.ctypes.data for a in prm], None, None, None, None)
    opts = RK4Opts(MEM_DEVICE, FP_STRICT, 0, 0x1FF, 0)
    engine._ck(engine.lib.sopot_b200_dpend_rhs(engine.ctx, ctypes.byref(p), 3, ctypes.byref(opts), _ptr(d_state), _ptr(d_out)))
    engine.sync()
    assert np.array_equal(d_out.download(), engine.dpend_rhs(1.0, 1.0, 1.0, 1.0, 9.81, np.full((4, 3), 0.25), fp_mode=FP_STRICT))


# ------------------------------------------------------------------------------------------------ config 2
def _dp_args(g):
    return [g[k] for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")]


@pytest.mark.parametrize("fp", [FP_STRICT, FP_CONTRACT])
def test_dpend_single_rhs_and_single_step(engine, port, fp):
    g = golden("dpend")
    rhs = engine.dpend_rhs(1.3, 0.7, 0.9, 1.1, 9.81, g["rhs_state"], fp_mode=fp)
    assert rel_err(rhs, g["ref_rhs"]).max() <= PER_STEP
    w = W.dpend_ensemble(4096, seed=21)
    args = _dp_args(w)
    one, _, _ = engine.dpend_rk4(*args, 4096, 0.0, 1e-3, 1, fp_mode=fp)
    ref1, _, _ = port.dpend_rk4(*args, 0.0, 1e-3, 1e-3, nthreads=4)
    assert rel_err(one, ref1).max() <= PER_STEP


@pytest.mark.parametrize("fp", [FP_STRICT, FP_CONTRACT])
def test_dpend_golden_shadow_twin(engine, fp):
    """<= 1e-9 at t = 1 s (hard gate); later snapshots: error <= C * (separation of the reference run
    from its own 1-ulp-perturbed twin) + 1e-9, C = 1e3 (10^4 steps each contribute an ulp-level
    perturbation of their own)."""
    g = golden("dpend")
    st, traj, status = engine.dpend_rk4(*_dp_args(g), 16, 0.0, 1e-3, 10000, fp_mode=fp, store_every=1000)
    assert np.all(status == ST_OK)
    err = rel_err(traj, g["ref_traj"], axis=1)                 # [snap][instance]
    twin = rel_err(g["ref_traj_twin"], g["ref_traj"], axis=1)
    assert err[0].max() <= FINAL, err[0]
    assert np.all(err <= 1e3 * twin + FINAL), (err / (twin + 1e-300)).max()
    assert np.array_equal(st, traj[-1])
    # the documented non-chaotic nominal case stays inside 1e-9 for the whole 10 s
    assert err[:, 0].max() <= FINAL


def test_dpend_bench_inputs_full_horizon_shadow_twin(engine, port):
    """The TIMED configuration against the oracle: the first 256 instances of bench.py's own inputs (W.dpend_ensemble(seed=2),
    theta ~ U(-3, 3): most of them chaotic), all 10 000 steps, both arithmetic modes.  Rule of SURVEY.md section 8c: error <=
    C * (separation of the checker's run from its own run with theta1 nudged by one ulp) + 1e-9 at every stored second, and
    <= 1e-9 outright at t = 1 s.  C is measured, not chosen: tools/dpend_twin_measure.py gives max (err - 1e-9) / twin = 0.149
    (contract) and 0.111 (strict) on these instances -- the engine's error is what a one-ulp change of the input does (median
    err / twin at 10 s: 1.06 and 0.78) -- so C = 2 is more than 10x the observed maximum (round 1 used 1e3)."""
    C = 2.0
    n = 256
    w = W.dpend_ensemble(1 << 22, seed=2)
    args = [np.ascontiguousarray(a[:n]) for a in _dp_args(w)]
    _, ref, _ = port.dpend_rk4(*args, 0.0, 10.0, 1e-3, store_every=1000, n_snap=10, nthreads=4)
    nudged = list(args)
    nudged[5] = np.nextafter(args[5], np.inf)
    _, ref_twin, _ = port.dpend_rk4(*nudged, 0.0, 10.0, 1e-3, store_every=1000, n_snap=10, nthreads=4)
    twin = rel_err(ref_twin, ref, axis=1)
    for fp in (FP_CONTRACT, FP_STRICT):
        _, traj, status = engine.dpend_rk4(*args, n, 0.0, 1e-3, 10000, fp_mode=fp, store_every=1000)
        err = rel_err(traj, ref, axis=1)
        assert np.all(status == ST_OK)
        assert err[0].max() <= FINAL, (fp, err[0].max())
        worst = (np.maximum(err - FINAL, 0.0) / np.maximum(twin, 1e-300)).max()
        print("fp_mode %d: max (err - 1e-9) / twin = %.3f, instances beyond 1e-9 at 10 s: %d" % (fp, worst, int((err[-1] > FINAL).sum())))
        assert np.all(err <= C * twin + FINAL), (fp, worst)


def test_dpend_full_horizon_sample_vs_oracle(engine, port):
    n = 1 << 16
    w = W.dpend_ensemble(n)
    args = _dp_args(w)
    st, _, status = engine.dpend_rk4(*args, n, 0.0, 1e-3, 1000, fp_mode=FP_CONTRACT)
    idx = np.arange(0, n, 257)
    ref, _, _ = port.dpend_rk4(*[a[idx] for a in args], 0.0, 1.0, 1e-3, nthreads=4)
    assert rel_err(st[:, idx], ref).max() <= FINAL
    assert np.all(status == ST_OK)
    # device-resident call gives the same bits as the host-staged call
    d, _, _ = engine.dpend_rk4(*args, n, 0.0, 1e-3, 1000, mem=MEM_DEVICE, fp_mode=FP_CONTRACT)
    assert np.array_equal(d.download(), st)


def test_dpend_contract_error_per_step_across_a_refresh_period(engine, port):
    """The contract-mode step carries rotated sin/cos pairs for 16 steps between full evaluations: the error against the
    oracle after k = 1 .. 35 steps (two refresh periods and a bit) stays within k * 1e-12, i.e. the per-step promise holds at
    every phase of the period, not only right after a refresh."""
    n = 2048
    w = W.dpend_ensemble(n, seed=33)
    args = _dp_args(w)
    worst = 0.0
    for k in (1, 2, 3, 7, 15, 16, 17, 31, 32, 35):
        got, _, status = engine.dpend_rk4(*args, n, 0.0, 1e-3, k, fp_mode=FP_CONTRACT)
        ref, _, _ = port.dpend_rk4(*args, 0.0, k * 1e-3, 1e-3, nthreads=4)
        err = rel_err(got, ref).max()
        worst = max(worst, err / k)
        assert np.all(status == ST_OK) and err <= k * PER_STEP, (k, err)
    assert worst <= PER_STEP


@pytest.mark.parametrize("fp", [FP_STRICT, FP_CONTRACT])
def test_dpend_jacobian_golden_and_batch(engine, port, fp):
    """Batched forward-mode Jacobians of the double pendulum (sopot_b200_dpend_jacobian) against the reference's computeJacobian
    (golden vectors) and, for a ragged batch through device pointers, against the oracle; broadcast parameters."""
    g = golden("dpend_jacobian")
    J = engine.dpend_jacobian(g["m1"], g["m2"], g["L1"], g["L2"], g["g"], g["state"], fp_mode=fp)
    scale = np.abs(g["J"]).reshape(4, 4, -1).max(axis=(0, 1))
    assert (np.abs(J - g["J"]) / scale).max() <= PER_STEP
    n = 10007
    rng = np.random.default_rng(23)
    st = np.stack([rng.uniform(-3, 3, n), rng.uniform(-3, 3, n), rng.uniform(-5, 5, n), rng.uniform(-5, 5, n)])
    m1, L2 = rng.uniform(0.5, 2, n), rng.uniform(0.5, 2, n)
    ref = port.dpend_jacobian(m1, np.full(n, 0.7), np.full(n, 1.3), L2, np.full(n, 9.81), st)
    dJ = engine.dpend_jacobian(engine.to_device(m1), 0.7, 1.3, engine.to_device(L2), 9.81, engine.to_device(st), fp_mode=fp, mem=MEM_DEVICE)
    scale = np.abs(ref).reshape(4, 4, -1).max(axis=(0, 1))
    assert (np.abs(dJ.download() - ref) / scale).max() <= PER_STEP


@pytest.mark.parametrize("fp", [FP_STRICT, FP_CONTRACT])
def test_dpend_degenerate_parameters_and_large_increments(engine, port, fp):
    """The contract-mode step carries sin/cos pairs scaled by g/L1 and L1/L2 and rotates them by small increments:
    g = 0 (first pair identically zero), m2 = 0 (kap = 0: a simple pendulum), very unequal arms, rates so large that
    the increments leave the small-angle range (|dt w| > 2^-5: full-sincos fallback at every stage) and a coarse step
    that defeats the tiny-increment shortcut -- all against the oracle."""
    n = 64
    rng = np.random.default_rng(5)
    base = dict(m1=rng.uniform(0.5, 2, n), m2=rng.uniform(0.5, 2, n), L1=rng.uniform(0.5, 2, n), L2=rng.uniform(0.5, 2, n),
                g=np.full(n, 9.81), th1=rng.uniform(-3, 3, n), th2=rng.uniform(-3, 3, n), w1=rng.uniform(-2, 2, n), w2=rng.uniform(-2, 2, n))
    cases = {"g0": dict(g=np.zeros(n)), "m2_0": dict(m2=np.zeros(n)), "arms": dict(L1=np.full(n, 0.05), L2=np.full(n, 3.0)),
             "fast": dict(w1=rng.uniform(-80, 80, n), w2=rng.uniform(-80, 80, n))}
    for name, over in cases.items():
        w = {**base, **over}
        args = _dp_args(w)
        # coarse steps: increments of 0.04 .. 0.4 rad (RK4 itself stays stable; at dt = 0.02 the 80 rad/s case blows up to
        # 1e80 in the reference as well)
        for dt, steps in ((1e-3, 40), (5e-3, 6) if name == "fast" else (2e-2, 3)):
            got, _, status = engine.dpend_rk4(*args, n, 0.0, dt, steps, fp_mode=fp)
            ref, _, _ = port.dpend_rk4(*args, 0.0, dt * steps, dt, nthreads=2)
            assert np.all(status == ST_OK), name
            assert rel_err(got, ref).max() <= 100 * PER_STEP, (name, dt, rel_err(got, ref).max())


# ------------------------------------------------------------------------------------------------ config 3
@pytest.mark.parametrize("fp", [FP_STRICT, FP_CONTRACT])
def test_cartpole_linearize_golden(engine, fp):
    g = golden("cartpole")
    P = g["u"].size
    A, B, status = engine.cartpole_linearize(g["mc"], g["m"], g["L"], g["g"], g["x"], g["u"], P, fp_mode=fp)
    scale = np.maximum(np.abs(g["ref_A"]).max(axis=0), 1.0)
    assert (np.abs(A - g["ref_A"]).max(axis=0) / scale).max() <= PER_STEP
    assert (np.abs(B - g["ref_B"]).max(axis=0) / np.maximum(np.abs(g["ref_B"]).max(axis=0), 1.0)).max() <= PER_STEP
    assert np.all(status == ST_OK)
    A4 = A.reshape(4, 4, P)
    # structure the reference tests assert (tests/inverted_pendulum_test.cpp:186-203): kinematic rows exact
    assert np.all(A4[0, 2] == 1.0) and np.all(A4[1, 3] == 1.0) and np.all(A4[0, [0, 1, 3]] == 0.0) and np.all(B[:2] == 0.0)


def test_cartpole_linearize_full_config(engine, port):
    P = 1 << 20
    w = W.cartpole_points(P)
    A, B, status = engine.cartpole_linearize(1.0, 0.1, 0.5, 9.81, w["x"], w["u"], P, fp_mode=FP_CONTRACT)   # broadcast plant
    idx = np.arange(0, P, 101)
    rA, rB = port.cartpole_linearize(w["x"][:, idx], w["u"][idx], 1.0, 0.1, 0.5, 9.81, nthreads=4)
    assert (np.abs(A[:, idx] - rA).max(axis=0) / np.maximum(np.abs(rA).max(axis=0), 1.0)).max() <= PER_STEP
    assert (np.abs(B[:, idx] - rB).max(axis=0) / np.maximum(np.abs(rB).max(axis=0), 1.0)).max() <= PER_STEP
    assert np.all(status == ST_OK)
    dx, du = engine.to_device(w["x"]), engine.to_device(w["u"])
    Ad, Bd, _ = engine.cartpole_linearize(1.0, 0.1, 0.5, 9.81, dx, du, P, mem=MEM_DEVICE, fp_mode=FP_CONTRACT)
    assert np.array_equal(Ad.download(), A) and np.array_equal(Bd.download(), B)


def test_cartpole_rk4_golden(engine):
    g = golden("cartpole")
    x = g["x"][:, :16]
    st, _, status = engine.cartpole_rk4(g["mc"][:16], g["m"][:16], g["L"][:16], g["g"][:16], x[0], x[1], x[2], x[3],
                                        g["u"][:16], 16, 0.0, 0.01, 200, fp_mode=FP_STRICT)
    assert rel_err(st, g["ref_rk4_final"]).max() <= FINAL
    assert np.all(status == ST_OK)


def test_cartpole_contract_step_vs_oracle(engine, port):
    """The contract-mode cart-pole step (closed-form 2x2 solve, carried sin/cos pair) against the oracle's pivoted
    elimination: swinging poles far from upright, rates and steps large enough to leave the small- and tiny-increment
    ranges (full-sincos fallbacks), more steps than the 16-step refresh period, and a massless cart (mc = 0: the generic
    step with the singular-pivot check runs instead of the closed form)."""
    n = 96
    rng = np.random.default_rng(8)
    mc, m, L = rng.uniform(0.5, 2.0, n), rng.uniform(0.05, 0.5, n), rng.uniform(0.3, 1.2, n)
    g = np.full(n, 9.81)
    mc[:8] = 0.0
    s0 = np.stack([rng.uniform(-1, 1, n), rng.uniform(-3.1, 3.1, n), rng.uniform(-1, 1, n), rng.uniform(-3, 3, n)])
    s0[1, :8] = rng.uniform(0.3, 1.2, 8)                       # massless carts: keep cos(theta) away from +-1 (D = m sin^2)
    force = rng.uniform(-5, 5, n)
    for dt, steps, w_scale in ((2e-3, 100, 1.0), (1e-2, 40, 1.0), (1e-3, 30, 25.0)):
        st0 = s0.copy(); st0[3] *= w_scale
        got, _, status = engine.cartpole_rk4(mc, m, L, g, st0[0], st0[1], st0[2], st0[3], force, n, 0.0, dt, steps, fp_mode=FP_CONTRACT)
        ref = port.cartpole_rk4(mc, m, L, g, st0, force, 0.0, dt * steps, dt, nthreads=2)
        # a massless cart whose pole swings through cos(theta) = +-1 hits the reference's singular-pivot throw: reported as
        # a status, not compared
        ok = status == ST_OK
        assert ok[8:].all() and ok.sum() >= n - 8
        assert rel_err(got[:, ok], ref[:, ok]).max() <= FINAL, (dt, steps, rel_err(got[:, ok], ref[:, ok]).max())
    # closed loop, contract mode, saturating controller on large initial angles
    K = np.tile(np.array([-3.0, 40.0, -4.0, 8.0])[:, None], (1, n))
    st0 = s0.copy(); st0[1] = rng.uniform(-0.6, 0.6, n); st0[3] = 0.0
    fin, stats, _, status = engine.cartpole_feedback_rk4(mc, m, L, g, st0[0], st0[1], st0[2], st0[3], K, n, 0.0, 0.002, 400,
                                                         u_min=-20.0, u_max=20.0, saturate=True, fp_mode=FP_CONTRACT)
    rfin, rstats = port.cartpole_feedback_rk4(mc, m, L, g, st0, K, None, -20.0, 20.0, 1, 0.0, 0.002, 400, nthreads=2)
    ok = np.isfinite(rfin).all(axis=0) & (np.abs(rfin).max(axis=0) < 1e3)
    assert ok.sum() > n // 2
    assert rel_err(fin[:, ok], rfin[:, ok]).max() <= 1e-7 and rel_err(stats[:, ok], rstats[:, ok]).max() <= 1e-7


# ------------------------------------------------------------------------------------------------ config 4
@pytest.mark.parametrize("fp", [FP_STRICT, FP_CONTRACT])
def test_rocket_rhs_golden(engine, fp):
    g = golden("rocket")
    out = engine.rocket_rhs(85.0, 0.0, 0.16, 9.80665, g["engine"], g["mass_tab"], g["rhs_state"], fp_mode=fp)
    ref = g["ref_rhs"]
    # derivative blocks have very different magnitudes: compare each physical block with its own scale
    for lo, hi in ((0, 1), (1, 4), (4, 7), (7, 11), (11, 14)):
        scale = np.maximum(np.abs(ref[lo:hi]).max(axis=0), 1e-6)
        assert (np.abs(out[lo:hi] - ref[lo:hi]).max(axis=0) / scale).max() <= 1e-11, (lo, hi)


@pytest.mark.parametrize("fp", [FP_STRICT, FP_CONTRACT])
def test_rocket_rhs_all_atmosphere_layers(engine, fp):
    """Every layer of the standard atmosphere, its boundaries and the two clamps (tests/golden/rocket_atmosphere.npz, from
    the reference): exercises the isothermal exponential, both lapse signs of the power law (contract mode: the lean
    log / exp of trig.cuh) and the layer search above the lowest layer."""
    g, ga = golden("rocket"), golden("rocket_atmosphere")
    out = engine.rocket_rhs(85.0, 0.0, 0.16, 9.80665, g["engine"], g["mass_tab"], ga["state"], fp_mode=fp)
    ref = ga["rhs"]
    for lo, hi in ((0, 1), (1, 4), (4, 7), (7, 11), (11, 14)):
        scale = np.maximum(np.abs(ref[lo:hi]).max(axis=0), 1e-6)
        assert (np.abs(out[lo:hi] - ref[lo:hi]).max(axis=0) / scale).max() <= 1e-11, (lo, hi)


@pytest.mark.parametrize("fp", [FP_STRICT, FP_CONTRACT])
def test_rocket_golden_trajectories(engine, fp):
    g = golden("rocket")
    st, stats, traj, status = engine.rocket_rk4(g["elev"], g["az"], g["diam"], g["g0"], g["engine"], g["mass_tab"], None,
                                                None, 8, 0.01, 3000, fp_mode=fp, store_every=500)
    assert np.all(status == ST_OK)
    assert rel_err(st, g["ref_final"]).max() <= FINAL
    assert rel_err(traj, g["ref_traj"], axis=1).max() <= FINAL
    ref = g["ref_stats"]
    assert np.all(np.abs(stats - ref) <= FINAL * np.maximum(np.abs(ref), 1.0))
    assert np.array_equal(stats[[1, 3]], ref[[1, 3]])          # apogee / max-speed times are exact step times


def test_rocket_inertia_tables_and_single_step(engine, port):
    g = golden("rocket")
    st, stats, _, _ = engine.rocket_rk4(g["elev"][:4], g["az"][:4], g["diam"][:4], g["g0"][:4], g["engine"], g["mass_tab"],
                                        g["ix_tab"], g["iyz_tab"], 4, 0.01, 1000, fp_mode=FP_STRICT)
    assert rel_err(st, g["ref_final_inertia"]).max() <= FINAL
    one, _, _, _ = engine.rocket_rk4(g["elev"], g["az"], g["diam"], g["g0"], g["engine"], g["mass_tab"], None, None, 8,
                                     0.01, 1, fp_mode=FP_CONTRACT, want_stats=False)
    ref1, _, _ = port.rocket_rk4(g["elev"], g["az"], g["diam"], g["g0"], g["engine"], g["mass_tab"], None, None, 0.01, 0.01)
    assert rel_err(one, ref1).max() <= PER_STEP
    # no tables at all: constants 100 / 10 / 100, no thrust (rocket_body.hpp:36-39) -> ballistic drop stays on the pad
    st0, _, _, _ = engine.rocket_rk4(g["elev"][:4], g["az"][:4], g["diam"][:4], g["g0"][:4], None, None, None, None, 4, 0.01, 100)
    ref0, _, _ = port.rocket_rk4(g["elev"][:4], g["az"][:4], g["diam"][:4], g["g0"][:4], None, None, None, None, 1.0, 0.01)
    assert rel_err(st0, ref0).max() <= FINAL


def test_rocket_state_functions_golden(engine):
    """Every state function the reference's registry answers for the full rocket (rocket_tags.hpp; what queryStateFunction<Tag>
    and SimulationResult::query<Tag>(t) return), from the device record, against the reference at states in every atmosphere
    layer (burning and coasting) and with coefficient tables loaded."""
    g, r = golden("rocket_state_functions"), golden("rocket")
    got = engine.rocket_state_functions(85.0, 0.0, 0.16, 9.80665, r["engine"], r["mass_tab"], g["state"])
    # vectors: relative to the vector's norm (components that cancel to ~0 carry the rounding of the large terms)
    for name, rows in engine.ROCKET_PROBE.items():
        ref, mine = g["values"][rows], got[rows]
        den = np.maximum(np.linalg.norm(np.atleast_2d(ref), axis=0), 1e-6)
        assert (np.abs(mine - ref) / den).max() <= PER_STEP, (name, (np.abs(mine - ref) / den).max())
    a = golden("rocket_aero")
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    got = engine.rocket_state_functions(84.0, 30.0, 0.16, 9.80665, a["engine"], a["mass_tab"], g["aero_state"], aero=mg.rocket_aero_tables())
    for name, rows in engine.ROCKET_PROBE.items():
        ref, mine = g["aero_values"][rows], got[rows]
        den = np.maximum(np.linalg.norm(np.atleast_2d(ref), axis=0), 1e-6)
        assert (np.abs(mine - ref) / den).max() <= PER_STEP, ("tables", name, (np.abs(mine - ref) / den).max())


def test_rocket_monte_carlo_sample_and_dispersion(engine, port):
    n = 1 << 14
    w = W.rocket_ensemble(n)
    st, stats, _, status = engine.rocket_rk4(w["elev"], w["az"], w["diam"], w["g0"], w["engine"], w["mass_tab"], None, None,
                                             n, 0.01, 3000, mem=MEM_DEVICE, fp_mode=FP_CONTRACT)
    hstats, hst = stats.download(), st.download()
    assert np.all(status.download() == ST_OK)
    idx = np.arange(0, n, 509)
    rfin, rstats, _ = port.rocket_rk4(w["elev"][idx], w["az"][idx], w["diam"][idx], w["g0"][idx], w["engine"], w["mass_tab"],
                                      None, None, 30.0, 0.01, nthreads=4)
    assert rel_err(hst[:, idx], rfin).max() <= FINAL
    assert np.all(np.abs(hstats[:, idx] - rstats) <= FINAL * np.maximum(np.abs(rstats), 1.0))
    # K4: device reduction of the per-trajectory rows against numpy
    disp = engine.dispersion_stats(stats)
    for r, d in enumerate(disp):
        row = hstats[r]
        assert d["count"] == n and d["nonfinite"] == 0
        assert d["min"] == row.min() and d["max"] == row.max()
        assert abs(d["mean"] - row.mean()) <= 1e-12 * max(abs(row.mean()), 1.0)
        assert abs(d["stddev"] - row.std()) <= 1e-6 * max(row.std(), 1e-9)
    # a narrow spread on a large mean (apogee-like: 1e5 m +- 1 m): the centred second moment keeps the standard deviation
    # to ~1e-10 relative where sumsq/count - mean^2 on raw moments would keep ~6 digits
    rng = np.random.default_rng(7)
    v = np.stack([1.0e5 + rng.normal(0.0, 1.0, 200001), 3.0e7 + rng.normal(0.0, 1e-2, 200001), np.full(200001, 12345.678)])
    for r, dd in enumerate(engine.dispersion_stats(v)):
        assert abs(dd["mean"] - v[r].mean()) <= 1e-13 * abs(v[r].mean())
        assert abs(dd["stddev"] - v[r].std()) <= 1e-9 * max(v[r].std(), 1e-6), (r, dd["stddev"], v[r].std())
    # non-finite entries are skipped and counted, ragged sizes reduce correctly
    v = np.arange(7 * 1001, dtype=np.float64).reshape(7, 1001)
    v[3, 17] = np.nan
    d = engine.dispersion_stats(v)
    assert d[3]["nonfinite"] == 1 and d[3]["count"] == 1000 and d[0]["sum"] == v[0].sum() and d[6]["max"] == v[6].max()


# ------------------------------------------------------------------------------------------------ config 5
GRID_CASES = [(2, 2, 0), (3, 3, 0), (4, 5, 0), (2, 5, 0), (3, 3, 1), (4, 4, 1), (3, 4, 2), (4, 4, 2)]


@pytest.mark.parametrize("r,c,t", GRID_CASES)
def test_grid2d_golden_bit_exact(engine, r, c, t):
    """Strict mode == the templated reference system bit for bit (rhs and 200 RK4Solver steps)."""
    g = golden("grid2d")
    key = f"g{r}x{c}t{t}"
    gp = engine.grid2d_params(r, c, t, 1.3, 0.5, 50.0, 0.5)
    assert np.array_equal(engine.grid2d_initial_state(gp).ravel(), g[key + "_init"])
    s = g[key + "_state"].reshape(-1, 4).copy()
    assert np.array_equal(engine.grid2d_rhs(gp, s, fp_mode=FP_STRICT).ravel(), g[key + "_rhs"])
    for kernel in ((None, "march", "tile") if t == 0 else (None,)):     # quad topology: both fused-step kernels
        out = engine.grid2d_rk4(gp, 1e-3, 200, s.copy(), fp_mode=FP_STRICT, kernel=kernel)
        assert np.array_equal(out.ravel(), g[key + "_rk4"]), kernel
        fast = engine.grid2d_rk4(gp, 1e-3, 200, s.copy(), fp_mode=FP_CONTRACT, kernel=kernel)
        assert rel_err(fast.ravel(), g[key + "_rk4"]).max() <= FINAL, kernel


def test_grid2d_degenerate_spring(engine):
    g = golden("grid2d")
    gp = engine.grid2d_params(3, 3, 0, 1.0, 0.5, 50.0, 0.5)
    out = engine.grid2d_rhs(gp, g["g3x3t0_degenerate_state"].reshape(-1, 4).copy(), fp_mode=FP_STRICT)
    assert np.array_equal(out.ravel(), g["g3x3t0_degenerate_rhs"], equal_nan=True)


@pytest.mark.parametrize("rows,cols,topo", [(64, 96, 0), (33, 257, 0), (48, 40, 1), (40, 48, 2), (513, 70, 0)])
def test_grid2d_medium_vs_oracle_bit_exact(engine, port, rows, cols, topo):
    p = W.GRID2D_PARAMS
    s0 = W.grid2d_state(rows, cols, p["spacing"])
    gp = engine.grid2d_params(rows, cols, topo, p["mass"], p["spacing"], p["k"], p["c"])
    ref_rhs = port.grid2d_rhs(rows, cols, topo, p["mass"], p["spacing"], p["k"], p["c"], s0)
    assert np.array_equal(engine.grid2d_rhs(gp, s0.copy(), fp_mode=FP_STRICT).ravel(), ref_rhs)
    ref = port.grid2d_rk4(rows, cols, topo, p["mass"], p["spacing"], p["k"], p["c"], p["dt"], 25, s0, nthreads=4)
    for kernel in ((None, "march", "tile") if topo == 0 else (None,)):
        got = engine.grid2d_rk4(gp, p["dt"], 25, s0.copy(), fp_mode=FP_STRICT, kernel=kernel)
        assert np.array_equal(got.ravel(), ref), kernel
        dev = engine.to_device(s0)
        engine.grid2d_rk4(gp, p["dt"], 25, dev, fp_mode=FP_STRICT, kernel=kernel)
        assert np.array_equal(dev.download().ravel(), ref), kernel
        fast = engine.grid2d_rk4(gp, p["dt"], 25, s0.copy(), fp_mode=FP_CONTRACT, kernel=kernel)
        assert rel_err(fast.ravel(), ref).max() <= FINAL, kernel


def test_grid2d_large_properties(engine, port):
    """2048^2 (too big for the oracle to integrate in seconds): size-independent properties --
    rest grid is a fixed point; total momentum stays at zero (internal forces cancel,
    tests/verify_grid_physics.cpp:116); top-left 64x64 block after 3 steps equals the oracle run on a
    512x512 crop (influence travels <= 4 rows per step, so the crop's far boundary cannot be felt)."""
    p = W.GRID2D_PARAMS
    n = 2048
    gp = engine.grid2d_params(n, n, 0, p["mass"], p["spacing"], p["k"], p["c"])
    rest = engine.grid2d_initial_state(gp, mem=MEM_DEVICE)
    engine.grid2d_rk4(gp, p["dt"], 5, rest, fp_mode=FP_STRICT)
    r = rest.download()
    assert np.array_equal(r, W.grid2d_state(n, n, p["spacing"], bump=False))
    s0 = W.grid2d_state(n, n, p["spacing"])
    dev = engine.to_device(s0)
    engine.grid2d_rk4(gp, p["dt"], 3, dev, fp_mode=FP_STRICT)
    out = dev.download().reshape(n, n, 4)
    assert np.abs(out[..., 2:].sum(axis=(0, 1))).max() < 1e-9
    crop = np.ascontiguousarray(s0.reshape(n, n, 4)[:512, :512]).reshape(-1, 4)
    ref = port.grid2d_rk4(512, 512, 0, p["mass"], p["spacing"], p["k"], p["c"], p["dt"], 3, crop, nthreads=8).reshape(512, 512, 4)
    assert np.array_equal(out[:64, :64], ref[:64, :64])
    # the size-based choice (marching kernel at this size), the forced tile kernel and the per-stage schedule: the same bits everywhere
    for kw in (dict(kernel="tile"), dict(kernel="march"), dict(per_stage=True)):
        dev2 = engine.to_device(s0)
        engine.grid2d_rk4(gp, p["dt"], 3, dev2, fp_mode=FP_STRICT, **kw)
        assert np.array_equal(dev2.download().reshape(n, n, 4), out), kw
        dev2.free()


@pytest.mark.parametrize("rows,cols,topo,steps", [(37, 45, 0, 7), (64, 64, 1, 4), (19, 130, 2, 5), (130, 33, 0, 6)])
def test_grid2d_fused_step_equals_per_stage_schedule(engine, port, rows, cols, topo, steps):
    """The fused RK4-step kernel (all four stages on a shared-memory tile with a 4-cell halo, the default) and the
    four-kernel per-stage schedule produce the same bits in strict mode (== the oracle) and agree to rounding in
    contract mode.  Ragged
    sizes exercise partial tiles on both axes, odd step counts the ping-pong copy-back."""
    p = W.GRID2D_PARAMS
    s0 = W.grid2d_state(rows, cols, p["spacing"])
    gp = engine.grid2d_params(rows, cols, topo, p["mass"], p["spacing"], p["k"], p["c"])
    staged = engine.grid2d_rk4(gp, p["dt"], steps, s0.copy(), fp_mode=FP_STRICT, per_stage=True)
    staged_c = engine.grid2d_rk4(gp, p["dt"], steps, s0.copy(), fp_mode=FP_CONTRACT, per_stage=True)
    for kernel in ((None, "march", "tile") if topo == 0 else (None,)):
        fused = engine.grid2d_rk4(gp, p["dt"], steps, s0.copy(), fp_mode=FP_STRICT, kernel=kernel)
        assert np.array_equal(fused, staged), kernel
        # contract mode: the kernels contract different expressions, so they agree to rounding, not to the bit
        fused_c = engine.grid2d_rk4(gp, p["dt"], steps, s0.copy(), fp_mode=FP_CONTRACT, kernel=kernel)
        assert rel_err(fused_c.ravel(), fused.ravel()).max() <= FINAL and rel_err(staged_c.ravel(), fused.ravel()).max() <= FINAL
    ref = port.grid2d_rk4(rows, cols, topo, p["mass"], p["spacing"], p["k"], p["c"], p["dt"], steps, s0, nthreads=4)
    assert np.array_equal(engine.grid2d_rk4(gp, p["dt"], steps, s0.copy(), fp_mode=FP_STRICT).ravel(), ref)


@pytest.mark.parametrize("spacing", [1e-160, 1e-12, 1e100])
def test_grid2d_strict_guarded_route_tiny_and_huge_operands(engine, port, spacing):
    """The strict marching kernel computes a stage with unguarded inline root / quotient sequences and falls back to the guarded
    functions when any operand of the warp leaves their range.  Grids on absurd length scales drive every stage down the
    guarded route (spacing 1e-160: numerators below 2^-500; 1e-12: springs shorter than the 2e-20 threshold on l2, i.e. the
    len_safe region; 1e100: l2 above 2^500): bit-identical to the per-stage schedule and to the oracle there too."""
    rows, cols, steps = 70, 100, 3
    s0 = W.grid2d_state(rows, cols, 0.5) * spacing / 0.5
    gp = engine.grid2d_params(rows, cols, 0, 1.0, spacing, 50.0, 0.5)
    fused = engine.grid2d_rk4(gp, 1e-3, steps, s0.copy(), fp_mode=FP_STRICT, kernel="march")
    staged = engine.grid2d_rk4(gp, 1e-3, steps, s0.copy(), fp_mode=FP_STRICT, per_stage=True)
    assert np.array_equal(fused, staged)
    assert np.array_equal(engine.grid2d_rk4(gp, 1e-3, steps, s0.copy(), fp_mode=FP_STRICT, kernel="tile"), staged)
    ref = port.grid2d_rk4(rows, cols, 0, 1.0, spacing, 50.0, 0.5, 1e-3, steps, s0, nthreads=4)
    assert np.array_equal(fused.ravel(), ref) and np.all(np.isfinite(ref))


def test_inline_ieee_division_and_sqrt_are_correctly_rounded(tmp_path):
    """csrc/ieee.cuh (the branch-light division / sqrt every strict kernel uses) against nvcc's own IEEE operators,
    bit for bit on 6e8 samples on the device: random mantissas and exponents, exact zeros, quotients next to 1,
    subnormals / inf / nan through the fallback."""
    import os
    import subprocess
    from conftest import ROOT
    exe = tmp_path / "exact_math_check"
    src = os.path.join(ROOT, "tests", "cuda", "exact_math_check.cu")
    subprocess.run(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe), src], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatches" in r.stdout


# ---------------------------------------------------------------- section 8f-1: LQR sweep and closed loop
def test_lqr_golden_bit_exact_strict(engine):
    """Batched LQR<4,1>::solve: strict mode == the reference's class bit for bit (gains, Riccati solution, status),
    including the 'singular' throw and the diverging iteration; contract mode within 1e-9 of it."""
    g = golden("control")
    P = g["A"].shape[1]
    for tag, r, dt, mi, tol in (("a", 0.01, 0.01, 1000, 1e-8), ("b", 0.5, 0.002, 300, 1e-10)):
        K, Pr, it, st = engine.lqr4_gain(g["A"], g["B"], P, g["Q"], r, dt, mi, tol, fp_mode=FP_STRICT, want_riccati=True)
        assert np.array_equal(K, g["K_" + tag]) and np.array_equal(Pr, g["P_" + tag]) and np.array_equal(st, g["st_" + tag])
        Kc, _, _, stc = engine.lqr4_gain(g["A"], g["B"], P, g["Q"], r, dt, mi, tol, fp_mode=FP_CONTRACT)
        assert not stc.any() and rel_err(Kc, g["K_" + tag]).max() <= 1e-6     # gain of an iteration stopped at tol 1e-8
    _, _, _, st = engine.lqr4_gain(g["A2"], g["B2"], 2, np.eye(4), 0.0, fp_mode=FP_STRICT)
    assert list(st) == [2, 2]
    K, _, _, st = engine.lqr4_gain(g["A2"], g["B2"], 2, np.eye(4), 1.0, fp_mode=FP_STRICT)
    assert np.array_equal(st, g["st_div"]) and np.array_equal(K, g["K_div"], equal_nan=True)


def test_closed_loop_golden_bit_exact_strict(engine):
    g = golden("control")
    n = g["cl_s0"].shape[1]
    pl = [g[k][:n] for k in ("mc", "m", "L", "g")]
    s0 = g["cl_s0"]
    fin, stats, _, st = engine.cartpole_feedback_rk4(*pl, s0[0], s0[1], s0[2], s0[3], g["cl_K"], n, 0.0, 0.002, 1500,
                                                     x_ref=g["cl_xref"], u_min=-15.0, u_max=15.0, saturate=True, fp_mode=FP_STRICT)
    assert not st.any() and np.abs(fin - g["cl_final"]).max() <= 1e-12      # sin/cos: libm-level differences only
    assert np.abs(stats - g["cl_stats"]).max() <= 1e-12
    fin, stats, _, _ = engine.cartpole_feedback_rk4(*pl, s0[0], s0[1], s0[2], s0[3], g["cl_K"], n, 0.0, 0.002, 700, fp_mode=FP_CONTRACT)
    assert rel_err(fin, g["cl_final_ns"]).max() <= FINAL and rel_err(stats, g["cl_stats_ns"]).max() <= FINAL


def test_linearize_lqr_closed_loop_pipeline_on_device(engine, port):
    """The whole sweep without leaving the GPU: linearize (K2) -> LQR (K5) -> closed-loop ensemble (K1), device
    pointers throughout, 20 000 plants; a sample is checked against the oracle and every loop must stabilise."""
    P = 20000
    rng = np.random.default_rng(11)
    mc, m, L = rng.uniform(0.8, 1.5, P), rng.uniform(0.05, 0.2, P), rng.uniform(0.4, 0.8, P)
    g = np.full(P, 9.81)
    x0 = np.zeros((4, P)); u0 = np.zeros(P)
    d = {k: engine.to_device(v) for k, v in dict(mc=mc, m=m, L=L, g=g).items()}
    dx, du = engine.to_device(x0), engine.to_device(u0)
    A, B, _ = engine.cartpole_linearize(d["mc"], d["m"], d["L"], d["g"], dx, du, P, mem=MEM_DEVICE, fp_mode=FP_STRICT)
    Q = np.diag([10.0, 100.0, 1.0, 10.0])
    K, _, it, st = engine.lqr4_gain(A, B, P, Q, 0.01, mem=MEM_DEVICE, fp_mode=FP_STRICT)
    assert not st.download().any()
    s0 = np.zeros((4, P)); s0[1] = rng.uniform(-0.1, 0.1, P); s0[0] = rng.uniform(-0.5, 0.5, P)
    ds = [engine.to_device(np.ascontiguousarray(s0[j])) for j in range(4)]
    fin, stats, _, st2 = engine.cartpole_feedback_rk4(d["mc"], d["m"], d["L"], d["g"], *ds, K, P, 0.0, 0.002, 2500,
                                                      u_min=-50.0, u_max=50.0, saturate=True, mem=MEM_DEVICE, fp_mode=FP_STRICT)
    fin, stats = fin.download(), stats.download()
    assert not st2.download().any()
    assert np.abs(fin[1]).max() < 0.01 and stats[0].max() < 0.2 and np.abs(fin[0]).max() < 0.05
    idx = np.arange(0, P, 997)
    rA, rB = port.cartpole_linearize(x0[:, idx], u0[idx], mc[idx], m[idx], L[idx], g[idx])
    assert np.abs(A.download()[:, idx] - rA).max() <= 1e-12
    rK, _, rit, rst = port.lqr4(A.download()[:, idx], B.download()[:, idx], Q, 0.01)
    assert np.array_equal(K.download()[:, idx], rK) and np.array_equal(it.download()[idx], rit)
    rfin, rstats = port.cartpole_feedback_rk4(mc[idx], m[idx], L[idx], g[idx], s0[:, idx], rK, None, -50.0, 50.0, 1, 0.0, 0.002, 2500, nthreads=4)
    assert np.abs(fin[:, idx] - rfin).max() <= 1e-9 and np.abs(stats[:, idx] - rstats).max() <= 1e-9


def test_rocket_aero_tables_golden(engine):
    """Rocket with bilinear Cd/Cl/Cs and damping tables in shared memory (RocketSysT<true>): trajectories, statistics and
    one derivative evaluation of a tumbling rocket against the reference's components (golden vectors)."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    aero = mg.rocket_aero_tables()
    g = golden("rocket_aero")
    n = g["elev"].size
    for mode in (FP_STRICT, FP_CONTRACT):
        for tag, on in (("on", True), ("off", False)):
            st, stats, _, status = engine.rocket_rk4(g["elev"], g["az"], g["diam"], g["g0"], g["engine"], g["mass_tab"], None, None, n,
                                                     0.01, 1000, fp_mode=mode, aero=engine.aero_tables(aero, engine_on=on))
            assert not status.any()
            assert rel_err(st, g["final_" + tag]).max() <= FINAL
            assert rel_err(stats[[0, 2, 4, 5]], g["stats_" + tag][[0, 2, 4, 5]]).max() <= FINAL
        rhs = engine.rocket_rhs(84.0, 30.0, 0.16, 9.80665, g["engine"], g["mass_tab"], g["rhs_state"], fp_mode=mode, aero=aero)
        assert rel_err(rhs, g["rhs"]).max() <= 1e-12


@pytest.mark.parametrize("n_links", [2, 6])
def test_lqr_block_kernel_golden_bit_exact_strict(engine, n_links):
    """LQR<6,1> / LQR<14,1> (one CTA per operating point, matrices in shared memory): strict mode == the reference class
    bit for bit on the 2- and 6-link upright linearizations (the reference's own 6-link test configuration)."""
    g = golden("control_n")
    k, ns = f"n{n_links}_", 2 * (n_links + 1)
    P = g[k + "A"].shape[1]
    K, Pr, it, st = engine.lqr_gain(ns, g[k + "A"], g[k + "B"], P, g[k + "Q"], 0.001, 0.005, 1000, 1e-8, fp_mode=FP_STRICT, want_riccati=True)
    assert np.array_equal(st, g[k + "st"]) and np.array_equal(K, g[k + "K"]) and np.array_equal(Pr, g[k + "P"])
    Kc, _, _, stc = engine.lqr_gain(ns, g[k + "A"], g[k + "B"], P, g[k + "Q"], 0.001, 0.005, 1000, 1e-8, fp_mode=FP_CONTRACT)
    assert not stc.any() and rel_err(Kc, g[k + "K"]).max() <= 1e-6
    # state_dim 4 routes to the thread-per-point kernel
    c = golden("control")
    K4, _, _, st4 = engine.lqr_gain(4, c["A"], c["B"], c["A"].shape[1], c["Q"], 0.01, 0.01, 1000, 1e-8, fp_mode=FP_STRICT)
    assert np.array_equal(K4, c["K_a"])


# ---------------------------------------------------------------- section 8f-2: unified 6-state grid
def _ugrid_prm():
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    return mg.UGRID_PARAMS


UGRID_KERNELS = (None, "march", "gather")


@pytest.mark.parametrize("rows,cols", [(3, 3), (4, 5), (2, 7), (6, 6), (10, 10)])
def test_unified_grid_golden(engine, rows, cols):
    """Unified 6-state grid (PointMass2D + Spring2D with attachment points and torque) against the reference's ValidatedSystem:
    initial state exact; derivative and 50 RK4 steps to 1e-12 / 1e-9 (sin/cos of the attachment angles: libm-level)."""
    g = golden("ugrid")
    for tag, prm in _ugrid_prm().items():
        key = f"u{rows}x{cols}_{tag}"
        gp = engine.ugrid_params(rows, cols, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], prm[7], tuple(prm[8:12]))
        assert np.array_equal(engine.ugrid_initial_state(gp), g[key + "_init"])
        for mode in (FP_STRICT, FP_CONTRACT):
            d = engine.ugrid_rhs(gp, g[key + "_state"].copy(), fp_mode=mode)
            assert np.abs(d - g[key + "_rhs"]).max() <= 1e-12 * np.abs(g[key + "_rhs"]).max()
            for kernel in UGRID_KERNELS:        # size-based choice, marching stage kernel forced, gather kernel forced
                out = engine.ugrid_rk4(gp, 1e-3, 50, g[key + "_state"].copy(), fp_mode=mode, kernel=kernel)
                assert rel_err(out.ravel(), g[key + "_rk4"].ravel()).max() <= FINAL, kernel


def test_unified_grid_integrators(engine, port):
    """Grid2DSimulator's other two integrators (wasm/wasm_grid2d.cpp:329-413): symplectic Euler and velocity Verlet on the
    device against the golden vectors (80 steps) and, on a size the reference is not instantiated for, against the oracle.
    Odd and even step counts exercise both ends of the ping-pong between the state and the scratch buffer."""
    from sopot_b200.capi import INTEGRATOR_SYMPLECTIC_EULER, INTEGRATOR_VELOCITY_VERLET
    g, gi = golden("ugrid"), golden("ugrid_integrators")
    prm = _ugrid_prm()["wasm"]
    ids = {"symplectic_euler": INTEGRATOR_SYMPLECTIC_EULER, "velocity_verlet": INTEGRATOR_VELOCITY_VERLET}
    for rows, cols in ((4, 5), (10, 10)):
        gp = engine.ugrid_params(rows, cols, *prm[:8], tuple(prm[8:12]))
        for name, integ in ids.items():
            for mode in (FP_STRICT, FP_CONTRACT):
                for kernel in UGRID_KERNELS:
                    out = engine.ugrid_rk4(gp, 1e-3, 80, g[f"u{rows}x{cols}_wasm_state"].copy(), fp_mode=mode, integrator=integ, kernel=kernel)
                    assert rel_err(out.ravel(), gi[f"u{rows}x{cols}_{name}"].ravel()).max() <= FINAL, kernel
    rows, cols = 45, 77
    gp = engine.ugrid_params(rows, cols, *prm[:8], tuple(prm[8:12]))
    s0 = engine.ugrid_initial_state(gp)
    rng = np.random.default_rng(5)
    st = s0 + rng.normal(0, 0.02, s0.shape); st[:, 4] = rng.normal(0, 0.2, rows * cols)
    for name, integ in ids.items():
        for n_steps in (1, 7, 12):
            ref = port.ugrid(rows, cols, prm, st, 2e-3, n_steps, name)
            strict_bits = None
            for kernel in UGRID_KERNELS:
                got = engine.ugrid_rk4(gp, 2e-3, n_steps, st.copy(), fp_mode=FP_STRICT, integrator=integ, kernel=kernel)
                assert rel_err(got.ravel(), ref.ravel()).max() <= FINAL, kernel
                assert strict_bits is None or np.array_equal(got, strict_bits), kernel      # strict: both kernels, the same bits
                strict_bits = got
                dev = engine.to_device(st)
                engine.ugrid_rk4(gp, 2e-3, n_steps, dev, fp_mode=FP_CONTRACT, integrator=integ, kernel=kernel)
                assert rel_err(dev.download().ravel(), ref.ravel()).max() <= FINAL, kernel
    with pytest.raises(Exception):
        engine.ugrid_rk4(gp, 1e-3, 1, st.copy(), integrator=9)


@pytest.mark.parametrize("rows,cols", [(2, 2), (9, 31), (8, 32), (17, 65), (70, 100), (129, 64)])
def test_unified_grid_stage_pairs_equal_per_stage_schedule(engine, port, rows, cols):
    """RK4 with two stages per kernel (shared-memory tile + 2-cell halo, opt-in) against the default per-stage
    schedule: bit-identical in strict mode, within tolerance in contract mode (FMA contraction may differ), and both
    against the oracle; tile-edge sizes (one short of / exactly / one past a 32 x 8 tile), odd and even step counts."""
    prm = _ugrid_prm()["repel" if rows % 2 else "wasm"]
    gp = engine.ugrid_params(rows, cols, *prm[:8], tuple(prm[8:12]))
    s0 = engine.ugrid_initial_state(gp)
    rng = np.random.default_rng(rows * 7 + cols)
    st = s0 + rng.normal(0, 0.02, s0.shape); st[:, 4] = rng.normal(0, 0.3, rows * cols); st[:, 5] = rng.normal(0, 0.3, rows * cols)
    for steps in (1, 4):
        ref = port.ugrid(rows, cols, prm, st, 1e-3, steps, "rk4")
        a = engine.ugrid_rk4(gp, 1e-3, steps, st.copy(), fp_mode=FP_STRICT, stage_pairs=True)
        b = engine.ugrid_rk4(gp, 1e-3, steps, st.copy(), fp_mode=FP_STRICT)
        assert np.array_equal(a, b)
        assert rel_err(a.ravel(), ref.ravel()).max() <= FINAL
        c = engine.ugrid_rk4(gp, 1e-3, steps, st.copy(), fp_mode=FP_CONTRACT, stage_pairs=True)
        d = engine.ugrid_rk4(gp, 1e-3, steps, st.copy(), fp_mode=FP_CONTRACT)
        assert rel_err(c.ravel(), ref.ravel()).max() <= FINAL and rel_err(d.ravel(), ref.ravel()).max() <= FINAL
        dev = engine.to_device(st)
        engine.ugrid_rk4(gp, 1e-3, steps, dev, fp_mode=FP_STRICT, stage_pairs=True)
        assert np.array_equal(dev.download(), a)


def test_unified_grid_triangle_topology(engine, port):
    """The "triangle" grid type: golden vectors of the reference (derivative, RK4, velocity Verlet), a size the reference is
    not instantiated for against the oracle, the opt-in stage-pair schedule (bit-identical in strict mode) and a rest state
    whose diagonal rest length equals the diagonal attachment distance."""
    from sopot_b200.capi import INTEGRATOR_VELOCITY_VERLET
    g, gt = golden("ugrid"), golden("ugrid_triangle")
    prm = _ugrid_prm()["wasm"]
    tri = list(prm) + [1.0, float(np.sqrt(2.0) * 0.5), np.pi / 4, -3 * np.pi / 4, 3 * np.pi / 4, -np.pi / 4]
    for rows, cols in ((3, 3), (4, 5), (2, 7), (6, 6)):
        gp = engine.ugrid_params(rows, cols, *prm[:8], tuple(prm[8:12]), triangle=True)
        st = g[f"u{rows}x{cols}_wasm_state"]
        key = f"x{rows}x{cols}"
        for mode in (FP_STRICT, FP_CONTRACT):
            d = engine.ugrid_rhs(gp, st.copy(), fp_mode=mode)
            assert np.abs(d - gt[key + "_rhs"]).max() <= 1e-12 * np.abs(gt[key + "_rhs"]).max()
            out = engine.ugrid_rk4(gp, 1e-3, 50, st.copy(), fp_mode=mode)
            assert rel_err(out.ravel(), gt[key + "_rk4"].ravel()).max() <= FINAL
            out = engine.ugrid_rk4(gp, 1e-3, 50, st.copy(), fp_mode=mode, integrator=INTEGRATOR_VELOCITY_VERLET)
            assert rel_err(out.ravel(), gt[key + "_verlet"].ravel()).max() <= FINAL
    rows, cols = 41, 67
    gp = engine.ugrid_params(rows, cols, *prm[:8], tuple(prm[8:12]), triangle=True)
    s0 = engine.ugrid_initial_state(gp)
    rng = np.random.default_rng(9)
    st = s0 + rng.normal(0, 0.02, s0.shape); st[:, 4] = rng.normal(0, 0.2, rows * cols)
    ref = port.ugrid(rows, cols, tri, st, 1e-3, 8, "rk4")
    a = engine.ugrid_rk4(gp, 1e-3, 8, st.copy(), fp_mode=FP_STRICT)
    b = engine.ugrid_rk4(gp, 1e-3, 8, st.copy(), fp_mode=FP_STRICT, stage_pairs=True)
    c = engine.ugrid_rk4(gp, 1e-3, 8, st.copy(), fp_mode=FP_CONTRACT)
    assert np.array_equal(a, b) and rel_err(a.ravel(), ref.ravel()).max() <= FINAL and rel_err(c.ravel(), ref.ravel()).max() <= FINAL
    # at rest: orthogonal rest length = spacing - 2 r, diagonal rest length = sqrt(2) spacing - 2 r
    gp = engine.ugrid_params(48, 48, prm[0], 0.05, 0.5, prm[3], 0.4, prm[5], 0.0, -1.0, tuple(prm[8:12]), triangle=True,
                             diag_rest_length=float(np.sqrt(2.0) * 0.5 - 0.1))
    s0 = engine.ugrid_initial_state(gp)
    out = engine.ugrid_rk4(gp, 1e-3, 5, s0.copy(), fp_mode=FP_STRICT)
    assert np.abs(out - s0).max() < 1e-12
    with pytest.raises(Exception):
        bad = engine.ugrid_params(4, 4, *prm[:8], tuple(prm[8:12])); bad.topology = 7
        engine.ugrid_rk4(bad, 1e-3, 1, engine.ugrid_initial_state(engine.ugrid_params(4, 4, *prm[:8], tuple(prm[8:12]))))


def test_unified_grid_medium_vs_oracle(engine, port):
    """Sizes the templated reference is not instantiated for: against the runtime-sized oracle; device-pointer path; a rest
    grid whose rest length equals the attachment distance stays at rest exactly."""
    prm = _ugrid_prm()["wasm"]
    for rows, cols in ((33, 70), (129, 64)):
        gp = engine.ugrid_params(rows, cols, *prm[:8], tuple(prm[8:12]))
        s0 = engine.ugrid_initial_state(gp)
        rng = np.random.default_rng(rows)
        st = s0 + rng.normal(0, 0.02, s0.shape); st[:, 4] = rng.normal(0, 0.2, rows * cols)
        ref = port.ugrid(rows, cols, prm, st, 1e-3, 10, "rk4")
        strict_bits = None
        for kernel in UGRID_KERNELS:
            got = engine.ugrid_rk4(gp, 1e-3, 10, st.copy(), fp_mode=FP_STRICT, kernel=kernel)
            assert rel_err(got.ravel(), ref.ravel()).max() <= FINAL, kernel
            assert strict_bits is None or np.array_equal(got, strict_bits), kernel
            strict_bits = got
            dev = engine.to_device(st)
            engine.ugrid_rk4(gp, 1e-3, 10, dev, fp_mode=FP_CONTRACT, kernel=kernel)
            assert rel_err(dev.download().ravel(), ref.ravel()).max() <= FINAL, kernel
    rest = list(prm); rest[4] = 0.4                                   # 0.5 spacing - 2 * 0.05 radius
    gp = engine.ugrid_params(64, 64, *rest[:8], tuple(rest[8:12]))
    s0 = engine.ugrid_initial_state(gp)
    out = engine.ugrid_rk4(gp, 1e-3, 5, s0.copy(), fp_mode=FP_STRICT)
    assert np.abs(out - s0).max() < 1e-13
