"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck), ragged sizes on purpose:
    compute-sanitizer --tool memcheck python tools/sanitize_driver.py
Results are compared with nothing here (parity is tests/); the point is that the tool sees every kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sopot_b200 import workloads as W  # noqa: E402
from sopot_b200.capi import FP_CONTRACT, FP_STRICT, Engine  # noqa: E402

only = sys.argv[1] if len(sys.argv) > 1 else "all"
eng = Engine(0)
if only in ("all", "ensemble"):
    for mode in (FP_CONTRACT, FP_STRICT):
        n = 1000 + 37
        w = W.ho_ensemble(n)
        eng.ho_rk4(w["m"], w["k"], w["c"], w["x0"], w["v0"], n, 0.0, 0.01, 50, fp_mode=mode, store_every=10)
        w = W.dpend_ensemble(n)
        eng.dpend_rk4(*[w[k] for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")], n, 0.0, 1e-3, 40, fp_mode=mode, store_every=7)
        w = W.cartpole_points(n)
        A, B, _ = eng.cartpole_linearize(w["mc"], w["m"], w["L"], w["g"], w["x"], w["u"], n, fp_mode=mode)
        eng.cartpole_rk4(1.0, 0.1, 0.5, 9.81, w["x"][0], 0.1 * w["x"][1], w["x"][2], w["x"][3], w["u"], n, 0.0, 0.01, 30, fp_mode=mode)
        x0 = np.zeros((4, 300)); u0 = np.zeros(300)
        A, B, _ = eng.cartpole_linearize(1.0, np.linspace(0.05, 0.2, 300), 0.5, 9.81, x0, u0, 300, fp_mode=mode)
        K, _, _, st = eng.lqr4_gain(A, B, 300, np.diag([10.0, 100.0, 1.0, 10.0]), 0.01, max_iter=200, fp_mode=mode, want_riccati=True)
        s0 = np.zeros((4, 300)); s0[1] = 0.05
        eng.cartpole_feedback_rk4(1.0, np.linspace(0.05, 0.2, 300), 0.5, 9.81, s0[0], s0[1], s0[2], s0[3], np.nan_to_num(K), 300, 0.0, 0.002, 100,
                                  u_min=-20.0, u_max=20.0, saturate=True, fp_mode=mode)
        w = W.rocket_ensemble(200)
        _, stats, _, _ = eng.rocket_rk4(w["elev"], w["az"], w["diam"], w["g0"], w["engine"], w["mass_tab"], None, None, 200, 0.01, 450, fp_mode=mode,
                                        store_every=100)
        eng.dispersion_stats(stats)
        mach = np.array([0.0, 0.5, 1.0]); ang = np.array([0.0, 5.0, 20.0])
        aero = dict(cd_power_on=(mach, ang, 0.3 + 0.1 * mach[:, None] + 0.0 * ang[None, :]), cl_power_off=(mach, ang, 0.01 * ang[None, :] + 0.0 * mach[:, None]),
                    cmq_yz=(np.array([-20.0, 0.0, 20.0]), mach, -400.0 + 0.0 * mach[None, :] + 0.0 * ang[:, None]))
        eng.rocket_rk4(w["elev"], w["az"], w["diam"], w["g0"], w["engine"], w["mass_tab"], None, None, 200, 0.01, 450, fp_mode=mode, aero=aero)
if only in ("all", "grid"):
    p = W.GRID2D_PARAMS
    for rows, cols, topo in ((37, 45, 0), (19, 70, 1), (33, 34, 2), (2, 2, 0)):
        gp = eng.grid2d_params(rows, cols, topo, p["mass"], p["spacing"], p["k"], p["c"])
        s0 = W.grid2d_state(rows, cols, p["spacing"])
        for mode in (FP_CONTRACT, FP_STRICT):
            for per_stage in (False, True):
                eng.grid2d_rk4(gp, p["dt"], 3, s0.copy(), fp_mode=mode, per_stage=per_stage)
            eng.grid2d_rhs(gp, s0.copy(), fp_mode=mode)
eng.sync()
print("sanitize_driver done:", only, eng.launch_count(), "launches")
eng.close()
