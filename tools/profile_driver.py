"""Small, short workloads for ncu captures (one config per call):  python tools/profile_driver.py <config>"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sopot_b200 import workloads as W  # noqa: E402
from sopot_b200.capi import FP_CONTRACT, FP_STRICT, MEM_DEVICE, Engine  # noqa: E402

cfg = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = Engine(0)
if cfg == "dpend":
    n = 1 << 20
    w = W.dpend_ensemble(n)
    keys = ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")
    d = [eng.to_device(w[k]) for k in keys]
    st = eng.empty((4, n))
    for _ in range(reps):
        eng.dpend_rk4(*d, n, 0.0, 1e-3, 1000, mem=MEM_DEVICE, fp_mode=FP_CONTRACT, want_status=False, state_out=st)
elif cfg == "ho":
    n = 1 << 20
    w = W.ho_ensemble(n)
    d = [eng.to_device(w[k]) for k in ("m", "k", "c", "x0", "v0")]
    st = eng.empty((2, n))
    for _ in range(reps):
        eng.ho_rk4(*d, n, 0.0, 0.01, 1000, mem=MEM_DEVICE, fp_mode=FP_CONTRACT, want_status=False, state_out=st)
elif cfg == "lin":
    P = 1 << 20
    w = W.cartpole_points(P)
    dx, du = eng.to_device(w["x"]), eng.to_device(w["u"])
    A, B = eng.empty((16, P)), eng.empty((4, P))
    for _ in range(reps):
        eng.cartpole_linearize(1.0, 0.1, 0.5, 9.81, dx, du, P, mem=MEM_DEVICE, A=A, B=B, want_status=False)
elif cfg == "rocket":
    n = 1 << 18
    w = W.rocket_ensemble(n)
    d = [eng.to_device(w[k]) for k in ("elev", "az", "diam", "g0")]
    st, stats = eng.empty((14, n)), eng.empty((8, n))
    for _ in range(reps):
        eng.rocket_rk4(*d, w["engine"], w["mass_tab"], None, None, n, 0.01, 600, mem=MEM_DEVICE, want_status=False,
                       state_out=st, stats_out=stats)
elif cfg == "rocket_1mi":                        # the bench ensemble itself: 1 Mi trajectories x 3000 steps
    n = 1 << 20
    w = W.rocket_ensemble(n)
    d = [eng.to_device(w[k]) for k in ("elev", "az", "diam", "g0")]
    st, stats = eng.empty((14, n)), eng.empty((8, n))
    for _ in range(reps):
        eng.rocket_rk4(*d, w["engine"], w["mass_tab"], None, None, n, 0.01, 3000, mem=MEM_DEVICE, want_status=False,
                       state_out=st, stats_out=stats)
elif cfg == "rocket_full":                       # the whole 30 s flight (88 % of it is coast), a quarter of the bench ensemble
    n = 1 << 18
    w = W.rocket_ensemble(n)
    d = [eng.to_device(w[k]) for k in ("elev", "az", "diam", "g0")]
    st, stats = eng.empty((14, n)), eng.empty((8, n))
    for _ in range(reps):
        eng.rocket_rk4(*d, w["engine"], w["mass_tab"], None, None, n, 0.01, 3000, mem=MEM_DEVICE, want_status=False,
                       state_out=st, stats_out=stats)
elif cfg in ("ugrid", "ugrid_pairs"):
    n = 4096
    ug = eng.ugrid_params(n, n, 1.0, 0.05, 0.5, 50.0, 0.5, 0.5)
    state = eng.ugrid_initial_state(ug, mem=MEM_DEVICE)
    for _ in range(reps):
        eng.ugrid_rk4(ug, 1e-3, 2, state, fp_mode=FP_CONTRACT, stage_pairs=cfg == "ugrid_pairs")
elif cfg == "closed":
    P = 1 << 20
    rng = np.random.default_rng(3)
    plant = {k: eng.to_device(v) for k, v in dict(mc=rng.uniform(0.8, 1.5, P), m=rng.uniform(0.05, 0.2, P),
                                                   L=rng.uniform(0.4, 0.8, P), g=np.full(P, 9.81)).items()}
    K = eng.to_device(np.tile(np.array([-3.0, 40.0, -4.0, 8.0])[:, None], (1, P)))
    s0 = [eng.to_device(v) for v in (rng.uniform(-0.5, 0.5, P), rng.uniform(-0.1, 0.1, P), np.zeros(P), np.zeros(P))]
    st = eng.empty((4, P))
    for _ in range(reps):
        eng.cartpole_feedback_rk4(plant["mc"], plant["m"], plant["L"], plant["g"], *s0, K, P, 0.0, 0.002, 500, u_min=-50.0, u_max=50.0,
                                  saturate=True, mem=MEM_DEVICE, want_status=False, want_stats=False, state_out=st)
elif cfg == "lqr":
    P = 1 << 14
    rng = np.random.default_rng(3)
    A, B, _ = eng.cartpole_linearize(eng.to_device(rng.uniform(0.8, 1.5, P)), eng.to_device(rng.uniform(0.05, 0.2, P)),
                                     eng.to_device(rng.uniform(0.4, 0.8, P)), eng.to_device(np.full(P, 9.81)),
                                     eng.to_device(np.zeros((4, P))), eng.to_device(np.zeros(P)), P, mem=MEM_DEVICE, want_status=False)
    for _ in range(reps):
        eng.lqr4_gain(A, B, P, np.diag([10.0, 100.0, 1.0, 10.0]), 0.01, mem=MEM_DEVICE, fp_mode=FP_CONTRACT)
elif cfg in ("grid", "grid_strict"):
    p = W.GRID2D_PARAMS
    n = 8192
    gp = eng.grid2d_params(n, n, 0, p["mass"], p["spacing"], p["k"], p["c"])
    state = eng.grid2d_initial_state(gp, mem=MEM_DEVICE)
    eng.grid2d_rk4(gp, p["dt"], reps, state, fp_mode=FP_STRICT if cfg == "grid_strict" else FP_CONTRACT,
                   per_stage=len(sys.argv) > 3 and sys.argv[3] == "per_stage")
else:
    raise SystemExit("unknown config " + cfg)
eng.sync()
print("done", cfg, eng.launch_count())
eng.close()
