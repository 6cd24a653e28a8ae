"""Time one config's kernel, device-resident, back-to-back launches between CUDA events:
   python tools/time_cfg.py rocket|ho|lin|grid|grid_strict|ugrid [n] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sopot_b200 import workloads as W
from sopot_b200.capi import FP_CONTRACT, FP_STRICT, MEM_DEVICE, Engine
cfg = sys.argv[1]
eng = Engine(0)
def timed(fn, reps):
    fn(); eng.sync()
    eng.event_record(0)
    for _ in range(reps): fn()
    eng.event_record(1)
    return eng.elapsed_ms(0, 1) / reps
if cfg == "rocket":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
    w = W.rocket_ensemble(n)
    d = [eng.to_device(w[k]) for k in ("elev", "az", "diam", "g0")]
    st, stats = eng.empty((14, n)), eng.empty((8, n))
    ms = timed(lambda: eng.rocket_rk4(*d, w["engine"], w["mass_tab"], None, None, n, 0.01, steps, mem=MEM_DEVICE,
                                      want_status=False, state_out=st, stats_out=stats, sync=False), 2)
    h = stats.download()
    print("rocket %d x %d: %.2f ms  %.3e inst-steps/s  apogee mean %.9f" % (n, steps, ms, n * steps / ms * 1e3, h[0].mean()))
elif cfg == "ho":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
    w = W.ho_ensemble(n)
    d = [eng.to_device(w[k]) for k in ("m", "k", "c", "x0", "v0")]
    st = eng.empty((2, n))
    ms = timed(lambda: eng.ho_rk4(*d, n, 0.0, 0.01, 1000, mem=MEM_DEVICE, fp_mode=FP_CONTRACT, want_status=False, state_out=st, sync=False), 5)
    print("ho %d x 1000: %.3f ms  %.3e inst-steps/s  checksum %.12e" % (n, ms, n * 1000 / ms * 1e3, float(st.download()[:, ::4097].sum())))
elif cfg == "lin":
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
    w = W.cartpole_points(P)
    dx, du = eng.to_device(w["x"]), eng.to_device(w["u"])
    A, B = eng.empty((16, P)), eng.empty((4, P))
    ms = timed(lambda: eng.cartpole_linearize(1.0, 0.1, 0.5, 9.81, dx, du, P, mem=MEM_DEVICE, A=A, B=B, want_status=False, sync=False), 20)
    print("lin %d: %.4f ms  %.1f GB/s (200 B/point)  checksum %.12e" % (P, ms, 200.0 * P / ms / 1e6, float(A.download()[:, ::4097].sum())))
elif cfg in ("grid", "grid_strict"):
    p = W.GRID2D_PARAMS
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    gp = eng.grid2d_params(n, n, 0, p["mass"], p["spacing"], p["k"], p["c"])
    state = eng.grid2d_initial_state(gp, mem=MEM_DEVICE)
    mode = FP_STRICT if cfg == "grid_strict" else FP_CONTRACT
    for per_stage in (False, True):
        ms = timed(lambda: eng.grid2d_rk4(gp, p["dt"], 10, state, fp_mode=mode, sync=False, per_stage=per_stage), 2) / 10
        print("%s %d^2 %s: %.3f ms/step  %.3e mass-steps/s  %.1f GB/s at 544 B/mass-step" %
              (cfg, n, "per-stage" if per_stage else "fused", ms, n * n / ms * 1e3, 544.0 * n * n / ms / 1e6))
elif cfg == "closed":
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
    rng = np.random.default_rng(3)
    plant = {k: eng.to_device(v) for k, v in dict(mc=rng.uniform(0.8, 1.5, P), m=rng.uniform(0.05, 0.2, P),
                                                   L=rng.uniform(0.4, 0.8, P), g=np.full(P, 9.81)).items()}
    dx, du = eng.to_device(np.zeros((4, P))), eng.to_device(np.zeros(P))
    A, B, _ = eng.cartpole_linearize(plant["mc"], plant["m"], plant["L"], plant["g"], dx, du, P, mem=MEM_DEVICE, want_status=False)
    K = eng.empty((4, P))
    eng.lqr4_gain(A, B, P, np.diag([10.0, 100.0, 1.0, 10.0]), 0.01, mem=MEM_DEVICE, fp_mode=FP_CONTRACT, K=K)
    s0 = [eng.to_device(v) for v in (rng.uniform(-0.5, 0.5, P), rng.uniform(-0.1, 0.1, P), np.zeros(P), np.zeros(P))]
    for mode, name in ((FP_CONTRACT, "contract"), (FP_STRICT, "strict")):
        st = eng.empty((4, P))
        ms = timed(lambda: eng.cartpole_feedback_rk4(plant["mc"], plant["m"], plant["L"], plant["g"], *s0, K, P, 0.0, 0.002, 2000,
                                                     u_min=-50.0, u_max=50.0, saturate=True, mem=MEM_DEVICE, want_status=False,
                                                     want_stats=False, fp_mode=mode, sync=False, state_out=st), 2)
        print("closed loop %d x 2000 %s: %.3f ms  %.3e inst-steps/s  checksum %.12e" % (P, name, ms, P * 2000 / ms * 1e3, float(st.download().sum())))
elif cfg == "lqr":
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
    rng = np.random.default_rng(3)
    plant = {k: eng.to_device(v) for k, v in dict(mc=rng.uniform(0.8, 1.5, P), m=rng.uniform(0.05, 0.2, P),
                                                   L=rng.uniform(0.4, 0.8, P), g=np.full(P, 9.81)).items()}
    dx, du = eng.to_device(np.zeros((4, P))), eng.to_device(np.zeros(P))
    A, B, _ = eng.cartpole_linearize(plant["mc"], plant["m"], plant["L"], plant["g"], dx, du, P, mem=MEM_DEVICE, want_status=False)
    K = eng.empty((4, P))
    Q = np.diag([10.0, 100.0, 1.0, 10.0])
    for mode, name in ((FP_CONTRACT, "contract"), (FP_STRICT, "strict")):
        ms = timed(lambda: eng.lqr4_gain(A, B, P, Q, 0.01, mem=MEM_DEVICE, fp_mode=mode, K=K, want_iters=False, sync=False), 2)
        _, _, iters, st = eng.lqr4_gain(A, B, P, Q, 0.01, mem=MEM_DEVICE, fp_mode=mode, K=K)
        print("lqr4 %d %s: %.3f ms  %.3e gains/s  mean iterations %.2f  failed %d  checksum %.12e" %
              (P, name, ms, P / ms * 1e3, float(iters.download().mean()), int(np.count_nonzero(st.download())), float(K.download().sum())))
elif cfg == "ugrid":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    ug = eng.ugrid_params(n, n, 1.0, 0.05, 0.5, 50.0, 0.5, 0.5)
    for mode, name in ((FP_STRICT, "strict"), (FP_CONTRACT, "contract")):
        for integ, iname in ((0, "rk4"), (1, "symplectic euler"), (2, "velocity verlet")):
            state = eng.ugrid_initial_state(ug, mem=MEM_DEVICE)
            ms = timed(lambda: eng.ugrid_rk4(ug, 1e-3, 6, state, fp_mode=mode, sync=False, integrator=integ), 2) / 6
            print("ugrid %d^2 %s %s: %.3f ms/step  %.3e mass-steps/s" % (n, name, iname, ms, n * n / ms * 1e3))
            if integ == 0:
                ms = timed(lambda: eng.ugrid_rk4(ug, 1e-3, 6, state, fp_mode=mode, sync=False, stage_pairs=True), 2) / 6
                print("ugrid %d^2 %s rk4 stage pairs: %.3f ms/step  %.3e mass-steps/s" % (n, name, ms, n * n / ms * 1e3))
            state.free()
eng.close()
