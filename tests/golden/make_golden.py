"""Regenerate tests/golden/*.npz from the UNMODIFIED reference compiled in the build container.

    python tests/golden/make_golden.py          (needs oracle/_ref/libsopot_ref.so, i.e. /root/reference)

Every array named ref_* below is an output of the reference's own code path (makeTypedODESystem ->
createRK4Solver().solve, makeLinearizer<4,1>, Rocket component system, makeGrid2DSystem), driven by
oracle/ref_harness.cpp; inputs are stored beside them so the fixtures do not depend on any RNG.
The reference's own tests contain almost no numeric goldens for this path;
the known answers quoted in BASELINE.md section 2 are asserted in tests/test_oracle.py as well.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle_py import Checker  # noqa: E402
from sopot_b200 import workloads as W  # noqa: E402

OUT = os.environ.get("SOPOT_GOLDEN_OUT") or os.path.dirname(os.path.abspath(__file__))     # fixtures are written here


def main():
    R = Checker("reference")
    # ---- config 1 ----
    w = W.ho_ensemble(64)
    w["m"][0], w["k"][0], w["c"][0], w["x0"][0] = 2.0, 9.0, 0.1, 1.5          # README nominal first
    fin, traj, tl = R.ho_rk4(w["m"], w["k"], w["c"], w["x0"], w["v0"], 0.0, 10.0, 0.01, 100, 10)
    np.savez(os.path.join(OUT, "ho.npz"), m=w["m"], k=w["k"], c=w["c"], x0=w["x0"], v0=w["v0"],
             ref_final=fin, ref_traj=traj, ref_t_last=tl)
    # ---- config 2 ----
    w = W.dpend_ensemble(16)
    w["th1"][0], w["th2"][0] = np.pi / 4, np.pi / 6                           # BASELINE.md nominal first
    w["th1"][1], w["th2"][1] = 3.0, 3.0                                       # most chaotic survey case
    args = [w[k] for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")]
    fin, traj, _ = R.dpend_rk4(*args, 0.0, 10.0, 1e-3, 1000, 10)
    st = np.stack([np.linspace(-3, 3, 32), np.linspace(3, -2, 32), np.linspace(-4, 4, 32), np.linspace(5, -5, 32)])
    rhs = R.dpend_rhs(1.3, 0.7, 0.9, 1.1, 9.81, st)
    # shadow twins: theta1 nudged by one ulp
    th1n = np.nextafter(w["th1"], np.inf)
    args_n = [w["m1"], w["m2"], w["L1"], w["L2"], w["g"], th1n, w["th2"], w["w1"], w["w2"]]
    _, traj_n, _ = R.dpend_rk4(*args_n, 0.0, 10.0, 1e-3, 1000, 10)
    np.savez(os.path.join(OUT, "dpend.npz"), **{k: w[k] for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")},
             ref_final=fin, ref_traj=traj, ref_traj_twin=traj_n, rhs_state=st, ref_rhs=rhs)
    # ---- config 3 ----
    w = W.cartpole_points(256)
    w["x"][:, 0] = 0.0; w["u"][0] = 0.0                                       # upright
    w["x"][:, 1] = [0.1, 0.3, -0.2, 0.5]; w["u"][1] = 1.5                     # BASELINE.md point
    A, B = R.cartpole_linearize(w["x"], w["u"], w["mc"], w["m"], w["L"], w["g"])
    fin = R.cartpole_rk4(w["mc"][:16], w["m"][:16], w["L"][:16], w["g"][:16], w["x"][:, :16], w["u"][:16], 0.0, 2.0, 0.01)
    np.savez(os.path.join(OUT, "cartpole.npz"), x=w["x"], u=w["u"], mc=w["mc"], m=w["m"], L=w["L"], g=w["g"],
             ref_A=A, ref_B=B, ref_rk4_final=fin)
    # ---- config 4 ----
    w = W.rocket_ensemble(8)
    w["elev"][0], w["az"][0], w["diam"][0], w["g0"][0] = 85.0, 0.0, 0.16, 9.80665
    fin, stats, traj = R.rocket_rk4(w["elev"], w["az"], w["diam"], w["g0"], w["engine"], w["mass_tab"], None, None,
                                    30.0, 0.01, 500, 6)
    h = np.concatenate([np.linspace(-100, 101000, 401), [0, 11000, 20000, 32000, 47000, 51000, 71000, 100000]])
    T, P, rho, a = R.atmosphere(h)
    tq = np.linspace(-0.1, 3.7, 77); pa = np.linspace(101325, 60000, 77)
    thrust = R.engine_thrust(w["engine"], tq, pa)
    rhs_state = traj[:, :, :4].transpose(1, 0, 2).reshape(14, -1).copy()      # states along real trajectories
    rhs = R.rocket_rhs(85.0, 0.0, 0.16, 9.80665, w["engine"], w["mass_tab"], rhs_state)
    ix = np.array([[0.0, 0.12], [3.5, 0.09]]); iyz = np.array([[0.0, 8.0], [3.5, 6.5]])
    fin2, stats2, _ = R.rocket_rk4(w["elev"][:4], w["az"][:4], w["diam"][:4], w["g0"][:4], w["engine"], w["mass_tab"],
                                   ix, iyz, 10.0, 0.01)
    np.savez(os.path.join(OUT, "rocket.npz"), elev=w["elev"], az=w["az"], diam=w["diam"], g0=w["g0"],
             engine=w["engine"], mass_tab=w["mass_tab"], ref_final=fin, ref_stats=stats, ref_traj=traj,
             atm_h=h, ref_atm_T=T, ref_atm_P=P, ref_atm_rho=rho, ref_atm_a=a, thrust_t=tq, thrust_patm=pa,
             ref_thrust=thrust, rhs_state=rhs_state, ref_rhs=rhs, ix_tab=ix, iyz_tab=iyz, ref_final_inertia=fin2,
             ref_stats_inertia=stats2)
    # ---- config 5 ----
    rng = np.random.Generator(np.random.MT19937(5))
    out = {}
    for (r, c, t) in [(2, 2, 0), (3, 3, 0), (4, 5, 0), (2, 5, 0), (3, 3, 1), (4, 4, 1), (3, 4, 2), (4, 4, 2)]:
        s0 = R.grid2d_initial_state(r, c, 0.5, topo=t)
        s = s0 + rng.normal(0, 0.1, s0.shape)
        key = f"g{r}x{c}t{t}"
        out[key + "_init"] = s0
        out[key + "_state"] = s
        out[key + "_rhs"] = R.grid2d_rhs(r, c, t, 1.3, 0.5, 50.0, 0.5, s)
        out[key + "_rk4"] = R.grid2d_rk4(r, c, t, 1.3, 0.5, 50.0, 0.5, 1e-3, 200, s)
    # a degenerate state: two coincident masses (len < 1e-10 clamp, indexed_spring_2d.hpp:125)
    s = R.grid2d_initial_state(3, 3, 0.5)
    s[4 * 4 + 0], s[4 * 4 + 1] = s[4 * 3 + 0], s[4 * 3 + 1]
    out["g3x3t0_degenerate_state"] = s
    out["g3x3t0_degenerate_rhs"] = R.grid2d_rhs(3, 3, 0, 1.0, 0.5, 50.0, 0.5, s)
    np.savez(os.path.join(OUT, "grid2d.npz"), **out)
    control_golden(R)
    cartn_linearize_golden(R)
    rocket_aero_golden(R)
    rocket_atmosphere_golden(R)
    rocket_state_functions_golden(R)
    dpend_jacobian_golden(R)
    control_n_golden(R)
    ugrid_golden(R)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


def control_golden(R):
    """Section 8f-1: the reference's LQR<4,1>::solve on its own linearizer output, and the closed loop
    StateFeedbackController<4,1> -> CartNPendulum<1> -> RK4Solver (one step per controller sample)."""
    P = 48
    w = W.cartpole_points(P, seed=31)
    x = w["x"].copy()
    x[1] *= 0.15                                   # near upright: stabilisable
    x[:, 0] = 0.0                                  # the upright equilibrium itself first
    mc = 1.0 + 0.2 * np.linspace(-1, 1, P); m = 0.1 * (1 + 0.3 * np.cos(np.arange(P))); L = 0.5 * (1 + 0.2 * np.sin(np.arange(P)))
    g = np.full(P, 9.81)
    A, B = R.cartpole_linearize(x, w["u"], mc, m, L, g)
    Q = np.diag([10.0, 100.0, 1.0, 10.0]) + 0.1     # full (non-diagonal) weight matrix
    Q = 0.5 * (Q + Q.T)
    out = dict(x=x, u=w["u"], mc=mc, m=m, L=L, g=g, A=A, B=B, Q=Q)
    for tag, r, dt, mi, tol in (("a", 0.01, 0.01, 1000, 1e-8), ("b", 0.5, 0.002, 300, 1e-10)):
        K, Pr, _, st = R.lqr4(A, B, Q, r, dt, mi, tol)
        out["K_" + tag], out["P_" + tag], out["st_" + tag] = K, Pr, st
    # degenerate problems: B = 0 with R = 0 -> "singular" throw; unstable + uncontrollable -> diverges
    A2 = np.zeros((16, 2)); B2 = np.zeros((4, 2))
    A2[0 * 4 + 0, 1] = 500.0
    K2, P2, _, st2 = R.lqr4(A2, B2, np.eye(4), 0.0, 0.01, 1000, 1e-8)
    K3, P3, _, st3 = R.lqr4(A2, B2, np.eye(4), 1.0, 0.01, 1000, 1e-8)
    out.update(A2=A2, B2=B2, st_sing=st2, st_div=st3, K_div=K3)
    n = 40
    s0 = np.zeros((4, n)); s0[0] = 0.2 * np.sin(np.arange(n)); s0[1] = 0.08 * np.linspace(-1, 1, n); s0[3] = 0.1 * np.cos(np.arange(n))
    Kc = out["K_a"][:, :n].copy()
    xref = np.zeros((4, n)); xref[0] = 0.05
    fin, stats = R.cartpole_feedback_rk4(mc[:n], m[:n], L[:n], g[:n], s0, Kc, xref, -15.0, 15.0, 1, 0.0, 0.002, 1500)
    fin_ns, stats_ns = R.cartpole_feedback_rk4(mc[:n], m[:n], L[:n], g[:n], s0, Kc, None, 0.0, 0.0, 0, 0.0, 0.002, 700)
    out.update(cl_s0=s0, cl_K=Kc, cl_xref=xref, cl_final=fin, cl_stats=stats, cl_final_ns=fin_ns, cl_stats_ns=stats_ns)
    np.savez(os.path.join(OUT, "control.npz"), **out)



def cartn_linearize_golden(R):
    """N-link linearization (makeLinearizer<2(N+1),1> around CartNPendulum<N,Dual<double,2N+3>>), N = 1, 2, 6."""
    out = {}
    rng = np.random.default_rng(17)
    for N, P in ((1, 37), (2, 21), (6, 9)):
        ns = 2 * (N + 1)
        m = rng.uniform(0.05, 0.3, (N, P)); L = rng.uniform(0.3, 0.8, (N, P)); mc = rng.uniform(0.5, 2.0, P); g = np.full(P, 9.81)
        x = rng.uniform(-0.5, 0.5, (ns, P)); x[:, 0] = 0.0; u = rng.uniform(-2, 2, P)
        A, B = R.cartn_linearize(N, mc, m, L, g, x, u)
        out.update({f"n{N}_mc": mc, f"n{N}_m": m, f"n{N}_L": L, f"n{N}_g": g, f"n{N}_x": x, f"n{N}_u": u, f"n{N}_A": A, f"n{N}_B": B})
    np.savez(os.path.join(OUT, "cartn_linearize.npz"), **out)


def rocket_aero_tables():
    """Synthetic coefficient tables in the reference's layouts (rocket/axisymmetric_aero.hpp:133-167, :276-281)."""
    mach = np.array([0.0, 0.3, 0.6, 0.9, 1.2]); ang = np.array([0.0, 2.0, 5.0, 10.0, 20.0])
    cd = 0.25 + 0.1 * mach[:, None] + 0.004 * ang[None, :]
    cl = 0.02 * ang[None, :] * (1 + 0.2 * mach[:, None])
    cs = 0.015 * ang[None, :] * (1 + 0.1 * mach[:, None])
    aoa_signed = np.array([-20.0, -5.0, 0.0, 5.0, 20.0])
    cmq = -400.0 - 5.0 * np.abs(aoa_signed)[:, None] - 30.0 * mach[None, :]
    return dict(cd_power_on=(mach, ang, cd), cd_power_off=(mach, ang, cd + 0.05), cl_power_on=(mach, ang, cl),
                cs_power_off=(mach, ang, cs), cmq_yz=(aoa_signed, mach, cmq))


def rocket_aero_golden(R):
    """Table-driven aerodynamics: full Rocket component system with Cd/Cl/Cs/cmq tables loaded through the CSV loaders."""
    aero = rocket_aero_tables()
    n = 12
    w = W.rocket_ensemble(n, seed=44)
    out = dict(elev=w["elev"], az=w["az"], diam=w["diam"], g0=w["g0"], engine=w["engine"], mass_tab=w["mass_tab"])
    for tag, on in (("on", True), ("off", False)):
        fin, stats, _ = R.rocket_rk4(w["elev"], w["az"], w["diam"], w["g0"], w["engine"], w["mass_tab"], None, None, 10.0, 0.01,
                                     aero=aero, engine_on=on)
        out["final_" + tag], out["stats_" + tag] = fin, stats
    st = out["final_on"].copy()
    st[11:] = np.array([0.05, -0.03, 0.02])[:, None]            # tumbling: exercises the damping table
    st[0] = 2.0                                                   # powered flight
    out["rhs_state"] = st
    out["rhs"] = R.rocket_rhs(84.0, 30.0, 0.16, 9.80665, w["engine"], w["mass_tab"], st, aero=aero)
    np.savez(os.path.join(OUT, "rocket_aero.npz"), **out)


def rocket_atmosphere_golden(R):
    """One derivative evaluation per atmosphere layer and at every layer boundary (standard_atmosphere.hpp:43-71): the
    trajectory fixtures never leave the lowest layer."""
    g = np.load(os.path.join(OUT, "rocket.npz"))
    alts = np.array([-50.0, 0.0, 1e-3, 5000.0, 10999.999, 11000.0, 11000.001, 15000.0, 20000.0, 20000.5, 26000.0, 32000.0, 32001.0,
                     40000.0, 47000.0, 47000.25, 49000.0, 51000.0, 51010.0, 60000.0, 71000.0, 71000.5, 85000.0, 99999.0, 100000.0,
                     120000.0])
    rng = np.random.default_rng(11)
    st = np.repeat(g["rhs_state"][:, :1], alts.size, axis=1)
    st[3] = alts
    st[0] = np.where(np.arange(alts.size) % 2 == 0, 1.2, 6.0)          # burning / coasting
    st[4:7] = rng.normal(0, 150.0, (3, alts.size))
    q = rng.normal(0, 1, (4, alts.size)); st[7:11] = q / np.linalg.norm(q, axis=0)
    st[11:14] = rng.normal(0, 0.3, (3, alts.size))
    np.savez(os.path.join(OUT, "rocket_atmosphere.npz"), state=st,
             rhs=R.rocket_rhs(85.0, 0.0, 0.16, 9.80665, g["engine"], g["mass_tab"], st))


def rocket_state_functions_golden(R):
    """Every state function the reference's registry provides for the full rocket (rocket_tags.hpp), at the states of the
    atmosphere fixture (all layers, burning and coasting) and, with coefficient tables, at the states of the aero fixture."""
    g = np.load(os.path.join(OUT, "rocket.npz"))
    st = np.load(os.path.join(OUT, "rocket_atmosphere.npz"))["state"]
    out = dict(state=st, values=R.rocket_state_functions(85.0, 0.0, 0.16, 9.80665, g["engine"], g["mass_tab"], st))
    a = np.load(os.path.join(OUT, "rocket_aero.npz"))
    out["aero_state"] = a["rhs_state"]
    out["aero_values"] = R.rocket_state_functions(84.0, 30.0, 0.16, 9.80665, a["engine"], a["mass_tab"], a["rhs_state"],
                                                  aero=rocket_aero_tables())
    np.savez(os.path.join(OUT, "rocket_state_functions.npz"), **out)


def dpend_jacobian_golden(R):
    """computeJacobian (core/typed_component.hpp:486-512) on DoublePendulum<Dual<double,4>>; point 0 is the reference test's
    (tests/double_pendulum_test.cpp:387-389)."""
    rng = np.random.default_rng(17)
    n = 48
    m1, m2, L1, L2 = (rng.uniform(0.5, 2, n) for _ in range(4)); g = np.full(n, 9.81)
    st = np.stack([rng.uniform(-3, 3, n), rng.uniform(-3, 3, n), rng.uniform(-5, 5, n), rng.uniform(-5, 5, n)])
    m1[0], m2[0], L1[0], L2[0] = 1, 1, 1, 1; st[:, 0] = [0.5, 0.3, 0.1, -0.1]       # the reference test's point
    np.savez(os.path.join(OUT, "dpend_jacobian.npz"), m1=m1, m2=m2, L1=L1, L2=L2, g=g, state=st,
             J=R.dpend_jacobian(m1, m2, L1, L2, g, st))


def control_n_golden(R):
    """LQR<6,1> / LQR<14,1> on the N-link linearizations at upright and the N-link closed loop (the reference's
    inverted-pendulum test configuration for 6 links: mc 2, m 0.1, L 0.3, Q = diag(1, 200.., 1, 10..), R 1e-3, dt 5e-3)."""
    out = {}
    for N, P in ((2, 5), (6, 4)):
        ns = 2 * (N + 1)
        m = np.full((N, P), 0.1) * (1 + 0.1 * np.arange(P))[None, :]; L = np.full((N, P), 0.3); mc = np.full(P, 2.0); g = np.full(P, 9.81)
        A, B = R.cartn_linearize(N, mc, m, L, g, np.zeros((ns, P)), np.zeros(P))
        Q = np.diag(np.array([1.0] + [200.0] * N + [1.0] + [10.0] * N))
        K, Pr, _, st = R.lqr_n(ns, A, B, Q, 0.001, 0.005, 1000, 1e-8)
        s0 = np.zeros((ns, P)); s0[1:N + 1] = 0.03
        fin, stats = R.cartn_feedback_rk4(N, mc, m, L, g, s0, K, None, -500.0, 500.0, 1, 0.0, 0.001, 1500)
        out.update({f"n{N}_mc": mc, f"n{N}_m": m, f"n{N}_L": L, f"n{N}_g": g, f"n{N}_A": A, f"n{N}_B": B, f"n{N}_Q": Q, f"n{N}_K": K,
                    f"n{N}_P": Pr, f"n{N}_st": st, f"n{N}_s0": s0, f"n{N}_final": fin, f"n{N}_stats": stats})
    np.savez(os.path.join(OUT, "control_n.npz"), **out)


UGRID_PI = 3.14159265358979323846
UGRID_PARAMS = {   # mass, radius, spacing, stiffness, rest_length, damping, min_distance, repulsion_stiffness, h_attach0/1, v_attach0/1
    "wasm": [1.0, 0.05, 0.5, 50.0, 0.5, 0.5, 0.0, -1.0, 0.0, UGRID_PI, UGRID_PI / 2, -UGRID_PI / 2],      # Grid2DSimulator's layout
    "repel": [0.7, 0.08, 0.4, 80.0, 0.3, 0.9, 0.35, -1.0, 0.0, UGRID_PI, 0.0, UGRID_PI],                  # Spring2D defaults + repulsion
}


UGRID_TRIANGLE = UGRID_PARAMS["wasm"] + [1.0, float(np.sqrt(2.0) * 0.5), UGRID_PI / 4, -3 * UGRID_PI / 4, 3 * UGRID_PI / 4, -UGRID_PI / 4]


def ugrid_golden(R):
    """Unified 6-state grid: ValidatedSystem<double, makeGridTopology<R,C>, PointMass2D, Spring2D> (the wasm Grid2DSimulator's
    system): initial state, one derivative evaluation of a perturbed state, 50 RK4Solver steps."""
    out = {}
    for rows, cols in ((3, 3), (4, 5), (2, 7), (6, 6), (10, 10)):
        for tag, prm in UGRID_PARAMS.items():
            key = f"u{rows}x{cols}_{tag}"
            s0 = R.ugrid(rows, cols, prm, mode="init")
            rng = np.random.default_rng(rows * 100 + cols)
            st = s0 + rng.normal(0, 0.03, s0.shape)
            st[:, 4] = rng.normal(0, 0.3, rows * cols); st[:, 5] = rng.normal(0, 0.5, rows * cols)
            out[key + "_init"], out[key + "_state"] = s0, st
            out[key + "_rhs"] = R.ugrid(rows, cols, prm, st)
            out[key + "_rk4"] = R.ugrid(rows, cols, prm, st, 1e-3, 50, "rk4")
    np.savez(os.path.join(OUT, "ugrid.npz"), **out)
    # the simulator's two other integrators (wasm/wasm_grid2d.cpp:329-413), 80 steps from the same perturbed states
    integ = {}
    for rows, cols in ((4, 5), (10, 10)):
        for mode in ("symplectic_euler", "velocity_verlet"):
            integ[f"u{rows}x{cols}_{mode}"] = R.ugrid(rows, cols, UGRID_PARAMS["wasm"], out[f"u{rows}x{cols}_wasm_state"], 1e-3, 80, mode)
    np.savez(os.path.join(OUT, "ugrid_integrators.npz"), **integ)
    # the "triangle" grid type (makeTriangleGridTopology + createTriangleSystem): diagonal springs of rest length sqrt(2) spacing
    tri = {}
    for rows, cols in ((3, 3), (4, 5), (2, 7), (6, 6)):
        prm = UGRID_TRIANGLE
        key = f"x{rows}x{cols}"
        st = out[f"u{rows}x{cols}_wasm_state"]
        tri[key + "_rhs"] = R.ugrid(rows, cols, prm, st)
        tri[key + "_rk4"] = R.ugrid(rows, cols, prm, st, 1e-3, 50, "rk4")
        tri[key + "_verlet"] = R.ugrid(rows, cols, prm, st, 1e-3, 50, "velocity_verlet")
    np.savez(os.path.join(OUT, "ugrid_triangle.npz"), **tri)


if __name__ == "__main__":
    main()
