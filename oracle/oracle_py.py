"""ctypes access to the CHECKER libraries -- test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module; nothing under sopot_b200/ does.

Two interchangeable checkers, same call surface:
  kind="port"       oracle/libsopot_oracle.so   plain-C restatement (oracle/sopot_oracle.c)
  kind="reference"  oracle/_ref/libsopot_ref.so the unmodified reference compiled from
                    /root/reference through oracle/ref_harness.cpp (exists where it was built)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_sz = ctypes.c_size_t
_d = ctypes.c_double
_i = ctypes.c_int


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_dp)


def _f64(a, n=None):
    a = np.asarray(a, dtype=np.float64)
    if n is not None and a.size == 1 and n != 1:
        a = np.full(n, float(a.reshape(-1)[0]))
    return np.ascontiguousarray(a)


def build(force: bool = False) -> None:
    """(Re)build the checker libraries with oracle/Makefile."""
    subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []) + ["all"], check=True,
                   stdout=subprocess.DEVNULL)


def have_reference() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libsopot_ref.so"))


class Checker:
    def __init__(self, kind: str = "port"):
        self.kind = kind
        if kind == "port":
            path, self.pfx = os.path.join(_HERE, "libsopot_oracle.so"), "orc_"
            if not os.path.exists(path):
                build()
        elif kind == "reference":
            path, self.pfx = os.path.join(_HERE, "_ref", "libsopot_ref.so"), "ref_"
        else:
            raise ValueError(kind)
        self.lib = ctypes.CDLL(path)
        self.path = path

    def _fn(self, name, argtypes, restype=_i):
        f = getattr(self.lib, self.pfx + name)
        f.argtypes, f.restype = argtypes, restype
        return f

    def hardware_threads(self) -> int:
        return int(self._fn("hardware_threads", [])())

    # ---------------- config 1 ----------------
    def ho_rk4(self, m, k, c, x0, v0, t0, t1, dt, store_every=0, n_snap=0, nthreads=1):
        k = _f64(k); n = k.size
        m, c, x0, v0 = (_f64(a, n) for a in (m, c, x0, v0))
        fin = np.empty((2, n)); tl = np.empty(n)
        traj = np.full((n_snap, 2, n), np.nan) if store_every else None
        self._fn("ho_rk4", [_sz] + [_dp] * 5 + [_d] * 3 + [_sz, _dp, _dp, _dp, _i])(
            n, _p(m), _p(k), _p(c), _p(x0), _p(v0), t0, t1, dt, store_every, _p(fin), _p(traj), _p(tl),
            nthreads)
        return fin, traj, tl

    # ---------------- config 2 ----------------
    def dpend_rk4(self, m1, m2, L1, L2, g, th1, th2, w1, w2, t0, t1, dt, store_every=0, n_snap=0,
                  nthreads=1):
        th1 = _f64(th1); n = th1.size
        m1, m2, L1, L2, g, th2, w1, w2 = (_f64(a, n) for a in (m1, m2, L1, L2, g, th2, w1, w2))
        fin = np.empty((4, n)); tl = np.empty(n)
        traj = np.full((n_snap, 4, n), np.nan) if store_every else None
        self._fn("dpend_rk4", [_sz] + [_dp] * 9 + [_d] * 3 + [_sz, _dp, _dp, _dp, _i])(
            n, _p(m1), _p(m2), _p(L1), _p(L2), _p(g), _p(th1), _p(th2), _p(w1), _p(w2), t0, t1, dt,
            store_every, _p(fin), _p(traj), _p(tl), nthreads)
        return fin, traj, tl

    def dpend_rhs(self, m1, m2, L1, L2, g, state):
        state = _f64(state); n = state.shape[1]
        m1, m2, L1, L2, g = (_f64(a, n) for a in (m1, m2, L1, L2, g))
        out = np.empty((4, n))
        self._fn("dpend_rhs", [_sz] + [_dp] * 7)(n, _p(m1), _p(m2), _p(L1), _p(L2), _p(g), _p(state),
                                                  _p(out))
        return out

    # ---------------- config 3 ----------------
    def cartpole_linearize(self, x, u, mc, m, L, g, nthreads=1):
        x = _f64(x); P = x.shape[1]
        u, mc, m, L, g = (_f64(a, P) for a in (u, mc, m, L, g))
        A = np.empty((16, P)); B = np.empty((4, P))
        if self.kind == "port":
            st = np.zeros(P, dtype=np.int32)
            self._fn("cartpole_linearize", [_sz] + [_dp] * 8 + [_ip, _i])(
                P, _p(x), _p(u), _p(mc), _p(m), _p(L), _p(g), _p(A), _p(B),
                st.ctypes.data_as(_ip), nthreads)
        else:
            self._fn("cartpole_linearize", [_sz] + [_dp] * 8 + [_i])(
                P, _p(x), _p(u), _p(mc), _p(m), _p(L), _p(g), _p(A), _p(B), nthreads)
        return A, B

    def dpend_jacobian(self, m1, m2, L1, L2, g, state):
        state = _f64(state); n = state.shape[1]
        m1, m2, L1, L2, g = (_f64(a, n).reshape(-1) for a in (m1, m2, L1, L2, g))
        J = np.empty((16, n))
        self._fn("dpend_jacobian", [_sz] + [_dp] * 7)(n, _p(m1), _p(m2), _p(L1), _p(L2), _p(g), _p(state), _p(J))
        return J

    def cartpole_rk4(self, mc, m, L, g, s0, force, t0, t1, dt, nthreads=1):
        s0 = _f64(s0); n = s0.shape[1]
        mc, m, L, g, force = (_f64(a, n) for a in (mc, m, L, g, force))
        fin = np.empty((4, n))
        self._fn("cartpole_rk4", [_sz] + [_dp] * 6 + [_d] * 3 + [_dp, _i])(
            n, _p(mc), _p(m), _p(L), _p(g), _p(s0), _p(force), t0, t1, dt, _p(fin), nthreads)
        return fin

    # ---------------- section 8f-1: LQR and closed loop ----------------
    def lqr4(self, A, B, Q16, R, dt=0.01, max_iter=1000, tol=1e-8, nthreads=1):
        A = _f64(A); P = A.shape[1]
        B, Q16 = _f64(B), _f64(Q16).reshape(16)
        K = np.empty((4, P)); Pric = np.empty((16, P))
        iters = np.zeros(P, dtype=np.int32); st = np.zeros(P, dtype=np.int32)
        self._fn("lqr4", [_sz, _dp, _dp, _dp, _d, _d, _sz, _d, _dp, _dp, _ip, _ip, _i])(
            P, _p(A), _p(B), _p(Q16), R, dt, max_iter, tol, _p(K), _p(Pric), iters.ctypes.data_as(_ip),
            st.ctypes.data_as(_ip), nthreads)
        return K, Pric, iters, st

    def cartpole_feedback_rk4(self, mc, m, L, g, s0, K, xref, umin, umax, saturate, t0, dt, n_steps, nthreads=1):
        s0 = _f64(s0); n = s0.shape[1]
        mc, m, L, g = (_f64(a, n) for a in (mc, m, L, g))
        K = _f64(K); xref = None if xref is None else _f64(xref)
        fin = np.empty((4, n)); stats = np.empty((2, n))
        self._fn("cartpole_feedback_rk4", [_sz] + [_dp] * 7 + [_d, _d, _i, _d, _d, _sz, _dp, _dp, _i])(
            n, _p(mc), _p(m), _p(L), _p(g), _p(s0), _p(K), _p(xref), umin, umax, int(saturate), t0, dt, n_steps,
            _p(fin), _p(stats), nthreads)
        return fin, stats

    def cartn_rk4(self, n_links, mc, m, L, g, s0, force, t0, t1, dt, nthreads=1):
        """CartNPendulum<N,double>, constant force: m, L [N][n]; s0 [2(N+1)][n]."""
        s0 = _f64(s0); n = s0.shape[1]
        m, L = _f64(m).reshape(n_links, n), _f64(L).reshape(n_links, n)
        mc, g, force = (_f64(a, n).reshape(-1) for a in (mc, g, force))
        fin = np.empty_like(s0)
        rc = self._fn("cartn_rk4", [_i, _sz] + [_dp] * 6 + [_d] * 3 + [_dp, _i])(
            n_links, n, _p(mc), _p(m), _p(L), _p(g), _p(s0), _p(force), t0, t1, dt, _p(fin), nthreads)
        if rc:
            raise KeyError("link count not instantiated")
        return fin

    def cartn_linearize(self, n_links, mc, m, L, g, x, u, nthreads=1):
        """makeLinearizer<2(N+1),1> around CartNPendulum<N,Dual>: A [(2N+2)^2][P], B [2N+2][P] (reference only)."""
        assert self.kind == "reference"
        x = _f64(x); P = x.shape[1]; ns = 2 * (n_links + 1)
        m, L = _f64(m).reshape(n_links, P), _f64(L).reshape(n_links, P)
        mc, g, u = (_f64(a, P).reshape(-1) for a in (mc, g, u))
        A = np.empty((ns * ns, P)); B = np.empty((ns, P))
        rc = self._fn("cartn_linearize", [_i, _sz] + [_dp] * 8 + [_i])(n_links, P, _p(mc), _p(m), _p(L), _p(g), _p(x), _p(u),
                                                                       _p(A), _p(B), nthreads)
        if rc:
            raise KeyError("link count not instantiated")
        return A, B

    def lqr_n(self, ns, A, B, Q, R, dt=0.01, max_iter=1000, tol=1e-8, nthreads=1):
        A = _f64(A); P = A.shape[1]
        B, Q = _f64(B), _f64(Q).reshape(ns * ns)
        K = np.empty((ns, P)); Pric = np.empty((ns * ns, P))
        iters = np.zeros(P, dtype=np.int32); st = np.zeros(P, dtype=np.int32)
        rc = self._fn("lqr_n", [_i, _sz, _dp, _dp, _dp, _d, _d, _sz, _d, _dp, _dp, _ip, _ip, _i])(
            ns, P, _p(A), _p(B), _p(Q), R, dt, max_iter, tol, _p(K), _p(Pric), iters.ctypes.data_as(_ip), st.ctypes.data_as(_ip), nthreads)
        assert rc == 0
        return K, Pric, iters, st

    def cartn_feedback_rk4(self, n_links, mc, m, L, g, s0, K, xref, umin, umax, saturate, t0, dt, n_steps, nthreads=1):
        s0 = _f64(s0); n = s0.shape[1]
        m, L = _f64(m).reshape(n_links, n), _f64(L).reshape(n_links, n)
        mc, g = (_f64(a, n).reshape(-1) for a in (mc, g))
        K = _f64(K); xref = None if xref is None else _f64(xref)
        fin = np.empty_like(s0); stats = np.empty((2, n))
        rc = self._fn("cartn_feedback_rk4", [_i, _sz] + [_dp] * 7 + [_d, _d, _i, _d, _d, _sz, _dp, _dp, _i])(
            n_links, n, _p(mc), _p(m), _p(L), _p(g), _p(s0), _p(K), _p(xref), umin, umax, int(saturate), t0, dt, n_steps,
            _p(fin), _p(stats), nthreads)
        assert rc == 0
        return fin, stats

    # ---------------- section 8f-2: unified 6-state grid ----------------
    def ugrid(self, rows, cols, prm, state=None, dt=0.0, n_steps=0, mode="rhs"):
        """prm: mass, radius, spacing, stiffness, rest_length, damping, min_distance, repulsion_stiffness, h_attach0/1, v_attach0/1
        [, topology (0 quad / 1 triangle), diagonal rest length, main-diagonal attach 0/1, anti-diagonal attach 0/1]."""
        prm = _f64(prm).reshape(-1)
        prm = np.concatenate([prm, np.zeros(18 - prm.size)])
        n = rows * cols * 6
        out = np.empty(n)
        st = None if state is None else _f64(state).reshape(n)
        rc = self._fn("ugrid", [_sz, _sz, _dp, _d, _sz, _dp, _dp, _i])(rows, cols, _p(prm), dt, n_steps, _p(st), _p(out),
                                                                        {"rhs": 0, "rk4": 1, "init": 2, "symplectic_euler": 3, "velocity_verlet": 4}[mode])
        if rc:
            raise KeyError("grid size not instantiated in the reference harness")
        return out.reshape(rows * cols, 6)

    # ---------------- plugin-path check: coupled oscillator ----------------
    def coupled_rk4(self, m1, m2, k, L0, c, x1, x2, t0, t1, dt, nthreads=1):
        k = _f64(k).reshape(-1); n = k.size
        m1, m2, L0, c, x1, x2 = (_f64(a, n).reshape(-1) for a in (m1, m2, L0, c, x1, x2))
        fin = np.empty((4, n)); en = np.empty(n)
        self._fn("coupled_rk4", [_sz] + [_dp] * 7 + [_d] * 3 + [_dp, _dp, _i])(
            n, _p(m1), _p(m2), _p(k), _p(L0), _p(c), _p(x1), _p(x2), t0, t1, dt, _p(fin), _p(en), nthreads)
        return fin, en

    # ---------------- config 4 ----------------
    @staticmethod
    def _tab(t, cols):
        if t is None:
            return 0, None
        t = _f64(t)
        assert t.ndim == 2 and t.shape[1] == cols
        return t.shape[0], t

    @staticmethod
    def aero_blob(tables: dict | None, engine_on: bool = True):
        """{name: (row_vals, col_vals, data)} -> flat array: engine_on, then 7 x (rows, cols, row_vals, col_vals, data)."""
        if tables is None:
            return None
        names = ("cd_power_on", "cd_power_off", "cl_power_on", "cl_power_off", "cs_power_on", "cs_power_off", "cmq_yz")
        assert set(tables) <= set(names), tables.keys()
        out = [np.array([1.0 if engine_on else 0.0])]
        for k in names:
            if k in tables:
                rv, cv, data = (np.asarray(x, dtype=np.float64) for x in tables[k])
                assert data.shape == (rv.size, cv.size)
                out += [np.array([rv.size, cv.size], dtype=np.float64), rv, cv, data.ravel()]
            else:
                out.append(np.zeros(2))
        return np.ascontiguousarray(np.concatenate(out))

    def rocket_rk4(self, elev, az, diam, g0, engine, mass_tab, ix_tab, iyz_tab, t_end, dt,
                   store_every=0, n_snap=0, want_stats=True, nthreads=1, aero=None, engine_on=True):
        elev = _f64(elev); n = elev.size
        az, diam, g0 = (_f64(a, n) for a in (az, diam, g0))
        er, engine = self._tab(engine, 8)
        mr, mass_tab = self._tab(mass_tab, 2)
        xr, ix_tab = self._tab(ix_tab, 2)
        yr, iyz_tab = self._tab(iyz_tab, 2)
        fin = np.empty((14, n))
        stats = np.empty((8, n)) if want_stats else None
        traj = np.full((n_snap, 14, n), np.nan) if store_every else None
        blob = self.aero_blob(aero, engine_on)
        rc = self._fn("rocket_rk4_aero", [_sz] + [_dp] * 4 + [_sz, _dp] * 4 + [_d, _d, _sz, _dp, _dp, _dp, _i, _dp])(
            n, _p(elev), _p(az), _p(diam), _p(g0), er, _p(engine), mr, _p(mass_tab), xr, _p(ix_tab),
            yr, _p(iyz_tab), t_end, dt, store_every, _p(fin), _p(traj), _p(stats), nthreads, _p(blob))
        assert rc == 0
        return fin, stats, traj

    def rocket_rhs(self, elev, az, diam, g0, engine, mass_tab, state, aero=None, engine_on=True):
        state = _f64(state); n = state.shape[1]
        er, engine = self._tab(engine, 8)
        mr, mass_tab = self._tab(mass_tab, 2)
        out = np.empty((14, n))
        blob = self.aero_blob(aero, engine_on)
        rc = self._fn("rocket_rhs_aero", [_sz, _d, _d, _d, _d, _sz, _dp, _sz, _dp, _dp, _dp, _dp])(
            n, elev, az, diam, g0, er, _p(engine), mr, _p(mass_tab), _p(state), _p(out), _p(blob))
        assert rc == 0
        return out

    def rocket_state_functions(self, elev, az, diam, g0, engine, mass_tab, state, aero=None, engine_on=True):
        """reference only: the registry's state functions at state[14][n] -> [28][n] (rows as in include/sopot_b200.h)"""
        assert self.kind == "reference"
        state = _f64(state); n = state.shape[1]
        er, engine = self._tab(engine, 8)
        mr, mass_tab = self._tab(mass_tab, 2)
        out = np.empty((28, n))
        blob = self.aero_blob(aero, engine_on)
        rc = self._fn("rocket_state_functions", [_sz, _d, _d, _d, _d, _sz, _dp, _sz, _dp, _dp, _dp, _dp])(
            n, elev, az, diam, g0, er, _p(engine), mr, _p(mass_tab), _p(state), _p(out), _p(blob))
        assert rc == 0
        return out

    def engine_thrust(self, engine, tq, patm):
        er, engine = self._tab(engine, 8)
        tq, patm = _f64(tq), _f64(patm)
        out = np.empty(tq.size)
        self._fn("engine_thrust", [_sz, _dp, _sz, _dp, _dp, _dp])(er, _p(engine), tq.size, _p(tq),
                                                                   _p(patm), _p(out))
        return out

    def atmosphere(self, h):
        h = _f64(h)
        T, P, rho, a = (np.empty(h.size) for _ in range(4))
        self._fn("atmosphere", [_sz] + [_dp] * 5)(h.size, _p(h), _p(T), _p(P), _p(rho), _p(a))
        return T, P, rho, a

    # ---------------- config 5 ----------------
    def grid2d_rhs(self, rows, cols, topo, mass, spacing, k, c, state):
        state = _f64(state).ravel()
        out = np.empty_like(state)
        if self.kind == "port":
            self._fn("grid2d_rhs", [_sz, _sz, _i, _d, _d, _d, _d, _dp, _dp])(
                rows, cols, topo, mass, spacing, k, c, _p(state), _p(out))
        else:
            rc = self._fn("grid2d", [_sz, _sz, _i, _d, _d, _d, _d, _dp, _i, _d, _sz, _dp])(
                rows, cols, topo, mass, spacing, k, c, _p(state), 0, 0.0, 0, _p(out))
            if rc < 0:
                raise KeyError("grid size not instantiated in the reference harness")
        return out

    def grid2d_rk4(self, rows, cols, topo, mass, spacing, k, c, dt, n_steps, state, nthreads=1):
        state = _f64(state).ravel().copy()
        if self.kind == "port":
            rc = self._fn("grid2d_rk4", [_sz, _sz, _i, _d, _d, _d, _d, _d, _sz, _dp, _i])(
                rows, cols, topo, mass, spacing, k, c, dt, n_steps, _p(state), nthreads)
            assert rc == n_steps
            return state
        out = np.empty_like(state)
        rc = self._fn("grid2d", [_sz, _sz, _i, _d, _d, _d, _d, _dp, _i, _d, _sz, _dp])(
            rows, cols, topo, mass, spacing, k, c, _p(state), 1, dt, n_steps, _p(out))
        if rc < 0:
            raise KeyError("grid size not instantiated in the reference harness")
        assert rc == n_steps, (rc, n_steps)
        return out

    def grid2d_initial_state(self, rows, cols, spacing, mass=1.0, k=1.0, c=0.0, topo=0):
        out = np.empty(4 * rows * cols)
        if self.kind == "port":
            self._fn("grid2d_initial_state", [_sz, _sz, _d, _dp])(rows, cols, spacing, _p(out))
        else:
            rc = self._fn("grid2d", [_sz, _sz, _i, _d, _d, _d, _d, _dp, _i, _d, _sz, _dp])(
                rows, cols, topo, mass, spacing, k, c, _p(out), 2, 0.0, 0, _p(out))
            if rc < 0:
                raise KeyError("grid size not instantiated in the reference harness")
        return out

    def rk4_num_steps(self, t0, t1, dt) -> int:
        if self.kind != "port":
            raise NotImplementedError
        return int(self._fn("rk4_num_steps", [_d, _d, _d], _sz)(t0, t1, dt))
