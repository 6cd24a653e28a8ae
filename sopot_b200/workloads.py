"""Seeded synthetic workloads for the five BASELINE.json configs.

Pure numpy, host side; the same arrays are handed to the CUDA engine and to the checker, so the
RNG itself is not part of any parity claim.  No reference data files exist for the rocket
(rocket/rocket.hpp:124-147 names CSVs that are not in the repository), so the engine table follows
the only recipe the reference contains: wasm/wasm_rocket.cpp:385-428 (p_c = 5 MPa, CF = 1.6,
expansion ratio 6, T_c = 3000 K, gamma = 1.24, M = 0.024 kg/mol, efficiency 0.96).
"""
from __future__ import annotations

import numpy as np


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.MT19937(seed))


def ho_ensemble(n: int, seed: int = 1) -> dict:
    """Config 1: README damped oscillator m=2,k=9,c=0.1,x0=1.5 with jittered k (5 %) and c (20 %)."""
    r = _rng(seed)
    u1, u2 = r.uniform(-1.0, 1.0, n), r.uniform(-1.0, 1.0, n)
    return dict(m=np.full(n, 2.0), k=9.0 * (1.0 + 0.05 * u1), c=0.1 * (1.0 + 0.2 * u2),
                x0=np.full(n, 1.5), v0=np.zeros(n), t0=0.0, t1=10.0, dt=0.01, n_steps=1000)


def dpend_ensemble(n: int, seed: int = 2) -> dict:
    """Config 2: double pendulum m=L=1, g=9.81, theta1/theta2 ~ U(-3,3), omega=0; dt=1e-3, 10 s."""
    r = _rng(seed)
    return dict(m1=np.ones(n), m2=np.ones(n), L1=np.ones(n), L2=np.ones(n), g=np.full(n, 9.81),
                th1=r.uniform(-3.0, 3.0, n), th2=r.uniform(-3.0, 3.0, n), w1=np.zeros(n),
                w2=np.zeros(n), t0=0.0, t1=10.0, dt=1e-3, n_steps=10000)


def cartpole_points(P: int, seed: int = 3) -> dict:
    """Config 3: cart-pole operating points for A/B linearization (mc=1, m=0.1, L=0.5, g=9.81)."""
    r = _rng(seed)
    x = np.stack([r.uniform(-1, 1, P), r.uniform(-np.pi, np.pi, P), r.uniform(-1, 1, P),
                  r.uniform(-2, 2, P)])
    return dict(x=np.ascontiguousarray(x), u=r.uniform(-5, 5, P), mc=np.ones(P), m=np.full(P, 0.1),
                L=np.full(P, 0.5), g=np.full(P, 9.81))


def rocket_engine_table(thrust_n: float = 1500.0, t_flat: float = 3.0, t_burn: float = 3.5,
                        rows: int = 36) -> np.ndarray:
    """[rows][8] engine table: time, throat_d, exit_d, p_comb, T_comb, gamma, mol_mass, efficiency."""
    t = np.linspace(0.0, t_burn, rows)
    F = np.where(t <= t_flat, thrust_n, thrust_n * (t_burn - t) / (t_burn - t_flat))
    F = np.maximum(F, 0.0)
    p_c, CF, eps = 5.0e6, 1.6, 6.0
    A_throat = np.where(F > 0, F / (p_c * CF), 1e-6)
    d_throat = 2.0 * np.sqrt(A_throat / 3.14159265359)
    d_exit = d_throat * np.sqrt(eps)
    tab = np.stack([t, d_throat, d_exit, np.full(rows, p_c), np.full(rows, 3000.0),
                    np.full(rows, 1.24), np.full(rows, 0.024), np.full(rows, 0.96)], axis=1)
    return np.ascontiguousarray(tab)


def rocket_mass_table(m_wet: float = 20.0, m_dry: float = 14.75, t_burn: float = 3.5,
                      rows: int = 36) -> np.ndarray:
    t = np.linspace(0.0, t_burn, rows)
    return np.ascontiguousarray(np.stack([t, m_wet + (m_dry - m_wet) * t / t_burn], axis=1))


def rocket_ensemble(n: int, seed: int = 4) -> dict:
    """Config 4: launcher elevation 85+-2 deg, azimuth U(0,360), diameter 0.16(1+-2 %), g0(1+-1e-3)."""
    r = _rng(seed)
    return dict(elev=85.0 + 2.0 * r.uniform(-1, 1, n), az=r.uniform(0.0, 360.0, n),
                diam=0.16 * (1.0 + 0.02 * r.uniform(-1, 1, n)),
                g0=9.80665 * (1.0 + 1e-3 * r.uniform(-1, 1, n)),
                engine=rocket_engine_table(), mass_tab=rocket_mass_table(), ix_tab=None, iyz_tab=None,
                t_end=30.0, dt=0.01, n_steps=3000)


def grid2d_state(rows: int, cols: int, spacing: float = 0.5, bump: bool = True) -> np.ndarray:
    """Config 5 initial state, AoS [rows*cols][x,y,vx,vy]: rest grid (connectivity_matrix_2d.hpp:171-180)
    + centre mass displaced (+0.1, 0) + Gaussian y-bump 0.5*exp(-r^2/0.5) about the grid centre
    (tests/grid_2d_test.cpp:267-276)."""
    c, r = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    s = np.zeros((rows, cols, 4))
    s[..., 0] = c * spacing
    s[..., 1] = r * spacing
    if bump:
        cx, cy = 0.5 * (cols - 1) * spacing, 0.5 * (rows - 1) * spacing
        dx, dy = c * spacing - cx, r * spacing - cy
        s[..., 1] += 0.5 * np.exp(-(dx * dx + dy * dy) / 0.5)
        s[rows // 2, cols // 2, 0] += 0.1
    return np.ascontiguousarray(s.reshape(rows * cols, 4))


GRID2D_PARAMS = dict(mass=1.0, spacing=0.5, k=50.0, c=0.5, dt=1e-3)
