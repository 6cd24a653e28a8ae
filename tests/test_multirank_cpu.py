"""CPU tests (gloo, world_size 2) of the host logic behind the N > 1 paths: ownership, NCCL-id broadcast
plumbing, the dispersion merge rule and the grid2d slab / halo schedule that grid2d.cu implements.
The slab stepping below is a numpy re-statement of the STAGE SCHEDULE only (which buffer's ghost rows are
refreshed when); its result must equal the single-domain oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from sopot_b200 import workloads as W
from sopot_b200.distributed import broadcast_bytes, halo_schedule, partition, slab


def test_partition_and_slab_cover_exactly():
    for n in (0, 1, 7, 8, 1 << 20, (1 << 22) + 3):
        for world in (1, 2, 3, 8):
            spans = [partition(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [slab(8192, r, 8) for r in (0, 7)] == [(0, 1024), (7168, 1024)]
    with pytest.raises(ValueError):
        slab(3, 0, 4)
    with pytest.raises(ValueError):
        partition(10, 2, 2)
    assert [s for _, s in halo_schedule()] == [0, 1, 2, 3, 4]


def _rhs_rows(S, ghost_top, ghost_bot, has_up, has_dn, mass, k, c, L):
    """Quad-grid derivative of a slab S[rows][cols][4] with ghost rows; same operation order as the
    reference aggregation: left(-), right(+), up(-), down(+)."""
    rows, cols, _ = S.shape
    ext = np.concatenate([ghost_top[None], S, ghost_bot[None]], axis=0)

    def spring(pi, pj):
        dx, dy = pj[..., 0] - pi[..., 0], pj[..., 1] - pi[..., 1]
        length = np.sqrt(dx * dx + dy * dy)
        safe = np.where(length < 1e-10, 1e-10, length)
        ux, uy = dx / safe, dy / safe
        rel = (pj[..., 2] - pi[..., 2]) * ux + (pj[..., 3] - pi[..., 3]) * uy
        F = k * (length - L) + c * rel
        return F * ux, F * uy

    Fx = np.zeros((rows, cols)); Fy = np.zeros((rows, cols))
    fx, fy = spring(S[:, :-1], S[:, 1:])
    Fx[:, 1:] -= fx; Fy[:, 1:] -= fy                      # left neighbour's spring: this mass is the second endpoint
    Fx[:, :-1] += fx; Fy[:, :-1] += fy
    fx, fy = spring(ext[:-1], ext[1:])                    # vertical springs incl. ghosts: index r = spring (r-1 -> r) in slab coords
    up_fx, up_fy = fx[:-1].copy(), fy[:-1].copy()         # spring from the row above, for slab rows 0..rows-1
    dn_fx, dn_fy = fx[1:].copy(), fy[1:].copy()           # spring to the row below
    if not has_up:
        up_fx[0] = 0; up_fy[0] = 0
    if not has_dn:
        dn_fx[-1] = 0; dn_fy[-1] = 0
    Fx = Fx - up_fx; Fy = Fy - up_fy
    Fx = Fx + dn_fx; Fy = Fy + dn_fy
    d = np.empty_like(S)
    d[..., 0], d[..., 1], d[..., 2], d[..., 3] = S[..., 2], S[..., 3], Fx / mass, Fy / mass
    return d


def _worker(rank, world, port, rows, cols, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    blob = broadcast_bytes(bytes(range(128)) if rank == 0 else None)
    assert blob == bytes(range(128))
    p = W.GRID2D_PARAMS
    full = W.grid2d_state(rows, cols, p["spacing"]).reshape(rows, cols, 4)
    rb, rl = slab(rows, rank, world)
    has_up, has_dn = rank > 0, rank < world - 1
    y = full[rb:rb + rl].copy()
    ghosts = {n: [np.zeros((cols, 4)), np.zeros((cols, 4))] for n in "yAB"}

    def exchange(name, buf):
        reqs, recv = [], {}
        for peer, send_row, slot in ((rank - 1, buf[0], 0), (rank + 1, buf[-1], 1)):
            if 0 <= peer < world:
                t = torch.from_numpy(np.ascontiguousarray(send_row))
                recv[slot] = torch.empty(cols, 4, dtype=torch.float64)
                reqs += [dist.isend(t, peer), dist.irecv(recv[slot], peer)]
        for r in reqs:
            r.wait()
        for slot, t in recv.items():
            ghosts[name][slot] = t.numpy().copy()

    f = lambda name, buf: _rhs_rows(buf, ghosts[name][0], ghosts[name][1], has_up, has_dn, p["mass"], p["k"], p["c"], p["spacing"] * 1.0)
    dt = p["dt"]
    exchange("y", y)
    for _ in range(steps):
        k1 = f("y", y) * dt; S = k1; A = y + 0.5 * k1; exchange("A", A)
        k2 = f("A", A) * dt; S = S + 2.0 * k2; B = y + 0.5 * k2; exchange("B", B)
        k3 = f("B", B) * dt; S = S + 2.0 * k3; A = y + k3; exchange("A", A)
        k4 = f("A", A) * dt; y = y + (S + k4) / 6.0; exchange("y", y)
    # dispersion merge rule: sums add, min/max reduce
    v = np.arange(1000.0)[rank::world]
    part = torch.tensor([v.size, v.sum(), (v * v).sum()], dtype=torch.float64)
    mn, mx = torch.tensor([v.min()]), torch.tensor([v.max()])
    dist.all_reduce(part); dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    assert part[0] == 1000 and part[1] == 499500.0 and mn[0] == 0 and mx[0] == 999
    np.save(os.path.join(out, f"slab{rank}.npy"), y)
    dist.destroy_process_group()


def test_slab_halo_schedule_matches_single_domain_oracle(port, tmp_path):
    rows, cols, steps, world = 9, 11, 5, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        free_port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, free_port, rows, cols, steps, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)], axis=0)
    p = W.GRID2D_PARAMS
    ref = port.grid2d_rk4(rows, cols, 0, p["mass"], p["spacing"], p["k"], p["c"], p["dt"], steps,
                          W.grid2d_state(rows, cols, p["spacing"]))
    assert np.array_equal(got.ravel(), ref)


# ---------------------------------------------------------------- unified 6-state grid: the same slab schedule (ugrid.cu)
UPRM = [1.0, 0.05, 0.5, 50.0, 0.5, 0.5, 0.0, -1.0, 0.0, 3.14159265358979323846, 3.14159265358979323846 / 2, -3.14159265358979323846 / 2]


def _ugrid_state(port, rows, cols):
    s0 = port.ugrid(rows, cols, UPRM, mode="init")
    rng = np.random.default_rng(rows * 31 + cols)
    st = s0 + rng.normal(0, 0.03, s0.shape)
    st[:, 4] = rng.normal(0, 0.3, rows * cols)
    return st


def _ugrid_worker(rank, world, port_no, rows, cols, steps, integrator, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    import torch
    sys.path.insert(0, ROOT)
    from oracle.oracle_py import Checker
    port = Checker("port")
    full = _ugrid_state(port, rows, cols).reshape(rows, cols, 6)
    rb, rl = slab(rows, rank, world)
    has_up, has_dn = rank > 0, rank < world - 1
    y = full[rb:rb + rl].copy()

    def exchange(buf):
        """boundary rows to the neighbours, theirs into (top, bottom) ghosts"""
        reqs, recv = [], {}
        for peer, send_row, slot in ((rank - 1, buf[0], 0), (rank + 1, buf[-1], 1)):
            if 0 <= peer < world:
                t = torch.from_numpy(np.ascontiguousarray(send_row))
                recv[slot] = torch.empty(cols, 6, dtype=torch.float64)
                reqs += [dist.isend(t, peer), dist.irecv(recv[slot], peer)]
        for r in reqs:
            r.wait()
        return recv.get(0), recv.get(1)

    def f(buf):
        """derivative of the slab rows: the oracle on the slab extended by its ghost rows (a ghost row only changes the
        ghost's own derivative, which is dropped -- exactly what the kernel's ghost loads do)"""
        top, bot = exchange(buf)
        parts = ([top.numpy()[None]] if has_up else []) + [buf] + ([bot.numpy()[None]] if has_dn else [])
        ext = np.concatenate(parts, axis=0)
        d = port.ugrid(ext.shape[0], cols, UPRM, ext.reshape(-1, 6)).reshape(ext.shape)
        return d[1 if has_up else 0: ext.shape[0] - (1 if has_dn else 0)]

    dt = 1e-3
    V, X = [2, 3, 5], [0, 1, 4]
    for _ in range(steps):
        if integrator == "rk4":                    # stage buffers A, B exchanged after the kernel that produced them
            k1 = f(y) * dt; S = k1; A = y + 0.5 * k1
            k2 = f(A) * dt; S = S + 2.0 * k2; B = y + 0.5 * k2
            k3 = f(B) * dt; S = S + 2.0 * k3; A = y + k3
            k4 = f(A) * dt; y = y + (S + k4) / 6.0
        elif integrator == "symplectic_euler":
            d = f(y); y = y.copy()
            y[..., V] += dt * d[..., V]; y[..., X] += dt * y[..., V]
        else:
            d = f(y); y = y.copy(); half_dt_sq = 0.5 * dt * dt
            y[..., X] += dt * y[..., V] + half_dt_sq * d[..., V]
            dn = f(y)
            y[..., V] += (0.5 * dt) * (d[..., V] + dn[..., V])
    np.save(os.path.join(out, f"uslab_{integrator}_{rank}.npy"), y)
    dist.destroy_process_group()


def _ugrid_deep_worker(rank, world, port_no, rows, cols, steps, integrator, out):
    """The default multi-rank schedule of ugrid.cu: H ghost rows at both ends of every buffer, ONE exchange of the H boundary
    rows per step, evaluation s of the step on the buffer shrunk by s rows (clamped to the global grid)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    import torch
    sys.path.insert(0, ROOT)
    from oracle.oracle_py import Checker
    port = Checker("port")
    full = _ugrid_state(port, rows, cols).reshape(rows, cols, 6)
    rb, rl = slab(rows, rank, world)
    H = {"rk4": 4, "velocity_verlet": 2, "symplectic_euler": 1}[integrator]
    ext_rows, g0 = rl + 2 * H, rb - H                      # global row of extended row 0 (negative on rank 0)
    Y = np.zeros((ext_rows, cols, 6)); Y[H:H + rl] = full[rb:rb + rl]

    def exchange(buf):
        reqs, recv = [], {}
        for peer, send, slot in ((rank - 1, buf[H:2 * H], 0), (rank + 1, buf[rl:rl + H], 1)):
            if 0 <= peer < world:
                recv[slot] = torch.empty(H, cols, 6, dtype=torch.float64)
                reqs += [dist.isend(torch.from_numpy(np.ascontiguousarray(send)), peer), dist.irecv(recv[slot], peer)]
        for r in reqs:
            r.wait()
        if 0 in recv:
            buf[:H] = recv[0].numpy()
        if 1 in recv:
            buf[H + rl:] = recv[1].numpy()

    def rng(s):                                             # rows on which evaluation s has valid inputs, inside the grid
        return max(s, -g0), min(ext_rows - s, rows - g0)

    def f(buf, lo, hi):                                     # derivative of rows [lo, hi): the oracle on them plus one context row
        a = lo - 1 if g0 + lo > 0 else lo
        b = hi + 1 if g0 + hi < rows else hi
        d = port.ugrid(b - a, cols, UPRM, buf[a:b].reshape(-1, 6)).reshape(b - a, cols, 6)
        return d[lo - a: lo - a + hi - lo]

    dt = 1e-3
    V, X = [2, 3, 5], [0, 1, 4]
    for _ in range(steps):
        exchange(Y)
        if integrator == "rk4":
            S, A, B = np.zeros_like(Y), np.zeros_like(Y), np.zeros_like(Y)
            lo, hi = rng(1); k = f(Y, lo, hi) * dt; S[lo:hi] = k; A[lo:hi] = Y[lo:hi] + 0.5 * k
            lo, hi = rng(2); k = f(A, lo, hi) * dt; S[lo:hi] = S[lo:hi] + 2.0 * k; B[lo:hi] = Y[lo:hi] + 0.5 * k
            lo, hi = rng(3); k = f(B, lo, hi) * dt; S[lo:hi] = S[lo:hi] + 2.0 * k; A[lo:hi] = Y[lo:hi] + k
            lo, hi = rng(4); k = f(A, lo, hi) * dt; Y[lo:hi] = Y[lo:hi] + (S[lo:hi] + k) / 6.0
        elif integrator == "symplectic_euler":
            lo, hi = rng(1); d = f(Y, lo, hi); N = Y.copy()
            N[lo:hi][..., V] += dt * d[..., V]; N[lo:hi][..., X] += dt * N[lo:hi][..., V]
            Y = N
        else:
            lo, hi = rng(1); d = f(Y, lo, hi); Bp = Y.copy(); An = np.zeros_like(Y); An[lo:hi] = d
            Bp[lo:hi][..., X] += dt * Y[lo:hi][..., V] + (0.5 * dt * dt) * d[..., V]
            lo, hi = rng(2); dn = f(Bp, lo, hi); N = Bp.copy()
            N[lo:hi][..., V] += (0.5 * dt) * (An[lo:hi][..., V] + dn[..., V])
            Y = N
    np.save(os.path.join(out, f"udeep_{integrator}_{rank}.npy"), Y[H:H + rl])
    dist.destroy_process_group()


@pytest.mark.parametrize("integrator", ["rk4", "symplectic_euler", "velocity_verlet"])
def test_unified_grid_one_exchange_per_step_matches_single_domain_oracle(port, tmp_path, integrator):
    rows, cols, steps, world = 9, 8, 3, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        free_port = s.getsockname()[1]
    mp.spawn(_ugrid_deep_worker, args=(world, free_port, rows, cols, steps, integrator, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"udeep_{integrator}_{r}.npy") for r in range(world)], axis=0)
    ref = port.ugrid(rows, cols, UPRM, _ugrid_state(port, rows, cols), 1e-3, steps, integrator)
    assert np.array_equal(got.reshape(-1, 6), ref)


@pytest.mark.parametrize("integrator", ["rk4", "symplectic_euler", "velocity_verlet"])
def test_unified_grid_slab_schedule_matches_single_domain_oracle(port, tmp_path, integrator):
    rows, cols, steps, world = 7, 9, 4, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        free_port = s.getsockname()[1]
    mp.spawn(_ugrid_worker, args=(world, free_port, rows, cols, steps, integrator, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"uslab_{integrator}_{r}.npy") for r in range(world)], axis=0)
    ref = port.ugrid(rows, cols, UPRM, _ugrid_state(port, rows, cols), 1e-3, steps, integrator)
    assert np.array_equal(got.reshape(-1, 6), ref)
