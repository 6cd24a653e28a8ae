"""Worker for tests/test_multigpu.py: run under torchrun with one rank per GPU.
Checks (1) dispersion statistics all-reduced over NCCL == numpy on the concatenated data,
(2) slab-decomposed grid2d (fused step with one 4-row halo exchange, and per-stage 1-row exchanges) == the
    single-GPU run bit for bit,
(3) ensemble sharding: concatenated shard results == one-GPU results bit for bit,
(4) slab-decomposed unified 6-state grid, all three integrators == the single-GPU run bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sopot_b200 import workloads as W  # noqa: E402
from sopot_b200.capi import FP_CONTRACT, FP_STRICT, Engine  # noqa: E402
from sopot_b200.distributed import dist_env, make_engine, partition, slab  # noqa: E402

rank, world, local = dist_env()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = make_engine()

# (1) dispersion over ranks
rng = np.random.default_rng(123)
full = rng.normal(5.0, 2.0, (6, 100003))
full[2, 77] = np.inf
b, e = partition(full.shape[1], rank, world)
d = eng.dispersion_stats(np.ascontiguousarray(full[:, b:e]))
for r in range(6):
    row = full[r][np.isfinite(full[r])]
    assert d[r]["count"] == row.size and d[r]["nonfinite"] == full.shape[1] - row.size, (rank, r, d[r])
    assert d[r]["min"] == row.min() and d[r]["max"] == row.max()
    assert abs(d[r]["mean"] - row.mean()) < 1e-12 * abs(row.mean()) and abs(d[r]["stddev"] - row.std()) < 1e-9

# narrow spread on a large mean: the centred second moment (second pass about the all-reduced mean) keeps the deviation
big = np.stack([1.0e5 + rng.normal(0.0, 1.0, 300007), 2.0e7 + rng.normal(0.0, 1e-2, 300007)])
b, e = partition(big.shape[1], rank, world)
for r, dd in enumerate(eng.dispersion_stats(np.ascontiguousarray(big[:, b:e]))):
    assert abs(dd["mean"] - big[r].mean()) <= 1e-13 * big[r].mean() and abs(dd["stddev"] - big[r].std()) <= 1e-9 * big[r].std(), (r, dd)

# the slab contract: anything but the even contiguous partition is rejected on every rank BEFORE the first exchange
from sopot_b200.capi import SopotB200Error  # noqa: E402
if world > 1:
    rows = 40 * world + 3
    rb, rl = slab(rows, rank, world)
    bad_rl = rl - 1 if rank == 0 else rl                       # rank 0 short by one row, the others shifted up by one
    bad_rb = rb if rank == 0 else rb - 1
    if rank == world - 1:
        bad_rl += 1
    pg = W.GRID2D_PARAMS
    try:
        eng.grid2d_rk4(eng.grid2d_params(rows, 32, 0, pg["mass"], pg["spacing"], pg["k"], pg["c"], row_begin=bad_rb, rows_local=bad_rl),
                       pg["dt"], 1, np.zeros((bad_rl * 32, 4)))
        raise AssertionError("uneven slab accepted")
    except SopotB200Error as ex:
        assert "even contiguous partition" in str(ex), str(ex)
    try:
        eng.ugrid_rk4(eng.ugrid_params(rows, 32, 1.0, 0.05, 0.5, 50.0, 0.5, 0.5, row_begin=bad_rb, rows_local=bad_rl), 1e-3, 1,
                      np.zeros((bad_rl * 32, 6)))
        raise AssertionError("uneven slab accepted")
    except SopotB200Error as ex:
        assert "even contiguous partition" in str(ex), str(ex)

# (2) grid2d slabs vs single domain (computed redundantly on every rank with a private single-rank engine)
p = W.GRID2D_PARAMS
# (slabs of more than 48 rows take the schedule that overlaps the halo exchange with the interior tiles; 49-row slabs end in a
#  one-row tile, the case where the last TWO tile rows touch the ghost rows)
for rows, cols, topo, steps in ((64, 96, 0, 12), (37, 130, 1, 6), (world, 64, 2, 3), (4 * world + 1, 70, 0, 5), (1024, 1024, 0, 4),
                                (49 * world + 1, 200, 0, 7), (64 * world + 5, 96, 1, 3), (51 * world, 40, 2, 2)):
    s0 = W.grid2d_state(rows, cols, p["spacing"]).reshape(rows, cols, 4)
    solo = Engine(local)
    ref = solo.grid2d_rk4(solo.grid2d_params(rows, cols, topo, p["mass"], p["spacing"], p["k"], p["c"]), p["dt"], steps,
                          s0.reshape(-1, 4).copy(), fp_mode=FP_STRICT).reshape(rows, cols, 4)
    solo.close()
    rb, rl = slab(rows, rank, world)
    gp = eng.grid2d_params(rows, cols, topo, p["mass"], p["spacing"], p["k"], p["c"], row_begin=rb, rows_local=rl)
    # fused step (one 4-row exchange per step) with the kernel chosen by size, with the marching kernel forced (it carries the
    # exchange over peer memory) and with the tile kernel forced (NCCL exchange between launches); per-stage schedule (four 1-row)
    for kw in (dict(), dict(kernel="march"), dict(kernel="tile"), dict(per_stage=True)):
        mine = eng.grid2d_rk4(gp, p["dt"], steps, s0[rb:rb + rl].reshape(-1, 4).copy(), fp_mode=FP_STRICT, **kw)
        assert np.array_equal(mine.reshape(rl, cols, 4), ref[rb:rb + rl]), (rank, rows, cols, topo, kw)
    # contract mode: the slabs agree with the single-rank run to rounding (the multi-rank step is a separately compiled
    # instantiation -- exchange over peer memory inside the kernel, boundary tiles -- whose FMA contraction may differ; only
    # strict arithmetic promises equal bits across decompositions)
    solo = Engine(local)
    ref_c = solo.grid2d_rk4(solo.grid2d_params(rows, cols, topo, p["mass"], p["spacing"], p["k"], p["c"]), p["dt"], steps,
                            s0.reshape(-1, 4).copy(), fp_mode=FP_CONTRACT).reshape(rows, cols, 4)
    solo.close()
    for kernel in (None, "march", "tile"):
        mine_c = eng.grid2d_rk4(gp, p["dt"], steps, s0[rb:rb + rl].reshape(-1, 4).copy(), fp_mode=FP_CONTRACT, kernel=kernel)
        assert np.abs(mine_c.reshape(rl, cols, 4) - ref_c[rb:rb + rl]).max() <= 1e-12 * np.abs(ref_c).max(), (rank, rows, cols, topo, kernel)
    dev = eng.to_device(s0[rb:rb + rl].reshape(-1, 4).copy())
    eng.grid2d_rk4(gp, p["dt"], steps, dev, fp_mode=FP_STRICT)
    assert np.array_equal(dev.download().reshape(rl, cols, 4), ref[rb:rb + rl])
    init = eng.grid2d_initial_state(gp).reshape(rl, cols, 4)
    assert np.array_equal(init, W.grid2d_state(rows, cols, p["spacing"], bump=False).reshape(rows, cols, 4)[rb:rb + rl])

# (3) ensemble sharding is exact: shard == slice of the whole
n = 50001
w = W.dpend_ensemble(n)
keys = ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")
b, e = partition(n, rank, world)
part, _, _ = eng.dpend_rk4(*[w[k][b:e] for k in keys], e - b, 0.0, 1e-3, 300, fp_mode=FP_CONTRACT)
solo = Engine(local)
whole, _, _ = solo.dpend_rk4(*[w[k] for k in keys], n, 0.0, 1e-3, 300, fp_mode=FP_CONTRACT)
solo.close()
assert np.array_equal(part, whole[:, b:e])

# (4) unified 6-state grid: slabs + 1-row halo per stencil input vs single domain, RK4 / symplectic Euler / velocity Verlet
PI = 3.14159265358979323846
uprm = (1.0, 0.05, 0.5, 50.0, 0.5, 0.5, 0.0, -1.0)
for rows, cols, steps, tri in ((world, 40, 3, False), (4 * world + 1, 70, 5, True), (257, 129, 4, False), (3 * world + 2, 33, 3, True)):
    solo = Engine(local)
    ug = solo.ugrid_params(rows, cols, *uprm, triangle=tri)
    s0 = solo.ugrid_initial_state(ug).reshape(rows, cols, 6)
    rng = np.random.default_rng(rows)
    s0 = s0 + rng.normal(0, 0.02, s0.shape); s0[..., 4] = rng.normal(0, 0.2, (rows, cols))
    rb, rl = slab(rows, rank, world)
    gp = eng.ugrid_params(rows, cols, *uprm, row_begin=rb, rows_local=rl, triangle=tri)
    for mode in (FP_STRICT, FP_CONTRACT):
        for integ in (0, 1, 2):
            ref = solo.ugrid_rk4(ug, 1e-3, steps, s0.reshape(-1, 6).copy(), fp_mode=mode, integrator=integ).reshape(rows, cols, 6)
            mine = eng.ugrid_rk4(gp, 1e-3, steps, s0[rb:rb + rl].reshape(-1, 6).copy(), fp_mode=mode, integrator=integ)
            assert np.array_equal(mine.reshape(rl, cols, 6), ref[rb:rb + rl]), (rank, rows, cols, mode, integ)
    init = eng.ugrid_initial_state(gp).reshape(rl, cols, 6)
    assert np.array_equal(init, solo.ugrid_initial_state(ug).reshape(rows, cols, 6)[rb:rb + rl])
    solo.close()

dist.barrier()
eng.close()
dist.destroy_process_group()
print(f"rank {rank}/{world} multigpu ok", flush=True)
