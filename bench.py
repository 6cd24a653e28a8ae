#!/usr/bin/env python
"""bench.py -- RK4 instance-steps/s (FP64) of the B200 ensemble engine, one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], the config the metric is quoted on): double-pendulum ensemble,
4 Mi instances per GPU, RK4 dt = 1e-3 over 10 s = 10 000 steps per instance.  One bench "step" is one
pass of the hot path over that batch = ONE launch of rk4_ensemble_kernel<DpendSys> (4.19e10
instance-steps).  Ensembles shard trivially: every rank integrates its own 4 Mi instances, there is no
data-path collective ("scaling": "weak"); torch.distributed (NCCL) is used only for the barrier and the
max-over-ranks of the device time.

  value     instance-steps/s with inputs resident in HBM (device pointers), CUDA-event timed.
  e2e       the same pass through the C-ABI with HOST buffers (pinned): H2D of the 9 parameter arrays and
            D2H of the final state + status inside the timed region.
  roofline  FP64: algorithmic flops (204 per instance-step, SURVEY.md section 8d: every + - * / and every
            transcendental = 1 flop) / kernel time, against the DFMA issue peak measured in this run by
            sopot_b200_measure_fp64_peak (MEASURED_PEAKS.json carries no FP64 figure).
  cpu_baseline  the reference's own CPU RK4 (oracle/_ref when it was built, else the oracle port) on
            all host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 1 << 22
N_STEPS = 10000
DT = 1e-3
FLOP_PER_INSTANCE_STEP = 204.0
# ncu counters of the headline kernel (DRAM bytes per launch, FP64 instructions per instance-step, three-register DFMAs)
# are NOT literals here: tools/ncu_counters.py writes them to profiles/dpend_bench_counters.json together with the SHA-256
# of the kernel's SASS, and load_counters() below hands them out only while the loaded library still holds that code.
COUNTERS_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "dpend_bench_counters.json")
ROCKET_OPCOUNT_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "rocket_opcount.json")
ROCKET_FLOP_AS_WRITTEN = 1870.0         # the reference's RHS as written (re-evaluated sub-trees)
METRIC = "RK4 instance-steps/s (FP64)"
UNIT = "instance-steps/s"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples drawing more than half the maximum power seen
        load = [s for s, p in zip(sm, power) if power and p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(seconds_target: float, force_kind: str | None = None):
    """Time the reference's CPU RK4 on the double-pendulum workload; returns (inst-steps/s, cores, kind, sample)."""
    from oracle.oracle_py import Checker, build, have_reference
    from sopot_b200 import workloads as W
    build()
    kind = force_kind or ("reference" if have_reference() else "port")
    chk = Checker(kind)
    cores = chk.hardware_threads()
    w = W.dpend_ensemble(64 * cores, seed=2)
    args = [w[k] for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")]
    t = time.perf_counter()
    chk.dpend_rk4(*args, 0.0, 0.5, DT, nthreads=cores)                    # probe: 500 steps
    probe = time.perf_counter() - t
    rate = 64 * cores * 500 / max(probe, 1e-6)
    n = int(max(cores * 8, min(N_PER_GPU, rate * seconds_target / N_STEPS)))
    n = (n // cores) * cores or cores
    w = W.dpend_ensemble(N_PER_GPU, seed=2)                                # the SAME inputs as the GPU run, first n
    args = [w[k][:n].copy() for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")]
    t = time.perf_counter()
    chk.dpend_rk4(*args, 0.0, 10.0, DT, nthreads=cores)
    el = time.perf_counter() - t
    sample = f"first {n} of the {N_PER_GPU} instances x {N_STEPS} steps, {cores} threads, {el:.1f} s wall"
    return n * N_STEPS / el, cores, kind, sample, el


def workload_config(n, world):
    """`config` of the JSON line, the same object in both arms (ours and --impl reference)."""
    return {"workload": "double pendulum ensemble (BASELINE configs[1]): m=L=1, g=9.81, theta~U(-3,3), "
                        "RK4 dt=1e-3 over 10 s = 10000 steps/instance, one kernel launch per step",
            "instances_per_gpu": n, "n_steps": N_STEPS, "dt": DT, "fp_mode": "contract",
            "l2": "inputs larger than L2 (9 arrays x %d MiB per step)" % (n * 8 >> 20),
            "parallelism": "ensemble sharded over %d GPU(s), no data-path collective" % world}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    rates, ms = [], []
    for i in range(args.warmup + args.steps):
        r, cores, kind, sample, el = cpu_reference_rate(4.0 if i < args.warmup else 12.0)
        if i >= args.warmup:
            rates.append(r); ms.append(el * 1e3)
    v = statistics.mean(rates)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": statistics.mean(ms), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # the product arm's config, verbatim (the driver compares the two); what this arm actually integrates per step -- a
            # bounded sample of that batch on the host cores -- is stated in cpu_baseline.sample
            "config": workload_config(args.instances, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def load_counters(n_instances):
    """(counters dict or None, reason).  The profile must belong to the code that is running: same SASS hash of
    rk4_ensemble_kernel<DpendSys,false> in the loaded library, and (for the DRAM traffic) the same batch size."""
    if not os.path.exists(COUNTERS_FILE):
        return None, "profiles/dpend_bench_counters.json not present"
    try:
        c = json.load(open(COUNTERS_FILE))
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from kernel_hash import kernel_hash
        from sopot_b200.capi import LIB_PATH
        live = kernel_hash(LIB_PATH, c["kernel_mangled"])["sass_sha256"]
    except Exception as e:
        return None, "cannot hash the loaded kernel: " + repr(e)
    if live != c["sass_sha256"]:
        return None, "counters are stale: captured on SASS %s..., the loaded library has %s..." % (c["sass_sha256"][:12], live[:12])
    c["traffic_valid"] = int(c["units_per_launch"]) == int(n_instances) * N_STEPS
    return c, "profiles/dpend_bench_counters.json (SASS hash matches the loaded library)"


def parity_check(eng, w, keys):
    """The timed configuration tied to the oracle: the first 256 instances of THIS bench's inputs, 1000 steps (t = 1 s, the
    hard gate of SURVEY.md section 8c for the chaotic config), contract mode on the GPU against the CPU checker.  Part of
    the checker leg (after the timed region); a miss fails the run."""
    from oracle.oracle_py import Checker, build, have_reference
    from sopot_b200.capi import FP_CONTRACT
    build()
    kind = "reference" if have_reference() else "port"
    chk = Checker(kind)
    m = 256
    args = [np.ascontiguousarray(w[k][:m]) for k in keys]
    got, _, status = eng.dpend_rk4(*args, m, 0.0, DT, 1000, fp_mode=FP_CONTRACT)
    ref, _, _ = chk.dpend_rk4(*args, 0.0, 1.0, DT, nthreads=min(8, chk.hardware_threads()))
    err = float((np.abs(got - ref).max(axis=0) / np.maximum(np.abs(ref).max(axis=0), 1e-300)).max())
    return {"instances": m, "steps": 1000, "checker": kind, "max_rel_err": err, "tolerance": 1e-9,
            "ok": bool(err <= 1e-9 and not np.any(status))}


def multi_gpu_configs(eng, rank, world, local, hbm_peak):
    """The communicating paths at N > 1 (every rank calls this; rank 0 gets the rows):
      * grid2d 8192^2 (BASELINE config 5) STRONG scaling: row slabs, halo exchange over NCCL inside the library; fused step
        (one 4-row exchange per step) and per-stage schedule (four 1-row exchanges), both arithmetic modes.  Beside the
        multi-rank time: the same slab on a single-rank context (same tiles, no exchange) = what the kernels alone cost,
        and the whole grid on rank 0 alone = this run's own N = 1, so efficiency needs no other run.
      * the unified 6-state grid 4096^2 the same way (RK4, contract).
      * rocket Monte Carlo (config 4), 1 Mi trajectories split over the ranks + the dispersion all-reduce.
      * a 512 x 512 strict run on the slabs against the single-rank bits (SHA-256 of the assembled state on rank 0).
    Device times are CUDA events on the library's stream, max over ranks."""
    import hashlib
    import torch
    import torch.distributed as dist
    from sopot_b200 import workloads as W
    from sopot_b200.capi import FP_CONTRACT, FP_STRICT, MEM_DEVICE, Engine
    from sopot_b200.distributed import partition, slab
    out = {}

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(x):
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    def timed(e, fn, steps):
        fn(3)                                              # warm-up: NCCL channels, function attributes, scratch
        e.sync(); dist.barrier(); torch.cuda.synchronize()
        e.event_record(2)
        fn(steps)
        e.event_record(3)
        return e.elapsed_ms(2, 3) / steps

    p = W.GRID2D_PARAMS
    n, steps = 8192, 10 * world          # a step of a 1/N slab is short: more of them, so that the ranks' start skew after the barrier amortises
    rb, rl = slab(n, rank, world)
    solo = Engine(local)                                   # single-rank context on the same GPU
    gp = eng.grid2d_params(n, n, 0, p["mass"], p["spacing"], p["k"], p["c"], row_begin=rb, rows_local=rl)
    gp_slab = solo.grid2d_params(rl, n, 0, p["mass"], p["spacing"], p["k"], p["c"])
    gp_full = solo.grid2d_params(n, n, 0, p["mass"], p["spacing"], p["k"], p["c"])
    state = eng.grid2d_initial_state(gp, mem=MEM_DEVICE)
    rows = {}
    for mode, mname in ((FP_CONTRACT, "contract"), (FP_STRICT, "strict")):
        for per_stage in (False, True):
            ms = timed(eng, lambda k: eng.grid2d_rk4(gp, p["dt"], k, state, fp_mode=mode, sync=False, per_stage=per_stage), steps)
            per_rank = all_ranks(ms)
            ms_slab = timed(solo, lambda k: solo.grid2d_rk4(gp_slab, p["dt"], k, state, fp_mode=mode, sync=False, per_stage=per_stage), steps)
            per_rank_slab = all_ranks(ms_slab)
            rows[f"{mname}_{'per_stage' if per_stage else 'fused'}"] = {
                "ms_per_step": max(per_rank), "ms_per_step_per_rank": per_rank,
                "slab_kernels_only_ms_per_rank": per_rank_slab,
                "exchange_and_sync_ms": max(per_rank) - max(per_rank_slab),
                "mass_steps_per_s": n * n / (max(per_rank) * 1e-3)}
    state.free()
    dist.barrier()
    if rank == 0:                                          # this run's own N = 1: the whole grid on one GPU
        full = solo.grid2d_initial_state(gp_full, mem=MEM_DEVICE)
        for mode, mname in ((FP_CONTRACT, "contract"), (FP_STRICT, "strict")):
            for per_stage in (False, True):
                key = f"{mname}_{'per_stage' if per_stage else 'fused'}"
                solo.grid2d_rk4(gp_full, p["dt"], 2, full, fp_mode=mode, per_stage=per_stage)
                solo.event_record(2)
                solo.grid2d_rk4(gp_full, p["dt"], 10, full, fp_mode=mode, sync=False, per_stage=per_stage)
                solo.event_record(3)
                n1 = solo.elapsed_ms(2, 3) / 10
                r = rows[key]
                r["n1_ms_per_step_same_run"] = n1
                r["speedup_vs_n1"] = n1 / r["ms_per_step"]
                r["strong_scaling_efficiency"] = n1 / r["ms_per_step"] / world
                # the limiter, named: kernels on a 1/N slab against the exchange (+ stream ordering) on top of them
                kern = max(r["slab_kernels_only_ms_per_rank"])
                r["limiter"] = ("slab kernels (%.3f ms = %.2fx the ideal n1/N %.3f ms: pipeline fill of the row chunks / halo recomputation "
                                "of the tiles weighs more on a 1/N slab, plus the tail of a smaller launch)"
                                % (kern, kern / (n1 / world), n1 / world)) if kern - n1 / world >= r["exchange_and_sync_ms"] else \
                               ("ncclSend/Recv group + stream ordering (%.3f ms on top of %.3f ms of kernels)" % (r["exchange_and_sync_ms"], kern))
        full.free()
    dist.barrier()
    out["grid2d_8192x8192_strong_scaling"] = dict(rows, n_gpus=world, steps_timed=steps,
                                                  halo_bytes_per_step_fused=2 * 4 * n * 32, halo_bytes_per_step_per_stage=2 * 4 * n * 32)

    # unified 6-state grid, 4096^2, RK4 contract: one exchange of 4 boundary rows per step
    nu = 4096
    ub, ul = slab(nu, rank, world)
    ug = eng.ugrid_params(nu, nu, 1.0, 0.05, 0.5, 50.0, 0.5, 0.5, row_begin=ub, rows_local=ul)
    ustate = eng.ugrid_initial_state(ug, mem=MEM_DEVICE)
    ms = max_over_ranks(timed(eng, lambda k: eng.ugrid_rk4(ug, 1e-3, k, ustate, fp_mode=FP_CONTRACT, sync=False), 5))
    ustate.free()
    row = {"ms_per_step": ms, "mass_steps_per_s": nu * nu / (ms * 1e-3), "n_gpus": world}
    if rank == 0:
        ugf = solo.ugrid_params(nu, nu, 1.0, 0.05, 0.5, 50.0, 0.5, 0.5)
        uf = solo.ugrid_initial_state(ugf, mem=MEM_DEVICE)
        solo.ugrid_rk4(ugf, 1e-3, 2, uf, fp_mode=FP_CONTRACT)
        solo.event_record(2)
        solo.ugrid_rk4(ugf, 1e-3, 5, uf, fp_mode=FP_CONTRACT, sync=False)
        solo.event_record(3)
        n1 = solo.elapsed_ms(2, 3) / 5
        uf.free()
        row.update(n1_ms_per_step_same_run=n1, speedup_vs_n1=n1 / ms, strong_scaling_efficiency=n1 / ms / world)
    dist.barrier()
    out["unified_grid_4096x4096_contract_rk4_strong_scaling"] = row

    # rocket Monte Carlo: 1 Mi trajectories over the ranks, then the path's one collective (dispersion all-reduce)
    n_total = 1 << 20
    b, e = partition(n_total, rank, world)
    w = W.rocket_ensemble(n_total)
    d = {k: eng.to_device(w[k][b:e]) for k in ("elev", "az", "diam", "g0")}
    st, stats = eng.empty((14, e - b)), eng.empty((8, e - b))

    def fly(_k):
        eng.rocket_rk4(d["elev"], d["az"], d["diam"], d["g0"], w["engine"], w["mass_tab"], None, None, e - b, 0.01, 3000,
                       mem=MEM_DEVICE, want_status=False, state_out=st, stats_out=stats, sync=False)
    fly(1); eng.sync(); dist.barrier(); torch.cuda.synchronize()
    eng.event_record(2); fly(1); eng.event_record(3)
    ms_fly = max_over_ranks(eng.elapsed_ms(2, 3))
    eng.dispersion_stats(stats)                            # warm-up of the reduction + NCCL channels
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    disp = eng.dispersion_stats(stats)                     # device reduction + 4 small ncclAllReduce + D2H (synchronous call)
    ms_disp = max_over_ranks((time.perf_counter() - t0) * 1e3)
    out["rocket_1Mi_x3000_sharded_with_dispersion_allreduce"] = {
        "n_gpus": world, "trajectories_total": n_total, "ms_integrate": ms_fly, "ms_dispersion_allreduce_call": ms_disp,
        "instance_steps_per_s": n_total * 3000 / ((ms_fly + ms_disp) * 1e-3),
        "count_all_ranks": disp[0]["count"], "count_ok": bool(disp[0]["count"] == n_total),
        "apogee_mean_m": disp[0]["mean"], "apogee_std_m": disp[0]["stddev"],
        "max_speed_mean": disp[2]["mean"], "downrange_east_std_m": disp[6]["stddev"]}
    for a in (st, stats, *d.values()):
        a.free()

    # bit-exactness of the slab decomposition, in this very run: 512 x 512 strict, 8 steps, both schedules
    R = C = 512
    s0 = W.grid2d_state(R, C, p["spacing"]).reshape(R, C, 4)
    ref = solo.grid2d_rk4(solo.grid2d_params(R, C, 0, p["mass"], p["spacing"], p["k"], p["c"]), p["dt"], 8,
                          s0.reshape(-1, 4).copy(), fp_mode=FP_STRICT).reshape(R, C, 4)
    rb2, rl2 = slab(R, rank, world)
    gps = eng.grid2d_params(R, C, 0, p["mass"], p["spacing"], p["k"], p["c"], row_begin=rb2, rows_local=rl2)
    check = {"grid": [R, C], "steps": 8, "fp_mode": "strict"}
    # fused: the kernel chosen by size (tile kernel + NCCL exchange at 512^2); fused_march: the marching kernel forced, i.e. the
    # route the 8192^2 timings above take (halo rows pushed over peer memory from inside the kernel); per_stage: four exchanges
    routes = (("fused", dict()), ("fused_march", dict(kernel="march")), ("per_stage", dict(per_stage=True)))
    for name, kw in routes:
        mine = eng.grid2d_rk4(gps, p["dt"], 8, s0[rb2:rb2 + rl2].reshape(-1, 4).copy(), fp_mode=FP_STRICT, **kw)
        parts = [None] * world
        dist.all_gather_object(parts, mine.tobytes())
        if rank == 0:
            check[name + "_sha256"] = hashlib.sha256(b"".join(parts)).hexdigest()
            check[name + "_equals_single_rank"] = check[name + "_sha256"] == hashlib.sha256(ref.tobytes()).hexdigest()
    if rank == 0:
        check["single_rank_sha256"] = hashlib.sha256(ref.tobytes()).hexdigest()
        check["ok"] = bool(all(check[name + "_equals_single_rank"] for name, _ in routes))
    out["grid2d_512x512_multirank_bitexact"] = check
    solo.close()
    dist.barrier()
    return out


def other_configs(eng, hbm_peak, fp64_peak):
    """One timed pass (after one warm-up) of the other four BASELINE configs at full size, N = 1."""
    from sopot_b200 import workloads as W
    from sopot_b200.capi import FP_CONTRACT, FP_STRICT, MEM_DEVICE
    out = {}

    def timed(fn, reps=2):
        """ms per call: one warm-up, then `reps` calls enqueued back to back between two CUDA events on the engine's
        stream (host-side launch overhead of a single call would otherwise sit inside the timed interval of a
        kernel that runs for tens of microseconds)."""
        fn(); eng.sync()
        eng.event_record(2)
        for _ in range(reps):
            fn()
        eng.event_record(3)
        return eng.elapsed_ms(2, 3) / reps

    # config 1: oscillator 1 Mi x 1000 steps, final-state mode (FP64 bound, 44 flop / instance-step)
    n = 1 << 20
    w = W.ho_ensemble(n)
    d = {k: eng.to_device(w[k]) for k in ("m", "k", "c", "x0", "v0")}
    st = eng.empty((2, n))
    ms = timed(lambda: eng.ho_rk4(d["m"], d["k"], d["c"], d["x0"], d["v0"], n, 0.0, 0.01, 1000, mem=MEM_DEVICE,
                                  fp_mode=FP_CONTRACT, want_status=False, state_out=st, sync=False), 3)
    out["ho_1Mi_x1000"] = {"ms": ms, "instance_steps_per_s": n * 1000 / (ms * 1e-3),
                           "fp64_frac": 44.0 * n * 1000 / (ms * 1e-3) / 1e12 / fp64_peak}
    # config 1, FULL-TRAJECTORY mode (HBM-write bound, 16 B per instance-step): every step (16.8 GB
    # written per launch) and every 10th step; the fraction is written bytes / time against the HBM copy peak
    for every in (1, 10):
        traj = eng.empty((1000 // every, 2, n))
        ms = timed(lambda: eng.ho_rk4(d["m"], d["k"], d["c"], d["x0"], d["v0"], n, 0.0, 0.01, 1000, mem=MEM_DEVICE,
                                      fp_mode=FP_CONTRACT, store_every=every, want_status=False, state_out=st, traj_out=traj,
                                      sync=False), 2)
        wbytes = 16.0 * n * (1000 // every)
        out[f"ho_1Mi_x1000_trajectory_every_{every}"] = {
            "ms": ms, "instance_steps_per_s": n * 1000 / (ms * 1e-3), "bytes_written": wbytes,
            "write_GBs": wbytes / (ms * 1e-3) / 1e9, "hbm_frac": wbytes / (ms * 1e-3) / 1e9 / hbm_peak,
            "fp64_frac": 44.0 * n * 1000 / (ms * 1e-3) / 1e12 / fp64_peak}
        traj.free()
    # config 2 in the bit-exact arithmetic mode (reference association, IEEE division, libdevice sin/cos; the headline is the
    # contract mode): 1 Mi instances x 1000 steps
    w = W.dpend_ensemble(n)
    dd = [eng.to_device(w[k]) for k in ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")]
    st4 = eng.empty((4, n))
    ms = timed(lambda: eng.dpend_rk4(*dd, n, 0.0, 1e-3, 1000, mem=MEM_DEVICE, fp_mode=FP_STRICT, want_status=False,
                                     state_out=st4, sync=False), 2)
    out["dpend_strict_1Mi_x1000"] = {"ms": ms, "instance_steps_per_s": n * 1000 / (ms * 1e-3)}
    # config 2 with a stored trajectory (32 B per stored instance-step, every 100th step as SURVEY.md section 8d recommends):
    # 1 Mi instances x 10 000 steps -> 100 snapshots = 3.4 GB written beside the FP64-bound integration
    traj = eng.empty((100, 4, n))
    ms = timed(lambda: eng.dpend_rk4(*dd, n, 0.0, 1e-3, 10000, mem=MEM_DEVICE, fp_mode=FP_CONTRACT, store_every=100,
                                     want_status=False, state_out=st4, traj_out=traj, sync=False), 1)
    out["dpend_1Mi_x10000_trajectory_every_100"] = {
        "ms": ms, "instance_steps_per_s": n * 10000 / (ms * 1e-3), "bytes_written": 32.0 * n * 100,
        "write_GBs": 32.0 * n * 100 / (ms * 1e-3) / 1e9, "hbm_frac": 32.0 * n * 100 / (ms * 1e-3) / 1e9 / hbm_peak,
        "fp64_frac": FLOP_PER_INSTANCE_STEP * n * 10000 / (ms * 1e-3) / 1e12 / fp64_peak}
    traj.free()
    for a in (st4, *dd):
        a.free()
    # config 3: batched linearization, HBM bound, 200 B / point.  4 Mi points per launch (0.84 GB of traffic, far
    # larger than the 126 MB L2, so every output byte really goes to HBM inside the timed region); the BASELINE
    # size of 1 Mi points (210 MB, of which part stays L2-resident) is reported beside it.
    for P, name in ((1 << 22, "cartpole_linearize_4Mi"), (1 << 20, "cartpole_linearize_1Mi")):
        w = W.cartpole_points(P)
        dx, du = eng.to_device(w["x"]), eng.to_device(w["u"])
        A, B = eng.empty((16, P)), eng.empty((4, P))
        ms = timed(lambda: eng.cartpole_linearize(1.0, 0.1, 0.5, 9.81, dx, du, P, mem=MEM_DEVICE, A=A, B=B,
                                                  want_status=False, sync=False), 20)
        out[name] = {"ms": ms, "points_per_s": P / (ms * 1e-3), "GBs": 200.0 * P / (ms * 1e-3) / 1e9,
                     "hbm_frac": 200.0 * P / (ms * 1e-3) / 1e9 / hbm_peak}
        if P != 1 << 20:
            for a in (dx, du, A, B):
                a.free()
    for a in (dx, du, A, B, st, *d.values()):
        a.free()
    # section 8f-1: the LQR sweep that consumes the Jacobians (64 Ki plants at upright, jittered parameters), then the
    # closed loop under those gains (64 Ki plants x 2000 steps).  FP64-bound: ~590 Riccati iterations per point.
    P = 1 << 16
    rng = np.random.default_rng(3)
    plant = {k: eng.to_device(v) for k, v in dict(mc=rng.uniform(0.8, 1.5, P), m=rng.uniform(0.05, 0.2, P),
                                                   L=rng.uniform(0.4, 0.8, P), g=np.full(P, 9.81)).items()}
    dx, du = eng.to_device(np.zeros((4, P))), eng.to_device(np.zeros(P))
    A, B, _ = eng.cartpole_linearize(plant["mc"], plant["m"], plant["L"], plant["g"], dx, du, P, mem=MEM_DEVICE, want_status=False)
    Q = np.diag([10.0, 100.0, 1.0, 10.0])
    K = eng.empty((4, P))
    ms = timed(lambda: eng.lqr4_gain(A, B, P, Q, 0.01, mem=MEM_DEVICE, fp_mode=FP_CONTRACT, K=K, want_iters=False, sync=False), 2)
    _, _, iters, st = eng.lqr4_gain(A, B, P, Q, 0.01, mem=MEM_DEVICE, fp_mode=FP_CONTRACT, K=K)
    mean_it = float(iters.download().mean())
    out["lqr4_sweep_64Ki"] = {"ms": ms, "gains_per_s": P / (ms * 1e-3), "mean_riccati_iterations": mean_it,
                              "failed": int(np.count_nonzero(st.download()))}
    s0 = [eng.to_device(v) for v in (rng.uniform(-0.5, 0.5, P), rng.uniform(-0.1, 0.1, P), np.zeros(P), np.zeros(P))]
    cl_state = eng.empty((4, P))
    ms = timed(lambda: eng.cartpole_feedback_rk4(plant["mc"], plant["m"], plant["L"], plant["g"], *s0, K, P, 0.0, 0.002, 2000,
                                                 u_min=-50.0, u_max=50.0, saturate=True, mem=MEM_DEVICE, want_status=False,
                                                 want_stats=False, sync=False, state_out=cl_state), 2)
    cl_state.free()
    out["cartpole_closed_loop_64Ki_x2000"] = {"ms": ms, "instance_steps_per_s": P * 2000 / (ms * 1e-3)}
    for a in (dx, du, A, B, K, iters, st, *s0, *plant.values()):
        a.free()
    # config 4: rocket Monte Carlo 1 Mi x 3000 steps + dispersion reduction
    n = 1 << 20
    w = W.rocket_ensemble(n)
    d = {k: eng.to_device(w[k]) for k in ("elev", "az", "diam", "g0")}
    st, stats = eng.empty((14, n)), eng.empty((8, n))
    ms = timed(lambda: eng.rocket_rk4(d["elev"], d["az"], d["diam"], d["g0"], w["engine"], w["mass_tab"], None, None, n,
                                      0.01, 3000, mem=MEM_DEVICE, want_status=False, state_out=st, stats_out=stats,
                                      sync=False), 1)
    disp = eng.dispersion_stats(stats)
    rate = n * 3000 / (ms * 1e-3)
    own = json.load(open(ROCKET_OPCOUNT_FILE)) if os.path.exists(ROCKET_OPCOUNT_FILE) else None
    out["rocket_1Mi_x3000"] = {
        "ms": ms, "instance_steps_per_s": rate, "apogee_mean_m": disp[0]["mean"], "apogee_std_m": disp[0]["stddev"],
        # FP64 roofline two ways: the flop count of the engine's OWN de-duplicated derivative
        # (counting scalar over RocketSysT::rhs, flight-phase weighted, tools/rocket_opcount.py) and, beside it, the
        # reference's RHS as written, which re-evaluates shared sub-trees and therefore flatters the fraction
        "roofline": {"bound": "fp64", "peak": fp64_peak, "unit": "TFLOP/s",
                     "flop_per_unit_own_count": own["flop_per_instance_step"] if own else None,
                     "achieved_own_count": own["flop_per_instance_step"] * rate / 1e12 if own else None,
                     "frac_own_count": own["flop_per_instance_step"] * rate / 1e12 / fp64_peak if own else None,
                     "own_count_source": "profiles/rocket_opcount.json" if own else "profiles/rocket_opcount.json not present",
                     "flop_per_unit_as_written": ROCKET_FLOP_AS_WRITTEN,
                     "achieved_as_written": ROCKET_FLOP_AS_WRITTEN * rate / 1e12,
                     "frac_as_written": ROCKET_FLOP_AS_WRITTEN * rate / 1e12 / fp64_peak}}
    for a in (st, stats, *d.values()):
        a.free()
    # config 5: grid2d 8192^2, 10 RK4 steps.  Default schedule = fused step kernel (64 B of HBM traffic per mass-step:
    # FP64/latency bound); per_stage = the four-kernel schedule (544 B per mass-step: HBM bound, the fraction below).
    p = W.GRID2D_PARAMS
    n = 8192
    gp = eng.grid2d_params(n, n, 0, p["mass"], p["spacing"], p["k"], p["c"])
    state = eng.grid2d_initial_state(gp, mem=MEM_DEVICE)
    for mode, name in ((FP_STRICT, "strict"), (FP_CONTRACT, "contract")):
        for per_stage in (False, True):
            ms = timed(lambda: eng.grid2d_rk4(gp, p["dt"], 10, state, fp_mode=mode, sync=False, per_stage=per_stage), 2) / 10.0
            row = {"ms_per_step": ms, "mass_steps_per_s": n * n / (ms * 1e-3)}
            if per_stage:
                row["GBs"] = 544.0 * n * n / (ms * 1e-3) / 1e9
                row["hbm_frac"] = row["GBs"] / hbm_peak
            else:
                row["hbm_bytes_per_mass_step"] = 64
                row["speedup_vs_hbm_bound_at_544B"] = (544.0 * n * n / (hbm_peak * 1e9)) / (ms * 1e-3)
            out[f"grid2d_8192x8192_{name}_{'per_stage' if per_stage else 'fused'}"] = row
    state.free()
    # section 8f-2: the unified 6-state grid (the system of the reference's Grid2DSimulator), 4096^2 masses, per-stage kernels,
    # 816 B of HBM traffic per mass-step (1.5x the 4-state grid)
    n = 4096
    pi = 3.14159265358979323846
    ug = eng.ugrid_params(n, n, 1.0, 0.05, 0.5, 50.0, 0.5, 0.5, 0.0, -1.0, (0.0, pi, pi / 2, -pi / 2))
    ustate = eng.ugrid_initial_state(ug, mem=MEM_DEVICE)
    for mode, name in ((FP_STRICT, "strict"), (FP_CONTRACT, "contract")):
        ms = timed(lambda: eng.ugrid_rk4(ug, 1e-3, 5, ustate, fp_mode=mode, sync=False), 2) / 5.0
        out[f"unified_grid_4096x4096_{name}"] = {"ms_per_step": ms, "mass_steps_per_s": n * n / (ms * 1e-3),
                                                 "GBs": 816.0 * n * n / (ms * 1e-3) / 1e9, "hbm_frac": 816.0 * n * n / (ms * 1e-3) / 1e9 / hbm_peak}
    ustate.free()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=N_PER_GPU, help="instances per GPU (default: the BASELINE config)")
    ap.add_argument("--no-extras", action="store_true", help="skip the other configs, the parity check and the CPU baseline")
    ap.add_argument("--extras-timeout", type=float, default=420.0, help="seconds after which the extras are abandoned")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = dist_env()
    import torch
    import torch.distributed as dist
    from sopot_b200 import workloads as W
    from sopot_b200.capi import FP_CONTRACT, MEM_DEVICE, MEM_HOST, Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [Engine.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    eng = Engine(local, nccl_id, rank, world)

    def barrier():
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.instances
    w = W.dpend_ensemble(n, seed=2 + rank)
    keys = ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")
    dev = {k: eng.to_device(w[k]) for k in keys}
    d_state = eng.empty((4, n))
    d_status = None

    def step_device():
        eng.dpend_rk4(*[dev[k] for k in keys], n, 0.0, DT, N_STEPS, mem=MEM_DEVICE, fp_mode=FP_CONTRACT,
                      want_status=False, state_out=d_state, sync=False)

    # ---- device-resident timing: W warm-up, then exactly K steps between barriers ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    eng.event_record(0)
    for _ in range(args.steps):
        step_device()
    eng.event_record(1)
    ms_total = eng.elapsed_ms(0, 1)
    barrier()
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the C ABI with pinned host buffers ----
    host = {k: eng.pinned((n,)) for k in keys}
    for k in keys:
        host[k][:] = w[k]
    h_state = eng.pinned((4, n))

    def step_host():
        eng.dpend_rk4(*[host[k] for k in keys], n, 0.0, DT, N_STEPS, mem=MEM_HOST, fp_mode=FP_CONTRACT,
                      want_status=False, state_out=h_state, sync=False)

    step_host()
    barrier()
    eng.event_record(4)
    for _ in range(args.steps):
        step_host()
    eng.event_record(5)
    ms_e2e = eng.elapsed_ms(4, 5)
    barrier()
    checksum = float(np.sum(h_state[:, :: max(1, n // 4096)]))       # the D2H result is really read
    assert np.isfinite(checksum)

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = (float(x) for x in t.cpu())

    total_units = float(n) * N_STEPS * world * args.steps
    value = total_units / (ms_total * 1e-3)
    e2e_value = total_units / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        fp64_peak = eng.measure_fp64_peak()
        hbm_peak_live = eng.measure_hbm_peak(1 << 30)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = hbm_peak_live, "measured live (sopot_b200_measure_hbm_peak)"
        if os.path.exists(peaks_path):
            hbm_peak, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json"
        kernel_ms = ms_total / max(launches, 1) if world == 1 else None
        counters, counters_src = load_counters(n)
        per_gpu_rate = float(n) * N_STEPS * args.steps / (ms_total * 1e-3)
        achieved = FLOP_PER_INSTANCE_STEP * per_gpu_rate / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 9 * n * 8, "d2h_bytes_per_step": 4 * n * 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu, same command); only valid
                         # for the BASELINE batch size the capture was taken at
                         "traffic": counters["dram_bytes_per_launch"] if counters and counters["traffic_valid"] else None,
                         "fp64_instr_per_unit": counters["fp64_instr_per_unit"] if counters else None,
                         "dfma3_per_unit": counters["dfma3_per_unit"] if counters else None,
                         "counters_source": counters_src,
                         "kernel": "rk4_ensemble_kernel<DpendSys,false>", "kernel_ms": kernel_ms,
                         "flop_per_unit": FLOP_PER_INSTANCE_STEP,
                         # share of the FP64 pipe's issue cycles the kernel's own instructions need at this rate
                         # (148 SMs x 4 sub-partitions, clock = median SM clock sampled during the timed region)
                         "fp64_issue_frac": (per_gpu_rate / 32.0 * counters["fp64_issue_cycles_per_warp_unit"] /
                                             (148 * 4 * (clocks or {}).get("sm_mhz", 1965.0) * 1e6))
                         if counters and clocks and clocks.get("sm_mhz") else None,
                         "peak_source": "DFMA issue rate measured in this run (sopot_b200_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src, "hbm_copy_live_gbs": hbm_peak_live},
        }
    # ---- beyond the headline: the other configs (N = 1) or the communicating paths (N > 1), the parity check of the timed
    # inputs against the oracle and the CPU baseline.  None of it may cost the headline line: a watchdog prints the line as
    # it stands and leaves if a collective of the extras hangs (one rank failing inside them would block the others).
    rc = 0
    done = threading.Event()

    def bail_out():
        if done.is_set():
            return
        if rank == 0:
            line.setdefault("other_configs", {"error": "extras timed out"})
            print(json.dumps(line), flush=True)
        os._exit(0)

    if not args.no_extras:
        watchdog = threading.Timer(args.extras_timeout, bail_out)
        watchdog.daemon = True
        watchdog.start()
        for a in list(dev.values()) + [d_state]:
            a.free()
        if world > 1:
            try:
                mg = multi_gpu_configs(eng, rank, world, local, hbm_peak if rank == 0 else None)
            except Exception as e:
                mg = {"error": repr(e)}
            if rank == 0:
                line["other_configs"] = mg
        elif rank == 0:
            try:
                line["other_configs"] = other_configs(eng, hbm_peak, fp64_peak)
            except Exception as e:
                line["other_configs"] = {"error": repr(e)}
        if rank == 0:
            try:
                line["parity_check"] = parity_check(eng, w, keys)
                if not line["parity_check"]["ok"]:
                    rc = 1
            except Exception as e:
                line["parity_check"] = {"error": repr(e), "ok": False}
                rc = 1
            if world == 1:
                try:
                    r, cores, kind, sample, _ = cpu_reference_rate(12.0)
                    line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
                except Exception as e:
                    line["cpu_baseline"] = {"error": repr(e)}
        done.set()
        watchdog.cancel()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return rc


if __name__ == "__main__":
    sys.exit(main())
