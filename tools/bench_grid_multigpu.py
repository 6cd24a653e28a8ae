"""Strong scaling of the grid2d cloth (BASELINE config 5) over the GPUs of one box: slab decomposition over rows, halo
exchange over NCCL inside libsopot_b200.so.  Launch with torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_grid_multigpu.py [n] [steps]
Prints one JSON line on rank 0: ms per RK4 step (max over ranks, CUDA events) for the fused and the per-stage schedule."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sopot_b200 import workloads as W  # noqa: E402
from sopot_b200.capi import FP_CONTRACT, FP_STRICT, MEM_DEVICE  # noqa: E402
from sopot_b200.distributed import dist_env, make_engine, slab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rank, world, local = dist_env()
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = make_engine()
p = W.GRID2D_PARAMS
rb, rl = slab(n, rank, world)
gp = eng.grid2d_params(n, n, 0, p["mass"], p["spacing"], p["k"], p["c"], row_begin=rb, rows_local=rl)
state = eng.grid2d_initial_state(gp, mem=MEM_DEVICE)
out = {"grid": n, "n_gpus": world, "steps": steps}
for mode, mname in ((FP_CONTRACT, "contract"), (FP_STRICT, "strict")):
    for per_stage in (False, True):
        eng.grid2d_rk4(gp, p["dt"], 3, state, fp_mode=mode, per_stage=per_stage)        # warm-up (NCCL channels, attributes)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.event_record(0)
        eng.grid2d_rk4(gp, p["dt"], steps, state, fp_mode=mode, per_stage=per_stage, sync=False)
        eng.event_record(1)
        ms = eng.elapsed_ms(0, 1) / steps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"{mname}_{'per_stage' if per_stage else 'fused'}_ms_per_step"] = float(t.item())
# the unified 6-state grid (Grid2DSimulator's system) on half the edge length: per-stage kernels, 1-row halo per stencil input
nu = n // 2
ug = eng.ugrid_params(nu, nu, 1.0, 0.05, 0.5, 50.0, 0.5, 0.5, row_begin=slab(nu, rank, world)[0], rows_local=slab(nu, rank, world)[1])
ustate = eng.ugrid_initial_state(ug, mem=MEM_DEVICE)
for integ, iname in ((0, "rk4"), (2, "velocity_verlet")):
    eng.ugrid_rk4(ug, 1e-3, 3, ustate, fp_mode=FP_CONTRACT, integrator=integ)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng.event_record(0)
    eng.ugrid_rk4(ug, 1e-3, steps, ustate, fp_mode=FP_CONTRACT, integrator=integ, sync=False)
    eng.event_record(1)
    t = torch.tensor([eng.elapsed_ms(0, 1) / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[f"unified_{nu}_contract_{iname}_ms_per_step"] = float(t.item())
if rank == 0:
    out["mass_steps_per_s_contract_fused"] = n * n / (out["contract_fused_ms_per_step"] * 1e-3)
    print(json.dumps(out), flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
