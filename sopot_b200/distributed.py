"""Host-side plumbing for N > 1: one process per GPU, torch.distributed for rendezvous only.

The data path needs exactly two exchanges, both inside libsopot_b200.so over NCCL:
  * ensembles: none while integrating; one all-reduce of <= 6*rows doubles for the dispersion statistics;
  * grid2d: one boundary row to each neighbour before every RK stage (slab decomposition over rows).
This module only decides who owns what and hands the library its NCCL bootstrap id.
"""
from __future__ import annotations

import os


def partition(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [begin, end) of n independent units for `rank`: sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def slab(rows: int, rank: int, world: int) -> tuple[int, int]:
    """Row slab (row_begin, rows_local) of a grid2d domain; every rank must own at least one row."""
    if rows < world:
        raise ValueError(f"{rows} rows cannot be split over {world} ranks")
    b, e = partition(rows, rank, world)
    return b, e - b


def halo_schedule():
    """Which stencil-input buffer has its ghost rows refreshed after which RK stage kernel (grid2d.cu):
    y before stage 1 (and after stage 4 of the previous step), A after stages 1 and 3, B after stage 2."""
    return [("y", 0), ("A", 1), ("B", 2), ("A", 3), ("y", 4)]


def dist_env() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload: bytes | None, src: int = 0) -> bytes:
    """Broadcast an opaque byte string (the 128-byte NCCL unique id) over the default process group."""
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def make_engine(device: int | None = None):
    """Engine for this rank; with WORLD_SIZE > 1 the ranks share one NCCL communicator owned by the library."""
    import torch.distributed as dist
    from .capi import Engine
    rank, world, local = dist_env()
    dev = local if device is None else device
    if world == 1:
        return Engine(dev)
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed must be initialised before make_engine() when WORLD_SIZE > 1")
    nccl_id = broadcast_bytes(Engine.nccl_unique_id() if rank == 0 else None)
    return Engine(dev, nccl_id, rank, world)
