"""GPU test of the N > 1 paths (NCCL inside libsopot_b200.so); needs >= 2 GPUs on the box, else skipped."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run_worker(port, extra_env=None):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"{n} GPU(s) visible; the multi-rank NCCL paths need 2")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    env = dict(os.environ, **(extra_env or {}))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("multigpu ok") == world


def test_two_ranks_dispersion_halo_and_sharding():
    """Default route: halo rows pushed into the neighbours' ghost slots over peer memory from inside the step kernel."""
    _run_worker(29631)


def test_two_ranks_halo_over_nccl():
    """The same worker with peer memory switched off (SOPOT_B200_NO_P2P=1): halo rows travel by NCCL send/recv between launches,
    the route a box without peer access takes.  Same assertions, same bits in strict mode."""
    _run_worker(29633, {"SOPOT_B200_NO_P2P": "1"})
