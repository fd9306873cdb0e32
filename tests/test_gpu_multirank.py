"""Class-parallel head and row-sharded gallery across REAL NCCL ranks (one process per GPU): needs >= 2 GPUs, skipped
on a single-GPU box (tests/test_parallel_gloo.py covers the host logic there; bench.py's N > 1 lines carry the same
sharded-vs-unsharded check as `parity`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2])
def test_sharded_head_and_gallery_over_nccl(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "nccl_rank_main.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:]); sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0
    for rank in range(world):
        assert f"RANK {rank} OK" in r.stdout
