"""Multi-GPU parity under pytest: spawns tests/dist_gpu_check.py with torch.distributed.run on 2 GPUs
(NCCL over NVLink) and requires its "ok" line.  Skipped on a box with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2])
def test_sharded_ops_on_gpus(world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(REPO, "tests", "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=570, cwd=REPO)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist_gpu_check ok on %d GPUs" % world in r.stdout, r.stdout[-2000:]
