"""Multi-GPU parity on real hardware (`-m gpu`, skipped on boxes with a single GPU): launches
tests/harness/dist_check.py under torchrun on 2 (and, when present, 4 / 8) GPUs -- sharded engine with
the fused NVLink pull exchange against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_engine_on_gpus(world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "harness", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DIST_CHECK_OK world=%d" % world in out.stdout
    assert '"fused_exchange": true' in out.stdout
