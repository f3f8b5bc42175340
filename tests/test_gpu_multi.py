"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): launches tests/multi_gpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_gpu_sharded_solve(peer):
    """peer=1: Krylov vectors exchanged by NVLink peer stores fused into the normalisation kernel; peer=0: NCCL allgather."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "2953" + peer, os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, BS_PEER_EXCHANGE=peer))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok: owned ") == 2
