"""Data-parallel parity over NCCL on >= 2 GPUs (skipped on a single-GPU box; run with ``gpurun --gpus 2``).
The worker (tests/dp_nccl_worker.py) is launched exactly as the driver launches bench.py: one process per GPU."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("kind,mode", [("model_CNN_ad", "dp"), ("model_ad", "dp"), ("model_CNN_ad", "syncbn")])
def test_two_rank_nccl_gradients_and_graph_path(kind, mode):
    """mode dp: reduced gradients == mean of the ranks' gradients, graph path == eager path.  mode syncbn: with
    torch.nn.SyncBatchNorm.convert_sync_batchnorm(model) two ranks reproduce one device on the global batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, DP_MODEL=kind, DP_MODE=mode, PYTHONDONTWRITEBYTECODE="1")
    for attempt in range(4):                        # a "free" port can be taken again before torchrun binds it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_nccl_worker.py")]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
        if r.returncode == 0 or "EADDRINUSE" not in r.stderr:
            break
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0 and "DP_NCCL_OK" in r.stdout
