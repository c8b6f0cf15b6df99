"""BASELINE.json configs[4] on real GPUs: the NCCL path of the data-parallel DQN trainer (gloo
world-2 covers the host logic on the CPU, tests/test_dist_cpu.py).  Needs >= 2 GPUs on the box
(`gpurun --gpus 2 -- python -m pytest tests/test_dp_nccl_gpu.py -m gpu`); skipped on one."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs 2 GPUs")
def test_nccl_gradient_is_the_mean_of_rank_gradients_and_weights_stay_in_sync(pb):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(port),
         os.path.join(ROOT, "tests", "nccl_dp_worker.py")],
        capture_output=True, text=True, timeout=420)
    assert res.returncode == 0, (res.stdout + res.stderr)[-4000:]
    assert f"DP_NCCL_OK world={world}" in res.stdout
