"""The multi-GPU path through NCCL: launched under torchrun on 2 GPUs of the box (skips on a 1-GPU box).
The world-size-2 host logic is covered on the CPU by tests/test_distributed_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_nash_allgather_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           "29533", os.path.join(ROOT, "tests", "multigpu_nccl_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    log = out.stdout + out.stderr
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "multigpu_nccl_check.log"), "w") as f:
        f.write(log)
    assert out.returncode == 0, log[-4000:]
    assert "MULTIGPU NCCL CHECK OK" in log
