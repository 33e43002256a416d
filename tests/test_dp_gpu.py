"""Data-parallel parity (tools/dp_check.py): on a box with >= 2 GPUs over NCCL, one rank per GPU; on a 1-GPU box two
ranks share the GPU and reduce through gloo -- same host logic, same arena aliasing, same checks."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_data_parallel_gradients_and_weights():
    n = torch.cuda.device_count()
    env = dict(os.environ)
    if n < 2:
        n = 2
        env["DP_CHECK_ONE_GPU"] = "1"
    else:
        n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "result=OK" in out.stdout
