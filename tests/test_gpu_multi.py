"""Multi-GPU (one process per GPU, NCCL): sharded-minibatch CD equals the single-GPU update.
Skipped on boxes with a single GPU; the rank logic is also covered on CPU (gloo) in test_host_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_data_parallel_update_equals_single_gpu(prec):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(REPO, "tools", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PREC=prec))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "dp_check ok" in out.stdout
