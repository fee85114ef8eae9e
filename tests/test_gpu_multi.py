"""GPU (needs >= 2 devices): launches tests/multi_gpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_sharded_solve_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(root, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=900)
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
