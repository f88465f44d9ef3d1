"""Launches tests/mgpu_pcg2d.py on 2 GPUs when the box has them (NCCL halo exchange + all-reduce)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_pcg2d.py")],
                       capture_output=True, text=True, timeout=600)
    assert "MGPU2D_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
