"""Launches tests/mgpu_pcg2d.py on 2 GPUs when the box has them (NCCL halo exchange + all-reduce)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,port", [("nccl", "29533"), ("p2p", "29537")])
def test_slab_partition_two_gpus(mode, port):
    """nccl: halo exchange and all-reduce over NCCL (CUDA-graph replayed); p2p: the persistent kernel that stores
    halo columns straight into the neighbours' memory and all-reduces through peer-mapped slots"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tests", "mgpu_pcg2d.py"), mode],
                       capture_output=True, text=True, timeout=600)
    assert "MGPU2D_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
