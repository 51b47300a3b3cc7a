"""Multi-GPU checks on the CUDA path (skipped below 2 GPUs): launches tests/multi_gpu_worker.py under torchrun with 2
ranks over NCCL - sharded render / training == unsharded with the device-reduced batch-global quantities, fused NVLink
peer all-reduce + Adam == NCCL all-reduce + Adam."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_sharded_paths_under_torchrun(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(out.stdout[-4000:], out.stderr[-2000:])
    assert out.returncode == 0
    assert out.stdout.strip().endswith("PASS")
