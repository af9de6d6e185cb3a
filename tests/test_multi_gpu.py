"""GPU, >= 2 devices: the library's distributed layer (NCCL communicator, slab halo exchange, distributed DDH, distributed GMRES)
under torch.distributed.run — scripts/multi_check.py holds the assertions. Skipped on a single-GPU box (the driver's GPU test
run); the host logic of the same layer is covered without GPUs by tests/test_host_setup.py and tests/test_parallel_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_distributed_layer_two_gpus():
    n = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "scripts", "multi_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
